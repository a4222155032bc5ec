"""Experiment: hand-patches a generated boids2d model_kernels.cu with the bulk-staged tile loop (ABL_MODE 7)."""
import sys
p = sys.argv[1]
s = open(p).read()
s = s.replace('const unsigned _tile_cap, const bool _tile_ok, const Boid& in, Boid& out) {',
              'const unsigned _tile_cap, const bool _tile_ok, const abl_btile_flat2& _bt, const Boid& in, Boid& out) {')
loop7 = '''            if (ABL_MODE == 7 && _bt.ok) {
                extern __shared__ __align__(16) unsigned char _abl_smem[];
                const unsigned char *const _tc = _abl_smem + ABL_BTILE_HDR_BYTES;
                const abl_float2 *const _t0 = reinterpret_cast<const abl_float2 *>(_tc);
                const abl_float2 *const _t1 = reinterpret_cast<const abl_float2 *>(_tc + (size_t)_tile_cap * sizeof(abl_float2));
                for (unsigned _k = 0; _k < _bt.N; _k += 2) {
                    const bool _hB = _k + 1u < _bt.N;
                    const unsigned _kB = _hB ? _k + 1u : _k;
                    const unsigned _eA = _k + (_k < _bt.T1 ? _bt.O0 : (_k < _bt.T2 ? _bt.O1 : _bt.O2));
                    const unsigned _eB = _kB + (_kB < _bt.T1 ? _bt.O0 : (_kB < _bt.T2 ? _bt.O1 : _bt.O2));
                    const abl_float2 _pA = _t0[_eA];
                    const abl_float2 _pB = _t0[_eB];
                    const abl_real _d2A = abl_sqnorm2(float2_sub(_pA, in.pos));
                    const abl_real _d2B = abl_sqnorm2(float2_sub(_pB, in.pos));
                    if (!(_d2A > _near_limit)) {
                        Boid nx;
                        nx.pos = _pA;
                        nx.velocity = _t1[_eA];
                        {
                            global_center = float2_add(global_center, nx.pos);
                            global_velocity = float2_add(global_velocity, nx.velocity);
                            interaction_count += 1;
                            if ((_d2A <= _sql.v[0])) {
                                collision_center = float2_add(collision_center, nx.pos);
                                collision_count += 1;
                            }
                        }
                    }
                    if (_hB && !(_d2B > _near_limit)) {
                        Boid nx;
                        nx.pos = _pB;
                        nx.velocity = _t1[_eB];
                        {
                            global_center = float2_add(global_center, nx.pos);
                            global_velocity = float2_add(global_velocity, nx.velocity);
                            interaction_count += 1;
                            if ((_d2B <= _sql.v[0])) {
                                collision_center = float2_add(collision_center, nx.pos);
                                collision_count += 1;
                            }
                        }
                    }
                }
            } else
'''
assert '            if (ABL_MODE == 2 && _tile_ok) {\n' in s
s = s.replace('            if (ABL_MODE == 2 && _tile_ok) {\n', loop7 + '            if (ABL_MODE == 2 && _tile_ok) {\n', 1)
s = s.replace('if (ABL_MODE == 3) _var0.rows2(_a, in.pos, true, _near_cull); else _var0.init2',
              'if (ABL_MODE == 3 || ABL_MODE == 7) _var0.rows2(_a, in.pos, true, _near_cull); else _var0.init2')
s = s.replace('                    if (ABL_MODE == 3) {\n                        const unsigned _var0T1',
              '                    if (ABL_MODE == 3 || ABL_MODE == 7) {\n                        const unsigned _var0T1')
s = s.replace('    if (ABL_MODE != 2 && !_active) return;', '    if (ABL_MODE != 2 && ABL_MODE != 7 && !_active) return;')
wr = '''    abl_btile_flat2 _bt;
    _bt.ok = false; _bt.T1 = _bt.T2 = _bt.N = _bt.O0 = _bt.O1 = _bt.O2 = 0;
    if (ABL_MODE == 7) {
        extern __shared__ __align__(16) unsigned char _abl_smem[];
        const abl_btile_cols _TC = { 2, {0, 1}, {(int)sizeof(abl_float2), (int)sizeof(abl_float2)}, {0u, (unsigned)sizeof(abl_float2)}, (unsigned)(2 * sizeof(abl_float2)) };
        if (threadIdx.x < 32) {
            unsigned _bf = 0, _bl = 0;
            const bool _any = abl_block_span(_a, _bf, _bl);
            const abl_real *const _pc = static_cast<const abl_real *>(_a.self.in[0]);
            abl_real _pf[2] = {0, 0}, _pl[2] = {0, 0};
            if (_any) { _pf[0] = __ldg(_pc + 2 * (size_t)_bf); _pf[1] = __ldg(_pc + 2 * (size_t)_bf + 1); _pl[0] = __ldg(_pc + 2 * (size_t)_bl); _pl[1] = __ldg(_pc + 2 * (size_t)_bl + 1); }
            abl_btile_plan<2>(_a, _pf, _pl, _any, _tile_cap, _abl_smem, _TC);
        }
        if (!_active) return;
        abl_near_iter<2> _rows;
        _rows.rows2(_a, in.pos, _active, _near_cull);
        _tile_ok = abl_btile_wait(_abl_smem);
        _bt = abl_btile_thread2(_rows, _abl_smem, _tile_ok);
    }
'''
assert '    Boid out = in;\n    abl_ctx _ctx;' in s
s = s.replace('    Boid out = in;\n    abl_ctx _ctx;', wr + '    Boid out = in;\n    abl_ctx _ctx;', 1)
s = s.replace('update_boid<ABL_MODE>(_ctx, _a, _i, _near_limit, _near_cull, _sql, _tile_cap, _tile_ok, in, out);',
              'update_boid<ABL_MODE>(_ctx, _a, _i, _near_limit, _near_cull, _sql, _tile_cap, _tile_ok, _bt, in, out);')
la = '''    if (getenv("ABL_EXP_MODE7")) {
        const int bs7 = getenv("ABL_EXP_BS") ? atoi(getenv("ABL_EXP_BS")) : 128;
        const unsigned entry7 = 2 * sizeof(abl_float2);
        unsigned cap7 = abl_tile_capacity(a, 3, bs7, entry7, 200u * 1024u);
        if (getenv("ABL_EXP_CAP")) cap7 = atoi(getenv("ABL_EXP_CAP"));
        if (cap7) {
            const size_t smem7 = ABL_BTILE_HDR_BYTES + (size_t)cap7 * entry7;
            cudaFuncSetAttribute(abl_kernel_update_boid<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7);
            abl_last_mode_update_boid = 7;
            return (int)abl_launch_kernel(a, abl_kernel_update_boid<7>, abl_grid_blocks(a, bs7), bs7, smem7, *a, limit, cull, sql, cap7);
        }
    }
'''
assert '    if (tile_cap) {\n        const unsigned tile_grid' in s
s = s.replace('    if (tile_cap) {\n        const unsigned tile_grid', la + '    if (tile_cap) {\n        const unsigned tile_grid', 1)
s = s.replace('    cudaGridDependencySynchronize();   // programmatic dependent launch: wait for the preceding kernel\n    bool _boundary;', '    if (ABL_MODE == 7) { extern __shared__ __align__(16) unsigned char _abl_smem0[]; abl_btile_begin(_abl_smem0); }\n    cudaGridDependencySynchronize();   // programmatic dependent launch: wait for the preceding kernel\n    bool _boundary;')
open(p, 'w').write(s)
