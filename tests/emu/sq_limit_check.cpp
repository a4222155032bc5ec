// TEST INFRASTRUCTURE: checks abl_near_sq_limit / abl_sq_cmp_limit (asset/cuda/abl_device.cuh, host code
// of the generated launchers) against the predicate they replace, around the bound and at random.
//   g++ -O1 -ffp-contract=off -std=c++17 [-DABL_USE_FLOAT] -I tests/emu -I include -I asset/cuda sq_limit_check.cpp
#include "cuda_runtime.h"
emu_idx threadIdx, blockIdx, blockDim, gridDim;
unsigned long long emu_threads_run = 0;
unsigned long long emu_ldg_count = 0;
char emu_last_kernel[256];
float emu_clock_ms = 0.f, emu_cost_ms[8];
int emu_fail_mode = -1;
#include "abl_device.cuh"

#ifdef ABL_USE_FLOAT
typedef uint32_t bits_t;
#else
typedef uint64_t bits_t;
#endif
static abl_real val(bits_t b) { abl_real x; memcpy(&x, &b, sizeof x); return x; }
static bits_t bits(abl_real x) { bits_t b; memcpy(&b, &x, sizeof b); return b; }
static uint64_t rng = 0x9e3779b97f4a7c15ull;
static uint64_t next() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }

static long checked = 0, failed = 0;
static bool reference(int op, abl_real s, abl_real C) {
  const abl_real d = abl_sqrt_narrow(s);
  return op == 0 ? d < C : op == 1 ? d <= C : op == 2 ? d > C : d >= C;
}
static void check(int op, abl_real C, abl_real limit, abl_real s) {
  const bool want = reference(op, s, C);
  const bool got = op <= 1 ? s <= limit : s >= limit;
  checked++;
  if (want != got) {
    if (failed++ < 10) fprintf(stderr, "op %d C=%.17g s=%.17g limit=%.17g: want %d got %d\n", op, (double)C, (double)s, (double)limit, want, got);
  }
}

int main() {
  const abl_real consts[] = {ABL_R(0.005), ABL_R(0.05), ABL_R(1.0), ABL_R(5.0), ABL_R(10.0), ABL_R(0.0), ABL_R(-1.0),
                             ABL_R(1e-30), ABL_R(3.4e38), ABL_R(1e30), (abl_real)INFINITY, ABL_R(1.5), ABL_R(0.12),
                             ABL_R(2.0), ABL_R(1e-20), ABL_R(123456.789)};
  for (abl_real C : consts) {
    for (int op = 0; op < 4; op++) {
      const abl_real limit = abl_sq_cmp_limit(op, C);
      // special values
      const abl_real specials[] = {ABL_R(0.0), (abl_real)INFINITY, (abl_real)NAN, val(1), val(2), C * C, C};
      for (abl_real s : specials) if (!(s < 0)) check(op, C, limit, s);
      // every representable value within 64 ulps of the bound
      if (limit == limit && limit >= 0 && limit != (abl_real)INFINITY) {
        const bits_t b = bits(limit);
        for (int k = -64; k <= 64; k++) {
          if (k < 0 && b < (bits_t)(-k)) continue;
          const abl_real s = val(b + (bits_t)k);
          if (s == s) check(op, C, limit, s);
        }
      }
      // random non-negative values: all exponents, and near C*C
      for (int k = 0; k < 200000; k++) {
        bits_t r = (bits_t)next() & (bits(( abl_real)INFINITY) - 1);
        check(op, C, limit, val(r));
        const abl_real near = C * C * (abl_real)(1.0 + ((double)(next() % 2001) - 1000.0) * 1e-7);
        if (near >= 0) check(op, C, limit, near);
      }
    }
    // the radius filter keeps a candidate unless dist > R  (NaN passes: `d2 > limit` is false)
    const abl_real lim = abl_near_sq_limit(C);
    for (int k = 0; k < 200000; k++) {
      const abl_real s = k & 1 ? val((bits_t)next() & (bits((abl_real)INFINITY) - 1))
                               : C * C * (abl_real)(1.0 + ((double)(next() % 2001) - 1000.0) * 1e-7);
      if (!(s >= 0)) continue;
      const bool want = !(abl_sqrt_narrow(s) > C), got = !(s > lim);
      checked++;
      if (want != got && failed++ < 10) fprintf(stderr, "near limit R=%.17g s=%.17g\n", (double)C, (double)s);
    }
  }
  printf("%ld checks, %ld failures\n", checked, failed);
  return failed ? 1 : 0;
}
