// tests/emu/emu_kernels.cpp — TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
//
// Compiles the generated device program of one model for the host and exposes its step
// launchers to Python: the three runtime calls abl_model_setup() makes are redirected to
// recorders below, so the harness learns the environment, the pools and — through the
// abl_step_desc the generated code registers — the launcher of every step function, exactly
// as the device runtime does (asset/cuda/abl_runtime.cu: abl_cuda_register_step).
#include "cuda_runtime.h"

emu_idx threadIdx, blockIdx, blockDim, gridDim;
unsigned long long emu_threads_run = 0;
unsigned long long emu_ldg_count = 0;
char emu_last_kernel[256];
float emu_clock_ms = 0.f, emu_cost_ms[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
int emu_fail_mode = -1;

#define abl_cuda_set_environment emu_set_environment
#define abl_cuda_add_pool emu_add_pool
#define abl_cuda_register_step emu_register_step
#define abl_cuda_step emu_step_unused

#include "model_kernels.cu"

// the dynamic shared memory of the generated kernels (ABL_MODE 1 acceptance masks; the tile
// kernels' buffer is declared but never used here)
unsigned _abl_masks[ABL_MASK_WORDS * 1024];
__attribute__((aligned(16))) unsigned char _abl_smem[64 * 1024];

namespace {
struct Env { int dim; double lo[3], hi[3], cell; bool set; } g_env;
struct StepRec { abl_step_desc desc; };
StepRec g_steps[64];
int g_n_steps = 0;
int g_n_pools = 0;
struct { int phase; unsigned *cnt, *idx, stride, *max; } g_nlist = {0, nullptr, nullptr, 0, nullptr};
abl_slab_view g_slab;   // zero-initialised: inactive
}

extern "C" int emu_set_environment(abl_runtime *, int dim, const double *env_min, const double *env_max, double granularity) {
  g_env.dim = dim;
  for (int a = 0; a < 3; a++) { g_env.lo[a] = a < dim ? env_min[a] : 0; g_env.hi[a] = a < dim ? env_max[a] : 0; }
  g_env.cell = granularity;
  g_env.set = true;
  return 0;
}
extern "C" int emu_add_pool(abl_runtime *, const abl_agent_desc *, int *pool) { *pool = g_n_pools++; return 0; }
extern "C" int emu_register_step(abl_runtime *, const abl_step_desc *desc, int *step) {
  g_steps[g_n_steps].desc = *desc;
  *step = g_n_steps++;
  return 0;
}
extern "C" int emu_step_unused(abl_runtime *, int) { return -1; }

extern "C" {

int emu_setup(void) {
  g_n_steps = 0; g_n_pools = 0; g_env.set = false;
  return abl_model_setup(nullptr);
}

// environment as registered by the model: dim, min[3], max[3], granularity (0 if none)
int emu_environment(int *dim, double *lo, double *hi, double *cell) {
  if (!g_env.set) return 1;
  *dim = g_env.dim;
  for (int a = 0; a < 3; a++) { lo[a] = g_env.lo[a]; hi[a] = g_env.hi[a]; }
  *cell = g_env.cell;
  return 0;
}

int emu_step_info(int s, int *self_pool, int *nbr_pool, double *radius, unsigned *written, int *uses_removal, int *added_pool) {
  if (s < 0 || s >= g_n_steps) return 1;
  const abl_step_desc &d = g_steps[s].desc;
  *self_pool = d.self_pool; *nbr_pool = d.nbr_pool; *radius = d.radius; *written = d.written_members;
  *uses_removal = d.uses_removal; *added_pool = d.added_pool;
  return 0;
}

// neighbour-list arguments of the next emu_run_step calls (abl_step_launch.nlist_*)
void emu_set_nlist(int phase, unsigned *cnt, unsigned *idx, unsigned stride, unsigned *max) {
  g_nlist.phase = phase; g_nlist.cnt = cnt; g_nlist.idx = idx; g_nlist.stride = stride; g_nlist.max = max;
}
int emu_step_nlist(int s) { return s >= 0 && s < g_n_steps ? g_steps[s].desc.nlist : 0; }

// Slab decomposition as the runtime's halo_fill_view sets it up for the fused send (no
// boundary-first scheduling, no device range): owned layers [begin, end), the layer ranges of
// the lower / upper peer ([0,0) = none), ghost flags, message areas and slot counters.
void emu_set_slab(int active, int dim, int pos_col, int n_layers, double origin, double inv_cell, int begin, int end,
                  int ghost, int lo_begin, int lo_end, int hi_begin, int hi_end, int lo_ghost, int hi_ghost,
                  unsigned char *msg_lo, unsigned char *msg_hi, unsigned *counters, unsigned capacity,
                  unsigned rec_words, int n_cols, const int *elem) {
  memset(&g_slab, 0, sizeof g_slab);
  if (!active) return;
  g_slab.active = 1; g_slab.dim = dim; g_slab.pos_col = pos_col; g_slab.n_layers = n_layers;
  g_slab.origin = origin; g_slab.inv_cell = inv_cell;
  g_slab.begin = begin; g_slab.end = end; g_slab.ghost = ghost;
  g_slab.lo_begin = lo_begin; g_slab.lo_end = lo_end; g_slab.hi_begin = hi_begin; g_slab.hi_end = hi_end;
  g_slab.lo_ghost = lo_ghost; g_slab.hi_ghost = hi_ghost;
  g_slab.msg[0] = msg_lo; g_slab.msg[1] = msg_hi;
  g_slab.count[0] = counters; g_slab.count[1] = counters + 1; g_slab.far = counters + 2;
  g_slab.late = counters + 3;
  g_slab.capacity = capacity; g_slab.rec_words = rec_words;
  g_slab.n_cols = n_cols;
  for (int c = 0; c < n_cols; c++) g_slab.elem[c] = elem[c];
}

int emu_real_size(void) { return (int)sizeof(abl_real); }
unsigned long long emu_thread_count(void) { return emu_threads_run; }
unsigned long long emu_load_count(void) { return emu_ldg_count; }
const char *emu_last_kernel_name(void) { return emu_last_kernel; }
void emu_set_fail_mode(int mode) { emu_fail_mode = mode; }
void emu_set_cost(int mode, float ms) { if (mode >= 0 && mode < 8) emu_cost_ms[mode] = ms; }

// Runs step function `s` over the given (already binned) pools: the launcher the generated code
// registered picks the kernel variant and the block size like on the device.
int emu_run_step(int s, const abl_pool_view *self, const abl_pool_view *nbr, const abl_grid_view *grid, int reach,
                 unsigned char *dead, unsigned char *add_flag, void *const *add_cols,
                 unsigned *bin_key, unsigned *bin_local, unsigned *bin_count,
                 unsigned long long seed, unsigned timestep, int block_size, int flat_loop) {
  if (s < 0 || s >= g_n_steps) return 1;
  abl_step_launch a;
  memset(&a, 0, sizeof a);
  a.self = *self;
  if (nbr) a.nbr = *nbr;
  a.grid = *grid;
  a.reach = reach;
  a.dead = dead;
  a.add_flag = add_flag;
  if (add_cols) for (int c = 0; c < ABL_MAX_COLUMNS; c++) a.add_cols[c] = add_cols[c];
  a.bin_key = bin_key; a.bin_local = bin_local; a.bin_count = bin_count;
  a.seed = seed;
  a.timestep = timestep;
  a.step_index = (unsigned)s;
  a.block_size = block_size;
  a.tile_neighbours = 0;   // block-cooperative kernels cannot be emulated sequentially
  a.flat_loop = flat_loop;
  a.slab = g_slab;
  a.nlist_phase = g_nlist.phase; a.nlist_cnt = g_nlist.cnt; a.nlist_idx = g_nlist.idx;
  a.nlist_stride = g_nlist.stride; a.nlist_max = g_nlist.max;
  a.pdl = 0;
  return a.self.n ? g_steps[s].desc.launch(&a) : 0;
}

}
