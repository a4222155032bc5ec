// abl_runtime.cu — B200-native OpenABL runtime (libabl_cuda.so), C ABI in abl_cuda.h.
//
// Owns: SoA agent pools with per-column ping-pong buffers in HBM, uniform-grid binning
// (single-digit radix = counting sort over the cell key, ties ordered by agent id), the
// step-function launch protocol, stream compaction for removeCurrent()/add(), reductions,
// and host AoS <-> device SoA transfer.  Generated step kernels (abl_device.cuh) are
// launched through the launcher registered with each step.
//
// Reference constructs replaced (see include/abl_cuda.h for the per-function mapping):
// dyn_array storage + double buffer (asset/c/libabl.h:11-57, CPrinter.cpp:210-228) and the
// brute-force neighbour scan (CPrinter.cpp:160-171).
//
// Kernel inventory (all HBM/L2-bound integer/byte work; no tensor cores by design):
//   k_aos_to_soa / k_soa_to_aos   host record transpose (upload / download)
//   k_bin_count                   cell key + per-cell histogram (returns arrival rank)
//   k_scan<...>                   single-pass decoupled-look-back exclusive scan
//   k_bin_scatter                 (id, source index) pairs into cell segments
//   k_bin_rank_move               rank by id inside the segment + gather all columns
//   k_compact_move / k_append     stream compaction for remove / add
//   k_reduce_*                    count / sum reductions
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "abl_cuda.h"
#include "abl_slab.cuh"

typedef unsigned int u32;
typedef unsigned long long u64;
typedef unsigned char u8;

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(ABL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                  __FILE__, __LINE__);                                                    \
  } while (0)

#define TRY(call)              \
  do {                         \
    int r_ = (call);           \
    if (r_ != ABL_OK) return r_; \
  } while (0)

// ---------------------------------------------------------------------------------------
// data structures
// ---------------------------------------------------------------------------------------
struct Column {
  int elem = 0;       // bytes per agent in this column (1, 4, 8 or 16)
  int comp = 0;       // bytes per scalar component
  int ncomp = 0;      // components packed in one element
  int host_off = -1;  // byte offset inside the host AoS record (-1: not part of it, i.e. ids)
  void *buf[2] = {nullptr, nullptr};
  int cur = 0;
};

struct Member {
  int type = 0;
  int first_col = 0;
  int ncols = 0;
  std::string name;
};

struct Pool {
  std::string name;
  std::vector<Member> members;
  std::vector<Column> cols;  // user columns followed by the id column
  int id_col = -1;
  int pos_member = -1;
  unsigned stride = 0;
  size_t n = 0, cap = 0;
  // single-precision shadow of the positions (ABL_MODE 8): float4 per record in pool order, valid while
  // shadow_serial == layout_serial (layout_serial counts every change of the order or the positions)
  float4 *shadow = nullptr;
  u32 *shadow_max = nullptr;
  size_t shadow_cap = 0;
  u64 shadow_serial = ~(u64)0;   // order_serial the shadow was built for
  u32 next_id = 0;
  // slab decomposition: [own_begin, own_end) is the owned part of the (binned) pool;
  // src_begin is where the live records start before the next binning compacts them
  u32 own_begin = 0, own_end = 0, src_begin = 0;
  u32 bnd_lo_end = 0, bnd_hi_begin = 0;   // pool indices delimiting the boundary parts of the owned range
  // The report of the last slab-mode binning (owned range, boundary parts, counters of the
  // preceding halo exchange) lands in a per-pool slice of mapped host memory.  With device-side
  // ranges the host consumes it lazily: at the latest when the next binning needs the pool size.
  int index = 0;                  // position in rt->pools
  bool report_pending = false;
  u32 report_stamp = 0;
  bool exch_deferred = false;     // a direct exchange was queued without the host knowing the owned range
  u32 exch_pad = 0;               // padding of that exchange
  u32 pad_used[2] = {0, 0};       // padding of the exchanges in flight, by parity of their sequence number
  bool halo_sync_check = false;   // verify the counts of the last exchange synchronously at the next binning (stand-alone exchanges)
  u32 key_base_hint = 0;  // copy of grid.key_base (cell_start is handed out with a virtual origin)
  bool own_valid = false;
  // direct halo transport (peer memory over NVLink, no NCCL, no host sync per exchange)
  u8 *halo_recv = nullptr;                 // own receive area: [from lower | from upper] x [parity 0 | 1]
  size_t halo_cap = 0;                     // records per block
  size_t halo_block = 0;                   // bytes per block
  u8 *halo_peer[2] = {nullptr, nullptr};   // receive areas of the lower / upper peer, mapped here
  bool halo_ipc[2] = {false, false};
  u32 *halo_ctr = nullptr;                 // [0,1] slot counters, [2,3] incoming counts, [4] timeout, [5] far, [6,7] sent
  u32 halo_seq = 0;                        // direct exchanges performed (parity selects the block)
  u32 halo_pad = 0;                        // host-side upper bound of arrivals (rest is sentinel padding)
  bool halo_pending = false;               // device-side counts of the last exchange not verified yet
  u32 halo_prev_ob = 0, halo_prev_own = 0, halo_last_arrivals = 0, halo_room = 0;
  u64 order_serial = 0;   // bumped whenever records are moved, added or removed (validity of cached neighbour lists)
  bool binned = false;
  bool ever_binned = false;
  bool counted = false;   // key/local/cell_count already hold the histogram of the current positions
  bool ever_removed = false;
  // binning scratch (per pool so that pools can be binned independently)
  u32 *key = nullptr, *local = nullptr;
  u64 *pairs = nullptr;
  u32 *cell_count = nullptr, *cell_start = nullptr;
  // remove / add scratch
  u8 *dead = nullptr;
  u8 *add_flag = nullptr;
  u32 *offsets = nullptr;
  size_t scratch_cap = 0;
};

struct ScanState {
  u64 *desc = nullptr;   // one descriptor per tile: (epoch<<2 | status) << 32 | value
  u32 *ctrl = nullptr;   // [0] dynamic tile counter, [1] finished tiles, [2] epoch
  u32 *tile_sum = nullptr;  // per-tile sums of the two-pass scan
  size_t max_tiles = 0;
};

// Cached neighbour lists of one step function (abl_step_desc.nlist): per agent of the stepped
// pool the pool indices of its accepted candidates, k-major (word k * stride + i), valid while
// neither pool of the for-near loop is reordered, resized or re-uploaded.
struct NeighbourLists {
  u32 *cnt = nullptr, *idx = nullptr;
  size_t cnt_cap = 0, idx_cap = 0;     // words
  u32 stride = 0, max_degree = 0;
  u64 self_serial = 0, nbr_serial = 0;
  size_t n_self = 0, n_nbr = 0;
  bool valid = false;
  bool off = false;                    // lists would exceed the memory budget: use the ordinary loops
  unsigned builds = 0;
};

static const int kPfWords = 56, kPfRows = 12;   // ABL_SHADOW_WORDS / ABL_SHADOW_ROWS of abl_device.cuh
struct Step {
  u32 *pf_buf = nullptr;     // scratch of the split pre-filter (ABL_MODE 9)
  size_t pf_cap = 0;         // words
  abl_step_desc desc;
  std::string name;
  int reach = 1;
  NeighbourLists nl;
};

struct ColTable {
  int ncols;
  int elem[ABL_MAX_COLUMNS + 1];
  int comp[ABL_MAX_COLUMNS + 1];
  int ncomp[ABL_MAX_COLUMNS + 1];
  int host_off[ABL_MAX_COLUMNS + 1];
  const void *in[ABL_MAX_COLUMNS + 1];
  void *out[ABL_MAX_COLUMNS + 1];
};

struct GridParams {
  int dim;
  int n_cell[3];
  double origin[3];
  double cell;
  double inv_cell;  // 1/cell in the precision of abl_float
  u32 n_cells;
  // window of cell layers (slowest axis) this runtime keeps cell ranges for: everything in a
  // single-GPU run, the own slab plus ghost layers under slab decomposition.  Keys stored in
  // the pool are relative to key_base; coordinates outside the window are clamped into it.
  int axis_lo, axis_hi;
  u32 key_base, n_local;
};

struct abl_runtime {
  abl_config cfg;
  int device = 0;
  int real_size = 8;
  cudaStream_t stream = nullptr;
  bool env_set = false;
  GridParams grid;
  std::vector<Pool> pools;
  std::vector<Step> steps;
  ScanState scan;
  void *stage = nullptr;       // device staging for AoS transfers
  int replay_reps = 0;         // abl_cuda_time_kernel: the next abl_cuda_step times its kernel over this many extra launches
  float replay_ms = 0.f;       // ... average duration of one of them
  u32 *rank_buf = nullptr;     // presence flags + their scan, one word per agent id each (commit_adds with many adds)
  size_t rank_cap = 0;         // words
  size_t stage_cap = 0;
  void *pinned = nullptr;      // pinned host bounce buffer
  size_t pinned_cap = 0;
  u32 *d_scalar = nullptr;     // small device scratch for totals / reductions
  u32 *h_scalar = nullptr;     // pinned mirror
  u32 *h_scalar_dev = nullptr; // the same memory as seen from the device (mapped)
  unsigned timestep = 0;
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_ts[2] = {nullptr, nullptr};
  cudaEvent_t ev_own = nullptr;  // signals that the owned-range words have reached the host
  bool ts_open = false;
  double last_ts_seconds = 0;
  std::vector<std::pair<void *, size_t>> pinned_ranges;
  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  bool slab = false;
  int layer_begin = 0, layer_end = 0;   // owned cell layers along the slab axis
  int n_slabs = 1, my_slab = 0;
  int slab_bounds[ABL_MAX_SLABS + 1] = {0};
  int ghost_layers = 1;
  void *xbuf[4] = {nullptr, nullptr, nullptr, nullptr};  // send L, send R, recv L, recv R
  size_t xcap[4] = {0, 0, 0, 0};
  u32 *xflags = nullptr;  // classification scratch
  // optional phase trace of exchange() (ABL_CUDA_TRACE=1): device and host time per phase
  bool trace = false;
  cudaEvent_t xev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double xdev[4] = {0, 0, 0, 0}, xhost[4] = {0, 0, 0, 0};
  unsigned long xcount = 0, xskip = 0;
  unsigned long long *trace_state = nullptr;   // device: [0] last stamp, [1..8] ns per bucket, [9..16] counts
  double own_wait_s = 0;
  unsigned long own_waits = 0;
  u32 x_out[2] = {0, 0};  // outgoing counts of the last exchange_pack (to lower, to upper)
  abl_runtime *peer_lo = nullptr, *peer_hi = nullptr;  // in-process transport (tests)
  // cudaFree synchronises the whole device; with the direct transport a neighbour driven by the
  // same host thread may be spinning in k_halo_wait, so buffers replaced while growing a pool
  // are released at the next explicit synchronisation point instead
  u32 bin_stamp = 0;           // number of the last slab-mode binning (stamps of the owned-range report)
  // ABL_CUDA_DEVICE_RANGE=1: step and exchange kernels read the owned range from cell_start on
  // the device, so the host can queue them without waiting for the binning report.  Off by
  // default: measured slightly slower on 2 x B200 (boids2d 1 M agents per GPU, 0.1511 against
  // 0.1467 ms per step, profiles/scaling/r1d_weak2_device_range_*.json)
  bool device_range = false;
  bool halo_overlap = true;    // ABL_CUDA_HALO_OVERLAP=0: publish after the whole step kernel instead of boundary-first
  // ABL_CUDA_HALO_ASYNC=0: the exchange kernel follows the step kernel on the runtime's stream.  Default: it runs
  // NEXT TO the step kernel on a high-priority side stream (forked before the step kernel is queued, joined before
  // the next binning) whenever the step kernel publishes by itself (boundary-first scheduling): the neighbours'
  // records arrive while the interior is still being stepped, so waiting for them, appending them and entering
  // them into the histogram costs the step nothing.
  bool halo_async = true;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pdl = true;             // ABL_CUDA_PDL=0 turns programmatic dependent launches off
  bool report_in_scan = false; // ABL_CUDA_REPORT_IN_SCAN=1: the owned range is reported by k_tile_scan instead of k_bin_scatter
  bool mbar_hint = false;      // ABL_CUDA_MBAR_HINT=1: bulk-tile kernels wait for their tile with a suspend-time hint (pdl bit 2)
  bool pdl_trigger = false;    // ABL_CUDA_PDL_TRIGGER=1: successors become resident while a kernel's last wave runs (measured neutral, off)
  bool nlist = true;           // ABL_CUDA_NLIST=0: ignore abl_step_desc.nlist (A/B against the ordinary loops)
  size_t nlist_budget = (size_t)8 << 30;   // ABL_CUDA_NLIST_MB: largest index array of one step function
  // Candidate loop of sparse 2-D step kernels: 1 (default) the flat loop — from a bulk-staged tile when
  // bulk_tile is set — chosen by the launcher's density rule (measured on B200 in round 2: flat beats
  // the cursor loop on boids2d f64/f32, game_of_life and predator_prey; the chunked loop wins only for
  // dense rows); ABL_CUDA_FLAT=0 pins the cursor loop; ABL_CUDA_TUNE=1 (-1) lets the launcher time the
  // plausible variants over its first launches instead.
  int flat_loop = 1;
  int bulk_tile = 1;           // ABL_CUDA_BULK=0: no TMA-staged tiles (ABL_MODE 7)
  int dense_tile = 1;          // ABL_CUDA_DENSE=0: no single-precision shadow of the positions (ABL_MODE 8)
  int split_prefilter = 1;     // ABL_CUDA_SPLIT=0: the shadow pre-filter stays inside the step kernel (no ABL_MODE 9)
  bool scan_two_pass = true;   // ABL_CUDA_SCAN=lookback selects the single-pass scan for the cell histogram
  std::vector<void *> garbage;
  bool defer_free = false;
  long long halo_timeout_ns = 10000000000ll;            // direct transport: wait for a neighbour at most this long
  size_t xflags_cap = 0;
  abl_step_timing last = {0, 0, 0, 0};
  unsigned launches = 0;
  // slab decomposition + run-time add(): ids of new agents are global, so a step function that
  // adds agents stays open until the caller has supplied the ranks of its parents among the
  // parents of all slabs (abl_cuda_pending_adds / abl_cuda_resolve_adds)
  int open_step = -1;
  std::vector<u64> open_list;               // (parent id << 32 | slot relative to the owned range), ascending
  void *open_staging[ABL_MAX_COLUMNS + 1];
  // timing mode: the four events of every abl_cuda_step are queued here and evaluated when the
  // caller asks (abl_cuda_last_timing) — no host synchronisation between the timed steps, so the
  // GPU is never idle between them and the stages are timed as they run in a real simulation
  std::vector<cudaEvent_t> timing_log;
  // combines rank-local reduction results across slabs (installed by the harness)
  abl_reduce_hook reduce_hook = nullptr;
  void *reduce_user = nullptr;
};

// Programmatic dependent launch: the kernel may be set up on the SMs while its predecessor in
// the stream is still draining; it calls cudaGridDependencySynchronize() before it touches
// memory.  Hides the launch latency between the short kernels of the binning chain.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ABL_CUDA_PDL_TRIGGER=1 (default 0): every kernel of the per-step chain lets the blocks of its successor become
// resident as soon as its own last wave is running (they wait in cudaGridDependencySynchronize).  Same-box A/B on
// a B200, boids2d 1 M: 0.0989 / 0.0993 ms per step without / with — neutral, hence off.
__constant__ int c_pdl_trigger;
// ABL_CUDA_BIN_PREFETCH (default 1): k_bin_rank_move requests the lines of a record before it ranks it
__constant__ int c_bin_prefetch;
// ABL_CUDA_BIN_SEGINFO=1 (default 0): k_bin_scatter leaves begin and size of every agent's cell segment per agent
// (coalesced), so that k_bin_rank_move reaches the segment's ids after one round trip instead of two
// (key -> cell_start -> ids becomes {begin, size} -> ids).  Same-box A/B on a B200, boids2d 1 M: 0.0916 / 0.0912 ms
// per timestep with the L2 flushed, 0.0828 / 0.0836 ms steady — the 8 bytes per agent cost what the round trip
// saves; off.
__constant__ int c_bin_seginfo;

static const int kScanBlock = 256;
static const int kScanItems = 16;                      // per thread
static const int kScanTile = kScanBlock * kScanItems;  // 4096

static inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// per-pool slice of the mapped host buffer the binning reports into (32 words each, after the
// first 64 words that other transfers use)
static u32 *report_host(abl_runtime *rt, const Pool &p) { return rt->h_scalar + 64 + 32 * (p.index % 28); }
static u32 *report_dev(abl_runtime *rt, const Pool &p) { return rt->h_scalar_dev + 64 + 32 * (p.index % 28); }


// ---------------------------------------------------------------------------------------
// kernels: transpose between host records and columns
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_scalar(void *dst, const void *src, int bytes) {
  switch (bytes) {
    case 1: *(u8 *)dst = *(const u8 *)src; break;
    case 4: *(u32 *)dst = *(const u32 *)src; break;
    default: *(u64 *)dst = *(const u64 *)src; break;
  }
}

__global__ void k_aos_to_soa(ColTable t, const u8 *aos, u32 stride, u32 n, u32 first_id) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8 *rec = aos + (size_t)i * stride;
  for (int c = 0; c < t.ncols; c++) {
    u8 *dst = (u8 *)t.out[c] + (size_t)i * t.elem[c];
    if (t.host_off[c] < 0) { *(u32 *)dst = first_id + i; continue; }
    for (int k = 0; k < t.ncomp[c]; k++)
      copy_scalar(dst + k * t.comp[c], rec + t.host_off[c] + k * t.comp[c], t.comp[c]);
  }
}

// dest_rank == nullptr: record i goes to slot id[i]; otherwise to dest_rank[id[i]]
__global__ void k_soa_to_aos(ColTable t, u8 *aos, u32 stride, u32 n, const u32 *ids,
                             const u32 *dest_rank) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 slot = ids[i];
  if (dest_rank) slot = dest_rank[slot];
  u8 *rec = aos + (size_t)slot * stride;
  for (int c = 0; c < t.ncols; c++) {
    if (t.host_off[c] < 0) continue;
    const u8 *src = (const u8 *)t.in[c] + (size_t)i * t.elem[c];
    for (int k = 0; k < t.ncomp[c]; k++)
      copy_scalar(rec + t.host_off[c] + k * t.comp[c], src + k * t.comp[c], t.comp[c]);
  }
}

__global__ void k_mark_present(const u32 *ids, u32 n, u32 *present) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) present[ids[i]] = 1;
}

// ---------------------------------------------------------------------------------------
// kernels: single-pass exclusive scan (decoupled look-back)
// ---------------------------------------------------------------------------------------
// Each CTA takes the next tile (dynamic tile index => earlier tiles are always running or
// done, so spinning on them cannot deadlock), scans it locally, publishes its aggregate,
// looks back over predecessors' descriptors to obtain its exclusive prefix, then publishes
// its inclusive prefix.  Descriptors carry an epoch so the array never needs clearing; the
// last CTA to finish resets the tile counter and bumps the epoch, which keeps the kernel
// re-launchable without any host-side memset (CUDA-graph friendly).
enum { SCAN_INVALID = 0, SCAN_AGGREGATE = 1, SCAN_PREFIX = 2 };

// Loads 16 consecutive items of thread `tid` (64 bytes of u32, 16 bytes of u8).
template <typename T> struct ScanLoad;
template <> struct ScanLoad<u32> {
  static __device__ __forceinline__ void load16(const u32 *in, size_t i, u32 v[16]) {
    const uint4 *p = reinterpret_cast<const uint4 *>(in + i);
    uint4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];  // four independent 16-byte loads
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
    v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
    v[12] = q3.x; v[13] = q3.y; v[14] = q3.z; v[15] = q3.w;
  }
  static __device__ __forceinline__ void zero16(u32 *in, size_t i) {
    uint4 *p = reinterpret_cast<uint4 *>(in + i);
    p[0] = p[1] = p[2] = p[3] = make_uint4(0, 0, 0, 0);
  }
};
template <> struct ScanLoad<u8> {
  static __device__ __forceinline__ void load16(const u8 *in, size_t i, u32 v[16]) {
    uint4 q = *reinterpret_cast<const uint4 *>(in + i);
    u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
  }
  static __device__ __forceinline__ void zero16(u8 *in, size_t i) {
    *reinterpret_cast<uint4 *>(in + i) = make_uint4(0, 0, 0, 0);
  }
};

// MODE 0: out[i] = exclusive prefix of in[i]
// MODE 1: in[] holds "dead" flags; scans (in[i] == 0), i.e. survivors
template <typename T, int MODE, bool ZERO_INPUT>
__global__ void __launch_bounds__(kScanBlock)
k_scan(T *in, u32 *out, u32 n, u64 *desc, u32 *ctrl, u32 *total_out) {
  __shared__ u32 s_tile;
  __shared__ u32 s_warp[kScanBlock / 32];
  __shared__ u32 s_prefix;
  __shared__ int s_first;
  __shared__ u32 s_sum;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&ctrl[0], 1u);
  __syncthreads();
  const u32 tile = s_tile;
  const u32 epoch = *((volatile u32 *)&ctrl[2]);
  const size_t i0 = (size_t)tile * kScanTile + (size_t)tid * kScanItems;

  // thread-local scan of 16 consecutive items (arrays are padded to a multiple of the
  // tile, so vector accesses past n stay in bounds; values past n count as 0)
  u32 v[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) v[k] = 0;
  if (i0 < n) {
    ScanLoad<T>::load16(in, i0, v);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      if (MODE == 1) v[k] = v[k] ? 0u : 1u;
      if (i0 + k >= n) v[k] = 0;
    }
    if (ZERO_INPUT) ScanLoad<T>::zero16(in, i0);
  }
  u32 t = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { u32 x = v[k]; v[k] = t; t += x; }  // v = local exclusive
  u32 inc = t;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  u32 wsum = lane < kScanBlock / 32 ? s_warp[lane] : 0;
  u32 winc = wsum;
#pragma unroll
  for (int d = 1; d < kScanBlock / 32; d <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, winc, d);
    if (lane >= d) winc += y;
  }
  const u32 warp_excl = __shfl_sync(0xffffffffu, winc - wsum, warp);
  const u32 aggregate = __shfl_sync(0xffffffffu, winc, kScanBlock / 32 - 1);
  const u32 thread_excl = warp_excl + inc - t;

  // publish, then block-wide look-back: 256 predecessor descriptors per round trip
  if (tid == 0) {
    u32 st = tile == 0 ? SCAN_PREFIX : SCAN_AGGREGATE;
    *((volatile u64 *)&desc[tile]) = ((u64)((epoch << 2) | st) << 32) | aggregate;
    s_prefix = 0;
  }
  u32 prefix = 0;
  if (tile > 0) {
    int idx = (int)tile - 1;
    for (;;) {
      if (tid == 0) { s_first = kScanBlock; s_sum = 0; }
      __syncthreads();
      int my = idx - tid;
      u32 status = SCAN_PREFIX, value = 0;  // tiles before 0 behave as prefix 0
      if (my >= 0) {
        for (;;) {
          u64 d = *((volatile u64 *)&desc[my]);
          u32 hi = (u32)(d >> 32);
          if ((hi >> 2) == epoch && (hi & 3u) != SCAN_INVALID) {
            status = hi & 3u;
            value = (u32)d;
            break;
          }
        }
      }
      if (status == SCAN_PREFIX) atomicMin(&s_first, tid);
      __syncthreads();
      const int first = s_first;  // nearest predecessor holding a full prefix
      u32 contrib = tid <= first ? value : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
      if (lane == 0 && contrib) atomicAdd(&s_sum, contrib);
      __syncthreads();
      prefix += s_sum;
      if (first < kScanBlock) break;
      idx -= kScanBlock;
      __syncthreads();
    }
    if (tid == 0)
      *((volatile u64 *)&desc[tile]) =
          ((u64)((epoch << 2) | SCAN_PREFIX) << 32) | (u32)(prefix + aggregate);
  }

  if (i0 < n) {
    const u32 b = prefix + thread_excl;
    uint4 *o = reinterpret_cast<uint4 *>(out + i0);
    o[0] = make_uint4(v[0] + b, v[1] + b, v[2] + b, v[3] + b);
    o[1] = make_uint4(v[4] + b, v[5] + b, v[6] + b, v[7] + b);
    o[2] = make_uint4(v[8] + b, v[9] + b, v[10] + b, v[11] + b);
    o[3] = make_uint4(v[12] + b, v[13] + b, v[14] + b, v[15] + b);
  }

  // bookkeeping: total + self-reset by the last CTA
  const u32 ntiles = gridDim.x;
  if (tid == 0) {
    if (tile == ntiles - 1 && total_out) *total_out = prefix + aggregate;
    __threadfence();
    u32 done = atomicAdd(&ctrl[1], 1u);
    if (done == ntiles - 1) {
      ctrl[0] = 0;
      ctrl[1] = 0;
      ctrl[2] = (epoch + 1) & 0x3fffffffu;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------
// Two-pass scan of the cell histogram (reduce-then-scan): k_tile_sum writes one sum per tile of
// 4096 counters, k_tile_scan lets every CTA add up the sums of the tiles before its own (a few
// hundred words that sit in L2) and scan its tile.  No CTA ever waits for another one, which
// for the few hundred tiles of a cell grid beats the look-back chain of the single-pass scan
// above (all of whose CTAs are resident at once and mostly poll).  The histogram is left in
// place: k_bin_scatter draws the slots of every cell segment by counting it back down to zero.
__global__ void __launch_bounds__(kScanBlock) k_tile_sum(const u32 *in, u32 n, u32 *tile_sum) {
  __shared__ u32 s_warp[kScanBlock / 32];
  if (c_pdl_trigger) cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t i0 = (size_t)blockIdx.x * kScanTile + (size_t)tid * kScanItems;
  u32 t = 0;
  if (i0 < n) {
    u32 v[kScanItems];
    ScanLoad<u32>::load16(in, i0, v);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) t += (i0 + k < n) ? v[k] : 0u;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if (lane == 0) s_warp[warp] = t;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < kScanBlock / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if (lane == 0) tile_sum[blockIdx.x] = w;
  }
}

// Slab decomposition: the two words of cell_start that delimit the owned range (and the
// counters of the last halo exchange) are written to page-locked host memory by the very
// threads that produce them (`report.host` mapped into the device address space), so the host
// learns them without a further kernel or copy.
// Every value is followed (after a system-scope fence) by a stamp word carrying the number of
// this binning, and the host simply polls the stamps in its own memory: no event, no driver
// call and no copy engine between two kernels of the stream.
// host words: [0] own begin, [1] own end, [2..7] halo counters, [8] lower boundary end,
// [9] upper boundary begin, [10] late, [11] sequence number of the exchange the counters belong
// to, [12] counters present; stamps at [16] own begin, [17] own end, [18], [19] boundaries,
// [20] counters.
struct ScanReport {
  u32 stamp;
  u32 *host;            // nullptr: nothing to report
  u32 lo_cell, hi_cell;
  u32 lo2_cell, hi2_cell;  // end of the lower / begin of the upper boundary part (boundary-first scheduling)
  const u32 *halo_ctr;  // may be nullptr
};

__global__ void __launch_bounds__(kScanBlock) k_tile_scan(u32 *in, u32 *out, u32 n, const u32 *tile_sum,
                                                          ScanReport report) {
  __shared__ u32 s_warp[kScanBlock / 32];
  __shared__ u32 s_pre[kScanBlock / 32];
  if (c_pdl_trigger) cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 tile = blockIdx.x;
  // sum of all tiles before this one
  u32 pre = 0;
  for (u32 q = tid; q < tile; q += kScanBlock) pre += tile_sum[q];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, d);
  if (lane == 0) s_pre[warp] = pre;
  const size_t i0 = (size_t)tile * kScanTile + (size_t)tid * kScanItems;
  u32 v[kScanItems];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) v[k] = 0;
  if (i0 < n) {
    ScanLoad<u32>::load16(in, i0, v);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (i0 + k >= n) v[k] = 0;
    // (the histogram stays: k_bin_scatter counts it back down to zero)
  }
  u32 t = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { u32 x = v[k]; v[k] = t; t += x; }
  u32 inc = t;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += y;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  u32 wsum = lane < kScanBlock / 32 ? s_warp[lane] : 0;
  u32 psum = lane < kScanBlock / 32 ? s_pre[lane] : 0;
  u32 winc = wsum;
#pragma unroll
  for (int d = 1; d < kScanBlock / 32; d <<= 1) {
    u32 y = __shfl_up_sync(0xffffffffu, winc, d);
    if (lane >= d) winc += y;
  }
#pragma unroll
  for (int d = 4; d > 0; d >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, d);
  const u32 prefix = __shfl_sync(0xffffffffu, psum, 0);
  const u32 warp_excl = __shfl_sync(0xffffffffu, winc - wsum, warp);
  if (i0 < n) {
    const u32 b = prefix + warp_excl + inc - t;
    uint4 *o = reinterpret_cast<uint4 *>(out + i0);
    o[0] = make_uint4(v[0] + b, v[1] + b, v[2] + b, v[3] + b);
    o[1] = make_uint4(v[4] + b, v[5] + b, v[6] + b, v[7] + b);
    o[2] = make_uint4(v[8] + b, v[9] + b, v[10] + b, v[11] + b);
    o[3] = make_uint4(v[12] + b, v[13] + b, v[14] + b, v[15] + b);
    if (report.host) {
#pragma unroll
      for (int k = 0; k < kScanItems; k++) {
        volatile u32 *hw = report.host;
        if (i0 + k == report.lo_cell) { hw[0] = v[k] + b; __threadfence_system(); hw[16] = report.stamp; }
        if (i0 + k == report.hi_cell) { hw[1] = v[k] + b; __threadfence_system(); hw[17] = report.stamp; }
        if (i0 + k == report.lo2_cell) { hw[8] = v[k] + b; __threadfence_system(); hw[18] = report.stamp; }
        if (i0 + k == report.hi2_cell) { hw[9] = v[k] + b; __threadfence_system(); hw[19] = report.stamp; }
      }
    }
  }
  if (report.host && tile == 0 && tid == 0) {
    // in lo, in hi, timeout, far, sent lo, sent hi of the last halo exchange
    for (int k = 0; k < 6; k++) report.host[2 + k] = report.halo_ctr ? report.halo_ctr[2 + k] : 0;
    report.host[10] = report.halo_ctr ? report.halo_ctr[12] : 0;  // late
    report.host[11] = report.halo_ctr ? report.halo_ctr[14] : 0;  // sequence number of that exchange
    report.host[12] = report.halo_ctr ? 1u : 0u;
    __threadfence_system();
    *(volatile u32 *)(report.host + 20) = report.stamp;
  }
}

// ---------------------------------------------------------------------------------------
// kernels: binning
// ---------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ int cell_coord(R p, R origin, R inv_cell, int n) {
  int c = (int)floor((p - origin) * inv_cell);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// POS2: position is one packed 2-vector column; otherwise three scalar columns.
// Key and arrival rank of the record at pool index `s`, stored at slot `o` of key/local.
template <typename R, int DIM>
__device__ __forceinline__ void bin_count_one(const void *px, const void *py, const void *pz, size_t s, size_t o,
                                              const GridParams &g, u32 *key, u32 *local, u32 *cell_count,
                                              const u32 *ids) {
  if (ids && ids[s] == ABL_SENTINEL_ID) {
    // padding record of a halo message: parked in the trash cell behind all real cells
    key[o] = g.n_local;
    // (one update per group of converged lanes: the padding records sit next to each other)
    const unsigned grp = __activemask();
    if ((threadIdx.x & 31u) == (unsigned)__ffs(grp) - 1u) atomicAdd(&cell_count[g.n_local], (u32)__popc(grp));
    return;
  }
  R x, y, z = 0;
  if (DIM == 2) {
    x = ((const R *)px)[2 * s];
    y = ((const R *)px)[2 * s + 1];
  } else {
    x = ((const R *)px)[s];
    y = ((const R *)py)[s];
    z = ((const R *)pz)[s];
  }
  int cx = cell_coord<R>(x, (R)g.origin[0], (R)g.inv_cell, g.n_cell[0]);
  int cy = cell_coord<R>(y, (R)g.origin[1], (R)g.inv_cell, g.n_cell[1]);
  if (DIM == 2) cy = min(max(cy, g.axis_lo), g.axis_hi - 1);
  u32 c = (u32)cy * (u32)g.n_cell[0] + (u32)cx;
  if (DIM == 3) {
    int cz = cell_coord<R>(z, (R)g.origin[2], (R)g.inv_cell, g.n_cell[2]);
    cz = min(max(cz, g.axis_lo), g.axis_hi - 1);
    c += (u32)cz * (u32)g.n_cell[0] * (u32)g.n_cell[1];
  }
  c -= g.key_base;
  key[o] = c;
  atomicAdd(&cell_count[c], 1u);   // result unused: a RED; slots are handed out by k_bin_scatter
}

template <typename R, int DIM>
__global__ void k_bin_count(const void *px, const void *py, const void *pz, u32 n, u32 src_begin,
                            u32 out_begin, GridParams g, u32 *key, u32 *local, u32 *cell_count,
                            const u32 *ids) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bin_count_one<R, DIM>(px, py, pz, (size_t)src_begin + i, (size_t)out_begin + i, g, key, local, cell_count, ids);
}

// seg_ids[slot] = id of the agent that drew slot `local` of its cell segment.  The slots are drawn
// here, by counting the histogram back down (the scan leaves it in place): every agent takes one,
// so the histogram is all zero again when the kernel ends — ready for the next fused histogram
// without a clearing pass — and the atomic round trip is paid by this short, fully occupied kernel
// instead of by the last instruction of the step kernel.
// Slab decomposition: the first thread also reports the owned range (ScanReport, see k_tile_scan) — from here
// rather than from the scan, whose few short blocks would all wait for the two round trips to host memory of
// the reporting threads (device timeline on 2 B200: scan 9 -> 16 us); this kernel runs longer than the report.
__global__ void k_bin_scatter(const u32 *key, u32 *local, const u32 *ids, u32 n, u32 src_begin,
                              const u32 *cell_start, u32 *seg_ids, u32 *cell_count, ScanReport report, u32 *seg_begin) {
  if (c_pdl_trigger) cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (report.host && t == 0) {
    volatile u32 *hw = report.host;
    hw[0] = cell_start[report.lo_cell];
    hw[1] = cell_start[report.hi_cell];
    for (int k = 0; k < 6; k++) hw[2 + k] = report.halo_ctr ? report.halo_ctr[2 + k] : 0;
    hw[8] = cell_start[report.lo2_cell];
    hw[9] = cell_start[report.hi2_cell];
    hw[10] = report.halo_ctr ? report.halo_ctr[12] : 0;  // late
    hw[11] = report.halo_ctr ? report.halo_ctr[14] : 0;  // sequence number of that exchange
    hw[12] = report.halo_ctr ? 1u : 0u;
    __threadfence_system();
    for (int k = 16; k <= 20; k++) hw[k] = report.stamp;
  }
  if (t >= n) return;
  const u32 c = key[t];
  // (one atomic per agent.  Aggregating the lanes of a warp that share a cell with match.any was measured: the
  // instruction iterates over the distinct keys of the warp — nearly 32 here — and the kernel went from 11 to
  // 20 us at 1 M agents.)
  const u32 l = atomicSub(&cell_count[c], 1u) - 1u;
  const u32 b = cell_start[c];
  const u32 id = ids[src_begin + t];
  seg_ids[b + l] = id;
  if (c_bin_seginfo) {
    // for k_bin_rank_move: where the agent's cell segment begins and how many ids it holds (padding records of a
    // halo exchange keep their slot instead: any distinct slot will do for them)
    seg_begin[t] = b;
    local[t] = id == ABL_SENTINEL_ID ? l : cell_start[c + 1] - b;
  } else {
    local[t] = l;
  }
}

__device__ __forceinline__ void copy_elem(void *dst, size_t di, const void *src, size_t si, int elem) {
  switch (elem) {
    case 1: ((u8 *)dst)[di] = ((const u8 *)src)[si]; break;
    case 4: ((u32 *)dst)[di] = ((const u32 *)src)[si]; break;
    case 8: ((uint2 *)dst)[di] = ((const uint2 *)src)[si]; break;
    default: ((uint4 *)dst)[di] = ((const uint4 *)src)[si]; break;
  }
}

// One thread per *source* agent.  Its final place is segment_begin + (number of ids in its
// cell segment smaller than its own id); the segment is contiguous, so the rank is a short
// scan of seg_ids.  All reads of the agent's record are coalesced; because agents move
// little between two binnings the pool is already almost in cell order and the writes land
// close to the reads (near-coalesced).
// Crowded cells (more than kRankCoop agents: a population gathered in one place, or agents beyond the
// environment's bounds clamped into an edge cell): every agent of such a cell would read the whole
// segment on its own.  Those agents are ranked by their warp instead, one after the other — the 32
// lanes count a strided share of the segment each (coalesced) and add up with one redux: 1/32 of the
// reads and compares, which keeps the rank well below what the neighbour loop over the same cell costs
// the step kernel.
static const u32 kRankCoop = 128;
__global__ void k_bin_rank_move(ColTable t, const u32 *seg_ids, const u32 *key, const u32 *local,
                                const u32 *ids, u32 n, u32 src_begin, const u32 *cell_start, const u32 *seg_begin) {
  if (c_pdl_trigger) cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;            // (no early return: whole warps reach the collectives below)
  const u32 src = src_begin + (valid ? i : 0u);
  u32 mine = ABL_SENTINEL_ID, b = 0, e = 0, slot = 0;
  if (valid) {
    mine = ids[src];
    // the record itself is needed only after three dependent round trips (key -> cell_start -> seg_ids): request
    // its lines now (one lane per 128-byte line and column)
    if (c_bin_prefetch) {
      for (int k = 0; k < t.ncols; k++) {
        const unsigned char *q = (const unsigned char *)t.in[k] + (size_t)src * t.elem[k];
        if ((threadIdx.x & 31u) == 0u || ((size_t)q & 127u) < (size_t)t.elem[k])
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
      }
    }
    if (c_bin_seginfo) {
      b = seg_begin[i];
      const u32 x = local[i];            // size of the segment; the drawn slot for padding records
      e = b + (mine == ABL_SENTINEL_ID ? 0u : x);
      slot = x;
    } else {
      const u32 c = key[i];
      b = cell_start[c];
      e = cell_start[c + 1];
      if (mine == ABL_SENTINEL_ID) slot = local[i];
    }
  }
  u32 rank = 0;
  const bool crowded = valid && mine != ABL_SENTINEL_ID && e - b > kRankCoop;
  if (valid && mine == ABL_SENTINEL_ID) rank = slot;  // padding records: any distinct slot will do
  else if (valid && !crowded) for (u32 q = b; q < e; q++) rank += (seg_ids[q] < mine) ? 1u : 0u;  // ids are unique
  unsigned todo = __ballot_sync(0xffffffffu, crowded);
  const u32 lane = threadIdx.x & 31u;
  while (todo) {
    const int owner = __ffs(todo) - 1;
    todo &= todo - 1u;
    const u32 ob = __shfl_sync(0xffffffffu, b, owner), oe = __shfl_sync(0xffffffffu, e, owner);
    const u32 oid = __shfl_sync(0xffffffffu, mine, owner);
    u32 cnt = 0;
    for (u32 q = ob + lane; q < oe; q += 32u) cnt += (seg_ids[q] < oid) ? 1u : 0u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((int)lane == owner) rank = cnt;
  }
  if (!valid) return;
  const u32 dst = b + rank;
  for (int k = 0; k < t.ncols; k++) copy_elem(t.out[k], dst, t.in[k], src, t.elem[k]);
}

// ---------------------------------------------------------------------------------------
// kernels: compaction (remove) and append (add)
// ---------------------------------------------------------------------------------------
// (`first`: pool index of the record flag 0 belongs to — the first owned record under slab
// decomposition, where the flags are relative to the owned range and survivors move to the
// front of the alternate buffers)
__global__ void k_compact_move(ColTable t, const u8 *dead, const u32 *offsets, u32 n, u32 first) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || dead[i]) return;
  u32 dst = offsets[i];
  for (int k = 0; k < t.ncols; k++) copy_elem(t.out[k], dst, t.in[k], (size_t)first + i, t.elem[k]);
}

// list[rank] = (parent id << 32 | parent slot) for every flagged parent
__global__ void k_collect_adds(const u8 *flag, const u32 *offsets, const u32 *parent_ids, u32 n,
                               u64 *list) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  list[offsets[i]] = ((u64)parent_ids[i] << 32) | i;
}

// New agents are appended in ascending parent-id order (canonical, independent of the
// current memory order of the parents).  m is small compared to n, rank by counting.
__global__ void k_append(ColTable t, const u64 *list, u32 m, u32 base, u32 first_id) {
  u32 a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  u64 mine = list[a];
  u32 rank = 0;
  for (u32 q = 0; q < m; q++) rank += (list[q] < mine) ? 1u : 0u;
  u32 src = (u32)mine;
  u32 dst = base + rank;
  for (int k = 0; k < t.ncols; k++) {
    if (t.host_off[k] < 0) ((u32 *)t.out[k])[dst] = first_id + rank;  // id column
    else copy_elem(t.out[k], dst, t.in[k], src, t.elem[k]);
  }
}

// Many adds in one step (m above kAppendScanThreshold): counting costs m compares per new agent.
// The rank of a parent among the adding parents is then read from an exclusive scan over a
// presence flag per agent id (k_mark_parents + the runtime's scan: O(next_id + m) instead of
// O(m^2)); same order, same ids.
static const u32 kAppendScanThreshold = 2048;
static u32 append_scan_threshold() {   // ABL_CUDA_APPEND_SCAN=<m>: tests force either path
  const char *e = getenv("ABL_CUDA_APPEND_SCAN");
  return e && *e ? (u32)strtoul(e, nullptr, 10) : kAppendScanThreshold;
}
__global__ void k_mark_parents(const u64 *list, u32 m, u32 *present) {
  u32 a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < m) present[(u32)(list[a] >> 32)] = 1u;
}
__global__ void k_append_by_id(ColTable t, const u64 *list, const u32 *rank_of_id, u32 m, u32 base, u32 first_id) {
  u32 a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  const u64 mine = list[a];
  const u32 rank = rank_of_id[(u32)(mine >> 32)];
  const u32 src = (u32)mine;
  const u32 dst = base + rank;
  for (int k = 0; k < t.ncols; k++) {
    if (t.host_off[k] < 0) ((u32 *)t.out[k])[dst] = first_id + rank;  // id column
    else copy_elem(t.out[k], dst, t.in[k], src, t.elem[k]);
  }
}

// Slab decomposition: the list arrives sorted by parent id from the host, together with the
// rank of every parent among the parents of ALL slabs (ids of new agents are global).
__global__ void k_append_ranked(ColTable t, const u64 *list, const u32 *global_rank, u32 m, u32 base, u32 first_id) {
  u32 a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  const u32 src = (u32)list[a];
  const u32 dst = base + a;
  for (int k = 0; k < t.ncols; k++) {
    if (t.host_off[k] < 0) ((u32 *)t.out[k])[dst] = first_id + global_rank[a];  // id column
    else copy_elem(t.out[k], dst, t.in[k], src, t.elem[k]);
  }
}

// ---------------------------------------------------------------------------------------
// kernels: reductions (integer results are exact; float sums use a fixed two-level tree)
// ---------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T block_sum(T v, T *smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  v = threadIdx.x < blockDim.x / 32 ? smem[threadIdx.x] : (T)0;
  if (warp == 0) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  }
  return v;
}

// KIND 0: sum of int column, 1: sum of bool column, 2: count int == value, 3: count bool == value
template <int KIND>
__global__ void k_reduce_int(const void *col, u32 n, int value, int *result) {
  __shared__ int smem[32];
  int acc = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (KIND == 0) acc += ((const int *)col)[i];
    else if (KIND == 1) acc += ((const u8 *)col)[i] ? 1 : 0;
    else if (KIND == 2) acc += ((const int *)col)[i] == value ? 1 : 0;
    else acc += (((const u8 *)col)[i] ? 1 : 0) == value ? 1 : 0;
  }
  acc = block_sum<int>(acc, smem);
  if (threadIdx.x == 0) atomicAdd(result, acc);  // integer addition: order-independent
}

template <typename R>
__global__ void k_reduce_real_partial(const R *col, int stride, int comp, u32 n, double *partial) {
  __shared__ double smem[32];
  double acc = 0;
  // fixed assignment of elements to threads and a fixed tree => deterministic result
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    acc += (double)col[(size_t)i * stride + comp];
  acc = block_sum<double>(acc, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

template <typename R>
__global__ void k_count_real(const R *col, u32 n, R value, int *result) {
  __shared__ int smem[32];
  int acc = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    acc += col[i] == value ? 1 : 0;
  acc = block_sum<int>(acc, smem);
  if (threadIdx.x == 0) atomicAdd(result, acc);
}

__global__ void k_final_sum(const double *partial, int m, double *out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < m; i++) s += partial[i];
    *out = s;
  }
}

// ---------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------
static inline u32 blocks_for(size_t n, int bs) { return (u32)((n + bs - 1) / bs); }

static int ensure_stage(abl_runtime *rt, size_t bytes) {
  if (rt->stage_cap >= bytes) return ABL_OK;
  if (rt->stage) CU(cudaFree(rt->stage));
  rt->stage = nullptr;
  size_t cap = round_up(bytes + bytes / 8, 1 << 20);
  CU(cudaMalloc(&rt->stage, cap));
  rt->stage_cap = cap;
  return ABL_OK;
}

static int ensure_scan(abl_runtime *rt, size_t n) {
  size_t tiles = (n + kScanTile - 1) / kScanTile + 1;
  if (!rt->scan.ctrl) {
    CU(cudaMalloc(&rt->scan.ctrl, 4 * sizeof(u32)));
    u32 init[4] = {0, 0, 1, 0};  // epoch starts at 1 so zeroed descriptors are invalid
    CU(cudaMemcpyAsync(rt->scan.ctrl, init, sizeof init, cudaMemcpyHostToDevice, rt->stream));
    CU(cudaStreamSynchronize(rt->stream));
  }
  if (rt->scan.max_tiles >= tiles) return ABL_OK;
  CU(cudaStreamSynchronize(rt->stream));
  if (rt->scan.desc) CU(cudaFree(rt->scan.desc));
  if (rt->scan.tile_sum) CU(cudaFree(rt->scan.tile_sum));
  size_t cap = tiles * 2;
  CU(cudaMalloc(&rt->scan.desc, cap * sizeof(u64)));
  CU(cudaMalloc(&rt->scan.tile_sum, cap * sizeof(u32)));
  CU(cudaMemsetAsync(rt->scan.desc, 0, cap * sizeof(u64), rt->stream));
  rt->scan.max_tiles = cap;
  return ABL_OK;
}

// Exclusive scan of `n` items; arrays must be padded to a multiple of kScanTile.
template <typename T, int MODE, bool ZERO>
static int run_scan(abl_runtime *rt, T *in, u32 *out, size_t n, u32 *total_out) {
  if (n == 0) {
    if (total_out) CU(cudaMemsetAsync(total_out, 0, sizeof(u32), rt->stream));
    return ABL_OK;
  }
  TRY(ensure_scan(rt, n));
  u32 tiles = (u32)((n + kScanTile - 1) / kScanTile);
  k_scan<T, MODE, ZERO><<<tiles, kScanBlock, 0, rt->stream>>>(in, out, (u32)n, rt->scan.desc,
                                                            rt->scan.ctrl, total_out);
  rt->launches++;
  CU(cudaGetLastError());
  return ABL_OK;
}

static int release_device(abl_runtime *rt, void *q) {
  if (!q) return ABL_OK;
  if (rt && rt->defer_free) rt->garbage.push_back(q);
  else CU(cudaFree(q));
  return ABL_OK;
}

static int collect_garbage(abl_runtime *rt) {
  for (void *q : rt->garbage) CU(cudaFree(q));
  rt->garbage.clear();
  return ABL_OK;
}

// ABL_CUDA_TRACE: device-side timeline.  A one-thread kernel between two stages adds the time
// since the previous stamp (globaltimer, so idle gaps count as well) to the bucket of the stage
// that has just ended.
__global__ void k_trace_stamp(unsigned long long *state, int bucket) {
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (state[0]) { state[1 + bucket] += now - state[0]; state[9 + bucket] += 1; }
  state[0] = now;
}
enum { TR_SCAN = 0, TR_MOVE = 1, TR_STEP = 2, TR_EXCHANGE = 3, TR_OTHER = 4 };
static void trace_stamp(abl_runtime *rt, int bucket) {
  if (!rt->trace || !rt->trace_state) return;
  k_trace_stamp<<<1, 1, 0, rt->stream>>>(rt->trace_state, bucket);
}

// exclusive scan of the cell histogram into cell_start (and zeroing of the histogram)
// `report` (optional): see ScanReport; *reported tells the caller whether the scan took care of it
static int run_cell_scan(abl_runtime *rt, u32 *count, u32 *start, size_t n, const ScanReport *report = nullptr,
                         bool *reported = nullptr) {
  if (reported) *reported = false;
  if (!rt->scan_two_pass) return run_scan<u32, 0, false>(rt, count, start, n, nullptr);
  TRY(ensure_scan(rt, n));
  const u32 tiles = (u32)((n + kScanTile - 1) / kScanTile);
  u32 *tile_sum = rt->scan.tile_sum;
  CU(launch_pdl(rt->pdl, k_tile_sum, dim3(tiles), dim3(kScanBlock), 0, rt->stream, (const u32 *)count, (u32)n, tile_sum));
  ScanReport rep;
  memset(&rep, 0, sizeof rep);
  if (report) { rep = *report; if (reported) *reported = true; }
  CU(launch_pdl(rt->pdl, k_tile_scan, dim3(tiles), dim3(kScanBlock), 0, rt->stream, count, start, (u32)n, (const u32 *)tile_sum, rep));
  rt->launches += 2;
  CU(cudaGetLastError());
  return ABL_OK;
}

static int free_pool_scratch(abl_runtime *rt, Pool &p) {
  void *ptrs[] = {p.key, p.local, p.pairs, p.dead, p.add_flag, p.offsets};
  for (void *q : ptrs) TRY(release_device(rt, q));
  p.key = p.local = nullptr; p.pairs = nullptr; p.dead = p.add_flag = nullptr; p.offsets = nullptr;
  p.scratch_cap = 0;
  return ABL_OK;
}

// Grows a pool (columns + scratch) to hold at least `want` agents, preserving contents.
static int reserve_pool(abl_runtime *rt, Pool &p, size_t want) {
  if (want <= p.cap && p.scratch_cap >= p.cap && p.cap > 0) return ABL_OK;
  size_t cap = p.cap;
  if (want > cap) {
    cap = round_up(std::max(want + want / 4, (size_t)kScanTile), kScanTile);
    CU(cudaStreamSynchronize(rt->stream));
    for (Column &c : p.cols) {
      for (int b = 0; b < 2; b++) {
        void *nb = nullptr;
        CU(cudaMalloc(&nb, cap * (size_t)c.elem));
        if (c.buf[b]) {
          if (b == c.cur && p.n)
            CU(cudaMemcpy(nb, c.buf[b], p.n * (size_t)c.elem, cudaMemcpyDeviceToDevice));
          TRY(release_device(rt, c.buf[b]));
        }
        c.buf[b] = nb;
      }
    }
    p.cap = cap;
  }
  if (p.scratch_cap < p.cap) {
    // dead / add flags may hold live data (set by a step kernel that has not been committed)
    u8 *dead = nullptr, *add_flag = nullptr;
    CU(cudaMalloc(&dead, p.cap));
    CU(cudaMalloc(&add_flag, p.cap));
    CU(cudaMemset(dead, 0, p.cap));
    CU(cudaMemset(add_flag, 0, p.cap));
    if (p.scratch_cap) {
      CU(cudaMemcpy(dead, p.dead, p.scratch_cap, cudaMemcpyDeviceToDevice));
      CU(cudaMemcpy(add_flag, p.add_flag, p.scratch_cap, cudaMemcpyDeviceToDevice));
    }
    TRY(free_pool_scratch(rt, p));
    p.dead = dead;
    p.add_flag = add_flag;
    CU(cudaMalloc(&p.key, p.cap * sizeof(u32)));
    CU(cudaMalloc(&p.local, p.cap * sizeof(u32)));
    CU(cudaMalloc(&p.pairs, p.cap * sizeof(u64)));
    CU(cudaMalloc(&p.offsets, (p.cap + kScanTile) * sizeof(u32)));
    p.scratch_cap = p.cap;
  }
  return ABL_OK;
}

static void fill_table(const Pool &p, ColTable &t, bool out_is_alt) {
  t.ncols = (int)p.cols.size();
  for (int c = 0; c < t.ncols; c++) {
    const Column &col = p.cols[c];
    t.elem[c] = col.elem; t.comp[c] = col.comp; t.ncomp[c] = col.ncomp;
    t.host_off[c] = col.host_off;
    t.in[c] = col.buf[col.cur];
    t.out[c] = out_is_alt ? col.buf[col.cur ^ 1] : col.buf[col.cur];
  }
}

// A fused histogram that will not be consumed by the next binning (the pool is about to be
// reordered or refilled) has to be wiped, otherwise the next histogram adds on top of it.
static int drop_fused_histogram(abl_runtime *rt, Pool &p) {
  if (p.counted && p.cell_count)
    CU(cudaMemsetAsync(p.cell_count, 0, ((size_t)rt->grid.n_local + 2) * sizeof(u32), rt->stream));
  p.counted = false;
  return ABL_OK;
}

static void flip_all(Pool &p) { for (Column &c : p.cols) c.cur ^= 1; }

// ---------------------------------------------------------------------------------------
// C ABI: life cycle
// ---------------------------------------------------------------------------------------
extern "C" int abl_cuda_abi_version(void) { return ABL_CUDA_ABI_VERSION; }

extern "C" const char *abl_cuda_last_error(void) { return g_err; }

extern "C" void abl_cuda_default_config(abl_config *cfg) {
  memset(cfg, 0, sizeof *cfg);
  cfg->device = -1;
  cfg->use_float = 0;
  cfg->seed = 0x0123456789abcdefull;
  cfg->deterministic = 1;
  cfg->tile_neighbours = 0;
  cfg->block_size = 0;
}

extern "C" int abl_cuda_create(abl_runtime **out, const abl_config *cfg) {
  if (!out) return fail(ABL_ERR_ARGUMENT, "abl_cuda_create: null output handle");
  *out = nullptr;
  abl_config c;
  if (cfg) c = *cfg; else abl_cuda_default_config(&c);
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (count == 0) return fail(ABL_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
  if (c.device >= 0) CU(cudaSetDevice(c.device));
  abl_runtime *rt = new abl_runtime;
  rt->cfg = c;
  if (rt->cfg.block_size < 0) rt->cfg.block_size = 0;
  CU(cudaGetDevice(&rt->device));
  rt->real_size = c.use_float ? 4 : 8;
  CU(cudaStreamCreateWithFlags(&rt->stream, cudaStreamNonBlocking));
  CU(cudaMalloc(&rt->d_scalar, 4096));
  CU(cudaMemset(rt->d_scalar, 0, 4096));
  CU(cudaHostAlloc(&rt->h_scalar, 4096, cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer(&rt->h_scalar_dev, rt->h_scalar, 0));
  for (int i = 0; i < 4; i++) CU(cudaEventCreate(&rt->ev[i]));
  for (int i = 0; i < 2; i++) CU(cudaEventCreate(&rt->ev_ts[i]));
  CU(cudaEventCreateWithFlags(&rt->ev_own, cudaEventDisableTiming));
  memset(&rt->grid, 0, sizeof rt->grid);
  if (const char *ms = getenv("ABL_CUDA_HALO_TIMEOUT_MS")) rt->halo_timeout_ns = atoll(ms) * 1000000ll;
  if (const char *dr = getenv("ABL_CUDA_DEVICE_RANGE")) rt->device_range = atoi(dr) != 0;
  if (const char *ov = getenv("ABL_CUDA_HALO_OVERLAP")) rt->halo_overlap = atoi(ov) != 0;
  if (const char *as = getenv("ABL_CUDA_HALO_ASYNC")) rt->halo_async = atoi(as) != 0;
  if (const char *pd = getenv("ABL_CUDA_PDL")) rt->pdl = atoi(pd) != 0;
  if (const char *pt = getenv("ABL_CUDA_PDL_TRIGGER")) rt->pdl_trigger = atoi(pt) != 0;
  if (const char *mh = getenv("ABL_CUDA_MBAR_HINT")) rt->mbar_hint = atoi(mh) != 0;
  if (const char *rs = getenv("ABL_CUDA_REPORT_IN_SCAN")) rt->report_in_scan = atoi(rs) != 0;
  {
    const int trig = rt->pdl && rt->pdl_trigger ? 1 : 0;
    CU(cudaMemcpyToSymbol(c_pdl_trigger, &trig, sizeof trig));
    const char *bp = getenv("ABL_CUDA_BIN_PREFETCH");
    const int pre = bp ? (atoi(bp) != 0 ? 1 : 0) : 1;
    CU(cudaMemcpyToSymbol(c_bin_prefetch, &pre, sizeof pre));
    const char *bsi = getenv("ABL_CUDA_BIN_SEGINFO");
    const int seginfo = bsi ? (atoi(bsi) != 0 ? 1 : 0) : 0;
    CU(cudaMemcpyToSymbol(c_bin_seginfo, &seginfo, sizeof seginfo));
  }
  if (const char *fl = getenv("ABL_CUDA_FLAT")) rt->flat_loop = atoi(fl) != 0 ? 1 : 0;
  if (const char *tn = getenv("ABL_CUDA_TUNE")) { if (atoi(tn) != 0) rt->flat_loop = -1; }
  if (const char *dn = getenv("ABL_CUDA_DENSE")) rt->dense_tile = atoi(dn) != 0 ? 1 : 0;
  if (const char *sp = getenv("ABL_CUDA_SPLIT")) rt->split_prefilter = atoi(sp) != 0 ? 1 : 0;
  if (const char *bk = getenv("ABL_CUDA_BULK")) rt->bulk_tile = atoi(bk) < 0 ? 0 : atoi(bk) > 2 ? 2 : atoi(bk);
  if (const char *nlv = getenv("ABL_CUDA_NLIST")) rt->nlist = atoi(nlv) != 0;
  if (const char *mb = getenv("ABL_CUDA_NLIST_MB")) rt->nlist_budget = (size_t)std::max(1, atoi(mb)) << 20;
  if (const char *sc = getenv("ABL_CUDA_SCAN")) rt->scan_two_pass = strcmp(sc, "lookback") != 0;
  if (getenv("ABL_CUDA_TRACE")) {
    rt->trace = true;
    for (int i = 0; i < 5; i++) CU(cudaEventCreate(&rt->xev[i]));
    CU(cudaMalloc(&rt->trace_state, 17 * sizeof(unsigned long long)));
    CU(cudaMemset(rt->trace_state, 0, 17 * sizeof(unsigned long long)));
  }
  *out = rt;
  return ABL_OK;
}

extern "C" int abl_cuda_destroy(abl_runtime *rt) {
  if (!rt) return ABL_OK;
  cudaSetDevice(rt->device);
  cudaStreamSynchronize(rt->stream);
  if (rt->trace && rt->xcount) {
    const char *names[4] = {"pack", "nccl", "headers+sync", "unpack"};
    char line[512];
    int off = snprintf(line, sizeof line, "abl_cuda[rank %d] exchange trace over %lu calls (us per call):", rt->rank, rt->xcount);
    for (int i = 0; i < 4; i++)
      off += snprintf(line + off, sizeof line - off, " %s dev %.1f host %.1f;", names[i], 1e3 * rt->xdev[i] / rt->xcount, 1e6 * rt->xhost[i] / rt->xcount);
    fprintf(stderr, "%s\n", line);
  }
  collect_garbage(rt);
  if (rt->trace && rt->trace_state) {
    unsigned long long st[17];
    if (cudaMemcpy(st, rt->trace_state, sizeof st, cudaMemcpyDeviceToHost) == cudaSuccess && st[9 + TR_STEP]) {
      const char *names[5] = {"histogram+scan", "scatter+rank/move", "step kernel", "exchange", "other"};
      char line[512];
      int off = snprintf(line, sizeof line, "abl_cuda[slab %d] device timeline, us per occurrence (idle gaps included):", rt->my_slab);
      for (int b = 0; b < 5; b++)
        if (st[9 + b]) off += snprintf(line + off, sizeof line - off, " %s %.1f;", names[b], 1e-3 * (double)st[1 + b] / (double)st[9 + b]);
      fprintf(stderr, "%s\n", line);
    }
    cudaFree(rt->trace_state);
  }
  if (rt->trace && rt->own_waits)
    fprintf(stderr, "abl_cuda[slab %d] host blocked %.2f us per binning waiting for the owned range (%lu binnings)\n",
            rt->my_slab, 1e6 * rt->own_wait_s / rt->own_waits, rt->own_waits);
  if (rt->trace) {
    for (Pool &p : rt->pools) {
      if (!p.halo_ctr) continue;
      u32 w[16];
      if (cudaMemcpy(w, p.halo_ctr, sizeof w, cudaMemcpyDeviceToHost) != cudaSuccess || !w[10]) continue;
      unsigned long long ns;
      memcpy(&ns, w + 8, sizeof ns);
      fprintf(stderr, "abl_cuda[slab %d] pool %s: %u direct exchanges, publish + wait for neighbours %.2f us per exchange\n",
              rt->my_slab, p.name.c_str(), w[10], 1e-3 * (double)ns / w[10]);
    }
  }
  for (Pool &p : rt->pools) {
    for (int d = 0; d < 2; d++) if (p.halo_peer[d] && p.halo_ipc[d]) cudaIpcCloseMemHandle(p.halo_peer[d]);
    if (p.halo_recv) cudaFree(p.halo_recv);
    if (p.halo_ctr) cudaFree(p.halo_ctr);
    for (Column &c : p.cols) for (int b = 0; b < 2; b++) if (c.buf[b]) cudaFree(c.buf[b]);
    free_pool_scratch(nullptr, p);
    if (p.cell_count) cudaFree(p.cell_count);
    if (p.cell_start) cudaFree(p.cell_start);
    if (p.shadow) cudaFree(p.shadow);
    if (p.shadow_max) cudaFree(p.shadow_max);
  }
  for (Step &st : rt->steps) {
    if (st.pf_buf) cudaFree(st.pf_buf);
    if (st.nl.cnt) cudaFree(st.nl.cnt);
    if (st.nl.idx) cudaFree(st.nl.idx);
  }
  for (auto &pr : rt->pinned_ranges) cudaHostUnregister(pr.first);
  rt->pinned_ranges.clear();
  if (rt->scan.desc) cudaFree(rt->scan.desc);
  if (rt->scan.tile_sum) cudaFree(rt->scan.tile_sum);
  if (rt->scan.ctrl) cudaFree(rt->scan.ctrl);
  if (rt->stage) cudaFree(rt->stage);
  if (rt->rank_buf) cudaFree(rt->rank_buf);
  if (rt->pinned) cudaFreeHost(rt->pinned);
  if (rt->d_scalar) cudaFree(rt->d_scalar);
  if (rt->h_scalar) cudaFreeHost(rt->h_scalar);
  for (int i = 0; i < 4; i++) if (rt->ev[i]) cudaEventDestroy(rt->ev[i]);
  for (cudaEvent_t e : rt->timing_log) cudaEventDestroy(e);
  rt->timing_log.clear();
  for (int i = 0; i < 2; i++) if (rt->ev_ts[i]) cudaEventDestroy(rt->ev_ts[i]);
  if (rt->ev_own) cudaEventDestroy(rt->ev_own);
  if (rt->ev_fork) cudaEventDestroy(rt->ev_fork);
  if (rt->ev_join) cudaEventDestroy(rt->ev_join);
  if (rt->side_stream) cudaStreamDestroy(rt->side_stream);
  cudaStreamDestroy(rt->stream);
  delete rt;
  return ABL_OK;
}

extern "C" void *abl_cuda_stream(abl_runtime *rt) { return rt ? (void *)rt->stream : nullptr; }

extern "C" int abl_cuda_synchronize(abl_runtime *rt) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  CU(cudaStreamSynchronize(rt->stream));
  TRY(collect_garbage(rt));
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// C ABI: environment and pools
// ---------------------------------------------------------------------------------------
extern "C" int abl_cuda_set_environment(abl_runtime *rt, int dim, const double *env_min,
                                        const double *env_max, double granularity) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (dim != 2 && dim != 3) return fail(ABL_ERR_ARGUMENT, "environment dimension must be 2 or 3");
  if (!(granularity > 0)) return fail(ABL_ERR_ARGUMENT, "granularity must be positive");
  GridParams &g = rt->grid;
  g.dim = dim;
  // Cells are a shade larger than the granularity (ABL_CELL_PAD_*, abl_cuda.h): the reference's
  // filter `sqrtf((float)s) > R` accepts true distances up to R(1 + 6e-8), and the cell index is a
  // floor of a rounded product, so with cell == R two agents the reference pairs up could land
  // two cells apart (p = 9.999999999999998, q = p + 5.0, cell 5.0 -> cells 1 and 3) and the 3^d
  // search would miss the pair.  With the padding |dx| <= R(1 + 6e-8) implies |d index| <= 1 with a
  // margin far above the rounding of the index computation.
  g.cell = granularity * (1.0 + (rt->real_size == 8 ? ABL_CELL_PAD_F64 : ABL_CELL_PAD_F32));
  g.inv_cell = rt->real_size == 8 ? 1.0 / g.cell : (double)(1.0f / (float)g.cell);
  u64 cells = 1;
  for (int a = 0; a < 3; a++) {
    if (a < dim) {
      double size = env_max[a] - env_min[a];
      if (size < 0) return fail(ABL_ERR_ARGUMENT, "environment max < min");
      long nc = (long)ceil(size / g.cell);
      if (nc < 1) nc = 1;
      g.n_cell[a] = (int)nc;
      g.origin[a] = env_min[a];
    } else {
      g.n_cell[a] = 1;
      g.origin[a] = 0;
    }
    cells *= (u64)g.n_cell[a];
  }
  if (cells > 0x7fffffffull)
    return fail(ABL_ERR_ARGUMENT, "grid of %llu cells exceeds 2^31; increase the granularity", cells);
  g.n_cells = (u32)cells;
  g.axis_lo = 0;
  g.axis_hi = dim == 3 ? g.n_cell[2] : g.n_cell[1];
  g.key_base = 0;
  g.n_local = g.n_cells;
  rt->env_set = true;
  for (Pool &p : rt->pools) p.binned = false;
  return ABL_OK;
}

extern "C" int abl_cuda_grid_cells(abl_runtime *rt, unsigned *n_cells, int n_cell_axis[3]) {
  if (!rt || !rt->env_set) return fail(ABL_ERR_STATE, "environment not set");
  if (n_cells) *n_cells = rt->grid.n_cells;
  if (n_cell_axis) for (int a = 0; a < 3; a++) n_cell_axis[a] = rt->grid.n_cell[a];
  return ABL_OK;
}

extern "C" int abl_cuda_add_pool(abl_runtime *rt, const abl_agent_desc *desc, int *pool) {
  if (!rt || !desc) return fail(ABL_ERR_ARGUMENT, "null argument");
  if (desc->n_members > ABL_MAX_MEMBERS) return fail(ABL_ERR_ARGUMENT, "too many members");
  Pool p;
  p.name = desc->name ? desc->name : "";
  p.stride = desc->stride;
  const int rs = rt->real_size;
  for (int m = 0; m < desc->n_members; m++) {
    const abl_member_desc &md = desc->members[m];
    Member mem;
    mem.type = md.type;
    mem.name = md.name ? md.name : "";
    mem.first_col = (int)p.cols.size();
    auto add_col = [&](int comp, int ncomp, int off) {
      Column c;
      c.comp = comp; c.ncomp = ncomp; c.elem = comp * ncomp; c.host_off = off;
      p.cols.push_back(c);
    };
    switch (md.type) {
      case ABL_TYPE_BOOL: add_col(1, 1, (int)md.offset); break;
      case ABL_TYPE_INT: add_col(4, 1, (int)md.offset); break;
      case ABL_TYPE_FLOAT: add_col(rs, 1, (int)md.offset); break;
      case ABL_TYPE_FLOAT2: add_col(rs, 2, (int)md.offset); break;
      case ABL_TYPE_FLOAT3:
        for (int k = 0; k < 3; k++) add_col(rs, 1, (int)md.offset + k * rs);
        break;
      default: return fail(ABL_ERR_ARGUMENT, "member %s: unsupported type %d", mem.name.c_str(), md.type);
    }
    mem.ncols = (int)p.cols.size() - mem.first_col;
    if (md.is_pos) p.pos_member = m;
    p.members.push_back(mem);
  }
  if ((int)p.cols.size() > ABL_MAX_COLUMNS) return fail(ABL_ERR_ARGUMENT, "too many columns");
  Column idc;
  idc.comp = 4; idc.ncomp = 1; idc.elem = 4; idc.host_off = -1;
  p.id_col = (int)p.cols.size();
  p.cols.push_back(idc);
  p.index = (int)rt->pools.size();
  rt->pools.push_back(p);
  if (pool) *pool = (int)rt->pools.size() - 1;
  return ABL_OK;
}

static int get_pool(abl_runtime *rt, int pool, Pool **out) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (pool < 0 || pool >= (int)rt->pools.size()) return fail(ABL_ERR_ARGUMENT, "bad pool index %d", pool);
  *out = &rt->pools[pool];
  return ABL_OK;
}

static int slab_bin_if_needed(abl_runtime *rt, Pool &p);
static int bin_pool(abl_runtime *rt, Pool &p, bool defer_report = false);
static int slab_settle(abl_runtime *rt, Pool &p);
static bool halo_direct(const Pool &p);
static int halo_reserve(abl_runtime *rt, Pool &p);
static int halo_fill_view(abl_runtime *rt, Pool &p, abl_slab_view &v, bool step_kernel, bool dev_range);
static int halo_finish(abl_runtime *rt, Pool &p, bool published, bool dev_range, bool forked = false);
extern "C" int abl_cuda_pool_size(abl_runtime *rt, int pool, size_t *n);

extern "C" int abl_cuda_pool_size(abl_runtime *rt, int pool, size_t *n) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  if (rt->slab && p->pos_member >= 0) {
    TRY(slab_bin_if_needed(rt, *p));
    if (n) *n = p->own_end - p->own_begin;
    return ABL_OK;
  }
  if (n) *n = p->n;
  return ABL_OK;
}

extern "C" int abl_cuda_unpin_host(abl_runtime *rt, void *ptr) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  for (size_t i = 0; i < rt->pinned_ranges.size(); i++) {
    if (rt->pinned_ranges[i].first == ptr) {
      CU(cudaStreamSynchronize(rt->stream));
      cudaHostUnregister(ptr);
      rt->pinned_ranges.erase(rt->pinned_ranges.begin() + i);
      break;
    }
  }
  return ABL_OK;
}

extern "C" int abl_cuda_pin_host(abl_runtime *rt, void *ptr, size_t bytes) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (!ptr || bytes < (1u << 16)) return ABL_OK;  // not worth it
  for (auto &pr : rt->pinned_ranges) {
    if (pr.first == ptr) {
      if (pr.second >= bytes) return ABL_OK;
      TRY(abl_cuda_unpin_host(rt, ptr));
      break;
    }
  }
  // best effort: when the pages cannot be locked the transfers simply stay pageable
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaSuccess) {
    rt->pinned_ranges.push_back({ptr, bytes});
  } else {
    cudaGetLastError();
    if (getenv("ABL_CUDA_VERBOSE"))
      fprintf(stderr, "abl_cuda: cudaHostRegister(%zu bytes) failed: %s (transfers stay pageable)\n", bytes, cudaGetErrorString(e));
  }
  return ABL_OK;
}

__global__ void k_set_ids(u32 *dst, const u32 *src, u32 n) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

static int upload_impl(abl_runtime *rt, int pool, const void *host_aos, const unsigned *ids, size_t n,
                       unsigned next_id) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  CU(cudaSetDevice(rt->device));
  if (n > 0x7fffffffu) return fail(ABL_ERR_ARGUMENT, "pool too large");
  p->n = 0;
  // with the direct halo transport arrivals are appended behind the owned records: leave room
  TRY(reserve_pool(rt, *p, std::max(n, (size_t)1) + (halo_direct(*p) ? 2 * p->halo_cap + 1024 : 0)));
  size_t bytes = n * (size_t)p->stride;
  if (n) {
    TRY(ensure_stage(rt, round_up(bytes, 256) + (ids ? n * sizeof(u32) : 0)));
    CU(cudaMemcpyAsync(rt->stage, host_aos, bytes, cudaMemcpyHostToDevice, rt->stream));
    ColTable t;
    fill_table(*p, t, false);
    k_aos_to_soa<<<blocks_for(n, 256), 256, 0, rt->stream>>>(t, (const u8 *)rt->stage, p->stride,
                                                             (u32)n, 0u);
    rt->launches++;
    if (ids) {
      u32 *d_ids = (u32 *)((u8 *)rt->stage + round_up(bytes, 256));
      CU(cudaMemcpyAsync(d_ids, ids, n * sizeof(u32), cudaMemcpyHostToDevice, rt->stream));
      k_set_ids<<<blocks_for(n, 256), 256, 0, rt->stream>>>((u32 *)t.out[p->id_col], d_ids, (u32)n);
      rt->launches++;
    }
    CU(cudaGetLastError());
  }
  p->n = n;
  p->order_serial++;
  p->next_id = ids ? next_id : (u32)n;
  p->binned = false;
  p->own_valid = false;
  p->report_pending = false;
  p->exch_deferred = false;
  p->halo_pending = false;
  p->halo_sync_check = false;
  p->src_begin = 0;
  p->own_begin = 0;
  p->own_end = (u32)n;
  TRY(drop_fused_histogram(rt, *p));
  p->ever_removed = false;
  CU(cudaStreamSynchronize(rt->stream));  // host buffer may be reused by the caller
  return ABL_OK;
}

static int slab_crop_to_owned(abl_runtime *rt, int pool);

// In slab mode every rank may upload the whole population: it keeps the agents of its own
// slab and obtains ghosts with the first exchange.
extern "C" int abl_cuda_upload(abl_runtime *rt, int pool, const void *host_aos, size_t n) {
  TRY(upload_impl(rt, pool, host_aos, nullptr, n, 0));
  if (rt->slab) TRY(slab_crop_to_owned(rt, pool));
  return ABL_OK;
}

extern "C" int abl_cuda_upload_with_ids(abl_runtime *rt, int pool, const void *host_aos,
                                        const unsigned *ids, size_t n, unsigned next_id) {
  if (n && !ids) return fail(ABL_ERR_ARGUMENT, "null ids");
  return upload_impl(rt, pool, host_aos, ids, n, next_id);
}

// ---------------------------------------------------------------------------------------
// scalable upload under slab decomposition: every slab uploads 1/N of the population BY INDEX,
// the records are routed to their owners on the devices (one all-to-all), never through a
// host that holds N copies.  Transit record: the host AoS record followed by the agent id,
// padded to a multiple of 8 bytes (the stride being a multiple of 4).
// ---------------------------------------------------------------------------------------
static int slab_layers(const abl_runtime *rt);
struct SlabBounds { int n; int b[ABL_MAX_SLABS + 1]; };
// bytes of a transit record: host record + id, padded so that 8-byte members stay aligned
static __host__ __device__ inline u32 transit_bytes(u32 stride) { return (stride + 4u + 7u) & ~7u; }

template <typename R>
__global__ void k_route_classify(const u8 *aos, u32 stride, u32 pos_offset, u32 n, R origin, R inv_cell, int n_layers,
                                 SlabBounds sb, u32 *owner, u32 *slot, u32 *counts) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const R p = *reinterpret_cast<const R *>(aos + (size_t)i * stride + pos_offset);
  const int layer = cell_coord<R>(p, origin, inv_cell, n_layers);
  int o = 0;
  for (int s = 1; s < sb.n; s++) o += layer >= sb.b[s] ? 1 : 0;
  owner[i] = (u32)o;
  slot[i] = atomicAdd(&counts[o], 1u);
}

__global__ void k_route_pack(const u8 *aos, u32 stride, u32 n, u32 first_id, const u32 *owner, const u32 *slot,
                             const u32 *offsets, u8 *out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 words = stride / 4u;
  const u32 *src = reinterpret_cast<const u32 *>(aos + (size_t)i * stride);
  u32 *dst = reinterpret_cast<u32 *>(out + (size_t)(offsets[owner[i]] + slot[i]) * transit_bytes(stride));
  for (u32 w = 0; w < words; w++) dst[w] = src[w];
  dst[words] = first_id + i;
}

__global__ void k_transit_to_soa(ColTable t, const u8 *transit, u32 stride, u32 n) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u8 *rec = transit + (size_t)i * transit_bytes(stride);
  for (int c = 0; c < t.ncols; c++) {
    u8 *dst = (u8 *)t.out[c] + (size_t)i * t.elem[c];
    if (t.host_off[c] < 0) { *(u32 *)dst = *reinterpret_cast<const u32 *>(rec + stride); continue; }   // the id column
    for (int k = 0; k < t.ncomp[c]; k++)
      copy_scalar(dst + k * t.comp[c], rec + t.host_off[c] + k * t.comp[c], t.comp[c]);
  }
}

extern "C" int abl_cuda_transit_record_bytes(abl_runtime *rt, int pool, size_t *bytes) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  if (p->stride % 4u) return fail(ABL_ERR_STATE, "pool %s: record size %u is not a multiple of 4", p->name.c_str(), p->stride);
  if (bytes) *bytes = (size_t)transit_bytes(p->stride);
  return ABL_OK;
}

// host records [first_id, first_id + n) of the population -> `dev_out` (device memory of this
// runtime's GPU, room for n transit records): grouped by owning slab, counts[s] records for slab s.
extern "C" int abl_cuda_partition_upload(abl_runtime *rt, int pool, const void *host_aos, size_t n, unsigned first_id,
                                         void *dev_out, unsigned *counts) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (!rt->slab) return fail(ABL_ERR_STATE, "partition_upload requires abl_cuda_set_slab");
  if (p.pos_member < 0) return fail(ABL_ERR_ARGUMENT, "pool %s has no position member", p.name.c_str());
  if (p.stride % 4u) return fail(ABL_ERR_STATE, "pool %s: record size %u is not a multiple of 4", p.name.c_str(), p.stride);
  if (n > 0x7fffffffu) return fail(ABL_ERR_ARGUMENT, "chunk too large");
  if (!counts || (n && (!host_aos || !dev_out))) return fail(ABL_ERR_ARGUMENT, "null argument");
  CU(cudaSetDevice(rt->device));
  const int N = rt->n_slabs;
  for (int s = 0; s < N; s++) counts[s] = 0;
  if (!n) return ABL_OK;
  const size_t bytes = n * (size_t)p.stride;
  const size_t aux = round_up(bytes, 256);
  TRY(ensure_stage(rt, aux + 2 * round_up(n * sizeof(u32), 256) + 2 * 256 * sizeof(u32)));
  u8 *aos = (u8 *)rt->stage;
  u32 *owner = (u32 *)(aos + aux);
  u32 *slot = (u32 *)((u8 *)owner + round_up(n * sizeof(u32), 256));
  u32 *cnt = (u32 *)((u8 *)slot + round_up(n * sizeof(u32), 256));
  u32 *off = cnt + 256;
  CU(cudaMemcpyAsync(aos, host_aos, bytes, cudaMemcpyHostToDevice, rt->stream));
  CU(cudaMemsetAsync(cnt, 0, 256 * sizeof(u32), rt->stream));
  SlabBounds sb;
  sb.n = N;
  for (int s = 0; s <= N; s++) sb.b[s] = rt->slab_bounds[s];
  const GridParams &g = rt->grid;
  const int axis = g.dim - 1;
  const Member &pm = p.members[p.pos_member];
  ColTable t;
  fill_table(p, t, false);
  // host offset of the slab-axis coordinate: a packed 2-vector column in 2D, the third of three scalar columns in 3D
  const u32 pos_off = g.dim == 2 ? (u32)t.host_off[pm.first_col] + (u32)axis * (u32)rt->real_size
                                 : (u32)t.host_off[pm.first_col + axis];
  const u32 nb = blocks_for(n, 256);
  if (rt->real_size == 8)
    k_route_classify<double><<<nb, 256, 0, rt->stream>>>(aos, p.stride, pos_off, (u32)n, g.origin[axis], g.inv_cell,
                                                        slab_layers(rt), sb, owner, slot, cnt);
  else
    k_route_classify<float><<<nb, 256, 0, rt->stream>>>(aos, p.stride, pos_off, (u32)n, (float)g.origin[axis],
                                                       (float)g.inv_cell, slab_layers(rt), sb, owner, slot, cnt);
  u32 h_cnt[ABL_MAX_SLABS];
  CU(cudaMemcpyAsync(h_cnt, cnt, N * sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  u32 h_off[ABL_MAX_SLABS];
  u32 run = 0;
  for (int s = 0; s < N; s++) { h_off[s] = run; run += h_cnt[s]; counts[s] = h_cnt[s]; }
  CU(cudaMemcpyAsync(off, h_off, N * sizeof(u32), cudaMemcpyHostToDevice, rt->stream));
  k_route_pack<<<nb, 256, 0, rt->stream>>>(aos, p.stride, (u32)n, first_id, owner, slot, off, (u8 *)dev_out);
  rt->launches += 2;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(rt->stream));   // the caller hands dev_out to another stream / device next
  return ABL_OK;
}

// n transit records in device memory (all of them owned by this slab) become the pool's population;
// next_id: first id a run-time add() may hand out (the size of the whole population).
extern "C" int abl_cuda_adopt_records(abl_runtime *rt, int pool, const void *dev_records, size_t n, unsigned next_id) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (p.stride % 4u) return fail(ABL_ERR_STATE, "pool %s: record size %u is not a multiple of 4", p.name.c_str(), p.stride);
  if (n > 0x7fffffffu) return fail(ABL_ERR_ARGUMENT, "pool too large");
  if (n && !dev_records) return fail(ABL_ERR_ARGUMENT, "null records");
  CU(cudaSetDevice(rt->device));
  p.n = 0;
  TRY(reserve_pool(rt, p, std::max(n, (size_t)1) + (halo_direct(p) ? 2 * p.halo_cap + 1024 : 0)));
  if (n) {
    ColTable t;
    fill_table(p, t, false);
    k_transit_to_soa<<<blocks_for(n, 256), 256, 0, rt->stream>>>(t, (const u8 *)dev_records, p.stride, (u32)n);
    rt->launches++;
    CU(cudaGetLastError());
  }
  p.n = n;
  p.order_serial++;
  p.next_id = next_id;
  p.binned = false;
  p.own_valid = false;
  p.report_pending = false;
  p.exch_deferred = false;
  p.halo_pending = false;
  p.halo_sync_check = false;
  p.src_begin = 0;
  p.own_begin = 0;
  p.own_end = (u32)n;
  TRY(drop_fused_histogram(rt, p));
  p.ever_removed = false;
  CU(cudaStreamSynchronize(rt->stream));   // the caller may release dev_records
  return ABL_OK;
}

extern "C" int abl_cuda_download(abl_runtime *rt, int pool, void *host_aos, size_t capacity,
                                 size_t *n_out) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  CU(cudaSetDevice(rt->device));
  // slab mode: only the owned agents are this rank's to report
  u32 first = 0;
  size_t count = p->n;
  if (rt->slab && p->pos_member >= 0) {
    TRY(slab_settle(rt, *p));
    first = p->own_begin;
    count = p->own_end - p->own_begin;
  } else if (p->src_begin) {
    first = p->src_begin;
  }
  if (n_out) *n_out = count;
  if (capacity < count) return fail(ABL_ERR_CAPACITY, "download: buffer holds %zu agents, pool has %zu", capacity, count);
  if (count == 0) return ABL_OK;
  size_t bytes = count * (size_t)p->stride;
  size_t rank_bytes = 0;
  const bool dense = !rt->slab && p->next_id == p->n;  // ids are a permutation of 0..n-1
  if (!dense) rank_bytes = round_up((size_t)p->next_id + 1, kScanTile) * sizeof(u32) * 2;
  TRY(ensure_stage(rt, round_up(bytes, 256) + rank_bytes));
  const u32 *ids = (const u32 *)p->cols[p->id_col].buf[p->cols[p->id_col].cur] + first;
  u32 *rank = nullptr;
  if (!dense) {
    u32 *present = (u32 *)((u8 *)rt->stage + round_up(bytes, 256));
    size_t padded = round_up((size_t)p->next_id + 1, kScanTile);
    rank = present + padded;
    CU(cudaMemsetAsync(present, 0, padded * sizeof(u32), rt->stream));
    k_mark_present<<<blocks_for(count, 256), 256, 0, rt->stream>>>(ids, (u32)count, present);
    rt->launches++;
    TRY((run_scan<u32, 0, false>(rt, present, rank, (size_t)p->next_id, nullptr)));
  }
  // padding bytes of the host records are zeroed for reproducible raw dumps
  CU(cudaMemsetAsync(rt->stage, 0, bytes, rt->stream));
  ColTable t;
  fill_table(*p, t, false);
  for (int c = 0; c < t.ncols; c++) t.in[c] = (const u8 *)t.in[c] + (size_t)first * t.elem[c];
  k_soa_to_aos<<<blocks_for(count, 256), 256, 0, rt->stream>>>(t, (u8 *)rt->stage, p->stride,
                                                               (u32)count, ids, rank);
  rt->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(host_aos, rt->stage, bytes, cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  return ABL_OK;
}

__global__ void k_ids_sorted(const u32 *ids, const u32 *rank, u32 n, u32 *out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[rank ? rank[ids[i]] : ids[i]] = ids[i];
}

// ids of the agents abl_cuda_download returns, in the same (ascending) order
extern "C" int abl_cuda_download_ids(abl_runtime *rt, int pool, unsigned *ids_out, size_t capacity,
                                     size_t *n_out) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  CU(cudaSetDevice(rt->device));
  u32 first = 0;
  size_t count = p->n;
  if (rt->slab && p->pos_member >= 0) {
    TRY(slab_settle(rt, *p));
    first = p->own_begin;
    count = p->own_end - p->own_begin;
  } else if (p->src_begin) {
    first = p->src_begin;
  }
  if (n_out) *n_out = count;
  if (capacity < count) return fail(ABL_ERR_CAPACITY, "download_ids: buffer too small");
  if (!count) return ABL_OK;
  const bool dense = !rt->slab && p->next_id == p->n;
  size_t padded = round_up((size_t)p->next_id + 1, kScanTile);
  TRY(ensure_stage(rt, round_up(count * sizeof(u32), 256) + padded * sizeof(u32) * 2));
  const u32 *ids = (const u32 *)p->cols[p->id_col].buf[p->cols[p->id_col].cur] + first;
  u32 *out = (u32 *)rt->stage;
  u32 *rank = nullptr;
  if (!dense) {
    u32 *present = (u32 *)((u8 *)rt->stage + round_up(count * sizeof(u32), 256));
    rank = present + padded;
    CU(cudaMemsetAsync(present, 0, padded * sizeof(u32), rt->stream));
    k_mark_present<<<blocks_for(count, 256), 256, 0, rt->stream>>>(ids, (u32)count, present);
    TRY((run_scan<u32, 0, false>(rt, present, rank, (size_t)p->next_id, nullptr)));
  }
  k_ids_sorted<<<blocks_for(count, 256), 256, 0, rt->stream>>>(ids, rank, (u32)count, out);
  rt->launches += 2;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ids_out, out, count * sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------------
static int ensure_grid_arrays(abl_runtime *rt, Pool &p) {
  if (p.cell_count) return ABL_OK;
  size_t padded = round_up((size_t)rt->grid.n_local + 2, kScanTile);
  CU(cudaMalloc(&p.cell_count, padded * sizeof(u32)));
  CU(cudaMalloc(&p.cell_start, padded * sizeof(u32)));
  CU(cudaMemsetAsync(p.cell_count, 0, padded * sizeof(u32), rt->stream));
  CU(cudaMemsetAsync(p.cell_start, 0, padded * sizeof(u32), rt->stream));
  return ABL_OK;
}

static int slab_consume_report(abl_runtime *rt, Pool &p, bool *redo);
static int slab_settle(abl_runtime *rt, Pool &p);
struct ScanReport;
static int slab_request_owned_range(abl_runtime *rt, Pool &p, bool reported, const ScanReport &rep);
static void slab_owned_cells(const abl_runtime *rt, u32 *lo_cell, u32 *hi_cell);
static void slab_boundary_cells(const abl_runtime *rt, u32 *lo2_cell, u32 *hi2_cell);

// histogram of `n` records starting at source index `src_begin`; keys/ranks go to slot
// `out_begin + i` of the pool's key/local arrays
static int launch_bin_count(abl_runtime *rt, Pool &p, u32 n, u32 src_begin, u32 out_begin) {
  const GridParams &g = rt->grid;
  const Member &pm = p.members[p.pos_member];
  const int bs = 256;
  if (!n) return ABL_OK;
  const void *px = p.cols[pm.first_col].buf[p.cols[pm.first_col].cur];
  const void *py = nullptr, *pz = nullptr;
  if (g.dim == 3) {
    py = p.cols[pm.first_col + 1].buf[p.cols[pm.first_col + 1].cur];
    pz = p.cols[pm.first_col + 2].buf[p.cols[pm.first_col + 2].cur];
  }
  u32 nb = blocks_for(n, bs);
  const u32 *ids = rt->slab ? (const u32 *)p.cols[p.id_col].buf[p.cols[p.id_col].cur] : nullptr;
  if (rt->real_size == 8) {
    if (g.dim == 2) k_bin_count<double, 2><<<nb, bs, 0, rt->stream>>>(px, py, pz, n, src_begin, out_begin, g, p.key, p.local, p.cell_count, ids);
    else k_bin_count<double, 3><<<nb, bs, 0, rt->stream>>>(px, py, pz, n, src_begin, out_begin, g, p.key, p.local, p.cell_count, ids);
  } else {
    if (g.dim == 2) k_bin_count<float, 2><<<nb, bs, 0, rt->stream>>>(px, py, pz, n, src_begin, out_begin, g, p.key, p.local, p.cell_count, ids);
    else k_bin_count<float, 3><<<nb, bs, 0, rt->stream>>>(px, py, pz, n, src_begin, out_begin, g, p.key, p.local, p.cell_count, ids);
  }
  rt->launches++;
  CU(cudaGetLastError());
  return ABL_OK;
}

// defer_report (slab mode): do not wait for the owned range; the caller queues kernels that
// read it on the device and the host catches up later (slab_consume_report)
static int bin_pool(abl_runtime *rt, Pool &p, bool defer_report) {
  if (!rt->env_set) return fail(ABL_ERR_STATE, "binning requires an environment");
  if (p.pos_member < 0) return fail(ABL_ERR_STATE, "pool %s has no position member", p.name.c_str());
  // the extent of this binning may still be on its way from the device
  if (rt->slab) TRY(slab_consume_report(rt, p, nullptr));
  TRY(reserve_pool(rt, p, std::max(p.n, (size_t)1)));
  TRY(ensure_grid_arrays(rt, p));
  const GridParams &g = rt->grid;
  const u32 n = (u32)p.n;
  const int bs = 256;
  trace_stamp(rt, TR_OTHER);
  // 1. histogram (skipped when the step kernel that produced the positions already did it)
  if (!p.counted) TRY(launch_bin_count(rt, p, n, p.src_begin, 0u));
  p.counted = false;
  // 2. cell_start[c] = number of agents in cells < c; entry n_cells = n.  The scan also
  //    clears the histogram for the next binning.
  bool scatter_reports = false;
  ScanReport scatter_rep;
  memset(&scatter_rep, 0, sizeof scatter_rep);
  if (rt->slab) {
    ScanReport rep;
    bool reported = false;
    slab_owned_cells(rt, &rep.lo_cell, &rep.hi_cell);
    slab_boundary_cells(rt, &rep.lo2_cell, &rep.hi2_cell);
    rep.host = report_dev(rt, p);
    rep.stamp = p.report_stamp = ++rt->bin_stamp;
    rep.halo_ctr = p.halo_pending ? p.halo_ctr : nullptr;
    p.halo_pending = false;
    // (the report travels with k_bin_scatter when there is one; ABL_CUDA_REPORT_IN_SCAN=1: with the scan)
    scatter_reports = n != 0 && !rt->report_in_scan;
    if (scatter_reports) scatter_rep = rep;
    TRY(run_cell_scan(rt, p.cell_count, p.cell_start, (size_t)g.n_local + 2, scatter_reports ? nullptr : &rep, &reported));
    if (!scatter_reports) TRY(slab_request_owned_range(rt, p, reported, rep));
    p.report_pending = true;
    p.own_valid = false;
  } else {
    TRY(run_cell_scan(rt, p.cell_count, p.cell_start, (size_t)g.n_local + 2));
  }
  trace_stamp(rt, TR_SCAN);
  if (n) {
    // 3. ids into their cell segments, 4. rank by id inside the segment + move the records
    const u32 *ids = (const u32 *)p.cols[p.id_col].buf[p.cols[p.id_col].cur];
    u32 *seg_ids = (u32 *)p.pairs;
    u32 nb = blocks_for(n, bs);
    CU(launch_pdl(rt->pdl, k_bin_scatter, dim3(nb), dim3(bs), 0, rt->stream, (const u32 *)p.key, p.local, ids, n,
                  p.src_begin, (const u32 *)p.cell_start, seg_ids, p.cell_count, scatter_rep, seg_ids + p.cap));
    ColTable t;
    fill_table(p, t, true);
    CU(launch_pdl(rt->pdl, k_bin_rank_move, dim3(nb), dim3(bs), 0, rt->stream, t, (const u32 *)seg_ids, (const u32 *)p.key,
                  (const u32 *)p.local, ids, n, p.src_begin, (const u32 *)p.cell_start, (const u32 *)(seg_ids + p.cap)));
    rt->launches += 2;
    CU(cudaGetLastError());
    flip_all(p);
  }
  p.order_serial++;
  trace_stamp(rt, TR_MOVE);
  const u32 src_before = p.src_begin;
  p.src_begin = 0;
  p.key_base_hint = g.key_base;
  p.binned = true;
  p.ever_binned = true;
  if (rt->slab) {
    if (defer_report && !p.halo_sync_check) return ABL_OK;
    bool redo = false;
    TRY(slab_consume_report(rt, p, &redo));
    p.halo_sync_check = false;
    if (redo) {
      // More halo records arrived than the padding the host assumed: they were unpacked (the
      // pool has room for two full messages) but not binned.  The source buffers of this
      // binning are still intact in the alternate halves: bin again over the full range.
      flip_all(p);
      p.src_begin = src_before;
      p.n = (size_t)p.halo_prev_own + p.halo_last_arrivals;
      p.binned = false;
      p.counted = false;
      CU(cudaMemsetAsync(p.cell_count, 0, ((size_t)g.n_local + 2) * sizeof(u32), rt->stream));
      return bin_pool(rt, p, defer_report);
    }
  } else {
    p.own_begin = 0;
    p.own_end = (u32)p.n;
    p.own_valid = true;
  }
  return ABL_OK;
}

extern "C" int abl_cuda_bin(abl_runtime *rt, int pool) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  CU(cudaSetDevice(rt->device));
  if (rt->timing) CU(cudaEventRecord(rt->ev[0], rt->stream));
  TRY(bin_pool(rt, *p));
  if (rt->timing) {
    CU(cudaEventRecord(rt->ev[1], rt->stream));
    CU(cudaEventSynchronize(rt->ev[1]));
    CU(cudaEventElapsedTime(&rt->last.bin_ms, rt->ev[0], rt->ev[1]));
  }
  return ABL_OK;
}

extern "C" int abl_cuda_debug_binning(abl_runtime *rt, int pool, unsigned *cell_start,
                                      size_t n_cells_plus_1, unsigned *ids, size_t n_ids) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  if (!p->binned) return fail(ABL_ERR_STATE, "pool is not binned");
  CU(cudaStreamSynchronize(rt->stream));
  if (cell_start) {
    size_t m = std::min(n_cells_plus_1, (size_t)rt->grid.n_local + 1);
    CU(cudaMemcpy(cell_start, p->cell_start, m * sizeof(u32), cudaMemcpyDeviceToHost));
  }
  if (ids) {
    size_t m = std::min(n_ids, p->n);
    CU(cudaMemcpy(ids, p->cols[p->id_col].buf[p->cols[p->id_col].cur], m * sizeof(u32),
                  cudaMemcpyDeviceToHost));
  }
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// step functions
// ---------------------------------------------------------------------------------------
extern "C" int abl_cuda_register_step(abl_runtime *rt, const abl_step_desc *desc, int *step) {
  if (!rt || !desc || !desc->launch) return fail(ABL_ERR_ARGUMENT, "bad step descriptor");
  Pool *p = nullptr;
  TRY(get_pool(rt, desc->self_pool, &p));
  if (desc->nbr_pool >= 0) {
    Pool *q = nullptr;
    TRY(get_pool(rt, desc->nbr_pool, &q));
    if (q->pos_member < 0) return fail(ABL_ERR_ARGUMENT, "step %s: neighbour pool has no position", desc->name);
    if (!rt->env_set) return fail(ABL_ERR_STATE, "register_step before set_environment");
  }
  if (desc->added_pool >= 0) {
    Pool *q = nullptr;
    TRY(get_pool(rt, desc->added_pool, &q));
  }
  Step s;
  s.desc = *desc;
  s.name = desc->name ? desc->name : "";
  s.desc.name = nullptr;
  s.reach = 1;
  if (desc->nbr_pool >= 0) {
    double r = desc->radius / rt->grid.cell;
    int reach = (int)ceil(r - 1e-12);
    s.reach = reach < 1 ? 1 : reach;
  }
  rt->steps.push_back(s);
  if (step) *step = (int)rt->steps.size() - 1;
  return ABL_OK;
}

static void fill_view(const Pool &p, abl_pool_view &v, uint32_t written_members, bool self) {
  memset(&v, 0, sizeof v);
  v.n = (unsigned)p.n;
  int ncols = (int)p.cols.size() - 1;
  for (int c = 0; c < ncols; c++) {
    v.in[c] = p.cols[c].buf[p.cols[c].cur];
    v.out[c] = p.cols[c].buf[p.cols[c].cur];
  }
  if (self) {
    for (size_t m = 0; m < p.members.size(); m++) {
      if (!(written_members >> m & 1u)) continue;
      const Member &mem = p.members[m];
      for (int c = mem.first_col; c < mem.first_col + mem.ncols; c++)
        v.out[c] = p.cols[c].buf[p.cols[c].cur ^ 1];
    }
  }
  v.id = (const unsigned *)p.cols[p.id_col].buf[p.cols[p.id_col].cur];
  v.cell_start = p.binned ? p.cell_start - p.key_base_hint : nullptr;
}

static int read_scalar(abl_runtime *rt, const u32 *d, u32 *out) {
  CU(cudaMemcpyAsync(rt->h_scalar, d, sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  *out = rt->h_scalar[0];
  return ABL_OK;
}

// Stream compaction of the records [first, first + n) by the flags dead[0..n).  Under slab
// decomposition that range is the owned part of the pool (ghosts are stale after a mutating step
// and are replaced by the exchange that follows): survivors move to the front of the alternate
// buffers and become the new owned range.
static int commit_removals(abl_runtime *rt, Pool &p, u32 first, u32 n, bool slab) {
  if (!n) return ABL_OK;
  TRY((run_scan<u8, 1, false>(rt, p.dead, p.offsets, n, rt->d_scalar)));
  u32 survivors = 0;
  TRY(read_scalar(rt, rt->d_scalar, &survivors));
  if (survivors == n) return ABL_OK;  // nobody died: nothing to move
  ColTable t;
  fill_table(p, t, true);
  k_compact_move<<<blocks_for(n, 256), 256, 0, rt->stream>>>(t, p.dead, p.offsets, n, first);
  rt->launches++;
  CU(cudaGetLastError());
  flip_all(p);
  p.order_serial++;
  p.ever_removed = true;
  p.binned = false;
  TRY(drop_fused_histogram(rt, p));  // stable compaction keeps cell order, but cell_start is stale
  p.n = survivors;
  if (slab) {
    p.own_begin = 0;
    p.own_end = survivors;
    p.src_begin = 0;
    p.own_valid = true;
  }
  return ABL_OK;
}

static int commit_adds(abl_runtime *rt, Pool &parent, Pool &target, void *const *staging) {
  const u32 n = (u32)parent.n;
  if (!n) return ABL_OK;
  TRY((run_scan<u8, 0, false>(rt, parent.add_flag, parent.offsets, n, rt->d_scalar)));
  u32 m = 0;
  TRY(read_scalar(rt, rt->d_scalar, &m));
  if (!m) return ABL_OK;
  const u32 *pids = (const u32 *)parent.cols[parent.id_col].buf[parent.cols[parent.id_col].cur];
  u64 *list = parent.pairs;  // free between binnings
  k_collect_adds<<<blocks_for(n, 256), 256, 0, rt->stream>>>(parent.add_flag, parent.offsets, pids, n, list);
  rt->launches++;
  CU(cudaGetLastError());
  u64 *tmp = nullptr;
  if (target.n + m > target.cap) {
    // growing the target reallocates its scratch (which holds `list` when parent == target)
    CU(cudaMalloc(&tmp, (size_t)m * sizeof(u64)));
    CU(cudaMemcpyAsync(tmp, list, (size_t)m * sizeof(u64), cudaMemcpyDeviceToDevice, rt->stream));
    CU(cudaStreamSynchronize(rt->stream));
    TRY(reserve_pool(rt, target, target.n + m));
    list = tmp;
  }
  ColTable t;
  fill_table(target, t, false);
  for (int c = 0; c < t.ncols; c++) t.in[c] = t.host_off[c] < 0 ? nullptr : staging[c];
  if (m > append_scan_threshold()) {
    const size_t padded = round_up((size_t)parent.next_id + 1, kScanTile);
    if (rt->rank_cap < 2 * padded) {
      CU(cudaStreamSynchronize(rt->stream));
      if (rt->rank_buf) CU(cudaFree(rt->rank_buf));
      rt->rank_buf = nullptr;
      rt->rank_cap = 0;
      const size_t cap = round_up(2 * padded + padded / 4, kScanTile);
      CU(cudaMalloc(&rt->rank_buf, cap * sizeof(u32)));
      rt->rank_cap = cap;
    }
    u32 *present = rt->rank_buf, *rank_of_id = rt->rank_buf + rt->rank_cap / 2;
    CU(cudaMemsetAsync(present, 0, padded * sizeof(u32), rt->stream));
    k_mark_parents<<<blocks_for(m, 256), 256, 0, rt->stream>>>(list, m, present);
    rt->launches++;
    TRY((run_scan<u32, 0, false>(rt, present, rank_of_id, (size_t)parent.next_id, nullptr)));
    k_append_by_id<<<blocks_for(m, 128), 128, 0, rt->stream>>>(t, list, rank_of_id, m, (u32)target.n, target.next_id);
  } else {
    k_append<<<blocks_for(m, 128), 128, 0, rt->stream>>>(t, list, m, (u32)target.n, target.next_id);
  }
  rt->launches++;
  CU(cudaGetLastError());
  if (tmp) {
    CU(cudaStreamSynchronize(rt->stream));
    CU(cudaFree(tmp));
  }
  target.n += m;
  target.order_serial++;
  target.next_id += m;
  target.binned = false;
  TRY(drop_fused_histogram(rt, target));
  return ABL_OK;
}

// ---- run-time add() under slab decomposition -------------------------------------------------
// New agents get ids next_id + (rank of their parent's id among the parents of ALL slabs), the
// numbering an undecomposed run produces (k_append).  The step function therefore stays open
// after its kernel: the parents' ids are brought to the host in ascending order, the caller
// exchanges them between the slabs and hands back global ranks and the global total.
static int slab_open_adds(abl_runtime *rt, int step, Pool &parent, u32 first, u32 n, void *const *staging) {
  rt->open_list.clear();
  u32 m = 0;
  if (n) {
    TRY((run_scan<u8, 0, false>(rt, parent.add_flag, parent.offsets, n, rt->d_scalar)));
    TRY(read_scalar(rt, rt->d_scalar, &m));
  }
  if (m) {
    const u32 *pids = (const u32 *)parent.cols[parent.id_col].buf[parent.cols[parent.id_col].cur] + first;
    k_collect_adds<<<blocks_for(n, 256), 256, 0, rt->stream>>>(parent.add_flag, parent.offsets, pids, n, parent.pairs);
    rt->launches++;
    CU(cudaGetLastError());
    rt->open_list.resize(m);
    CU(cudaMemcpyAsync(rt->open_list.data(), parent.pairs, (size_t)m * sizeof(u64), cudaMemcpyDeviceToHost, rt->stream));
    CU(cudaStreamSynchronize(rt->stream));
    std::sort(rt->open_list.begin(), rt->open_list.end());
  }
  memcpy(rt->open_staging, staging, sizeof rt->open_staging);
  rt->open_step = step;
  return ABL_OK;
}

static int slab_after_mutation(abl_runtime *rt, int pool_index);

extern "C" int abl_cuda_pending_adds(abl_runtime *rt, int *open, unsigned *count) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (open) *open = rt->open_step >= 0 ? 1 : 0;
  if (count) *count = rt->open_step >= 0 ? (unsigned)rt->open_list.size() : 0u;
  return ABL_OK;
}

extern "C" int abl_cuda_pending_add_parents(abl_runtime *rt, unsigned *parent_ids, size_t capacity) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (rt->open_step < 0) return fail(ABL_ERR_STATE, "no step function is waiting for abl_cuda_resolve_adds");
  if (capacity < rt->open_list.size()) return fail(ABL_ERR_CAPACITY, "parent id buffer too small");
  for (size_t i = 0; i < rt->open_list.size(); i++) parent_ids[i] = (unsigned)(rt->open_list[i] >> 32);
  return ABL_OK;
}

extern "C" int abl_cuda_resolve_adds(abl_runtime *rt, const unsigned *global_rank, unsigned global_total) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (rt->open_step < 0) return fail(ABL_ERR_STATE, "no step function is waiting for abl_cuda_resolve_adds");
  CU(cudaSetDevice(rt->device));
  Step &s = rt->steps[rt->open_step];
  Pool &p = rt->pools[s.desc.self_pool];   // (added pool == stepped pool: checked by abl_cuda_step)
  const u32 m = (u32)rt->open_list.size();
  if (m && !global_rank) return fail(ABL_ERR_ARGUMENT, "null rank array");
  for (u32 i = 0; i < m; i++)
    if (global_rank[i] >= global_total) return fail(ABL_ERR_ARGUMENT, "global rank %u out of range (total %u)", global_rank[i], global_total);
  rt->open_step = -1;
  if (s.desc.uses_removal) TRY(commit_removals(rt, p, p.own_begin, p.own_end - p.own_begin, true));
  if (m) {
    const size_t want = (size_t)p.own_end + m;
    if (want > p.cap) {
      p.n = std::max(p.n, (size_t)p.own_end);
      TRY(drop_fused_histogram(rt, p));
      TRY(reserve_pool(rt, p, want));
    }
    // (pairs / offsets are free between binnings)
    CU(cudaMemcpyAsync(p.pairs, rt->open_list.data(), (size_t)m * sizeof(u64), cudaMemcpyHostToDevice, rt->stream));
    CU(cudaMemcpyAsync(p.offsets, global_rank, (size_t)m * sizeof(u32), cudaMemcpyHostToDevice, rt->stream));
    ColTable t;
    fill_table(p, t, false);
    for (int c = 0; c < t.ncols; c++) t.in[c] = t.host_off[c] < 0 ? nullptr : rt->open_staging[c];
    k_append_ranked<<<blocks_for(m, 128), 128, 0, rt->stream>>>(t, p.pairs, p.offsets, m, p.own_end, p.next_id);
    rt->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(rt->stream));   // the host arrays may go away
    p.own_end += m;
    p.n = std::max(p.n, (size_t)p.own_end);
    p.binned = false;
    TRY(drop_fused_histogram(rt, p));
  }
  p.next_id += global_total;
  return slab_after_mutation(rt, s.desc.self_pool);
}

// Evaluates the queued step timings: rt->last becomes the MEAN per abl_cuda_step call over the
// steps since the last evaluation (a caller that asks after every step gets that step's times).
static int drain_timing(abl_runtime *rt) {
  const size_t quads = rt->timing_log.size() / 4;
  if (!quads) return ABL_OK;
  CU(cudaEventSynchronize(rt->timing_log.back()));
  double bin = 0, kernel = 0, commit = 0;
  for (size_t q = 0; q < quads; q++) {
    cudaEvent_t *e = &rt->timing_log[4 * q];
    float a = 0, b = 0, c = 0;
    CU(cudaEventElapsedTime(&a, e[0], e[1]));
    CU(cudaEventElapsedTime(&b, e[1], e[2]));
    CU(cudaEventElapsedTime(&c, e[2], e[3]));
    bin += a; kernel += b; commit += c;
  }
  for (cudaEvent_t e : rt->timing_log) cudaEventDestroy(e);
  rt->timing_log.clear();
  rt->last.bin_ms = (float)(bin / quads);
  rt->last.kernel_ms = (float)(kernel / quads);
  rt->last.commit_ms = (float)(commit / quads);
  return ABL_OK;
}

// Builds the neighbour lists of step `s` for the views in `a` (count pass -> size by the largest
// count -> fill pass; both passes are launches of the generated kernel that store nothing of the
// step itself).  One host synchronisation, once per build.
// Single-precision shadow of a pool's positions for the pre-filter of ABL_MODE 8 (abl_device.cuh):
// float4 (x, y, z, 0) per record in pool order plus the largest coordinate magnitude (the error bound
// of the pre-filter scales with it).  NaN coordinates do not enter the maximum; such candidates pass
// the pre-filter by themselves (every comparison with NaN is false) and are decided by the exact test.
template <int DIM>
__global__ void k_shadow(const double *px, const double *py, const double *pz, u32 n, float4 *shadow, u32 *max_bits) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  float m = 0.0f;
  if (i < n) {
    float4 f;
    if (DIM == 2) {
      const double2 p = reinterpret_cast<const double2 *>(px)[i];
      f = make_float4((float)p.x, (float)p.y, 0.0f, 0.0f);
    } else {
      f = make_float4((float)px[i], (float)py[i], (float)pz[i], 0.0f);
    }
    shadow[i] = f;
    m = fmaxf(fabsf(f.x), fmaxf(fabsf(f.y), fabsf(f.z)));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31u) == 0u && m > 0.0f) atomicMax(max_bits, __float_as_uint(m));   // non-negative floats order like their bits
}

static int refresh_shadow(abl_runtime *rt, Pool &p) {
  if (p.shadow && p.shadow_serial == p.order_serial && p.shadow_cap >= p.n) return ABL_OK;
  if (p.shadow_cap < std::max(p.n, (size_t)1)) {
    if (p.shadow) TRY(release_device(rt, p.shadow));
    p.shadow = nullptr;
    p.shadow_cap = std::max(p.cap, std::max(p.n, (size_t)1));
    CU(cudaMalloc(&p.shadow, p.shadow_cap * sizeof(float4)));
  }
  if (!p.shadow_max) CU(cudaMalloc(&p.shadow_max, sizeof(u32)));
  CU(cudaMemsetAsync(p.shadow_max, 0, sizeof(u32), rt->stream));
  if (p.n) {
    const Member &pm = p.members[p.pos_member];
    const double *px = (const double *)p.cols[pm.first_col].buf[p.cols[pm.first_col].cur];
    const u32 nb = blocks_for(p.n, 256);
    if (rt->grid.dim == 2) {
      k_shadow<2><<<nb, 256, 0, rt->stream>>>(px, nullptr, nullptr, (u32)p.n, p.shadow, p.shadow_max);
    } else {
      const double *py = (const double *)p.cols[pm.first_col + 1].buf[p.cols[pm.first_col + 1].cur];
      const double *pz = (const double *)p.cols[pm.first_col + 2].buf[p.cols[pm.first_col + 2].cur];
      k_shadow<3><<<nb, 256, 0, rt->stream>>>(px, py, pz, (u32)p.n, p.shadow, p.shadow_max);
    }
    rt->launches++;
    CU(cudaGetLastError());
  }
  p.shadow_serial = p.order_serial;
  return ABL_OK;
}

static int build_neighbour_lists(abl_runtime *rt, Step &s, const abl_step_launch &a, const Pool &self, const Pool &nbr) {
  NeighbourLists &nl = s.nl;
  nl.valid = false;
  const size_t n = a.self.n;
  if (nl.cnt_cap < n) {
    if (nl.cnt) CU(cudaFree(nl.cnt));
    nl.cnt = nullptr;
    nl.cnt_cap = round_up(n + n / 8, 1024);
    CU(cudaMalloc(&nl.cnt, nl.cnt_cap * sizeof(u32)));
  }
  u32 *d_max = rt->d_scalar + 1020;   // a word of the 4 KB scratch no other path uses
  CU(cudaMemsetAsync(d_max, 0, sizeof(u32), rt->stream));
  abl_step_launch b = a;
  b.bin_key = b.bin_local = b.bin_count = nullptr;
  b.nlist_phase = 1;
  b.nlist_cnt = nl.cnt;
  b.nlist_idx = nullptr;
  b.nlist_stride = 0;
  b.nlist_max = d_max;
  int rc = s.desc.launch(&b);
  rt->launches++;
  if (rc != 0) return fail(ABL_ERR_CUDA, "step %s: neighbour-list count pass failed: %s", s.name.c_str(), cudaGetErrorString((cudaError_t)rc));
  u32 max_degree = 0;
  TRY(read_scalar(rt, d_max, &max_degree));
  const size_t words = (size_t)std::max(max_degree, 1u) * n;
  if (words * sizeof(u32) > rt->nlist_budget) {
    nl.off = true;   // e.g. a crowd in one cell: the k-major layout pads every agent to the largest list
    if (getenv("ABL_CUDA_VERBOSE"))
      fprintf(stderr, "abl_cuda: step %s: neighbour lists need %.1f MB (largest list %u), over the budget: ordinary loop\n",
              s.name.c_str(), words * 4.0 / 1048576.0, max_degree);
    return ABL_OK;
  }
  if (nl.idx_cap < words) {
    if (nl.idx) CU(cudaFree(nl.idx));
    nl.idx = nullptr;
    nl.idx_cap = 0;
    // the lists are an optimisation: when the device has no room for them the step keeps the ordinary loop
    if (cudaMalloc(&nl.idx, words * sizeof(u32)) != cudaSuccess) {
      cudaGetLastError();
      nl.idx = nullptr;
      nl.off = true;
      if (getenv("ABL_CUDA_VERBOSE"))
        fprintf(stderr, "abl_cuda: step %s: no device memory for %.1f MB of neighbour lists: ordinary loop\n",
                s.name.c_str(), words * 4.0 / 1048576.0);
      return ABL_OK;
    }
    nl.idx_cap = words;
  }
  b.nlist_phase = 2;
  b.nlist_idx = nl.idx;
  b.nlist_stride = (u32)n;
  rc = s.desc.launch(&b);
  rt->launches++;
  if (rc != 0) return fail(ABL_ERR_CUDA, "step %s: neighbour-list fill pass failed: %s", s.name.c_str(), cudaGetErrorString((cudaError_t)rc));
  CU(cudaGetLastError());
  nl.stride = (u32)n;
  nl.max_degree = max_degree;
  nl.self_serial = self.order_serial;
  nl.nbr_serial = nbr.order_serial;
  nl.n_self = self.n;
  nl.n_nbr = nbr.n;
  nl.valid = true;
  nl.builds++;
  if (getenv("ABL_CUDA_VERBOSE"))
    fprintf(stderr, "abl_cuda: step %s: neighbour lists built (%zu agents, largest list %u, %.1f MB)\n", s.name.c_str(), n,
            max_degree, words * 4.0 / 1048576.0);
  return ABL_OK;
}

extern "C" int abl_cuda_step(abl_runtime *rt, int step) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (step < 0 || step >= (int)rt->steps.size()) return fail(ABL_ERR_ARGUMENT, "bad step index %d", step);
  CU(cudaSetDevice(rt->device));
  Step &s = rt->steps[step];
  Pool &self = rt->pools[s.desc.self_pool];
  Pool *nbr = s.desc.nbr_pool >= 0 ? &rt->pools[s.desc.nbr_pool] : nullptr;
  Pool *added = s.desc.added_pool >= 0 ? &rt->pools[s.desc.added_pool] : nullptr;
  if (rt->open_step >= 0)
    return fail(ABL_ERR_STATE, "step %s added agents under slab decomposition: call abl_cuda_resolve_adds before the next step",
                rt->steps[rt->open_step].name.c_str());
  // a step that removes or adds agents: under slab decomposition the runtime commits over the
  // owned range and exchanges in a separate pass afterwards (no fused send, no device range)
  const bool mutating = s.desc.uses_removal || added;
  const bool slab_pool = rt->slab && self.pos_member >= 0;
  if (rt->slab && mutating) {
    if (!slab_pool) return fail(ABL_ERR_STATE, "step %s: run-time add/remove of agents without a position is not supported with slab decomposition", s.name.c_str());
    if (added && added != &self)
      return fail(ABL_ERR_STATE, "step %s: adding agents of another type is not supported with slab decomposition", s.name.c_str());
  }

  if (rt->timing) CU(cudaEventRecord(rt->ev[0], rt->stream));
  // slab mode with the direct transport: this step's kernel packs the halo itself, and (device
  // range) reads the owned range of its pool from cell_start, so that it can be queued without
  // the host waiting for the binning
  const bool direct = rt->slab && self.pos_member >= 0 && halo_direct(self) && s.desc.written_members && !mutating;
  const bool dev_range = direct && rt->device_range && !rt->timing;
  if (nbr) {
    if (!nbr->binned) TRY(bin_pool(rt, *nbr, dev_range && nbr == &self));
    // keep the iterating pool in cell order too: neighbouring threads then walk
    // neighbouring cells (coalescing / L1 reuse)
    if (&self != nbr && self.pos_member >= 0 && !self.binned) TRY(bin_pool(rt, self, dev_range));
  }
  if (rt->slab) {
    if (self.pos_member >= 0 && !self.binned) TRY(bin_pool(rt, self, dev_range));
    // a pool whose range the host has to know (no device range for this launch)
    if (self.pos_member >= 0 && !dev_range && self.report_pending) TRY(slab_consume_report(rt, self, nullptr));
  }
  // the device range can only be used while the host's view is behind (report pending); once it
  // has caught up (first step after an upload, timing mode) the classic host-side range is used
  const bool use_dev_range = dev_range && self.report_pending;
  if (rt->timing) CU(cudaEventRecord(rt->ev[1], rt->stream));

  if (self.n) {
    TRY(reserve_pool(rt, self, self.n));
    if (direct) TRY(halo_reserve(rt, self));  // may move the columns: before any view is taken
    abl_step_launch a;
    memset(&a, 0, sizeof a);
    fill_view(self, a.self, s.desc.written_members, true);
    if (direct) TRY(halo_fill_view(rt, self, a.slab, true, use_dev_range));
    if (rt->slab && self.pos_member >= 0 && !use_dev_range) {
      // the step function runs over the owned range only
      const u32 ob = self.own_begin;
      int ncols = (int)self.cols.size() - 1;
      for (int c = 0; c < ncols; c++) {
        a.self.in[c] = (const u8 *)a.self.in[c] + (size_t)ob * self.cols[c].elem;
        a.self.out[c] = (u8 *)a.self.out[c] + (size_t)ob * self.cols[c].elem;
      }
      a.self.id += ob;
      a.self.n = self.own_end - ob;
    }
    if (nbr) fill_view(*nbr, a.nbr, 0, false);
    a.grid.dim = rt->grid.dim;
    for (int k = 0; k < 3; k++) { a.grid.n_cell[k] = rt->grid.n_cell[k]; a.grid.origin[k] = rt->grid.origin[k]; }
    a.grid.cell_size = rt->grid.cell;
    a.grid.inv_cell_size = rt->grid.inv_cell;
    a.grid.n_cells = rt->grid.n_cells;
    a.grid.axis_lo = rt->grid.axis_lo;
    a.grid.axis_hi = rt->grid.axis_hi;
    a.grid.key_base = rt->grid.key_base;
    a.reach = s.reach;
    a.dead = s.desc.uses_removal ? self.dead : nullptr;
    void *staging[ABL_MAX_COLUMNS + 1];
    memset(staging, 0, sizeof staging);
    if (added) {
      // staging for one new agent per parent: the alternate buffers of the *target* pool
      // cannot be used (target may be the parent itself and be written by this step), so
      // dedicated staging columns are carved out of the transfer staging buffer.
      size_t bytes = 0;
      int ncols = (int)added->cols.size() - 1;
      for (int c = 0; c < ncols; c++) bytes += round_up(self.n * (size_t)added->cols[c].elem, 256);
      TRY(ensure_stage(rt, bytes));
      size_t off = 0;
      for (int c = 0; c < ncols; c++) {
        staging[c] = (u8 *)rt->stage + off;
        a.add_cols[c] = staging[c];
        off += round_up(self.n * (size_t)added->cols[c].elem, 256);
      }
      a.add_flag = self.add_flag;
    }
    // Fuse the cell histogram of the *next* binning into this kernel's epilogue when it
    // rewrites the positions of a pool that is used for neighbour search and no commit
    // stage reorders the pool afterwards.
    const bool writes_pos = self.pos_member >= 0 && (s.desc.written_members >> self.pos_member & 1u);
    bool fuse = writes_pos && rt->env_set && !s.desc.uses_removal && added != &self && self.cell_count != nullptr && self.ever_binned;
    if (writes_pos) TRY(drop_fused_histogram(rt, self));  // never consumed: start over
    if (fuse) {
      a.bin_key = self.key;
      a.bin_local = self.local;
      a.bin_count = self.cell_count;
    }
    a.seed = rt->cfg.seed;
    a.timestep = rt->timestep;
    a.step_index = (unsigned)step;
    a.block_size = rt->cfg.block_size;
    a.tile_neighbours = rt->cfg.tile_neighbours;
    a.flat_loop = rt->flat_loop;
    a.bulk_tile = rt->bulk_tile;
    // dense for-near loops pre-filter on a single-precision shadow of the neighbours' positions
    // (only when the launcher's rule picks that variant for this launch: it is asked first)
    if (s.desc.shadow && rt->dense_tile && nbr && rt->real_size == 8 && nbr->pos_member >= 0 && a.self.n) {
      a.probe = 1;
      const int wants = s.desc.launch(&a);
      a.probe = 0;
      if (wants == 1 || wants == 2) {
        TRY(refresh_shadow(rt, *nbr));
        a.nbr_shadow = nbr->shadow;
        a.nbr_shadow_max = nbr->shadow_max;
      }
      if (wants == 2 && rt->split_prefilter) {
        // scratch columns of the pre-filter kernel: words per agent = masks + row table + header + 64-bit map; when the
        // memory is not to be had the launcher stays with the in-kernel pre-filter
        const size_t per = (size_t)kPfWords + kPfRows + 1 + 2, need = (size_t)a.self.n * per;
        if (s.pf_cap < need) {
          if (s.pf_buf) TRY(release_device(rt, s.pf_buf));
          s.pf_buf = nullptr;
          s.pf_cap = need + need / 16;
          if (cudaMalloc(&s.pf_buf, s.pf_cap * sizeof(u32)) != cudaSuccess) { cudaGetLastError(); s.pf_buf = nullptr; s.pf_cap = 0; }
        }
        if (s.pf_buf) {
          const size_t n = a.self.n;
          a.pf_sbits = (unsigned long long *)s.pf_buf;          // the 8-byte aligned part first
          a.pf_masks = s.pf_buf + 2 * n;
          a.pf_rows = a.pf_masks + (size_t)kPfWords * n;
          a.pf_hdr = a.pf_rows + (size_t)kPfRows * n;
          a.pf_stride = (unsigned)n;
        }
      }
    }
    a.pdl = (rt->pdl ? (rt->pdl_trigger ? 3 : 1) : 0) | (rt->mbar_hint ? 4 : 0);
    a.stream = (void *)rt->stream;
    // cached neighbour lists: neither pool of this step's for-near loop ever moves (the code
    // generator's guarantee), so the accepted candidates are found once and walked afterwards
    if (s.desc.nlist && rt->nlist && nbr && !rt->slab && !s.nl.off) {
      NeighbourLists &nl = s.nl;
      const bool current = nl.valid && nl.self_serial == self.order_serial && nl.nbr_serial == nbr->order_serial &&
                           nl.n_self == self.n && nl.n_nbr == nbr->n && nl.stride == a.self.n;
      if (!current) TRY(build_neighbour_lists(rt, s, a, self, *nbr));
      if (nl.valid) {
        a.nlist_cnt = nl.cnt;
        a.nlist_idx = nl.idx;
        a.nlist_stride = nl.stride;
      }
    }
    trace_stamp(rt, TR_OTHER);
    // the exchange of this step next to its kernel (see halo_async): fork point
    // (a column the step does not write keeps ONE buffer: the exchange would store arrivals behind the owned range
    // of the very array whose ghost records the neighbour loop may still be reading — only forked when the loop
    // reads another pool or every member is rewritten, i.e. arrivals land in output buffers)
    const unsigned all_members = self.members.size() >= 32 ? 0xffffffffu : ((1u << self.members.size()) - 1u);
    const bool arrivals_hit_outputs = nbr != &self || (s.desc.written_members & all_members) == all_members;
    // (stage timing keeps the fork: the commit interval then shows what the step really waits for — the join)
    const bool fork_exchange = direct && !use_dev_range && rt->halo_async && !rt->trace &&
                               a.slab.active && a.slab.boundary_first && a.self.n && arrivals_hit_outputs;
    if (fork_exchange) {
      if (!rt->side_stream) {
        int lo_prio = 0, hi_prio = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CU(cudaStreamCreateWithPriority(&rt->side_stream, cudaStreamNonBlocking, hi_prio));
        CU(cudaEventCreateWithFlags(&rt->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&rt->ev_join, cudaEventDisableTiming));
      }
      CU(cudaEventRecord(rt->ev_fork, rt->stream));
    }
    if (rt->replay_reps > 0) {
      // abl_cuda_time_kernel: the kernel's own duration, undisturbed by events between the kernels of the chain.
      // The launch is repeated on the same input buffers (a step kernel reads `in` and writes `out`: every
      // repetition stores the same values), one untimed launch first, then `reps` between two events; the fused
      // histogram the repetitions add to is cleared again before the launch that counts.
      const int reps = rt->replay_reps;
      rt->replay_reps = 0;
      rt->replay_ms = 0.f;
      if (a.self.n && !rt->slab && !mutating) {
        cudaEvent_t e0, e1;
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        int rrc = s.desc.launch(&a);
        CU(cudaEventRecord(e0, rt->stream));
        for (int r = 0; r < reps && rrc == 0; r++) rrc = s.desc.launch(&a);
        CU(cudaEventRecord(e1, rt->stream));
        if (a.bin_count) CU(cudaMemsetAsync(self.cell_count, 0, ((size_t)rt->grid.n_local + 2) * sizeof(u32), rt->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (rrc != 0) return fail(ABL_ERR_CUDA, "step %s: kernel launch failed: %s", s.name.c_str(), cudaGetErrorString((cudaError_t)rrc));
        rt->replay_ms = ms / (float)reps;
        rt->launches += (unsigned)(reps + 1) * (a.pf_masks ? 2 : 1);
      }
    }
    int rc = a.self.n ? s.desc.launch(&a) : 0;
    rt->launches += a.pf_masks ? 2 : 1;   // (ABL_MODE 9: the pre-filter kernel and the step kernel)
    trace_stamp(rt, TR_STEP);
    if (rc != 0) return fail(ABL_ERR_CUDA, "step %s: kernel launch failed: %s", s.name.c_str(),
                             cudaGetErrorString((cudaError_t)rc));
    if (rt->timing) CU(cudaEventRecord(rt->ev[2], rt->stream));

    // commit: flip written columns
    for (size_t m = 0; m < self.members.size(); m++) {
      if (!(s.desc.written_members >> m & 1u)) continue;
      const Member &mem = self.members[m];
      for (int c = mem.first_col; c < mem.first_col + mem.ncols; c++) self.cols[c].cur ^= 1;
      if ((int)m == self.pos_member) self.binned = false;
    }
    if (fuse) self.counted = true;
    if (slab_pool && mutating) {
      const u32 n_own = self.own_end - self.own_begin;
      // (the caller resolves the global ids, then abl_cuda_resolve_adds removes, appends and exchanges)
      if (added) return slab_open_adds(rt, step, self, self.own_begin, n_own, staging);
      TRY(commit_removals(rt, self, self.own_begin, n_own, true));
      return slab_after_mutation(rt, s.desc.self_pool);
    }
    const size_t n_stepped = self.n;
    if (added) TRY(commit_adds(rt, self, *added, staging));
    if (s.desc.uses_removal) {
      // agents the step appended to its own pool carry no removal flag of their own
      if (self.n > n_stepped) CU(cudaMemsetAsync(self.dead + n_stepped, 0, self.n - n_stepped, rt->stream));
      TRY(commit_removals(rt, self, 0, (u32)self.n, false));
    }
    // slab mode: ghosts of this pool are stale (and agents may have left the slab)
    if (direct) {
      TRY(halo_finish(rt, self, !use_dev_range && a.slab.boundary_first && a.self.n, use_dev_range, fork_exchange));
      trace_stamp(rt, TR_EXCHANGE);
    }
    else if (rt->slab && !rt->peer_lo && !rt->peer_hi && s.desc.written_members && self.pos_member >= 0)
      TRY(abl_cuda_exchange(rt, s.desc.self_pool));
  } else {
    if (rt->timing) CU(cudaEventRecord(rt->ev[2], rt->stream));
    if (slab_pool && mutating) {  // an empty slab still takes part in the id resolution and the exchange
      void *none[ABL_MAX_COLUMNS + 1];
      memset(none, 0, sizeof none);
      if (added) return slab_open_adds(rt, step, self, 0, 0, none);
      return slab_after_mutation(rt, s.desc.self_pool);
    }
    if (direct) {  // an empty slab still has to answer its neighbours
      TRY(halo_reserve(rt, self));
      TRY(halo_finish(rt, self, false, use_dev_range));
    }
  }
  if (rt->timing) {
    CU(cudaEventRecord(rt->ev[3], rt->stream));
    for (int i = 0; i < 4; i++) {
      rt->timing_log.push_back(rt->ev[i]);
      CU(cudaEventCreate(&rt->ev[i]));
    }
    if (rt->timing_log.size() >= 4 * 2048) TRY(drain_timing(rt));   // bounded number of live events
  }
  return ABL_OK;
}

extern "C" int abl_cuda_begin_timestep(abl_runtime *rt) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  CU(cudaEventRecord(rt->ev_ts[0], rt->stream));
  rt->ts_open = true;
  return ABL_OK;
}

extern "C" int abl_cuda_end_timestep(abl_runtime *rt) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  rt->timestep++;
  if (rt->ts_open) {
    CU(cudaEventRecord(rt->ev_ts[1], rt->stream));
  }
  return ABL_OK;
}

extern "C" int abl_cuda_last_exec_time(abl_runtime *rt, double *seconds) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  // time of the parallel part of the current timestep so far (the sequential step runs
  // after all step functions, reference MasonPrinter.cpp:535-541)
  float ms = 0;
  if (rt->ts_open) {
    CU(cudaEventRecord(rt->ev_ts[1], rt->stream));
    CU(cudaEventSynchronize(rt->ev_ts[1]));
    CU(cudaEventElapsedTime(&ms, rt->ev_ts[0], rt->ev_ts[1]));
  }
  if (seconds) *seconds = ms / 1000.0;
  return ABL_OK;
}

extern "C" int abl_cuda_time_kernel(abl_runtime *rt, int step, int reps, float *ms_per_launch) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  if (reps < 1) return fail(ABL_ERR_ARGUMENT, "time_kernel: reps must be positive");
  rt->replay_reps = reps;
  rt->replay_ms = 0.f;
  int rc = abl_cuda_step(rt, step);
  rt->replay_reps = 0;
  if (ms_per_launch) *ms_per_launch = rt->replay_ms;
  return rc;
}

extern "C" int abl_cuda_enable_timing(abl_runtime *rt, int on) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  TRY(drain_timing(rt));
  rt->timing = on != 0;
  return ABL_OK;
}

extern "C" int abl_cuda_last_timing(abl_runtime *rt, abl_step_timing *t) {
  if (!rt || !t) return fail(ABL_ERR_ARGUMENT, "null argument");
  TRY(drain_timing(rt));
  *t = rt->last;
  t->launches = rt->launches;
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------
extern "C" int abl_cuda_set_reduce_hook(abl_runtime *rt, abl_reduce_hook hook, void *user) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  rt->reduce_hook = hook;
  rt->reduce_user = user;
  return ABL_OK;
}

// Under slab decomposition a reduction covers the owned agents of this runtime; the hook (if
// the harness installed one) turns the rank-local value into the global one.
static int combine_int(abl_runtime *rt, int *v) {
  if (!rt->slab || !rt->reduce_hook || !v) return ABL_OK;
  long long x = *v;
  if (rt->reduce_hook(rt->reduce_user, &x, 1, nullptr, 0) != 0) return fail(ABL_ERR_COMM, "reduction hook failed");
  *v = (int)x;
  return ABL_OK;
}
static int combine_real(abl_runtime *rt, double *v) {
  if (!rt->slab || !rt->reduce_hook || !v) return ABL_OK;
  if (rt->reduce_hook(rt->reduce_user, nullptr, 0, v, 1) != 0) return fail(ABL_ERR_COMM, "reduction hook failed");
  return ABL_OK;
}

extern "C" int abl_cuda_count(abl_runtime *rt, int pool, int *result) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  size_t n = 0;
  TRY(abl_cuda_pool_size(rt, pool, &n));  // owned agents only under slab decomposition
  int r = (int)n;
  TRY(combine_int(rt, &r));
  if (result) *result = r;
  return ABL_OK;
}

// Range of live records a reduction runs over (rank-local under slab decomposition: the
// caller combines the per-rank values).
static int reduce_range(abl_runtime *rt, Pool &p, u32 *first, u32 *n) {
  if (rt->slab && p.pos_member >= 0) {
    TRY(slab_settle(rt, p));
    *first = p.own_begin;
    *n = p.own_end - p.own_begin;
  } else {
    *first = p.src_begin;
    *n = (u32)p.n;
  }
  return ABL_OK;
}

static int get_member(abl_runtime *rt, int pool, int member, Pool **p, Member **m) {
  TRY(get_pool(rt, pool, p));
  if (member < 0 || member >= (int)(*p)->members.size()) return fail(ABL_ERR_ARGUMENT, "bad member index %d", member);
  *m = &(*p)->members[member];
  return ABL_OK;
}

static int reduce_int(abl_runtime *rt, Pool &p, Member &m, int kind, int value, int *result) {
  int *d = (int *)rt->d_scalar;
  CU(cudaMemsetAsync(d, 0, sizeof(int), rt->stream));
  u32 first = 0, n = 0;
  TRY(reduce_range(rt, p, &first, &n));
  const void *col = (const u8 *)p.cols[m.first_col].buf[p.cols[m.first_col].cur] + (size_t)first * p.cols[m.first_col].elem;
  if (n) {
    u32 nb = std::min(blocks_for(n, 256), 148u * 8u);
    switch (kind) {
      case 0: k_reduce_int<0><<<nb, 256, 0, rt->stream>>>(col, n, value, d); break;
      case 1: k_reduce_int<1><<<nb, 256, 0, rt->stream>>>(col, n, value, d); break;
      case 2: k_reduce_int<2><<<nb, 256, 0, rt->stream>>>(col, n, value, d); break;
      default: k_reduce_int<3><<<nb, 256, 0, rt->stream>>>(col, n, value, d); break;
    }
    rt->launches++;
    CU(cudaGetLastError());
  }
  u32 out = 0;
  TRY(read_scalar(rt, rt->d_scalar, &out));
  int r = (int)out;
  TRY(combine_int(rt, &r));
  if (result) *result = r;
  return ABL_OK;
}

extern "C" int abl_cuda_sum_int(abl_runtime *rt, int pool, int member, int *result) {
  Pool *p = nullptr; Member *m = nullptr;
  TRY(get_member(rt, pool, member, &p, &m));
  if (m->type == ABL_TYPE_INT) return reduce_int(rt, *p, *m, 0, 0, result);
  if (m->type == ABL_TYPE_BOOL) return reduce_int(rt, *p, *m, 1, 0, result);
  return fail(ABL_ERR_ARGUMENT, "sum_int on non-integer member %s", m->name.c_str());
}

extern "C" int abl_cuda_count_member_int(abl_runtime *rt, int pool, int member, int value, int *result) {
  Pool *p = nullptr; Member *m = nullptr;
  TRY(get_member(rt, pool, member, &p, &m));
  if (m->type == ABL_TYPE_INT) return reduce_int(rt, *p, *m, 2, value, result);
  if (m->type == ABL_TYPE_BOOL) return reduce_int(rt, *p, *m, 3, value ? 1 : 0, result);
  return fail(ABL_ERR_ARGUMENT, "count_member_int on non-integer member %s", m->name.c_str());
}

extern "C" int abl_cuda_count_member_float(abl_runtime *rt, int pool, int member, double value, int *result) {
  Pool *p = nullptr; Member *m = nullptr;
  TRY(get_member(rt, pool, member, &p, &m));
  if (m->type != ABL_TYPE_FLOAT) return fail(ABL_ERR_ARGUMENT, "count_member_float on non-float member");
  int *d = (int *)rt->d_scalar;
  CU(cudaMemsetAsync(d, 0, sizeof(int), rt->stream));
  u32 first = 0, n = 0;
  TRY(reduce_range(rt, *p, &first, &n));
  const void *col = (const u8 *)p->cols[m->first_col].buf[p->cols[m->first_col].cur] + (size_t)first * p->cols[m->first_col].elem;
  if (n) {
    u32 nb = std::min(blocks_for(n, 256), 148u * 8u);
    if (rt->real_size == 8) k_count_real<double><<<nb, 256, 0, rt->stream>>>((const double *)col, n, value, d);
    else k_count_real<float><<<nb, 256, 0, rt->stream>>>((const float *)col, n, (float)value, d);
    rt->launches++;
    CU(cudaGetLastError());
  }
  u32 out = 0;
  TRY(read_scalar(rt, rt->d_scalar, &out));
  int r = (int)out;
  TRY(combine_int(rt, &r));
  if (result) *result = r;
  return ABL_OK;
}

extern "C" int abl_cuda_sum_float(abl_runtime *rt, int pool, int member, int component, double *result) {
  Pool *p = nullptr; Member *m = nullptr;
  TRY(get_member(rt, pool, member, &p, &m));
  int col_index = m->first_col, stride = 1, comp = 0;
  if (m->type == ABL_TYPE_FLOAT2) { stride = 2; comp = component; }
  else if (m->type == ABL_TYPE_FLOAT3) { col_index += component; }
  else if (m->type != ABL_TYPE_FLOAT) return fail(ABL_ERR_ARGUMENT, "sum_float on non-float member");
  if (comp < 0 || comp > 1 || component < 0 || component > 2) return fail(ABL_ERR_ARGUMENT, "bad component");
  u32 first = 0, n = 0;
  TRY(reduce_range(rt, *p, &first, &n));
  const void *col = (const u8 *)p->cols[col_index].buf[p->cols[col_index].cur] + (size_t)first * p->cols[col_index].elem;
  const u32 nb = 148 * 2;
  double *partial = (double *)(rt->d_scalar + 16);  // 8-byte aligned region inside scratch
  double *d_out = (double *)(rt->d_scalar + 2);
  if (rt->real_size == 8) k_reduce_real_partial<double><<<nb, 256, 0, rt->stream>>>((const double *)col, stride, comp, n, partial);
  else k_reduce_real_partial<float><<<nb, 256, 0, rt->stream>>>((const float *)col, stride, comp, n, partial);
  k_final_sum<<<1, 32, 0, rt->stream>>>(partial, (int)nb, d_out);
  rt->launches += 2;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(rt->h_scalar, d_out, sizeof(double), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  double r;
  memcpy(&r, rt->h_scalar, sizeof(double));
  TRY(combine_real(rt, &r));
  if (result) *result = r;
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// multi-GPU: slab decomposition with halo + migration exchange over NCCL
// ---------------------------------------------------------------------------------------
// The grid is cut into slabs of whole cell layers along the slowest-varying axis (y in 2D,
// z in 3D).  Rank r owns the agents whose cell layer lies in [layer_begin, layer_end) and
// additionally holds read-only ghost copies of the `ghost_layers` adjacent layers on each
// side.  Because the pool is kept sorted by global cell key, the array always looks like
//     [ ghosts below | owned | ghosts above ]
// and ownership is nothing but a key range: own_begin = cell_start[layer_begin * row],
// own_end = cell_start[layer_end * row].  Step kernels run over the owned range only and see
// ghosts through the ordinary neighbour search, in the same (cell key, id) order as a
// single-GPU run, so results are bit-identical for any number of GPUs.
//
// After a step function has written members of a pool its ghosts are stale.  exchange():
//   1. classify every formerly owned agent by the layer of its (new) position:
//      layer <  layer_begin + G  -> copy goes to rank-1      (G = ghost_layers)
//      layer >= layer_end   - G  -> copy goes to rank+1
//      (this covers both migration and halo refresh: the receiver decides by key range
//      whether an arrival is owned or a ghost; the sender keeps its record, which turns
//      into a ghost by the same rule if the agent left the slab)
//   2. pack the selected records column-wise (warp-scan compaction), swap counts, then
//      payloads, with grouped ncclSend/ncclRecv on the runtime's stream (NVLink)
//   3. drop the old ghosts (the owned range is contiguous), append the arrivals and let the
//      next binning sort everything back into [ghosts | owned | ghosts].
static int slab_bin_if_needed(abl_runtime *rt, Pool &p) {
  TRY(slab_settle(rt, p));
  return ABL_OK;
}

static int slab_row_cells(const abl_runtime *rt) {
  const GridParams &g = rt->grid;
  return g.dim == 3 ? g.n_cell[0] * g.n_cell[1] : g.n_cell[0];
}
static int slab_layers(const abl_runtime *rt) {
  return rt->grid.dim == 3 ? rt->grid.n_cell[2] : rt->grid.n_cell[1];
}

static int slab_crop_to_owned(abl_runtime *rt, int pool) {
  Pool &p = rt->pools[pool];
  if (p.pos_member < 0) return ABL_OK;
  TRY(bin_pool(rt, p));
  p.src_begin = p.own_begin;
  p.n = p.own_end - p.own_begin;
  p.binned = false;  // the next binning compacts the owned records to the front
  return ABL_OK;
}

// The owned range is two words of cell_start.  They are copied to the host right after the
// scan and awaited only after the remaining binning kernels have been enqueued, so the host
// learns them while the GPU is still busy and can enqueue the step kernel without a gap.
__global__ void k_gather_bin_words(const u32 *cell_start, u32 lo_cell, u32 hi_cell, u32 lo2_cell, u32 hi2_cell,
                                   const u32 *halo_ctr, u32 *out, u32 stamp) {
  // `out` is page-locked host memory mapped into the device address space: the words reach the
  // host without a copy-engine operation between two kernels of the stream
  if (threadIdx.x == 0) {
    out[0] = cell_start[lo_cell];
    out[1] = cell_start[hi_cell];
    for (int k = 0; k < 6; k++) out[2 + k] = halo_ctr ? halo_ctr[2 + k] : 0;  // in lo, in hi, timeout, far, sent lo, sent hi
    out[8] = cell_start[lo2_cell];
    out[9] = cell_start[hi2_cell];
    out[10] = halo_ctr ? halo_ctr[12] : 0;  // late
    out[11] = halo_ctr ? halo_ctr[14] : 0;
    out[12] = halo_ctr ? 1u : 0u;
    __threadfence_system();
    for (int k = 16; k <= 20; k++) *(volatile u32 *)(out + k) = stamp;
  }
}

static void slab_owned_cells(const abl_runtime *rt, u32 *lo_cell, u32 *hi_cell) {
  const int row = slab_row_cells(rt);
  *lo_cell = (u32)rt->layer_begin * (u32)row - rt->grid.key_base;
  *hi_cell = (u32)rt->layer_end * (u32)row - rt->grid.key_base;
}
// Boundary parts for boundary-first scheduling: the 2 * ghost outermost owned layers on each
// side (an agent further inside cannot reach the halo zone by moving less than `ghost` layers).
// Slabs too thin for an interior are all boundary.
static void slab_boundary_cells(const abl_runtime *rt, u32 *lo2_cell, u32 *hi2_cell) {
  const int row = slab_row_cells(rt);
  const int w = 2 * rt->ghost_layers;
  int lo2 = rt->layer_begin + w, hi2 = rt->layer_end - w;
  if (lo2 >= hi2) lo2 = hi2 = rt->layer_end;   // everything belongs to the lower boundary part
  *lo2_cell = (u32)lo2 * (u32)row - rt->grid.key_base;
  *hi2_cell = (u32)hi2 * (u32)row - rt->grid.key_base;
}

// `reported`: the scan kernel has already written the words (two-pass scan); otherwise a
// one-thread kernel gathers them
static int slab_request_owned_range(abl_runtime *rt, Pool &p, bool reported, const ScanReport &rep) {
  if (!reported) {
    k_gather_bin_words<<<1, 32, 0, rt->stream>>>(p.cell_start, rep.lo_cell, rep.hi_cell, rep.lo2_cell, rep.hi2_cell,
                                                 rep.halo_ctr, rep.host, rep.stamp);
    rt->launches++;
  }
  return ABL_OK;
}

// Waits for the report of the pool's last binning (normally long there) and brings the host's
// view up to date: owned range, boundary parts, and — if a direct exchange was queued with a
// device-side range — the extent the next binning has to cover.  The counters of the halo
// exchange that preceded the reported binning are checked here.  `redo` (synchronous callers
// only): set when more records arrived than the padding covered, so that the caller can bin
// again; without it that condition is an error.
static int slab_consume_report(abl_runtime *rt, Pool &p, bool *redo) {
  if (redo) *redo = false;
  if (!p.report_pending) return ABL_OK;
  p.report_pending = false;
  {
    // poll the stamps the scan kernel writes into this (mapped, page-locked) memory
    const volatile u32 *hv = report_host(rt, p);
    const u32 stamp = p.report_stamp;
    auto t0 = std::chrono::steady_clock::now();
    unsigned long spins = 0;
    for (;;) {
      if (hv[16] == stamp && hv[17] == stamp && hv[18] == stamp && hv[19] == stamp && hv[20] == stamp) break;
      if ((++spins & 0xfffu) == 0) {
        // a failed kernel never writes its stamps: surface the error instead of spinning
        cudaError_t e = cudaStreamQuery(rt->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) CU(e);
        if (e == cudaSuccess && !(hv[16] == stamp && hv[17] == stamp && hv[18] == stamp && hv[19] == stamp && hv[20] == stamp))
          return fail(ABL_ERR_STATE, "pool %s: the binning finished without reporting the owned range", p.name.c_str());
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
          return fail(ABL_ERR_COMM, "pool %s: no owned-range report from the device after 120 s", p.name.c_str());
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (rt->trace) {
      // how long the host waits here tells who is ahead: ~0 means the GPU is waiting for the host
      rt->own_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      rt->own_waits++;
    }
  }
  const u32 *h = report_host(rt, p);
  p.own_begin = h[0];
  p.own_end = h[1];
  p.bnd_lo_end = h[8];
  p.bnd_hi_begin = h[9];
  p.own_valid = true;
  if (h[12]) {
    // the report carries the counters of halo exchange number h[11]
    const u32 seq = h[11];
    const u32 in_lo = h[2], in_hi = h[3], flags = h[4], far = h[5], sent_lo = h[6], sent_hi = h[7];
    if (flags & 1u) return fail(ABL_ERR_COMM, "pool %s: timed out waiting for a neighbour's halo message", p.name.c_str());
    if (h[10])
      return fail(ABL_ERR_COMM, "%u agents of pool %s crossed more than %d cell layers in one step and reached the halo "
                  "zone after the messages had been published; set ABL_CUDA_HALO_OVERLAP=0", h[10], p.name.c_str(),
                  rt->ghost_layers);
    if (far)
      return fail(ABL_ERR_COMM, "%u agents of pool %s moved farther than a neighbouring slab in one step "
                  "(only neighbour and periodic wrap-around migration is supported)", far, p.name.c_str());
    if (in_lo > p.halo_cap || in_hi > p.halo_cap || sent_lo > p.halo_cap || sent_hi > p.halo_cap)
      return fail(ABL_ERR_COMM, "pool %s: halo message of %u records exceeds the capacity of %zu; raise the "
                  "capacity passed to abl_cuda_halo_setup", p.name.c_str(), std::max(std::max(in_lo, in_hi), std::max(sent_lo, sent_hi)), p.halo_cap);
    const u32 arrivals = in_lo + in_hi;
    if (flags & 2u)
      return fail(ABL_ERR_COMM, "pool %s: %u halo records arrived but the pool had no room for them", p.name.c_str(), arrivals);
    const u32 pad_used = p.pad_used[seq & 1u];
    // adapt the padding to what actually arrives: the counts of an exchange are verified one
    // binning later, so the head room has to absorb the growth of two steps
    p.halo_pad = (u32)round_up(2 * (size_t)arrivals + 1024, 256);
    p.halo_last_arrivals = arrivals;
    if (arrivals > pad_used) {
      if (redo) *redo = true;
      else
        return fail(ABL_ERR_COMM, "pool %s: %u halo records arrived in exchange %u but the binning was sized for %u "
                    "(arrivals more than doubled within two steps)", p.name.c_str(), arrivals, seq, pad_used);
    }
  }
  if (p.exch_deferred) {
    // a direct exchange was queued behind the reported binning: old ghosts are dead, arrivals and
    // padding sit behind the owned range
    p.exch_deferred = false;
    p.halo_prev_ob = p.own_begin;
    p.halo_prev_own = p.own_end - p.own_begin;
    p.src_begin = p.own_begin;
    p.n = (size_t)(p.own_end - p.own_begin) + p.exch_pad;
  }
  return ABL_OK;
}

// the host needs the owned range of a pool (downloads, reductions, staged / NCCL exchange)
static int slab_settle(abl_runtime *rt, Pool &p) {
  if (!rt->slab || p.pos_member < 0) return ABL_OK;
  if (p.report_pending && !p.exch_deferred) TRY(slab_consume_report(rt, p, nullptr));
  if (!p.binned) TRY(bin_pool(rt, p));
  return ABL_OK;
}

struct SlabTable {
  int n_slabs, my;
  int bounds[ABL_MAX_SLABS + 1];
  int ghost;
};

// One pass over the formerly owned agents: classify by the layer of the current position and
// append the selected records to the outgoing messages (atomic slot allocation; the order
// inside a message is irrelevant because the receiver's next binning sorts by (cell, id)).
// The record is copied to the lower / upper peer.  Peers form a ring when there are more than
// two slabs, so that agents leaving through one end of a periodic world (wraparound / teleport
// to the opposite bound) reach the slab at the other end.
// Message layout: u32 count (16-byte header), then packed records of `rec_words` 32-bit words:
// the columns of one agent back to back (1-byte columns widened to a word), id last.
static const u32 kMsgHeader = 16;
static const u32 kMsgFirstRecords = 8192;   // records that fit the fixed-size first message

template <typename R>
__global__ void k_slab_pack(ColTable t, const R *axis_col, int stride, int comp, u32 n, u32 first,
                            R origin, R inv_cell, int n_layers, SlabTable tab, bool has_lo, bool has_hi,
                            u8 *msg_lo, u8 *msg_hi, u32 rec_words, u32 *far) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t src = (size_t)first + i;
  R v = axis_col[src * stride + comp];
  const int layer = cell_coord<R>(v, origin, inv_cell, n_layers);
  const int me = tab.my, N = tab.n_slabs;
  const int lb = tab.bounds[me], le = tab.bounds[me + 1];
  int owner = 0;
  while (owner + 1 < N && layer >= tab.bounds[owner + 1]) owner++;
  bool lo = false, hi = false;
  if (owner == me) {
    // still mine: neighbours need a ghost copy of my boundary layers
    lo = me > 0 && layer < lb + tab.ghost;
    hi = me < N - 1 && layer >= le - tab.ghost;
  } else if (owner == me - 1) {
    lo = true;
  } else if (owner == me + 1) {
    hi = true;
  } else if (me == 0 && owner == N - 1) {
    lo = true;   // around the ring
  } else if (me == N - 1 && owner == 0) {
    hi = true;
  } else {
    atomicAdd(far, 1u);  // moved farther than a neighbouring slab: unsupported
  }
  lo = lo && has_lo;
  hi = hi && has_hi;
  for (int dir = 0; dir < 2; dir++) {
    if (!(dir == 0 ? lo : hi)) continue;
    u8 *msg = dir == 0 ? msg_lo : msg_hi;
    u32 slot = atomicAdd((u32 *)msg, 1u);
    u32 *rec = (u32 *)(msg + kMsgHeader) + (size_t)slot * rec_words;
    u32 w = 0;
    for (int k = 0; k < t.ncols; k++) {
      const int e = t.elem[k];
      if (e == 1) { rec[w++] = ((const u8 *)t.in[k])[src]; continue; }
      const u32 *p = (const u32 *)t.in[k] + src * (e / 4);
      for (int q = 0; q < e / 4; q++) rec[w++] = p[q];
    }
  }
}

__global__ void k_slab_unpack(ColTable t, const u8 *msg, u32 count, u32 dst_first, u32 rec_words) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const u32 *rec = (const u32 *)(msg + kMsgHeader) + (size_t)i * rec_words;
  const size_t dst = (size_t)dst_first + i;
  u32 w = 0;
  for (int k = 0; k < t.ncols; k++) {
    const int e = t.elem[k];
    if (e == 1) { ((u8 *)t.out[k])[dst] = (u8)rec[w++]; continue; }
    u32 *p = (u32 *)t.out[k] + dst * (e / 4);
    for (int q = 0; q < e / 4; q++) p[q] = rec[w++];
  }
}

__global__ void k_reset_words(u32 *a, u32 *b, u32 *c) {
  if (threadIdx.x == 0) { *a = 0; *b = 0; *c = 0; }
}
__global__ void k_gather_words(const u32 *a, const u32 *b, const u32 *c, const u32 *d, const u32 *e, u32 *out) {
  if (threadIdx.x == 0) {
    out[0] = a ? *a : 0; out[1] = b ? *b : 0; out[2] = c ? *c : 0; out[3] = d ? *d : 0; out[4] = e ? *e : 0;
  }
}

static u32 slab_rec_words(const Pool &p) {
  u32 w = 0;
  for (const Column &c : p.cols) w += c.elem == 1 ? 1 : c.elem / 4;
  return w;
}
static size_t slab_msg_bytes(const Pool &p, size_t count) {
  return kMsgHeader + count * (size_t)slab_rec_words(p) * 4;
}

static int ensure_xbuf(abl_runtime *rt, int which, size_t bytes) {
  if (rt->xcap[which] >= bytes) return ABL_OK;
  CU(cudaStreamSynchronize(rt->stream));
  if (rt->xbuf[which]) CU(cudaFree(rt->xbuf[which]));
  size_t cap = round_up(bytes + bytes / 4 + 4096, 1 << 16);
  CU(cudaMalloc(&rt->xbuf[which], cap));
  rt->xcap[which] = cap;
  return ABL_OK;
}

#define NCCL(call)                                                                          \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess)                                                                  \
      return fail(ABL_ERR_COMM, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r_),     \
                  __FILE__, __LINE__);                                                      \
  } while (0)

extern "C" int abl_cuda_nccl_unique_id(void *id128) {
  if (!id128) return fail(ABL_ERR_ARGUMENT, "null id buffer");
  ncclUniqueId id;
  NCCL(ncclGetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  memcpy(id128, &id, sizeof id);
  return ABL_OK;
}

extern "C" int abl_cuda_comm_init_nccl(abl_runtime *rt, const void *id128, int rank, int world) {
  if (!rt || !id128) return fail(ABL_ERR_ARGUMENT, "null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail(ABL_ERR_ARGUMENT, "bad rank/world");
  CU(cudaSetDevice(rt->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NCCL(ncclCommInitRank(&rt->comm, world, id, rank));
  rt->rank = rank;
  rt->world = world;
  return ABL_OK;
}

extern "C" int abl_cuda_slab_axis_layers(abl_runtime *rt, int *n_layers) {
  if (!rt || !rt->env_set) return fail(ABL_ERR_STATE, "environment not set");
  if (n_layers) *n_layers = slab_layers(rt);
  return ABL_OK;
}

// Also usable without NCCL (world == 1 or comm == NULL): then exchange() only re-sorts, which
// is what the single-process tests of the slab bookkeeping use.
extern "C" int abl_cuda_set_slab(abl_runtime *rt, const int *layer_bounds, int n_slabs, int my_slab) {
  if (!rt || !rt->env_set) return fail(ABL_ERR_STATE, "set_slab requires an environment");
  if (!layer_bounds || n_slabs < 1 || n_slabs > ABL_MAX_SLABS || my_slab < 0 || my_slab >= n_slabs)
    return fail(ABL_ERR_ARGUMENT, "bad slab table");
  const int layers = slab_layers(rt);
  if (layer_bounds[0] != 0 || layer_bounds[n_slabs] != layers)
    return fail(ABL_ERR_ARGUMENT, "slab table must cover layers 0..%d", layers);
  for (int s = 0; s < n_slabs; s++)
    if (layer_bounds[s] >= layer_bounds[s + 1]) return fail(ABL_ERR_ARGUMENT, "empty slab %d", s);
  rt->slab = true;
  rt->n_slabs = n_slabs;
  rt->my_slab = my_slab;
  for (int s = 0; s <= n_slabs; s++) rt->slab_bounds[s] = layer_bounds[s];
  rt->layer_begin = layer_bounds[my_slab];
  rt->layer_end = layer_bounds[my_slab + 1];
  int g = 1;
  for (const Step &s : rt->steps) g = std::max(g, s.reach);
  rt->ghost_layers = g;
  // cell ranges are kept for the own slab plus ghost layers only: the cost of binning does not
  // grow with the size of the whole world
  GridParams &gp = rt->grid;
  gp.axis_lo = std::max(0, rt->layer_begin - g);
  gp.axis_hi = std::min(layers, rt->layer_end + g);
  gp.key_base = (u32)gp.axis_lo * (u32)slab_row_cells(rt);
  gp.n_local = (u32)(gp.axis_hi - gp.axis_lo) * (u32)slab_row_cells(rt);
  CU(cudaStreamSynchronize(rt->stream));
  for (Pool &p : rt->pools) {
    p.binned = false;
    p.own_valid = false;
    p.report_pending = false;
    p.exch_deferred = false;
    p.counted = false;
    if (p.cell_count) { CU(cudaFree(p.cell_count)); p.cell_count = nullptr; }
    if (p.cell_start) { CU(cudaFree(p.cell_start)); p.cell_start = nullptr; }
  }
  return ABL_OK;
}

extern "C" int abl_cuda_owned_size(abl_runtime *rt, int pool, size_t *n) {
  Pool *p = nullptr;
  TRY(get_pool(rt, pool, &p));
  if (rt->slab && p->pos_member >= 0) {
    TRY(slab_settle(rt, *p));
    if (n) *n = p->own_end - p->own_begin;
  } else if (n) {
    *n = p->n;
  }
  return ABL_OK;
}

// ---- exchange, phase 1: classify + pack into xbuf[0] (to lower) / xbuf[1] (to upper) ----------
static int exchange_pack(abl_runtime *rt, Pool &p, bool has_lo, bool has_hi) {
  // (the owned range of the binning the step ran on — not a fresh binning of the new positions)
  if (p.report_pending && !p.exch_deferred) TRY(slab_consume_report(rt, p, nullptr));
  if (!p.own_valid) TRY(bin_pool(rt, p));  // fresh upload: establish the owned range
  const u32 ob = p.own_begin, oe = p.own_end, n_own = oe - ob;
  const GridParams &g = rt->grid;
  const int axis = g.dim - 1;
  const Member &pm = p.members[p.pos_member];
  // worst case every owned agent is sent; the first message always has room for kMsgFirstRecords
  const size_t cap = std::max<size_t>(n_own, kMsgFirstRecords);
  TRY(ensure_xbuf(rt, 0, slab_msg_bytes(p, cap)));
  TRY(ensure_xbuf(rt, 1, slab_msg_bytes(p, cap)));
  u32 *far = rt->d_scalar + 10;
  k_reset_words<<<1, 32, 0, rt->stream>>>((u32 *)rt->xbuf[0], (u32 *)rt->xbuf[1], far);
  rt->launches++;
  if (n_own && (has_lo || has_hi)) {
    int col = pm.first_col, stride = 1, comp = 0;
    if (g.dim == 2) { stride = 2; comp = 1; } else { col += 2; }
    const void *axis_col = p.cols[col].buf[p.cols[col].cur];
    SlabTable tab;
    tab.n_slabs = rt->n_slabs;
    tab.my = rt->my_slab;
    tab.ghost = rt->ghost_layers;
    for (int s = 0; s <= rt->n_slabs; s++) tab.bounds[s] = rt->slab_bounds[s];
    ColTable t;
    fill_table(p, t, false);
    const u32 rw = slab_rec_words(p);
    u32 nb = blocks_for(n_own, 256);
    if (rt->real_size == 8)
      k_slab_pack<double><<<nb, 256, 0, rt->stream>>>(t, (const double *)axis_col, stride, comp, n_own, ob,
          g.origin[axis], g.inv_cell, slab_layers(rt), tab, has_lo, has_hi, (u8 *)rt->xbuf[0], (u8 *)rt->xbuf[1], rw, far);
    else
      k_slab_pack<float><<<nb, 256, 0, rt->stream>>>(t, (const float *)axis_col, stride, comp, n_own, ob,
          (float)g.origin[axis], (float)g.inv_cell, slab_layers(rt), tab, has_lo, has_hi, (u8 *)rt->xbuf[0], (u8 *)rt->xbuf[1], rw, far);
    rt->launches++;
    CU(cudaGetLastError());
  }
  return ABL_OK;
}

// ---- exchange, phase 3: arrivals (messages in xbuf[2/3]) are appended behind the owned range --
static int exchange_unpack(abl_runtime *rt, Pool &p, const u32 incoming[2]) {
  const u32 ob = p.own_begin, oe = p.own_end, n_own = oe - ob;
  const u32 arrivals = incoming[0] + incoming[1];
  if ((size_t)oe + arrivals > p.cap) {
    p.n = std::max(p.n, (size_t)oe);
    TRY(drop_fused_histogram(rt, p));  // growing reallocates the key/rank scratch
    TRY(reserve_pool(rt, p, (size_t)oe + arrivals));
  }
  ColTable t;
  fill_table(p, t, false);
  for (int c = 0; c < t.ncols; c++) t.out[c] = const_cast<void *>(t.in[c]);
  const u32 rw = slab_rec_words(p);
  if (incoming[0]) {
    k_slab_unpack<<<blocks_for(incoming[0], 256), 256, 0, rt->stream>>>(t, (const u8 *)rt->xbuf[2], incoming[0], oe, rw);
    rt->launches++;
  }
  if (incoming[1]) {
    k_slab_unpack<<<blocks_for(incoming[1], 256), 256, 0, rt->stream>>>(t, (const u8 *)rt->xbuf[3], incoming[1], oe + incoming[0], rw);
    rt->launches++;
  }
  CU(cudaGetLastError());
  // old ghosts (below ob, above oe) are dead; the next binning compacts [ob, oe + arrivals)
  p.src_begin = ob;
  p.n = (size_t)n_own + arrivals;
  p.binned = false;
  // the step kernel already produced keys and the histogram of the owned agents (fused
  // epilogue); only the arrivals are still missing
  if (p.counted && arrivals) TRY(launch_bin_count(rt, p, arrivals, oe, n_own));
  return ABL_OK;
}

static int exchange_check_far(abl_runtime *rt, const Pool &p, u32 far) {
  if (far)
    return fail(ABL_ERR_COMM, "%u agents of pool %s moved farther than a neighbouring slab in one step "
                "(only neighbour and periodic wrap-around migration is supported)", far, p.name.c_str());
  return ABL_OK;
}

// Grows a receive buffer while keeping the part of the message that has already arrived.
static int grow_recv(abl_runtime *rt, int which, size_t bytes, size_t keep) {
  if (rt->xcap[which] >= bytes) return ABL_OK;
  void *nb = nullptr;
  size_t cap = round_up(bytes + bytes / 4, 1 << 16);
  CU(cudaMalloc(&nb, cap));
  if (keep) CU(cudaMemcpyAsync(nb, rt->xbuf[which], keep, cudaMemcpyDeviceToDevice, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  if (rt->xbuf[which]) CU(cudaFree(rt->xbuf[which]));
  rt->xbuf[which] = nb;
  rt->xcap[which] = cap;
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// direct halo transport: records are written straight into the neighbour's memory
// ---------------------------------------------------------------------------------------
// Every rank owns a receive area of four blocks, [from lower | from upper] x [parity 0 | 1];
// a block is a 64-byte header {seq, count} followed by packed records.  The neighbours map the
// area (CUDA IPC across processes, plain pointers inside one process) and the *step kernel
// itself* appends the halo / migration records of exchange number `seq` to block
// [direction][seq & 1] over NVLink (abl_slab_epilogue).  One small kernel follows on the same
// stream (k_halo_exchange): block 0 fences and writes the header {count, seq} into the
// neighbours' blocks; every block spins until both own headers carry `seq` (bounded by a
// time-out); then the arrivals are appended behind the owned range, padded with sentinel
// records up to the host's estimate `pad`, and entered into the cell histogram.
// No NCCL call and no host synchronisation is involved: the host never learns the counts of
// the current exchange.  It bins `owned + pad` records (sentinels fall into a trash cell
// behind all real cells) and reads the true counts together with the owned range one binning
// later (slab_update_owned_range), where capacity overruns, far migrations and time-outs are
// reported and the padding is adapted.  Two parities suffice as flow control: a rank can only
// start exchange seq+2 after it has consumed seq+1 from both neighbours, which they publish
// after having consumed seq.
struct HaloHeader {
  u32 seq, count;
  u32 reserved[14];
};
static_assert(sizeof(HaloHeader) == ABL_MSG_HEADER, "header size");

static bool halo_direct(const Pool &p) { return p.halo_recv != nullptr; }

static u8 *halo_block_of(u8 *area, size_t block_bytes, int from_dir, u32 seq) {
  return area + ((size_t)from_dir * 2 + (seq & 1u)) * block_bytes;
}

// One kernel per exchange: block 0 publishes this rank's counts, every block waits for both
// neighbours' headers, then all blocks unpack (grid-stride) and — when the step kernel has
// already produced the histogram of the owned agents — add keys and arrival ranks of the
// arrivals and of the padding, so that the next binning starts with the scan.
// Arrivals go to pool slots [dst_first, dst_first + in_lo + in_hi); slots up to `pad` behind
// them receive the sentinel id.  More arrivals than the host expected are still stored (bounded
// by `room`); the next binning then notices and bins once more.
struct HaloExchangeArgs {
  u32 *ctr;
  HaloHeader *to_lo, *to_hi;
  const u8 *from_lo, *from_hi;
  u32 seq, cap;
  long long timeout_ns;
  u32 dst_first, pad, room, rec_words;
  // device-side range (dev_range != 0): dst_first / room / key_first are derived from the pool's
  // cell_start (owned range = [cs[lo_cell], cs[hi_cell])) instead of being passed by the host
  int dev_range;
  const u32 *cs;
  u32 lo_cell, hi_cell, pool_cap;
  int publish;     // 0: the step kernel has published already (boundary-first scheduling)
  int trace;       // accumulate the wait time of block 0 in ctr[8..9] (ns) and calls in ctr[10]
  // fused histogram of the arrivals (count != 0)
  int count;
  const void *px, *py, *pz;
  u32 key_first;   // slot of the first arrival in key/local
  u32 *key, *local, *cell_count;
};

template <typename R, int DIM>
__global__ void __launch_bounds__(256) k_halo_exchange(ColTable t, HaloExchangeArgs a, GridParams g) {
  __shared__ u32 s_in[2];
  if (a.dev_range) {
    const u32 ob = a.cs[a.lo_cell], oe = a.cs[a.hi_cell];
    a.dst_first = oe;
    a.key_first = oe - ob;
    a.room = a.pool_cap - oe;
  }
  unsigned long long trace_t0 = 0;
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
  // (publish == 2: the step kernel publishes unless it had no boundary block to do it)
  if (blockIdx.x == 0 && threadIdx.x == 0) a.ctr[14] = a.seq;
  if (blockIdx.x == 0 && threadIdx.x == 0 && (a.publish == 1 || (a.publish == 2 && a.ctr[13] != a.seq))) {
    const u32 c0 = a.ctr[0], c1 = a.ctr[1];
    a.ctr[6] = c0;
    a.ctr[7] = c1;
    a.ctr[0] = 0;
    a.ctr[1] = 0;
    __threadfence_system();  // the records (written by the preceding kernel) before the header
    if (a.to_lo) {
      *(volatile u32 *)&a.to_lo->count = min(c0, a.cap);
      __threadfence_system();
      *(volatile u32 *)&a.to_lo->seq = a.seq;
    }
    if (a.to_hi) {
      *(volatile u32 *)&a.to_hi->count = min(c1, a.cap);
      __threadfence_system();
      *(volatile u32 *)&a.to_hi->seq = a.seq;
    }
  }
  if (threadIdx.x < 2) {
    const int dir = threadIdx.x;
    const HaloHeader *h = (const HaloHeader *)(dir == 0 ? a.from_lo : a.from_hi);
    u32 count = 0;
    if (h) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (;;) {
        if (*(const volatile u32 *)&h->seq == a.seq) {
          __threadfence_system();
          count = *(const volatile u32 *)&h->count;
          break;
        }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if ((long long)(t1 - t0) > a.timeout_ns) { atomicOr(a.ctr + 4, 1u); break; }
        __nanosleep(100);
      }
    }
    s_in[dir] = count;
    if (blockIdx.x == 0) a.ctr[2 + dir] = count;
  }
  __syncthreads();
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    // ABL_CUDA_TRACE: time block 0 spent publishing and waiting for the neighbours, and calls
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicAdd((unsigned long long *)(a.ctr + 8), t1 - trace_t0);
    atomicAdd(a.ctr + 10, 1u);
  }
  const u32 in_lo = s_in[0], in_hi = s_in[1];
  u32 total = in_lo + in_hi;
  if (total > a.room && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.ctr + 4, 2u);   // no room: reported by the host
  if (total < a.pad) total = a.pad;
  if (total > a.room) total = a.room;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const size_t dst = (size_t)a.dst_first + i;
    if (i >= in_lo + in_hi) {
      ((u32 *)t.out[t.ncols - 1])[dst] = ABL_SENTINEL_ID;
    } else {
      // word-major message: word w of record r at w * cap + r (see abl_slab_send)
      const u32 *rec = i < in_lo ? (const u32 *)(a.from_lo + ABL_MSG_HEADER) + i
                                 : (const u32 *)(a.from_hi + ABL_MSG_HEADER) + (i - in_lo);
      size_t w = 0;
      for (int k = 0; k < t.ncols; k++) {
        const int e = t.elem[k];
        if (e == 1) { ((u8 *)t.out[k])[dst] = (u8)rec[w]; w += a.cap; continue; }
        u32 *q = (u32 *)t.out[k] + dst * (e / 4);
        for (int j = 0; j < e / 4; j++) { q[j] = rec[w]; w += a.cap; }
      }
    }
    // only the part the host will bin (`pad` records) enters the histogram; a surplus makes
    // the next binning start over anyway
    if (a.count && i < a.pad)
      bin_count_one<R, DIM>(a.px, a.py, a.pz, dst, (size_t)a.key_first + i, g, a.key, a.local, a.cell_count,
                            (const u32 *)t.out[t.ncols - 1]);
  }
}

// stand-alone classify + pack for exchanges that do not follow a step kernel (first exchange
// after an upload): same routing and record format as abl_slab_epilogue
template <typename R>
__global__ void k_halo_pack(ColTable t, abl_slab_view s, u32 n, u32 first) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t src = (size_t)first + i;
  R v;
  if (s.dim == 2) v = ((const R *)t.in[s.pos_col])[2 * src + 1];
  else v = ((const R *)t.in[s.pos_col + 2])[src];
  const int layer = cell_coord<R>(v, (R)s.origin, (R)s.inv_cell, s.n_layers);
  const unsigned route = abl_slab_route(s, layer);
  if (route & 4u) { atomicAdd(s.far, 1u); return; }
  for (int dir = 0; dir < 2; dir++) {
    if (!((route >> dir) & 1u)) continue;
    const u32 slot = atomicAdd(s.count[dir], 1u);
    if (slot >= s.capacity || !s.msg[dir]) continue;
    u32 *rec = (u32 *)s.msg[dir] + slot;   // word-major, see abl_slab_send
    size_t w = 0;
    for (int k = 0; k < t.ncols; k++) {
      const int e = t.elem[k];
      if (e == 1) { rec[w] = ((const u8 *)t.in[k])[src]; w += s.capacity; continue; }
      const u32 *q = (const u32 *)t.in[k] + src * (e / 4);
      for (int j = 0; j < e / 4; j++) { rec[w] = q[j]; w += s.capacity; }
    }
  }
}

// Describes exchange number halo_seq + 1 of `p` for the kernel that packs it.
static int halo_fill_view(abl_runtime *rt, Pool &p, abl_slab_view &v, bool step_kernel, bool dev_range) {
  memset(&v, 0, sizeof v);
  const GridParams &g = rt->grid;
  const int axis = g.dim - 1, N = rt->n_slabs, me = rt->my_slab;
  const Member &pm = p.members[p.pos_member];
  const u32 seq = p.halo_seq + 1;
  v.active = 1;
  v.dim = g.dim;
  v.pos_col = pm.first_col;
  v.n_layers = slab_layers(rt);
  v.origin = g.origin[axis];
  v.inv_cell = g.inv_cell;
  v.begin = rt->layer_begin;
  v.end = rt->layer_end;
  v.ghost = rt->ghost_layers;
  const bool ring = N > 2;
  const int lo = me > 0 ? me - 1 : (ring ? N - 1 : -1), hi = me < N - 1 ? me + 1 : (ring ? 0 : -1);
  if (lo >= 0 && p.halo_peer[0]) {
    v.lo_begin = rt->slab_bounds[lo];
    v.lo_end = rt->slab_bounds[lo + 1];
    v.lo_ghost = me > 0;  // around the ring only migrants travel, ghosts are for true neighbours
    // what I send downwards arrives "from upper" over there
    v.msg[0] = halo_block_of(p.halo_peer[0], p.halo_block, 1, seq) + ABL_MSG_HEADER;
  }
  if (hi >= 0 && p.halo_peer[1]) {
    v.hi_begin = rt->slab_bounds[hi];
    v.hi_end = rt->slab_bounds[hi + 1];
    v.hi_ghost = me < N - 1;
    v.msg[1] = halo_block_of(p.halo_peer[1], p.halo_block, 0, seq) + ABL_MSG_HEADER;
  }
  v.count[0] = p.halo_ctr + 0;
  v.count[1] = p.halo_ctr + 1;
  v.far = p.halo_ctr + 5;
  v.capacity = (unsigned)p.halo_cap;
  v.rec_words = (unsigned)slab_rec_words(p);
  v.n_cols = (int)p.cols.size() - 1;
  for (int c = 0; c < v.n_cols; c++) v.elem[c] = p.cols[c].elem;
  // boundary-first scheduling with the publish inside the step kernel
  const u32 n_own = p.own_end - p.own_begin;
  const u32 lo_end = std::min(std::max(p.bnd_lo_end, p.own_begin), p.own_end) - p.own_begin;
  const u32 hi_begin = std::min(std::max(p.bnd_hi_begin, p.own_begin + lo_end), p.own_end) - p.own_begin;
  if (dev_range) {
    // the kernel reads the ranges from cell_start itself (global cell keys; the view's
    // cell_start pointer has a virtual origin)
    u32 lo, hi, lo2, hi2;
    slab_owned_cells(rt, &lo, &hi);
    slab_boundary_cells(rt, &lo2, &hi2);
    v.device_range = 1;
    v.lo_key = lo + rt->grid.key_base;
    v.hi_key = hi + rt->grid.key_base;
    v.lo2_key = lo2 + rt->grid.key_base;
    v.hi2_key = hi2 + rt->grid.key_base;
    v.boundary_first = rt->halo_overlap ? 1 : 0;
  } else if (step_kernel && rt->halo_overlap && n_own && (lo_end > 0 || hi_begin < n_own)) {
    v.boundary_first = 1;
    v.idx_lo_end = lo_end;
    v.idx_hi_begin = hi_begin;
  }
  v.published = p.halo_ctr + 13;
  if (v.boundary_first) {
    v.done = p.halo_ctr + 11;
    v.late = p.halo_ctr + 12;
    v.sent = p.halo_ctr + 6;
    v.hdr[0] = v.msg[0] ? (unsigned *)(v.msg[0] - ABL_MSG_HEADER) : nullptr;
    v.hdr[1] = v.msg[1] ? (unsigned *)(v.msg[1] - ABL_MSG_HEADER) : nullptr;
    v.seq = seq;
  }
  return ABL_OK;
}

// Room for the arrivals must exist before the packing kernel runs (growing a pool
// synchronises and moves its columns).
static int halo_reserve(abl_runtime *rt, Pool &p) {
  if (p.halo_pad == 0) {
    p.halo_pad = (u32)std::min<size_t>(2 * p.halo_cap, round_up(p.halo_cap / 2 + 256, 256));
    if (const char *e = getenv("ABL_CUDA_HALO_PAD")) p.halo_pad = (u32)std::max(1, atoi(e));  // tests
  }
  // (with a device-side range the host only knows an upper bound of the owned range: the pool size)
  const size_t own_end = p.report_pending ? p.n : (size_t)p.own_end;
  const size_t want = own_end + 2 * (size_t)p.halo_pad + 1024;
  if (want > p.cap) {
    p.n = std::max(p.n, own_end);
    TRY(drop_fused_histogram(rt, p));
    TRY(reserve_pool(rt, p, want));
  }
  return ABL_OK;
}

// publish + wait + unpack of exchange number ++halo_seq; records have been packed by the
// kernel launched just before
static int halo_finish(abl_runtime *rt, Pool &p, bool published, bool dev_range, bool forked) {
  const u32 seq = ++p.halo_seq;
  // (dev_range: the host's owned range is stale; the kernel reads it from cell_start)
  const u32 ob = p.own_begin, oe = p.own_end, n_own = oe - ob;
  HaloHeader *to_lo = p.halo_peer[0] ? (HaloHeader *)halo_block_of(p.halo_peer[0], p.halo_block, 1, seq) : nullptr;
  HaloHeader *to_hi = p.halo_peer[1] ? (HaloHeader *)halo_block_of(p.halo_peer[1], p.halo_block, 0, seq) : nullptr;
  const u8 *from_lo = p.halo_peer[0] ? halo_block_of(p.halo_recv, p.halo_block, 0, seq) : nullptr;
  const u8 *from_hi = p.halo_peer[1] ? halo_block_of(p.halo_recv, p.halo_block, 1, seq) : nullptr;
  ColTable t;
  fill_table(p, t, false);
  const u32 pad = p.halo_pad;
  p.pad_used[seq & 1u] = pad;
  HaloExchangeArgs a;
  memset(&a, 0, sizeof a);
  a.ctr = p.halo_ctr;
  a.to_lo = to_lo; a.to_hi = to_hi;
  a.from_lo = from_lo; a.from_hi = from_hi;
  a.seq = seq; a.cap = (u32)p.halo_cap;
  a.timeout_ns = rt->halo_timeout_ns;
  a.trace = rt->trace ? 1 : 0;
  a.pad = pad; a.rec_words = (u32)slab_rec_words(p);
  if (dev_range) {
    a.dev_range = 1;
    a.cs = p.cell_start;
    slab_owned_cells(rt, &a.lo_cell, &a.hi_cell);
    a.pool_cap = (u32)std::min<size_t>(p.cap, 0x7fffffffu);
    a.publish = 2;
  } else {
    a.publish = published ? 0 : 1;
    a.dst_first = oe;
    a.room = (u32)std::min<size_t>(p.cap - oe, 0x7fffffffu);
  }
  const GridParams &g = rt->grid;
  // the step kernel already produced keys and the histogram of the owned agents (fused
  // epilogue); arrivals and padding are added by the exchange kernel
  a.count = p.counted ? 1 : 0;
  if (a.count) {
    const Member &pm = p.members[p.pos_member];
    a.px = p.cols[pm.first_col].buf[p.cols[pm.first_col].cur];
    if (g.dim == 3) {
      a.py = p.cols[pm.first_col + 1].buf[p.cols[pm.first_col + 1].cur];
      a.pz = p.cols[pm.first_col + 2].buf[p.cols[pm.first_col + 2].cur];
    }
    a.key_first = n_own;
    a.key = p.key; a.local = p.local; a.cell_count = p.cell_count;
  }
  // Every block polls the headers before it unpacks.  When several slabs share one GPU (local
  // peers, single host thread) the spinning blocks of one slab must leave room for the kernels
  // of the others, so the grid stays small there; across GPUs it covers the device.
  const bool local_peers = (p.halo_peer[0] && !p.halo_ipc[0]) || (p.halo_peer[1] && !p.halo_ipc[1]);
  // (forked: the kernel runs next to the step kernel and waits there for the neighbours — a few blocks only)
  const u32 nb = std::max(1u, std::min(blocks_for(pad, 256), forked ? 16u : local_peers ? 32u : 4u * 148u));
  cudaStream_t xs = rt->stream;
  if (forked) {
    // everything queued before the step kernel (the binning that used key / local / histogram) first
    CU(cudaStreamWaitEvent(rt->side_stream, rt->ev_fork, 0));
    xs = rt->side_stream;
  }
  if (rt->real_size == 8) {
    if (g.dim == 2) k_halo_exchange<double, 2><<<nb, 256, 0, xs>>>(t, a, g);
    else k_halo_exchange<double, 3><<<nb, 256, 0, xs>>>(t, a, g);
  } else {
    if (g.dim == 2) k_halo_exchange<float, 2><<<nb, 256, 0, xs>>>(t, a, g);
    else k_halo_exchange<float, 3><<<nb, 256, 0, xs>>>(t, a, g);
  }
  rt->launches++;
  CU(cudaGetLastError());
  if (forked) {
    CU(cudaEventRecord(rt->ev_join, rt->side_stream));
    CU(cudaStreamWaitEvent(rt->stream, rt->ev_join, 0));
  }
  p.halo_pending = true;
  p.binned = false;
  if (dev_range) {
    // src_begin and the extent of the next binning follow from the report still in flight
    p.exch_deferred = true;
    p.exch_pad = pad;
  } else {
    p.halo_prev_ob = ob;
    p.halo_prev_own = n_own;
    p.src_begin = ob;
    p.n = (size_t)n_own + pad;
  }
  return ABL_OK;
}

static int halo_exchange_standalone(abl_runtime *rt, Pool &p) {
  // (the owned range of the binning the step ran on — not a fresh binning of the new positions)
  if (p.report_pending && !p.exch_deferred) TRY(slab_consume_report(rt, p, nullptr));
  if (!p.own_valid) TRY(bin_pool(rt, p));
  TRY(halo_reserve(rt, p));
  abl_slab_view v;
  TRY(halo_fill_view(rt, p, v, false, false));
  const u32 n_own = p.own_end - p.own_begin;
  if (n_own) {
    ColTable t;
    fill_table(p, t, false);
    if (rt->real_size == 8) k_halo_pack<double><<<blocks_for(n_own, 256), 256, 0, rt->stream>>>(t, v, n_own, p.own_begin);
    else k_halo_pack<float><<<blocks_for(n_own, 256), 256, 0, rt->stream>>>(t, v, n_own, p.own_begin);
    rt->launches++;
    CU(cudaGetLastError());
  }
  // the host sized this exchange blindly: let the next binning verify the counts synchronously
  // (it can still bin again if more arrived than the padding covered)
  p.halo_sync_check = true;
  return halo_finish(rt, p, false, false);
}

extern "C" int abl_cuda_halo_setup(abl_runtime *rt, int pool, size_t capacity_records, void *handle_out) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (!rt->slab) return fail(ABL_ERR_STATE, "halo_setup requires abl_cuda_set_slab");
  if (p.pos_member < 0) return fail(ABL_ERR_ARGUMENT, "pool %s has no position member", p.name.c_str());
  if (p.halo_recv) return fail(ABL_ERR_STATE, "pool %s: halo transport already set up", p.name.c_str());
  CU(cudaSetDevice(rt->device));
  p.halo_cap = round_up(capacity_records ? capacity_records : 65536, 32);   // rows of the word-major message stay 128-byte aligned
  p.halo_block = round_up(ABL_MSG_HEADER + p.halo_cap * (size_t)slab_rec_words(p) * 4, 256);
  CU(cudaMalloc(&p.halo_recv, 4 * p.halo_block));
  CU(cudaMemset(p.halo_recv, 0, 4 * p.halo_block));
  rt->defer_free = true;
  CU(cudaMalloc(&p.halo_ctr, 16 * sizeof(u32)));
  CU(cudaMemset(p.halo_ctr, 0, 16 * sizeof(u32)));
  // CUDA loads kernels lazily and the first launch of a kernel may synchronise the context.
  // That must not happen between k_halo_wait and the neighbours' publish when several slabs
  // are driven by one host thread (it would wait for a spinning kernel): load them now.
  {
    const void *kernels[] = {(const void *)k_halo_exchange<double, 2>, (const void *)k_halo_exchange<double, 3>,
                             (const void *)k_halo_exchange<float, 2>, (const void *)k_halo_exchange<float, 3>,
                             (const void *)k_halo_pack<double>, (const void *)k_halo_pack<float>,
                             (const void *)k_bin_count<double, 2>, (const void *)k_bin_count<double, 3>,
                             (const void *)k_bin_count<float, 2>, (const void *)k_bin_count<float, 3>,
                             (const void *)k_gather_bin_words, (const void *)k_bin_scatter,
                             (const void *)k_bin_rank_move};
    cudaFuncAttributes attr;
    for (const void *k : kernels) CU(cudaFuncGetAttributes(&attr, k));
  }
  if (handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == ABL_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, p.halo_recv));
    memcpy(handle_out, &h, sizeof h);
  }
  return ABL_OK;
}

extern "C" int abl_cuda_halo_connect(abl_runtime *rt, int pool, const void *lower_handle, const void *upper_handle) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (!p.halo_recv) return fail(ABL_ERR_STATE, "halo_connect before halo_setup");
  CU(cudaSetDevice(rt->device));
  const void *hs[2] = {lower_handle, upper_handle};
  for (int d = 0; d < 2; d++) {
    if (!hs[d]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hs[d], sizeof h);
    void *ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p.halo_peer[d] = (u8 *)ptr;
    p.halo_ipc[d] = true;
  }
  return ABL_OK;
}

extern "C" int abl_cuda_halo_connect_local(abl_runtime *rt, int pool, abl_runtime *lower, abl_runtime *upper) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (!p.halo_recv) return fail(ABL_ERR_STATE, "halo_connect before halo_setup");
  abl_runtime *peers[2] = {lower, upper};
  for (int d = 0; d < 2; d++) {
    if (!peers[d]) continue;
    Pool *q = nullptr;
    TRY(get_pool(peers[d], pool, &q));
    if (!q->halo_recv || q->halo_block != p.halo_block)
      return fail(ABL_ERR_STATE, "halo_connect_local: the peer has no matching receive area");
    if (peers[d]->device != rt->device) {
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, rt->device, peers[d]->device));
      if (!can) return fail(ABL_ERR_COMM, "device %d cannot access device %d", rt->device, peers[d]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[d]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
      cudaGetLastError();
    }
    p.halo_peer[d] = q->halo_recv;
    p.halo_ipc[d] = false;
  }
  return ABL_OK;
}

extern "C" int abl_cuda_exchange(abl_runtime *rt, int pool) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (!rt->slab || p.pos_member < 0) return ABL_OK;
  if (halo_direct(p)) {
    CU(cudaSetDevice(rt->device));
    return halo_exchange_standalone(rt, p);
  }
  if (rt->peer_lo || rt->peer_hi)
    return fail(ABL_ERR_STATE, "in-process peers: use abl_cuda_exchange_begin/end on all runtimes");
  CU(cudaSetDevice(rt->device));
  const int N = rt->world, me = rt->rank;
  const bool ring = N > 2;
  const bool has_lo = rt->comm && N > 1 && (me > 0 || ring), has_hi = rt->comm && N > 1 && (me < N - 1 || ring);
  const int lo_peer = (me - 1 + N) % N, hi_peer = (me + 1) % N;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double h0 = 0, h1 = 0, h2 = 0, h3 = 0;
  if (rt->trace) { cudaEventRecord(rt->xev[0], rt->stream); h0 = now(); }
  TRY(exchange_pack(rt, p, has_lo, has_hi));
  if (rt->trace) { cudaEventRecord(rt->xev[1], rt->stream); h1 = now(); }
  u32 incoming[2] = {0, 0};
  if (has_lo || has_hi) {
    // One NCCL group moves fixed-size first messages (header + up to kMsgFirstRecords records),
    // so neither side needs the other's count beforehand; one host synchronisation then reads
    // all four headers.  Oversized messages (rare) send their tail in a second group.
    // Issue order per rank: (send to lower, receive from upper), (send to upper, receive from
    // lower) — with two ranks both peers are the same process and NCCL matches operations
    // between a pair in issue order.
    const size_t first_bytes = slab_msg_bytes(p, kMsgFirstRecords);
    TRY(ensure_xbuf(rt, 2, first_bytes));
    TRY(ensure_xbuf(rt, 3, first_bytes));
    NCCL(ncclGroupStart());
    if (has_lo) NCCL(ncclSend(rt->xbuf[0], first_bytes, ncclUint8, lo_peer, rt->comm, rt->stream));
    if (has_hi) NCCL(ncclRecv(rt->xbuf[3], first_bytes, ncclUint8, hi_peer, rt->comm, rt->stream));
    if (has_hi) NCCL(ncclSend(rt->xbuf[1], first_bytes, ncclUint8, hi_peer, rt->comm, rt->stream));
    if (has_lo) NCCL(ncclRecv(rt->xbuf[2], first_bytes, ncclUint8, lo_peer, rt->comm, rt->stream));
    NCCL(ncclGroupEnd());
    if (rt->trace) { cudaEventRecord(rt->xev[2], rt->stream); h2 = now(); }
    u32 *h = rt->h_scalar + 8;  // [0] out lo, [1] out hi, [2] in lo, [3] in hi, [4] far
    k_gather_words<<<1, 32, 0, rt->stream>>>((const u32 *)rt->xbuf[0], (const u32 *)rt->xbuf[1],
        has_lo ? (const u32 *)rt->xbuf[2] : nullptr, has_hi ? (const u32 *)rt->xbuf[3] : nullptr,
        rt->d_scalar + 10, rt->d_scalar + 24);
    rt->launches++;
    CU(cudaMemcpyAsync(h, rt->d_scalar + 24, 5 * sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
    if (rt->trace) cudaEventRecord(rt->xev[3], rt->stream);
    CU(cudaStreamSynchronize(rt->stream));
    if (rt->trace) h3 = now();
    TRY(exchange_check_far(rt, p, h[4]));
    const u32 out_lo = has_lo ? h[0] : 0, out_hi = has_hi ? h[1] : 0;
    incoming[0] = has_lo ? h[2] : 0;
    incoming[1] = has_hi ? h[3] : 0;
    const u32 K = kMsgFirstRecords;
    if (out_lo > K || out_hi > K || incoming[0] > K || incoming[1] > K) {
      const size_t rec = (size_t)slab_rec_words(p) * 4;
      if (incoming[0] > K) TRY(grow_recv(rt, 2, slab_msg_bytes(p, incoming[0]), first_bytes));
      if (incoming[1] > K) TRY(grow_recv(rt, 3, slab_msg_bytes(p, incoming[1]), first_bytes));
      NCCL(ncclGroupStart());
      if (out_lo > K) NCCL(ncclSend((u8 *)rt->xbuf[0] + first_bytes, (out_lo - K) * rec, ncclUint8, lo_peer, rt->comm, rt->stream));
      if (incoming[1] > K) NCCL(ncclRecv((u8 *)rt->xbuf[3] + first_bytes, (incoming[1] - K) * rec, ncclUint8, hi_peer, rt->comm, rt->stream));
      if (out_hi > K) NCCL(ncclSend((u8 *)rt->xbuf[1] + first_bytes, (out_hi - K) * rec, ncclUint8, hi_peer, rt->comm, rt->stream));
      if (incoming[0] > K) NCCL(ncclRecv((u8 *)rt->xbuf[2] + first_bytes, (incoming[0] - K) * rec, ncclUint8, lo_peer, rt->comm, rt->stream));
      NCCL(ncclGroupEnd());
    }
  }
  int rc = exchange_unpack(rt, p, incoming);
  if (rt->trace && (has_lo || has_hi) && rc == ABL_OK) {
    cudaEventRecord(rt->xev[4], rt->stream);
    double h4 = now();
    cudaEventSynchronize(rt->xev[4]);
    float ms;
    if (++rt->xskip > 64) {  // connection set-up and warm-up calls are not representative
      for (int i = 0; i < 4; i++) { cudaEventElapsedTime(&ms, rt->xev[i], rt->xev[i + 1]); rt->xdev[i] += ms; }
      rt->xhost[0] += h1 - h0; rt->xhost[1] += h2 - h1; rt->xhost[2] += h3 - h2; rt->xhost[3] += h4 - h3;
      rt->xcount++;
    }
  }
  return rc;
}

// After a step function removed or added agents: the owned range changed under the neighbours'
// feet, so ghosts are refreshed even when no member was written.  With in-process peers the
// caller exchanges (exchange_begin / exchange_end on every runtime).
static int slab_after_mutation(abl_runtime *rt, int pool_index) {
  Pool &p = rt->pools[pool_index];
  if (halo_direct(p)) return halo_exchange_standalone(rt, p);
  if (rt->peer_lo || rt->peer_hi) return ABL_OK;
  return abl_cuda_exchange(rt, pool_index);
}

// ---- in-process transport: several runtimes (slabs) driven by one host thread -------------
// Used to exercise the slab bookkeeping on a single GPU: the caller runs the step on every
// slab, then exchange_begin on every slab, then exchange_end on every slab.
extern "C" int abl_cuda_set_local_peers(abl_runtime *rt, abl_runtime *lower, abl_runtime *upper) {
  if (!rt) return fail(ABL_ERR_ARGUMENT, "null runtime");
  rt->peer_lo = lower;
  rt->peer_hi = upper;
  return ABL_OK;
}

extern "C" int abl_cuda_exchange_begin(abl_runtime *rt, int pool) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  if (!rt->slab || pp->pos_member < 0) return ABL_OK;
  CU(cudaSetDevice(rt->device));
  TRY(exchange_pack(rt, *pp, rt->peer_lo != nullptr, rt->peer_hi != nullptr));
  u32 *h = rt->h_scalar + 8;
  CU(cudaMemcpyAsync(h + 0, rt->xbuf[0], sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaMemcpyAsync(h + 1, rt->xbuf[1], sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaMemcpyAsync(h + 4, rt->d_scalar + 10, sizeof(u32), cudaMemcpyDeviceToHost, rt->stream));
  CU(cudaStreamSynchronize(rt->stream));
  rt->x_out[0] = h[0];
  rt->x_out[1] = h[1];
  return exchange_check_far(rt, *pp, h[4]);
}

extern "C" int abl_cuda_exchange_end(abl_runtime *rt, int pool) {
  Pool *pp = nullptr;
  TRY(get_pool(rt, pool, &pp));
  Pool &p = *pp;
  if (!rt->slab || p.pos_member < 0) return ABL_OK;
  CU(cudaSetDevice(rt->device));
  u32 incoming[2] = {0, 0};
  if (rt->peer_lo) {  // what the lower slab sent upwards
    incoming[0] = rt->peer_lo->x_out[1];
    size_t bytes = slab_msg_bytes(p, incoming[0]);
    TRY(ensure_xbuf(rt, 2, bytes));
    CU(cudaMemcpyAsync(rt->xbuf[2], rt->peer_lo->xbuf[1], bytes, cudaMemcpyDefault, rt->stream));
  }
  if (rt->peer_hi) {  // what the upper slab sent downwards
    incoming[1] = rt->peer_hi->x_out[0];
    size_t bytes = slab_msg_bytes(p, incoming[1]);
    TRY(ensure_xbuf(rt, 3, bytes));
    CU(cudaMemcpyAsync(rt->xbuf[3], rt->peer_hi->xbuf[0], bytes, cudaMemcpyDefault, rt->stream));
  }
  TRY(exchange_unpack(rt, p, incoming));
  CU(cudaStreamSynchronize(rt->stream));
  return ABL_OK;
}

// ---------------------------------------------------------------------------------------
// one process, several GPUs: the driver behind `-C cuda.gpus=N` (see abl_cuda.h)
// ---------------------------------------------------------------------------------------
namespace {
struct GroupSlab {
  abl_runtime *rt = nullptr;
  int device = 0;
  int rc = ABL_OK;
  std::string err;
  std::vector<u8 *> send;                  // per type: transit records grouped by owner (device memory)
  std::vector<std::vector<u32>> counts;    // per type: records for every slab
};

template <typename F> int for_each_slab(std::vector<GroupSlab> &slabs, F f) {
  std::vector<std::thread> th;
  for (size_t r = 0; r < slabs.size(); r++)
    th.emplace_back([&, r] {
      GroupSlab &g = slabs[r];
      if (g.rc != ABL_OK) return;
      cudaSetDevice(g.device);
      g.rc = f((int)r, g);
      if (g.rc != ABL_OK) g.err = abl_cuda_last_error();
    });
  for (auto &t : th) t.join();
  for (GroupSlab &g : slabs)
    if (g.rc != ABL_OK) return fail(g.rc, "device %d: %s", g.device, g.err.c_str());
  return ABL_OK;
}
}  // namespace

extern "C" int abl_cuda_group_simulate(const abl_config *cfg_in, int n_gpus, int (*setup)(abl_runtime *),
                                       int (*timestep)(abl_runtime *), int timesteps, const abl_group_population *pop) {
  if (!cfg_in || !setup || !timestep || !pop || n_gpus < 1 || n_gpus > ABL_MAX_SLABS)
    return fail(ABL_ERR_ARGUMENT, "bad arguments to group_simulate");
  int visible = 0;
  CU(cudaGetDeviceCount(&visible));
  // (ABL_CUDA_OVERSUBSCRIBE=1: more slabs than devices, several slabs share a GPU — how the single-GPU test tier runs this driver)
  const bool share = getenv("ABL_CUDA_OVERSUBSCRIBE") && atoi(getenv("ABL_CUDA_OVERSUBSCRIBE")) != 0;
  if (visible < 1 || (n_gpus > visible && !share)) return fail(ABL_ERR_ARGUMENT, "%d GPUs requested, %d visible", n_gpus, visible);
  const int N = n_gpus, T = pop->n_types;
  std::vector<GroupSlab> slabs(N);
  auto cleanup = [&] { for (GroupSlab &g : slabs) { for (u8 *p : g.send) if (p) { cudaSetDevice(g.device); cudaFree(p); } if (g.rt) abl_cuda_destroy(g.rt); g.rt = nullptr; } };
  // 1. one runtime per device, same pools and steps everywhere
  for (int r = 0; r < N; r++) {
    slabs[r].device = ((cfg_in->device >= 0 ? cfg_in->device : 0) + r) % visible;
    slabs[r].send.assign(T, nullptr);
    slabs[r].counts.assign(T, std::vector<u32>(N, 0));
  }
  int rc = for_each_slab(slabs, [&](int, GroupSlab &g) -> int {
    abl_config cfg = *cfg_in;
    cfg.device = g.device;
    TRY(abl_cuda_create(&g.rt, &cfg));
    if (setup(g.rt) != 0) return fail(ABL_ERR_STATE, "model setup failed: %s", abl_cuda_last_error());
    for (const Step &s : g.rt->steps)
      if (s.desc.added_pool >= 0)
        return fail(ABL_ERR_STATE, "step %s adds agents at run time: not supported with cuda.gpus > 1 (the ids of new agents are "
                                   "resolved across the slabs by the host; use the torchrun harness openabl_b200.slab.RankSlab)", s.name.c_str());
    return ABL_OK;
  });
  if (rc != ABL_OK) { cleanup(); return rc; }
  // 2. slabs of whole cell layers, ring of peer-mapped receive areas
  int layers = 0;
  rc = abl_cuda_slab_axis_layers(slabs[0].rt, &layers);
  if (rc == ABL_OK && layers < N) rc = fail(ABL_ERR_ARGUMENT, "more GPUs (%d) than cell layers (%d)", N, layers);
  if (rc != ABL_OK) { cleanup(); return rc; }
  std::vector<int> bounds(N + 1);
  for (int r = 0; r <= N; r++) bounds[r] = (int)(((long long)layers * r) / N);
  rc = for_each_slab(slabs, [&](int r, GroupSlab &g) -> int { return abl_cuda_set_slab(g.rt, bounds.data(), N, r); });
  if (rc == ABL_OK && N > 1) {
    rc = for_each_slab(slabs, [&](int, GroupSlab &g) -> int {
      for (int t = 0; t < T; t++) {
        if (!pop->has_position[t]) continue;
        const size_t cap = std::max<size_t>(16384, 4 * pop->len[t] / (size_t)std::max(layers, 1) + 4096);
        TRY(abl_cuda_halo_setup(g.rt, pop->pool[t], cap, nullptr));
      }
      return ABL_OK;
    });
    if (rc == ABL_OK)
      rc = for_each_slab(slabs, [&](int r, GroupSlab &g) -> int {
        const bool ring = N > 2;
        abl_runtime *lo = r > 0 ? slabs[r - 1].rt : (ring ? slabs[N - 1].rt : nullptr);
        abl_runtime *hi = r + 1 < N ? slabs[r + 1].rt : (ring ? slabs[0].rt : nullptr);
        for (int t = 0; t < T; t++)
          if (pop->has_position[t]) TRY(abl_cuda_halo_connect_local(g.rt, pop->pool[t], lo, hi));
        return ABL_OK;
      });
  }
  if (rc != ABL_OK) { cleanup(); return rc; }
  // 3. upload: slab r takes the r-th part of every array by index, classifies it on its device ...
  rc = for_each_slab(slabs, [&](int r, GroupSlab &g) -> int {
    for (int t = 0; t < T; t++) {
      const size_t n = pop->len[t];
      if (!pop->has_position[t]) { TRY(abl_cuda_upload(g.rt, pop->pool[t], pop->data[t], n)); continue; }
      const size_t b = n * (size_t)r / N, e = n * (size_t)(r + 1) / N;
      if (e > b) CU(cudaMalloc(&g.send[t], (e - b) * (size_t)transit_bytes((u32)pop->stride[t])));
      TRY(abl_cuda_partition_upload(g.rt, pop->pool[t], (const u8 *)pop->data[t] + b * pop->stride[t], e - b, (unsigned)b,
                                    g.send[t], g.counts[t].data()));
    }
    return ABL_OK;
  });
  // ... and every slab fetches the records it owns from all of them (peer copies), then adopts them
  if (rc == ABL_OK)
    rc = for_each_slab(slabs, [&](int d, GroupSlab &g) -> int {
      for (int t = 0; t < T; t++) {
        if (!pop->has_position[t]) continue;
        const size_t rec = transit_bytes((u32)pop->stride[t]);
        size_t total = 0;
        for (int r = 0; r < N; r++) total += slabs[r].counts[t][d];
        u8 *recv = nullptr;
        if (total) CU(cudaMalloc(&recv, total * rec));
        size_t at = 0;
        for (int r = 0; r < N; r++) {
          const size_t cnt = slabs[r].counts[t][d];
          if (!cnt) continue;
          size_t before = 0;
          for (int q = 0; q < d; q++) before += slabs[r].counts[t][q];
          CU(cudaMemcpyPeer(recv + at * rec, g.device, slabs[r].send[t] + before * rec, slabs[r].device, cnt * rec));
          at += cnt;
        }
        int arc = abl_cuda_adopt_records(g.rt, pop->pool[t], recv, total, (unsigned)pop->len[t]);
        if (recv) cudaFree(recv);
        TRY(arc);
      }
      return ABL_OK;
    });
  if (rc == ABL_OK)
    rc = for_each_slab(slabs, [&](int, GroupSlab &g) -> int {
      for (u8 *&p : g.send) if (p) { cudaFree(p); p = nullptr; }
      if (N > 1)
        for (int t = 0; t < T; t++)
          if (pop->has_position[t]) TRY(abl_cuda_exchange(g.rt, pop->pool[t]));   // ghosts of the initial state
      return ABL_OK;
    });
  // 4. the simulation: every device runs its slab, coupled only through the halo messages
  if (rc == ABL_OK)
    rc = for_each_slab(slabs, [&](int, GroupSlab &g) -> int {
      for (int s = 0; s < timesteps; s++)
        if (timestep(g.rt) != 0) return fail(ABL_ERR_STATE, "timestep %d failed: %s", s, abl_cuda_last_error());
      return abl_cuda_synchronize(g.rt);
    });
  // 5. download: owned records of every slab, placed by agent id
  if (rc == ABL_OK) {
    for (int t = 0; t < T && rc == ABL_OK; t++) {
      const size_t stride = pop->stride[t], n = pop->len[t];
      if (!pop->has_position[t]) { size_t got = 0; rc = abl_cuda_download(slabs[0].rt, pop->pool[t], pop->data[t], n, &got); continue; }
      std::vector<std::vector<u8>> rec(N);
      std::vector<std::vector<unsigned>> ids(N);
      rc = for_each_slab(slabs, [&](int r, GroupSlab &g) -> int {
        size_t own = 0;
        TRY(abl_cuda_owned_size(g.rt, pop->pool[t], &own));
        rec[r].resize(own * stride);
        ids[r].resize(own);
        size_t got = 0;
        TRY(abl_cuda_download(g.rt, pop->pool[t], rec[r].data(), own, &got));
        TRY(abl_cuda_download_ids(g.rt, pop->pool[t], ids[r].data(), own, &got));
        return ABL_OK;
      });
      if (rc != ABL_OK) break;
      size_t total = 0;
      for (int r = 0; r < N; r++) total += ids[r].size();
      if (total != n) { rc = fail(ABL_ERR_STATE, "%zu agents came back, %zu were uploaded", total, n); break; }
      for (int r = 0; r < N && rc == ABL_OK; r++)
        for (size_t k = 0; k < ids[r].size(); k++) {
          if (ids[r][k] >= n) { rc = fail(ABL_ERR_STATE, "agent id %u out of range", ids[r][k]); break; }
          memcpy((u8 *)pop->data[t] + (size_t)ids[r][k] * stride, rec[r].data() + k * stride, stride);
        }
    }
  }
  std::string keep = g_err;
  cleanup();
  if (rc != ABL_OK) snprintf(g_err, sizeof g_err, "%s", keep.c_str());
  return rc;
}
