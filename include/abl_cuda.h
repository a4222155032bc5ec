/* abl_cuda.h — C ABI of the B200-native OpenABL runtime (asset/cuda, libabl_cuda.so).
 *
 * This is the drop-in boundary on the run-time side.  The reference `c` backend has no
 * FFI: its generated main.c owns agent storage (`dyn_array`, reference
 * asset/c/libabl.h:11-57) and runs the simulate loop inline (reference
 * src/backend/CPrinter.cpp:189-233).  The `cuda` backend keeps that host program shape
 * and replaces exactly those two pieces by calls into this library; every entry point
 * below cites the reference construct it stands in for.
 *
 * Conventions: plain C types only; every function returns 0 on success or a non-zero
 * ABL_ERR_* code, with a human-readable message available from abl_cuda_last_error().
 * One host thread drives one runtime; one runtime drives one CUDA device.  Device memory
 * is owned by the runtime, host buffers stay owned by the caller.
 *
 * There is NO CPU fallback: if no CUDA device is usable abl_cuda_create() fails.
 */
#ifndef ABL_CUDA_H
#define ABL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABL_CUDA_ABI_VERSION 1
#define ABL_MAX_MEMBERS 16   /* members per agent type */
#define ABL_MAX_COLUMNS 32   /* SoA columns per agent type (a float3 member is 3 columns) */
#define ABL_MAX_SLABS 64     /* slabs (GPUs) of one decomposed simulation */

enum {
  ABL_OK = 0,
  ABL_ERR_CUDA = 1,        /* a CUDA call failed / no device */
  ABL_ERR_ARGUMENT = 2,    /* bad handle, index or descriptor */
  ABL_ERR_STATE = 3,       /* call order violated (e.g. step before environment) */
  ABL_ERR_CAPACITY = 4,    /* host buffer too small */
  ABL_ERR_COMM = 5         /* multi-GPU exchange failed */
};

/* Member kinds.  Numbering equals the reference's type_id (asset/c/libabl.h:188-196) so a
 * generated type table can be passed through unchanged. */
enum {
  ABL_TYPE_END = 0,
  ABL_TYPE_BOOL = 1,
  ABL_TYPE_INT = 2,
  ABL_TYPE_FLOAT = 3,
  ABL_TYPE_STRING = 4, /* not storable in agents */
  ABL_TYPE_FLOAT2 = 5,
  ABL_TYPE_FLOAT3 = 6
};

/* One agent member inside the host AoS record; mirrors `type_info`
 * (reference asset/c/libabl.h:198-203). */
typedef struct {
  int type;            /* ABL_TYPE_* */
  unsigned offset;     /* byte offset inside the host record */
  const char *name;
  int is_pos;          /* the `position` member */
} abl_member_desc;

/* One agent type; mirrors the per-agent `type_info[]` + `agent_info` pair
 * (reference src/backend/CPrinter.cpp:250-265, 286-302). */
typedef struct {
  const char *name;
  const abl_member_desc *members;
  int n_members;
  unsigned stride;     /* sizeof(host record) */
} abl_agent_desc;

typedef struct abl_runtime abl_runtime;

typedef struct {
  int device;            /* CUDA device ordinal; -1 = current device */
  int use_float;         /* 1: abl_float is float (reference -DLIBABL_USE_FLOAT=1), 0: double */
  uint64_t seed;         /* seed of the in-step counter-based RNG */
  int deterministic;     /* reserved, must be 1: neighbour order = (cell, agent id) */
  int tile_neighbours;   /* 1: step kernels stage neighbour cell ranges in shared memory (default 0: measured
                            neutral to slightly slower than the L1-cached global loop on B200) */
  int block_size;        /* threads per CTA for step kernels (0 = automatic: 128, or 256 for dense neighbourhoods) */
} abl_config;

/* ---- life cycle -------------------------------------------------------------------- */
int abl_cuda_abi_version(void);
void abl_cuda_default_config(abl_config *cfg);
int abl_cuda_create(abl_runtime **rt, const abl_config *cfg);
int abl_cuda_destroy(abl_runtime *rt);
const char *abl_cuda_last_error(void);

/* ---- environment (reference EnvironmentDeclaration: envMin/envMax/envGranularity,
 *      src/AST.hpp:658-662; grid sizing rule ceil(size/cell) per axis) -------------------
 * The cell size is granularity * (1 + ABL_CELL_PAD_*): with cell == radius exactly, two agents at a
 * distance the reference's float-rounded filter still accepts (up to R(1 + 6e-8)) could be binned two
 * cells apart (the cell index is the floor of a rounded product) and the 3^d search would miss the
 * pair.  The padding is far above the rounding of the index computation in either precision and
 * costs 2e-6 (double) / 2e-3 (float) more candidates. */
#define ABL_CELL_PAD_F64 (1.0 / 1048576.0)   /* 2^-20 */
#define ABL_CELL_PAD_F32 (1.0 / 1024.0)      /* 2^-10 */
int abl_cuda_set_environment(abl_runtime *rt, int dim, const double *env_min,
                             const double *env_max, double granularity);

/* ---- agent pools: SoA, per-column double buffered, replaces `agents_T` / `agents_T_dbuf`
 *      dyn_arrays (reference src/backend/CPrinter.cpp:286-302) --------------------------- */
int abl_cuda_add_pool(abl_runtime *rt, const abl_agent_desc *desc, int *pool);
/* Host AoS -> device SoA.  Agent i receives id i (ids are the tie-break of the cell sort
 * and the order of every download). */
int abl_cuda_upload(abl_runtime *rt, int pool, const void *host_aos, size_t n);
/* Same with caller-chosen agent ids (slab decomposition: every rank uploads its part of a
 * globally numbered population); next_id = 1 + the largest id in the whole population. */
int abl_cuda_upload_with_ids(abl_runtime *rt, int pool, const void *host_aos, const unsigned *ids,
                             size_t n, unsigned next_id);
/* Device SoA -> host AoS in ascending agent-id order (= original index order while no agent
 * has been removed; reference save() order, asset/c/libabl.c:96-109). */
int abl_cuda_download(abl_runtime *rt, int pool, void *host_aos, size_t capacity, size_t *n);
/* Ids of the agents abl_cuda_download returns, in the same (ascending) order. */
int abl_cuda_download_ids(abl_runtime *rt, int pool, unsigned *ids, size_t capacity, size_t *n);
int abl_cuda_pool_size(abl_runtime *rt, int pool, size_t *n);
/* Page-locks a caller-owned host buffer (the record array of an agent type) so that upload /
 * download are direct DMA transfers.  Optional; idempotent for an unchanged (ptr, bytes).  The
 * buffer must be unpinned before it is reallocated or freed. */
int abl_cuda_pin_host(abl_runtime *rt, void *ptr, size_t bytes);
int abl_cuda_unpin_host(abl_runtime *rt, void *ptr);

/* ---- step functions ------------------------------------------------------------------ */

/* Everything a generated step kernel needs, passed by value at launch. */
typedef struct {
  unsigned n;                              /* live agents */
  const void *in[ABL_MAX_COLUMNS];         /* current column base pointers */
  void *out[ABL_MAX_COLUMNS];              /* where this step writes (== in[] if unwritten) */
  const unsigned *id;                      /* agent ids, same order as the columns */
  const unsigned *cell_start;              /* [n_cells + 1] exclusive prefix, NULL if unbinned */
} abl_pool_view;

typedef struct {
  int dim;
  int n_cell[3];
  double origin[3];
  double cell_size;
  double inv_cell_size;   /* 1/cell_size rounded in the precision of abl_float */
  unsigned n_cells;
  /* window of cell layers along the slowest axis held by this runtime (all layers unless the
   * simulation is slab-decomposed) and the key of its first cell; cell_start is handed to
   * kernels with a virtual origin so that it is indexed by global cell key */
  int axis_lo, axis_hi;
  unsigned key_base;
} abl_grid_view;

/* Slab decomposition as seen by a step kernel: after storing its outputs the kernel itself
 * classifies the agent by the cell layer of its (new) position and appends the record to the
 * outgoing halo / migration messages.  The message buffers may live in the peer GPU's memory
 * (CUDA IPC mapping, written over NVLink), which makes the step kernel and the halo send one
 * fused kernel.  active == 0 when the simulation is not decomposed (or the runtime packs in a
 * separate pass). */
typedef struct {
  int active;
  int dim;                  /* 2 or 3: the slab axis is y or z */
  int pos_col;              /* first column of the position member */
  int n_layers;             /* cell layers along the slab axis */
  double origin, inv_cell;  /* of the slab axis */
  int begin, end, ghost;    /* owned layers [begin, end), ghost width in layers */
  int lo_begin, lo_end;     /* layers owned by the lower peer ([0,0): none) */
  int hi_begin, hi_end;
  int lo_ghost, hi_ghost;   /* 1: that peer is a true neighbour and wants ghost copies */
  unsigned char *msg[2];    /* record area of the outgoing message to the lower / upper peer */
  unsigned *count[2];       /* slot counters (local memory) */
  unsigned capacity;        /* records per message */
  unsigned rec_words;       /* 32-bit words per packed record */
  unsigned *far;            /* counts agents that moved farther than a neighbouring slab */
  int n_cols;               /* columns of the pool (without the id) */
  int elem[ABL_MAX_COLUMNS];
  /* Boundary-first scheduling (boundary_first != 0): the kernel's first nb_lo + nb_hi thread
   * blocks step the agents of the outermost cell layers of the slab — indices [0, idx_lo_end)
   * and [idx_hi_begin, n) of the launched range — and the last of them to finish publishes the
   * messages (counts + sequence number into the neighbours' headers) while the remaining
   * blocks are still stepping the interior [idx_lo_end, idx_hi_begin).  Interior agents are at
   * least two layers away from the halo zone; one that still needs routing is counted in
   * `late` (reported as an error by the runtime). */
  int boundary_first;
  unsigned idx_lo_end, idx_hi_begin;
  unsigned nb_lo, nb_hi;
  /* device_range != 0: the launch covers the whole pool (self.n is an upper bound, pointers are
   * not offset) and the kernel reads its range from the pool's own cell_start: owned agents are
   * [cell_start[lo_key], cell_start[hi_key]), the boundary parts end / begin at
   * cell_start[lo2_key] / cell_start[hi2_key] (global cell keys).  The host does not have to
   * know the range when it queues the kernel. */
  int device_range;
  unsigned lo_key, hi_key, lo2_key, hi2_key;
  unsigned *published;      /* sequence number of the last exchange the step kernel published itself */
  unsigned *done;           /* finished boundary blocks (reset by the publishing block) */
  unsigned *late;
  unsigned *sent;           /* [2] copies of the final counts, for the host's bookkeeping */
  unsigned *hdr[2];         /* headers {seq, count} of the neighbours' receive blocks (NULL: no such neighbour) */
  unsigned seq;
} abl_slab_view;

typedef struct {
  abl_pool_view self;                      /* pool the step function iterates */
  abl_pool_view nbr;                       /* pool of its for-near loop (n = 0 if none) */
  abl_grid_view grid;
  int reach;                               /* cells to visit on each side: ceil(radius/cell) */
  unsigned char *dead;                     /* [self.n] removeCurrent() flags, or NULL */
  unsigned char *add_flag;                 /* [self.n] add() flags, or NULL */
  void *add_cols[ABL_MAX_COLUMNS];         /* staging columns of the added type, [self.n] */
  /* when non-NULL the kernel also bins the positions it writes: bin_key[i] = cell key and
   * bin_count[key] += 1 (fused histogram of the next binning; bin_local is unused since the
   * slots inside a cell segment are handed out by the scatter pass) */
  unsigned *bin_key;
  unsigned *bin_local;
  unsigned *bin_count;
  abl_slab_view slab;                      /* fused halo/migration pack (see above) */
  uint64_t seed;
  unsigned timestep;
  unsigned step_index;
  int block_size;
  int tile_neighbours;
  int flat_loop;                           /* candidate loop of the step kernels: 0 cursor loop, 1 flat loop (sparse 2-D loops, ABL_MODE 3) — both
                                              leave dense neighbourhoods to the chunked loop by the density rule; -1 (runtime default) the launcher
                                              times the plausible bit-identical variants (cursor, flat, chunked) over its first launches and keeps the fastest */
  int bulk_tile;                           /* 1 (default; ABL_CUDA_BULK=0 clears it): sparse 2-D for-near loops run from a shared-memory tile whose rows
                                              the TMA engine copies (cp.async.bulk + mbarrier, ABL_MODE 7) when the launcher's rule picks the flat loop and
                                              a tile entry has 32 bytes or more; 2 (ABL_CUDA_BULK=2): for every entry size */
  /* Single-precision shadow of the neighbour pool's positions (double-precision builds, steps registered with
   * abl_step_desc.shadow != 0; NULL: none, e.g. ABL_CUDA_DENSE=0): one float4 (x, y, z, 0) per agent in pool order,
   * refreshed by the runtime whenever the pool was re-binned or its positions were rewritten, and the largest
   * coordinate magnitude in it (float bits).  Dense for-near loops (the launcher's chunked case) pre-filter their
   * candidates on it and re-test the survivors exactly (ABL_MODE 8). */
  const void *nbr_shadow;
  const unsigned *nbr_shadow_max;
  int probe;   /* != 0: launch nothing; return 1 if this launch would use the shadow (ABL_MODE 8), 2 if it would also
                  use the pre-filter scratch below (ABL_MODE 9), else 0 */
  /* Scratch of the split pre-filter (ABL_MODE 9; NULL: none): the step's pre-filter kernel leaves, per agent of the
   * launched range and column-wise with stride pf_stride, up to ABL_SHADOW_WORDS acceptance masks (pf_masks), the
   * first pool index of up to ABL_SHADOW_ROWS row ranges (pf_rows), the map of the words that start a new range
   * (pf_sbits) and a header (pf_hdr: number of words; bit 31: the candidates did not fit); the step kernel, launched
   * right behind it, walks them. */
  unsigned *pf_masks, *pf_rows, *pf_hdr;
  unsigned long long *pf_sbits;
  unsigned pf_stride;
  /* > 0 (set by the generated launcher for dense for-near loops of reach 1): the squared radius bound; the neighbour
   * iterator then narrows every row of cells along x to the cells within reach of the agent (abl_device.cuh: row_reach) */
  double row_cull;
  int pdl;                                 /* bit 0: launch with programmatic stream serialization (the kernel calls cudaGridDependencySynchronize first); bit 1: the kernel triggers the launch of its successor at once (ABL_CUDA_PDL_TRIGGER); bit 2: bulk-tile kernels wait for their tile with a suspend-time hint (ABL_CUDA_MBAR_HINT) */
  void *stream;                            /* cudaStream_t */
  /* Cached neighbour lists (steps registered with abl_step_desc.nlist != 0: neither pool of the
   * for-near loop ever moves, so the set of accepted candidates of every agent is the same in
   * every timestep).  nlist_phase 1: the kernel only counts the accepted candidates of each
   * agent into nlist_cnt[i] and folds the maximum into *nlist_max; phase 2: it stores their pool
   * indices, in visiting order, at nlist_idx[k * nlist_stride + i] (k-major: the lanes of a warp
   * read consecutive words); phase 0 with nlist_cnt != NULL: the step kernel walks that list
   * instead of the cell rows — no position loads, no filter.  Same candidates in the same order:
   * bit-identical results. */
  int nlist_phase;
  unsigned *nlist_cnt;
  unsigned *nlist_idx;
  unsigned nlist_stride;
  unsigned *nlist_max;
} abl_step_launch;

typedef int (*abl_step_launcher)(const abl_step_launch *args);

/* Static facts about one `step` function — the analysis products a GPU backend consumes
 * (reference FunctionDeclaration::accessedAgent/accessedMembers/usesRuntimeRemoval/
 * runtimeAddedAgent, src/AST.hpp:536-544). */
typedef struct {
  const char *name;
  int self_pool;
  int nbr_pool;              /* -1: no for-near loop */
  double radius;             /* folded for-near radius */
  uint32_t written_members;  /* bit m set: member m of `out` is assigned */
  int uses_removal;
  int added_pool;            /* -1: no run-time add() */
  abl_step_launcher launch;
  int shadow;                /* 1: the step's kernels include the shadow pre-filter variant (ABL_MODE 8) */
  int nlist;                 /* 1: list kernels were generated and no step function of the model moves, adds or
                                removes agents of either pool of the for-near loop (`-C cuda.nlist=true`) */
} abl_step_desc;

int abl_cuda_register_step(abl_runtime *rt, const abl_step_desc *desc, int *step);
/* One parallel step function over its pool: (re)bin the neighbour pool if its positions
 * changed, launch the kernel, flip the written columns, compact removals / append adds.
 * Replaces one `#pragma omp parallel for` block + buffer swap of the reference simulate
 * loop (src/backend/CPrinter.cpp:207-228). */
int abl_cuda_step(abl_runtime *rt, int step);
/* Marks the start of a new timestep (advances the RNG counter, starts the timer used by
 * getLastExecTime()). */
int abl_cuda_begin_timestep(abl_runtime *rt);
int abl_cuda_end_timestep(abl_runtime *rt);
int abl_cuda_synchronize(abl_runtime *rt);

/* ---- reductions for `sequential step` functions (semantics: reference
 *      src/backend/MasonPrinter.cpp:623-682; absent from the `c` backend) ---------------- */
int abl_cuda_count(abl_runtime *rt, int pool, int *result);
int abl_cuda_count_member_int(abl_runtime *rt, int pool, int member, int value, int *result);
int abl_cuda_count_member_float(abl_runtime *rt, int pool, int member, double value, int *result);
int abl_cuda_sum_int(abl_runtime *rt, int pool, int member, int *result);        /* int and bool */
int abl_cuda_sum_float(abl_runtime *rt, int pool, int member, int component, double *result);
int abl_cuda_last_exec_time(abl_runtime *rt, double *seconds);

/* ---- introspection used by tests / bench (binning is an internal stage of abl_cuda_step) */
/* Forces binning of `pool` now. */
int abl_cuda_bin(abl_runtime *rt, int pool);
/* Copies the pool's current cell_start (n_cells+1 entries) and agent ids in device order. */
int abl_cuda_debug_binning(abl_runtime *rt, int pool, unsigned *cell_start, size_t n_cells_plus_1,
                           unsigned *ids, size_t n_ids);
int abl_cuda_grid_cells(abl_runtime *rt, unsigned *n_cells, int n_cell_axis[3]);
/* Device time of the most recent abl_cuda_step / abl_cuda_bin stages, CUDA events, ms. */
typedef struct {
  float bin_ms;
  float kernel_ms;
  float commit_ms;
  unsigned launches;       /* kernels launched by the runtime since creation */
} abl_step_timing;
int abl_cuda_enable_timing(abl_runtime *rt, int on);
/* Runs step function `step` once (like abl_cuda_step) and, before the launch that counts, times `reps` further
 * launches of its kernel on the same input buffers between two events on the runtime's stream: the average
 * duration of ONE launch, without the idle time an event between two kernels of the per-step chain adds.
 * *ms_per_launch = 0 when the step cannot be repeated (it adds or removes agents, slab decomposition, empty pool). */
int abl_cuda_time_kernel(abl_runtime *rt, int step, int reps, float *ms_per_launch);
int abl_cuda_last_timing(abl_runtime *rt, abl_step_timing *t);
void *abl_cuda_stream(abl_runtime *rt);

/* ---- multi-GPU slab decomposition (one runtime per GPU, one process per GPU) ----------
 * The population is split into slabs of whole cell layers along the slowest-varying grid
 * axis.  Each runtime owns the agents whose cell layer lies in [layer_begin, layer_end) and
 * keeps read-only ghost copies of the adjacent layer(s) on each side.  The transport is
 * supplied by the caller as NCCL (abl_cuda_comm_init_nccl) — the runtime issues grouped
 * ncclSend/ncclRecv on its own stream. */
int abl_cuda_nccl_unique_id(void *id128);   /* 128-byte ncclUniqueId, to be broadcast by the caller */
int abl_cuda_comm_init_nccl(abl_runtime *rt, const void *id128, int rank, int world);
/* layer_bounds[0..n_slabs]: slab s owns cell layers [layer_bounds[s], layer_bounds[s+1]) along
 * the slab axis; this runtime is slab `my_slab`.  Slabs are neighbours on a ring (periodic
 * worlds: agents that leave through one end re-enter at the other). */
int abl_cuda_set_slab(abl_runtime *rt, const int *layer_bounds, int n_slabs, int my_slab);
int abl_cuda_slab_axis_layers(abl_runtime *rt, int *n_layers);
/* After a step (or upload): send agents that left the slab to the neighbour ranks, receive
 * arrivals, then refresh ghost layers of `pool`.  Collective over neighbouring ranks. */
int abl_cuda_exchange(abl_runtime *rt, int pool);
int abl_cuda_owned_size(abl_runtime *rt, int pool, size_t *n);
/* Direct halo transport over peer memory.  halo_setup allocates this runtime's receive area for
 * `pool` (capacity_records per direction, 0 = default) and returns its 64-byte CUDA IPC handle;
 * the caller passes every rank its ring neighbours' handles (halo_connect), after which step
 * kernels write halo and migrating records straight into the neighbour's memory over NVLink and
 * abl_cuda_step needs neither NCCL nor a host synchronisation for the exchange.  Runtimes in one
 * process are connected with halo_connect_local instead. */
#define ABL_IPC_HANDLE_BYTES 64
int abl_cuda_halo_setup(abl_runtime *rt, int pool, size_t capacity_records, void *handle_out);
int abl_cuda_halo_connect(abl_runtime *rt, int pool, const void *lower_handle, const void *upper_handle);
int abl_cuda_halo_connect_local(abl_runtime *rt, int pool, abl_runtime *lower, abl_runtime *upper);
/* In-process transport between runtimes driven by one host thread (several slabs on one
 * GPU; used by the single-GPU tests of the decomposition).  With local peers set the caller
 * runs a step on every slab, then exchange_begin on every slab, then exchange_end on every
 * slab; abl_cuda_step does not exchange by itself in this mode. */
int abl_cuda_set_local_peers(abl_runtime *rt, abl_runtime *lower, abl_runtime *upper);
/* ---- scalable upload + one-process multi-GPU driver -------------------------------------------------
 * Under slab decomposition no participant has to touch the whole population: slab r uploads the r-th part
 * of the host records BY INDEX (abl_cuda_partition_upload: H2D, owner of every record from its position,
 * records grouped by owner in device memory), the parts are exchanged between the devices (NCCL all-to-all
 * under torchrun, peer copies inside one process) and adopted (abl_cuda_adopt_records).  Transit record =
 * host record + u32 id (abl_cuda_transit_record_bytes). */
int abl_cuda_transit_record_bytes(abl_runtime *rt, int pool, size_t *bytes);
int abl_cuda_partition_upload(abl_runtime *rt, int pool, const void *host_aos, size_t n, unsigned first_id,
                              void *dev_out, unsigned *counts /* [n_slabs] */);
int abl_cuda_adopt_records(abl_runtime *rt, int pool, const void *dev_records, size_t n, unsigned next_id);

/* `simulate(timesteps)` of a generated program on n_gpus devices of this machine (`-C cuda.gpus=N`,
 * ABL_CUDA_GPUS=N): one runtime and one host thread per device, slabs of whole cell layers along the
 * slowest axis, halo cells and migrating agents written into the neighbour's memory by the step kernels
 * (peer access over NVLink), scalable upload as above, download merged by agent id into the host arrays.
 * The reference's precedent for putting the decomposition into the generated program:
 * src/backend/DMasonPrinter.cpp:30-32 (dmason.grid_rows/cols), src/backend/FlameMainPrinter.cpp:40-41.
 * setup(rt): registers pools and steps (abl_model_setup); timestep(rt): one timestep of the parallel step
 * functions, callable from several threads at once, one runtime each.  Models whose timestep needs the
 * host between step functions (run-time add(), a sequential step) are refused. */
typedef struct {
  int n_types;
  const int *pool;          /* [n_types] pool index of every agent type (the same in every runtime) */
  void *const *data;        /* [n_types] host record arrays */
  const size_t *len;        /* [n_types] records in them */
  const size_t *stride;     /* [n_types] */
  const int *has_position;  /* [n_types] agent types without a position are replicated on every device */
} abl_group_population;
int abl_cuda_group_simulate(const abl_config *cfg, int n_gpus, int (*setup)(abl_runtime *rt),
                            int (*timestep)(abl_runtime *rt), int timesteps, const abl_group_population *pop);

int abl_cuda_exchange_begin(abl_runtime *rt, int pool);
int abl_cuda_exchange_end(abl_runtime *rt, int pool);


/* ---- run-time add() / removeCurrent() under slab decomposition ---------------------------
 * (semantics of add/remove: reference src/backend/MasonPrinter.cpp:159-178, 358-367; they are
 * absent from the `c` backend, CBackend.cpp:30-32.)  Removal is local to the owning slab:
 * abl_cuda_step compacts the owned range and refreshes the neighbours' ghosts.  New agents get
 * the ids an undecomposed run gives them: next_id + rank of the parent's id among the parents
 * of ALL slabs.  A step function that adds agents is therefore left open by abl_cuda_step; the
 * caller fetches the local parents' ids (ascending), combines them across the slabs and hands
 * back, for every local parent, its rank among all parents plus the global number of adds:
 *     abl_cuda_step(rt, s);                              on every slab
 *     abl_cuda_pending_adds(rt, &open, &m);              open == 1 after an adding step
 *     abl_cuda_pending_add_parents(rt, ids, m);
 *     ... all-gather of ids, rank of each local id in the sorted union ...
 *     abl_cuda_resolve_adds(rt, ranks, total);           appends, removes, exchanges
 * Any other abl_cuda_step while a step is open fails with ABL_ERR_STATE. */
int abl_cuda_pending_adds(abl_runtime *rt, int *open, unsigned *count);
int abl_cuda_pending_add_parents(abl_runtime *rt, unsigned *parent_ids, size_t capacity);
int abl_cuda_resolve_adds(abl_runtime *rt, const unsigned *global_rank, unsigned global_total);

/* Reductions (count / sum / count_member) cover the owned agents of one runtime under slab
 * decomposition.  A hook installed here is called with the rank-local values and replaces them
 * by the sums over all slabs (e.g. an all-reduce of the calling harness); it returns 0 on
 * success.  Without a hook the caller combines the per-rank results itself. */
typedef int (*abl_reduce_hook)(void *user, long long *ints, int n_ints, double *reals, int n_reals);
int abl_cuda_set_reduce_hook(abl_runtime *rt, abl_reduce_hook hook, void *user);

#ifdef __cplusplus
}
#endif
#endif /* ABL_CUDA_H */
