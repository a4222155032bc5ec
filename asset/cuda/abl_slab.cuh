// abl_slab.cuh — slab-decomposition helpers shared by the runtime (abl_runtime.cu) and by
// generated step kernels (abl_device.cuh): routing of an agent to the neighbouring slabs and
// the packed message record format.
#pragma once

#include "abl_cuda.h"

#define ABL_SLAB_HD __host__ __device__ __forceinline__

// Message: 64-byte header, then `rec_words` rows of `capacity` 32-bit words (word-major): row w
// holds word w of every record — the columns of an agent in pool order (1-byte columns widened
// to a word), the agent id last.  Lanes of a warp write consecutive slots, so every store
// instruction of the packing code is one contiguous 128-byte write into the peer's memory.
#define ABL_MSG_HEADER 64u
#define ABL_SENTINEL_ID 0xffffffffu   // padding record: sorted into the trash cell by binning

// Where does an agent whose position now lies in cell layer `layer` have to be copied to?
// bit 0: lower peer, bit 1: upper peer, bit 2: nowhere reachable (moved too far).
//  - still inside the own slab: ghost copies for true neighbours if within `ghost` layers of
//    the respective boundary;
//  - inside a peer's slab: migration (the sender keeps its record; it becomes a ghost or is
//    dropped by key range at the next binning).  Peers form a ring, so agents teleported from
//    one end of a periodic world to the other reach the slab over there.
ABL_SLAB_HD unsigned abl_slab_route(const abl_slab_view &s, int layer) {
  if (layer >= s.begin && layer < s.end) {
    unsigned r = 0;
    if (s.lo_ghost && layer < s.begin + s.ghost) r |= 1u;
    if (s.hi_ghost && layer >= s.end - s.ghost) r |= 2u;
    return r;
  }
  if (layer >= s.lo_begin && layer < s.lo_end) return 1u;
  if (layer >= s.hi_begin && layer < s.hi_end) return 2u;
  return 4u;
}
