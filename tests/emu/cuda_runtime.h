// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE, not part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the *generated* device program
// (model_kernels.cu + asset/cuda/abl_device.cuh) for the host, so that the `-m "not gpu"`
// tests can execute the very kernels the code generator prints — one simulated thread after
// the other — and compare them bit for bit with the parity oracle.  This checks the code
// generator's kernel semantics (candidate order, filter, body, stores, fused histogram) in a
// container without a GPU; timing, memory behaviour and everything that needs threads of a
// block to cooperate (the shared-memory tile kernels, the slab publish) are out of its reach
// and stay with the `-m gpu` tests.
//
// Execution model: cudaLaunchKernelEx runs kernel(args...) for blockIdx.x = 0..grid-1 and
// threadIdx.x = 0..block-1 sequentially.  That is faithful for kernels whose threads only meet
// through atomics on global memory (ABL_MODE 0 and 1, the flat loop); __syncthreads() aborts.
// Arithmetic: compile with -ffp-contract=off; x86-64 SSE2 then rounds like the device code
// compiled with -fmad=false (IEEE add/mul/div/sqrt; transcendental functions may differ).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 1 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes;
  cudaStream_t stream;
  cudaLaunchAttribute *attrs;
  unsigned numAttrs;
};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

// ---- the one simulated thread --------------------------------------------------------------
struct emu_idx { unsigned x, y, z; };
extern emu_idx threadIdx, blockIdx, blockDim, gridDim;
extern unsigned long long emu_threads_run;
extern char emu_last_kernel[256];   // mangled name of the kernel launched last (shows the ABL_MODE instance)

// events read a simulated clock that every launch advances by emu_cost_ms[ABL_MODE of the kernel]
// (set from Python), so the launchers' run-time tuner can be driven either way
struct emu_event { float t; };
typedef emu_event *cudaEvent_t;
extern float emu_clock_ms, emu_cost_ms[8];
extern int emu_fail_mode;   // launches of this ABL_MODE fail (-1: none)
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0.f}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = emu_clock_ms; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = b->t - a->t; return cudaSuccess; }

static inline void cudaGridDependencySynchronize() {}
static inline void cudaTriggerProgrammaticLaunchCompletion() {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __syncthreads() {
  fprintf(stderr, "emu: __syncthreads() reached — block-cooperative kernels cannot run on the sequential emulator\n");
  abort();
}
extern unsigned long long emu_ldg_count;   // read-only loads executed (a proxy for candidates visited)
template <typename T> static inline T __ldg(const T *p) { emu_ldg_count++; return *p; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicExch(unsigned *p, unsigned v) { unsigned o = *p; *p = v; return o; }
static inline unsigned atomicMin(unsigned *p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
// every simulated thread is alone in its warp, at its natural lane
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T> static inline T __shfl_sync(unsigned, T v, unsigned) { return v; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return v; }

// CUDA's integer min/max overload set
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }

template <typename... KArgs, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*kernel)(KArgs...), Args... args) {
  blockDim.x = cfg->blockDim.x; blockDim.y = blockDim.z = 1;
  gridDim.x = cfg->gridDim.x; gridDim.y = gridDim.z = 1;
  threadIdx.y = threadIdx.z = blockIdx.y = blockIdx.z = 0;
  Dl_info info;
  emu_last_kernel[0] = 0;
  if (dladdr(reinterpret_cast<void *>(kernel), &info) && info.dli_sname) {
    strncpy(emu_last_kernel, info.dli_sname, sizeof emu_last_kernel - 1);
    emu_last_kernel[sizeof emu_last_kernel - 1] = 0;
  }
  if (const char *m = strstr(emu_last_kernel, "ILi")) {
    const int mode = atoi(m + 3);
    if (mode == emu_fail_mode) return 1;   // simulated launch failure (e.g. too much dynamic shared memory)
    emu_clock_ms += emu_cost_ms[mode >= 0 && mode < 8 ? mode : 0];
  }
  for (unsigned b = 0; b < cfg->gridDim.x; b++) {
    blockIdx.x = b;
    for (unsigned t = 0; t < cfg->blockDim.x; t++) {
      threadIdx.x = t;
      kernel(args...);
      emu_threads_run++;
    }
  }
  return cudaSuccess;
}
