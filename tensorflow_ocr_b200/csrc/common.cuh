// common.cuh — shared helpers for libplhead.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <utility>

#include "../../include/plhead.h"

namespace plh {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

extern std::atomic<long long> g_launch_count;

inline int launch_status() {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? PLH_OK : (int)e;
}

// Every kernel of the hot chains is launched with programmatic dependent launch (PDL): the next
// kernel's CTAs are scheduled while the previous kernel drains, and wait at
// cudaGridDependencySynchronize() (the first statement of each kernel) until the predecessor has
// completed and flushed.  This removes most of the launch latency between the ~10 small dependent
// kernels of a head step; PLH_NO_PDL=1 in the environment disables it (A/B measurements).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("PLH_NO_PDL") ? 0 : 1;
  return v == 1;
}

// launch_plain: ordinary stream order (no early launch of this kernel's CTAs under its predecessor)
template <bool PDL = true, typename... KArgs, typename... Args>
inline int launch_as(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (PDL && pdl_enabled()) ? 1 : 0;
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  return e == cudaSuccess ? PLH_OK : (int)e;
}
template <typename... KArgs, typename... Args>
inline int launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_as<true>(kern, grid, block, smem, s, static_cast<Args&&>(args)...);
}
template <typename... KArgs, typename... Args>
inline int launch_plain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  return launch_as<false>(kern, grid, block, smem, s, static_cast<Args&&>(args)...);
}

// Optional device-side timeline (compile with -DPLH_TIMELINE): first-CTA start / last-CTA end of every
// hot-chain kernel in %globaltimer nanoseconds, read back with plh_timeline_read().
#ifdef PLH_TIMELINE
static __device__ unsigned long long* g_tlp;  // one copy per translation unit, set by tl_set_ptr()
static inline int tl_set_ptr(unsigned long long* p) { return (int)cudaMemcpyToSymbol(g_tlp, &p, sizeof(p)); }
__device__ __forceinline__ void tl_start(int id) {
  if (threadIdx.x == 0 && g_tlp) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMin(&g_tlp[2 * id], t);
  }
}
__device__ __forceinline__ void tl_end(int id) {
  if (threadIdx.x == 0 && g_tlp) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMax(&g_tlp[2 * id + 1], t);
  }
}
#else
__device__ __forceinline__ void tl_start(int) {}
__device__ __forceinline__ void tl_end(int) {}
#endif

// first statement of every PDL-launched kernel
__device__ __forceinline__ void pdl_wait_and_release() {
  cudaGridDependencySynchronize();             // predecessor complete, its writes visible
  cudaTriggerProgrammaticLaunchCompletion();   // let the successor start scheduling
}
// Split form for long kernels on few SMs: a successor released at the top would park its CTAs on the
// idle SMs for the whole kernel and keep the other branch's kernels off them, so release late.
__device__ __forceinline__ void pdl_wait() { cudaGridDependencySynchronize(); }
__device__ __forceinline__ void pdl_release() { cudaTriggerProgrammaticLaunchCompletion(); }

// Dynamic shared memory above 48 KB is an opt-in per (kernel, DEVICE): one process may drive several
// devices (one host thread each), so the bookkeeping is per device, not per process.  `state` is one
// static object per launch site; the largest size granted so far is remembered per device.
struct SmemOptIn {
  std::atomic<int> granted[64];
  SmemOptIn() { for (auto& g : granted) g.store(0, std::memory_order_relaxed); }
};
template <typename K>
inline int ensure_dynamic_smem(SmemOptIn& state, K kern, size_t bytes) {
  if (bytes <= 48 * 1024) return PLH_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  const bool tracked = dev >= 0 && dev < 64;
  if (tracked && state.granted[dev].load(std::memory_order_acquire) >= (int)bytes) return PLH_OK;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);  // idempotent; benign if raced
  if (e != cudaSuccess) return (int)e;
  if (tracked) state.granted[dev].store((int)bytes, std::memory_order_release);
  return PLH_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 128-bit streaming loads/stores: inputs are read once and outputs written once,
// so keep them out of L1 (L2 still caches them for the later passes).
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream2(float2* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// The pixel score the OHEM selection ranks: softmax probability of the NEGATIVE
// class with TensorFlow's formula exp(x-max)/sum (slim.softmax, nets/model.py:216-217).
// Accurate expf + IEEE division on purpose: every kernel that needs the score calls
// this one function, so the mask is self-consistent bit for bit.
__device__ __forceinline__ float neg_class_score(float x0, float x1) {
  const float m = fmaxf(x0, x1);
  const float e0 = expf(x0 - m);
  const float e1 = expf(x1 - m);
  return __fdiv_rn(e0, __fadd_rn(e0, e1));
}

// Deterministic block-wide sum of per-CTA partial vectors.  Thread t owns value t % NV and
// CTAs g, g+G, ... (g = t / NV, G = T / NV groups).  Loads are issued in independent batches of
// 8 (the L2 latency would otherwise serialise into a long single-CTA tail), added in a fixed
// order into one fp64 accumulator per thread, then summed over the groups in a fixed order.
// Result in s_out[NV] (shared, valid after the trailing __syncthreads).  s_tmp: (T / NV) * NV doubles.
template <int NV, int T>
__device__ __forceinline__ void block_final_reduce(const float* partials, int stride, unsigned ncta, double* s_out,
                                                   double* s_tmp) {
  constexpr int G = T / NV;
  const int g = threadIdx.x / NV, i = threadIdx.x - g * NV;
  if (g < G) {
    double acc = 0.0;
    for (unsigned c0 = g; c0 < ncta; c0 += G * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned c = c0 + u * G;
        v[u] = c < ncta ? __ldcg(partials + (size_t)c * stride + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += (double)v[u];
    }
    s_tmp[g * NV + i] = acc;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int gg = 0; gg < G; ++gg) s += s_tmp[gg * NV + threadIdx.x];
    s_out[threadIdx.x] = s;
  }
  __syncthreads();
}

}  // namespace plh
