// Gradient all-reduce of the data-parallel training step over NVLink / NVSwitch peer memory (SURVEY.md 8 row
// e-train): the ONE exchange step of the path, written as our own kernels on NVSwitch multicast memory instead of an
// NCCL call.
//
// Why not NCCL here (measured, profiles/r2_dp_variants_n2_16tiles.log): the backward kernels are persistent, one CTA
// per SM with the shared memory full and a static tile schedule.  An NCCL all-reduce needs SMs of its own, so it
// starts only when a kernel drains, and the next persistent kernel then runs with its CTAs on those SMs starting
// late -- with a static schedule the whole kernel finishes late by the all-reduce's duration.  Bucketing, stream
// priorities and reserving SMs all left ~0.5 ms of a 14.2 ms step exposed: the collective's time (124 MB through ring /
// tree kernels, ~0.4 ms) was simply added to the step, overlapped or not.
//
// What runs instead: make the exchange itself ~5x shorter and run it once, after the last gradient is final.  The
// gradient arena of every rank lives in symmetric memory (torch symmetric memory: every rank maps every peer's arena
// and -- on NVSwitch -- one MULTICAST address that stands for all of them).  Rank r owns 1/N of the arena and, for its
// part,
//   * NVLS path:  v = multimem.ld_reduce.add.f32 [mc + i]   (the switch adds the N copies in flight)
//                 multimem.st [mc + i], v                    (the switch writes all N copies)
//     -- per GPU 2 x 124/N MB cross NVLink instead of 2 x 124 (N-1)/N MB, and no SM does an addition;
//   * peer path:  v = g_0[i] + g_1[i] + ... (fixed order, peer loads), then one store per peer.
// Synchronisation is by monotonically increasing step numbers in a small symmetric flag pad (never reset, CUDA-graph
// friendly: the step number is read from device memory):
//   produced[b][k] = step   written to every peer by rank k once its gradients are final (signal kernel)
//   done[b][k]     = step   written to every peer by rank k once its part is reduced and broadcast
// The reduce kernel spins until produced[b][*] >= step, the gate kernel in front of the optimizer until
// done[b][*] >= step; every wait is bounded (trap after UB_WAIT_LIMIT_NS).  All four kernels run on the compute stream,
// so the data-parallel step is the single-GPU schedule plus four launches and captures into one CUDA graph.
#pragma once
#include "ptx.cuh"

namespace ub {

constexpr int AR_MAX_RANKS = 16;
constexpr int AR_MAX_BUCKETS = 32;

struct ArParams {
  float* peer[AR_MAX_RANKS];          // arena base of every rank, as mapped into this process
  unsigned int* flags[AR_MAX_RANKS];  // flag pad of every rank: [2][AR_MAX_BUCKETS][AR_MAX_RANKS] (produced, done)
  float* mc;                          // multicast address of the arenas (NVLS) or nullptr
  const unsigned int* step;           // device counter, bumped once per training step
  unsigned int* done_ctr;             // local: CTAs of the running reduce kernel that have finished
  int rank, world, bucket;
  long long begin, end;               // element range of the bucket (multiples of 4)
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// bounded spin until *p >= want (wrap-safe comparison of step numbers)
__device__ __forceinline__ void ar_wait_flag(const unsigned int* p, unsigned int want, unsigned int tag) {
  if ((int)(ld_acquire_sys(p) - want) >= 0) return;
  const uint64_t t0 = globaltimer_ns();
  unsigned int spins = 0;
  while ((int)(ld_acquire_sys(p) - want) < 0) {
    __nanosleep(200);
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > UB_WAIT_LIMIT_NS) {
      atomicExch(&g_device_error, 0xA0000000u | (tag << 16) | (blockIdx.x & 0xffff));
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ unsigned int* ar_flag(unsigned int* pad, int kind, int bucket, int rank) {
  return pad + ((kind * AR_MAX_BUCKETS + bucket) * AR_MAX_RANKS + rank);
}

// "my gradients of bucket b are final": one thread per peer
__global__ void ar_signal_kernel(const ArParams p) {
  const int k = threadIdx.x;
  if (k < p.world) {
    __threadfence_system();
    st_release_sys(ar_flag(p.flags[k], 0, p.bucket, p.rank), *p.step);
  }
}

// "every part of every bucket [0, n_buckets) has been reduced and written into my arena": one thread per (bucket, rank)
__global__ void ar_gate_kernel(const ArParams p, int n_buckets) {
  const unsigned int step = *p.step;
  for (int i = threadIdx.x; i < n_buckets * p.world; i += blockDim.x)
    ar_wait_flag(ar_flag(p.flags[p.rank], 1, i / p.world, i % p.world), step, 2);
}

template <bool NVLS>
__global__ void __launch_bounds__(512) ar_reduce_kernel(const ArParams p) {
  const unsigned int step = *p.step;
  if (threadIdx.x < p.world) ar_wait_flag(ar_flag(p.flags[p.rank], 0, p.bucket, threadIdx.x), step, 1);
  __syncthreads();
  // this rank's part of the bucket, in float4 units
  const long long n4 = (p.end - p.begin) >> 2;
  const long long per = (n4 + p.world - 1) / p.world;
  const long long lo = p.begin / 4 + per * p.rank;
  const long long hi = min(p.begin / 4 + n4, lo + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 2 * stride) {
    const long long j = i + stride;
    if (NVLS) {
      // two independent switch round trips in flight per thread
      const float4 a = multimem_ld_reduce_f4(p.mc + 4 * i);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < hi) b = multimem_ld_reduce_f4(p.mc + 4 * j);
      multimem_st_f4(p.mc + 4 * i, a);
      if (j < hi) multimem_st_f4(p.mc + 4 * j, b);
    } else {
      float4 a = ld_relaxed_sys_f4(p.peer[0] + 4 * i);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < hi) b = ld_relaxed_sys_f4(p.peer[0] + 4 * j);
      for (int k = 1; k < p.world; ++k) {      // fixed order: bit-identical on every rank and from run to run
        const float4 c = ld_relaxed_sys_f4(p.peer[k] + 4 * i);
        a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
        if (j < hi) {
          const float4 d = ld_relaxed_sys_f4(p.peer[k] + 4 * j);
          b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
        }
      }
      for (int k = 0; k < p.world; ++k) {
        st_relaxed_sys_f4(p.peer[k] + 4 * i, a);
        if (j < hi) st_relaxed_sys_f4(p.peer[k] + 4 * j, b);
      }
    }
  }
  // all stores of this CTA are out before it counts itself done; the last CTA tells every peer
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(p.done_ctr, 1u) == gridDim.x - 1) {
    *p.done_ctr = 0;
    __threadfence_system();
    for (int k = 0; k < p.world; ++k) st_release_sys(ar_flag(p.flags[k], 1, p.bucket, p.rank), step);
  }
}

// one thread: ++step (first kernel of every training step)
__global__ void ar_bump_kernel(unsigned int* step) { *step += 1; }

}  // namespace ub
