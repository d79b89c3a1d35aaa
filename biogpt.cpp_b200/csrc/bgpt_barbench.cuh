// bgpt_barbench.cuh -- micro-benchmark of grid-barrier variants (debug tool; the winner lives in
// bgpt_mega.cuh).  One CTA per SM, 512 threads, `iters` back-to-back barriers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long bb_ld_acquire(const unsigned long long * p) {
    unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long bb_ld_relaxed(const unsigned long long * p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void bb_st_release(unsigned long long * p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void bb_red_release(unsigned long long * p) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" :: "l"(p) : "memory");
}

template <int V>
__device__ __forceinline__ void bb_barrier(unsigned long long * w, unsigned long long k /*1-based*/) {
    const unsigned long long n = gridDim.x;
    __syncthreads();
    if (V == 0) {          // counter, everyone polls the counter
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(w, 1ULL); while (bb_ld_acquire(w) < k * n) { } __threadfence(); }
    } else if (V == 1) {   // counter + separate release flag written by the last arriver
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned long long old = atomicAdd(w, 1ULL);
            if (old + 1 == k * n) bb_st_release(w + 16, k); else while (bb_ld_acquire(w + 16) < k) { }
            __threadfence();
        }
    } else if (V == 2) {   // per-CTA flags, gridDim.x polling threads per CTA
        if (threadIdx.x == 0) bb_st_release(w + 32 + blockIdx.x, k);
        for (unsigned t = threadIdx.x; t < gridDim.x; t += blockDim.x) while (bb_ld_acquire(w + 32 + t) < k) { }
    } else if (V == 3) {   // red (no return) + poll counter
        if (threadIdx.x == 0) { bb_red_release(w); while (bb_ld_acquire(w) < k * n) { } }
    } else if (V == 4) {   // red + relaxed polling, one fence at the end
        if (threadIdx.x == 0) { bb_red_release(w); while (bb_ld_relaxed(w) < k * n) { } __threadfence(); }
    } else if (V == 5) {   // per-CTA flags, one warp polls (lane strides)
        if (threadIdx.x == 0) bb_st_release(w + 32 + blockIdx.x, k);
        if (threadIdx.x < 32) for (unsigned t = threadIdx.x; t < gridDim.x; t += 32) while (bb_ld_relaxed(w + 32 + t) < k) { }
        if (threadIdx.x == 0) __threadfence();
    } else if (V == 6) {   // two-level: 16 groups; group counter, then top counter by group leader; flag fan-out
        const unsigned g = blockIdx.x & 15, gsz = (gridDim.x - g + 15) / 16;
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned long long old = atomicAdd(w + 64 + g * 16, 1ULL);
            if (old + 1 == k * gsz) {
                const unsigned long long o2 = atomicAdd(w, 1ULL);
                if (o2 + 1 == k * 16) bb_st_release(w + 16, k);
            }
            while (bb_ld_acquire(w + 16) < k) { }
            __threadfence();
        }
    }
    __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(512, 1) k_barbench(unsigned long long * w, int iters, float * sink, const float * chase) {
    float acc = 0.f;
    for (int i = 1; i <= iters; i++) {
        bb_barrier<V>(w, (unsigned long long) i);
        if (chase) acc += __ldcg(chase + ((blockIdx.x * 37 + i * 101) & 1023));     // one dependent L2 load per phase
    }
    if (threadIdx.x == 0 && sink) sink[blockIdx.x] = acc;
}
