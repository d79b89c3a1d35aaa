// tools/cluster_probe.cu -- what thread-block clusters and distributed shared memory cost on this GPU.
// Design input for the generation-5 decode kernel (csrc/bgpt_mega5.cuh): how many clusters of 2/4/8 CTAs of 512 threads with
// ~200 KB of shared memory are co-resident, what a cluster barrier, a DSMEM all-gather (st.async + mbarrier complete_tx) and
// an L2 tagged-word exchange with / without cluster-level forwarding cost per round.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/cluster_probe tools/cluster_probe.cu
//   tools/_build/cluster_probe            (every wait carries a watchdog: a lost signal prints "TIMEOUT", it cannot hang)
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <time.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define NT 512
#define WATCHDOG_CYCLES 400000000LL      // ~0.2 s

__device__ int g_abort = 0;

__device__ __forceinline__ uint32_t s32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t * b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t * b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint64_t * b, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
        if (done) return true;
        if (clock64() - t0 > WATCHDOG_CYCLES || *(volatile int *) &g_abort) { g_abort = 1; return false; }
    }
}
__device__ __forceinline__ void st_async64(uint32_t raddr, uint64_t v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" :: "r"(raddr), "l"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void put64(unsigned long long * p, uint32_t payload, uint32_t tag) {
    const unsigned long long w = ((unsigned long long) tag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ bool poll2(const unsigned long long * p, uint32_t tag, uint32_t & a, uint32_t & b) {
    unsigned long long w0, w1;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
        if ((uint32_t) (w0 >> 32) == tag && (uint32_t) (w1 >> 32) == tag) break;
        if (clock64() - t0 > WATCHDOG_CYCLES || *(volatile int *) &g_abort) { g_abort = 1; a = b = 0; return false; }
    }
    a = (uint32_t) w0; b = (uint32_t) w1;
    return true;
}

// ---- T2: hardware cluster barrier --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) k_clsync(int iters, long long * cyc) {
    cl_sync();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) cl_sync();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

// ---- T3: DSMEM all-gather: every CTA sends `words` 8-byte words to every rank of its cluster (st.async, self included) ----
__global__ void __launch_bounds__(NT, 1) k_allgather(int CL, int words, int iters, long long * cyc, float * sink) {
    extern __shared__ __align__(16) unsigned long long buf[];            // [2][CL][words]
    __shared__ __align__(8) uint64_t mbar[2];
    const uint32_t rank = cl_rank();
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(&mbar[0], (uint32_t) (CL * words * 8)); mbar_expect(&mbar[1], (uint32_t) (CL * words * 8));
    }
    cl_sync();
    unsigned acc = blockIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const int par = i & 1;
        // payload depends on what the previous round delivered: a true dependent chain
        for (int u = threadIdx.x; u < CL * words; u += NT) {
            const int dst = u / words, w = u - dst * words;
            const uint32_t la = s32(buf + ((size_t) par * CL + rank) * words + w);
            st_async64(mapa(la, dst), ((unsigned long long) i << 32) | (acc & 0xffffu), mapa(s32(&mbar[par]), dst));
        }
        if (!mbar_wait(&mbar[par], (i >> 1) & 1)) break;
        // consume: every thread reads one word of every source
        unsigned a2 = 0;
        for (int s = 0; s < CL; s++) a2 += (unsigned) buf[((size_t) par * CL + s) * words + (threadIdx.x % words)];
        acc = a2 + 1;
        __syncthreads();                                                  // everyone has read: the buffer / barrier may be re-armed
        if (threadIdx.x == 0) mbar_expect(&mbar[par], (uint32_t) (CL * words * 8));
    }
    if (threadIdx.x == 0) { cyc[blockIdx.x] = clock64() - t0; sink[blockIdx.x] = (float) acc; }
    cl_sync();                                                            // no CTA exits while peers may still write into it
}

// ---- T4: all-to-all exchange of 1024 f32 through L2 tagged words, R replicas ------------------------------------------------
// MODE 0: every CTA polls all 1024 words itself (generation 4).  MODE 1: rank r of a cluster polls its 1024/CL share and
// forwards the payload to all ranks of the cluster through DSMEM (st.async + complete_tx).
template <int MODE>
__global__ void __launch_bounds__(NT, 1) k_xch(int CL, int R, unsigned long long * xw, int iters, long long * cyc, float * sink) {
    __shared__ __align__(16) float s_x[2][1024];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ float s_out[8];
    const unsigned nC = gridDim.x, c = blockIdx.x;
    const int o0 = (int) ((c * 1024u) / nC), o1 = (int) (((c + 1u) * 1024u) / nC);      // rows this CTA produces
    const uint32_t rank = MODE ? cl_rank() : 0u;
    const int rep = MODE ? (int) (cl_id() % (unsigned) R) : (int) (c % (unsigned) R);
    if (MODE) {
        if (threadIdx.x == 0) {
            mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect(&mbar[0], 4096u); mbar_expect(&mbar[1], 4096u);
        }
        cl_sync();
    }
    float acc = (float) c;
    const long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        const int par = i & 1;
        unsigned long long * X = xw + (size_t) par * R * 1024;
        // produce: gather in shared memory, one warp writes rows x replicas
        if (threadIdx.x < o1 - o0) s_out[threadIdx.x] = acc + (float) threadIdx.x;
        __syncthreads();
        if (threadIdx.x < 32) {
            const int row = threadIdx.x & 7;
            if (row < o1 - o0) for (int r = threadIdx.x >> 3; r < R; r += 4) put64(X + (size_t) r * 1024 + o0 + row, __float_as_uint(s_out[row]), (uint32_t) i);
        }
        bool ok = true;
        if (MODE == 0) {
            uint32_t a, b;
            ok = poll2(X + (size_t) rep * 1024 + 2 * threadIdx.x, (uint32_t) i, a, b);
            *(float2 *) &s_x[par][2 * threadIdx.x] = make_float2(__uint_as_float(a), __uint_as_float(b));
            __syncthreads();
        } else {
            const int share = 1024 / CL;                                                   // words this rank polls
            if (2 * (int) threadIdx.x < share) {
                const int w = (int) rank * share + 2 * threadIdx.x;
                uint32_t a, b;
                ok = poll2(X + (size_t) rep * 1024 + w, (uint32_t) i, a, b);
                const unsigned long long v = ((unsigned long long) b << 32) | a;
                const uint32_t la = s32(&s_x[par][w]), lb = s32(&mbar[par]);
                for (int d = 0; d < CL; d++) st_async64(mapa(la, d), v, mapa(lb, d));
            }
            ok = mbar_wait(&mbar[par], ((i - 1) >> 1) & 1) && ok;
        }
        if (!ok) break;
        // consume: a value that depends on the whole vector's arrival
        acc = s_x[par][(threadIdx.x * 7 + i) & 1023] * 0.5f + 1.0f;
        if (MODE) { __syncthreads(); if (threadIdx.x == 0) mbar_expect(&mbar[par], 4096u); }
    }
    if (threadIdx.x == 0) { cyc[blockIdx.x] = clock64() - t0; sink[blockIdx.x] = acc; }
    if (MODE) cl_sync();
}

static cudaError_t launch(const void * fn, int grid, int CL, size_t smem, void ** args, bool coop) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[2]; int na = 0;
    if (CL > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; na++; }
    if (coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; na++; }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelExC(&cfg, fn, args);
}
static double med_cycles(const long long * d_cyc, int n) {
    std::vector<long long> h(n);
    CK(cudaMemcpy(h.data(), d_cyc, n * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    return (double) h[n / 2];
}
static int aborted() { int v = 0; CK(cudaMemcpyFromSymbol(&v, g_abort, sizeof v)); if (v) { int z = 0; CK(cudaMemcpyToSymbol(g_abort, &z, sizeof z)); } return v; }

// run one kernel with a host-side deadline: a kernel that is still running after 8 s is reported and the process exits
// (destroying the context is the only way to stop it)
static cudaError_t sync_with_deadline(const char * what) {
    cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); CK(cudaEventRecord(ev, 0));
    for (int ms = 0; ms < 8000; ms += 5) {
        cudaError_t e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) { cudaEventDestroy(ev); return e; }
        struct timespec ts = { 0, 5000000 }; nanosleep(&ts, nullptr);
    }
    printf("HUNG: %s did not finish within 8 s -- exiting\n", what); fflush(stdout);
    _exit(3);
}

int main(int argc, char ** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const char * only = argc > 1 ? argv[1] : "";          // "", "T1", "T2", "T3", "T4", "T5"
    auto want = [&](const char * t) { return !only[0] || !strcmp(only, t); };
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz / 1e6;
    printf("device: %s, %d SMs, %.3f GHz, smem optin %zu\n", prop.name, prop.multiProcessorCount, ghz, prop.sharedMemPerBlockOptin);
    long long * d_cyc; float * d_sink; unsigned long long * d_x;
    CK(cudaMalloc(&d_cyc, 1024 * sizeof(long long))); CK(cudaMalloc(&d_sink, 1024 * sizeof(float)));
    CK(cudaMalloc(&d_x, 2 * 32 * 1024 * sizeof(unsigned long long))); CK(cudaMemset(d_x, 0, 2 * 32 * 1024 * sizeof(unsigned long long)));

    // ---- T1: co-resident clusters
    int maxcl[17] = {0};
    for (size_t smem : { (size_t) 100 * 1024, (size_t) 200 * 1024, (size_t) 220 * 1024 }) {
        for (int CL : {1, 2, 4, 8, 16}) {
            const void * fn = (const void *) k_allgather;
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            if (CL > 8) { if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; } }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(CL * 64); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
            printf("T1 max active clusters: cluster %2d x %d threads, %3zu KB smem -> %d clusters = %d CTAs (%s)\n", CL, NT, smem / 1024, n, n * CL, cudaGetErrorString(e));
            cudaGetLastError();
            if (smem == (size_t) 200 * 1024 && n > 0) maxcl[CL] = n;
        }
    }
    const int iters = 2000;
    // ---- T2: cluster barrier
    if (want("T2")) for (int CL : {2, 4, 8}) {
        if (maxcl[CL] <= 0) continue;
        int it = iters; void * args[] = { &it, &d_cyc };
        const int grid = maxcl[CL] * CL;
        cudaError_t e = launch((const void *) k_clsync, grid, CL, 0, args, false);
        if (e == cudaSuccess) e = sync_with_deadline("kernel");
        printf("T2 barrier.cluster arrive+wait, cluster %d, grid %d: %.0f cycles = %.0f ns per barrier (%s)\n", CL, grid, med_cycles(d_cyc, grid) / iters,
               med_cycles(d_cyc, grid) / iters / ghz, cudaGetErrorString(e));
    }
    // ---- T5: cluster + cooperative attribute together
    if (want("T5")) {
        int it = 10; void * args[] = { &it, &d_cyc };
        cudaError_t e = launch((const void *) k_clsync, std::max(1, maxcl[8]) * 8, 8, 0, args, true);
        if (e == cudaSuccess) e = sync_with_deadline("kernel");
        printf("T5 cluster 8 + cooperative launch attribute: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
    }
    // ---- T3: DSMEM all-gather
    if (want("T3")) for (int CL : {2, 4, 8}) {
        if (maxcl[CL] <= 0) continue;
        for (int words : {8, 16, 64, 128}) {
            int it = iters, cl = CL, w = words; void * args[] = { &cl, &w, &it, &d_cyc, &d_sink };
            const size_t smem = (size_t) 2 * CL * words * 8;
            CK(cudaFuncSetAttribute((const void *) k_allgather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            const int grid = maxcl[CL] * CL;
            cudaError_t e = launch((const void *) k_allgather, grid, CL, std::max(smem, (size_t) 200 * 1024), args, false);
            if (e == cudaSuccess) e = sync_with_deadline("kernel");
            const int ab = aborted();
            printf("T3 DSMEM all-gather, cluster %d, %4d B per source, grid %d: %.0f cycles = %.0f ns per round%s (%s)\n", CL, words * 8, grid,
                   med_cycles(d_cyc, grid) / iters, med_cycles(d_cyc, grid) / iters / ghz, ab ? " TIMEOUT" : "", cudaGetErrorString(e));
        }
    }
    // ---- T4: L2 exchange, direct vs cluster-forwarded
    if (want("T4")) for (int R : {8, 16}) {
        for (int grid0 : {148, 128}) {
            int it = iters, cl = 1, r = R; void * args[] = { &cl, &r, &d_x, &it, &d_cyc, &d_sink };
            CK(cudaMemset(d_x, 0, 2 * 32 * 1024 * sizeof(unsigned long long)));
            cudaError_t e = launch((const void *) k_xch<0>, grid0, 1, 0, args, true);
            if (e == cudaSuccess) e = sync_with_deadline("kernel");
            const int ab = aborted();
            printf("T4 L2 exchange, every CTA polls 8 KB, R=%d, grid %d (no clusters): %.0f ns per exchange%s (%s)\n", R, grid0, med_cycles(d_cyc, grid0) / iters / ghz,
                   ab ? " TIMEOUT" : "", cudaGetErrorString(e));
        }
        for (int CL : {2, 4, 8}) {
            if (maxcl[CL] <= 0) continue;
            const int grid = maxcl[CL] * CL;
            for (int mode = 0; mode < 2; mode++) {
                int it = iters, cl = CL, r = R; void * args[] = { &cl, &r, &d_x, &it, &d_cyc, &d_sink };
                CK(cudaMemset(d_x, 0, 2 * 32 * 1024 * sizeof(unsigned long long)));
                cudaError_t e = launch(mode ? (const void *) k_xch<1> : (const void *) k_xch<0>, grid, CL, 0, args, false);
                if (e == cudaSuccess) e = sync_with_deadline("kernel");
                const int ab = aborted();
                printf("T4 L2 exchange, %s, R=%d, cluster %d, grid %d: %.0f ns per exchange%s (%s)\n", mode ? "rank polls 1/CL and forwards by DSMEM" : "every CTA polls 8 KB",
                       R, CL, grid, med_cycles(d_cyc, grid) / iters / ghz, ab ? " TIMEOUT" : "", cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
