// tu_tcw.cu -- instantiations and host-side launchers of the warp-specialised, TMA-fed tcgen05 matmuls (bgpt_tcw.cuh)
#include "bgpt_tcw.cuh"
#include "bgpt_tu.h"
#include <cstdio>
#include <cstdlib>

// cuTensorMapEncodeTiled comes from the driver; the library links the CUDA runtime statically and no libcuda, so the entry point is
// looked up through the runtime
typedef CUresult (*tw_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tw_encode_fn tw_get_encode() {
    static tw_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void * p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (tw_encode_fn) p;
        else cudaGetLastError();
    }
    return fn;
}
// fp16 matrix [rows][K] with `pitch` bytes between rows; box = 64 elements (128 bytes, swizzled) x box_rows; rows beyond `rows` read as zero
static bool tw_encode(CUtensorMap * m, const void * base, uint64_t K, uint64_t rows, uint64_t pitch, uint32_t box_rows) {
    tw_encode_fn enc = tw_get_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = { K, rows };
    const cuuint64_t strides[1] = { pitch };
    const cuuint32_t box[2] = { TW_BK, box_rows };
    const cuuint32_t es[2] = { 1, 1 };
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool tw_first_use(unsigned long long & mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    if (mask & (1ULL << dev)) return false;
    mask |= 1ULL << dev;
    return true;
}
static int tw_tpt() {                                  // tokens per epilogue thread of k_tcw_exact: 4 (16 epilogue warps, default) or 8
    static int v = 0;
    if (!v) { const char * e = getenv("BGPT_TCW_TPT"); v = (e && atoi(e) == 8) ? 8 : 4; }
    return v;
}
static bool tw_f16_split() {                           // k_tcw_f16: per-stage accumulators + rounded f32 adds (default) or one accumulator
    static int v = -1;
    if (v < 0) { const char * e = getenv("BGPT_F16_TC_SPLIT"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v != 0;
}
template <int TPT> static const void * tw_exact_fn_t(int wtype) {
    switch (wtype) {
        case BG_Q4_0: return (const void *) k_tcw_exact<BG_Q4_0, TPT>; case BG_Q4_1: return (const void *) k_tcw_exact<BG_Q4_1, TPT>;
        case BG_Q5_0: return (const void *) k_tcw_exact<BG_Q5_0, TPT>; case BG_Q5_1: return (const void *) k_tcw_exact<BG_Q5_1, TPT>;
        case BG_Q8_0: return (const void *) k_tcw_exact<BG_Q8_0, TPT>;
    }
    return nullptr;
}
static const void * tw_exact_fn(int wtype, int tpt) { return tpt == 8 ? tw_exact_fn_t<8>(wtype) : tw_exact_fn_t<4>(wtype); }
static const size_t TWX_SMEM = 1024 + (size_t) TWX_STAGES * TWX_STAGE_BYTES + 256;
static const size_t TWH_SMEM = 1024 + (size_t) TWH_STAGES * TWH_STAGE_BYTES + 256;
static void tw_init_attrs() {
    static unsigned long long done = 0;
    if (!tw_first_use(done)) return;
    for (int t : { BG_Q4_0, BG_Q4_1, BG_Q5_0, BG_Q5_1, BG_Q8_0 })
        for (int tpt : { 4, 8 }) cudaFuncSetAttribute(tw_exact_fn(t, tpt), cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TWX_SMEM);
    cudaFuncSetAttribute((const void *) k_tcw_f16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TWH_SMEM);
    cudaFuncSetAttribute((const void *) k_tcw_f16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TWH_SMEM);
    cudaGetLastError();
}

bool bgpt_tcw_available() { return tw_get_encode() != nullptr; }

cudaError_t bgpt_tcw_decode(int wtype, cudaStream_t s, const GemvArgs & g, int K, void * out16, float * sw, float * mw) {
    TwDecodeArgs a{};
    for (int i = 0; i < 3; i++) a.W[i] = g.W[i];
    a.rows_per = g.rows_per; a.M = g.M; a.G = g.G; a.stride = g.stride; a.off_qh = g.off_qh; a.off_d = g.off_d; a.off_m = g.off_m;
    a.out = (__half *) out16; a.sw = sw; a.mw = mw; a.K = K;
    const long long threads = (long long) g.M * g.G * 8;
    const unsigned blocks = (unsigned) ((threads + 255) / 256);
    switch (wtype) {
        case BG_Q4_0: k_tcw_decode<BG_Q4_0><<<blocks, 256, 0, s>>>(a); break;
        case BG_Q4_1: k_tcw_decode<BG_Q4_1><<<blocks, 256, 0, s>>>(a); break;
        case BG_Q5_0: k_tcw_decode<BG_Q5_0><<<blocks, 256, 0, s>>>(a); break;
        case BG_Q5_1: k_tcw_decode<BG_Q5_1><<<blocks, 256, 0, s>>>(a); break;
        case BG_Q8_0: k_tcw_decode<BG_Q8_0><<<blocks, 256, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t bgpt_tcw_expand(cudaStream_t s, const GemvArgs & g, int K, int n_pad, void * out16, float * sa, float * ss, int hasm) {
    TwExpandArgs a{};
    a.act = g.act; a.act_bytes = g.act_bytes; a.off_dd = g.off_dd; a.off_s = g.off_s; a.G = g.G; a.K = K;
    a.n = g.n; a.n_pad = n_pad; a.tok0 = g.tok0; a.out = (__half *) out16; a.sa = sa; a.ss = ss; a.hasm = hasm;
    const long long threads = (long long) n_pad * g.G * 8;
    k_tcw_expand<<<(unsigned) ((threads + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

// Y = W . act through k_tcw_exact.  a16 / sw / mw: the prompt-operand cache of the (stacked) weight; b16 / sa / ss: the expanded
// activations of this matmul (n_pad = tokens rounded up to 16).  Needs M % 128 == 0 and K % 64 == 0.
cudaError_t bgpt_tcw_gemm_exact(int wtype, cudaStream_t s, const void * a16, const float * sw, const float * mw, const void * b16,
                                const float * sa, const float * ss, int M, int K, int n, int tok0, int n_pad, const Epi & epi, int n_sm) {
    tw_init_attrs();
    const int tpt = tw_tpt();
    const void * fn = tw_exact_fn(wtype, tpt);
    if (!fn || M % TW_ROWS || K % TW_BK || n_pad % TWX_TOK) return cudaErrorInvalidValue;
    TwxArgs P{};
    if (!tw_encode(&P.tmA, a16, (uint64_t) K, (uint64_t) M, (uint64_t) K * 2, TW_ROWS)) return cudaErrorInvalidValue;
    if (!tw_encode(&P.tmB, b16, (uint64_t) K, (uint64_t) n_pad * 4, (uint64_t) K * 2, TWX_BROWS)) return cudaErrorInvalidValue;
    P.sw = sw; P.mw = mw; P.sa = sa; P.ss = ss;
    P.M = M; P.n = n; P.n_pad = n_pad; P.tok0 = tok0; P.nkb = K / TW_BK;
    P.n_row_tiles = M / TW_ROWS; P.n_tok_tiles = n_pad / TWX_TOK; P.epi = epi;
    { static const int dbg = getenv("BGPT_TCW_DBG") ? atoi(getenv("BGPT_TCW_DBG")) : 0; P.dbg = dbg; }
    const int tiles = P.n_row_tiles * P.n_tok_tiles;
    void * args[] = { &P };
    return cudaLaunchKernel(fn, dim3((unsigned) (tiles < n_sm ? tiles : n_sm)), dim3(64 + 32 * 4 * (TWX_TOK / tpt)), args, TWX_SMEM, s);
}

cudaError_t bgpt_tcw_act_h(cudaStream_t s, const GemvArgs & g, int K, void * out16) {
    const int cnt = g.n - g.tok0;
    const long long threads = (long long) cnt * (K / 8);
    k_tcw_act_h<<<(unsigned) ((threads + 255) / 256), 256, 0, s>>>(g.act, g.act_bytes, K, g.n, g.tok0, (__half *) out16);
    return cudaGetLastError();
}

// Y = W . act for F16 weights through k_tcw_f16: W[i] are the device rows (fp16, K permuted inside the row, `stride` bytes apart),
// b16 the activations as fp16 [n - tok0][K] in the same order.  Needs rows_per % 128 == 0 and K % 64 == 0.
cudaError_t bgpt_tcw_gemm_f16(cudaStream_t s, const GemvArgs & g, int K, const void * b16, int n_sm) {
    tw_init_attrs();
    if (g.rows_per % TW_ROWS || K % TW_BK) return cudaErrorInvalidValue;
    TwhArgs P{};
    const int nmat = g.M / g.rows_per;
    for (int i = 0; i < 3; i++)
        if (!tw_encode(&P.tmA[i], g.W[i < nmat ? i : 0], (uint64_t) K, (uint64_t) g.rows_per, (uint64_t) g.stride, TW_ROWS)) return cudaErrorInvalidValue;
    const int cnt = g.n - g.tok0;
    if (!tw_encode(&P.tmB, b16, (uint64_t) K, (uint64_t) cnt, (uint64_t) K * 2, TWH_TOK)) return cudaErrorInvalidValue;
    P.rows_per = g.rows_per; P.M = g.M; P.n = g.n; P.tok0 = g.tok0; P.nkb = K / TW_BK;
    P.n_row_tiles = g.M / TW_ROWS; P.n_tok_tiles = (cnt + TWH_TOK - 1) / TWH_TOK; P.epi = g.epi;
    const int tiles = P.n_row_tiles * P.n_tok_tiles;
    void * args[] = { &P };
    return cudaLaunchKernel(tw_f16_split() ? (const void *) k_tcw_f16<true> : (const void *) k_tcw_f16<false>, dim3((unsigned) (tiles < n_sm ? tiles : n_sm)), dim3(TW_THREADS), args, TWH_SMEM, s);
}
