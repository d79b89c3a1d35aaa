// bgpt_rows.cuh -- persistent kernel for 2..8 token rows (quantised weights, BioGPT-base shapes): the reference's own prompt
// chunking (n_batch = 8, BASELINE configs[2]) and 8 lock-step streams per GPU (configs[3]) as ONE launch per eval.
//
// The fused skinny-batch schedule (bgpt_skinny.cuh) runs such an eval as 195 dependent launches; chained by programmatic
// dependent launch inside a CUDA graph they still cost ~4.3 us each against 1-3 us of work (profiles/README.md, round 2: 848 us
// per 8-row eval, 1043 us per 8-stream step).  Here the same stages run inside one kernel of 128 CTAs x 1024 threads (one CTA
// per SM, all co-resident) and a stage boundary is a counter in L2: producers `red.release` it after their last store, consumers
// poll it with `ld.acquire` (1.25 us per boundary, tools/barrier_bench.py).  The weights of a CTA's next tiles are already in
// flight (TMA bulk copies into a ring of shared-memory slots, issued up to three stages ahead) while it waits.
//
//  stage (per layer)             who                                      waits for          produces
//  A  LayerNorm0 + quantise      CTA r < n: row r                         G of layer l-1     record r (global)                 count n
//  B  q,k,v + bias (+ q scale)   all: rows 8c..8c+7 of q, k and v         A                  q (global), K/V cache rows        count 128
//  C  attention                  CTA u < 16 n: (head u % 16, row u / 16)  B                  blocks 2h, 2h+1 of record r       count 16 n
//  D  out_proj + bias + residual all: rows 8c..8c+7                       C                  x1                                count 128
//  E  LayerNorm1 + quantise      CTA r < n                                D                  record r                          count n
//  F  fc1 + bias + GELU + quant  all: rows 32c..32c+31 = block c of fc2   E                  block c of the d_ff records       count 128
//  G  fc2 + bias + residual      all: rows 8c..8c+7 (K = 4096)            F                  x                                 count 128
//  final LayerNorm (A of "layer" L) + lm_head in 32-row tiles, 331-332 rows per CTA -> logits
//
// A warp works on (weight row, 4 token rows): lanes 8q..8q+7 hold the 8 running sums of token q -- m5_row_dot with one weight row
// and four activation records, i.e. the arithmetic of every other path (bgpt_cuda.cu, "lane order").  LayerNorm / quantise are
// generation 5's (m5_layer_norm8, m5_quant8), attention is k_sk_attn's body.  tests/test_gpu_eval.py compares this kernel with the
// oracle and with the skinny schedule bit for bit.
#pragma once
#include "bgpt_skinny.cuh"
#include "bgpt_mega5.cuh"

#define RW_NT 1024
#define RW_NW (RW_NT / 32)
#define RW_NC 128
#define RW_MAXS 8
#define RW_NST 8              // counters per layer
#define RW_LMRT 32            // lm_head rows per tile
enum { RW_A = 0, RW_B, RW_C, RW_D, RW_E, RW_F, RW_G };

struct RowsParams {
    MegaParams b;
    MegaLayer layers[M5_MAXL];
    const int * tokens;            // device: n token ids
    const DevState * st;           // device: n_past (one CUDA graph serves every position)
    int n, mode;                   // rows; 0: prompt rows of one sequence, 1: lock-step streams (bg_row_info)
    unsigned long long stream_stride;   // floats between the KV caches of two streams
    unsigned int * cnt;            // [(n_layer + 1) * RW_NST] arrival counters, zeroed before the launch
    int * err;                     // [0] 0 or the code of the first wait that timed out, [2..3] watchdog limit in cycles
    uint8_t * rec_d; uint8_t * rec_f;   // activation records [n][actb_d], [n][actb_f]
    long long * trace;             // optional: clock64 stamps of CTA 0, [(n_layer + 1) * RW_NST]
    int nslot, slot_bytes;
    int sm_w, sm_rec, sm_total;
};

__device__ __forceinline__ void rw_wait(const unsigned int * c, unsigned int target, int * err, int code) {
    if (threadIdx.x == 0) {
        unsigned spins = 0; long long t0 = 0;
        for (;;) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if (v >= target) break;
            if ((++spins & 1023u) == 0) { if (t0 == 0) t0 = clock64(); else if (m5_give_up(err, code, t0)) break; }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void rw_arrive(unsigned int * c) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(c) : "memory");
}

// m5_row_dot with a smaller unroll: 1024 threads leave 64 registers
template <int FMT, int G>
__device__ __forceinline__ float rw_row_dot(const uint8_t * wrow, const uint8_t * rec, const M4MM & D) {
    constexpr bool IS8    = (FMT == BG_Q8_0);
    constexpr bool HASQH  = (FMT == BG_Q5_0 || FMT == BG_Q5_1);
    constexpr bool HASM   = (FMT == BG_Q4_1 || FMT == BG_Q5_1);
    constexpr bool HASOFF = (FMT == BG_Q4_0 || FMT == BG_Q5_0);
    const int l = threadIdx.x & 7, j = l & 3, hi = l >> 2, sh = hi * 4;
    float acc = 0.0f, summ = 0.0f;
#pragma unroll 2
    for (int g = 0; g < G; g++) {
        const uint4 wq = IS8 ? *(const uint4 *) (wrow + ((g * 2 + hi) * 4 + j) * 16) : *(const uint4 *) (wrow + (g * 4 + j) * 16);
        uint32_t qh = 0;
        if (HASQH) qh = *(const uint32_t *) (wrow + D.off_qh + (g * 4 + j) * 4);
        const uint4 aw = *(const uint4 *) (rec + (g * 8 + l) * 16);
        int4 an = make_int4(0, 0, 0, 0);
        if (HASOFF) an = *(const int4 *) (rec + D.off_n + (g * 8 + l) * 16);
        const uint2 dh = *(const uint2 *) (wrow + D.off_d + g * 8);
        const float4 da = *(const float4 *) (rec + D.off_dd + g * 16);
        const uint32_t ww[4] = { wq.x, wq.y, wq.z, wq.w }, aa[4] = { aw.x, aw.y, aw.z, aw.w };
        const int nn[4] = { an.x, an.y, an.z, an.w };
        const float dd[4] = { da.x, da.y, da.z, da.w };
        const uint16_t dw[4] = { (uint16_t) (dh.x & 0xFFFF), (uint16_t) (dh.x >> 16), (uint16_t) (dh.y & 0xFFFF), (uint16_t) (dh.y >> 16) };
        uint16_t mw[4] = { 0, 0, 0, 0 }; float ss[4] = { 0.f, 0.f, 0.f, 0.f };
        if (HASM) {
            const uint2 mh = *(const uint2 *) (wrow + D.off_m + g * 8);
            const float4 sa = *(const float4 *) (rec + D.off_s + g * 16);
            mw[0] = (uint16_t) (mh.x & 0xFFFF); mw[1] = (uint16_t) (mh.x >> 16); mw[2] = (uint16_t) (mh.y & 0xFFFF); mw[3] = (uint16_t) (mh.y >> 16);
            ss[0] = sa.x; ss[1] = sa.y; ss[2] = sa.z; ss[3] = sa.w;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t code;
            if (IS8) code = ww[i];
            else {
                code = (ww[i] >> sh) & 0x0F0F0F0Fu;
                if (HASQH) code |= bg_spread4((qh >> (8 * i + sh)) & 0xFu);
            }
            const float pr = (float) __dp4a((int) code, (int) aa[i], nn[i]);
            const float s = __fmul_rn(bg_h2f(dw[i]), dd[i]);
            acc = fmaf(s, pr, acc);
            if (HASM) summ = fmaf(bg_h2f(mw[i]), ss[i], summ);
        }
    }
    float r = __fadd_rn(acc, __shfl_xor_sync(FULLMASK, acc, 4));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(FULLMASK, r, 1));
    if (HASM) r = __fadd_rn(r, summ);
    return r;
}

#define RWSTAMP(L_, ST_) do { if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) P.trace[(L_) * RW_NST + (ST_)] = clock64(); } while (0)

template <int FMT>
__global__ void __launch_bounds__(RW_NT, 1) k_rows(const __grid_constant__ RowsParams P) {
    const MegaParams & p = P.b;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar[M4_NSLOT];
    __shared__ double sredA[RW_NW], sredB[RW_NW];
    __shared__ float sredF[RW_NW];
    __shared__ __align__(16) float s_g[RW_MAXS * 32];             // fc1: [token][row of the CTA] before GELU
    __shared__ __align__(16) float s_out[SK_DK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    const int S = P.n, mode = P.mode;
    const int n_past = P.st->n_past;
    int * const err = P.err;
    uint8_t * s_w = smem + P.sm_w;
    uint8_t * s_rec = smem + P.sm_rec;                             // the stage's activation records; attention: scores | partial sums | V tail
    float * sc = (float *) s_rec;
    float * red = sc + 1024;
    float * tailv = red + 32 * SK_DK;
    const int n_pos = p.n_positions;
    const int o0 = cta * 8;
    const unsigned uC = (unsigned) RW_NC, uc = (unsigned) cta;
    const int v0 = (int) ((uc * (unsigned) p.n_vocab) / uC), v1 = (int) (((uc + 1u) * (unsigned) p.n_vocab) / uC);
    const int n_lm = (v1 - v0 + RW_LMRT - 1) / RW_LMRT;
    const int n_lt = 4 * p.n_layer;
    const int n_tiles = n_lt + n_lm;
    const int lm_t0 = mode == 0 ? S - 1 : 0;                       // the reference returns the last row of a prompt batch (biogpt.cpp:803, 844)
    const int au_h = cta & (M5_NH - 1), au_row = cta >> 4;         // attention unit of this CTA
    const bool has_au = cta < M5_NH * S;

    // ---- weight ring: tile n = layer n >> 2, stage n & 3 (B, D, F, G), then the lm_head tiles
    struct TileSrc { const uint8_t * s0, * s1, * s2; uint32_t b0, b1, b2; };
    auto describe_tile = [&](int n) -> TileSrc {
        TileSrc t{nullptr, nullptr, nullptr, 0u, 0u, 0u};
        if (n >= n_tiles) return t;
        if (n < n_lt) {
            const MegaLayer & L = P.layers[n >> 2];
            const int k = n & 3;
            if (k == 0) {
                const size_t off = (size_t) o0 * p.stride_d;
                t.s0 = L.q_w + off; t.s1 = L.k_w + off; t.s2 = L.v_w + off;
                t.b0 = t.b1 = t.b2 = 8u * (uint32_t) p.stride_d;
            } else if (k == 1) { t.s0 = L.o_w + (size_t) o0 * p.stride_d; t.b0 = 8u * (uint32_t) p.stride_d; }
            else if (k == 2)   { t.s0 = L.fc1_w + (size_t) cta * 32 * p.stride_d; t.b0 = 32u * (uint32_t) p.stride_d; }
            else               { t.s0 = L.fc2_w + (size_t) o0 * p.stride_f; t.b0 = 8u * (uint32_t) p.stride_f; }
        } else {
            const int r = v0 + (n - n_lt) * RW_LMRT;
            t.s0 = p.lm_head + (size_t) r * p.stride_d; t.b0 = (uint32_t) min(RW_LMRT, v1 - r) * (uint32_t) p.stride_d;
        }
        return t;
    };
    const uint64_t pol_w = m4_policy_evict_first();
    auto fire_tile = [&](int n) {
        const TileSrc t = describe_tile(n);
        if (t.b0 == 0) return;
        const int slot = n % P.nslot;
        uint8_t * dst = s_w + (size_t) slot * P.slot_bytes;
        m4_mbar_expect(&mbar[slot], t.b0 + t.b1 + t.b2);
        m4_bulk_g2s(dst, t.s0, t.b0, &mbar[slot], pol_w);
        if (t.b1) m4_bulk_g2s(dst + t.b0, t.s1, t.b1, &mbar[slot], pol_w);
        if (t.b2) m4_bulk_g2s(dst + t.b0 + t.b1, t.s2, t.b2, &mbar[slot], pol_w);
    };
    constexpr int ISSUER = RW_NT - 32;
    uint32_t wphase = 0;
    if (tid == 0) {
        for (int i = 0; i < P.nslot; i++) m4_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == ISSUER) {
#pragma unroll 1
        for (int n = 0; n < P.nslot - 1; n++) fire_tile(n);
    }

    // LayerNorm + quantise of row `row` into its d_model record (stage A / E): 4 warps, 8 elements per thread
    auto ln_row = [&](int row, const float * xsrc, const float * lnw, const float * lnb) {
        if (tid < M5_PT) {
            float v[8];
            if (xsrc) {
                const float4 a = __ldcg((const float4 *) (xsrc + (size_t) row * M5_D + 8 * tid));
                const float4 b = __ldcg((const float4 *) (xsrc + (size_t) row * M5_D + 8 * tid + 4));
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
                // embedding (biogpt.cpp:663-686): get_rows(embed_tokens) * sqrt(d_model) + get_rows(embed_pos, pos + 2)
                int stream, pos, T; bg_row_info(mode, S, n_past, row, stream, pos, T);
                int tok = P.tokens[row]; tok = tok < 0 ? 0 : (tok >= p.n_vocab ? p.n_vocab - 1 : tok);
                int prow = pos + 2; prow = prow >= p.n_pos_rows ? p.n_pos_rows - 1 : prow;
                const size_t rb = bg_file_row_bytes(FMT, M5_D);
                const uint8_t * tr = p.embed_tok + rb * (size_t) tok;
                const uint8_t * pr = p.embed_pos + rb * (size_t) prow;
#pragma unroll
                for (int i = 0; i < 8; i++)
                    v[i] = __fadd_rn(__fmul_rn(bg_dequant_elem(FMT, tr, 8 * tid + i), p.emb_scale), bg_dequant_elem(FMT, pr, 8 * tid + i));
                *(float4 *) (p.x + (size_t) row * M5_D + 8 * tid) = make_float4(v[0], v[1], v[2], v[3]);
                *(float4 *) (p.x + (size_t) row * M5_D + 8 * tid + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            float y[8];
            m5_layer_norm8<false>(v, lnw, lnb, p.eps, sredA, sredB, y, nullptr);
            m5_quant8<FMT>(y, P.rec_d + (size_t) row * p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off);
        }
    };
    // the cached K/V rows of this CTA's attention unit towards the L2 (layer Ln)
    auto prefetch_kv = [&](int Ln) {
        if (!has_au || Ln >= p.n_layer) return;
        int stream, pos, T; bg_row_info(mode, S, n_past, au_row, stream, pos, T);
        const size_t base = (size_t) Ln * n_pos * M5_D + (size_t) stream * P.stream_stride + (size_t) au_h * SK_DK;
        for (int t = tid; t < n_past; t += RW_NT) {
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.kcache + base + (size_t) t * M5_D));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.kcache + base + (size_t) t * M5_D + 32));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.vcache + base + (size_t) t * M5_D));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p.vcache + base + (size_t) t * M5_D + 32));
        }
    };
    prefetch_kv(0);
    RWSTAMP(p.n_layer, 7);

#pragma unroll 1
    for (int tn = 0; tn < n_tiles; tn++) {
        const bool lm = tn >= n_lt;
        const int kind = lm ? 4 : (tn & 3);                        // 0 q,k,v  1 out_proj  2 fc1  3 fc2  4 lm_head
        const int l = lm ? p.n_layer : (tn >> 2);
        const MegaLayer & L = P.layers[lm ? 0 : l];
        unsigned int * C = P.cnt + (size_t) l * RW_NST;
        const int wcode = ((kind + 1) << 16) | (l << 8);
        float * kc = p.kcache + (size_t) (lm ? 0 : l) * n_pos * M5_D;
        float * vc = p.vcache + (size_t) (lm ? 0 : l) * n_pos * M5_D;

        // ---- stage A / E / final LayerNorm: one row per CTA
        if (kind == 0 || kind == 2 || tn == n_lt) {
            const int first = tn == n_lt ? lm_t0 : 0;
            if (cta >= first && cta < S) {
                if (kind == 2) rw_wait(C + RW_D, RW_NC, err, wcode | 3);
                else if (tn > 0) rw_wait(P.cnt + (size_t) (l - 1) * RW_NST + RW_G, RW_NC, err, wcode | 3);
                if (kind == 2) ln_row(cta, p.x1, L.ln1_w, L.ln1_b);
                else if (tn == n_lt) ln_row(cta, p.x, p.lnf_w, p.lnf_b);
                else ln_row(cta, tn == 0 ? nullptr : p.x, L.ln0_w, L.ln0_b);
                rw_arrive(C + (kind == 2 ? RW_E : RW_A));
                RWSTAMP(l, kind == 2 ? RW_E : RW_A);
            }
            if (kind == 0) prefetch_kv(l + 1);
        }
        if (lm && tn > n_lt) __syncthreads();                      // everyone is done with the previous lm_head tile before its slot is refilled
        if (tid == ISSUER) fire_tile(tn - 1 + P.nslot);            // the slot of tile tn-1 is free: the stage ended with a block barrier

        // ---- geometry of the tile
        int rt;
        if (kind == 0) rt = 24; else if (kind == 2) rt = 32; else if (kind == 4) rt = min(RW_LMRT, v1 - (v0 + (tn - n_lt) * RW_LMRT)); else rt = 8;
        const bool Kff = kind == 3;
        M4MM D;
        D.G = Kff ? p.Gf : p.Gd; D.gsh = Kff ? 5 : 3; D.stride = Kff ? p.stride_f : p.stride_d;
        D.off_qh = Kff ? p.offqh_f : p.offqh_d; D.off_d = Kff ? p.offd_f : p.offd_d; D.off_m = Kff ? p.offm_f : p.offm_d;
        D.off_n = Kff ? p.offn_f : p.offn_d; D.off_dd = Kff ? p.offdd_f : p.offdd_d; D.off_s = Kff ? p.offs_f : p.offs_d;
        const int rb = Kff ? p.actb_f : p.actb_d;
        const int t0 = lm ? lm_t0 : 0, nt = S - t0;                // token rows [t0, S) take part
        const int slot = tn % P.nslot;
        const uint8_t * wt = s_w + (size_t) slot * P.slot_bytes;

        // ---- wait for the producers of this stage's records, and for the weights
        if (!lm || tn == n_lt) {
            unsigned int target; const unsigned int * c;
            if (kind == 0)      { c = C + RW_A; target = (unsigned) S; }
            else if (kind == 1) { c = C + RW_C; target = (unsigned) (M5_NH * S); }
            else if (kind == 2) { c = C + RW_E; target = (unsigned) S; }
            else if (kind == 3) { c = C + RW_F; target = RW_NC; }
            else                { c = C + RW_A; target = (unsigned) nt; }
            rw_wait(c, target, err, wcode | 1);
            const uint4 * gsrc = (const uint4 *) ((Kff ? P.rec_f : P.rec_d) + (size_t) t0 * rb);
            const int nv16 = (nt * rb) >> 4;
            for (int i = tid; i < nv16; i += RW_NT) ((uint4 *) s_rec)[i] = __ldcg(gsrc + i);
        }
        if (tid == 0) m5_mbar_wait(&mbar[slot], (wphase >> slot) & 1u, err, wcode | 2);
        wphase ^= 1u << slot;
        __syncthreads();

        // ---- dot products: warp-unit u = (quad of token rows, weight row)
        {
            const int nq = (nt + 3) >> 2, nunits = rt * nq;
            const int tq = lane >> 3;
#pragma unroll 1
            for (int u = warp; u < nunits; u += RW_NW) {
                const int q = u / rt, rl = u - q * rt;
                const int r = q * 4 + tq, rc = r < nt ? r : nt - 1, tok = t0 + r;
                const bool owner = (lane & 7) == 0 && r < nt;
                float bias = 0.f, resid = 0.f;
                int row = 0;
                if (kind == 0) {
                    const int mat = rl >> 3; row = o0 + (rl & 7);
                    if (owner) bias = (mat == 0 ? L.q_b : mat == 1 ? L.k_b : L.v_b)[row];
                } else if (kind == 2) { row = cta * 32 + rl; if (owner) bias = L.fc1_b[row]; }
                else if (kind == 4)   { row = v0 + (tn - n_lt) * RW_LMRT + rl; }
                else {
                    row = o0 + rl;
                    if (owner) { bias = (kind == 1 ? L.o_b : L.fc2_b)[row]; resid = __ldcg((kind == 1 ? p.x : p.x1) + (size_t) tok * M5_D + row); }
                }
                const uint8_t * wrow = wt + (size_t) rl * D.stride;
                const uint8_t * rec = s_rec + (size_t) rc * rb;
                const float dot = Kff ? rw_row_dot<FMT, 32>(wrow, rec, D) : rw_row_dot<FMT, 8>(wrow, rec, D);
                if (owner) {
                    if (kind == 0) {
                        const int mat = rl >> 3;
                        const float t = __fadd_rn(bias, dot);
                        if (mat == 0) p.q[(size_t) tok * M5_D + row] = __fmul_rn(t, p.qscale);
                        else {
                            int stream, pos, T; bg_row_info(mode, S, n_past, tok, stream, pos, T);
                            (mat == 1 ? kc : vc)[(size_t) stream * P.stream_stride + (size_t) pos * M5_D + row] = t;
                        }
                    } else if (kind == 1) p.x1[(size_t) tok * M5_D + row] = __fadd_rn(__fadd_rn(dot, bias), resid);
                    else if (kind == 2)   s_g[r * 32 + rl] = __fadd_rn(bias, dot);
                    else if (kind == 3)   p.x[(size_t) tok * M5_D + row] = __fadd_rn(__fadd_rn(dot, bias), resid);
                    else                  p.logits[(size_t) r * p.n_vocab + row] = dot;
                }
            }
        }
        if (kind == 2) {
            // fp16-table GELU, then the CTA's 32 values of every token row are block `cta` of the row's d_ff record
            __syncthreads();
            for (int i = tid; i < nt * 32; i += RW_NT) s_g[i] = bg_h2f(p.gelu[bg_f2h(s_g[i])]);
            __syncthreads();
            if (warp < nt) {
                const int l8 = lane & 7;
                const float4 v = *(const float4 *) (s_g + warp * 32 + 4 * l8);
                sk_quant_block<FMT>(v, cta, l8, P.rec_f + (size_t) warp * p.actb_f, p.offn_f, p.offdd_f, p.offs_f, p.code_off, lane < 8);
            }
        }
        if (!lm) {
            rw_arrive(C + (kind == 0 ? RW_B : kind == 1 ? RW_D : kind == 2 ? RW_F : RW_G));
            RWSTAMP(l, kind == 0 ? RW_B : kind == 1 ? RW_D : kind == 2 ? RW_F : RW_G);
        }

        // ---- stage C: softmax(K q) V of one (head, token row); body of k_sk_attn (bgpt_skinny.cuh)
        if (kind == 0 && has_au) {
            rw_wait(C + RW_B, RW_NC, err, wcode | 4);
            const int h = au_h, row = au_row;
            int stream, pos, T; bg_row_info(mode, S, n_past, row, stream, pos, T);
            const float * Kb = kc + (size_t) stream * P.stream_stride + (size_t) h * SK_DK;
            const float * Vb = vc + (size_t) stream * P.stream_stride + (size_t) h * SK_DK;
            const float q0 = __ldcg(p.q + (size_t) row * M5_D + h * SK_DK + lane), q1 = __ldcg(p.q + (size_t) row * M5_D + h * SK_DK + 32 + lane);
#pragma unroll 1
            for (int tb = warp; tb < T; tb += RW_NW * 16) {
                float kr[16][2];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int t = tb + u * RW_NW;
                    kr[u][0] = (t < T) ? __ldcg(Kb + (size_t) t * M5_D + lane) : 0.0f;
                    kr[u][1] = (t < T) ? __ldcg(Kb + (size_t) t * M5_D + 32 + lane) : 0.0f;
                }
                float s[16];
#pragma unroll
                for (int u = 0; u < 16; u++) { float x = 0.0f; x = fmaf(kr[u][0], q0, x); x = fmaf(kr[u][1], q1, x); s[u] = x; }
                // 16 reduce trees (xor 16, 8, 4, 1, 2: GGML_F32x8_REDUCE) as one transposing butterfly
                float a8[8], a4[4], a2[2];
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b0 = lane & 1;
#pragma unroll
                for (int i = 0; i < 8; i++) { const float mine = b4 ? s[8 + i] : s[i], send = b4 ? s[i] : s[8 + i]; a8[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 16)); }
#pragma unroll
                for (int i = 0; i < 4; i++) { const float mine = b3 ? a8[4 + i] : a8[i], send = b3 ? a8[i] : a8[4 + i]; a4[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 8)); }
#pragma unroll
                for (int i = 0; i < 2; i++) { const float mine = b2 ? a4[2 + i] : a4[i], send = b2 ? a4[i] : a4[2 + i]; a2[i] = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 4)); }
                const float mine = b0 ? a2[1] : a2[0], send = b0 ? a2[0] : a2[1];
                const float a1 = __fadd_rn(mine, __shfl_xor_sync(FULLMASK, send, 1));
                const float dotv = __fadd_rn(a1, __shfl_xor_sync(FULLMASK, a1, 2));
                const int u = ((lane >> 1) & 14) | (lane & 1);
                const int t = tb + u * RW_NW;
                if (t < T && !(lane & 2)) sc[t] = dotv;
            }
            const int np = T & ~31;
            for (int i = tid; i < (T - np) * SK_DK; i += RW_NT) tailv[i] = __ldcg(Vb + (size_t) (np + i / SK_DK) * M5_D + (i % SK_DK));
            __syncthreads();
            {   // softmax over sc[0..T): max, fp16-table exp, sum in double (exact for fp16 values), scale (ggml.c:12955-12974)
                const float x0 = tid < T ? sc[tid] : -INFINITY;
                float mx = x0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
                if (lane == 0) sredF[warp] = mx;
                __syncthreads();
                mx = sredF[lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULLMASK, mx, o));
                float e0 = 0.f;
                if (tid < T) e0 = bg_h2f(p.exp_tab[bg_f2h(__fsub_rn(x0, mx))]);
                double sm = (double) e0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(FULLMASK, sm, o);
                if (lane == 0) sredA[warp] = sm;
                __syncthreads();
                double tot = sredA[lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULLMASK, tot, o);
                const float inv = (float) (1.0 / tot);
                if (tid < T) sc[tid] = __fmul_rn(e0, inv);
            }
            __syncthreads();
            {   // V: thread (r = tid / 32, 2 columns): running sum r over t = r, r + 32, ... < np (ggml_vec_dot_f32's 32 lanes)
                const int vr = tid >> 5, vcn = tid & 31;
                const float * vp = Vb + (size_t) vr * M5_D + 2 * vcn;
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
                for (int s0 = 0; s0 < np; s0 += 512) {
                    float2 vv[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int t = s0 + 32 * k + vr;
                        vv[k] = (t < np) ? __ldcg((const float2 *) (vp + (size_t) (s0 + 32 * k) * M5_D)) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int t = s0 + 32 * k + vr;
                        if (t < np) {
                            const float pw = sc[t];
                            acc.x = fmaf(vv[k].x, pw, acc.x); acc.y = fmaf(vv[k].y, pw, acc.y);
                        }
                    }
                }
                *(float2 *) (red + vr * SK_DK + 2 * vcn) = acc;
            }
            __syncthreads();
            if (tid < SK_DK) {
                float x0[8];
#pragma unroll
                for (int l8 = 0; l8 < 8; l8++) {
                    const float a02 = __fadd_rn(red[(0 * 8 + l8) * SK_DK + tid], red[(2 * 8 + l8) * SK_DK + tid]);
                    const float a13 = __fadd_rn(red[(1 * 8 + l8) * SK_DK + tid], red[(3 * 8 + l8) * SK_DK + tid]);
                    x0[l8] = __fadd_rn(a02, a13);
                }
                const float t0s = __fadd_rn(x0[0], x0[4]), t1s = __fadd_rn(x0[1], x0[5]);
                const float t2s = __fadd_rn(x0[2], x0[6]), t3s = __fadd_rn(x0[3], x0[7]);
                float sumf = __fadd_rn(__fadd_rn(t0s, t1s), __fadd_rn(t2s, t3s));
                // the as-built scalar tail of ggml_vec_dot_f32: products unfused in groups of 4, then <= 3 fused
                const int nv = np + ((T - np) & ~3);
                int t = np;
#pragma unroll 1
                for (; t < nv; t++) sumf = __fadd_rn(sumf, __fmul_rn(tailv[(t - np) * SK_DK + tid], sc[t]));
#pragma unroll 1
                for (; t < T;  t++) sumf = fmaf(tailv[(t - np) * SK_DK + tid], sc[t], sumf);
                s_out[tid] = sumf;
            }
            __syncthreads();
            if (warp == 0) {                                         // lanes 0..15: block lane / 8, word lane % 8
                const int blk = (lane >> 3) & 1, l8 = lane & 7;
                const float4 v = *(const float4 *) (s_out + blk * 32 + 4 * l8);
                sk_quant_block<FMT>(v, h * 2 + blk, l8, P.rec_d + (size_t) row * p.actb_d, p.offn_d, p.offdd_d, p.offs_d, p.code_off, lane < 16);
            }
            rw_arrive(C + RW_C);
            RWSTAMP(l, RW_C);
        }
    }
    RWSTAMP(p.n_layer, RW_B);
}
