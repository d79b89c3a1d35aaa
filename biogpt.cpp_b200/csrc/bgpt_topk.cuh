// bgpt_topk.cuh -- the K largest logits on the device, for biogpt_sample_top_k_top_p (biogpt.cpp:908-980).
//
// The reference sorts all n_vocab (logit / temp, id) pairs with std::partial_sort, keeps top_k (default 40), and draws from their
// softmax with std::discrete_distribution on std::mt19937.  Only the top_k pairs influence the draw, so the device returns exactly
// those -- K (logit, id) pairs sorted by logit descending, 8 K bytes instead of 4 n_vocab -- and the host runs the reference's
// double-precision softmax / top-p / RNG draw on them unchanged (host/biogpt_b200.cpp).  partial_sort leaves the order of EQUAL
// values unspecified: when the K-th value is tied with the (K+1)-th, or two of the K values are equal, the kernel says so
// (info[1] = 0) and the caller falls back to the full-logit path, so the drawn id is the reference's in every case.
//
// One CTA: the keys (order-preserving uint32 images of the floats) are staged in shared memory once (170 KB for 42384 logits),
// the K-th largest key is found by a 4-pass radix select on 8-bit digits, the <= K survivors are ranked by counting.  Logits
// share their leading digits, so the histogram of a pass is one or two hot bins: every warp counts into its OWN histogram and
// aggregates equal digits with match.any first (one shared-memory atomic per distinct digit per warp instead of 32 colliding ones;
// the naive version spent 50 us in those collisions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TOPK_NT 1024
#define TOPK_MAXK 128

__device__ __forceinline__ uint32_t topk_key(float f) {            // ascending float order -> ascending uint32 order
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Large vocabularies go through three short launches instead of one CTA scanning every logit (the 4-pass radix select below measures
// ~50 us even on 7 k keys: zeroing 32 per-warp histograms, match.any, a serial scan of 256 bins per pass): k_topk_part lets every
// slice of TOPK_SLICE logits rank itself by counting (166 CTAs for 42384 logits) and keep its K + 1 best, k_topk_rank<false> merges
// TOPK_FAN slices' candidates per CTA the same way, k_topk_rank<true> ranks the remaining few hundred and writes the result packet.
// The global K + 1 largest values all lie in the union of the slices' K + 1 largest at every level, so the selection, its order
// and the tie flags are the ones of the single-kernel form.
// rank of entry e (key u) among the n4 (multiple of 4) keys s_k[]: how many precede it in (key descending, position ascending)
// order = #{ j < e : key_j >= u } + #{ j > e : key_j > u } -- one compare and one add per key, no tie-break inside the loops
__device__ __forceinline__ int topk_rank_of(const uint32_t * s_k, int n4, int e, uint32_t u) {
    int rank = 0;
    const int e4 = e & ~3;
#pragma unroll 4
    for (int j = 0; j < e4; j += 4) {
        const uint4 o = *(const uint4 *) (s_k + j);
        rank += (o.x >= u) + (o.y >= u) + (o.z >= u) + (o.w >= u);
    }
    {
        const uint4 o = *(const uint4 *) (s_k + e4);
        const int k = e - e4;
        rank += (k > 0 ? (o.x >= u) : 0) + (k > 1 ? (o.y >= u) : (k < 1 ? (o.y > u) : 0)) + (k > 2 ? (o.z >= u) : (k < 2 ? (o.z > u) : 0)) + (k < 3 ? (o.w > u) : 0);
    }
#pragma unroll 4
    for (int j = e4 + 4; j < n4; j += 4) {
        const uint4 o = *(const uint4 *) (s_k + j);
        rank += (o.x > u) + (o.y > u) + (o.z > u) + (o.w > u);
    }
    return rank;
}

#define TOPK_SLICE 256
// grid = ceil(n / TOPK_SLICE), block = TOPK_SLICE: thread i owns logit i of the slice; its rank = number of slice entries before it
// in (value descending, index ascending) order; ranks below kc are written to cand_*[slice][rank] (pads: -inf, index -1)
static __global__ void __launch_bounds__(TOPK_SLICE) k_topk_part(const float * __restrict__ logits, int n, int kc,
                                                                  float * __restrict__ cand_val, int * __restrict__ cand_idx) {
    __shared__ __align__(16) uint32_t s_k[TOPK_SLICE];
    const int tid = threadIdx.x, i = blockIdx.x * TOPK_SLICE + tid;
    logits += (size_t) blockIdx.y * n;                             // blockIdx.y: the logit row (lock-step streams), candidates per row
    cand_val += (size_t) blockIdx.y * gridDim.x * kc; cand_idx += (size_t) blockIdx.y * gridDim.x * kc;
    const float f = i < n ? logits[i] : -INFINITY;
    const uint32_t u = i < n ? topk_key(f) : 0u;
    s_k[tid] = u;
    __syncthreads();
    const int rank = topk_rank_of(s_k, TOPK_SLICE, tid, u);
    if (rank < kc) {
        cand_val[(size_t) blockIdx.x * kc + rank] = f;
        cand_idx[(size_t) blockIdx.x * kc + rank] = i < n ? i : -1;
    }
}

// second and third launch of the large-vocabulary form: rank-by-counting over a slice of `per` candidates (value, index) held in shared
// memory, keep the kc best.  !FINAL: grid = slices, the survivors go to out_*[slice][rank] (pads: -inf, -1).  FINAL: one CTA over all
// remaining candidates: the k best in order, info = { entries, exact, error code of the forward pass, sequence number } -- `exact`
// is 0 when two neighbours among the k + 1 largest values are equal (a duplicate inside the top k, or the k-th tied with the
// (k+1)-th: std::partial_sort's choice or order would be ambiguous).  The outputs may live in mapped pinned HOST memory: the sequence
// number is written last, behind a system-scope fence, so a host thread that polls it sees a complete packet.
#define TOPK_RANK_NT 512
#define TOPK_FAN 16                     // slices merged per second-level CTA
#define TOPK_RANK_MAX (TOPK_FAN * (TOPK_MAXK + 1))
template <bool FINAL>
static __global__ void __launch_bounds__(TOPK_RANK_NT) k_topk_rank(const float * __restrict__ val, const int * __restrict__ idx, int n_c, int per, int kc,
                                                                    float * __restrict__ out_val, int * __restrict__ out_idx,
                                                                    int k, int * __restrict__ info, const int * __restrict__ err, unsigned seq,
                                                                    int * __restrict__ n_c_dev, int row_stride_bytes = 0) {
    __shared__ __align__(16) uint32_t s_k[TOPK_RANK_MAX + 4];
    __shared__ uint32_t s_sorted[TOPK_MAXK + 2];
    const int tid = threadIdx.x;
    // blockIdx.y: the logit row (lock-step streams): n_c candidates per row in, gridDim.x * kc per row out (FINAL: one packet per row,
    // row_stride_bytes apart, each with its own serial)
    val += (size_t) blockIdx.y * n_c; idx += (size_t) blockIdx.y * n_c;
    if (FINAL) {
        out_val = (float *) ((uint8_t *) out_val + (size_t) blockIdx.y * row_stride_bytes);
        out_idx = (int *) ((uint8_t *) out_idx + (size_t) blockIdx.y * row_stride_bytes);
        info = (int *) ((uint8_t *) info + (size_t) blockIdx.y * row_stride_bytes);
    } else { out_val += (size_t) blockIdx.y * gridDim.x * kc; out_idx += (size_t) blockIdx.y * gridDim.x * kc; }
    const int b0 = blockIdx.x * per;
    // FINAL after k_topk_filter: the number of candidates is the filter's counter; more than fit = a plateau of equal logits:
    // the packet says "not exact" and the caller takes the full-row path
    bool overflow = false;
    if (FINAL && n_c_dev) { n_c = *(volatile int *) n_c_dev; per = n_c; overflow = n_c > TOPK_RANK_MAX; }
    int cnt = n_c - b0; cnt = cnt < per ? cnt : per; cnt = cnt < TOPK_RANK_MAX ? cnt : TOPK_RANK_MAX;
    const int cnt4 = (cnt + 3) & ~3;
    for (int i = tid; i < cnt4; i += TOPK_RANK_NT) s_k[i] = i < cnt ? topk_key(val[b0 + i]) : 0u;
    if (FINAL) for (int i = tid; i < TOPK_MAXK + 2; i += TOPK_RANK_NT) s_sorted[i] = 0u;
    __syncthreads();
    for (int e = tid; e < cnt; e += TOPK_RANK_NT) {
        const uint32_t u = s_k[e];
        const int rank = topk_rank_of(s_k, cnt4, e, u);
        if (rank < kc) {
            if (FINAL) {
                s_sorted[rank] = u;
                if (rank < k) { out_val[rank] = val[b0 + e]; out_idx[rank] = idx[b0 + e]; }
            } else {
                out_val[(size_t) blockIdx.x * kc + rank] = val[b0 + e];
                out_idx[(size_t) blockIdx.x * kc + rank] = idx[b0 + e];
            }
        }
    }
    if (!FINAL) {
        for (int r = cnt + tid; r < kc; r += TOPK_RANK_NT) { out_val[(size_t) blockIdx.x * kc + r] = -INFINITY; out_idx[(size_t) blockIdx.x * kc + r] = -1; }
        return;
    }
    __threadfence_system();
    __syncthreads();
    const int k_eff = k < cnt ? k : cnt;
    int dup = 0;
    for (int j = tid; j < k_eff && j + 1 < cnt; j += TOPK_RANK_NT) dup |= (s_sorted[j] == s_sorted[j + 1]) ? 1 : 0;
    dup = __syncthreads_or(dup);
    if (tid == 0) {
        if (n_c_dev) *n_c_dev = 0;                                   // ready for the next call's filter
        info[0] = k_eff; info[1] = (dup || overflow) ? 0 : 1; info[2] = err ? *(volatile const int *) err : 0;
        __threadfence_system();
        *(volatile unsigned *) (info + 3) = seq;
    }
}

// Single-token steps on a persistent decode kernel leave one maximum per CTA (the argmax candidates of its lm_head rows).  The k-th
// largest of those n_max maxima, t0, is a lower bound of the k-th largest logit (k logits -- the maxima themselves -- are >= t0),
// so the k + 1 largest logits that matter (the (k+1)-th only if it ties or beats t0) are all among { logit >= t0 }: typically a few
// hundred entries.  grid = ceil(n / 256): every CTA derives t0 itself (rank-counting n_max <= 256 values), then appends its
// logits >= t0 to the candidate list; k_topk_rank<true> ranks the list, writes the packet and resets the counter.
static __global__ void __launch_bounds__(TOPK_SLICE) k_topk_filter(const float * __restrict__ logits, int n, int k,
                                                                    const float * __restrict__ slice_max, int n_max,
                                                                    float * __restrict__ out_val, int * __restrict__ out_idx,
                                                                    int * __restrict__ counter, int cap) {
    __shared__ __align__(16) uint32_t s_m[TOPK_SLICE];
    __shared__ uint32_t s_t0;
    const int tid = threadIdx.x, i = blockIdx.x * TOPK_SLICE + tid;
    const float f = i < n ? logits[i] : 0.0f;
    const uint32_t mk = tid < n_max ? topk_key(slice_max[tid]) : 0u;
    s_m[tid] = mk;
    __syncthreads();
    if (tid < n_max && topk_rank_of(s_m, (n_max + 3) & ~3, tid, mk) == k - 1) s_t0 = mk;     // ranks are unique: exactly one writer
    __syncthreads();
    if (i < n && topk_key(f) >= s_t0) {
        const int p = atomicAdd(counter, 1);
        if (p < cap) { out_val[p] = f; out_idx[p] = i; }
    }
}

// The same selection as k_topk_filter + k_topk_rank<true>, as the TAIL of a persistent decode kernel: the CTA that finishes its
// lm_head rows last (an atomic ticket) derives t0 from the n_max per-CTA maxima, scans the logit row in L2 for { logit >= t0 },
// ranks the survivors by counting and writes the result packet -- no extra launch, no extra grid-wide exchange.  Called by every
// thread of that CTA (NT threads); scratch: TOPK_TAIL_SMEM bytes of shared memory, 16-byte aligned; n_max <= 256.
#define TOPK_TAIL_SMEM (2 * (TOPK_RANK_MAX + 4) * 4 + (2 * (TOPK_MAXK + 4) + 256) * 4)
template <int NT>
__device__ __noinline__ void topk_tail(const float * logits, int n, int k, const float * cta_max, int n_max, uint8_t * scratch,
                                          float * out_val, int * out_idx, int * info, const int * err, unsigned seq) {
    __shared__ int s_cnt; __shared__ uint32_t s_t0;
    uint32_t * s_k = (uint32_t *) scratch;
    int * s_i = (int *) (s_k + TOPK_RANK_MAX + 4);
    uint32_t * s_sorted = (uint32_t *) (s_i + TOPK_RANK_MAX + 4);
    uint32_t * s_m = s_sorted + TOPK_MAXK + 4;
    const int tid = threadIdx.x;
    const int n_max4 = (n_max + 3) & ~3;
    // the row is read in rounds of TB 16-byte loads per thread, all in flight together (a load per loop iteration would cost one L2
    // round trip each: 21 x 0.7 us); the first round is issued before the threshold is known
    constexpr int TB = 11;
    const int n4 = n >> 2;
    float4 v[TB];
#pragma unroll
    for (int j = 0; j < TB; j++) { const int i = j * NT + tid; if (i < n4) v[j] = __ldcg((const float4 *) logits + i); }
    for (int i = tid; i < n_max4; i += NT) s_m[i] = i < n_max ? topk_key(__ldcg(cta_max + i)) : 0u;
    for (int i = tid; i < TOPK_MAXK + 2; i += NT) s_sorted[i] = 0u;
    if (tid == 0) { s_cnt = 0; s_t0 = 0u; }
    __syncthreads();
    for (int i = tid; i < n_max; i += NT) { const uint32_t mk = s_m[i]; if (topk_rank_of(s_m, n_max4, i, mk) == k - 1) s_t0 = mk; }   // ranks are unique
    __syncthreads();
    const uint32_t t0 = s_t0;
#pragma unroll 1
    for (int base = 0; base < n4; base += TB * NT) {
        if (base > 0) {
#pragma unroll
            for (int j = 0; j < TB; j++) { const int i = base + j * NT + tid; if (i < n4) v[j] = __ldcg((const float4 *) logits + i); }
        }
#pragma unroll
        for (int j = 0; j < TB; j++) {
            const int i = base + j * NT + tid;
            if (i >= n4) break;
            const float f[4] = { v[j].x, v[j].y, v[j].z, v[j].w };
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t u = topk_key(f[c]);
                if (u >= t0) { const int p = atomicAdd(&s_cnt, 1); if (p < TOPK_RANK_MAX) { s_k[p] = u; s_i[p] = 4 * i + c; } }
            }
        }
    }
    for (int i = 4 * n4 + tid; i < n; i += NT) {
        const uint32_t u = topk_key(__ldcg(logits + i));
        if (u >= t0) { const int p = atomicAdd(&s_cnt, 1); if (p < TOPK_RANK_MAX) { s_k[p] = u; s_i[p] = i; } }
    }
    __syncthreads();
    const bool overflow = s_cnt > TOPK_RANK_MAX;                   // a plateau of equal logits: "not exact", the caller takes the full row
    const int cnt = overflow ? TOPK_RANK_MAX : s_cnt, cnt4 = (cnt + 3) & ~3;
    if (tid < cnt4 - cnt) s_k[cnt + tid] = 0u;
    __syncthreads();
    // the winners are staged in shared memory (values in s_sorted, ids in s_oi) and leave as consecutive words from ONE warp: the packet
    // may live in mapped pinned host memory, where 2 k scattered 4-byte stores from 40 warps are 2 k PCIe writes
    int * s_oi = (int *) (s_m + 256);
    for (int e = tid; e < cnt; e += NT) {
        const uint32_t u = s_k[e];
        const int rank = topk_rank_of(s_k, cnt4, e, u);
        if (rank < k + 1) { s_sorted[rank] = u; if (rank < k) s_oi[rank] = s_i[e]; }
    }
    __syncthreads();
    const int k_eff = k < cnt ? k : cnt;
    int dup = 0;
    for (int j = tid; j < k_eff && j + 1 < cnt; j += NT) dup |= (s_sorted[j] == s_sorted[j + 1]) ? 1 : 0;
    dup = __syncthreads_or(dup);
    if (tid < 32) {
        for (int j = tid; j < k_eff; j += 32) {
            const uint32_t u = s_sorted[j];
            out_val[j] = __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
            out_idx[j] = s_oi[j];
        }
        if (tid == 0) { info[0] = k_eff; info[1] = (dup || overflow) ? 0 : 1; info[2] = err ? *(volatile const int *) err : 0; }
        __syncwarp();
        if (tid == 0) { __threadfence_system(); *(volatile unsigned *) (info + 3) = seq; }
    }
}

// out_val / out_idx: K entries; info[0] = entries written (min(K, n)), info[1] = 1 if the selection and its order are unambiguous
// (info: 4 ints, see the end of the kernel).
// remap (optional): out_idx[rank] = remap[position] -- the input is a candidate list, not the logit row.
static __global__ void __launch_bounds__(TOPK_NT) k_topk(const float * __restrict__ logits, int n, int k, int staged,
                                                          float * __restrict__ out_val, int * __restrict__ out_idx, int * __restrict__ info,
                                                          const int * __restrict__ remap, const int * __restrict__ err, unsigned seq) {
    extern __shared__ uint32_t s_keys[];                            // n keys when `staged`
    __shared__ unsigned hist[256];
    __shared__ unsigned whist[TOPK_NT / 32][256];                  // per-warp histograms
    __shared__ uint32_t s_prefix, s_remaining;
    __shared__ unsigned s_cnt, s_eq;
    __shared__ uint32_t c_key[TOPK_MAXK]; __shared__ int c_idx[TOPK_MAXK];
    __shared__ int s_dup;
    const int tid = threadIdx.x;
    if (k > n) k = n;
    if (k > TOPK_MAXK) k = TOPK_MAXK;
    if (staged) for (int i = tid; i < n; i += TOPK_NT) s_keys[i] = topk_key(logits[i]);
    if (tid == 0) { s_prefix = 0u; s_remaining = (uint32_t) k; s_cnt = 0u; s_eq = 0u; s_dup = 0; }
    __syncthreads();
    uint32_t mask = 0u;
    for (int pass = 3; pass >= 0; pass--) {
        for (int b = tid; b < (TOPK_NT / 32) * 256; b += TOPK_NT) (&whist[0][0])[b] = 0u;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        unsigned * mine = whist[tid >> 5];
        const int n_pad = (n + 31) & ~31;                          // whole warps take part in match.any
        for (int i = tid; i < n_pad; i += TOPK_NT) {
            const uint32_t u = i < n ? (staged ? s_keys[i] : topk_key(logits[i])) : 0u;
            const bool live = i < n && (u & mask) == prefix;
            const unsigned digit = live ? ((u >> (8 * pass)) & 255u) : 256u + (tid & 31);     // dead lanes match nobody
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (live && (int) (__ffs(peers) - 1) == (tid & 31)) atomicAdd(&mine[digit], (unsigned) __popc(peers));
        }
        __syncthreads();
        for (int b = tid; b < 256; b += TOPK_NT) {
            unsigned c = 0;
#pragma unroll 8
            for (int w = 0; w < TOPK_NT / 32; w++) c += whist[w][b];
            hist[b] = c;
        }
        __syncthreads();
        if (tid == 0) {                                             // the digit of the remaining-th largest key among the survivors
            uint32_t need = s_remaining, c = 0u; int bin = 0;
            for (int b = 255; b >= 0; b--) { if (c + hist[b] >= need) { bin = b; break; } c += hist[b]; }
            s_remaining = need - c;
            s_prefix = prefix | ((uint32_t) bin << (8 * pass));
        }
        mask |= 0xFFu << (8 * pass);
        __syncthreads();
    }
    const uint32_t thr = s_prefix;                                  // the K-th largest key; s_remaining of the keys equal to it are needed
    const uint32_t need_eq = s_remaining;
    for (int i = tid; i < n; i += TOPK_NT) {
        const uint32_t u = staged ? s_keys[i] : topk_key(logits[i]);
        if (u > thr) { const unsigned p = atomicAdd(&s_cnt, 1u); if (p < TOPK_MAXK) { c_key[p] = u; c_idx[p] = i; } }
        else if (u == thr) atomicAdd(&s_eq, 1u);
    }
    __syncthreads();
    const unsigned n_gt = s_cnt;                                    // = k - need_eq
    // the keys equal to the threshold: lowest indices first (any choice is flagged when there are more than needed)
    if (tid == 0) s_cnt = n_gt;
    __syncthreads();
    if (s_eq <= need_eq) {
        for (int i = tid; i < n; i += TOPK_NT) {
            const uint32_t u = staged ? s_keys[i] : topk_key(logits[i]);
            if (u == thr) { const unsigned p = atomicAdd(&s_cnt, 1u); if (p < TOPK_MAXK) { c_key[p] = u; c_idx[p] = i; } }
        }
    } else if (tid == 0) {
        unsigned got = 0;
        for (int i = 0; i < n && got < need_eq; i++) {
            const uint32_t u = staged ? s_keys[i] : topk_key(logits[i]);
            if (u == thr) { c_key[n_gt + got] = u; c_idx[n_gt + got] = i; got++; }
        }
        s_dup = 1;                                                  // boundary tie: which of the equal values partial_sort keeps is unspecified
    }
    __syncthreads();
    // rank by counting: key descending, index ascending among equals (equal keys are flagged)
    if (tid < k) {
        const uint32_t mk = c_key[tid]; const int mi = c_idx[tid];
        int rank = 0;
        for (int j = 0; j < k; j++) {
            const uint32_t ok = c_key[j];
            if (ok > mk || (ok == mk && c_idx[j] < mi)) rank++;
            if (j != tid && ok == mk) s_dup = 1;
        }
        const uint32_t u = mk;
        const uint32_t bits = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
        out_val[rank] = __uint_as_float(bits);
        out_idx[rank] = remap ? remap[mi] : mi;
    }
    // info = { entries, exact, error code of the forward pass (the persistent kernel's watchdog flag), sequence number }.  The outputs
    // may live in mapped pinned HOST memory: the sequence number is written last, behind a system-scope fence, so a host thread that
    // polls it sees a complete packet without a stream synchronisation.
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        info[0] = k; info[1] = s_dup ? 0 : 1; info[2] = err ? *(volatile const int *) err : 0;
        __threadfence_system();
        *(volatile unsigned *) (info + 3) = seq;
    }
}
