// K4: diagonal sequence matching + top-N selection + Recall@N counters.
//
// Replaces lens/run_model.py:248-254 (conv2d(S, eye(L)) / L, transposed) and
// lens/src/metrics.py:213-224 (argsort(0)[-K:] + hit test), called from
// lens/run_model.py:301-302 for N in {1,5,10,15,20,25}.  HBM-bound: every entry of
// the similarity matrix S[B][Q][P] is needed L times, the L reads of a row hit L2
// (one stream's S is Q*P*4 bytes), DRAM sees it once.
//
// D[b][r][q] = (sum_{j<L} S[b][q+j][r+j]) / L : the partial sums are small integers
// (spike counts), exact in fp32 in any order; the one division is IEEE (__fdiv_rn),
// like numpy's float32 / int.  Selection order is (value desc, index desc) -- the
// order np.argsort(kind='stable')[-N:][::-1] produces -- implemented as a 64-bit max
// over keys (orderable(value) << 32 | index).
#include "common.cuh"

#include <algorithm>
#include <climits>

namespace lens {

constexpr int kMatchThreads = 256;
constexpr int kCandChunk = 2048;
constexpr int kMaxTopN = 64;
constexpr int kUnrollR = 4;       // candidate batches (of 32 places) in flight per warp in the warp-per-query kernel

__device__ __forceinline__ uint32_t f32_orderable(float v)
{
    uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// grid = B * Qo.  One CTA scores every database place for one (stream, query).
__global__ void __launch_bounds__(kMatchThreads)
seqmatch_topk_kernel(const float *__restrict__ S, int Q, int P, int L, int N,
                     float *__restrict__ D_out, float *__restrict__ top_val,
                     int32_t *__restrict__ top_idx)
{
    __shared__ unsigned long long cand[kCandChunk];      // keys of the current chunk of places
    __shared__ unsigned long long best[kMaxTopN];        // running top-N (descending)
    __shared__ unsigned long long next_best[kMaxTopN];
    __shared__ unsigned long long warp_best[kMatchThreads / 32];
    __shared__ unsigned long long s_winner;
    const int Qo = Q - L + 1, Po = P - L + 1;
    const int b = blockIdx.x / Qo, q = blockIdx.x - b * Qo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *Sb = S + (size_t)b * Q * P;
    const float fl = (float)L;
    for (int i = tid; i < kMaxTopN; i += kMatchThreads) best[i] = 0ull;   // 0 = empty slot

    for (int r0 = 0; r0 < Po; r0 += kCandChunk) {
        const int nr = min(kCandChunk, Po - r0);
        for (int i = tid; i < nr; i += kMatchThreads) {
            const int r = r0 + i;
            float acc = 0.0f;
            for (int j = 0; j < L; ++j) acc += __ldg(Sb + (size_t)(q + j) * P + (r + j));
            const float d = __fdiv_rn(acc, fl);
            if (D_out) D_out[((size_t)b * Po + r) * Qo + q] = d;
            cand[i] = ((unsigned long long)f32_orderable(d) << 32) | (uint32_t)r;
        }
        __syncthreads();
        // N rounds of block-wide max over (running best U chunk); keys are unique
        // (distinct place index) so exactly one slot holds the winner.
        for (int n = 0; n < N; ++n) {
            unsigned long long mine = 0ull;
            int mine_pos = -1;
            for (int i = tid; i < N + nr; i += kMatchThreads) {
                const unsigned long long k = (i < N) ? best[i] : cand[i - N];
                if (k > mine) { mine = k; mine_pos = i; }
            }
            unsigned long long wb = mine;
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, wb, o);
                wb = other > wb ? other : wb;
            }
            if (lane == 0) warp_best[warp] = wb;
            __syncthreads();
            if (tid == 0) {
                unsigned long long w = 0ull;
                for (int i = 0; i < kMatchThreads / 32; ++i) w = warp_best[i] > w ? warp_best[i] : w;
                s_winner = w;
                next_best[n] = w;
            }
            __syncthreads();
            if (mine != 0ull && mine == s_winner) {
                if (mine_pos < N) best[mine_pos] = 0ull;
                else cand[mine_pos - N] = 0ull;
            }
            __syncthreads();
        }
        for (int i = tid; i < N; i += kMatchThreads) best[i] = next_best[i];
        __syncthreads();
    }
    for (int n = tid; n < N; n += kMatchThreads) {
        const unsigned long long k = best[n];
        const size_t o = ((size_t)b * Qo + q) * N + n;
        if (k == 0ull) {
            top_val[o] = -INFINITY;
            top_idx[o] = -1;
        } else {
            top_val[o] = f32_from_orderable((uint32_t)(k >> 32));
            top_idx[o] = (int32_t)(uint32_t)(k & 0xffffffffull);
        }
    }
}

// ---- warp-per-query variant (N <= 32) ------------------------------------------------------------
// One warp scores one (stream, query): lane i of the warp holds the i-th largest key seen so far.
// Candidates arrive 32 at a time (coalesced loads of the L diagonal terms); a batch is merged only if
// one of its keys beats the current 32nd largest (warp vote), by a 32-element bitonic sort of the batch
// and a bitonic merge with the running list -- shuffles only, no shared memory, no block barriers.
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m)
{
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src)
{
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}

// kSeg = false is the plain one-warp-per-query kernel (40 registers: six CTAs per SM; the segmented instantiation
// needs 48 and would cost the large-batch case a sixth of its resident warps, measured 4.5 -> 7.8 ms on config 3)
template <bool kSeg>
__global__ void __launch_bounds__(256)
seqmatch_topk_warp_kernel(const float *__restrict__ S, long long n_queries, int Q, int P, int L, int N, int n_seg,
                          int seg_len, float *__restrict__ D_out, float *__restrict__ top_val,
                          int32_t *__restrict__ top_idx)
{
    // With n_seg > 1 (few queries, many places) a warp ranks only the places [seg * seg_len, (seg + 1) * seg_len) of
    // its query and writes list `seg` of [n_seg][n_queries][N]; lens_seqmatch_topk merges the lists afterwards.
    const int lane = threadIdx.x & 31;
    const int Qo = Q - L + 1, Po_all = P - L + 1;
    const float fl = (float)L;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = warp0; w < (kSeg ? n_queries * n_seg : n_queries); w += n_warps) {
        const int seg = kSeg ? (int)(w / n_queries) : 0;
        const long long g = kSeg ? w - (long long)seg * n_queries : w;
        const int b = (int)(g / Qo), q = (int)(g - (long long)b * Qo);
        const float *Sb = S + ((size_t)b * Q + q) * P;
        const int Po = kSeg ? min(Po_all, (seg + 1) * seg_len) : Po_all;
        unsigned long long best = 0ull;                    // lane i: i-th largest key, 0 = empty
        // kUnrollR batches of 32 candidates per iteration: all their L x kUnrollR loads are issued before
        // the first use, which is what keeps enough bytes in flight per warp to approach the HBM rate
        for (int r0 = kSeg ? seg * seg_len : 0; r0 < Po; r0 += 32 * kUnrollR) {
            float acc[kUnrollR];
#pragma unroll
            for (int u = 0; u < kUnrollR; ++u) acc[u] = 0.0f;
            if (r0 + 32 * kUnrollR <= Po) {
                for (int j = 0; j < L; ++j) {
                    const float *row = Sb + (size_t)j * P + (r0 + lane + j);
#pragma unroll
                    for (int u = 0; u < kUnrollR; ++u) acc[u] += __ldg(row + 32 * u);
                }
            } else {
                for (int j = 0; j < L; ++j) {
                    const float *row = Sb + (size_t)j * P + (r0 + lane + j);
#pragma unroll
                    for (int u = 0; u < kUnrollR; ++u)
                        if (r0 + 32 * u + lane < Po) acc[u] += __ldg(row + 32 * u);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnrollR; ++u) {
                const int r = r0 + 32 * u + lane;
                unsigned long long key = 0ull;
                if (r < Po) {
                    const float d = __fdiv_rn(acc[u], fl);
                    if (D_out) D_out[((size_t)b * Po_all + r) * Qo + q] = d;
                    key = ((unsigned long long)f32_orderable(d) << 32) | (uint32_t)r;
                }
                const unsigned long long kth = shfl_u64(best, 31);          // current 32nd largest
                if (!__any_sync(0xffffffffu, key > kth)) continue;          // nothing in this batch can enter
                // bitonic sort of the batch, descending
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const unsigned long long other = shfl_xor_u64(key, j);
                        const bool take_max = (((lane & j) == 0) == ((lane & k) == 0));
                        key = take_max ? (key > other ? key : other) : (key < other ? key : other);
                    }
                }
                // top 32 of (running list U batch): elementwise max against the reversed batch is bitonic
                const unsigned long long rev = shfl_u64(key, 31 - lane);
                unsigned long long c = best > rev ? best : rev;
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    const unsigned long long other = shfl_xor_u64(c, j);
                    c = ((lane & j) == 0) ? (c > other ? c : other) : (c < other ? c : other);
                }
                best = c;
            }
        }
        if (lane < N) {
            const size_t o = (size_t)w * N + lane;             // w = seg * n_queries + g
            if (best == 0ull) {
                top_val[o] = -INFINITY;
                top_idx[o] = -1;
            } else {
                top_val[o] = f32_from_orderable((uint32_t)(best >> 32));
                top_idx[o] = (int32_t)(uint32_t)(best & 0xffffffffull);
            }
        }
    }
}

// Final merge of per-shard top-N lists (place-sharded database, BASELINE config 5): every rank ranks its own
// range of places, the W lists of a query are all-gathered, and the global top-N is the N best of their
// union under the same order (value desc, place index desc).  One warp per query; the <= W*N candidate keys
// sit in registers (kMergeSlots per lane), N rounds of warp-wide 64-bit max.
constexpr int kMergeSlots = 16;      // W * N <= 32 * kMergeSlots

__global__ void __launch_bounds__(256) topn_merge_kernel(const float *__restrict__ val, const int32_t *__restrict__ idx,
                                                         int W, long long M, int N, float *__restrict__ out_val,
                                                         int32_t *__restrict__ out_idx)
{
    const int lane = threadIdx.x & 31;
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const int n_cand = W * N;
    unsigned long long key[kMergeSlots];
#pragma unroll
    for (int s = 0; s < kMergeSlots; ++s) {
        const int c = s * 32 + lane;
        key[s] = 0ull;
        if (c < n_cand) {
            const int w = c / N, n = c - w * N;
            const size_t o = ((size_t)w * M + m) * N + n;
            const int32_t i = idx[o];
            if (i >= 0) key[s] = ((unsigned long long)f32_orderable(val[o]) << 32) | (uint32_t)i;
        }
    }
    for (int n = 0; n < N; ++n) {
        unsigned long long mine = 0ull;
#pragma unroll
        for (int s = 0; s < kMergeSlots; ++s) mine = key[s] > mine ? key[s] : mine;
        unsigned long long best = mine;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = shfl_xor_u64(best, o);
            best = other > best ? other : best;
        }
        if (best != 0ull && mine == best) {            // keys are unique (distinct place indices): one owner
#pragma unroll
            for (int s = 0; s < kMergeSlots; ++s)
                if (key[s] == best) key[s] = 0ull;
        }
        if (lane == 0) {
            const size_t o = (size_t)m * N + n;
            out_val[o] = best ? f32_from_orderable((uint32_t)(best >> 32)) : -INFINITY;
            out_idx[o] = best ? (int32_t)(uint32_t)(best & 0xffffffffull) : -1;
        }
    }
}

struct RecallParams {
    const int32_t *top_idx;
    int B, Qo, Po, N;
    const uint8_t *gt_dense;
    int64_t gt_stream_stride;
    const int32_t *gt_center;
    int gt_tol;
    int ns[8];
    int n_ns;
    unsigned long long *hits, *n_valid;
};

// One thread per (stream, query).  metrics.py:214-216 drops queries without any
// positive; :218-224 counts a query as recalled when one of its K best is positive.
__global__ void __launch_bounds__(256) recall_kernel(RecallParams p)
{
    __shared__ unsigned int s_hits[8];
    __shared__ unsigned int s_valid;
    if (threadIdx.x < 8) s_hits[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < (int64_t)p.B * p.Qo) {
        const int b = (int)(g / p.Qo), q = (int)(g - (int64_t)b * p.Qo);
        const int32_t *ti = p.top_idx + (size_t)g * p.N;
        bool valid;
        int first_hit = INT_MAX;   // rank of the best-ranked positive among the top N
        if (p.gt_dense) {
            const uint8_t *gt = p.gt_dense + (size_t)b * p.gt_stream_stride;
            valid = false;
            for (int r = 0; r < p.Po; ++r) valid |= gt[(size_t)r * p.Qo + q] != 0;
            if (valid)
                for (int n = 0; n < p.N; ++n) {
                    const int r = ti[n];
                    if (r >= 0 && gt[(size_t)r * p.Qo + q] != 0) { first_hit = n; break; }
                }
        } else {
            const int c = p.gt_center[g];
            valid = c >= 0;
            if (valid)
                for (int n = 0; n < p.N; ++n) {
                    const int r = ti[n];
                    if (r >= 0 && abs(r - c) <= p.gt_tol) { first_hit = n; break; }
                }
        }
        if (valid) {
            atomicAdd(&s_valid, 1u);
            for (int i = 0; i < p.n_ns; ++i)
                if (first_hit < p.ns[i]) atomicAdd(&s_hits[i], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < p.n_ns && s_hits[threadIdx.x])
        atomicAdd(p.hits + threadIdx.x, (unsigned long long)s_hits[threadIdx.x]);
    if (threadIdx.x == 0 && s_valid) atomicAdd(p.n_valid, (unsigned long long)s_valid);
}

// Tie-aware bounds of Recall@K: lens/src/metrics.py:218 ranks every query column with numpy's
// default argsort, which is free to order equal similarities either way, so the reference's own number
// is only defined up to the choice among ties.  One warp per query column walks the distinct similarity
// values from the top: g = entries strictly above the current value (all of them are in any top-K with
// K > g), c = entries equal to it.  For the K whose K-th entry falls on this value (g < K <= g + c):
//   certain hit  <=> a positive lies strictly above, or the K - g picks among the c ties cannot avoid one
//   possible hit <=> a positive lies strictly above or among the ties.
struct BoundsParams {
    const float *D;            // [Po][Qo]
    const uint8_t *GT;         // [Po][Qo]
    int Po, Qo, n_ns;
    int ns[8];
    unsigned long long *lo, *hi, *n_valid;
};

__global__ void __launch_bounds__(256) recall_bounds_kernel(BoundsParams p)
{
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.Qo) return;
    int any = 0;
    for (int r = lane; r < p.Po; r += 32) any |= p.GT[(size_t)r * p.Qo + q] != 0;
    if (!__any_sync(0xffffffffu, any)) return;                   // metrics.py:214-216: no positive, dropped
    const int kmax = p.ns[p.n_ns - 1];
    float cur = INFINITY;
    int g = 0, gpos = 0, next_k = 0;                              // next_k: first K not decided yet
    unsigned lo_mask = 0u, hi_mask = 0u;
    while (next_k < p.n_ns) {
        // largest value strictly below `cur`, its multiplicity and its positives
        float m = -INFINITY;
        for (int r = lane; r < p.Po; r += 32) {
            const float v = p.D[(size_t)r * p.Qo + q];
            if (v < cur) m = fmaxf(m, v);
        }
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        int c = 0, cpos = 0;
        for (int r = lane; r < p.Po; r += 32) {
            const float v = p.D[(size_t)r * p.Qo + q];
            if (v == m && v < cur) { ++c; cpos += p.GT[(size_t)r * p.Qo + q] != 0; }
        }
        c = __reduce_add_sync(0xffffffffu, c);
        cpos = __reduce_add_sync(0xffffffffu, cpos);
        if (c == 0) {
            // the column is exhausted (K > Po, or only NaNs are left): everything ranked so far is certain
            for (; next_k < p.n_ns; ++next_k)
                if (gpos > 0) { lo_mask |= 1u << next_k; hi_mask |= 1u << next_k; }
            break;
        }
        for (; next_k < p.n_ns && p.ns[next_k] <= g + c; ++next_k) {
            const int room = p.ns[next_k] - g;                    // picks among the c ties
            if (gpos > 0) { lo_mask |= 1u << next_k; hi_mask |= 1u << next_k; }
            else if (cpos > 0) {
                hi_mask |= 1u << next_k;
                if (c - cpos < room) lo_mask |= 1u << next_k;
            }
        }
        g += c; gpos += cpos; cur = m;
        if (g >= kmax) break;
    }
    if (lane == 0) {
        atomicAdd(p.n_valid, 1ull);
        for (int i = 0; i < p.n_ns; ++i) {
            if (lo_mask >> i & 1u) atomicAdd(p.lo + i, 1ull);
            if (hi_mask >> i & 1u) atomicAdd(p.hi + i, 1ull);
        }
    }
}

// createPR, matching = 'single' (lens/src/metrics.py:21-139): one CTA, columns strided over threads.
// best[q] / hit[q] live in shared memory (Qo <= 8192).
__global__ void __launch_bounds__(256) pr_counts_kernel(const float *__restrict__ S, const uint8_t *__restrict__ GT,
                                                        int Po, int Qo, int n_thresh,
                                                        unsigned long long *__restrict__ tp,
                                                        unsigned long long *__restrict__ fp,
                                                        unsigned long long *__restrict__ gtp)
{
    extern __shared__ float s_best[];                       // [Qo] best similarity per query
    uint8_t *s_hit = reinterpret_cast<uint8_t *>(s_best + Qo);   // [Qo] GT at the best match
    __shared__ float s_max[8], s_min[8];
    __shared__ unsigned int s_gtp;
    if (threadIdx.x == 0) s_gtp = 0;
    __syncthreads();
    float vmax = -INFINITY, vmin = INFINITY;
    unsigned int my_gtp = 0;
    for (int q = threadIdx.x; q < Qo; q += blockDim.x) {
        float best = -INFINITY;
        int arg = 0;
        bool any = false;
        for (int r = 0; r < Po; ++r) {
            const float v = S[(size_t)r * Qo + q];
            if (v > best) { best = v; arg = r; }             // strict: first maximum wins (np.argmax)
            any |= GT[(size_t)r * Qo + q] != 0;
        }
        s_best[q] = best;
        s_hit[q] = GT[(size_t)arg * Qo + q] != 0;
        my_gtp += any;
        vmax = fmaxf(vmax, best); vmin = fminf(vmin, best);
    }
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        my_gtp += __shfl_xor_sync(0xffffffffu, my_gtp, o);
    }
    if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = vmax; s_min[threadIdx.x >> 5] = vmin; atomicAdd(&s_gtp, my_gtp); }
    __syncthreads();
    for (int i = 0; i < 8; ++i) { vmax = fmaxf(vmax, s_max[i]); vmin = fminf(vmin, s_min[i]); }
    // np.linspace(start, stop, n): y_i = i * step + start in float64, last element = stop
    const double start = (double)vmax, stop = (double)vmin;
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(n_thresh - 1));
    for (int i = threadIdx.x; i < n_thresh; i += blockDim.x) {
        // numpy: y = arange(n) * step; y += start  (two separately rounded float64 operations, no FMA)
        const double t = (i == n_thresh - 1) ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
        unsigned long long a = 0, b = 0;
        for (int q = 0; q < Qo; ++q) {
            if ((double)s_best[q] >= t) { if (s_hit[q]) ++a; else ++b; }
        }
        tp[i] = a; fp[i] = b;
    }
    if (threadIdx.x == 0) gtp[0] = s_gtp;
}

// createPR, matching = 'multi' (lens/src/metrics.py:63-91): every entry of the matrix counts.  Pass 1: extreme
// values (as order-preserving integers) and the number of ground-truth positives; pass 2: one CTA per (threshold,
// slice of the matrix) counts the entries >= threshold that are / are not positives.  Thresholds as in the
// 'single' kernel: np.linspace(max, min, n) evaluated in float64 without FMA contraction.
__global__ void __launch_bounds__(256) pr_multi_minmax_kernel(const float *__restrict__ S, const uint8_t *__restrict__ GT,
                                                              long long n, unsigned int *__restrict__ mm,
                                                              unsigned long long *__restrict__ gtp)
{
    unsigned int vmax = 0u, vmin = 0xffffffffu;
    unsigned long long pos = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int k = f32_orderable(S[i]);
        vmax = max(vmax, k); vmin = min(vmin, k);
        pos += GT[i] != 0;
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(mm, vmax);
        atomicMin(mm + 1, vmin);
        if (pos) atomicAdd(gtp, pos);
    }
}

__global__ void __launch_bounds__(256) pr_multi_counts_kernel(const float *__restrict__ S, const uint8_t *__restrict__ GT,
                                                              long long n, int n_thresh, const unsigned int *__restrict__ mm,
                                                              unsigned long long *__restrict__ tp,
                                                              unsigned long long *__restrict__ fp)
{
    const int i = blockIdx.x;
    const double start = (double)f32_from_orderable(mm[0]), stop = (double)f32_from_orderable(mm[1]);
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(n_thresh - 1));
    const double t = (i == n_thresh - 1) ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
    const long long per = ceil_div64(n, gridDim.y), e0 = per * blockIdx.y, e1 = min(n, e0 + per);
    unsigned long long a = 0, b = 0;
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x)
        if ((double)S[e] >= t) { if (GT[e]) ++a; else ++b; }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (a) atomicAdd(tp + i, a);
        if (b) atomicAdd(fp + i, b);
    }
}

// SAD baseline (lens/src/sad.py:38): one warp per (query, reference) pair, 16 pixels per lane-iteration
// with byte-wise SIMD sum of absolute differences; integer accumulation (exact), stored as fp32.
__global__ void __launch_bounds__(256) sad_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                                                  int Q, int R, int npix, float *__restrict__ dist)
{
    const int lane = threadIdx.x & 31;
    const long long pair = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (pair >= (long long)Q * R) return;
    const int q = (int)(pair / R), r = (int)(pair - (long long)q * R);
    const uint8_t *pa = a + (size_t)q * npix, *pb = b + (size_t)r * npix;
    unsigned int acc = 0;
    if ((npix & 15) == 0 && ((((uintptr_t)a | (uintptr_t)b)) & 15) == 0) {
        const uint4 *va = reinterpret_cast<const uint4 *>(pa), *vb = reinterpret_cast<const uint4 *>(pb);
        for (int i = lane; i < npix / 16; i += 32) {
            const uint4 x = __ldg(va + i), y = __ldg(vb + i);
            acc += __vsadu4(x.x, y.x) + __vsadu4(x.y, y.y) + __vsadu4(x.z, y.z) + __vsadu4(x.w, y.w);
        }
    } else {
        for (int i = lane; i < npix; i += 32) acc += (unsigned)abs((int)pa[i] - (int)pb[i]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dist[pair] = (float)acc;
}

__global__ void reciprocal_kernel(const float *__restrict__ in, int64_t n, float *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fdiv_rn(1.0f, in[i]);
}

// Online matcher (run_speck.py:155-226): a readout adds its spike counts to the running sums; every
// `div` readouts the sums, floor-divided, become one sequence row.
__global__ void online_accumulate_kernel(int32_t *__restrict__ sum, const float *__restrict__ counts, int P, int div,
                                         int32_t *__restrict__ row_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int32_t s = sum[i] + (int32_t)counts[i];
    sum[i] = s;
    if (row_out) {
        int32_t q = s / div;
        if ((s % div != 0) && ((s < 0) != (div < 0))) --q;      // floor division like numpy's //
        row_out[i] = q;
    }
}

// result[p][r] = (1/L) sum_a seq[r+o-a][p+o-a] ('same' crop of the full convolution with eye(L)), then the
// first maximum of every column r.  One CTA per column: coalesced reads along p for each of the L taps.
__global__ void __launch_bounds__(256) online_match_kernel(const int32_t *__restrict__ seq, int R, int P, int L,
                                                           double *__restrict__ result, int32_t *__restrict__ argmax)
{
    const int r = blockIdx.x, o = (L - 1) / 2;
    double best = 0.0;
    int best_p = 0x7fffffff;
    bool have = false;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        long long acc = 0;
        for (int a = 0; a < L; ++a) {
            const int rr = r + o - a, pp = p + o - a;
            if (rr >= 0 && rr < R && pp >= 0 && pp < P) acc += seq[(size_t)rr * P + pp];
        }
        const double v = __ddiv_rn((double)acc, (double)L);
        result[(size_t)p * R + r] = v;
        if (!have || v > best) { best = v; best_p = p; have = true; }   // p ascending per thread: first max kept
    }
    __shared__ double sv[256];
    __shared__ int sp[256];
    sv[threadIdx.x] = best; sp[threadIdx.x] = best_p;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v2 = sv[threadIdx.x + s];
            const int p2 = sp[threadIdx.x + s];
            const int p1 = sp[threadIdx.x];
            if (p2 != 0x7fffffff && (p1 == 0x7fffffff || v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && p2 < p1))) {
                sv[threadIdx.x] = v2; sp[threadIdx.x] = p2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) argmax[r] = sp[0];
}

}  // namespace lens

using namespace lens;

extern "C" int lens_online_accumulate(int32_t *sum, const float *counts, int P, int div, int32_t *row_out, void *stream)
{
    LENS_CHECK_ARG(sum && counts, "lens_online_accumulate: NULL buffer");
    LENS_CHECK_ARG(P > 0 && div != 0, "lens_online_accumulate: need P > 0 and div != 0");
    online_accumulate_kernel<<<(unsigned)((P + 255) / 256), 256, 0, as_stream(stream)>>>(sum, counts, P, div, row_out);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_online_match(const int32_t *seq, int R, int P, int L, double *result, int32_t *argmax, void *stream)
{
    LENS_CHECK_ARG(seq && result && argmax, "lens_online_match: NULL buffer");
    LENS_CHECK_ARG(P > 0 && R >= 1 && R <= 64 && L >= 1 && L <= 64, "lens_online_match: need P > 0, 1 <= R, L <= 64");
    online_match_kernel<<<(unsigned)R, 256, 0, as_stream(stream)>>>(seq, R, P, L, result, argmax);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_sad_matrix(const uint8_t *a, const uint8_t *b, int Q, int R, int npix, float *dist, void *stream)
{
    LENS_CHECK_ARG(a && b && dist, "lens_sad_matrix: NULL buffer");
    LENS_CHECK_ARG(Q > 0 && R > 0 && npix > 0 && npix <= 65793, "lens_sad_matrix: need Q, R > 0 and 0 < npix <= 65793 "
                   "(sums must stay below 2^24)");
    const long long pairs = (long long)Q * R;
    sad_kernel<<<(unsigned)((pairs * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(a, b, Q, R, npix, dist);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_reciprocal(const float *in, int64_t n, float *out, void *stream)
{
    LENS_CHECK_ARG(in && out && n >= 0, "lens_reciprocal: bad arguments");
    if (n == 0) return 0;
    reciprocal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(in, n, out);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_pr_counts(const float *S, const uint8_t *GT, int Po, int Qo, int n_thresh, int64_t *tp,
                              int64_t *fp, int64_t *gtp, void *stream)
{
    LENS_CHECK_ARG(S && GT && tp && fp && gtp, "lens_pr_counts: NULL buffer");
    LENS_CHECK_ARG(Po > 0 && Qo > 0 && Qo <= 8192, "lens_pr_counts: need 0 < Qo <= 8192, Po > 0");
    LENS_CHECK_ARG(n_thresh > 1, "lens_pr_counts: n_thresh must be > 1");
    pr_counts_kernel<<<1, 256, (size_t)Qo * 5, as_stream(stream)>>>(
        S, GT, Po, Qo, n_thresh, reinterpret_cast<unsigned long long *>(tp),
        reinterpret_cast<unsigned long long *>(fp), reinterpret_cast<unsigned long long *>(gtp));
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_seqmatch_topk(const float *S, int B, int Q, int P, int L, int N, float *D_out,
                                  float *top_val, int32_t *top_idx, void *stream)
{
    LENS_CHECK_ARG(B >= 0 && Q > 0 && P > 0, "lens_seqmatch_topk: bad sizes");
    LENS_CHECK_ARG(L >= 1 && L <= Q && L <= P, "lens_seqmatch_topk: need 1 <= L <= min(Q, P), got L=%d", L);
    LENS_CHECK_ARG(N >= 1 && N <= kMaxTopN, "lens_seqmatch_topk: N=%d outside [1, %d]", N, kMaxTopN);
    LENS_CHECK_ARG(B == 0 || (S && top_val && top_idx), "lens_seqmatch_topk: NULL buffer");
    LENS_CHECK_ARG((int64_t)B * (Q - L + 1) <= 2147483647LL, "lens_seqmatch_topk: too many (stream, query) pairs");
    if (B == 0) return 0;
    const long long n_queries = (long long)B * (Q - L + 1);
    if (N <= 32) {
        // Few queries against many places (config 5: 1 024 streams x 100 000 places) would leave most SMs without
        // a warp: the places of a query are then split into segments ranked by different warps, and the per-segment
        // lists are merged by the kernel that merges per-GPU lists (same order: value desc, place index desc).
        const int Po = P - L + 1;
        const long long sms = std::max(sm_count(), 1), want_warps = sms * 32;
        int n_seg = 1;
        if (n_queries < want_warps && Po >= 8192)
            n_seg = (int)std::min<long long>(std::min<long long>(ceil_div64(want_warps, n_queries), (32 * kMergeSlots) / N),
                                             Po / 4096);
        n_seg = std::max(n_seg, 1);
        const int seg_len = ceil_div(ceil_div(Po, n_seg), 32 * kUnrollR) * (32 * kUnrollR);
        n_seg = ceil_div(Po, seg_len);
        const long long n_work = n_queries * n_seg;
        const long long blocks = std::min<long long>((n_work + 7) / 8, sms * 16);
        cudaStream_t st = as_stream(stream);
        if (n_seg == 1) {
            seqmatch_topk_warp_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(S, n_queries, Q, P, L, N, 1, seg_len, D_out,
                                                                              top_val, top_idx);
        } else {
            // stream-ordered scratch for the per-segment lists (no synchronisation)
            float *seg_val = nullptr;
            int32_t *seg_idx = nullptr;
            const size_t n_list = (size_t)n_work * N;
            LENS_CUDA(cudaMallocAsync(&seg_val, n_list * (sizeof(float) + sizeof(int32_t)), st));   // one block: values | indices
            seg_idx = reinterpret_cast<int32_t *>(seg_val + n_list);
            seqmatch_topk_warp_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(S, n_queries, Q, P, L, N, n_seg, seg_len, D_out,
                                                                             seg_val, seg_idx);
            LENS_LAUNCH_CHECK();
            topn_merge_kernel<<<(unsigned)ceil_div64(n_queries, 8), 256, 0, st>>>(seg_val, seg_idx, n_seg, n_queries, N,
                                                                                   top_val, top_idx);
            LENS_CUDA(cudaFreeAsync(seg_val, st));
        }
    } else {
        dim3 grid((unsigned)n_queries);
        seqmatch_topk_kernel<<<grid, kMatchThreads, 0, as_stream(stream)>>>(S, Q, P, L, N, D_out, top_val, top_idx);
    }
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_recall(const int32_t *top_idx, int B, int Qo, int Po, int N,
                           const uint8_t *gt_dense, int64_t gt_stream_stride,
                           const int32_t *gt_center, int gt_tol, const int *ns, int n_ns,
                           int64_t *hits, int64_t *n_valid, void *stream)
{
    LENS_CHECK_ARG(B >= 0 && Qo > 0 && Po > 0 && N >= 1, "lens_recall: bad sizes");
    LENS_CHECK_ARG((gt_dense != nullptr) != (gt_center != nullptr),
                   "lens_recall: exactly one of gt_dense / gt_center must be given");
    LENS_CHECK_ARG(ns && n_ns >= 1 && n_ns <= 8, "lens_recall: n_ns must be in [1, 8]");
    LENS_CHECK_ARG(top_idx && hits && n_valid, "lens_recall: NULL buffer");
    RecallParams p;
    p.top_idx = top_idx; p.B = B; p.Qo = Qo; p.Po = Po; p.N = N;
    p.gt_dense = gt_dense; p.gt_stream_stride = gt_stream_stride;
    p.gt_center = gt_center; p.gt_tol = gt_tol; p.n_ns = n_ns;
    for (int i = 0; i < 8; ++i) p.ns[i] = 0;
    for (int i = 0; i < n_ns; ++i) {
        LENS_CHECK_ARG(ns[i] >= 1 && ns[i] <= N, "lens_recall: ns[%d]=%d outside [1, N=%d]", i, ns[i], N);
        p.ns[i] = ns[i];
    }
    p.hits = reinterpret_cast<unsigned long long *>(hits);
    p.n_valid = reinterpret_cast<unsigned long long *>(n_valid);
    if (B == 0) return 0;
    int64_t total = (int64_t)B * Qo;
    recall_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(p);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_recall_bounds(const float *D, const uint8_t *GT, int Po, int Qo, const int *ns, int n_ns,
                                  int64_t *lo, int64_t *hi, int64_t *n_valid, void *stream)
{
    LENS_CHECK_ARG(Po > 0 && Qo > 0, "lens_recall_bounds: bad sizes");
    LENS_CHECK_ARG(D && GT && lo && hi && n_valid, "lens_recall_bounds: NULL buffer");
    LENS_CHECK_ARG(ns && n_ns >= 1 && n_ns <= 8, "lens_recall_bounds: n_ns must be in [1, 8]");
    BoundsParams p;
    p.D = D; p.GT = GT; p.Po = Po; p.Qo = Qo; p.n_ns = n_ns;
    for (int i = 0; i < 8; ++i) p.ns[i] = 0;
    for (int i = 0; i < n_ns; ++i) {
        LENS_CHECK_ARG(ns[i] >= 1 && (i == 0 || ns[i] > ns[i - 1]), "lens_recall_bounds: ns must be ascending and >= 1");
        p.ns[i] = ns[i];
    }
    p.lo = reinterpret_cast<unsigned long long *>(lo);
    p.hi = reinterpret_cast<unsigned long long *>(hi);
    p.n_valid = reinterpret_cast<unsigned long long *>(n_valid);
    recall_bounds_kernel<<<(unsigned)ceil_div(Qo, 8), 256, 0, as_stream(stream)>>>(p);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_topn_merge(const float *val, const int32_t *idx, int W, int64_t M, int N, float *out_val,
                               int32_t *out_idx, void *stream)
{
    LENS_CHECK_ARG(W >= 1 && M >= 0 && N >= 1, "lens_topn_merge: bad sizes");
    LENS_CHECK_ARG((int64_t)W * N <= 32 * kMergeSlots, "lens_topn_merge: W * N = %lld exceeds %d", (long long)W * N,
                   32 * kMergeSlots);
    if (M == 0) return 0;
    LENS_CHECK_ARG(val && idx && out_val && out_idx, "lens_topn_merge: NULL buffer");
    topn_merge_kernel<<<(unsigned)ceil_div64(M, 8), 256, 0, as_stream(stream)>>>(val, idx, W, M, N, out_val, out_idx);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_pr_counts_multi(const float *S, const uint8_t *GT, int Po, int Qo, int n_thresh, int64_t *tp,
                                    int64_t *fp, int64_t *gtp, void *stream)
{
    LENS_CHECK_ARG(Po > 0 && Qo > 0, "lens_pr_counts_multi: bad sizes");
    LENS_CHECK_ARG(n_thresh > 1 && n_thresh <= 65535, "lens_pr_counts_multi: n_thresh must be in (1, 65535]");
    LENS_CHECK_ARG(S && GT && tp && fp && gtp, "lens_pr_counts_multi: NULL buffer");
    cudaStream_t st = as_stream(stream);
    const long long n = (long long)Po * Qo;
    unsigned int *mm = nullptr;                       // {max, min} of S as order-preserving integers
    LENS_CUDA(cudaMallocAsync(&mm, 2 * sizeof(unsigned int), st));
    LENS_CUDA(cudaMemsetAsync(mm, 0x00, sizeof(unsigned int), st));
    LENS_CUDA(cudaMemsetAsync(mm + 1, 0xff, sizeof(unsigned int), st));
    LENS_CUDA(cudaMemsetAsync(tp, 0, (size_t)n_thresh * sizeof(int64_t), st));
    LENS_CUDA(cudaMemsetAsync(fp, 0, (size_t)n_thresh * sizeof(int64_t), st));
    LENS_CUDA(cudaMemsetAsync(gtp, 0, sizeof(int64_t), st));
    const int blocks = (int)std::min<long long>(ceil_div64(n, 256), (long long)std::max(sm_count(), 1) * 8);
    pr_multi_minmax_kernel<<<blocks, 256, 0, st>>>(S, GT, n, mm, reinterpret_cast<unsigned long long *>(gtp));
    LENS_LAUNCH_CHECK();
    const int slices = (int)std::max<long long>(1, std::min<long long>(ceil_div64(n, 65536), 64));
    pr_multi_counts_kernel<<<dim3((unsigned)n_thresh, (unsigned)slices), 256, 0, st>>>(
        S, GT, n, n_thresh, mm, reinterpret_cast<unsigned long long *>(tp), reinterpret_cast<unsigned long long *>(fp));
    LENS_LAUNCH_CHECK();
    LENS_CUDA(cudaFreeAsync(mm, st));
    return 0;
}
