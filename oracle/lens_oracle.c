/*
 * lens_oracle.c -- CPU restatement of the LENS inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it
 * may be imported / linked / executed only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, and there only as the
 * checker or the timed CPU baseline.  The product (lens_b200/) never calls it.
 *
 * Parity pinning: the reference ships no tests and its SNN arithmetic lives in
 * the third-party `sinabs` package (requirements.txt:8, `sinabs>=2.0.0`,
 * unpinned, NOT installable in the build container).  This file is checked
 * against golden vectors produced by running the reference's own
 * run_model.py / dataset.py / metrics.py with a torch restatement of the
 * sinabs layers (tests/golden/make_golden.py): "sinabs-restated" pinning.
 *
 * Each function cites the reference site (relative to /root/reference) it
 * restates.  Plain C99 + __int128 (gcc), no dependencies.
 *
 * Arithmetic contract (the one place the reference is implementation-defined):
 * the two Linear contractions (run_model.py:143,145 -> torch F.linear, i.e.
 * whatever summation order the BLAS picks) are evaluated EXACTLY here -- every
 * spike count is a small integer, every weight an fp32 number, the sum is
 * accumulated in integer arithmetic and rounded ONCE to fp32 (round to nearest
 * even).  That is the correctly rounded value of the expression the reference
 * writes down, it is independent of summation order, and any fp32 BLAS result
 * lies within a few ulp of it.  Everything else (IAF update, clamp, division,
 * comparisons) is elementwise IEEE fp32 exactly as torch evaluates it.
 * Compile with -ffp-contract=off (see Makefile) so no FMA is formed.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_SPIKE 127 /* per-step spike count that stays exact in every path */

/* ------------------------------------------------------------------------- */
/* R0: event binning -- lens/collect_data.py:193-202 (create_images) and the
 * duplicate lens/tools/manual_eventframe_generator.py:6-14:
 *     frame = zeros((roi, roi), int); for ev: frame[ev.y-1, ev.x-1] += 1
 *     imwrite(frame.astype(uint8))            -> counts wrap modulo 256
 * One frame per `window_us` slice of the time axis (collect_data.py:186-191
 * drains the sink every `timebin` ms).  Index -1 wraps to the last row/column
 * (torch negative indexing).  Events whose shifted index lies outside the range
 * the reference's tensor accepts, [-roi, roi-1], are cropped (on-chip ROI,
 * collect_data.py:230-233; the Speck crop is 81 columns wide, so x == roi occurs).  win_events[w] counts the in-ROI events of window
 * w so the caller can drop empty windows like create_images does (:194).
 * frames: [n_win, roi, roi] u8, pooled (R1 applied): [n_win, dims*dims] u8.   */
int lens_oracle_bin_events(const uint32_t *t_us, const uint16_t *x, const uint16_t *y,
                           int64_t n_events, uint32_t t0_us, uint32_t window_us,
                           int roi_x0, int roi_y0, int roi, int k, int index_shift,
                           int wrap_u8, uint8_t *frames, uint8_t *pooled,
                           int32_t *win_events, int64_t n_win)
{
    if (roi <= 0 || k <= 0 || window_us == 0) return -1;
    int dims = roi / k;
    int c = (k / 2) - 1; /* run_model.py:132 centre_coordinate */
    if (c < 0) c += k;   /* k == 1: index -1 wraps to 0 (python negative index) */
    int64_t npix = (int64_t)roi * roi;
    int64_t *acc = (int64_t *)calloc((size_t)npix, sizeof(int64_t));
    if (!acc) return -2;
    int64_t e = 0;
    for (int64_t w = 0; w < n_win; ++w) {
        memset(acc, 0, (size_t)npix * sizeof(int64_t));
        int32_t cnt = 0;
        uint64_t lo = (uint64_t)t0_us + (uint64_t)w * window_us;
        uint64_t hi = lo + window_us;
        while (e < n_events && (uint64_t)t_us[e] < lo) ++e; /* before t0: dropped */
        for (; e < n_events && (uint64_t)t_us[e] < hi; ++e) {
            /* every index the reference's roi x roi tensor accepts: [-roi, roi-1], negatives wrap
             * (x == roi lands in column roi-1 when index_shift is 1); the rest would raise IndexError */
            int col = (int)x[e] - roi_x0 - index_shift, row = (int)y[e] - roi_y0 - index_shift;
            if (col < -roi || col >= roi || row < -roi || row >= roi) continue;
            if (col < 0) col += roi;
            if (row < 0) row += roi;
            acc[(int64_t)row * roi + col] += 1;
            ++cnt;
        }
        uint8_t *f = frames + w * npix;
        for (int64_t p = 0; p < npix; ++p) {
            int64_t v = acc[p];
            f[p] = wrap_u8 ? (uint8_t)(v & 255) : (uint8_t)(v > 255 ? 255 : v);
        }
        if (pooled) { /* R1: y[i,j] = x[k*i+c, k*j+c] */
            uint8_t *pw = pooled + w * (int64_t)dims * dims;
            for (int i = 0; i < dims; ++i)
                for (int j = 0; j < dims; ++j)
                    pw[i * dims + j] = f[(int64_t)(k * i + c) * roi + (k * j + c)];
        }
        if (win_events) win_events[w] = cnt;
    }
    free(acc);
    return 0;
}

/* R1: one-hot strided Conv2d(1,1,k,stride=k) -- lens/run_model.py:130-137:
 * kernel[c,c] = 1 with c = (k//2)-1  =>  y[i,j] = x[k*i+c, k*j+c] (a pixel pick).
 * frames [n, roi, roi] u8 -> pooled [n, dims*dims] u8.                         */
int lens_oracle_pool(const uint8_t *frames, int64_t n, int roi, int k, uint8_t *pooled)
{
    if (roi <= 0 || k <= 0) return -1;
    int dims = roi / k;
    int c = (k / 2) - 1;
    if (c < 0) c += k;
    for (int64_t f = 0; f < n; ++f)
        for (int i = 0; i < dims; ++i)
            for (int j = 0; j < dims; ++j)
                pooled[f * dims * dims + i * dims + j] =
                    frames[f * (int64_t)roi * roi + (int64_t)(k * i + c) * roi + (k * j + c)];
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Exact contraction support.                                                  */

typedef struct {
    int n_out, n_in;
    int32_t *mant; /* [n_in][n_out] signed 24-bit significand (transposed for the sparse loop) */
    uint8_t *shift;/* [n_in][n_out] left shift relative to the row's smallest ulp exponent     */
    int64_t *w64;  /* [n_in][n_out] mant * 2^shift for rows that fit 64-bit accumulation, else 0 */
    int *emin;     /* [n_out] exponent of the row's smallest ulp: w = mant * 2^(emin+shift)    */
    float *scale;  /* [n_out] 2^emin when that is a normal float, else 0 (-> ldexpf path)      */
    int *wide;     /* [n_out] 1 if the row needs the 128-bit accumulator                       */
    int any_wide, unsupported;
} exact_layer;

static void exact_layer_free(exact_layer *L)
{
    free(L->mant); free(L->shift); free(L->w64); free(L->emin); free(L->wide); free(L->scale);
    memset(L, 0, sizeof(*L));
}

/* w[n_out][n_in] fp32 -> per-row fixed point. */
static int exact_layer_init(exact_layer *L, const float *w, int n_out, int n_in)
{
    memset(L, 0, sizeof(*L));
    L->n_out = n_out; L->n_in = n_in;
    L->mant = (int32_t *)calloc((size_t)n_out * n_in, sizeof(int32_t));
    L->shift = (uint8_t *)calloc((size_t)n_out * n_in, 1);
    L->w64 = (int64_t *)calloc((size_t)n_out * n_in, sizeof(int64_t));
    L->emin = (int *)calloc((size_t)n_out, sizeof(int));
    L->wide = (int *)calloc((size_t)n_out, sizeof(int));
    L->scale = (float *)calloc((size_t)n_out, sizeof(float));
    if (!L->scale || !L->mant || !L->shift || !L->w64 || !L->emin || !L->wide) return -2;
    for (int n = 0; n < n_out; ++n) {
        int lo = 1 << 30, hi = -(1 << 30);
        for (int kk = 0; kk < n_in; ++kk) {
            float v = w[(size_t)n * n_in + kk];
            if (v == 0.0f || !isfinite(v)) continue;
            int e; float m = frexpf(v, &e); (void)m;      /* |v| in [2^(e-1), 2^e) */
            int ulp = e - 24;
            if (ulp < lo) lo = ulp;
            if (e > hi) hi = e;
        }
        if (hi < lo) { L->emin[n] = 0; continue; }        /* all-zero row */
        L->emin[n] = lo;
        if (lo >= -126 && lo <= 127) L->scale[n] = ldexpf(1.0f, lo);   /* exact power of two */
        int span = hi - lo;                                /* bits needed for the largest |w| */
        /* s <= 127 (7 bits), n_in terms (<= 2^16) -> headroom 23 bits */
        if (span + 23 > 62) { L->wide[n] = 1; L->any_wide = 1; }
        if (span + 23 > 126) L->unsupported = 1;
        for (int kk = 0; kk < n_in; ++kk) {
            float v = w[(size_t)n * n_in + kk];
            if (v == 0.0f || !isfinite(v)) continue;
            int e; float m = frexpf(v, &e);
            L->mant[(size_t)kk * n_out + n] = (int32_t)ldexpf(m, 24); /* exact: 24-bit significand */
            int sh = (e - 24) - lo;
            L->shift[(size_t)kk * n_out + n] = (uint8_t)(sh > 200 ? 200 : sh);
            if (!L->wide[n])
                L->w64[(size_t)kk * n_out + n] =
                    (int64_t)L->mant[(size_t)kk * n_out + n] * ((int64_t)1 << sh);
        }
    }
    return 0;
}

/* Correctly rounded (RNE) fp32 value of a * 2^e for a signed 128-bit integer. */
static float i128_to_f32(__int128 a, int e)
{
    if (a == 0) return 0.0f;
    int neg = a < 0;
    unsigned __int128 u = neg ? (unsigned __int128)(-a) : (unsigned __int128)a;
    int top = 127;
    while (!((u >> top) & 1)) --top;                       /* index of the leading one */
    if (top > 23) {
        int drop = top - 23;
        unsigned __int128 keep = u >> drop;
        unsigned __int128 rem = u & ((((unsigned __int128)1) << drop) - 1);
        unsigned __int128 half = ((unsigned __int128)1) << (drop - 1);
        if (rem > half || (rem == half && (keep & 1))) ++keep;
        /* keep may become 2^24: still exactly representable */
        float r = ldexpf((float)(uint32_t)keep, e + drop);
        return neg ? -r : r;
    }
    float r = ldexpf((float)(uint32_t)u, e);
    return neg ? -r : r;
}

/* ------------------------------------------------------------------------- */
/* R3-R6: the converted sinabs network and the per-query loop.                 */

typedef struct lens_oracle_snn {
    int I, F, P, T, B;
    float thr, vmin;
    exact_layer L1, L2;
    float *U;             /* [T][I] */
    float *v0, *v1, *v2;  /* [B][I], [B][F], [B][P] membrane potentials (carried across calls) */
    int64_t overflow;     /* number of per-step spike counts above ORACLE_MAX_SPIKE */
} lens_oracle_snn;

/* R4: one timestep of sinabs IAFSqueeze (lif_forward_single with alpha_mem = 1,
 * MultiSpike, MembraneSubtract, min_v_mem clip); see tests/golden/sinabs_stub.py
 * for the torch form.  Returns the number of spikes emitted.                    */
static inline float iaf_step(float *v, float x, float thr, float vmin)
{
    float vv = 1.0f * (*v) + x;                         /* v = alpha * v + input             */
    float s = (vv > 0.0f) ? truncf(vv / thr) : 0.0f;    /* (v > 0) * div(v, thr, 'trunc')    */
    vv = vv - s * thr;                                  /* MembraneSubtract                  */
    float r = vv - vmin;                                /* relu(v - min_v_mem) + min_v_mem   */
    r = r > 0.0f ? r : 0.0f;
    *v = r + vmin;
    return s;
}

/* model assembly, R3: lens/run_model.py:139-156 + lens/src/blitnet.py:59-64
 * (bias-free Linear I->F, F->P; IAF after the pooling conv, after each Linear). */
lens_oracle_snn *lens_oracle_snn_create(int I, int F, int P, int T, float thr, float vmin,
                                        const float *W_feat /*[F][I]*/,
                                        const float *W_out /*[P][F]*/,
                                        const float *U /*[T][I], may be NULL*/, int n_streams)
{
    lens_oracle_snn *h = (lens_oracle_snn *)calloc(1, sizeof(*h));
    if (!h) return NULL;
    h->I = I; h->F = F; h->P = P; h->T = T; h->B = n_streams; h->thr = thr; h->vmin = vmin;
    if (exact_layer_init(&h->L1, W_feat, F, I) || exact_layer_init(&h->L2, W_out, P, F)) goto fail;
    if (h->L1.unsupported || h->L2.unsupported) goto fail;
    if (U) {
        h->U = (float *)malloc((size_t)T * I * sizeof(float));
        if (!h->U) goto fail;
        memcpy(h->U, U, (size_t)T * I * sizeof(float));
    }
    h->v0 = (float *)calloc((size_t)n_streams * I, sizeof(float));
    h->v1 = (float *)calloc((size_t)n_streams * F, sizeof(float));
    h->v2 = (float *)calloc((size_t)n_streams * P, sizeof(float));
    if (!h->v0 || !h->v1 || !h->v2) goto fail;
    return h;
fail:
    exact_layer_free(&h->L1); exact_layer_free(&h->L2);
    free(h->U); free(h->v0); free(h->v1); free(h->v2); free(h);
    return NULL;
}

void lens_oracle_snn_destroy(lens_oracle_snn *h)
{
    if (!h) return;
    exact_layer_free(&h->L1); exact_layer_free(&h->L2);
    free(h->U); free(h->v0); free(h->v1); free(h->v2); free(h);
}

/* sinabs Network.reset_states(): zero every membrane potential. */
void lens_oracle_snn_reset(lens_oracle_snn *h)
{
    memset(h->v0, 0, (size_t)h->B * h->I * sizeof(float));
    memset(h->v1, 0, (size_t)h->B * h->F * sizeof(float));
    memset(h->v2, 0, (size_t)h->B * h->P * sizeof(float));
    h->overflow = 0;
}

void lens_oracle_snn_get_state(const lens_oracle_snn *h, float *v0, float *v1, float *v2)
{
    if (v0) memcpy(v0, h->v0, (size_t)h->B * h->I * sizeof(float));
    if (v1) memcpy(v1, h->v1, (size_t)h->B * h->F * sizeof(float));
    if (v2) memcpy(v2, h->v2, (size_t)h->B * h->P * sizeof(float));
}

int64_t lens_oracle_snn_overflow(const lens_oracle_snn *h) { return h->overflow; }

/* x[n] = RN_f32( sum_k s[k] * w[n][k] ), s given as a sparse list (idx, cnt). */
static void exact_contract(const exact_layer *L, const int *idx, const int *cnt, int n_act,
                           float *x, int64_t *acc64, __int128 *acc128)
{
    int n_out = L->n_out;
    memset(acc64, 0, (size_t)n_out * sizeof(int64_t));
    if (L->any_wide) memset(acc128, 0, (size_t)n_out * sizeof(__int128));
    for (int a = 0; a < n_act; ++a) {
        const int32_t *m = L->mant + (size_t)idx[a] * n_out;
        const uint8_t *sh = L->shift + (size_t)idx[a] * n_out;
        const int64_t *w = L->w64 + (size_t)idx[a] * n_out;
        int64_t s = cnt[a];
        for (int n = 0; n < n_out; ++n) acc64[n] += s * w[n];
        if (L->any_wide)
            for (int n = 0; n < n_out; ++n)
                if (L->wide[n])
                    acc128[n] += (__int128)(s * m[n]) * ((__int128)1 << sh[n]);
    }
    for (int n = 0; n < n_out; ++n) {
        if (L->any_wide && L->wide[n]) x[n] = i128_to_f32(acc128[n], L->emin[n]);
        else if (L->scale[n] != 0.0f) x[n] = (float)acc64[n] * L->scale[n]; /* int64 -> f32: one RNE
                                                    rounding; the power-of-two scaling is exact   */
        else x[n] = ldexpf((float)acc64[n], L->emin[n]);
    }
}

/* One timestep of the whole network for one stream.  in_spk[i] = input to IAF#0
 * (already pooled).  Returns through s2 (float [P]) the output spikes.          */
static void snn_step(lens_oracle_snn *h, int b, const float *in, float *x1, float *x2,
                     int *idx, int *cnt, int64_t *acc64, __int128 *acc128,
                     uint8_t *hid_out, float *s2)
{
    int I = h->I, F = h->F, P = h->P;
    float *v0 = h->v0 + (size_t)b * I, *v1 = h->v1 + (size_t)b * F, *v2 = h->v2 + (size_t)b * P;
    int na = 0;
    for (int i = 0; i < I; ++i) {          /* IAF#0 (after the pooling conv, run_model.py:140-141) */
        float s = iaf_step(&v0[i], in[i], h->thr, h->vmin);
        if (s != 0.0f) {
            if (s > ORACLE_MAX_SPIKE) { h->overflow++; s = ORACLE_MAX_SPIKE; }
            idx[na] = i; cnt[na] = (int)s; ++na;
        }
    }
    exact_contract(&h->L1, idx, cnt, na, x1, acc64, acc128);   /* feature_layer.w, :143 */
    na = 0;
    for (int f = 0; f < F; ++f) {          /* IAF#1 (:144) */
        float s = iaf_step(&v1[f], x1[f], h->thr, h->vmin);
        if (hid_out) hid_out[f] = (uint8_t)(s > 255 ? 255 : s);
        if (s != 0.0f) {
            if (s > ORACLE_MAX_SPIKE) { h->overflow++; s = ORACLE_MAX_SPIKE; }
            idx[na] = f; cnt[na] = (int)s; ++na;
        }
    }
    exact_contract(&h->L2, idx, cnt, na, x2, acc64, acc128);   /* output_layer.w, :145 */
    for (int p = 0; p < P; ++p)            /* IAF#2 (add_spiking_output=True, :155) */
        s2[p] = iaf_step(&v2[p], x2[p], h->thr, h->vmin);
}

/* R2 + R6: the per-query loop, lens/run_model.py:229-246, fed by
 * lens/src/dataset.py:14-51,118-125 (p = u8/255; spikes = U < p with the SAME
 * U[T, roi*roi] -- torch.manual_seed(50) -- for every image).  `pooled` holds the
 * already sub-sampled pixels [B][Q][I] (R1 commutes with the raster because the
 * raster is elementwise), U the matching columns of the seed-50 matrix.
 * counts[B][Q][P] = spikes.sum(dim=0) per query (:239); state carried across
 * queries and calls (no reset_states() anywhere in the reference).
 * hidden_steps (nullable) [B][Q*T][F] u8, out_steps (nullable) [B][Q*T][P] u8. */
int lens_oracle_snn_forward(lens_oracle_snn *h, const uint8_t *pooled, int B, int Q,
                            float *counts, uint8_t *hidden_steps, uint8_t *out_steps)
{
    if (!h || !h->U || B > h->B) return -1;
    int I = h->I, F = h->F, P = h->P, T = h->T;
    int nmax = (I > F ? I : F);
    float *in = (float *)malloc((size_t)I * sizeof(float));
    float *x1 = (float *)malloc((size_t)F * sizeof(float));
    float *x2 = (float *)malloc((size_t)P * sizeof(float));
    float *s2 = (float *)malloc((size_t)P * sizeof(float));
    float *pr = (float *)malloc((size_t)I * sizeof(float));
    int *idx = (int *)malloc((size_t)nmax * sizeof(int));
    int *cnt = (int *)malloc((size_t)nmax * sizeof(int));
    int nacc = (F > P ? F : P);
    int64_t *acc64 = (int64_t *)malloc((size_t)nacc * sizeof(int64_t));
    __int128 *acc128 = (__int128 *)malloc((size_t)nacc * sizeof(__int128));
    if (!in || !x1 || !x2 || !s2 || !pr || !idx || !cnt || !acc64 || !acc128) return -2;
    for (int b = 0; b < B; ++b) {
        for (int q = 0; q < Q; ++q) {
            const uint8_t *px = pooled + ((size_t)b * Q + q) * I;
            float *c = counts + ((size_t)b * Q + q) * P;
            for (int p = 0; p < P; ++p) c[p] = 0.0f;
            for (int i = 0; i < I; ++i) pr[i] = (float)px[i] / 255.0f;  /* dataset.py:23 */
            for (int t = 0; t < T; ++t) {
                const float *u = h->U + (size_t)t * I;
                for (int i = 0; i < I; ++i) in[i] = (u[i] < pr[i]) ? 1.0f : 0.0f; /* :121 */
                size_t step = (size_t)b * Q * T + (size_t)q * T + t;
                snn_step(h, b, in, x1, x2, idx, cnt, acc64, acc128,
                         hidden_steps ? hidden_steps + step * F : NULL, s2);
                for (int p = 0; p < P; ++p) c[p] += s2[p];              /* run_model.py:239 */
                if (out_steps) {
                    uint8_t *o = out_steps + step * P;
                    for (int p = 0; p < P; ++p) o[p] = (uint8_t)(s2[p] > 255 ? 255 : s2[p]);
                }
            }
        }
    }
    free(in); free(x1); free(x2); free(s2); free(pr); free(idx); free(cnt); free(acc64); free(acc128);
    return 0;
}

/* Operator seam: sinabs_model(x) with x float [B][steps][I] (pooled pixels of
 * an arbitrary float raster, run_model.py:238); spikes_out [B][steps][P] float. */
int lens_oracle_snn_forward_float(lens_oracle_snn *h, const float *x, int B, int steps,
                                  float *spikes_out)
{
    if (!h || B > h->B) return -1;
    int I = h->I, F = h->F, P = h->P;
    int nmax = (I > F ? I : F), nacc = (F > P ? F : P);
    float *x1 = (float *)malloc((size_t)F * sizeof(float));
    float *x2 = (float *)malloc((size_t)P * sizeof(float));
    int *idx = (int *)malloc((size_t)nmax * sizeof(int));
    int *cnt = (int *)malloc((size_t)nmax * sizeof(int));
    int64_t *acc64 = (int64_t *)malloc((size_t)nacc * sizeof(int64_t));
    __int128 *acc128 = (__int128 *)malloc((size_t)nacc * sizeof(__int128));
    if (!x1 || !x2 || !idx || !cnt || !acc64 || !acc128) return -2;
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < steps; ++t)
            snn_step(h, b, x + ((size_t)b * steps + t) * I, x1, x2, idx, cnt, acc64, acc128, NULL,
                     spikes_out + ((size_t)b * steps + t) * P);
    free(x1); free(x2); free(idx); free(cnt); free(acc64); free(acc128);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* R7: sequence matching -- lens/run_model.py:248-254:
 *   D = conv2d(S[1,1,Q,P] as f32, eye(L)) / L, transposed
 *   => D[r][q] = (sum_{j<L} S[q+j][r+j]) / L,  shape [P-L+1][Q-L+1]
 * (rows = database, columns = query).  The L partial sums are integers, hence
 * exact in fp32 in any order; the single division is IEEE fp32 (numpy f32 / int).
 * S: [Q][P] float.  D: [P-L+1][Q-L+1] float.  L >= 1.                          */
int lens_oracle_seqmatch(const float *S, int Q, int P, int L, float *D)
{
    if (L < 1 || L > Q || L > P) return -1;
    int Qo = Q - L + 1, Po = P - L + 1;
    float fl = (float)L;
    for (int r = 0; r < Po; ++r)
        for (int q = 0; q < Qo; ++q) {
            float acc = 0.0f;
            for (int j = 0; j < L; ++j) acc += S[(size_t)(q + j) * P + (r + j)];
            D[(size_t)r * Qo + q] = acc / fl;
        }
    return 0;
}

/* R9 (selection part): the K best database rows of every query column of
 * D[Po][Qo] under the documented deterministic tie rule
 *   order by (value descending, row index descending)
 * which is what np.argsort(D, 0, kind='stable')[-K:][::-1] yields
 * (lens/src/metrics.py:218 uses the default, unstable kind: see DESIGN.md H5).
 * top_idx/top_val: [Qo][K]; entries beyond Po are idx -1 / val -inf.            */
int lens_oracle_topk(const float *D, int Po, int Qo, int K, int32_t *top_idx, float *top_val)
{
    if (K < 1) return -1;
    for (int q = 0; q < Qo; ++q) {
        int32_t *ti = top_idx + (size_t)q * K;
        float *tv = top_val + (size_t)q * K;
        int n = 0;
        for (int r = 0; r < Po; ++r) {
            float v = D[(size_t)r * Qo + q];
            /* insert (v, r): later rows win ties */
            int pos = n;
            while (pos > 0 && tv[pos - 1] <= v) --pos;
            if (pos >= K) continue;
            int last = (n < K) ? n : K - 1;
            for (int m = last; m > pos; --m) { tv[m] = tv[m - 1]; ti[m] = ti[m - 1]; }
            tv[pos] = v; ti[pos] = r;
            if (n < K) ++n;
        }
        for (int m = n; m < K; ++m) { tv[m] = -INFINITY; ti[m] = -1; }
    }
    return 0;
}
