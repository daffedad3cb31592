/*
 * cl4_oracle.c — CPU restatement of the CL4WSIS pseudo-label hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under cl4wsis_b200/ (the product) imports, links or calls it.
 *
 * Parity status: PINNED.  The reference has no golden vectors of its own
 * (SURVEY.md §4), so the oracle is pinned against outputs of the reference
 * itself, run in the build container by tests/golden/make_golden.py and
 * committed as tests/golden/ (.npz), plus the survey's known answers (§8c ①-⑨).
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout).  Arithmetic is fp32 where the reference's is, with the
 * exact operation order where bit-exactness is required (group_pixels, NMS).
 * Build with -ffp-contract=off: every fused multiply-add below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CL4O_OK 0
#define CL4O_EINVAL (-1)
#define CL4O_ENOMEM (-2)

/* 8-neighbourhood tap order of the 3x3 shift stencils, centre skipped:
 * wss/modules.py:30-40 (LocalAffinity._init_aff), :69-79 (LocalAffinityCopy). */
static const int TAP_DY[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static const int TAP_DX[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
/* 9-tap order of LocalStDev (centre included): wss/modules.py:92-102. */
static const int TAP9_DY[9] = {-1, -1, -1, 0, 0, 0, 1, 1, 1};
static const int TAP9_DX[9] = {-1, 0, 1, -1, 0, 1, -1, 0, 1};

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

int cl4o_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void cl4o_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- *
 * Bilinear resize, align_corners=True — the first line of PAMR.forward
 * (wss/modules.py:134, F.interpolate(mask, size=x.size()[-2:], "bilinear",
 * align_corners=True)).  Identity when sizes match (SURVEY §8c ①).
 * ------------------------------------------------------------------------- */
int cl4o_resize_bilinear_ac(const float* in, float* out, int planes, int h, int w, int H, int W) {
    if (planes < 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return CL4O_EINVAL;
    const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < planes; ++p) {
        const float* src = in + (size_t)p * h * w;
        float* dst = out + (size_t)p * H * W;
        for (int y = 0; y < H; ++y) {
            const float fy = sy * (float)y;
            int y0 = (int)fy;
            if (y0 > h - 1) y0 = h - 1;
            const int y1 = y0 + ((y0 < h - 1) ? 1 : 0);
            const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
            for (int x = 0; x < W; ++x) {
                const float fx = sx * (float)x;
                int x0 = (int)fx;
                if (x0 > w - 1) x0 = w - 1;
                const int x1 = x0 + ((x0 < w - 1) ? 1 : 0);
                const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
                const float top = lx0 * src[y0 * w + x0] + lx1 * src[y0 * w + x1];
                const float bot = lx0 * src[y1 * w + x0] + lx1 * src[y1 * w + x1];
                dst[y * W + x] = ly0 * top + ly1 * bot;
            }
        }
    }
    return CL4O_OK;
}

/* ------------------------------------------------------------------------- *
 * Helper stencils on their own: wss/modules.py:17-119.
 *   mode 0  LocalAffinity.forward      x - shift_p(x)   (:47-62; conv2d with the
 *           +1 centre / -1 neighbour kernel of :26-45 = one rounded subtraction)
 *   mode 1  LocalAffinityAbs.forward   |x - shift_p(x)| (:115-119)
 *   mode 2  LocalAffinityCopy.forward  shift_p(x)       (:65-83)
 * x [planes,H,W] -> out [planes,8*D,H,W], p = dilation_index*8 + tap.
 * ------------------------------------------------------------------------- */
int cl4o_local_affinity(const float* x, float* out, int planes, int H, int W, const int* dil, int D, int mode) {
    if (planes < 0 || H <= 0 || W <= 0 || D <= 0 || mode < 0 || mode > 2) return CL4O_EINVAL;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int pl = 0; pl < planes; ++pl) {
        const float* src = x + (size_t)pl * HW;
        for (int di = 0; di < D; ++di)
            for (int t = 0; t < 8; ++t) {
                float* dst = out + ((size_t)pl * 8 * D + di * 8 + t) * HW;
                for (int y = 0; y < H; ++y) {
                    const int yy = clampi(y + TAP_DY[t] * dil[di], 0, H - 1);
                    for (int xx = 0; xx < W; ++xx) {
                        const float nb = src[(size_t)yy * W + clampi(xx + TAP_DX[t] * dil[di], 0, W - 1)];
                        const float c = src[(size_t)y * W + xx];
                        dst[(size_t)y * W + xx] = mode == 2 ? nb : (mode == 1 ? fabsf(c - nb) : c - nb);
                    }
                }
            }
    }
    return CL4O_OK;
}

/* LocalStDev.forward: wss/modules.py:86-112 — x.std(2) (unbiased) over the 9*D
 * gathered samples, the centre once per dilation.  x [planes,H,W] -> out [planes,H,W]. */
int cl4o_local_stdev(const float* x, float* out, int planes, int H, int W, const int* dil, int D) {
    if (planes < 0 || H <= 0 || W <= 0 || D <= 0) return CL4O_EINVAL;
    const size_t HW = (size_t)H * W;
    const int N = 9 * D;
#pragma omp parallel for schedule(static)
    for (int pl = 0; pl < planes; ++pl) {
        const float* src = x + (size_t)pl * HW;
        for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
                double s = 0.0;
                for (int di = 0; di < D; ++di)
                    for (int t = 0; t < 9; ++t)
                        s += (double)src[(size_t)clampi(y + TAP9_DY[t] * dil[di], 0, H - 1) * W +
                                         clampi(xx + TAP9_DX[t] * dil[di], 0, W - 1)];
                const double mean = s / N;
                double ss = 0.0;
                for (int di = 0; di < D; ++di)
                    for (int t = 0; t < 9; ++t) {
                        const double dlt = (double)src[(size_t)clampi(y + TAP9_DY[t] * dil[di], 0, H - 1) * W +
                                                       clampi(xx + TAP9_DX[t] * dil[di], 0, W - 1)] - mean;
                        ss += dlt * dlt;
                    }
                out[(size_t)pl * HW + (size_t)y * W + xx] = (float)sqrt(ss / (double)(N - 1));
            }
    }
    return CL4O_OK;
}

/* ------------------------------------------------------------------------- *
 * PAMR affinity weights: wss/modules.py:141-145.
 *   x_std = LocalStDev(x)           unbiased std over the 9*D samples (:86-112)
 *   a     = |x - shift_p(x)|        LocalAffinityAbs (:115-119 over :47-62)
 *   l_p   = mean_k( -a_kp / (1e-8 + 0.1*std_k) )      (:143-144)
 *   w_p   = softmax_p(l_p)                             (:145)
 * replicate padding == clamped source coordinates (:57-58).
 * Output layout: w[b][p][y][x], p = dilation_index*8 + tap.
 * ------------------------------------------------------------------------- */
int cl4o_pamr_weights(const float* img, float* w, int B, int K, int H, int W, const int* dil, int D) {
    if (B < 0 || K <= 0 || H <= 0 || W <= 0 || D <= 0 || D > 16) return CL4O_EINVAL;
    for (int i = 0; i < D; ++i)
        if (dil[i] <= 0) return CL4O_EINVAL;
    const int P = 8 * D, N = 9 * D;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int y = 0; y < H; ++y) {
            float logit[8 * 16];
            for (int x = 0; x < W; ++x) {
                for (int p = 0; p < P; ++p) logit[p] = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float* pl = img + ((size_t)b * K + k) * HW;
                    /* torch.std on CPU accumulates in double (Welford); a two-pass
                     * double computation agrees to the last float bit in practice. */
                    double s = 0.0;
                    for (int di = 0; di < D; ++di)
                        for (int t = 0; t < 9; ++t) {
                            const int yy = clampi(y + TAP9_DY[t] * dil[di], 0, H - 1);
                            const int xx = clampi(x + TAP9_DX[t] * dil[di], 0, W - 1);
                            s += (double)pl[(size_t)yy * W + xx];
                        }
                    const double mean = s / N;
                    double ss = 0.0;
                    for (int di = 0; di < D; ++di)
                        for (int t = 0; t < 9; ++t) {
                            const int yy = clampi(y + TAP9_DY[t] * dil[di], 0, H - 1);
                            const int xx = clampi(x + TAP9_DX[t] * dil[di], 0, W - 1);
                            const double dlt = (double)pl[(size_t)yy * W + xx] - mean;
                            ss += dlt * dlt;
                        }
                    const float sd = (float)sqrt(ss / (double)(N - 1));
                    const float den = 1e-8f + 0.1f * sd;
                    const float c = pl[(size_t)y * W + x];
                    for (int di = 0; di < D; ++di)
                        for (int t = 0; t < 8; ++t) {
                            const int yy = clampi(y + TAP_DY[t] * dil[di], 0, H - 1);
                            const int xx = clampi(x + TAP_DX[t] * dil[di], 0, W - 1);
                            const float a = fabsf(c - pl[(size_t)yy * W + xx]);
                            logit[di * 8 + t] += (-a) / den;
                        }
                }
                float mx = -INFINITY;
                for (int p = 0; p < P; ++p) {
                    logit[p] = logit[p] / (float)K;
                    if (logit[p] > mx) mx = logit[p];
                }
                float z = 0.f;
                for (int p = 0; p < P; ++p) {
                    logit[p] = expf(logit[p] - mx);
                    z += logit[p];
                }
                for (int p = 0; p < P; ++p)
                    w[((size_t)b * P + p) * HW + (size_t)y * W + x] = logit[p] / z;
            }
        }
    }
    return CL4O_OK;
}

/* One propagation sweep: wss/modules.py:147-149,
 *   m = aff_m(mask)  (gather of the 8*D clamped neighbours, :65-83 over :47-62)
 *   mask = (m * w).sum(2)
 */
static void pamr_sweep(const float* w, const float* min_, float* mout, int B, int C, int H, int W,
                       const int* dil, int D) {
    const int P = 8 * D;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int y = 0; y < H; ++y) {
            for (int c = 0; c < C; ++c) {
                const float* src = min_ + ((size_t)b * C + c) * HW;
                float* dst = mout + ((size_t)b * C + c) * HW + (size_t)y * W;
                for (int x = 0; x < W; ++x) dst[x] = 0.f;
                for (int di = 0; di < D; ++di) {
                    const int d = dil[di];
                    for (int t = 0; t < 8; ++t) {
                        const int yy = clampi(y + TAP_DY[t] * d, 0, H - 1);
                        const float* srow = src + (size_t)yy * W;
                        const float* wrow = w + ((size_t)b * P + di * 8 + t) * HW + (size_t)y * W;
                        const int sx = TAP_DX[t] * d;
                        for (int x = 0; x < W; ++x) {
                            const int xx = clampi(x + sx, 0, W - 1);
                            dst[x] += wrow[x] * srow[xx];
                        }
                    }
                }
            }
        }
    }
}

/* PAMR.forward(x, mask): wss/modules.py:133-152.  mask is [B,C,h,w]; it is
 * resized to the image's HxW first (:134).  out is [B,C,H,W]. */
int cl4o_pamr(const float* img, const float* mask, float* out, int B, int K, int C, int H, int W, int h,
              int w_, const int* dil, int D, int num_iter) {
    if (B < 0 || C <= 0 || num_iter < 0) return CL4O_EINVAL;
    const size_t HW = (size_t)H * W;
    const size_t nm = (size_t)B * C * HW;
    const int P = 8 * D;
    float* wts = (float*)malloc(sizeof(float) * (size_t)B * P * HW + 16);
    float* a = (float*)malloc(sizeof(float) * nm + 16);
    float* bbuf = (float*)malloc(sizeof(float) * nm + 16);
    if (!wts || !a || !bbuf) {
        free(wts); free(a); free(bbuf);
        return CL4O_ENOMEM;
    }
    int rc = cl4o_resize_bilinear_ac(mask, a, B * C, h, w_, H, W);
    if (rc == CL4O_OK) rc = cl4o_pamr_weights(img, wts, B, K, H, W, dil, D);
    if (rc == CL4O_OK) {
        float *cur = a, *nxt = bbuf;
        for (int it = 0; it < num_iter; ++it) {
            pamr_sweep(wts, cur, nxt, B, C, H, W, dil, D);
            float* t = cur; cur = nxt; nxt = t;
        }
        memcpy(out, cur, sizeof(float) * nm);
    }
    free(wts); free(a); free(bbuf);
    return rc;
}

/* ------------------------------------------------------------------------- *
 * k x k stride-1 max pooling with implicit -inf padding, NaN-propagating like
 * ATen's max_pool2d (used by wss/utils.py:8-9 and modules/utils.py:483-484).
 * Separable: rows then columns.
 * ------------------------------------------------------------------------- */
static inline float nanmax(float a, float b) { return (b > a || isnan(b)) ? b : a; }

static int maxpool_plane(const float* in, float* out, float* tmp, int H, int W, int k) {
    const int r = (k - 1) / 2;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float m = -INFINITY;
            const int x0 = x - r < 0 ? 0 : x - r, x1 = x + r > W - 1 ? W - 1 : x + r;
            for (int xx = x0; xx <= x1; ++xx) m = nanmax(m, in[(size_t)y * W + xx]);
            tmp[(size_t)y * W + x] = m;
        }
    for (int y = 0; y < H; ++y) {
        const int y0 = y - r < 0 ? 0 : y - r, y1 = y + r > H - 1 ? H - 1 : y + r;
        for (int x = 0; x < W; ++x) {
            float m = -INFINITY;
            for (int yy = y0; yy <= y1; ++yy) m = nanmax(m, tmp[(size_t)yy * W + x]);
            out[(size_t)y * W + x] = m;
        }
    }
    return CL4O_OK;
}

/* Total order used for top-k: larger score first, NaN above everything (as
 * torch.topk), ties broken by the LOWER flat index.  The reference leaves the
 * tie order unspecified; this is the documented choice. */
static inline int key_better(float sa, int ia, float sb, int ib) {
    const int na = isnan(sa), nb = isnan(sb);
    if (na || nb) {
        if (na && nb) return ia < ib;
        return na;
    }
    if (sa != sb) return sa > sb;
    return ia < ib;
}

/* peak_extract(heat, kernel, K): wss/utils.py:3-25.
 *   hmax = max_pool2d(heat, k, 1, (k-1)//2); keep = (hmax == heat)
 *   peak = heat * keep; topk over H*W per (b,c), sorted
 *   ys = int(float(idx) / W), xs = idx % W
 * outputs [B,C,K]: scores f32, ys i32, xs i32. */
int cl4o_peak_extract(const float* heat, float* scores, int* ys, int* xs, int B, int C, int H, int W,
                      int kernel, int K) {
    if (B < 0 || C < 0 || H <= 0 || W <= 0 || kernel <= 0 || (kernel & 1) == 0) return CL4O_EINVAL;
    if (K <= 0 || (size_t)K > (size_t)H * W) return CL4O_EINVAL;
    const size_t HW = (size_t)H * W;
    int rc = CL4O_OK;
#pragma omp parallel for schedule(dynamic)
    for (int pl = 0; pl < B * C; ++pl) {
        float* pooled = (float*)malloc(sizeof(float) * HW);
        float* tmp = (float*)malloc(sizeof(float) * HW);
        float* hs = (float*)malloc(sizeof(float) * K);
        int* hi = (int*)malloc(sizeof(int) * K);
        if (!pooled || !tmp || !hs || !hi) {
            rc = CL4O_ENOMEM;
            free(pooled); free(tmp); free(hs); free(hi);
            continue;
        }
        const float* in = heat + (size_t)pl * HW;
        maxpool_plane(in, pooled, tmp, H, W, kernel);
        int n = 0; /* hs/hi: the K best so far, kept sorted best-first */
        for (size_t i = 0; i < HW; ++i) {
            const float keep = (pooled[i] == in[i]) ? 1.f : 0.f;
            const float s = in[i] * keep;
            if (n == K && !key_better(s, (int)i, hs[K - 1], hi[K - 1])) continue;
            int pos = (n < K) ? n : K - 1;
            while (pos > 0 && key_better(s, (int)i, hs[pos - 1], hi[pos - 1])) {
                hs[pos] = hs[pos - 1];
                hi[pos] = hi[pos - 1];
                --pos;
            }
            hs[pos] = s;
            hi[pos] = (int)i;
            if (n < K) ++n;
        }
        for (int j = 0; j < K; ++j) {
            scores[(size_t)pl * K + j] = hs[j];
            ys[(size_t)pl * K + j] = (int)((float)hi[j] / (float)W);
            xs[(size_t)pl * K + j] = hi[j] % W;
        }
        free(pooled); free(tmp); free(hs); free(hi);
    }
    return rc;
}

/* find_instance_center, top_k=None branch: modules/utils.py:476-495.
 *   t = F.threshold(x, thr, -1)     keep x where x > thr, else -1   (:480)
 *   p = max_pool2d(t, k, 1, (k-1)//2)                               (:483-484)
 *   t[t != p] = -1; centres = nonzero(t > 0) in row-major order     (:485,:492)
 * Writes up to max_out (y,x) int64 pairs; returns the TOTAL count (>= 0) or a
 * negative error. */
long long cl4o_center_nms(const float* heat, float thr, int kernel, int H, int W, long long* ctr,
                          long long max_out) {
    if (H <= 0 || W <= 0 || kernel <= 0 || (kernel & 1) == 0) return CL4O_EINVAL;
    const size_t HW = (size_t)H * W;
    float* t = (float*)malloc(sizeof(float) * HW);
    float* p = (float*)malloc(sizeof(float) * HW);
    float* tmp = (float*)malloc(sizeof(float) * HW);
    if (!t || !p || !tmp) {
        free(t); free(p); free(tmp);
        return CL4O_ENOMEM;
    }
    for (size_t i = 0; i < HW; ++i) t[i] = (heat[i] <= thr) ? -1.f : heat[i];
    maxpool_plane(t, p, tmp, H, W, kernel);
    long long n = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const float v = (t[i] != p[i]) ? -1.f : t[i];
            if (v > 0.f) {
                if (n < max_out) {
                    ctr[2 * n] = y;
                    ctr[2 * n + 1] = x;
                }
                ++n;
            }
        }
    free(t); free(p); free(tmp);
    return n;
}

/* group_pixels(ctr, offsets): modules/utils.py:505-542, with the optional
 * (fg * ins_seg).long() of get_instance_segmentation (:606) folded in when fg
 * is non-NULL.
 *   loc = (y, x) + offsets[:, y, x]                    fp32 add      (:527)
 *   dist_k = torch.norm(ctr_k - loc, dim=-1)                        (:536)
 *          = sqrt_rn(fma_rn(dx, dx, rn(dy*dy)))  (ATen CPU 2-vector norm,
 *            0 mismatches / 2M pairs; re-pinned by tests)
 *   id = 1 + argmin_k dist_k, first minimum wins                    (:540)
 */
int cl4o_group_pixels(const long long* ctr, int Kc, const float* offsets, const unsigned char* fg,
                      long long* ids, int H, int W) {
    if (Kc <= 0 || H <= 0 || W <= 0) return CL4O_EINVAL;
    const size_t HW = (size_t)H * W;
    float* cy = (float*)malloc(sizeof(float) * Kc);
    float* cx = (float*)malloc(sizeof(float) * Kc);
    if (!cy || !cx) {
        free(cy); free(cx);
        return CL4O_ENOMEM;
    }
    for (int k = 0; k < Kc; ++k) {
        cy[k] = (float)ctr[2 * k];
        cx[k] = (float)ctr[2 * k + 1];
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const float ly = (float)y + offsets[i];
            const float lx = (float)x + offsets[HW + i];
            float best = 0.f;
            int bk = 0;
            for (int k = 0; k < Kc; ++k) {
                const float dy = cy[k] - ly, dx = cx[k] - lx;
                const float d = sqrtf(fmaf(dx, dx, dy * dy));
                /* ATen argmin: first index wins ties; a NaN is "smaller" than
                 * everything and the first NaN wins. */
                if (k == 0) { best = d; bk = 0; }
                else if (!isnan(best) && (d < best || isnan(d))) { best = d; bk = k; }
            }
            long long id = (long long)bk + 1;
            if (fg) id = fg[i] ? id : 0;
            ids[i] = id;
        }
    free(cy); free(cx);
    return CL4O_OK;
}
