/*
 * Plain-C restatement of the reference's detection-head grid path (CPU oracle).
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Never linked into the
 * product library (multigriddet_b200/csrc); built by oracle/Makefile into
 * oracle/_build/libmgd_oracle.so and used by tests/, smoke() and bench.py's
 * cpu_baseline leg as the fast checker for full-size batches.
 *
 * Parity status: PINNED.  tests/test_oracle_c.py checks this file bit-for-bit
 * against oracle/mgd_oracle.py (itself pinned to the reference sources and to
 * tests/golden/) when NumPy runs with its AVX dispatch disabled, i.e. when
 * np.exp/np.log/np.tanh are glibc's expf/logf/tanhf -- the same libm calls
 * made here.  scipy.special.expit is 1/(1+expf(-x)) in float on every host.
 *
 * Reference lines restated (paths relative to /root/reference):
 *   encode      multigriddet/data/generators.py:2486-2494, 2514-2544, 3393-3473
 *   decode      multigriddet/postprocess/multigrid_decode.py:100-183
 *   letterbox   multigriddet/postprocess/multigrid_decode.py:185-235
 *   threshold   multigriddet/postprocess/multigrid_decode.py:262-278, 336-345
 *   NMS         multigriddet/postprocess/nms.py:83-148 (IoU), 151-231 (DIoU)
 *   xyxy        multigriddet/postprocess/multigrid_decode.py:397-422
 * Third-party arithmetic restated from its published algorithm:
 *   numpy float32 add.reduce = pairwise sum with 8 unrolled accumulators
 *   (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE 128), numpy 2.3.5;
 *   scipy.special.softmax = exp(x - max) / sum (scipy 1.18.1).
 *
 * Compile with -O2 -ffp-contract=off (no FMA contraction: the integer results
 * depend on IEEE single operations in the reference's order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_ANCHORS 64

/* ---- numpy float32 pairwise sum (contiguous) ------------------------------ */
static float np_sum_f32(const float *a, long n)
{
    if (n < 8) {
        float res = 0.f;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        long i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum_f32(a, n2) + np_sum_f32(a + n2, n - n2);
    }
}

static float expit_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

/* ---- encode --------------------------------------------------------------- */

/* generators.py:2486-2494 + np.round(.,3) + stable argsort(-iol)[0] */
static int best_anchor(float bw, float bh, const double *anchors, int K, int anchors_f64)
{
    int best = 0;
    if (!anchors_f64) {
        float top = -INFINITY;
        for (int g = 0; g < K; g++) {
            float aw = (float)anchors[2 * g], ah = (float)anchors[2 * g + 1];
            float iw = fminf(bw, aw), ih = fminf(bh, ah);
            float inter = iw * ih;
            float box_area = bw * bh;
            float anc_area = aw * ah;
            float iol = inter / fmaxf(box_area, anc_area);
            float r = rintf(iol * 1000.0f) / 1000.0f;
            if (r > top) { top = r; best = g; }
        }
    } else {
        double top = -INFINITY;
        for (int g = 0; g < K; g++) {
            double aw = anchors[2 * g], ah = anchors[2 * g + 1];
            double iw = fmin((double)bw, aw), ih = fmin((double)bh, ah);
            double inter = iw * ih;
            double box_area = (double)(float)(bw * bh);       /* f32 product, :2489 */
            double anc_area = aw * ah;
            double iol = inter / fmax(box_area, anc_area);
            double r = rint(iol * 1000.0) / 1000.0;
            if (r > top) { top = r; best = g; }
        }
    }
    return best;
}

/*
 * boxes (B,N,5) f32; anchors double[K*2] layer-major; layer_na[L]; grid_hw[L*2];
 * y_true[L] -> (B, g0, g1, 5+A_l+C) f32, fully overwritten.
 * returns 0, or 1 if any class id >= C (generators.py:3409).
 */
int orc_encode(const float *boxes, int B, int N, const double *anchors,
               const int *layer_na, int L, int anchors_f64, int C,
               int in_h, int in_w, const int *grid_hw, float **y_true,
               long long *stats /* [n_valid, n_skipped] or NULL */)
{
    int K = 0, first_of[16];
    for (int l = 0; l < L; l++) { first_of[l] = K; K += layer_na[l]; }
    for (long i = 0; i < (long)B * N; i++)
        if (!(boxes[i * 5 + 4] < (float)C)) return 1;
    for (int l = 0; l < L; l++) {
        size_t n = (size_t)B * grid_hw[2 * l] * grid_hw[2 * l + 1] * (5 + layer_na[l] + C);
        memset(y_true[l], 0, n * sizeof(float));
    }
    long long n_valid = 0, n_skipped = 0;
    for (int b = 0; b < B; b++) {
        for (int t = 0; t < N; t++) {
            const float *bx = boxes + ((size_t)b * N + t) * 5;
            float sx = bx[0] + bx[2], sy = bx[1] + bx[3];
            float ctr_x = floorf(sx / 2.0f), ctr_y = floorf(sy / 2.0f);   /* :3415 */
            float bw = bx[2] - bx[0], bh = bx[3] - bx[1];                 /* :3416 */
            if (bw * bh <= 0.0f) continue;                                /* :3431 */
            n_valid++;
            int g = best_anchor(bw, bh, anchors, K, anchors_f64);
            int layer = 0;
            while (layer + 1 < L && g >= first_of[layer + 1]) layer++;
            int k = g - first_of[layer];
            int g0 = grid_hw[2 * layer], g1 = grid_hw[2 * layer + 1];
            int D = 5 + layer_na[layer] + C;
            int cls = (int)bx[4];
            double gx = (double)ctr_x * ((double)g0 / (double)in_h);      /* :3438 */
            double gy = (double)ctr_y * ((double)g1 / (double)in_w);      /* :3439 */
            int col = (int)gx, row = (int)gy;                             /* :3441 */
            double fx = gx - col, fy = gy - row;
            float tw, th;
            if (!anchors_f64) {
                float rw = bw / (float)anchors[2 * g], rh = bh / (float)anchors[2 * g + 1];
                tw = (rw < 1e-3f) ? (float)log(1e-3) : logf(rw);          /* :3446-3449 */
                th = (rh < 1e-3f) ? (float)log(1e-3) : logf(rh);
            } else {
                double rw = (double)bw / anchors[2 * g], rh = (double)bh / anchors[2 * g + 1];
                tw = (float)log(rw < 1e-3 ? 1e-3 : rw);
                th = (float)log(rh < 1e-3 ? 1e-3 : rh);
            }
            int written = 0;
            for (int dx = -1; dx <= 1; dx++) {                            /* :3454 */
                int cc = col + dx;
                for (int dy = -1; dy <= 1; dy++) {                        /* :3456 */
                    int rr = row + dy;
                    if (cc < 0 || cc >= g0 || rr < 0 || rr >= g1) continue;
                    float *cell = y_true[layer] + (((size_t)b * g0 + rr) * g1 + cc) * D;
                    if (cell[4] == 1.0f && written >= 3) { n_skipped++; continue; }  /* :3463 */
                    memset(cell, 0, D * sizeof(float));
                    cell[0] = (float)((double)(-dx) + fx);
                    cell[1] = (float)((double)(-dy) + fy);
                    cell[2] = tw;
                    cell[3] = th;
                    cell[4] = 1.0f;
                    cell[5 + k] = 1.0f;
                    cell[5 + layer_na[layer] + cls] = 1.0f;
                    written++;
                }
            }
        }
    }
    if (stats) { stats[0] = n_valid; stats[1] = n_skipped; }
    return 0;
}

/* ---- decode + NMS --------------------------------------------------------- */

typedef struct {
    double x, y, w, h;   /* top-left + size, original-image pixels */
    double score;        /* float32 value in a float64 container */
    int cls;
    int index;           /* flat candidate index (layer-major, row-major) */
} cand_t;

static int cand_order(const void *pa, const void *pb)
{
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    return (a->index > b->index) - (a->index < b->index);
}

/* nms.py:121-148 / 189-231 in float64 */
static double pair_metric(const cand_t *a, const cand_t *b, int use_diou)
{
    double x1 = a->x, y1 = a->y, w1 = a->w, h1 = a->h;
    double x2 = b->x, y2 = b->y, w2 = b->w, h2 = b->h;
    double iw = fmax(0.0, fmin(x1 + w1, x2 + w2) - fmax(x1, x2));
    double ih = fmax(0.0, fmin(y1 + h1, y2 + h2) - fmax(y1, y2));
    double inter = iw * ih;
    double uni = w1 * h1 + w2 * h2 - inter;
    double iou = inter / (uni + 1e-8);
    if (!use_diou) return iou;
    double dxc = (x1 + w1 / 2) - (x2 + w2 / 2), dyc = (y1 + h1 / 2) - (y2 + h2 / 2);
    double dist = dxc * dxc + dyc * dyc;
    double ex = fmax(x1 + w1, x2 + w2) - fmin(x1, x2);
    double ey = fmax(y1 + h1, y2 + h2) - fmin(y1, y2);
    double diag = ex * ex + ey * ey;
    return iou - dist / (diag + 1e-8);
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * preds[L] -> (B, gh, gw, 5+A_l+C) f32.  image_hw (B,2) int32 (h, w).
 * method: 0 = IoU greedy, 1 = DIoU greedy.  Outputs are padded to max_boxes.
 * Also writes, if non-NULL, all_scores (B, cells) f32 and all_cls (B, cells) i32
 * for every cell (dense decode), for value-level parity checks.
 */
int orc_decode_nms(const float *const *preds, int B, int L, const int *grid_hw,
                   const double *anchors, const int *layer_na, int anchors_f64,
                   int C, int in_h, int in_w, const int *image_hw,
                   int use_softmax, int rescore, double confidence,
                   double nms_threshold, int method, int per_class, int max_boxes,
                   double *out_xywh, int *out_xyxy, double *out_scores,
                   int *out_classes, int *out_index, int *out_counts,
                   int *out_ncand, float *all_scores, int *all_cls)
{
    int first_of[16], K = 0;
    long cells = 0, cell_off[16];
    for (int l = 0; l < L; l++) {
        first_of[l] = K; K += layer_na[l];
        cell_off[l] = cells; cells += (long)grid_hw[2 * l] * grid_hw[2 * l + 1];
    }
    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)cells);
    unsigned char *dead = (unsigned char *)malloc((size_t)cells);
    float *ebuf = (float *)malloc(sizeof(float) * (size_t)(C > ORC_MAX_ANCHORS ? C : ORC_MAX_ANCHORS));
    if (!cand || !dead || !ebuf) return 2;
    for (int b = 0; b < B; b++) {
        /* letterbox constants, float32: multigrid_decode.py:205-216 */
        float mh = (float)in_h, mw = (float)in_w;
        float ih = (float)image_hw[2 * b], iw = (float)image_hw[2 * b + 1];
        float ratio = fminf(mh / ih, mw / iw);
        float new_h = rintf(ih * ratio), new_w = rintf(iw * ratio);
        float off_h = (mh - new_h) / 2.0f / mh, off_w = (mw - new_w) / 2.0f / mw;
        float sc_h = mh / new_h, sc_w = mw / new_w;
        int n = 0;
        for (int l = 0; l < L; l++) {
            int gh = grid_hw[2 * l], gw = grid_hw[2 * l + 1], A = layer_na[l];
            int D = 5 + A + C;
            const float *base = preds[l] + (size_t)b * gh * gw * D;
            for (int r = 0; r < gh; r++) for (int c = 0; c < gw; c++) {
                const float *p = base + ((size_t)r * gw + c) * D;
                /* anchor / class probabilities: :140-145 */
                float pa_max, pc_max; int ka = 0, kc = 0;
                for (int pass = 0; pass < 2; pass++) {
                    const float *x = pass ? p + 5 + A : p + 5;
                    int nn = pass ? C : A;
                    float best = -INFINITY; int arg = 0;
                    if (use_softmax) {
                        float m = x[0];
                        for (int i = 1; i < nn; i++) if (x[i] > m) m = x[i];
                        for (int i = 0; i < nn; i++) ebuf[i] = expf(x[i] - m);
                        float s = np_sum_f32(ebuf, nn);
                        for (int i = 0; i < nn; i++) {
                            float q = ebuf[i] / s;
                            if (q > best) { best = q; arg = i; }
                        }
                    } else {
                        for (int i = 0; i < nn; i++) {
                            float q = expit_f32(x[i]);
                            if (q > best) { best = q; arg = i; }
                        }
                    }
                    if (pass) { pc_max = best; kc = arg; } else { pa_max = best; ka = arg; }
                }
                float score = expit_f32(p[4]);                            /* :147 */
                if (rescore) score = (score * pa_max) * pc_max;           /* :170 */
                long flat = cell_off[l] + (long)r * gw + c;
                if (all_scores) all_scores[(size_t)b * cells + flat] = score;
                if (all_cls) all_cls[(size_t)b * cells + flat] = kc;
                if (!((double)score >= confidence)) continue;             /* :271 */
                /* box: :151-163 */
                float ux = 0.15f * p[0], uy = 0.15f * p[1];
                float ax = tanhf(ux) + expit_f32(ux), ay = tanhf(uy) + expit_f32(uy);
                double bx = ((double)ax + (double)c) / (double)gh;        /* :154-155 */
                double by = ((double)ay + (double)r) / (double)gw;
                int g = first_of[l] + ka;
                double bw, bh;
                if (!anchors_f64) {
                    float w32 = (float)anchors[2 * g] * expf(p[2]);
                    float h32 = (float)anchors[2 * g + 1] * expf(p[3]);
                    bw = (double)(float)((double)w32 / (double)in_h);     /* :163 f32 in place */
                    bh = (double)(float)((double)h32 / (double)in_w);
                } else {
                    bw = anchors[2 * g] * (double)expf(p[2]) / (double)in_h;
                    bh = anchors[2 * g + 1] * (double)expf(p[3]) / (double)in_w;
                }
                /* correct_boxes: :219-228 */
                bx = (bx - (double)off_w) * (double)sc_w;
                by = (by - (double)off_h) * (double)sc_h;
                bw = bw * (double)sc_w;
                bh = bh * (double)sc_h;
                bx = bx - bw / 2.0;
                by = by - bh / 2.0;
                bx *= (double)iw; by *= (double)ih;
                bw *= (double)iw; bh *= (double)ih;
                cand[n].x = bx; cand[n].y = by; cand[n].w = bw; cand[n].h = bh;
                cand[n].score = (double)score; cand[n].cls = kc; cand[n].index = (int)flat;
                n++;
            }
        }
        if (out_ncand) out_ncand[b] = n;
        qsort(cand, (size_t)n, sizeof(cand_t), cand_order);
        memset(dead, 0, (size_t)n);
        int kept = 0;
        for (int i = 0; i < n && kept < max_boxes; i++) {
            if (dead[i]) continue;
            size_t o = (size_t)b * max_boxes + kept;
            out_xywh[o * 4 + 0] = cand[i].x; out_xywh[o * 4 + 1] = cand[i].y;
            out_xywh[o * 4 + 2] = cand[i].w; out_xywh[o * 4 + 3] = cand[i].h;
            double W = (double)image_hw[2 * b + 1], H = (double)image_hw[2 * b];
            out_xyxy[o * 4 + 0] = (int)floor(clipd(cand[i].x, 0, W) + 0.5);
            out_xyxy[o * 4 + 1] = (int)floor(clipd(cand[i].y, 0, H) + 0.5);
            out_xyxy[o * 4 + 2] = (int)floor(clipd(cand[i].x + cand[i].w, 0, W) + 0.5);
            out_xyxy[o * 4 + 3] = (int)floor(clipd(cand[i].y + cand[i].h, 0, H) + 0.5);
            out_scores[o] = cand[i].score;
            out_classes[o] = cand[i].cls;
            out_index[o] = cand[i].index;
            kept++;
            for (int j = i + 1; j < n; j++) {
                if (dead[j]) continue;
                if (per_class && cand[j].cls != cand[i].cls) continue;
                if (!(pair_metric(&cand[i], &cand[j], method) < nms_threshold)) dead[j] = 1;
            }
        }
        out_counts[b] = kept;
        for (int q = kept; q < max_boxes; q++) {
            size_t o = (size_t)b * max_boxes + q;
            for (int e = 0; e < 4; e++) { out_xywh[o * 4 + e] = 0.0; out_xyxy[o * 4 + e] = 0; }
            out_scores[o] = 0.0; out_classes[o] = -1; out_index[o] = -1;
        }
    }
    free(cand); free(dead); free(ebuf);
    return 0;
}

/* plain libm values, so tests can compare the product's emulations with them */
float orc_expf(float x) { return expf(x); }
float orc_logf(float x) { return logf(x); }
float orc_tanhf(float x) { return tanhf(x); }
float orc_np_sum_f32(const float *a, long n) { return np_sum_f32(a, n); }
