/*
 * jmme_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the mounted reference (/root/reference) holds only README.md:1-4 — no JM
 * sources, tests or golden vectors exist to pin this restatement against.
 * What this file restates is therefore
 *   [STD] the normative H.264 arithmetic (6-tap / bilinear luma interpolation §8.4.2.2.1,
 *         se(v)/ue(v) code lengths), and
 *   [MEM] the JM lencod encoder conventions recalled in SURVEY.md Appendix A (spiral scan,
 *         strict-< tie-break, (0,0) pre-test, 16x16 bonus, Q16 lambda, SATD rounding),
 *         frozen in DESIGN.md §2.
 * Each function names the JM function (Gen A ‖ Gen B name) it follows;
 * none of those files is present under /root/reference, so no file:line can be given.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libjmme_cuda.so) never links or calls it.
 *
 * Single-threaded, scalar, written for clarity.  Implements include/jmme.h.
 */
#include "jmme.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_MVD 4096                 /* |cand - pred| in quarter-pel units stays below this */
#define MAX_PRED JMME_MAX_PRED_QPEL   /* |pred| limit, quarter-pel                            */

static int g_threads = 1;            /* worker threads, see jmme_oracle_set_threads */

/* ---- block geometry (JM blc_size) -------------------------------------------- */
static const int blc_w[8] = {0, 16, 16, 8, 8, 8, 4, 4};
static const int blc_h[8] = {0, 16, 8, 16, 8, 4, 8, 4};
static const int blk_base[8] = {0, 0, 1, 3, 5, 9, 17, 25};

struct jmme_ctx {
    jmme_params p;
    int w16, h16, mb_w, mb_h, pad, pstride, pheight, lambda_factor;
    int n_planes;                    /* 16 with subpel, else 1 */
    int metric[3];                   /* JMME_DIST_* of the integer, half-pel and quarter-pel stage */
    int lf[3];                       /* lambda factor of each stage in the context's cost domain (SSE: lambda^2) */
    int max_pred, cmax;              /* |pred| limit of this context; limit of the window centre (samples)  */
    int cpad, cstride, cheight;      /* chroma_me: padded integer chroma planes (w16/2 + 2 cpad) x (h16/2 + 2 cpad) */
    uint8_t *cplanes[JMME_MAX_REFS][2];
    int cref_set[JMME_MAX_REFS];
    uint8_t *cur_c[2];               /* current chroma, padded to w16/2 x h16/2 by replication */
    int cur_c_set;
    uint8_t *planes[JMME_MAX_REFS];  /* [n_planes] padded planes, contiguous */
    uint8_t *planes_l1;              /* bi-pred: the list-1 picture, same layout */
    int ref_set[JMME_MAX_REFS];
    int32_t *mvbits;                 /* centred table, index v + MAX_MVD */
    int16_t *spx, *spy;              /* spiral */
    int ncand;
    int16_t *med_pred;               /* JMME_PRED_MEDIAN: predictors [ref][mb][41][2] of the last search */
    int16_t *fmv;                    /* ... and the field committed so far: [4*mb_h][4*mb_w][2] */
    int8_t *fref;                    /*                                     [4*mb_h][4*mb_w]    */
    char err[256];
};

/* ---- (a1) Init_Motion_Search_Module ‖ InitializeMotionSearch ----------------------------- */
static int se_bits(int v)            /* length of the signed Exp-Golomb code of v [STD 9.1] */
{
    int a = v < 0 ? -v : v, k = 0;
    if (a == 0) return 1;
    while ((1 << (k + 1)) <= a) k++;        /* k = floor(log2 |v|) */
    return 2 * k + 3;
}
static int ue_bits(int r)            /* length of the unsigned Exp-Golomb code of r */
{
    int k = 0;
    while ((1 << (k + 1)) <= r + 1) k++;
    return 2 * k + 1;
}
static void build_spiral(int R, int16_t *x, int16_t *y)     /* SURVEY A.5 */
{
    int k = 0, l, i;
    x[k] = 0; y[k] = 0; k++;
    for (l = 1; l <= R; l++) {
        for (i = -l + 1; i < l; i++) {
            x[k] = (int16_t)i;  y[k] = (int16_t)-l; k++;
            x[k] = (int16_t)i;  y[k] = (int16_t)l;  k++;
        }
        for (i = -l; i <= l; i++) {
            x[k] = (int16_t)-l; y[k] = (int16_t)i;  k++;
            x[k] = (int16_t)l;  y[k] = (int16_t)i;  k++;
        }
    }
}

int jmme_InitMotionSearchModule(int R, int max_mvd, int32_t *mvbits, int n_refbits,
                                int32_t *refbits, int16_t *sx, int16_t *sy)
{
    int v;
    if (R < 0 || max_mvd < 0 || n_refbits < 0) return JMME_ERR_PARAM;
    if (mvbits) for (v = -max_mvd; v <= max_mvd; v++) mvbits[v + max_mvd] = se_bits(v);
    if (refbits) for (v = 0; v < n_refbits; v++) refbits[v] = ue_bits(v);
    if (sx && sy) build_spiral(R, sx, sy);
    return JMME_OK;
}

/* ---- (a2) LAMBDA_FACTOR / WEIGHTED_COST / MV_COST / REF_COST ----------------------------- */
static const int QP2QUANT[40] = {1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, 7, 8, 9,
                                 10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51, 57,
                                 64, 72, 81, 91};
static double lambda_motion(int qp, int rdopt)
{
    int q = qp - 12;
    if (q < 0) q = 0;
    if (q > 39) q = 39;
    if (rdopt) return sqrt(0.85 * pow(2.0, (double)q / 3.0));
    return (double)QP2QUANT[q];
}
int jmme_lambda_factor(int qp, int rdopt) { return (int)(65536.0 * lambda_motion(qp, rdopt) + 0.5); }

static int weighted_cost(int f, int bits) { return (int)(((int64_t)f * bits) >> 16); }

/* Cost domains.  0 = JM <= 10: J = D + ((lambda_factor * bits) >> 16), lambda_factor Q16.
 * 1 = JM >= 12 with JCOST_CALC_SCALEUP: LAMBDA_ACCURACY_BITS = 5, lambda_factor = (int)(32 lambda + 0.5),
 * J = (D << 5) + lambda_factor * bits — the rate is not truncated, so ties fall differently. */
#define SCALEUP_BITS 5
static int wcost(int domain, int f, int bits) { return domain ? f * bits : weighted_cost(f, bits); }
static int dscale(int domain, int d) { return domain ? d << SCALEUP_BITS : d; }
/* lambda factor of a stage from lambda_motion; an SSE stage works with lambda^2 (JM: lambda_me = lambda_md for
 * SSE, its square root otherwise) */
static int stage_lambda_factor(int domain, double lambda, int metric)
{
    if (metric == JMME_DIST_SSE) lambda *= lambda;
    return (int)((domain ? 32.0 : 65536.0) * lambda + 0.5);
}

/* MV_COST(f,s,cx,cy,px,py) = WEIGHTED_COST(f, mvbits[(cx<<s)-px] + mvbits[(cy<<s)-py]) is written
 * out at its call sites below (s = 2 for integer candidates, 0 for sub-pel ones). */
/* reference rate, charged with the lambda factor of the last stage that ran */
static int ref_cost(const jmme_ctx *c, int ref)
{
    const int f = c->lf[c->p.subpel ? 2 : 0], dom = c->p.cost_domain;
    if (c->p.rdopt) return wcost(dom, f, ue_bits(ref));
    /* (int)(2*lambda*min(ref,1)); lambda is an integer when !rdopt */
    return ref ? (dom ? 2 * f : (int)((2 * (int64_t)f) >> 16)) : 0;
}

/* ---- (a12) UnifiedOneForthPix ‖ getSubImagesLuma [STD 8.4.2.2.1] ------------------------- */
static int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Builds n_planes (1 or 16) planes of (w+2pad) x (h+2pad); plane index yfrac*4+xfrac.
 * Integer samples are edge-replicated without bound (a9: UMV clamp). */
static int build_planes(const uint8_t *luma, int w_in, int h_in, int stride, int w, int h, int pad,
                        int n_planes, uint8_t *out)
{
    int ps = w + 2 * pad, ph = h + 2 * pad, x, y;
    size_t psz = (size_t)ps * ph;
    /* working arrays carry a 3-sample apron so that every tap is in range */
    int ap = 3, as = ps + 2 * ap, ah = ph + 2 * ap;
    int *G = (int *)malloc(sizeof(int) * (size_t)as * ah);   /* integer samples            */
    int *B1 = NULL, *H1 = NULL, *J1 = NULL;
    uint8_t *Pb, *Ph, *Pj;
    if (!G) return JMME_ERR_NOMEM;
#define AT(A, xx, yy) A[(size_t)((yy) + ap) * as + ((xx) + ap)]
    for (y = -ap; y < ph + ap; y++)
        for (x = -ap; x < ps + ap; x++) {
            int sx = clampi(x - pad, 0, w_in - 1), sy = clampi(y - pad, 0, h_in - 1);
            AT(G, x, y) = luma[(size_t)sy * stride + sx];
        }
    for (y = 0; y < ph; y++)
        for (x = 0; x < ps; x++) out[(size_t)y * ps + x] = (uint8_t)AT(G, x, y);
    if (n_planes == 1) { free(G); return JMME_OK; }

    B1 = (int *)malloc(sizeof(int) * (size_t)as * ah);       /* unrounded horizontal half  */
    H1 = (int *)malloc(sizeof(int) * (size_t)as * ah);       /* unrounded vertical half    */
    J1 = (int *)malloc(sizeof(int) * (size_t)as * ah);       /* unrounded centre           */
    if (!B1 || !H1 || !J1) { free(G); free(B1); free(H1); free(J1); return JMME_ERR_NOMEM; }
    /* samples beyond the apron are replicas of the border column/row, so clamping the tap
     * coordinate to the apron is exact */
#define GC(xx, yy) AT(G, clampi(xx, -ap, ps + ap - 1), clampi(yy, -ap, ph + ap - 1))
#pragma omp parallel for private(x) num_threads(g_threads)
    for (y = -ap; y < ph + ap; y++)
        for (x = -ap; x < ps + ap; x++) {
            AT(B1, x, y) = GC(x - 2, y) - 5 * GC(x - 1, y) + 20 * GC(x, y) + 20 * GC(x + 1, y)
                           - 5 * GC(x + 2, y) + GC(x + 3, y);
            AT(H1, x, y) = GC(x, y - 2) - 5 * GC(x, y - 1) + 20 * GC(x, y) + 20 * GC(x, y + 1)
                           - 5 * GC(x, y + 2) + GC(x, y + 3);
        }
#define B1C(xx, yy) AT(B1, xx, clampi(yy, -ap, ph + ap - 1))
#pragma omp parallel for private(x) num_threads(g_threads)
    for (y = 0; y < ph + 1; y++)
        for (x = 0; x < ps + 1; x++)
            AT(J1, x, y) = B1C(x, y - 2) - 5 * B1C(x, y - 1) + 20 * B1C(x, y) + 20 * B1C(x, y + 1)
                           - 5 * B1C(x, y + 2) + B1C(x, y + 3);
#define PL(xf, yf) (out + psz * (size_t)((yf) * 4 + (xf)))
#define Gs(xx, yy) AT(G, xx, yy)
#define bs(xx, yy) clip255((AT(B1, xx, yy) + 16) >> 5)
#define hs(xx, yy) clip255((AT(H1, xx, yy) + 16) >> 5)
#define js(xx, yy) clip255((AT(J1, xx, yy) + 512) >> 10)
    Pb = PL(2, 0); Ph = PL(0, 2); Pj = PL(2, 2);
#pragma omp parallel for private(x) num_threads(g_threads)
    for (y = 0; y < ph; y++)
        for (x = 0; x < ps; x++) {
            size_t o = (size_t)y * ps + x;
            int g = Gs(x, y), b = bs(x, y), h = hs(x, y), j = js(x, y);
            int gr = Gs(x + 1, y), gd = Gs(x, y + 1);       /* H and M of the standard       */
            int m = hs(x + 1, y), s = bs(x, y + 1);
            Pb[o] = (uint8_t)b; Ph[o] = (uint8_t)h; Pj[o] = (uint8_t)j;
            PL(1, 0)[o] = (uint8_t)((g + b + 1) >> 1);      /* a */
            PL(3, 0)[o] = (uint8_t)((gr + b + 1) >> 1);     /* c */
            PL(0, 1)[o] = (uint8_t)((g + h + 1) >> 1);      /* d */
            PL(0, 3)[o] = (uint8_t)((gd + h + 1) >> 1);     /* n */
            PL(2, 1)[o] = (uint8_t)((b + j + 1) >> 1);      /* f */
            PL(2, 3)[o] = (uint8_t)((j + s + 1) >> 1);      /* q */
            PL(1, 2)[o] = (uint8_t)((h + j + 1) >> 1);      /* i */
            PL(3, 2)[o] = (uint8_t)((j + m + 1) >> 1);      /* k */
            PL(1, 1)[o] = (uint8_t)((b + h + 1) >> 1);      /* e */
            PL(3, 1)[o] = (uint8_t)((b + m + 1) >> 1);      /* g */
            PL(1, 3)[o] = (uint8_t)((h + s + 1) >> 1);      /* p */
            PL(3, 3)[o] = (uint8_t)((m + s + 1) >> 1);      /* r */
        }
    free(G); free(B1); free(H1); free(J1);
    return JMME_OK;
}

int jmme_getSubImagesLuma(const uint8_t *luma, int width, int height, int stride, int pad,
                          uint8_t *out_planes)
{
    if (!luma || !out_planes || width <= 0 || height <= 0 || (width & 15) || (height & 15) ||
        pad < 0 || stride < width)
        return JMME_ERR_PARAM;
    return build_planes(luma, width, height, stride, width, height, pad, 16, out_planes);
}

/* ---- (a11) SATD ‖ HadamardSAD4x4 --------------------------------------------------------- */
static int satd4x4(const int *d, int satd_round)
{
    int m[16], t[16], i, s = 0;
    for (i = 0; i < 4; i++) {                       /* rows */
        int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
        int s0 = a + e, s1 = b + c, d0 = a - e, d1 = b - c;
        t[4 * i] = s0 + s1; t[4 * i + 1] = d0 + d1; t[4 * i + 2] = s0 - s1; t[4 * i + 3] = d0 - d1;
    }
    for (i = 0; i < 4; i++) {                       /* columns */
        int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
        int s0 = a + e, s1 = b + c, d0 = a - e, d1 = b - c;
        m[i] = s0 + s1; m[4 + i] = d0 + d1; m[8 + i] = s0 - s1; m[12 + i] = d0 - d1;
    }
    for (i = 0; i < 16; i++) s += m[i] < 0 ? -m[i] : m[i];
    return satd_round ? (s + 1) >> 1 : s >> 1;
}
int jmme_SATD(const int16_t *diff, int n, int satd_round, int32_t *out)
{
    int i, k, d[16];
    if (!diff || !out || n < 0) return JMME_ERR_PARAM;
    for (i = 0; i < n; i++) {
        for (k = 0; k < 16; k++) d[k] = diff[16 * i + k];
        out[i] = satd4x4(d, satd_round);
    }
    return JMME_OK;
}

/* ---- (f2) HadamardSAD8x8: sum |H8 D H8'| over an 8x8 difference block, (s + 2) >> 2 ------------------- */
static int satd8x8(const int *d, int satd_round)
{
    int m[64], i, j, s = 0;
    for (i = 0; i < 64; i++) m[i] = d[i];
    for (j = 0; j < 2; j++) {                       /* rows, then columns (after the transpose below) */
        int t[64];
        for (i = 0; i < 8; i++) {
            const int *r = m + 8 * i;
            int a0 = r[0] + r[4], a1 = r[1] + r[5], a2 = r[2] + r[6], a3 = r[3] + r[7];
            int a4 = r[0] - r[4], a5 = r[1] - r[5], a6 = r[2] - r[6], a7 = r[3] - r[7];
            int b0 = a0 + a2, b1 = a1 + a3, b2 = a0 - a2, b3 = a1 - a3;
            int b4 = a4 + a6, b5 = a5 + a7, b6 = a4 - a6, b7 = a5 - a7;
            /* transposed store: the second pass transforms the columns */
            t[0 * 8 + i] = b0 + b1; t[1 * 8 + i] = b0 - b1; t[2 * 8 + i] = b2 + b3; t[3 * 8 + i] = b2 - b3;
            t[4 * 8 + i] = b4 + b5; t[5 * 8 + i] = b4 - b5; t[6 * 8 + i] = b6 + b7; t[7 * 8 + i] = b6 - b7;
        }
        memcpy(m, t, sizeof m);
    }
    for (i = 0; i < 64; i++) s += m[i] < 0 ? -m[i] : m[i];
    return satd_round ? (s + 2) >> 2 : s >> 2;
}
int jmme_HadamardSAD8x8(const int16_t *diff, int n, int satd_round, int32_t *out)
{
    int i, k, d[64];
    if (!diff || !out || n < 0) return JMME_ERR_PARAM;
    for (i = 0; i < n; i++) {
        for (k = 0; k < 64; k++) d[k] = diff[64 * i + k];
        out[i] = satd8x8(d, satd_round);
    }
    return JMME_OK;
}

/* distortion of a bw x bh block of differences cur - ref under `metric` (JMME_DIST_*); t8: 8x8 transform tiles
 * (both dimensions >= 8), else 4x4 tiles; blocks smaller than 4x4 in a dimension (chroma of the small luma
 * partitions) fall back to SAD under the Hadamard metric (frozen choice) */
static int block_dist(const uint8_t *cur, int cs, const uint8_t *ref, int rs, int bw, int bh, int metric, int t8,
                      int satd_round)
{
    int x, y, x0, y0, s = 0, d[64];
    if (metric == JMME_DIST_HADAMARD && (bw < 4 || bh < 4)) metric = JMME_DIST_SAD;
    if (metric != JMME_DIST_HADAMARD) {
        for (y = 0; y < bh; y++)
            for (x = 0; x < bw; x++) {
                const int e = (int)cur[y * cs + x] - (int)ref[y * rs + x];
                s += metric == JMME_DIST_SSE ? e * e : (e < 0 ? -e : e);
            }
        return s;
    }
    if (t8 && bw >= 8 && bh >= 8) {
        for (y0 = 0; y0 < bh; y0 += 8)
            for (x0 = 0; x0 < bw; x0 += 8) {
                for (y = 0; y < 8; y++)
                    for (x = 0; x < 8; x++) d[8 * y + x] = (int)cur[(y0 + y) * cs + x0 + x] - (int)ref[(y0 + y) * rs + x0 + x];
                s += satd8x8(d, satd_round);
            }
        return s;
    }
    for (y0 = 0; y0 < bh; y0 += 4)
        for (x0 = 0; x0 < bw; x0 += 4) {
            for (y = 0; y < 4; y++)
                for (x = 0; x < 4; x++) d[4 * y + x] = (int)cur[(y0 + y) * cs + x0 + x] - (int)ref[(y0 + y) * rs + x0 + x];
            s += satd4x4(d, satd_round);
        }
    return s;
}

/* ---- (f3) chroma samples at eighth-pel positions [STD 8.4.2.2.2] ------------------------------------------ */
/* plane: padded integer chroma plane, (0,0) of the picture at (cpad, cpad); (x, y) integer chroma position,
 * (xf, yf) in 0..7.  The caller keeps x+1, y+1 inside the padded plane. */
static int chroma_sample(const uint8_t *plane, int cstride, int cpad, int x, int y, int xf, int yf)
{
    const uint8_t *a = plane + (size_t)(y + cpad) * cstride + (x + cpad);
    return ((8 - xf) * (8 - yf) * a[0] + xf * (8 - yf) * a[1] + (8 - xf) * yf * a[cstride] + xf * yf * a[cstride + 1] + 32) >> 6;
}
int jmme_getSubImagesChroma(const uint8_t *chroma, int width, int height, int stride, int pad, uint8_t *out)
{
    int x, y, xf, yf;
    const int ps = width + 2 * pad, ph = height + 2 * pad;
    if (!chroma || !out || width <= 0 || height <= 0 || pad < 0 || stride < width) return JMME_ERR_PARAM;
    for (yf = 0; yf < 8; yf++)
        for (xf = 0; xf < 8; xf++) {
            uint8_t *o = out + (size_t)ps * ph * (yf * 8 + xf);
            for (y = 0; y < ph; y++)
                for (x = 0; x < ps; x++) {
#define CS(xx, yy) chroma[(size_t)clampi((yy) - pad, 0, height - 1) * stride + clampi((xx) - pad, 0, width - 1)]
                    o[(size_t)y * ps + x] = (uint8_t)(((8 - xf) * (8 - yf) * CS(x, y) + xf * (8 - yf) * CS(x + 1, y) +
                                                       (8 - xf) * yf * CS(x, y + 1) + xf * yf * CS(x + 1, y + 1) + 32) >> 6);
#undef CS
                }
        }
    return JMME_OK;
}

/* ---- context ----------------------------------------------------------------------------- */
static int set_err(jmme_ctx *c, int code, const char *msg)
{
    if (c) snprintf(c->err, sizeof c->err, "%s", msg);
    return code;
}
void jmme_default_params(jmme_params *p)
{
    memset(p, 0, sizeof *p);
    p->search_range = 32; p->num_refs = 1; p->blocktype_mask = JMME_MASK_ALL;
    p->qp = 28; p->rdopt = 0; p->use_hadamard = 1; p->subpel = 0;
    p->search_mode = JMME_SEARCH_FASTFULL; p->pred_policy = JMME_PRED_ZERO;
}
/* replication border: the window centre moves up to cmax samples from the MB, the window R more, 16 for the MB
 * and the sub-pel / 6-tap margin */
static int pad_for(int R, int cmax) { return (cmax + R + 16 + 15) & ~15; }

int jmme_create(jmme_ctx **out, const jmme_params *p)
{
    jmme_ctx *c;
    int v;
    if (!out || !p) return JMME_ERR_PARAM;
    *out = NULL;
    if (p->width <= 0 || p->height <= 0 || p->search_range < 1 ||
        p->search_range > JMME_MAX_SEARCH_RANGE || p->num_refs < 1 || p->num_refs > JMME_MAX_REFS ||
        (p->blocktype_mask & ~JMME_MASK_ALL) || !(p->blocktype_mask & JMME_MASK_ALL) ||
        p->qp < 0 || p->qp > 51 || p->lambda_factor < 0 ||
        p->search_mode < 0 || p->search_mode > 1 || p->pred_policy < 0 || p->pred_policy > 3 ||
        p->satd_round < 0 || p->satd_round > 1 || p->slice_rows < 0 || p->cost_domain < 0 || p->cost_domain > 1 ||
        p->me_distortion < 0 || p->me_distortion > 1 || p->transform8x8 < 0 || p->transform8x8 > 1 ||
        p->chroma_me < 0 || p->chroma_me > 1 || p->jm_center < 0 || p->jm_center > 1 ||
        (p->max_pred_qpel && (p->max_pred_qpel < 4 || p->max_pred_qpel > JMME_MAX_PRED_QPEL)))
        return JMME_ERR_PARAM;
    if (p->me_distortion &&
        (p->me_distortion_fpel < 0 || p->me_distortion_fpel > 2 || p->me_distortion_hpel < 0 || p->me_distortion_hpel > 2 ||
         p->me_distortion_qpel < 0 || p->me_distortion_qpel > 2))
        return JMME_ERR_PARAM;
    /* the integer stage builds its surfaces from per-pixel sums (SAD or SSE); a Hadamard integer stage is refused */
    if (p->me_distortion && p->me_distortion_fpel == JMME_DIST_HADAMARD) return JMME_ERR_UNSUPPORTED;
    if (p->chroma_me && !p->subpel) return JMME_ERR_PARAM;      /* chroma enters at the sub-pel stages */
    c = (jmme_ctx *)calloc(1, sizeof *c);
    if (!c) return JMME_ERR_NOMEM;
    c->p = *p;
    c->w16 = (p->width + 15) & ~15; c->h16 = (p->height + 15) & ~15;
    c->mb_w = c->w16 / 16; c->mb_h = c->h16 / 16;
    if (c->p.mb_row_end == 0) c->p.mb_row_end = c->mb_h;
    if (c->p.mb_row_begin < 0 || c->p.mb_row_end > c->mb_h || c->p.mb_row_begin >= c->p.mb_row_end) {
        free(c); return JMME_ERR_PARAM;
    }
    if (p->pred_policy == JMME_PRED_MEDIAN) {       /* stripes start and end on slice boundaries */
        const int k = p->slice_rows ? p->slice_rows : c->mb_h;
        if (c->p.mb_row_begin % k || (c->p.mb_row_end % k && c->p.mb_row_end != c->mb_h)) {
            free(c); return JMME_ERR_PARAM;
        }
    }
    c->max_pred = p->max_pred_qpel ? p->max_pred_qpel : MAX_PRED;
    /* centre limit: +-R, or (JM: rdopt = 1 does not clamp) as far as the largest predictor reaches */
    c->cmax = (p->jm_center && p->rdopt) ? c->max_pred / 4 : p->search_range;
    if (c->cmax < p->search_range) c->cmax = p->search_range;
    c->pad = pad_for(p->search_range, c->cmax);
    c->pstride = c->w16 + 2 * c->pad; c->pheight = c->h16 + 2 * c->pad;
    {
        const int lf16 = p->lambda_factor ? p->lambda_factor : jmme_lambda_factor(p->qp, p->rdopt);
        const double lambda = p->lambda_factor ? (double)p->lambda_factor / 65536.0 : lambda_motion(p->qp, p->rdopt);
        int st;
        if (lf16 > (96 << 16)) { free(c); return JMME_ERR_PARAM; }
        c->metric[0] = p->me_distortion ? p->me_distortion_fpel : JMME_DIST_SAD;
        c->metric[1] = p->me_distortion ? p->me_distortion_hpel : (p->use_hadamard ? JMME_DIST_HADAMARD : JMME_DIST_SAD);
        c->metric[2] = p->me_distortion ? p->me_distortion_qpel : (p->use_hadamard ? JMME_DIST_HADAMARD : JMME_DIST_SAD);
        for (st = 0; st < 3; st++) c->lf[st] = stage_lambda_factor(p->cost_domain, lambda, c->metric[st]);
        c->lambda_factor = c->lf[0];
    }
    c->cpad = c->pad / 2; c->cstride = c->w16 / 2 + 2 * c->cpad; c->cheight = c->h16 / 2 + 2 * c->cpad;
    c->n_planes = p->subpel ? 16 : 1;
    c->ncand = (2 * p->search_range + 1) * (2 * p->search_range + 1);
    c->mvbits = (int32_t *)malloc(sizeof(int32_t) * (2 * MAX_MVD + 1));
    c->spx = (int16_t *)malloc(sizeof(int16_t) * c->ncand);
    c->spy = (int16_t *)malloc(sizeof(int16_t) * c->ncand);
    if (!c->mvbits || !c->spx || !c->spy) { jmme_destroy(c); return JMME_ERR_NOMEM; }
    for (v = -MAX_MVD; v <= MAX_MVD; v++) c->mvbits[v + MAX_MVD] = se_bits(v);
    build_spiral(p->search_range, c->spx, c->spy);
    *out = c;
    return JMME_OK;
}
int jmme_destroy(jmme_ctx *c)
{
    int r;
    if (!c) return JMME_OK;
    for (r = 0; r < JMME_MAX_REFS; r++) { free(c->planes[r]); free(c->cplanes[r][0]); free(c->cplanes[r][1]); }
    free(c->cur_c[0]); free(c->cur_c[1]); free(c->planes_l1);
    free(c->med_pred); free(c->fmv); free(c->fref);
    free(c->mvbits); free(c->spx); free(c->spy); free(c);
    return JMME_OK;
}
const char *jmme_strerror(int code)
{
    switch (code) {
    case JMME_OK: return "ok";
    case JMME_ERR_PARAM: return "invalid parameter";
    case JMME_ERR_CUDA: return "CUDA error";
    case JMME_ERR_NOMEM: return "out of memory";
    case JMME_ERR_UNSUPPORTED: return "unsupported configuration";
    case JMME_ERR_STATE: return "invalid state (reference not set?)";
    case JMME_ERR_NODEVICE: return "no CUDA device";
    default: return "unknown error";
    }
}
const char *jmme_last_error(const jmme_ctx *c) { return c ? c->err : ""; }
const char *jmme_backend(void) { return "cpu-oracle"; }
/* oracle-only extension (not part of jmme.h): worker threads used by jmme_set_reference and
 * jmme_search_frame.  1 = the single-threaded JM-like baseline (default); bench.py's
 * `--impl reference` arm raises it to the host's core count.  Returns the value in effect. */
int jmme_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n < 1) n = omp_get_num_procs();
    g_threads = n;
#else
    (void)n;
    g_threads = 1;
#endif
    return g_threads;
}
int jmme_abi_version(void) { return JMME_ABI_VERSION; }
int jmme_mb_width(const jmme_ctx *c) { return c ? c->mb_w : 0; }
int jmme_mb_height(const jmme_ctx *c) { return c ? c->mb_h : 0; }
int jmme_pad(const jmme_ctx *c) { return c ? c->pad : 0; }
int jmme_lambda_factor_of(const jmme_ctx *c) { return c ? c->lambda_factor : 0; }
int jmme_set_peer_fields_dev(jmme_ctx *c, void *const *p, int n)
{
    (void)p; (void)n;
    return set_err(c, JMME_ERR_UNSUPPORTED, "device pointers: CUDA library only");
}
int jmme_set_multicast_field_dev(jmme_ctx *c, void *p)
{
    (void)p;
    return c ? JMME_ERR_UNSUPPORTED : JMME_ERR_PARAM;
}
int jmme_push_stripe_dev(jmme_ctx *c, const void *l, void *const *p, int n, void *st)
{ (void)l; (void)p; (void)n; (void)st; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no device path"); }
int64_t jmme_launch_count(const jmme_ctx *c) { (void)c; return 0; }
int jmme_set_tuning(jmme_ctx *c, const jmme_tuning *t) { (void)t; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no launch tuning"); }
int jmme_get_tuning(const jmme_ctx *c, jmme_tuning *t) { (void)c; if (t) memset(t, 0, sizeof *t); return JMME_ERR_UNSUPPORTED; }
const char *jmme_last_kernel(const jmme_ctx *c) { (void)c; return "cpu-oracle"; }
int jmme_set_profiling(jmme_ctx *c, int e) { (void)e; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no kernels"); }
int jmme_get_kernel_times(jmme_ctx *c, float ms[4]) { (void)ms; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no kernels"); }

int jmme_set_reference(jmme_ctx *c, int r, const uint8_t *luma, int stride)
{
    size_t sz;
    int rc;
    if (!c || !luma || r < 0 || r >= c->p.num_refs || stride < c->p.width) return JMME_ERR_PARAM;
    sz = (size_t)c->pstride * c->pheight * c->n_planes;
    if (!c->planes[r]) c->planes[r] = (uint8_t *)malloc(sz);
    if (!c->planes[r]) return set_err(c, JMME_ERR_NOMEM, "plane allocation failed");
    rc = build_planes(luma, c->p.width, c->p.height, stride, c->w16, c->h16, c->pad, c->n_planes,
                      c->planes[r]);
    if (rc == JMME_OK) c->ref_set[r] = 1;
    return rc;
}
/* replicate a (w x h) chroma picture into a plane of (pw x ph) with `pad` border samples on every side */
static uint8_t *pad_chroma(const uint8_t *src, int w, int h, int stride, int pw, int ph, int pad, uint8_t *dst)
{
    int x, y;
    if (!dst) dst = (uint8_t *)malloc((size_t)pw * ph);
    if (!dst) return NULL;
    for (y = 0; y < ph; y++)
        for (x = 0; x < pw; x++)
            dst[(size_t)y * pw + x] = src[(size_t)clampi(y - pad, 0, h - 1) * stride + clampi(x - pad, 0, w - 1)];
    return dst;
}
int jmme_set_reference_chroma(jmme_ctx *c, int r, const uint8_t *cb, const uint8_t *cr, int stride)
{
    const uint8_t *src[2];
    int k;
    if (!c || !cb || !cr || r < 0 || r >= c->p.num_refs || stride < (c->p.width + 1) / 2) return JMME_ERR_PARAM;
    if (!c->p.chroma_me) return set_err(c, JMME_ERR_STATE, "chroma_me is off");
    src[0] = cb; src[1] = cr;
    for (k = 0; k < 2; k++) {
        c->cplanes[r][k] = pad_chroma(src[k], (c->p.width + 1) / 2, (c->p.height + 1) / 2, stride, c->cstride, c->cheight,
                                      c->cpad, c->cplanes[r][k]);
        if (!c->cplanes[r][k]) return set_err(c, JMME_ERR_NOMEM, "chroma plane allocation failed");
    }
    c->cref_set[r] = 1;
    return JMME_OK;
}
int jmme_set_current_chroma(jmme_ctx *c, const uint8_t *cb, const uint8_t *cr, int stride)
{
    const uint8_t *src[2];
    int k;
    if (!c || !cb || !cr || stride < (c->p.width + 1) / 2) return JMME_ERR_PARAM;
    if (!c->p.chroma_me) return set_err(c, JMME_ERR_STATE, "chroma_me is off");
    src[0] = cb; src[1] = cr;
    for (k = 0; k < 2; k++) {
        c->cur_c[k] = pad_chroma(src[k], (c->p.width + 1) / 2, (c->p.height + 1) / 2, stride, c->w16 / 2, c->h16 / 2, 0, c->cur_c[k]);
        if (!c->cur_c[k]) return set_err(c, JMME_ERR_NOMEM, "chroma allocation failed");
    }
    c->cur_c_set = 1;
    return JMME_OK;
}
int jmme_get_subimage(jmme_ctx *c, int r, int xf, int yf, uint8_t *dst, int dst_stride)
{
    int y;
    const uint8_t *src;
    if (!c || !dst || r < 0 || r >= c->p.num_refs || xf < 0 || xf > 3 || yf < 0 || yf > 3 ||
        dst_stride < c->pstride)
        return JMME_ERR_PARAM;
    if (!c->ref_set[r]) return JMME_ERR_STATE;
    if ((xf || yf) && c->n_planes != 16) return JMME_ERR_STATE;
    src = c->planes[r] + (size_t)c->pstride * c->pheight * (yf * 4 + xf);
    for (y = 0; y < c->pheight; y++) memcpy(dst + (size_t)y * dst_stride, src + (size_t)y * c->pstride, c->pstride);
    return JMME_OK;
}
int jmme_set_reference_dev(jmme_ctx *c, int r, const void *d, int s, void *st)
{ (void)r; (void)d; (void)s; (void)st; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no device path"); }
int jmme_set_reference_chroma_dev(jmme_ctx *c, int r, const void *a, const void *b, int s, void *st)
{ (void)r; (void)a; (void)b; (void)s; (void)st; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no device path"); }
int jmme_set_current_chroma_dev(jmme_ctx *c, const void *a, const void *b, int s, void *st)
{ (void)a; (void)b; (void)s; (void)st; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no device path"); }
int jmme_search_frame_dev(jmme_ctx *c, const void *a, int s, const void *b, void *o, void *o2, void *st)
{ (void)a; (void)s; (void)b; (void)o; (void)o2; (void)st; return set_err(c, JMME_ERR_UNSUPPORTED, "oracle has no device path"); }

/* ---- distortion -------------------------------------------------------------------------- */
/* padded-plane sample fetch; (x,y) relative to sample (0,0) of the unpadded picture */
static const uint8_t *plane_at(const uint8_t *base, int pstride, int pad, int x, int y)
{
    return base + (size_t)(y + pad) * pstride + (x + pad);
}
/* what the stages of one block search need beyond the pictures */
typedef struct stage_cfg {
    int domain;                      /* cost domain 0 / 1 */
    int metric[3], lf[3], bonus[3];  /* per stage (integer, half, quarter): JMME_DIST_*, lambda factor, 16x16 (0,0) bonus */
    int satd_round, t8;              /* SATD rounding; 8x8 transform tiles for blocks >= 8x8 */
    int chroma;                      /* 1: sub-pel stages add both chroma blocks */
    const uint8_t *cplane[2];        /* padded integer chroma planes of the reference */
    int cstride, cpad;
    const uint8_t *cur_c[2];         /* current chroma, stride cur_cs */
    int cur_cs;
} stage_cfg;

/* distortion of a block at quarter-pel MV (qx,qy) under stage st's metric: luma from the 16 quarter-pel planes
 * (a10/a11), plus — chroma ME, sub-pel stages — both chroma blocks sampled at eighth-pel positions */
static int subpel_dist(const uint8_t *cur, int cs, const uint8_t *planes, int pstride, int pheight,
                       int pad, int bx, int by, int bw, int bh, int qx, int qy, const stage_cfg *g, int st)
{
    const uint8_t *pl = planes + (size_t)pstride * pheight * ((qy & 3) * 4 + (qx & 3));
    const uint8_t *ref = plane_at(pl, pstride, pad, bx + (qx >> 2), by + (qy >> 2));
    int s = block_dist(cur, cs, ref, pstride, bw, bh, g->metric[st], g->t8, g->satd_round);
    if (g->chroma) {
        /* 4:2:0: the luma vector in quarter-pel units is the chroma vector in eighth-pel units [STD 8.4.1.4] */
        const int cw = bw / 2, ch = bh / 2, cx0 = bx / 2 + (qx >> 3), cy0 = by / 2 + (qy >> 3);
        uint8_t pred[64];
        int k, x, y;
        for (k = 0; k < 2; k++) {
            for (y = 0; y < ch; y++)
                for (x = 0; x < cw; x++)
                    pred[8 * y + x] = (uint8_t)chroma_sample(g->cplane[k], g->cstride, g->cpad, cx0 + x, cy0 + y, qx & 7, qy & 7);
            s += block_dist(g->cur_c[k] + (size_t)(by / 2) * g->cur_cs + bx / 2, g->cur_cs, pred, 8, cw, ch,
                            g->metric[st], 0, g->satd_round);
        }
    }
    return s;
}

/* ---- (a6) SetupFastFullPelSearch + SetupLargerBlocks ------------------------------------- */
/* 4x4 SADs at every spiral position, then 4x8/8x4 <- 4x4, 8x8 <- 8x4, 16x8/8x16 <- 8x8,
 * 16x16 <- 16x8.  out[41][ncand]. */
static void setup_fastfull(const uint8_t *cur, int cs, const uint8_t *ref00, int rs, int mbx, int mby,
                           int cx, int cy, int ncand, const int16_t *spx, const int16_t *spy,
                           int bonus, int metric, int32_t *out)
{
    int pos, b, i, j;
    for (pos = 0; pos < ncand; pos++) {
        int mx = cx + spx[pos], my = cy + spy[pos];
        const uint8_t *r = ref00 + (size_t)(16 * mby + my) * rs + (16 * mbx + mx);
        int s44[4][4], s84[4][2], s48[2][4], s88[2][2];
        for (j = 0; j < 4; j++)
            for (i = 0; i < 4; i++)
                s44[j][i] = block_dist(cur + 4 * j * cs + 4 * i, cs, r + 4 * j * rs + 4 * i, rs, 4, 4, metric, 0, 0);
        for (j = 0; j < 4; j++) for (i = 0; i < 2; i++) s84[j][i] = s44[j][2 * i] + s44[j][2 * i + 1];
        for (j = 0; j < 2; j++) for (i = 0; i < 4; i++) s48[j][i] = s44[2 * j][i] + s44[2 * j + 1][i];
        for (j = 0; j < 2; j++) for (i = 0; i < 2; i++) s88[j][i] = s84[2 * j][i] + s84[2 * j + 1][i];
#define O(blk) out[(size_t)(blk) * ncand + pos]
        for (j = 0; j < 4; j++) for (i = 0; i < 4; i++) O(blk_base[7] + 4 * j + i) = s44[j][i];
        for (j = 0; j < 2; j++) for (i = 0; i < 4; i++) O(blk_base[6] + 4 * j + i) = s48[j][i];
        for (j = 0; j < 4; j++) for (i = 0; i < 2; i++) O(blk_base[5] + 2 * j + i) = s84[j][i];
        for (j = 0; j < 2; j++) for (i = 0; i < 2; i++) O(blk_base[4] + 2 * j + i) = s88[j][i];
        O(blk_base[3] + 0) = s88[0][0] + s88[1][0];           /* 8x16 left, right */
        O(blk_base[3] + 1) = s88[0][1] + s88[1][1];
        O(blk_base[2] + 0) = s88[0][0] + s88[0][1];           /* 16x8 top, bottom */
        O(blk_base[2] + 1) = s88[1][0] + s88[1][1];
        b = O(blk_base[2]) + O(blk_base[2] + 1);
        if (mx == 0 && my == 0) b -= bonus;                   /* 16x16 (0,0) bias, SURVEY A.6 */
        O(0) = b;
#undef O
    }
}
int jmme_SetupFastFullPelSearch(const uint8_t *cur, int cs, const uint8_t *ref, int rs, int mbx,
                                int mby, int cx, int cy, int R, int bonus, int32_t *out)
{
    int n = (2 * R + 1) * (2 * R + 1);
    int16_t *sx, *sy;
    if (!cur || !ref || !out || R < 1 || R > JMME_MAX_SEARCH_RANGE) return JMME_ERR_PARAM;
    sx = (int16_t *)malloc(2 * n); sy = (int16_t *)malloc(2 * n);
    if (!sx || !sy) { free(sx); free(sy); return JMME_ERR_NOMEM; }
    build_spiral(R, sx, sy);
    setup_fastfull(cur, cs, ref, rs, mbx, mby, cx, cy, n, sx, sy, bonus, JMME_DIST_SAD, out);
    free(sx); free(sy);
    return JMME_OK;
}

/* ---- (a7) FastFullPelBlockMotionSearch ---------------------------------------------------- */
static void fastfull_block(const int32_t *sad, int ncand, const int16_t *spx, const int16_t *spy,
                           const int32_t *mvbits, int domain, int f, int cx, int cy, int px, int py, int pretest,
                           int bonus, int *best_pos, int *best_cost)
{
    int pos, min = INT_MAX, bp = 0;
    /* bonus: the 16x16 block's (0,0) bias when it has not been folded into the surface already */
#define COST(p_, mx, my) (dscale(domain, sad[p_]) + wcost(domain, f, mvbits[4 * (mx) - px + MAX_MVD] + mvbits[4 * (my) - py + MAX_MVD]) \
                          - (((mx) == 0 && (my) == 0) ? bonus : 0))
    if (pretest) {                                  /* MV (0,0) first when !rdopt */
        for (pos = 0; pos < ncand; pos++)
            if (cx + spx[pos] == 0 && cy + spy[pos] == 0) break;
        if (pos < ncand) { min = COST(pos, 0, 0); bp = pos; }
    }
    for (pos = 0; pos < ncand; pos++) {
        int c = COST(pos, cx + spx[pos], cy + spy[pos]);
        if (c < min) { min = c; bp = pos; }
    }
#undef COST
    *best_pos = bp; *best_cost = min;
}
int jmme_FastFullPelBlockMotionSearch(const int32_t *sad, int R, int cx, int cy, int px, int py, int f,
                                      int pretest, int16_t *mvx, int16_t *mvy, int32_t *cost)
{
    int n = (2 * R + 1) * (2 * R + 1), bp, bc, v;
    int16_t *sx, *sy;
    int32_t *mvb;
    if (!sad || !mvx || !mvy || !cost || R < 1 || R > JMME_MAX_SEARCH_RANGE) return JMME_ERR_PARAM;
    if (abs(px) > MAX_PRED || abs(py) > MAX_PRED) return JMME_ERR_PARAM;
    sx = (int16_t *)malloc(2 * n); sy = (int16_t *)malloc(2 * n);
    mvb = (int32_t *)malloc(sizeof(int32_t) * (2 * MAX_MVD + 1));
    if (!sx || !sy || !mvb) { free(sx); free(sy); free(mvb); return JMME_ERR_NOMEM; }
    for (v = -MAX_MVD; v <= MAX_MVD; v++) mvb[v + MAX_MVD] = se_bits(v);
    build_spiral(R, sx, sy);
    fastfull_block(sad, n, sx, sy, mvb, 0, f, cx, cy, px, py, pretest, 0, &bp, &bc);
    *mvx = (int16_t)(cx + sx[bp]); *mvy = (int16_t)(cy + sy[bp]); *cost = bc;
    free(sx); free(sy); free(mvb);
    return JMME_OK;
}

/* ---- (a8) FullPelBlockMotionSearch -------------------------------------------------------- */
static void full_block(const uint8_t *cur, int cs, const uint8_t *ref00, int rs, int bx, int by, int bw,
                       int bh, int cx, int cy, int px, int py, int ncand, const int16_t *spx,
                       const int16_t *spy, const int32_t *mvbits, int domain, int metric, int f, int bonus, int *bmx,
                       int *bmy, int *bcost)
{
    int pos, min = INT_MAX, mx0 = cx, my0 = cy;
    for (pos = 0; pos < ncand; pos++) {
        int mx = cx + spx[pos], my = cy + spy[pos];
        int c = wcost(domain, f, mvbits[4 * mx - px + MAX_MVD] + mvbits[4 * my - py + MAX_MVD]);
        c += dscale(domain, block_dist(cur + (size_t)by * cs + bx, cs, ref00 + (size_t)(by + my) * rs + (bx + mx), rs, bw, bh,
                                       metric, 0, 0));
        if (mx == 0 && my == 0) c -= bonus;
        if (c < min) { min = c; mx0 = mx; my0 = my; }
    }
    *bmx = mx0; *bmy = my0; *bcost = min;
}
int jmme_FullPelBlockMotionSearch(const uint8_t *cur, int cs, const uint8_t *ref, int rs, int bx, int by,
                                  int bw, int bh, int px, int py, int R, int f, int bonus, int16_t *mvx,
                                  int16_t *mvy, int32_t *cost)
{
    int n = (2 * R + 1) * (2 * R + 1), v, mx, my, mc, cx, cy;
    int16_t *sx, *sy;
    int32_t *mvb;
    if (!cur || !ref || !mvx || !mvy || !cost || R < 1 || R > JMME_MAX_SEARCH_RANGE) return JMME_ERR_PARAM;
    if (abs(px) > MAX_PRED || abs(py) > MAX_PRED) return JMME_ERR_PARAM;
    sx = (int16_t *)malloc(2 * n); sy = (int16_t *)malloc(2 * n);
    mvb = (int32_t *)malloc(sizeof(int32_t) * (2 * MAX_MVD + 1));
    if (!sx || !sy || !mvb) { free(sx); free(sy); free(mvb); return JMME_ERR_NOMEM; }
    for (v = -MAX_MVD; v <= MAX_MVD; v++) mvb[v + MAX_MVD] = se_bits(v);
    build_spiral(R, sx, sy);
    cx = clampi(px / 4, -R, R); cy = clampi(py / 4, -R, R);
    full_block(cur, cs, ref, rs, bx, by, bw, bh, cx, cy, px, py, n, sx, sy, mvb, 0, JMME_DIST_SAD, f, bonus, &mx, &my, &mc);
    *mvx = (int16_t)mx; *mvy = (int16_t)my; *cost = mc;
    free(sx); free(sy); free(mvb);
    return JMME_OK;
}

/* ---- (a10) SubPelBlockMotionSearch -------------------------------------------------------- */
/* In: integer MV (quarter-pel units, multiple of 4) and its integer-search cost.
 * Half-pel: spiral positions 1..8, step 2; quarter-pel: positions 1..8, step 1, around the half-pel winner.
 * A stage whose metric differs from the previous stage's (or any sub-pel stage with chroma ME) starts at
 * position 0 with the running minimum reset (JM start_me_refinement_hp / _qp).  Strict <.  No early
 * termination (result-neutral in JM except in combination with the (0,0) bonus; the frozen spec is the plain
 * argmin). */
static void subpel_block(const uint8_t *cur, int cs, const uint8_t *planes, int pstride, int pheight,
                         int pad, int bx, int by, int bw, int bh, int px, int py, const int32_t *mvbits,
                         const stage_cfg *g, const int16_t *spx, const int16_t *spy, int *mvx, int *mvy, int *cost)
{
    int pos, step, bmx = *mvx, bmy = *mvy, min = *cost, prev = g->metric[0];
    for (step = 2; step >= 1; step--) {
        const int st = step == 2 ? 1 : 2;
        const int restart = g->chroma || g->metric[st] != prev;
        int ox = bmx, oy = bmy, best = 0;
        if (restart) min = INT_MAX;
        for (pos = restart ? 0 : 1; pos < 9; pos++) {
            int qx = ox + step * spx[pos], qy = oy + step * spy[pos];
            int c = wcost(g->domain, g->lf[st], mvbits[qx - px + MAX_MVD] + mvbits[qy - py + MAX_MVD]);
            c += dscale(g->domain, subpel_dist(cur + (size_t)by * cs + bx, cs, planes, pstride, pheight, pad, bx, by, bw, bh,
                                               qx, qy, g, st));
            if (qx == 0 && qy == 0) c -= g->bonus[st];
            if (c < min) { min = c; best = pos; }
        }
        bmx = ox + step * spx[best]; bmy = oy + step * spy[best];
        prev = g->metric[st];
    }
    *mvx = bmx; *mvy = bmy; *cost = min;
}
int jmme_SubPelBlockMotionSearch(const uint8_t *cur, int cs, const uint8_t *planes, int width, int height,
                                 int pad, int bx, int by, int bw, int bh, int px, int py, int f,
                                 int hadamard, int satd_round, int bonus, int16_t *mvx, int16_t *mvy,
                                 int32_t *cost)
{
    int16_t sx[9], sy[9];
    int32_t *mvb;
    int v, x, y, c, st;
    stage_cfg g;
    if (!cur || !planes || !mvx || !mvy || !cost) return JMME_ERR_PARAM;
    if (abs(px) > MAX_PRED || abs(py) > MAX_PRED) return JMME_ERR_PARAM;
    mvb = (int32_t *)malloc(sizeof(int32_t) * (2 * MAX_MVD + 1));
    if (!mvb) return JMME_ERR_NOMEM;
    for (v = -MAX_MVD; v <= MAX_MVD; v++) mvb[v + MAX_MVD] = se_bits(v);
    build_spiral(1, sx, sy);
    x = *mvx; y = *mvy; c = *cost;
    memset(&g, 0, sizeof g);                        /* the leaf is the legacy (JM <= 10) form: domain 0 */
    g.metric[0] = JMME_DIST_SAD; g.metric[1] = g.metric[2] = hadamard ? JMME_DIST_HADAMARD : JMME_DIST_SAD;
    for (st = 0; st < 3; st++) { g.lf[st] = f; g.bonus[st] = bonus; }
    g.satd_round = satd_round;
    subpel_block(cur, cs, planes, width + 2 * pad, height + 2 * pad, pad, bx, by, bw, bh, px, py, mvb, &g,
                 sx, sy, &x, &y, &c);
    *mvx = (int16_t)x; *mvy = (int16_t)y; *cost = c;
    free(mvb);
    return JMME_OK;
}

/* ---- (a3) SetMotionVectorPredictor ‖ GetMotionVectorPredictorNormal [STD 8.4.1.3] ------------ */
static int median3(int a, int b, int c)
{
    int mn = a < b ? (a < c ? a : c) : (b < c ? b : c);
    int mx = a > b ? (a > c ? a : c) : (b > c ? b : c);
    return a + b + c - mn - mx;
}
int jmme_SetMotionVectorPredictor(int blocktype, int part, int ref_idx, const int16_t mvA[2], int refA, int availA,
                                  const int16_t mvB[2], int refB, int availB, const int16_t mvC[2], int refC,
                                  int availC, int16_t pred[2])
{
    int a[2], b[2], c[2], k, n;
    if (blocktype < 1 || blocktype > 7 || !mvA || !mvB || !mvC || !pred) return JMME_ERR_PARAM;
    for (k = 0; k < 2; k++) { a[k] = availA ? mvA[k] : 0; b[k] = availB ? mvB[k] : 0; c[k] = availC ? mvC[k] : 0; }
    if (!availA) refA = -1;
    if (!availB) refB = -1;
    if (!availC) refC = -1;
    if (refA < 0) a[0] = a[1] = 0;                     /* intra / other list: no vector */
    if (refB < 0) b[0] = b[1] = 0;
    if (refC < 0) c[0] = c[1] = 0;
    /* directional prediction of 16x8 and 8x16 */
    if (blocktype == 2) {
        if (part == 0 && refB == ref_idx) { pred[0] = (int16_t)b[0]; pred[1] = (int16_t)b[1]; return JMME_OK; }
        if (part == 1 && refA == ref_idx) { pred[0] = (int16_t)a[0]; pred[1] = (int16_t)a[1]; return JMME_OK; }
    } else if (blocktype == 3) {
        if (part == 0 && refA == ref_idx) { pred[0] = (int16_t)a[0]; pred[1] = (int16_t)a[1]; return JMME_OK; }
        if (part == 1 && refC == ref_idx) { pred[0] = (int16_t)c[0]; pred[1] = (int16_t)c[1]; return JMME_OK; }
    }
    /* median prediction */
    if (!availB && !availC && availA) {
        b[0] = c[0] = a[0]; b[1] = c[1] = a[1]; refB = refC = refA;
    }
    n = (refA == ref_idx) + (refB == ref_idx) + (refC == ref_idx);
    if (n == 1) {
        const int *m = refA == ref_idx ? a : (refB == ref_idx ? b : c);
        pred[0] = (int16_t)m[0]; pred[1] = (int16_t)m[1];
    } else {
        pred[0] = (int16_t)median3(a[0], b[0], c[0]);
        pred[1] = (int16_t)median3(a[1], b[1], c[1]);
    }
    return JMME_OK;
}

/* Is field cell (x,y) (4x4 units, frame coordinates) decoded before block `blk` of MB (mbx,mby)? */
static int cell_available(const jmme_ctx *c, int x, int y, int mbx, int mby, int t, int x0, int y0)
{
    int mx, my, lx, ly, w = blc_w[t], h = blc_h[t];
    if (x < 0 || y < 0 || x >= 4 * c->mb_w || y >= 4 * c->mb_h) return 0;
    mx = x >> 2; my = y >> 2;
    if (my != mby) return my < mby;                 /* rows above: decoded; rows below: not */
    if (mx != mbx) return mx < mbx;                 /* same row: only MBs to the left */
    lx = (x & 3) * 4; ly = (y & 3) * 4;             /* same MB: partition decoding order */
    if (t == 1) return 0;
    if (t == 2) return (ly >= 8) < (y0 >= 8);
    if (t == 3) return (lx >= 8) < (x0 >= 8);
    {
        int q = 2 * (y0 >= 8) + (x0 >= 8), qc = 2 * (ly >= 8) + (lx >= 8);
        int s = ((y0 & 7) / h) * (8 / w) + ((x0 & 7) / w), sc = ((ly & 7) / h) * (8 / w) + ((lx & 7) / w);
        if (qc != q) return qc < q;
        return sc < s;
    }
}

/* Predictor of block (t; i,j) of MB (mbx,mby) for reference r from the field (mv4, ref4).
 * slice_top: first MB row of the slice (rows above it are unavailable).
 * p16: NULL = neighbour cells inside this MB are read from the field (a field of an earlier pass);
 *      else they carry (p16, r): the in-frame median policy, DESIGN.md §2. */
static void predict_block(const jmme_ctx *c, const int16_t *mv4, const int8_t *ref4, int r, int mbx, int mby, int t,
                          int i, int j, int slice_top, const int16_t *p16, int16_t *o)
{
    static const int16_t zero[2] = {0, 0};
    const int fw = 4 * c->mb_w, bw = blc_w[t], bh = blc_h[t];
    const int x0 = i * bw, y0 = j * bh;
    const int cx = 4 * mbx + x0 / 4, cy = 4 * mby + y0 / 4, wc = bw / 4;
    const int nx[4] = {cx - 1, cx, cx + wc, cx - 1}, ny[4] = {cy, cy - 1, cy - 1, cy - 1};
    const int16_t *mv[4];
    int rf[4], av[4], k, part = t == 2 ? j : (t == 3 ? i : 0);
    for (k = 0; k < 4; k++) {
        av[k] = cell_available(c, nx[k], ny[k], mbx, mby, t, x0, y0) && ny[k] >= 4 * slice_top;
        mv[k] = av[k] ? mv4 + ((size_t)ny[k] * fw + nx[k]) * 2 : zero;
        rf[k] = av[k] ? ref4[(size_t)ny[k] * fw + nx[k]] : -1;
        if (av[k] && p16 && (nx[k] >> 2) == mbx && (ny[k] >> 2) == mby) { mv[k] = p16; rf[k] = r; }
    }
    if (!av[2]) { av[2] = av[3]; mv[2] = mv[3]; rf[2] = rf[3]; }     /* C := D */
    jmme_SetMotionVectorPredictor(t, part, r, mv[0], rf[0], av[0], mv[1], rf[1], av[1], mv[2], rf[2], av[2], o);
}

/* all 41 predictors of one MB and reference; o = int16 [41][2] */
static void predict_mb(const jmme_ctx *c, const int16_t *mv4, const int8_t *ref4, int r, int mbx, int mby,
                       int slice_top, int in_frame, int16_t *o)
{
    int t, i, j;
    for (t = 1; t <= 7; t++) {
        const int nbx = 16 / blc_w[t], nby = 16 / blc_h[t];
        for (j = 0; j < nby; j++)
            for (i = 0; i < nbx; i++)     /* block 0 (16x16) comes first: o[0..1] is p16 for the others */
                predict_block(c, mv4, ref4, r, mbx, mby, t, i, j, slice_top, in_frame && t > 1 ? o : NULL,
                              o + 2 * (blk_base[t] + j * nbx + i));
    }
}

int jmme_predict_frame(jmme_ctx *c, const int16_t *mv4, const int8_t *ref4, int16_t *pred)
{
    int r, mb;
    if (!c || !mv4 || !ref4 || !pred) return JMME_ERR_PARAM;
    for (r = 0; r < c->p.num_refs; r++)
        for (mb = 0; mb < c->mb_w * c->mb_h; mb++)
            predict_mb(c, mv4, ref4, r, mb % c->mb_w, mb / c->mb_w, 0, 0,
                       pred + ((size_t)r * c->mb_w * c->mb_h + mb) * JMME_BLOCKS_PER_MB * 2);
    return JMME_OK;
}

/* ME-only mode decision: DESIGN.md §2 */
static void commit_mb(const jmme_ctx *c, const jmme_mbresult *m, int mb, int16_t *mv4, int8_t *ref4, uint8_t *mode)
{
    const int fw = 4 * c->mb_w;
    {
        const int mbx = mb % c->mb_w, mby = mb / c->mb_w;
        int64_t J[4], best;
        int sub[4] = {0, 0, 0, 0}, q, t, k, md, cell;
#define ON(tt) ((c->p.blocktype_mask >> (tt)) & 1)
        J[0] = ON(1) ? m->cost[0] : INT64_MAX;
        J[1] = ON(2) ? (int64_t)m->cost[1] + m->cost[2] : INT64_MAX;
        J[2] = ON(3) ? (int64_t)m->cost[3] + m->cost[4] : INT64_MAX;
        J[3] = 0;
        for (q = 0; q < 4; q++) {
            int64_t bq = INT64_MAX;
            for (t = 4; t <= 7; t++) {
                const int bw = blc_w[t], bh = blc_h[t], nbx = 16 / bw, per = (8 / bw) * (8 / bh);
                int64_t s = 0;
                if (!ON(t)) continue;
                for (k = 0; k < per; k++) {
                    const int sx = (q & 1) * (8 / bw) + k % (8 / bw), sy = (q >> 1) * (8 / bh) + k / (8 / bw);
                    s += m->cost[blk_base[t] + sy * nbx + sx];
                }
                if (s < bq) { bq = s; sub[q] = t; }
            }
            if (bq == INT64_MAX) { J[3] = INT64_MAX; break; }
            J[3] += bq;
        }
        md = 0; best = J[0];
        for (k = 1; k < 4; k++) if (J[k] < best) { best = J[k]; md = k; }
        if (mode) {
            mode[5 * mb] = (uint8_t)(md == 3 ? 8 : md + 1);
            for (q = 0; q < 4; q++) mode[5 * mb + 1 + q] = (uint8_t)(md == 3 ? sub[q] : 0);
        }
        for (cell = 0; cell < 16; cell++) {
            const int cx4 = cell & 3, cy4 = cell >> 2, q8 = 2 * (cy4 >> 1) + (cx4 >> 1);
            const int tt = md == 3 ? sub[q8] : md + 1;
            const int bw = blc_w[tt], bh = blc_h[tt], nbx = 16 / bw;
            const int blk = blk_base[tt] + ((4 * cy4) / bh) * nbx + (4 * cx4) / bw;
            const size_t o = (size_t)(4 * mby + cy4) * fw + 4 * mbx + cx4;
            mv4[2 * o] = m->mv[blk][0]; mv4[2 * o + 1] = m->mv[blk][1]; ref4[o] = m->ref_idx[blk];
        }
#undef ON
    }
}
int jmme_commit_field(jmme_ctx *c, const jmme_mbresult *res, int16_t *mv4, int8_t *ref4, uint8_t *mode)
{
    int mb;
    if (!c || !res || !mv4 || !ref4 || !mode) return JMME_ERR_PARAM;
    for (mb = 0; mb < c->mb_w * c->mb_h; mb++) commit_mb(c, &res[mb], mb, mv4, ref4, mode);
    return JMME_OK;
}

/* ---- (a4,a5) PartitionMotionSearch / BlockMotionSearch over a frame ----------------------- */
/* every (reference, blocktype, block) of one macroblock; bsad = scratch [41][ncand] (FASTFULL) */
static void search_mb(const jmme_ctx *c, const uint8_t *cur, const int16_t *pred, jmme_mbresult *out,
                      jmme_mbresult *out_per_ref, int mbx, int mby, int32_t *bsad)
{
    const int nmb = c->mb_w * c->mb_h, mb = mby * c->mb_w + mbx;
    const int npb = c->p.pred_policy >= JMME_PRED_PER_BLOCK ? JMME_BLOCKS_PER_MB : 1;
    const int dom = c->p.cost_domain;
    jmme_mbresult *o = &out[mb];
    stage_cfg g;
    int r, t, b, st;
    memset(&g, 0, sizeof g);
    g.domain = dom; g.satd_round = c->p.satd_round; g.t8 = c->p.transform8x8; g.chroma = c->p.chroma_me;
    for (st = 0; st < 3; st++) { g.metric[st] = c->metric[st]; g.lf[st] = c->lf[st]; }
    g.cstride = c->cstride; g.cpad = c->cpad; g.cur_c[0] = c->cur_c[0]; g.cur_c[1] = c->cur_c[1]; g.cur_cs = c->w16 / 2;
    for (b = 0; b < JMME_BLOCKS_PER_MB; b++) {
        o->mv[b][0] = o->mv[b][1] = 0; o->cost[b] = INT_MAX; o->ref_idx[b] = -1;
    }
    memset(o->reserved, 0, sizeof o->reserved);
    for (r = 0; r < c->p.num_refs; r++) {
        const uint8_t *ref00 = plane_at(c->planes[r], c->pstride, c->pad, 0, 0);   /* integer plane */
        const int16_t *pr = pred ? pred + ((size_t)r * nmb + mb) * npb * 2 : NULL;
        const int has_bonus = !c->p.rdopt && r == 0;       /* 16x16 (0,0) bias: !rdopt, reference 0 */
        jmme_mbresult *opr = out_per_ref ? &out_per_ref[(size_t)r * nmb + mb] : NULL;
        const int p16x = pr ? pr[0] : 0, p16y = pr ? pr[1] : 0;
        const int cx = clampi(p16x / 4, -c->cmax, c->cmax), cy = clampi(p16y / 4, -c->cmax, c->cmax);
        g.cplane[0] = c->cplanes[r][0]; g.cplane[1] = c->cplanes[r][1];
        if (opr) {
            for (b = 0; b < JMME_BLOCKS_PER_MB; b++) {
                opr->mv[b][0] = opr->mv[b][1] = 0; opr->cost[b] = INT_MAX; opr->ref_idx[b] = -1;
            }
            memset(opr->reserved, 0, sizeof opr->reserved);
        }
        if (c->p.search_mode == JMME_SEARCH_FASTFULL)
            setup_fastfull(cur + (size_t)16 * mby * c->w16 + 16 * mbx, c->w16, ref00, c->pstride, mbx, mby, cx,
                           cy, c->ncand, c->spx, c->spy, 0, c->metric[0], bsad);
        for (t = 1; t <= 7; t++) {
            const int bw = blc_w[t], bh = blc_h[t], nbx = 16 / bw, nby = 16 / bh;
            int j, i;
            if (!(c->p.blocktype_mask & (1 << t))) continue;
            for (st = 0; st < 3; st++) g.bonus[st] = (t == 1 && has_bonus) ? wcost(dom, c->lf[st], 16) : 0;
            for (j = 0; j < nby; j++)
                for (i = 0; i < nbx; i++) {
                    const int blk = blk_base[t] + j * nbx + i;
                    const int px = pr ? pr[(npb == 1 ? 0 : blk) * 2] : 0;
                    const int py = pr ? pr[(npb == 1 ? 0 : blk) * 2 + 1] : 0;
                    const int bx = 16 * mbx + i * bw, by = 16 * mby + j * bh;
                    int mvx, mvy, cost, total;
                    if (c->p.search_mode == JMME_SEARCH_FASTFULL) {
                        int bp;
                        fastfull_block(bsad + (size_t)blk * c->ncand, c->ncand, c->spx, c->spy, c->mvbits, dom,
                                       c->lf[0], cx, cy, px, py, !c->p.rdopt, g.bonus[0], &bp, &cost);
                        mvx = cx + c->spx[bp]; mvy = cy + c->spy[bp];
                    } else {
                        const int bcx = clampi(px / 4, -c->cmax, c->cmax), bcy = clampi(py / 4, -c->cmax, c->cmax);
                        full_block(cur, c->w16, ref00, c->pstride, bx, by, bw, bh, bcx, bcy, px, py, c->ncand,
                                   c->spx, c->spy, c->mvbits, dom, c->metric[0], c->lf[0], g.bonus[0], &mvx, &mvy, &cost);
                    }
                    mvx *= 4; mvy *= 4;
                    if (c->p.subpel)
                        subpel_block(cur, c->w16, c->planes[r], c->pstride, c->pheight, c->pad, bx, by, bw, bh, px,
                                     py, c->mvbits, &g, c->spx, c->spy, &mvx, &mvy, &cost);
                    if (opr) {
                        opr->mv[blk][0] = (int16_t)mvx; opr->mv[blk][1] = (int16_t)mvy;
                        opr->cost[blk] = cost; opr->ref_idx[blk] = (int8_t)r;
                    }
                    total = cost + ref_cost(c, r);
                    if (total < o->cost[blk]) {                 /* lowest ref wins ties */
                        o->mv[blk][0] = (int16_t)mvx; o->mv[blk][1] = (int16_t)mvy;
                        o->cost[blk] = total; o->ref_idx[blk] = (int8_t)r;
                    }
                }
        }
    }
}

int jmme_search_frame(jmme_ctx *c, const uint8_t *cur_in, int stride, const int16_t *pred,
                      jmme_mbresult *out, jmme_mbresult *out_per_ref)
{
    int r, x, y, nmb, npb, n_stripe, fail = 0;
    uint8_t *cur;                    /* current picture padded to x16 by replication */
    if (!c || !cur_in || !out || stride < c->p.width) return JMME_ERR_PARAM;
    if (c->p.pred_policy == JMME_PRED_MEDIAN) pred = NULL;
    else if (c->p.pred_policy != JMME_PRED_ZERO && !pred) return set_err(c, JMME_ERR_PARAM, "pred required");
    for (r = 0; r < c->p.num_refs; r++)
        if (!c->ref_set[r]) return set_err(c, JMME_ERR_STATE, "reference not set");
    if (c->p.chroma_me) {
        if (!c->cur_c_set) return set_err(c, JMME_ERR_STATE, "current chroma not set");
        for (r = 0; r < c->p.num_refs; r++)
            if (!c->cref_set[r]) return set_err(c, JMME_ERR_STATE, "reference chroma not set");
    }
    nmb = c->mb_w * c->mb_h;
    npb = c->p.pred_policy == JMME_PRED_PER_BLOCK ? JMME_BLOCKS_PER_MB : 1;
    if (c->p.pred_policy == JMME_PRED_MEDIAN && !c->med_pred) {
        c->med_pred = (int16_t *)calloc((size_t)c->p.num_refs * nmb * JMME_BLOCKS_PER_MB * 2, sizeof(int16_t));
        c->fmv = (int16_t *)calloc((size_t)nmb * 16 * 2, sizeof(int16_t));
        c->fref = (int8_t *)malloc((size_t)nmb * 16);
        if (!c->med_pred || !c->fmv || !c->fref) return JMME_ERR_NOMEM;
    }
    if (pred && c->p.pred_policy != JMME_PRED_ZERO) {       /* only the stripe's MB rows are read */
        const size_t per_mb = (size_t)npb * 2, n = (size_t)(c->p.mb_row_end - c->p.mb_row_begin) * c->mb_w * per_mb;
        size_t i;
        for (r = 0; r < c->p.num_refs; r++) {
            const int16_t *q = pred + ((size_t)r * nmb + (size_t)c->p.mb_row_begin * c->mb_w) * per_mb;
            for (i = 0; i < n; i++)
                if (q[i] > c->max_pred || q[i] < -c->max_pred) return set_err(c, JMME_ERR_PARAM, "pred out of range");
        }
    }
    if (c->p.pred_policy == JMME_PRED_ZERO) pred = NULL;
    cur = (uint8_t *)malloc((size_t)c->w16 * c->h16);
    if (!cur) return JMME_ERR_NOMEM;
    for (y = 0; y < c->h16; y++)
        for (x = 0; x < c->w16; x++)
            cur[(size_t)y * c->w16 + x] =
                cur_in[(size_t)clampi(y, 0, c->p.height - 1) * stride + clampi(x, 0, c->p.width - 1)];
    n_stripe = (c->p.mb_row_end - c->p.mb_row_begin) * c->mb_w;
#pragma omp parallel num_threads(g_threads)
    {
        int i;
        int32_t *bsad = NULL;
        if (c->p.search_mode == JMME_SEARCH_FASTFULL) {
            bsad = (int32_t *)malloc(sizeof(int32_t) * JMME_BLOCKS_PER_MB * (size_t)c->ncand);
            if (!bsad) {
#pragma omp atomic write
                fail = 1;
            }
        }
#pragma omp barrier
        if (!fail && c->p.pred_policy == JMME_PRED_MEDIAN) {
            /* in-frame median: slices are independent, MBs of a slice in raster order:
             * predict from the field committed so far -> search -> commit */
            const int k = c->p.slice_rows ? c->p.slice_rows : c->mb_h;
            const int n_slices = (c->p.mb_row_end - c->p.mb_row_begin + k - 1) / k;
#pragma omp for schedule(dynamic, 1)
            for (i = 0; i < n_slices; i++) {
                const int top = c->p.mb_row_begin + i * k;
                const int bot = top + k < c->p.mb_row_end ? top + k : c->p.mb_row_end;
                int mbx, mby, rr;
                for (mby = top; mby < bot; mby++)
                    for (mbx = 0; mbx < c->mb_w; mbx++) {
                        const int mb = mby * c->mb_w + mbx;
                        for (rr = 0; rr < c->p.num_refs; rr++)
                            predict_mb(c, c->fmv, c->fref, rr, mbx, mby, top, 1,
                                       c->med_pred + ((size_t)rr * nmb + mb) * JMME_BLOCKS_PER_MB * 2);
                        search_mb(c, cur, c->med_pred, out, out_per_ref, mbx, mby, bsad);
                        commit_mb(c, &out[mb], mb, c->fmv, c->fref, NULL);
                    }
            }
        } else if (!fail) {
#pragma omp for schedule(dynamic, 4)
            for (i = 0; i < n_stripe; i++)
                search_mb(c, cur, pred, out, out_per_ref, i % c->mb_w, c->p.mb_row_begin + i / c->mb_w, bsad);
        }
        free(bsad);
    }
    free(cur);
    return fail ? JMME_ERR_NOMEM : JMME_OK;
}

/* ---- (f2) bi-predictive refinement: JM BiPredBlockMotionSearch, restated per the header's definition -------- */
int jmme_set_reference_l1(jmme_ctx *c, const uint8_t *luma, int stride)
{
    size_t sz;
    if (!c || !luma || stride < c->p.width) return JMME_ERR_PARAM;
    sz = (size_t)c->pstride * c->pheight * c->n_planes;
    if (!c->planes_l1) c->planes_l1 = (uint8_t *)malloc(sz);
    if (!c->planes_l1) return set_err(c, JMME_ERR_NOMEM, "plane allocation failed");
    return build_planes(luma, c->p.width, c->p.height, stride, c->w16, c->h16, c->pad, c->n_planes, c->planes_l1);
}

/* sample block of reference planes `pl` at quarter-pel vector (qx, qy): pointer to its top-left sample */
static const uint8_t *pred_block(const jmme_ctx *c, const uint8_t *pl, int bx, int by, int qx, int qy)
{
    const uint8_t *plane = pl + (size_t)c->pstride * c->pheight * ((qy & 3) * 4 + (qx & 3));
    return plane_at(plane, c->pstride, c->pad, bx + (qx >> 2), by + (qy >> 2));
}

int jmme_search_frame_bipred(jmme_ctx *c, const uint8_t *cur_in, int stride, const jmme_mbresult *l0, const jmme_mbresult *l1,
                             const int16_t *pred0, const int16_t *pred1, int range, int iterations, jmme_bipred *out)
{
    const int npb = c && c->p.pred_policy >= JMME_PRED_PER_BLOCK ? JMME_BLOCKS_PER_MB : 1;
    int16_t *spx, *spy;
    uint8_t *cur;
    int x, y, mb, nmb, ncand, bad = 0;
    if (!c || !cur_in || !l0 || !l1 || !out || stride < c->p.width || range < 1 || range > 15 || iterations < 1 || iterations > 8)
        return JMME_ERR_PARAM;
    if (!c->planes_l1) return set_err(c, JMME_ERR_STATE, "list-1 reference not set");
    for (x = 0; x < c->p.num_refs; x++)
        if (!c->ref_set[x]) return set_err(c, JMME_ERR_STATE, "reference not set");
    nmb = c->mb_w * c->mb_h;
    ncand = (2 * range + 1) * (2 * range + 1);
    spx = (int16_t *)malloc(2 * ncand); spy = (int16_t *)malloc(2 * ncand);
    cur = (uint8_t *)malloc((size_t)c->w16 * c->h16);
    if (!spx || !spy || !cur) { free(spx); free(spy); free(cur); return JMME_ERR_NOMEM; }
    build_spiral(range, spx, spy);
    for (y = 0; y < c->h16; y++)
        for (x = 0; x < c->w16; x++)
            cur[(size_t)y * c->w16 + x] = cur_in[(size_t)clampi(y, 0, c->p.height - 1) * stride + clampi(x, 0, c->p.width - 1)];
#pragma omp parallel for schedule(dynamic, 4) num_threads(g_threads)
    for (mb = c->p.mb_row_begin * c->mb_w; mb < c->p.mb_row_end * c->mb_w; mb++) {
        const int mbx = mb % c->mb_w, mby = mb / c->mb_w;
        jmme_bipred *o = &out[mb];
        int t, i, j;
        memset(o, 0, sizeof *o);
        for (i = 0; i < JMME_BLOCKS_PER_MB; i++) { o->cost[i] = INT_MAX; o->ref0[i] = -1; }
        for (t = 1; t <= 7; t++) {
            const int bw = blc_w[t], bh = blc_h[t], nbx = 16 / bw, nby = 16 / bh;
            if (!(c->p.blocktype_mask & (1 << t))) continue;
            for (j = 0; j < nby; j++)
                for (i = 0; i < nbx; i++) {
                    const int blk = blk_base[t] + j * nbx + i, bx = 16 * mbx + i * bw, by = 16 * mby + j * bh;
                    const int r0 = l0[mb].ref_idx[blk];
                    const uint8_t *pl[2];
                    int mv[2][2], pr[2][2], it, best_cost = INT_MAX;
                    if (r0 < 0 || r0 >= c->p.num_refs || l1[mb].ref_idx[blk] < 0) {
#pragma omp atomic write
                        bad = 1;
                        continue;
                    }
                    pl[0] = c->planes[r0]; pl[1] = c->planes_l1;
                    mv[0][0] = l0[mb].mv[blk][0]; mv[0][1] = l0[mb].mv[blk][1];
                    mv[1][0] = l1[mb].mv[blk][0]; mv[1][1] = l1[mb].mv[blk][1];
                    if (c->n_planes == 1 && ((mv[0][0] | mv[0][1] | mv[1][0] | mv[1][1]) & 3)) {
#pragma omp atomic write
                        bad = 1;
                        continue;
                    }
                    {
                        const int16_t *p0 = pred0 ? pred0 + ((size_t)r0 * nmb + mb) * npb * 2 + (npb == 1 ? 0 : 2 * blk) : NULL;
                        const int16_t *p1 = pred1 ? pred1 + (size_t)mb * npb * 2 + (npb == 1 ? 0 : 2 * blk) : NULL;
                        pr[0][0] = p0 ? p0[0] : 0; pr[0][1] = p0 ? p0[1] : 0;
                        pr[1][0] = p1 ? p1[0] : 0; pr[1][1] = p1 ? p1[1] : 0;
                    }
                    for (it = 0; it < iterations; it++) {
                        const int s = it & 1, f = 1 - s;
                        const uint8_t *fb = pred_block(c, pl[f], bx, by, mv[f][0], mv[f][1]);
                        const int fbits = c->mvbits[mv[f][0] - pr[f][0] + MAX_MVD] + c->mvbits[mv[f][1] - pr[f][1] + MAX_MVD];
                        int pos, bp = -1, min = INT_MAX;
                        for (pos = 0; pos < ncand; pos++) {
                            const int qx = mv[s][0] + 4 * spx[pos], qy = mv[s][1] + 4 * spy[pos];
                            const uint8_t *sb;
                            int xx, yy, d = 0, cst;
                            if ((qx >> 2) < -(c->pad - 1) || (qx >> 2) > c->pad - 1 || (qy >> 2) < -(c->pad - 1) || (qy >> 2) > c->pad - 1)
                                continue;
                            sb = pred_block(c, pl[s], bx, by, qx, qy);
                            for (yy = 0; yy < bh; yy++)
                                for (xx = 0; xx < bw; xx++) {
                                    const int p = (fb[(size_t)yy * c->pstride + xx] + sb[(size_t)yy * c->pstride + xx] + 1) >> 1;
                                    const int e = (int)cur[(size_t)(by + yy) * c->w16 + bx + xx] - p;
                                    d += c->metric[0] == JMME_DIST_SSE ? e * e : (e < 0 ? -e : e);
                                }
                            cst = dscale(c->p.cost_domain, d) +
                                  wcost(c->p.cost_domain, c->lf[0], fbits + c->mvbits[qx - pr[s][0] + MAX_MVD] + c->mvbits[qy - pr[s][1] + MAX_MVD]);
                            if (cst < min) { min = cst; bp = pos; }
                        }
                        if (bp >= 0) { mv[s][0] += 4 * spx[bp]; mv[s][1] += 4 * spy[bp]; best_cost = min; }
                    }
                    o->mv0[blk][0] = (int16_t)mv[0][0]; o->mv0[blk][1] = (int16_t)mv[0][1];
                    o->mv1[blk][0] = (int16_t)mv[1][0]; o->mv1[blk][1] = (int16_t)mv[1][1];
                    o->cost[blk] = best_cost; o->ref0[blk] = (int8_t)r0;
                }
        }
    }
    free(spx); free(spy); free(cur);
    return bad ? set_err(c, JMME_ERR_PARAM, "l0 / l1 records: reference index or vector phase not usable") : JMME_OK;
}

int jmme_get_predictors(jmme_ctx *c, int16_t *pred)
{
    if (!c || !pred) return JMME_ERR_PARAM;
    if (c->p.pred_policy != JMME_PRED_MEDIAN || !c->med_pred) return set_err(c, JMME_ERR_STATE, "no median search yet");
    memcpy(pred, c->med_pred, sizeof(int16_t) * 2 * JMME_BLOCKS_PER_MB * (size_t)c->p.num_refs * c->mb_w * c->mb_h);
    return JMME_OK;
}
