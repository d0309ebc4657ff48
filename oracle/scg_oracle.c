/* TEST INFRASTRUCTURE ONLY -- scalar CPU restatement (C, OpenMP) of the Gaussian rasterizer
 * hot path: preprocess -> tile binning (64-bit tile|depth keys, stable radix sort) -> per-tile
 * alpha compositing, forward and explicit (hand-derived) backward.
 *
 * PARITY UNPINNED: the reference tree vendors no rasterizer source, tests or golden vectors
 * (SURVEY.md sections 0 and 8c).  The algorithm is the one published by the third-party package
 * the reference imports at gaussian_renderer/__init__.py:15 (`ashawkey/diff-gaussian-rasterization`,
 * un-pinned, reference README.md:23-25); it is restated here from SURVEY.md Appendix A.1-A.10.
 * The sub-steps that ARE reference-owned follow the in-tree Python:
 *   SH polynomial/constants ....... reference utils/sh_utils.py:26-43,74-100
 *   +0.5 and clamp_min(0) ......... reference gaussian_renderer/__init__.py:79-83
 *   R(q), Sigma = (R S)(R S)^T .... reference utils/general_utils.py:84-116, scene/gaussian_model.py:37-41
 *   6-vector order xx,xy,xz,yy,yz,zz reference utils/general_utils.py:73-78
 *   homogeneous projection, +1e-7 . reference utils/graphics_utils.py:22-29
 *   row-vector (transposed) matrices reference scene/cameras.py:60-63
 * This file is validated against oracle/torch_oracle.py (autograd, float64) in tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file.  It is never linked into the product.
 *
 * Build twice: -DREAL=float (liboracle_f32.so, mimics kernel rounding) and -DREAL=double
 * (liboracle_f64.so, gradient truth at sizes the torch oracle cannot reach).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif

#define BLK 16
#define NEAR_Z ((REAL)0.2)
#define ALPHA_MIN ((REAL)(1.0 / 255.0))
#define ALPHA_MAX ((REAL)0.99)
#define T_EPS ((REAL)1e-4)
#define DILATION ((REAL)0.3)

static const double C0 = 0.28209479177387814, C1 = 0.4886025119029199;
static const double C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                             -1.0925484305920792, 0.5462742152960396};
static const double C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
                             0.3731763325901154, -0.4570457994644658, 1.445305721320277,
                             -0.5900435899266435};

typedef struct {
    int P, M, D, W, H;
    REAL tanfovx, tanfovy, scale_modifier;
    const REAL *bg, *viewmatrix, *projmatrix, *campos;
    const REAL *means3D, *opacities, *shs, *colors_precomp, *scales, *rotations, *cov3D_precomp;
} scgo_inputs;

typedef struct {
    scgo_inputs in;            /* shallow copy: caller keeps arrays alive until scgo_free */
    int gx, gy;
    /* per Gaussian */
    REAL *depth, *xy, *conic, *rgb, *cov3D;
    int *radii, *rect, *tiles;
    unsigned char *clamped;
    REAL *geom_margin;         /* distance (pixels) of the pre-ceil radius and of the four rect edges to the value at
                                  which the integer decision (ceil / truncation to a tile index) would change */
    /* binning */
    int64_t R;
    uint32_t *point_list;
    int64_t *ranges;           /* [Tn][2] */
    /* per pixel */
    int *n_contrib;
    REAL *final_T;
} scgo_state;

static REAL rmin(REAL a, REAL b) { return a < b ? a : b; }
static REAL rmax(REAL a, REAL b) { return a > b ? a : b; }
static int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* basis values b[16] and gradients db[16][3] at unit direction (x,y,z);
 * reference utils/sh_utils.py:74-100 */
static void sh_basis(int D, double x, double y, double z, double *b, double (*db)[3]) {
    for (int k = 0; k < 16; k++) { b[k] = 0; db[k][0] = db[k][1] = db[k][2] = 0; }
    b[0] = C0;
    if (D < 1) return;
    b[1] = -C1 * y; db[1][1] = -C1;
    b[2] = C1 * z;  db[2][2] = C1;
    b[3] = -C1 * x; db[3][0] = -C1;
    if (D < 2) return;
    double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = C2[0] * xy; db[4][0] = C2[0] * y; db[4][1] = C2[0] * x;
    b[5] = C2[1] * yz; db[5][1] = C2[1] * z; db[5][2] = C2[1] * y;
    b[6] = C2[2] * (2 * zz - xx - yy); db[6][0] = C2[2] * -2 * x; db[6][1] = C2[2] * -2 * y; db[6][2] = C2[2] * 4 * z;
    b[7] = C2[3] * xz; db[7][0] = C2[3] * z; db[7][2] = C2[3] * x;
    b[8] = C2[4] * (xx - yy); db[8][0] = C2[4] * 2 * x; db[8][1] = C2[4] * -2 * y;
    if (D < 3) return;
    b[9] = C3[0] * y * (3 * xx - yy); db[9][0] = C3[0] * 6 * xy; db[9][1] = C3[0] * (3 * xx - 3 * yy);
    b[10] = C3[1] * xy * z; db[10][0] = C3[1] * yz; db[10][1] = C3[1] * xz; db[10][2] = C3[1] * xy;
    b[11] = C3[2] * y * (4 * zz - xx - yy);
    db[11][0] = C3[2] * -2 * xy; db[11][1] = C3[2] * (4 * zz - xx - 3 * yy); db[11][2] = C3[2] * 8 * yz;
    b[12] = C3[3] * z * (2 * zz - 3 * xx - 3 * yy);
    db[12][0] = C3[3] * -6 * xz; db[12][1] = C3[3] * -6 * yz; db[12][2] = C3[3] * (6 * zz - 3 * xx - 3 * yy);
    b[13] = C3[4] * x * (4 * zz - xx - yy);
    db[13][0] = C3[4] * (4 * zz - 3 * xx - yy); db[13][1] = C3[4] * -2 * xy; db[13][2] = C3[4] * 8 * xz;
    b[14] = C3[5] * z * (xx - yy); db[14][0] = C3[5] * 2 * xz; db[14][1] = C3[5] * -2 * yz; db[14][2] = C3[5] * (xx - yy);
    b[15] = C3[6] * x * (xx - 3 * yy); db[15][0] = C3[6] * (3 * xx - 3 * yy); db[15][1] = C3[6] * -6 * xy;
}

static void quat_to_R(const REAL *q, REAL R[9]) {   /* reference utils/general_utils.py:96-104 */
    REAL r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - r * z); R[2] = 2 * (x * z + r * y);
    R[3] = 2 * (x * y + r * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (x * z - r * y); R[7] = 2 * (y * z + r * x); R[8] = 1 - 2 * (x * x + y * y);
}

/* Sigma6 = (R S)(R S)^T, order xx,xy,xz,yy,yz,zz */
static void cov3d_of(const REAL *scale, REAL mod, const REAL *q, REAL *c6) {
    REAL R[9], L[9];
    quat_to_R(q, R);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) L[i * 3 + j] = R[i * 3 + j] * (mod * scale[j]);
    REAL S[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        REAL a = 0; for (int k = 0; k < 3; k++) a += L[i * 3 + k] * L[j * 3 + k]; S[i * 3 + j] = a;
    }
    c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
}

/* Everything the projection of one Gaussian needs, shared by fwd and bwd. */
typedef struct {
    REAL t[3];       /* view-space point (unclamped) */
    REAL tx, ty;     /* clamped */
    int xin, yin;    /* clamp inactive? */
    REAL M0[3], M1[3];
    REAL a, b, c;    /* dilated cov2D */
    REAL fx, fy;
} proj_t;

static void project_cov(const scgo_inputs *in, const REAL *p, const REAL *c6, proj_t *o) {
    const REAL *V = in->viewmatrix;   /* V[k*4+j]: row-vector convention hom @ V */
    for (int j = 0; j < 3; j++) o->t[j] = p[0] * V[0 * 4 + j] + p[1] * V[1 * 4 + j] + p[2] * V[2 * 4 + j] + V[3 * 4 + j];
    o->fx = in->W / (2 * in->tanfovx);
    o->fy = in->H / (2 * in->tanfovy);
    REAL limx = (REAL)1.3 * in->tanfovx, limy = (REAL)1.3 * in->tanfovy;
    REAL tz = o->t[2];
    REAL txtz = o->t[0] / tz, tytz = o->t[1] / tz;
    o->xin = (txtz >= -limx && txtz <= limx);
    o->yin = (tytz >= -limy && tytz <= limy);
    o->tx = rmin(limx, rmax(-limx, txtz)) * tz;
    o->ty = rmin(limy, rmax(-limy, tytz)) * tz;
    REAL J[6] = {o->fx / tz, 0, -o->fx * o->tx / (tz * tz), 0, o->fy / tz, -o->fy * o->ty / (tz * tz)};
    /* W3[r][k] = V[k*4+r] (true W2C rotation); M = J W3 */
    for (int k = 0; k < 3; k++) {
        o->M0[k] = J[0] * V[k * 4 + 0] + J[1] * V[k * 4 + 1] + J[2] * V[k * 4 + 2];
        o->M1[k] = J[3] * V[k * 4 + 0] + J[4] * V[k * 4 + 1] + J[5] * V[k * 4 + 2];
    }
    REAL S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    REAL SM0[3], SM1[3];
    for (int i = 0; i < 3; i++) {
        SM0[i] = S[i * 3] * o->M0[0] + S[i * 3 + 1] * o->M0[1] + S[i * 3 + 2] * o->M0[2];
        SM1[i] = S[i * 3] * o->M1[0] + S[i * 3 + 1] * o->M1[1] + S[i * 3 + 2] * o->M1[2];
    }
    o->a = o->M0[0] * SM0[0] + o->M0[1] * SM0[1] + o->M0[2] * SM0[2] + DILATION;
    o->b = o->M0[0] * SM1[0] + o->M0[1] * SM1[1] + o->M0[2] * SM1[2];
    o->c = o->M1[0] * SM1[0] + o->M1[1] * SM1[1] + o->M1[2] * SM1[2] + DILATION;
}

/* stable LSD radix sort of (key64, val32) on the low `bits` bits */
static void radix_sort64(uint64_t *k, uint32_t *v, int64_t n, int bits) {
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (n > 0 ? n : 1));
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * (n > 0 ? n : 1));
    uint64_t *ks = k, *kd = k2; uint32_t *vs = v, *vd = v2;
    for (int sh = 0; sh < bits; sh += 8) {
        int64_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; i++) cnt[((ks[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; i++) { int64_t o = cnt[(ks[i] >> sh) & 255]++; kd[o] = ks[i]; vd[o] = vs[i]; }
        uint64_t *tk = ks; ks = kd; kd = tk; uint32_t *tv = vs; vs = vd; vd = tv;
    }
    if (ks != k) { memcpy(k, ks, sizeof(uint64_t) * n); memcpy(v, vs, sizeof(uint32_t) * n); }
    free(k2); free(v2);
}

void scgo_free(scgo_state *s) {
    if (!s) return;
    free(s->depth); free(s->xy); free(s->conic); free(s->rgb); free(s->cov3D); free(s->radii);
    free(s->rect); free(s->tiles); free(s->clamped); free(s->point_list); free(s->ranges);
    free(s->n_contrib); free(s->final_T); free(s->geom_margin); free(s);
}

int scgo_real_size(void) { return (int)sizeof(REAL); }
/* number of OpenMP threads of the following calls (a launcher may have exported OMP_NUM_THREADS=1) */
int scgo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
int64_t scgo_num_rendered(const scgo_state *s) { return s->R; }
const uint32_t *scgo_point_list(const scgo_state *s) { return s->point_list; }
const int64_t *scgo_ranges(const scgo_state *s) { return s->ranges; }
const REAL *scgo_means2D(const scgo_state *s) { return s->xy; }
const REAL *scgo_conic(const scgo_state *s) { return s->conic; }
const REAL *scgo_rgb(const scgo_state *s) { return s->rgb; }
const REAL *scgo_depths(const scgo_state *s) { return s->depth; }
const int *scgo_tiles_touched(const scgo_state *s) { return s->tiles; }
const int *scgo_n_contrib(const scgo_state *s) { return s->n_contrib; }

/* Optional per-Gaussian 2D state to use INSTEAD of this file's own preprocess (staged parity: the binning, the blend
 * and the whole backward are then checked on bit-identical 2D parameters, e.g. the ones a CUDA preprocess produced,
 * while the preprocess itself is compared value by value).  radii[i] <= 0 marks a culled Gaussian. */
typedef struct {
    const REAL *xy, *conic, *rgb, *depth;      /* [P,2] [P,3] [P,3] [P] */
    const int *radii;                          /* [P] */
    const unsigned char *clamped;              /* [P,3] */
} scgo_override;

/* A.1-A.8.  out_color[3*H*W], out_depth[H*W], out_alpha[H*W], radii[P]. */
scgo_state *scgo_forward_ex(const scgo_inputs *in, const scgo_override *ov, REAL *out_color, REAL *out_depth,
                            REAL *out_alpha, int *radii_out) {
    const int P = in->P, W = in->W, H = in->H;
    scgo_state *s = (scgo_state *)calloc(1, sizeof(scgo_state));
    s->in = *in;
    s->gx = (W + BLK - 1) / BLK; s->gy = (H + BLK - 1) / BLK;
    const int Tn = s->gx * s->gy;
    const int Pa = P > 0 ? P : 1;
    s->depth = (REAL *)calloc(Pa, sizeof(REAL)); s->xy = (REAL *)calloc(2 * Pa, sizeof(REAL));
    s->conic = (REAL *)calloc(3 * Pa, sizeof(REAL)); s->rgb = (REAL *)calloc(3 * Pa, sizeof(REAL));
    s->cov3D = (REAL *)calloc(6 * Pa, sizeof(REAL)); s->radii = (int *)calloc(Pa, sizeof(int));
    s->rect = (int *)calloc(4 * Pa, sizeof(int)); s->tiles = (int *)calloc(Pa, sizeof(int));
    s->clamped = (unsigned char *)calloc(3 * Pa, 1);
    s->geom_margin = (REAL *)calloc(Pa, sizeof(REAL));
    s->ranges = (int64_t *)calloc(2 * (size_t)Tn, sizeof(int64_t));
    s->n_contrib = (int *)calloc((size_t)W * H, sizeof(int));
    s->final_T = (REAL *)calloc((size_t)W * H, sizeof(REAL));
    memset(out_color, 0, sizeof(REAL) * 3 * W * H);
    memset(out_depth, 0, sizeof(REAL) * W * H);
    memset(out_alpha, 0, sizeof(REAL) * W * H);
    if (P == 0) return s;     /* section 8b: zero images, not background-filled */
    const int use_sh = (in->colors_precomp == NULL);
    const int use_cov = (in->cov3D_precomp != NULL);

#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        const REAL *p = in->means3D + 3 * i;
        const REAL *V = in->viewmatrix, *PM = in->projmatrix;
        const int forced = ov != NULL;
        if (forced && ov->radii[i] <= 0) continue;
        REAL zv = p[0] * V[2] + p[1] * V[6] + p[2] * V[10] + V[14];
        if (!forced && zv <= NEAR_Z) continue;                        /* A.1 */
        REAL hom[4];
        for (int j = 0; j < 4; j++) hom[j] = p[0] * PM[j] + p[1] * PM[4 + j] + p[2] * PM[8 + j] + PM[12 + j];
        REAL pw = 1 / (hom[3] + (REAL)1e-7);
        REAL ndcx = hom[0] * pw, ndcy = hom[1] * pw;                 /* A.2 */
        REAL *c6 = s->cov3D + 6 * i;
        if (use_cov) memcpy(c6, in->cov3D_precomp + 6 * i, 6 * sizeof(REAL));
        else cov3d_of(in->scales + 3 * i, in->scale_modifier, in->rotations + 4 * i, c6);   /* A.3 */
        proj_t pr; project_cov(in, p, c6, &pr);                       /* A.4 */
        REAL det = pr.a * pr.c - pr.b * pr.b;
        if (!forced && det == 0) continue;
        REAL dinv = 1 / det;
        REAL mid = (REAL)0.5 * (pr.a + pr.c);
        REAL disc = (REAL)sqrt((double)rmax((REAL)0.1, mid * mid - det));
        REAL lam = rmax(mid + disc, mid - disc);
        /* in REAL precision, as a float kernel would evaluate it */
        int rad = (int)ceil((double)((REAL)3 * (REAL)sqrt((double)lam)));
        REAL px = ((ndcx + 1) * W - 1) * (REAL)0.5, py = ((ndcy + 1) * H - 1) * (REAL)0.5;
        if (forced) { rad = ov->radii[i]; px = ov->xy[2 * i]; py = ov->xy[2 * i + 1]; }
        int x0 = iclamp((int)((px - rad) / BLK), 0, s->gx), y0 = iclamp((int)((py - rad) / BLK), 0, s->gy);
        int x1 = iclamp((int)((px + rad + BLK - 1) / BLK), 0, s->gx), y1 = iclamp((int)((py + rad + BLK - 1) / BLK), 0, s->gy);
        {   /* margins of the integer decisions above (test infrastructure: scgo_margins) */
            double rv = 3.0 * sqrt((double)lam), gm = fabs(rv - floor(rv + 0.5));
            double e[4] = {(double)px - rad, (double)py - rad, (double)px + rad + BLK - 1, (double)py + rad + BLK - 1};
            for (int q = 0; q < 4; q++) { double f = e[q] / BLK, d = fabs(f - floor(f + 0.5)) * BLK; if (d < gm) gm = d; }
            s->geom_margin[i] = (REAL)gm;
        }
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (forced) {
            for (int ch = 0; ch < 3; ch++) { s->rgb[3 * i + ch] = ov->rgb[3 * i + ch]; s->clamped[3 * i + ch] = ov->clamped[3 * i + ch]; }
        } else if (use_sh) {                                          /* A.5 */
            double dx = p[0] - in->campos[0], dy = p[1] - in->campos[1], dz = p[2] - in->campos[2];
            double n = sqrt(dx * dx + dy * dy + dz * dz);
            double b[16], db[16][3];
            sh_basis(in->D, dx / n, dy / n, dz / n, b, db);
            int nk = (in->D + 1) * (in->D + 1);
            for (int ch = 0; ch < 3; ch++) {
                double acc = 0;
                for (int k = 0; k < nk; k++) acc += b[k] * in->shs[((size_t)i * in->M + k) * 3 + ch];
                REAL v = (REAL)(acc + 0.5);
                s->clamped[3 * i + ch] = v < 0;
                s->rgb[3 * i + ch] = rmax(v, 0);
            }
        } else for (int ch = 0; ch < 3; ch++) s->rgb[3 * i + ch] = in->colors_precomp[3 * i + ch];
        s->depth[i] = forced ? ov->depth[i] : zv; s->radii[i] = rad; s->xy[2 * i] = px; s->xy[2 * i + 1] = py;
        if (forced) { s->conic[3 * i] = ov->conic[3 * i]; s->conic[3 * i + 1] = ov->conic[3 * i + 1]; s->conic[3 * i + 2] = ov->conic[3 * i + 2]; }
        else { s->conic[3 * i] = pr.c * dinv; s->conic[3 * i + 1] = -pr.b * dinv; s->conic[3 * i + 2] = pr.a * dinv; }
        s->rect[4 * i] = x0; s->rect[4 * i + 1] = y0; s->rect[4 * i + 2] = x1; s->rect[4 * i + 3] = y1;
        s->tiles[i] = (x1 - x0) * (y1 - y0);
    }
    if (radii_out) memcpy(radii_out, s->radii, sizeof(int) * P);

    /* A.6: keys in Gaussian-index emission order, stable sort on tile|depth bits */
    int64_t R = 0;
    for (int i = 0; i < P; i++) R += s->tiles[i];
    s->R = R;
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (R > 0 ? R : 1));
    s->point_list = (uint32_t *)malloc(sizeof(uint32_t) * (R > 0 ? R : 1));
    int64_t off = 0;
    for (int i = 0; i < P; i++) {
        if (!s->tiles[i]) continue;
        float df = (float)s->depth[i]; uint32_t dbits; memcpy(&dbits, &df, 4);
        for (int y = s->rect[4 * i + 1]; y < s->rect[4 * i + 3]; y++)
            for (int x = s->rect[4 * i]; x < s->rect[4 * i + 2]; x++) {
                keys[off] = ((uint64_t)(y * s->gx + x) << 32) | dbits;
                s->point_list[off++] = (uint32_t)i;
            }
    }
    int tb = 0; while ((1 << tb) <= Tn) tb++;
    radix_sort64(keys, s->point_list, R, 32 + tb);
    /* A.7 */
    for (int64_t i = 0; i < R; i++) {
        int64_t t = (int64_t)(keys[i] >> 32);
        if (i == 0 || (int64_t)(keys[i - 1] >> 32) != t) s->ranges[2 * t] = i;
        if (i == R - 1 || (int64_t)(keys[i + 1] >> 32) != t) s->ranges[2 * t + 1] = i + 1;
    }
    free(keys);

    /* A.8 */
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < Tn; t++) {
        int tx = t % s->gx, ty = t / s->gx;
        int64_t r0 = s->ranges[2 * t], r1 = s->ranges[2 * t + 1];
        for (int ly = 0; ly < BLK; ly++) for (int lx = 0; lx < BLK; lx++) {
            int px = tx * BLK + lx, py = ty * BLK + ly;
            if (px >= W || py >= H) continue;
            REAL T = 1, C[3] = {0, 0, 0}, Dsum = 0, Wsum = 0; int last = 0, cnt = 0;
            for (int64_t j = r0; j < r1; j++) {
                cnt++;
                uint32_t g = s->point_list[j];
                REAL dx = s->xy[2 * g] - px, dy = s->xy[2 * g + 1] - py;
                const REAL *co = s->conic + 3 * g;
                REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0) continue;
                REAL alpha = rmin(ALPHA_MAX, in->opacities[g] * (REAL)exp((double)power));
                if (alpha < ALPHA_MIN) continue;
                REAL test_T = T * (1 - alpha);
                if (test_T < T_EPS) break;
                REAL w = alpha * T;
                for (int ch = 0; ch < 3; ch++) C[ch] += s->rgb[3 * g + ch] * w;
                Dsum += s->depth[g] * w; Wsum += w;
                T = test_T; last = cnt;
            }
            size_t pid = (size_t)py * W + px;
            for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * W * H + pid] = C[ch] + T * in->bg[ch];
            out_depth[pid] = Dsum; out_alpha[pid] = Wsum;
            s->n_contrib[pid] = last; s->final_T[pid] = T;
        }
    }
    return s;
}

scgo_state *scgo_forward(const scgo_inputs *in, REAL *out_color, REAL *out_depth, REAL *out_alpha, int *radii_out) {
    return scgo_forward_ex(in, NULL, out_color, out_depth, out_alpha, radii_out);
}

/* A.9 + A.10.  Upstream: dL_dcolor[3*H*W], dL_ddepth[H*W], dL_dalpha[H*W].
 * Outputs (all fully written): dmeans3D[P*3], dmeans2D[P*3] (NDC units, z = 0), dsh[P*M*3],
 * dcolors[P*3], dopac[P], dscales[P*3], drots[P*4], dcov3D[P*6].  Any may be NULL. */
void scgo_backward(const scgo_state *s, const REAL *gC, const REAL *gD, const REAL *gA,
                   REAL *dmeans3D, REAL *dmeans2D, REAL *dsh, REAL *dcolors, REAL *dopac,
                   REAL *dscales, REAL *drots, REAL *dcov3D) {
    const scgo_inputs *in = &s->in;
    const int P = in->P, W = in->W, H = in->H, Tn = s->gx * s->gy;
    if (dmeans3D) memset(dmeans3D, 0, sizeof(REAL) * 3 * P);
    if (dmeans2D) memset(dmeans2D, 0, sizeof(REAL) * 3 * P);
    if (dsh) memset(dsh, 0, sizeof(REAL) * 3 * (size_t)in->M * P);
    if (dcolors) memset(dcolors, 0, sizeof(REAL) * 3 * P);
    if (dopac) memset(dopac, 0, sizeof(REAL) * P);
    if (dscales) memset(dscales, 0, sizeof(REAL) * 3 * P);
    if (drots) memset(drots, 0, sizeof(REAL) * 4 * P);
    if (dcov3D) memset(dcov3D, 0, sizeof(REAL) * 6 * P);
    if (P == 0) return;
    /* per-Gaussian screen-space accumulators: mean2D(2, pixel units) conic(3, true dL/dB)
     * opacity(1) rgb(3) depth(1) */
    double *acc = (double *)calloc((size_t)P * 10, sizeof(double));

#pragma omp parallel
    {
        double *loc = NULL; int64_t loc_cap = 0;
#pragma omp for schedule(dynamic, 4)
        for (int t = 0; t < Tn; t++) {
            int tx = t % s->gx, ty = t / s->gx;
            int64_t r0 = s->ranges[2 * t], r1 = s->ranges[2 * t + 1], n = r1 - r0;
            if (n <= 0) continue;
            if (n > loc_cap) { free(loc); loc_cap = n * 2; loc = (double *)malloc(sizeof(double) * 10 * loc_cap); }
            memset(loc, 0, sizeof(double) * 10 * n);
            for (int ly = 0; ly < BLK; ly++) for (int lx = 0; lx < BLK; lx++) {
                int px = tx * BLK + lx, py = ty * BLK + ly;
                if (px >= W || py >= H) continue;
                size_t pid = (size_t)py * W + px;
                const REAL T_final = s->final_T[pid];
                REAL T = T_final;
                REAL g3[3] = {gC[pid], gC[(size_t)W * H + pid], gC[2 * (size_t)W * H + pid]};
                REAL gd = gD[pid], ga = gA[pid];
                REAL bgdot = in->bg[0] * g3[0] + in->bg[1] * g3[1] + in->bg[2] * g3[2];
                REAL arec[3] = {0, 0, 0}, drec = 0, alrec = 0, last_alpha = 0, last_c[3] = {0, 0, 0}, last_d = 0;
                for (int64_t j = r0 + s->n_contrib[pid] - 1; j >= r0; j--) {
                    uint32_t g = s->point_list[j];
                    REAL dx = s->xy[2 * g] - px, dy = s->xy[2 * g + 1] - py;
                    const REAL *co = s->conic + 3 * g;
                    REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0) continue;
                    REAL G = (REAL)exp((double)power);
                    REAL o = in->opacities[g];
                    REAL alpha = rmin(ALPHA_MAX, o * G);
                    if (alpha < ALPHA_MIN) continue;
                    T = T / (1 - alpha);
                    REAL w = alpha * T;
                    REAL dL_dalpha = 0;
                    double *L = loc + 10 * (j - r0);
                    for (int ch = 0; ch < 3; ch++) {
                        REAL c = s->rgb[3 * g + ch];
                        arec[ch] = last_alpha * last_c[ch] + (1 - last_alpha) * arec[ch];
                        last_c[ch] = c;
                        dL_dalpha += (c - arec[ch]) * g3[ch];
                        L[6 + ch] += w * g3[ch];
                    }
                    REAL dpt = s->depth[g];
                    drec = last_alpha * last_d + (1 - last_alpha) * drec; last_d = dpt;
                    dL_dalpha += (dpt - drec) * gd;
                    L[9] += w * gd;
                    alrec = last_alpha + (1 - last_alpha) * alrec;
                    dL_dalpha += (1 - alrec) * ga;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1 - alpha)) * bgdot;
                    REAL dL_dG = o * dL_dalpha;       /* propagated even when alpha was capped (A.9) */
                    REAL gdx = G * dx, gdy = G * dy;
                    L[0] += dL_dG * (-gdx * co[0] - gdy * co[1]);
                    L[1] += dL_dG * (-gdy * co[2] - gdx * co[1]);
                    L[2] += (REAL)-0.5 * gdx * dx * dL_dG;
                    L[3] += -gdx * dy * dL_dG;          /* true dL/dB (twice the external's slot) */
                    L[4] += (REAL)-0.5 * gdy * dy * dL_dG;
                    L[5] += G * dL_dalpha;
                }
            }
            for (int64_t j = 0; j < n; j++) {
                uint32_t g = s->point_list[r0 + j];
                for (int k = 0; k < 10; k++) if (loc[10 * j + k] != 0) {
#pragma omp atomic
                    acc[(size_t)g * 10 + k] += loc[10 * j + k];
                }
            }
        }
        free(loc);
    }

    const int use_sh = (in->colors_precomp == NULL);
    const int use_cov = (in->cov3D_precomp != NULL);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (s->radii[i] <= 0) continue;
        const double *A = acc + (size_t)i * 10;
        const REAL *p = in->means3D + 3 * i;
        const REAL *V = in->viewmatrix, *PM = in->projmatrix;
        REAL dmean[3] = {0, 0, 0};
        REAL dm2x = (REAL)(A[0] * 0.5 * W), dm2y = (REAL)(A[1] * 0.5 * H);   /* NDC units (A.9) */
        if (dmeans2D) { dmeans2D[3 * i] = dm2x; dmeans2D[3 * i + 1] = dm2y; }
        if (dopac) dopac[i] = (REAL)A[5];
        /* (1) conic -> cov2D -> Sigma, J -> t -> mean */
        const REAL *c6 = s->cov3D + 6 * i;
        proj_t pr; project_cov(in, p, c6, &pr);
        REAL a = pr.a, b = pr.b, c = pr.c;
        REAL den = a * c - b * b;
        REAL k = 1 / (den * den + (REAL)1e-7);
        REAL dA = (REAL)A[2], dB = (REAL)A[3], dC = (REAL)A[4];
        REAL dLa = k * (-c * c * dA + b * c * dB - b * b * dC);
        REAL dLb = k * (2 * b * c * dA - (den + 2 * b * b) * dB + 2 * a * b * dC);
        REAL dLc = k * (-b * b * dA + a * b * dB - a * a * dC);
        const REAL *M0 = pr.M0, *M1 = pr.M1;
        REAL d6[6];
        d6[0] = dLa * M0[0] * M0[0] + dLb * M0[0] * M1[0] + dLc * M1[0] * M1[0];
        d6[3] = dLa * M0[1] * M0[1] + dLb * M0[1] * M1[1] + dLc * M1[1] * M1[1];
        d6[5] = dLa * M0[2] * M0[2] + dLb * M0[2] * M1[2] + dLc * M1[2] * M1[2];
        d6[1] = 2 * dLa * M0[0] * M0[1] + dLb * (M0[0] * M1[1] + M0[1] * M1[0]) + 2 * dLc * M1[0] * M1[1];
        d6[2] = 2 * dLa * M0[0] * M0[2] + dLb * (M0[0] * M1[2] + M0[2] * M1[0]) + 2 * dLc * M1[0] * M1[2];
        d6[4] = 2 * dLa * M0[1] * M0[2] + dLb * (M0[1] * M1[2] + M0[2] * M1[1]) + 2 * dLc * M1[1] * M1[2];
        if (dcov3D && use_cov) memcpy(dcov3D + 6 * i, d6, sizeof(d6));
        REAL S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
        REAL SM0[3], SM1[3], dM0[3], dM1[3];
        for (int r = 0; r < 3; r++) {
            SM0[r] = S[r * 3] * M0[0] + S[r * 3 + 1] * M0[1] + S[r * 3 + 2] * M0[2];
            SM1[r] = S[r * 3] * M1[0] + S[r * 3 + 1] * M1[1] + S[r * 3 + 2] * M1[2];
        }
        for (int r = 0; r < 3; r++) { dM0[r] = 2 * dLa * SM0[r] + dLb * SM1[r]; dM1[r] = dLb * SM0[r] + 2 * dLc * SM1[r]; }
        /* dJ[r][kk] = sum_i dM[r][i] * W3[kk][i], W3[kk][i] = V[i*4+kk] */
        REAL dJ00 = dM0[0] * V[0] + dM0[1] * V[4] + dM0[2] * V[8];
        REAL dJ02 = dM0[0] * V[2] + dM0[1] * V[6] + dM0[2] * V[10];
        REAL dJ11 = dM1[0] * V[1] + dM1[1] * V[5] + dM1[2] * V[9];
        REAL dJ12 = dM1[0] * V[2] + dM1[1] * V[6] + dM1[2] * V[10];
        REAL tz = pr.t[2], tzi = 1 / tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        REAL dtx = pr.xin ? -pr.fx * tz2 * dJ02 : 0;
        REAL dty = pr.yin ? -pr.fy * tz2 * dJ12 : 0;
        REAL dtz = -pr.fx * tz2 * dJ00 - pr.fy * tz2 * dJ11 + 2 * pr.fx * pr.tx * tz3 * dJ02 + 2 * pr.fy * pr.ty * tz3 * dJ12;
        for (int kk = 0; kk < 3; kk++) dmean[kk] += V[kk * 4 + 0] * dtx + V[kk * 4 + 1] * dty + V[kk * 4 + 2] * dtz;
        /* (2) NDC mean -> mean3D */
        REAL hom[4];
        for (int j = 0; j < 4; j++) hom[j] = p[0] * PM[j] + p[1] * PM[4 + j] + p[2] * PM[8 + j] + PM[12 + j];
        REAL pw = 1 / (hom[3] + (REAL)1e-7);
        for (int kk = 0; kk < 3; kk++)
            dmean[kk] += dm2x * (PM[kk * 4 + 0] * pw - hom[0] * pw * pw * PM[kk * 4 + 3])
                       + dm2y * (PM[kk * 4 + 1] * pw - hom[1] * pw * pw * PM[kk * 4 + 3]);
        /* (3) depth -> mean3D */
        for (int kk = 0; kk < 3; kk++) dmean[kk] += V[kk * 4 + 2] * (REAL)A[9];
        /* (4) colour */
        if (use_sh) {
            double vx = p[0] - in->campos[0], vy = p[1] - in->campos[1], vz = p[2] - in->campos[2];
            double n = sqrt(vx * vx + vy * vy + vz * vz);
            double dir[3] = {vx / n, vy / n, vz / n};
            double bb[16], db[16][3];
            sh_basis(in->D, dir[0], dir[1], dir[2], bb, db);
            int nk = (in->D + 1) * (in->D + 1);
            double ddir[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ch++) {
                double g = s->clamped[3 * i + ch] ? 0.0 : A[6 + ch];
                for (int kq = 0; kq < nk; kq++) {
                    size_t idx = ((size_t)i * in->M + kq) * 3 + ch;
                    if (dsh) dsh[idx] = (REAL)(bb[kq] * g);
                    for (int ax = 0; ax < 3; ax++) ddir[ax] += g * db[kq][ax] * in->shs[idx];
                }
            }
            double dot = dir[0] * ddir[0] + dir[1] * ddir[1] + dir[2] * ddir[2];
            for (int ax = 0; ax < 3; ax++) dmean[ax] += (REAL)((ddir[ax] - dir[ax] * dot) / n);
        } else if (dcolors) for (int ch = 0; ch < 3; ch++) dcolors[3 * i + ch] = (REAL)A[6 + ch];
        /* (5) Sigma -> scale, rotation */
        if (!use_cov) {
            const REAL *sc = in->scales + 3 * i, *q = in->rotations + 4 * i;
            REAL R[9]; quat_to_R(q, R);
            REAL sp[3] = {in->scale_modifier * sc[0], in->scale_modifier * sc[1], in->scale_modifier * sc[2]};
            REAL Sg[9] = {2 * d6[0], d6[1], d6[2], d6[1], 2 * d6[3], d6[4], d6[2], d6[4], 2 * d6[5]};
            REAL dLm[9], L[9];
            for (int r = 0; r < 3; r++) for (int j = 0; j < 3; j++) L[r * 3 + j] = R[r * 3 + j] * sp[j];
            for (int r = 0; r < 3; r++) for (int j = 0; j < 3; j++)
                dLm[r * 3 + j] = Sg[r * 3] * L[j] + Sg[r * 3 + 1] * L[3 + j] + Sg[r * 3 + 2] * L[6 + j];
            REAL Dm[9];
            for (int j = 0; j < 3; j++) {
                REAL ds = dLm[j] * R[j] + dLm[3 + j] * R[3 + j] + dLm[6 + j] * R[6 + j];
                if (dscales) dscales[3 * i + j] = in->scale_modifier * ds;
                for (int r = 0; r < 3; r++) Dm[r * 3 + j] = dLm[r * 3 + j] * sp[j];
            }
            REAL r = q[0], x = q[1], y = q[2], z = q[3];
            if (drots) {
                drots[4 * i + 0] = 2 * (-z * Dm[1] + y * Dm[2] + z * Dm[3] - x * Dm[5] - y * Dm[6] + x * Dm[7]);
                drots[4 * i + 1] = 2 * (y * Dm[1] + z * Dm[2] + y * Dm[3] - 2 * x * Dm[4] - r * Dm[5] + z * Dm[6] + r * Dm[7] - 2 * x * Dm[8]);
                drots[4 * i + 2] = 2 * (-2 * y * Dm[0] + x * Dm[1] + r * Dm[2] + x * Dm[3] + z * Dm[5] - r * Dm[6] + z * Dm[7] - 2 * y * Dm[8]);
                drots[4 * i + 3] = 2 * (-2 * z * Dm[0] - r * Dm[1] + x * Dm[2] + r * Dm[3] - 2 * z * Dm[4] + y * Dm[5] + x * Dm[6] + y * Dm[7]);
            }
        }
        if (dmeans3D) for (int kk = 0; kk < 3; kk++) dmeans3D[3 * i + kk] = dmean[kk];
    }
    free(acc);
}

/* ------------------------------------------------------------------------------------------------------------
 * Margins of the discrete decisions (parity-test support).  The rasterizer takes four kinds of yes/no decisions
 * whose outcome a last-bit difference in fp32 arithmetic can flip:
 *   (a) alpha = opacity * exp(power) >= 1/255           (A.8 skip)
 *   (b) test_T = T (1 - alpha) >= 1e-4                  (A.8 early stop)
 *   (c) power <= 0                                      (A.8 skip)
 *   (d) radius = ceil(3 sqrt(lambda)), tile rect = trunc((pix -/+ radius) / 16)
 * How close is "close enough to flip" follows from an ERROR MODEL of two fp32 implementations of the same formulas
 * (this file vs a CUDA kernel with FMA contraction and hardware exp2), stated by three constants:
 *   pos_err    absolute uncertainty of a projected mean, pixels   (a few ulp of the largest pixel coordinate)
 *   conic_err  relative uncertainty of the conic coefficients (hence of power)
 *   base_err   relative uncertainty of exp() and of the products around it
 * For one (Gaussian, pixel) entry the relative uncertainty of alpha is
 *   u = base_err + conic_err |power| + pos_err (|A dx + B dy| + |B dx + C dy|)          (|grad power| . pos_err)
 * decision (a) is uncertain when |255 alpha - 1| < u; the relative uncertainty of T accumulates along the walk,
 * U += u alpha / (1 - alpha) per blended entry, and decision (b) is uncertain when |test_T / 1e-4 - 1| < U + u alpha / (1 - alpha) + base_err;
 * (c) when |power| < conic_err |power| + pos_err |grad power| + 1e-12 for an entry that would contribute; (d) when the
 * pre-ceil radius or a rect edge lies within geom_err pixels of the value where the integer changes.
 * A pixel is FLIP-PRONE when some entry it visits (the terminating one included) has an uncertain decision, or when
 * a Gaussian uncertain in (d) could reach it.  A Gaussian is FLIP-AFFECTED when it (nearly) contributes to a
 * flip-prone pixel -- every gradient of such a Gaussian changes with the flip, through T and the suffix blend -- or is
 * uncertain in (d); it is flagged OWN (value 2) when the uncertain decision is its own.
 * pix_margin[H*W]: min over the decisions of the pixel of margin / uncertainty (< 1 <=> flip-prone): by how much the
 * error model would have to be scaled for the pixel to change class (calibration);  gauss_margin[P]: min of
 * pix_margin over the pixels the Gaussian (nearly) contributes to. */
void scgo_margins(const scgo_state *s, double base_err, double conic_err, double pos_err, double geom_err,
                  REAL *pix_margin, unsigned char *pix_flag, REAL *gauss_margin, unsigned char *gauss_flag) {
    const scgo_inputs *in = &s->in;
    const int P = in->P, W = in->W, H = in->H, Tn = s->gx * s->gy;
    const size_t N = (size_t)W * H;
    for (size_t i = 0; i < N; i++) pix_margin[i] = (REAL)1e30;
    for (int i = 0; i < P; i++) gauss_margin[i] = (REAL)1e30;
    memset(pix_flag, 0, N);
    memset(gauss_flag, 0, P > 0 ? P : 0);
    if (P == 0) return;
    unsigned char *own = (unsigned char *)calloc(P, 1);
    /* pass 1: per-pixel minimum of margin / uncertainty over the entries the forward visits */
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < Tn; t++) {
        int tx = t % s->gx, ty = t / s->gx;
        int64_t r0 = s->ranges[2 * t], r1 = s->ranges[2 * t + 1];
        for (int ly = 0; ly < BLK; ly++) for (int lx = 0; lx < BLK; lx++) {
            int px = tx * BLK + lx, py = ty * BLK + ly;
            if (px >= W || py >= H) continue;
            double m = 1e30, U = 0;
            REAL T = 1;
            for (int64_t j = r0; j < r1; j++) {
                uint32_t g = s->point_list[j];
                REAL dx = s->xy[2 * g] - px, dy = s->xy[2 * g + 1] - py;
                const REAL *co = s->conic + 3 * g;
                REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                double gp = fabs((double)co[0] * dx + (double)co[1] * dy) + fabs((double)co[1] * dx + (double)co[2] * dy);
                double up = conic_err * fabs((double)power) + pos_err * gp;       /* absolute uncertainty of power */
                double u = base_err + up;
                double araw = (double)in->opacities[g] * exp((double)power);
                double mine = 1e30;
                if (araw * 255.0 >= 1.0 - u) { double q = fabs((double)power) / (up + 1e-12); if (q < mine) mine = q; }   /* (c) */
                if (power <= 0) {
                    double q = fabs(araw * 255.0 - 1.0) / u;                                                            /* (a) */
                    if (q < mine) mine = q;
                }
                if (mine < 1.0) own[g] = 1;     /* benign race: all writers store 1 */
                if (mine < m) m = mine;
                if (power > 0) continue;
                REAL alpha = rmin(ALPHA_MAX, (REAL)araw);
                if (alpha < ALPHA_MIN) continue;
                REAL test_T = T * (1 - alpha);
                double ua = (araw < (double)ALPHA_MAX) ? u * alpha / (1 - alpha) : 0.0;     /* the cap is exact */
                double q = fabs((double)test_T / 1e-4 - 1.0) / (U + ua + base_err);                                      /* (b) */
                if (q < 1.0) own[g] = 1;
                if (q < m) m = q;
                if (test_T < T_EPS) break;
                T = test_T; U += ua;
            }
            size_t pid = (size_t)py * W + px;
            pix_margin[pid] = (REAL)m;
            if (m < 1.0) pix_flag[pid] = 1;
        }
    }
    /* (d): Gaussians whose radius / rect could differ by one: every pixel of the rect grown by one tile that they
     * could reach is flip-prone, and they are flip-affected themselves */
    for (int i = 0; i < P; i++) {
        if (s->radii[i] <= 0 || (double)s->geom_margin[i] >= geom_err) continue;
        own[i] = 1;
        int x0 = iclamp(s->rect[4 * i] - 1, 0, s->gx) * BLK, y0 = iclamp(s->rect[4 * i + 1] - 1, 0, s->gy) * BLK;
        int x1 = iclamp(s->rect[4 * i + 2] + 1, 0, s->gx) * BLK, y1 = iclamp(s->rect[4 * i + 3] + 1, 0, s->gy) * BLK;
        if (x1 > W) x1 = W;
        if (y1 > H) y1 = H;
        const REAL *co = s->conic + 3 * i;
        for (int py = y0; py < y1; py++) for (int px = x0; px < x1; px++) {
            REAL dx = s->xy[2 * i] - px, dy = s->xy[2 * i + 1] - py;
            REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if ((double)in->opacities[i] * exp((double)power) * 255.0 >= 0.99) {
                pix_flag[(size_t)py * W + px] = 1;
                pix_margin[(size_t)py * W + px] = (REAL)((double)s->geom_margin[i] / geom_err);
            }
        }
    }
    /* pass 2: per Gaussian, the smallest pix_margin among the pixels it (nearly) contributes to.  The whole list is
     * walked: a flipped early stop lets entries behind the oracle's terminating one contribute. */
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < Tn; t++) {
        int tx = t % s->gx, ty = t / s->gx;
        int64_t r0 = s->ranges[2 * t], r1 = s->ranges[2 * t + 1], n = r1 - r0;
        if (n <= 0) continue;
        REAL *loc = (REAL *)malloc(sizeof(REAL) * n);
        for (int64_t j = 0; j < n; j++) loc[j] = (REAL)1e30;
        for (int ly = 0; ly < BLK; ly++) for (int lx = 0; lx < BLK; lx++) {
            int px = tx * BLK + lx, py = ty * BLK + ly;
            if (px >= W || py >= H) continue;
            REAL pm = pix_margin[(size_t)py * W + px];
            if (pm > (REAL)64) continue;          /* (margins beyond 64x the error model are not tracked per Gaussian) */
            for (int64_t j = r0; j < r1; j++) {
                if (loc[j - r0] <= pm) continue;
                uint32_t g = s->point_list[j];
                REAL dx = s->xy[2 * g] - px, dy = s->xy[2 * g + 1] - py;
                const REAL *co = s->conic + 3 * g;
                REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if ((double)in->opacities[g] * exp((double)power) * 255.0 >= 0.99) loc[j - r0] = pm;
            }
        }
        for (int64_t j = 0; j < n; j++) if (loc[j] < (REAL)1e30) {
            uint32_t g = s->point_list[r0 + j];
#pragma omp critical(scgo_gm)
            { if (loc[j] < gauss_margin[g]) gauss_margin[g] = loc[j]; }
        }
        free(loc);
    }
    for (int i = 0; i < P; i++) {
        if (s->radii[i] > 0 && (double)s->geom_margin[i] < geom_err) {
            REAL q = (REAL)((double)s->geom_margin[i] / geom_err);
            if (q < gauss_margin[i]) gauss_margin[i] = q;
        }
        gauss_flag[i] = own[i] ? 2 : (gauss_margin[i] < (REAL)1 ? 1 : 0);
    }
    free(own);
}
const REAL *scgo_geom_margin(const scgo_state *s) { return s->geom_margin; }
