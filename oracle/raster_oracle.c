/*
 * ORACLE (test infrastructure, never a product path).
 *
 * Plain-C restatement of the tile rasteriser behind
 * diff_gaussian_rasterization.GaussianRasterizer, as called by Free-SurGS
 * (reference gaussian_renderer/__init__.py:68,69,131; settings scene/pose_optimizer.py:619-632).
 *
 * PARITY UNPINNED: the rasteriser source is a third-party dependency that is absent from
 * /root/reference (requirements.txt:26 -> ingra14m/depth-diff-gaussian-rasterization @ HEAD, no
 * pin; .gitmodules:4-6 submodule directory missing).  This file restates the published
 * algorithm (kernels K1..K9) including its closed-form backward with the
 * upstream regulariser 1/(det^2 + 1e-7); it is validated against the float64 torch.autograd
 * oracle (oracle/raster_oracle.py) by tests/test_oracle_c.py.
 *
 * Built twice from this one source: -DORACLE_DOUBLE (truth at scale) and float (the
 * same-precision "port" that bench.py times as the CPU baseline).  OpenMP over Gaussians/tiles.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_DOUBLE
typedef double real;
#define FN(name) name##_f64
#define R_EXP exp
#define R_SQRT sqrt
#define R_CEIL ceil
#else
typedef float real;
#define FN(name) name##_f32
#define R_EXP expf
#define R_SQRT sqrtf
#define R_CEIL ceilf
#endif

#define BLOCK 16
#define NCH 3

static const real SH_C0 = (real)0.28209479177387814;
static const real SH_C1 = (real)0.4886025119029199;
static const real SH_C2[5] = {(real)1.0925484305920792, (real)-1.0925484305920792, (real)0.31539156525252005,
                              (real)-1.0925484305920792, (real)0.5462742152960396};
static const real SH_C3[7] = {(real)-0.5900435899266435, (real)2.890611442640554, (real)-0.4570457994644658,
                              (real)0.3731763325901154, (real)-0.4570457994644658, (real)1.445305721320277,
                              (real)-0.5900435899266435};

typedef struct {
    /* borrowed inputs */
    int P, sh_deg, n_coeffs, W, H;
    const real *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp;
    real scale_modifier, tanfovx, tanfovy;
    real view[16], proj[16], campos[3], bg[3];
    /* geometry state (K1) */
    real *xy, *depth, *conic, *cov3D, *rgb;
    int *radii, *rect;           /* rect: minx,miny,maxx,maxy */
    unsigned char *clamped;      /* 3 per Gaussian */
    /* binning state (K2-K5) */
    int64_t R;
    uint32_t *point_list;
    int64_t *ranges;             /* 2 per tile */
    /* image state (K6) */
    real *final_T;
    int *n_contrib;
} Ctx;

static void xf43(const real *p, const real *M, real *o) {
    o[0] = M[0] * p[0] + M[4] * p[1] + M[8] * p[2] + M[12];
    o[1] = M[1] * p[0] + M[5] * p[1] + M[9] * p[2] + M[13];
    o[2] = M[2] * p[0] + M[6] * p[1] + M[10] * p[2] + M[14];
}
static void xf44(const real *p, const real *M, real *o) {
    xf43(p, M, o);
    o[3] = M[3] * p[0] + M[7] * p[1] + M[11] * p[2] + M[15];
}

/* R (row-major, math convention) from an un-normalised quaternion (w,x,y,z) used as given */
static void quat_R(const real *q, real *R) {
    real r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - r * z); R[2] = 2 * (x * z + r * y);
    R[3] = 2 * (x * y + r * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (x * z - r * y); R[7] = 2 * (y * z + r * x); R[8] = 1 - 2 * (x * x + y * y);
}

/* Sigma = R diag(s^2) R^T, six upper-triangular entries (xx,xy,xz,yy,yz,zz) */
static void cov3d_from_sr(const real *scale, real mod, const real *q, real *c6) {
    real R[9], s2[3];
    quat_R(q, R);
    for (int k = 0; k < 3; ++k) { real s = mod * scale[k]; s2[k] = s * s; }
    real S[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            real acc = 0;
            for (int k = 0; k < 3; ++k) acc += R[i * 3 + k] * s2[k] * R[j * 3 + k];
            S[i * 3 + j] = acc;
        }
    c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
}

/* shared by K1 and K8: the clamped view-space point, M = J*W3 (2x3) and a,b,c */
static void cov2d_terms(const Ctx *c, const real *mean, const real *c6, real *t, real *M, real *abc,
                        real *xmask, real *ymask) {
    const real *V = c->view;
    real fx = c->W / (2 * c->tanfovx), fy = c->H / (2 * c->tanfovy);
    xf43(mean, V, t);
    real limx = (real)1.3 * c->tanfovx, limy = (real)1.3 * c->tanfovy;
    real txtz = t[0] / t[2], tytz = t[1] / t[2];
    *xmask = (txtz < -limx || txtz > limx) ? 0 : 1;
    *ymask = (tytz < -limy || tytz > limy) ? 0 : 1;
    t[0] = fmin(limx, fmax(-limx, txtz)) * t[2];
    t[1] = fmin(limy, fmax(-limy, tytz)) * t[2];
    real J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]);
    real J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
    /* W3[r][col] = V[col*4 + r] */
    for (int j = 0; j < 3; ++j) {
        M[j] = J00 * V[4 * j + 0] + J02 * V[4 * j + 2];
        M[3 + j] = J11 * V[4 * j + 1] + J12 * V[4 * j + 2];
    }
    real S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    real SM0[3], SM1[3];
    for (int i = 0; i < 3; ++i) {
        SM0[i] = S[i * 3] * M[0] + S[i * 3 + 1] * M[1] + S[i * 3 + 2] * M[2];
        SM1[i] = S[i * 3] * M[3] + S[i * 3 + 1] * M[4] + S[i * 3 + 2] * M[5];
    }
    abc[0] = M[0] * SM0[0] + M[1] * SM0[1] + M[2] * SM0[2] + (real)0.3;
    abc[1] = M[0] * SM1[0] + M[1] * SM1[1] + M[2] * SM1[2];
    abc[2] = M[3] * SM1[0] + M[4] * SM1[1] + M[5] * SM1[2] + (real)0.3;
}

static void sh_basis(int deg, const real *d, real *B) {
    real x = d[0], y = d[1], z = d[2];
    B[0] = SH_C0;
    if (deg > 0) {
        B[1] = -SH_C1 * y; B[2] = SH_C1 * z; B[3] = -SH_C1 * x;
        if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            B[4] = SH_C2[0] * xy; B[5] = SH_C2[1] * yz; B[6] = SH_C2[2] * (2 * zz - xx - yy);
            B[7] = SH_C2[3] * xz; B[8] = SH_C2[4] * (xx - yy);
            if (deg > 2) {
                B[9] = SH_C3[0] * y * (3 * xx - yy); B[10] = SH_C3[1] * xy * z;
                B[11] = SH_C3[2] * y * (4 * zz - xx - yy); B[12] = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy);
                B[13] = SH_C3[4] * x * (4 * zz - xx - yy); B[14] = SH_C3[5] * z * (xx - yy);
                B[15] = SH_C3[6] * x * (xx - 3 * yy);
            }
        }
    }
}

/* d(basis)/d(dir) : dB[k][3] */
static void sh_basis_grad(int deg, const real *d, real dB[16][3]) {
    real x = d[0], y = d[1], z = d[2];
    memset(dB, 0, sizeof(real) * 48);
    if (deg > 0) {
        dB[1][1] = -SH_C1; dB[2][2] = SH_C1; dB[3][0] = -SH_C1;
        if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            dB[4][0] = SH_C2[0] * y; dB[4][1] = SH_C2[0] * x;
            dB[5][1] = SH_C2[1] * z; dB[5][2] = SH_C2[1] * y;
            dB[6][0] = SH_C2[2] * -2 * x; dB[6][1] = SH_C2[2] * -2 * y; dB[6][2] = SH_C2[2] * 4 * z;
            dB[7][0] = SH_C2[3] * z; dB[7][2] = SH_C2[3] * x;
            dB[8][0] = SH_C2[4] * 2 * x; dB[8][1] = SH_C2[4] * -2 * y;
            if (deg > 2) {
                dB[9][0] = SH_C3[0] * 6 * xy; dB[9][1] = SH_C3[0] * (3 * xx - 3 * yy);
                dB[10][0] = SH_C3[1] * yz; dB[10][1] = SH_C3[1] * xz; dB[10][2] = SH_C3[1] * xy;
                dB[11][0] = SH_C3[2] * -2 * xy; dB[11][1] = SH_C3[2] * (4 * zz - xx - 3 * yy); dB[11][2] = SH_C3[2] * 8 * yz;
                dB[12][0] = SH_C3[3] * -6 * xz; dB[12][1] = SH_C3[3] * -6 * yz; dB[12][2] = SH_C3[3] * (6 * zz - 3 * xx - 3 * yy);
                dB[13][0] = SH_C3[4] * (4 * zz - 3 * xx - yy); dB[13][1] = SH_C3[4] * -2 * xy; dB[13][2] = SH_C3[4] * 8 * xz;
                dB[14][0] = SH_C3[5] * 2 * xz; dB[14][1] = SH_C3[5] * -2 * yz; dB[14][2] = SH_C3[5] * (xx - yy);
                dB[15][0] = SH_C3[6] * (3 * xx - 3 * yy); dB[15][1] = SH_C3[6] * -6 * xy;
            }
        }
    }
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void radix_sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int bits) {
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    for (int shift = 0; shift < bits; shift += 8) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; ++i) cnt[((keys[i] >> shift) & 0xFF) + 1]++;
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (int64_t i = 0; i < n; ++i) {
            int64_t d = cnt[(keys[i] >> shift) & 0xFF]++;
            k2[d] = keys[i]; v2[d] = vals[i];
        }
        uint64_t *tk = keys; keys = k2; k2 = tk;
        uint32_t *tv = vals; vals = v2; v2 = tv;
    }
    /* passes = ceil(bits/8); if odd the result lives in the scratch arrays */
    if (((bits + 7) / 8) & 1) {
        memcpy(k2, keys, sizeof(uint64_t) * (size_t)n);
        memcpy(v2, vals, sizeof(uint32_t) * (size_t)n);
        free(keys); free(vals);
    } else {
        free(k2); free(v2);
    }
}

void FN(fsgs_oracle_free)(void *vctx) {
    Ctx *c = (Ctx *)vctx;
    if (!c) return;
    free(c->xy); free(c->depth); free(c->conic); free(c->cov3D); free(c->rgb); free(c->radii);
    free(c->rect); free(c->clamped); free(c->point_list); free(c->ranges); free(c->final_T);
    free(c->n_contrib); free(c);
}

/* K1..K6.  Returns an opaque context that the backward consumes. */
void *FN(fsgs_oracle_forward)(int P, int sh_deg, int n_coeffs, const real *means3D, const real *shs,
                              const real *colors_precomp, const real *opacities, const real *scales,
                              real scale_modifier, const real *rotations, const real *cov3D_precomp,
                              const real *viewmatrix, const real *projmatrix, const real *campos, int W,
                              int H, real tanfovx, real tanfovy, const real *bg, real *out_color,
                              real *out_depth, int *radii_out, int64_t *num_rendered, real *pix_margin,
                              const float *sort_depth, real margin_kappa) {
    /* margin_kappa: the threshold margins written to pix_margin are relative distances divided by
       (1 + margin_kappa * g), g = |d power / d centre| in 1/pixel (for the T test: the alpha-weighted sum of g over the
       entries in front).  A splat centre is a float32 PIXEL coordinate, so any float32 implementation carries an
       absolute error of a few ulp(W) in it, which moves log(alpha) by g times that -- steep (small) splats are
       fragile over a wider band than flat ones.  0 = plain relative distances. */
    /* sort_depth (optional, P floats): float32 view depths to take the SORT KEYS from instead of rounding this
       build's own depths.  The list order of near-equal depths is decided by float32 rounding and therefore differs
       between float32 implementations; a test that checks compositing against this oracle passes the depths of the
       implementation under test here (and checks separately that they are within float32 rounding of the truth),
       so the two composite every tile in the same order.  Only the keys change, never a composited value. */
    Ctx *c = (Ctx *)calloc(1, sizeof(Ctx));
    c->P = P; c->sh_deg = sh_deg; c->n_coeffs = n_coeffs; c->W = W; c->H = H;
    c->means3D = means3D; c->shs = shs; c->colors_precomp = colors_precomp; c->opacities = opacities;
    c->scales = scales; c->rotations = rotations; c->cov3D_precomp = cov3D_precomp;
    c->scale_modifier = scale_modifier; c->tanfovx = tanfovx; c->tanfovy = tanfovy;
    memcpy(c->view, viewmatrix, sizeof(real) * 16);
    memcpy(c->proj, projmatrix, sizeof(real) * 16);
    memcpy(c->campos, campos, sizeof(real) * 3);
    memcpy(c->bg, bg, sizeof(real) * 3);
    const int gx = (W + BLOCK - 1) / BLOCK, gy = (H + BLOCK - 1) / BLOCK;
    const int64_t ntiles = (int64_t)gx * gy;
    size_t Pn = (size_t)(P > 0 ? P : 1);
    c->xy = (real *)calloc(Pn * 2, sizeof(real));
    c->depth = (real *)calloc(Pn, sizeof(real));
    c->conic = (real *)calloc(Pn * 3, sizeof(real));
    c->cov3D = (real *)calloc(Pn * 6, sizeof(real));
    c->rgb = (real *)calloc(Pn * 3, sizeof(real));
    c->radii = (int *)calloc(Pn, sizeof(int));
    c->rect = (int *)calloc(Pn * 4, sizeof(int));
    c->clamped = (unsigned char *)calloc(Pn * 3, 1);
    int64_t *offsets = (int64_t *)calloc(Pn + 1, sizeof(int64_t));

    /* ---- K1 preprocess ---- */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const real *m = means3D + 3 * i;
        real pv[3], ph[4];
        xf43(m, c->view, pv);
        if (pv[2] <= (real)0.2) continue;
        xf44(m, c->proj, ph);
        real pw = 1 / (ph[3] + (real)0.0000001);
        real pproj[2] = {ph[0] * pw, ph[1] * pw};
        real *c6 = c->cov3D + 6 * i;
        if (cov3D_precomp) memcpy(c6, cov3D_precomp + 6 * i, sizeof(real) * 6);
        else cov3d_from_sr(scales + 3 * i, scale_modifier, rotations + 4 * i, c6);
        real t[3], M[6], abc[3], xm, ym;
        cov2d_terms(c, m, c6, t, M, abc, &xm, &ym);
        real det = abc[0] * abc[2] - abc[1] * abc[1];
        if (det == 0) continue;
        real det_inv = 1 / det;
        real conic[3] = {abc[2] * det_inv, -abc[1] * det_inv, abc[0] * det_inv};
        real mid = (real)0.5 * (abc[0] + abc[2]);
        real lam1 = mid + R_SQRT(fmax((real)0.1, mid * mid - det));
        real lam2 = mid - R_SQRT(fmax((real)0.1, mid * mid - det));
        real my_radius = R_CEIL(3 * R_SQRT(fmax(lam1, lam2)));
        real px = ((pproj[0] + 1) * W - 1) * (real)0.5, py = ((pproj[1] + 1) * H - 1) * (real)0.5;
        int rminx = clampi((int)((px - my_radius) / BLOCK), 0, gx);
        int rminy = clampi((int)((py - my_radius) / BLOCK), 0, gy);
        int rmaxx = clampi((int)((px + my_radius + BLOCK - 1) / BLOCK), 0, gx);
        int rmaxy = clampi((int)((py + my_radius + BLOCK - 1) / BLOCK), 0, gy);
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue;
        if (shs) {
            real dir[3] = {m[0] - c->campos[0], m[1] - c->campos[1], m[2] - c->campos[2]};
            real inv = 1 / R_SQRT(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
            dir[0] *= inv; dir[1] *= inv; dir[2] *= inv;
            real B[16];
            sh_basis(sh_deg, dir, B);
            int nb = (sh_deg + 1) * (sh_deg + 1);
            const real *sh = shs + (size_t)i * n_coeffs * 3;
            for (int ch = 0; ch < 3; ++ch) {
                real v = 0;
                for (int k = 0; k < nb; ++k) v += B[k] * sh[k * 3 + ch];
                v += (real)0.5;
                c->clamped[3 * i + ch] = v < 0;
                c->rgb[3 * i + ch] = v < 0 ? 0 : v;
            }
        } else {
            for (int ch = 0; ch < 3; ++ch) c->rgb[3 * i + ch] = colors_precomp[3 * i + ch];
        }
        c->depth[i] = pv[2];
        c->radii[i] = (int)my_radius;
        c->xy[2 * i] = px; c->xy[2 * i + 1] = py;
        c->conic[3 * i] = conic[0]; c->conic[3 * i + 1] = conic[1]; c->conic[3 * i + 2] = conic[2];
        c->rect[4 * i] = rminx; c->rect[4 * i + 1] = rminy; c->rect[4 * i + 2] = rmaxx; c->rect[4 * i + 3] = rmaxy;
        offsets[i + 1] = (int64_t)(rmaxx - rminx) * (rmaxy - rminy);
    }
    for (int i = 0; i < P; ++i) radii_out[i] = c->radii[i];

    /* ---- K2 scan, K3 duplicate with keys ---- */
    for (int i = 0; i < P; ++i) offsets[i + 1] += offsets[i];
    int64_t R = offsets[P];
    c->R = R;
    *num_rendered = R;
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(R > 0 ? R : 1));
    c->point_list = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        if (c->radii[i] <= 0) continue;
        int64_t off = offsets[i];
        /* (a NaN entry = no override for that Gaussian) */
        float df = (sort_depth && sort_depth[i] == sort_depth[i]) ? sort_depth[i] : (float)c->depth[i];
        uint32_t dbits;
        memcpy(&dbits, &df, 4);
        for (int y = c->rect[4 * i + 1]; y < c->rect[4 * i + 3]; ++y)
            for (int x = c->rect[4 * i]; x < c->rect[4 * i + 2]; ++x) {
                keys[off] = ((uint64_t)(y * gx + x) << 32) | dbits;
                c->point_list[off] = (uint32_t)i;
                off++;
            }
    }
    /* ---- K4 stable LSD radix sort on (tile | depth bits), K5 ranges ---- */
    int tile_bits = 0;
    while ((1LL << tile_bits) < ntiles) tile_bits++;
    radix_sort_pairs(keys, c->point_list, R, 32 + tile_bits + 1);
    c->ranges = (int64_t *)calloc((size_t)ntiles * 2, sizeof(int64_t));
    for (int64_t k = 0; k < R; ++k) {
        int64_t tile = (int64_t)(keys[k] >> 32);
        if (k == 0 || (int64_t)(keys[k - 1] >> 32) != tile) c->ranges[2 * tile] = k;
        if (k == R - 1 || (int64_t)(keys[k + 1] >> 32) != tile) c->ranges[2 * tile + 1] = k + 1;
    }
    free(keys);
    free(offsets);

    /* ---- K6 composite ---- */
    c->final_T = (real *)malloc(sizeof(real) * (size_t)W * H);
    c->n_contrib = (int *)malloc(sizeof(int) * (size_t)W * H);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        int tx0 = (int)(tile % gx) * BLOCK, ty0 = (int)(tile / gx) * BLOCK;
        int64_t s = c->ranges[2 * tile], e = c->ranges[2 * tile + 1];
        for (int py = ty0; py < ty0 + BLOCK && py < H; ++py)
            for (int px = tx0; px < tx0 + BLOCK && px < W; ++px) {
                real T = 1, C[NCH] = {0, 0, 0}, D = 0;
                real margin = (real)1e30;   /* relative distance of the closest threshold decision */
                /* relative depth gap of the closest pair of consecutive entries that both reach alpha >= 1/255
                   here: a float32 implementation whose depths round differently may composite them in the other order */
                real order_margin = (real)1e30, prev_depth = (real)-1, gsum = 0;
                int contributor = 0, last = 0;
                for (int64_t k = s; k < e; ++k) {
                    contributor++;
                    uint32_t id = c->point_list[k];
                    real dx = c->xy[2 * id] - (real)px, dy = c->xy[2 * id + 1] - (real)py;
                    const real *con = c->conic + 3 * id;
                    real power = (real)-0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy;
                    if (fabs(power) < (real)1e-12) margin = 0;
                    if (power > 0) continue;
                    real alpha = fmin((real)0.99, opacities[id] * R_EXP(power));
                    const real g = fabs(con[0] * dx + con[1] * dy) + fabs(con[2] * dy + con[1] * dx);
                    margin = fmin(margin, fabs(alpha - (real)(1.0 / 255.0)) * 255 / (1 + margin_kappa * g));
                    if (alpha < (real)(1.0 / 255.0)) continue;
                    if (prev_depth >= 0) order_margin = fmin(order_margin, (c->depth[id] - prev_depth) / c->depth[id]);
                    prev_depth = c->depth[id];
                    real test_T = T * (1 - alpha);
                    gsum += g * alpha / (1 - alpha);
                    margin = fmin(margin, fabs(test_T - (real)0.0001) * 10000 / (1 + margin_kappa * gsum));
                    if (test_T < (real)0.0001) break;
                    for (int ch = 0; ch < NCH; ++ch) C[ch] += c->rgb[3 * id + ch] * alpha * T;
                    D += c->depth[id] * alpha * T;
                    T = test_T;
                    last = contributor;
                }
                int64_t pix = (int64_t)py * W + px;
                c->final_T[pix] = T;
                c->n_contrib[pix] = last;
                if (pix_margin) { pix_margin[pix] = margin; pix_margin[(int64_t)W * H + pix] = order_margin; }
                for (int ch = 0; ch < NCH; ++ch) out_color[(int64_t)ch * H * W + pix] = C[ch] + T * c->bg[ch];
                out_depth[pix] = D;
            }
    }
    return c;
}

static void atomic_add(real *p, real v) {
#pragma omp atomic
    *p += v;
}

/* K7..K9.  All outputs are overwritten. dL_dmeans2D is [P,3] (z stays 0). */
void FN(fsgs_oracle_backward)(void *vctx, const real *dL_dcolor, const real *dL_ddepth_img,
                              real *dL_dmeans2D, real *dL_dcolors, real *dL_dopacity, real *dL_dmeans3D,
                              real *dL_dcov3D, real *dL_dsh, real *dL_dscales, real *dL_drots) {
    Ctx *c = (Ctx *)vctx;
    const int P = c->P, W = c->W, H = c->H;
    const int gx = (W + BLOCK - 1) / BLOCK, gy = (H + BLOCK - 1) / BLOCK;
    const int64_t ntiles = (int64_t)gx * gy;
    size_t Pn = (size_t)(P > 0 ? P : 1);
    memset(dL_dmeans2D, 0, sizeof(real) * Pn * 3);
    memset(dL_dcolors, 0, sizeof(real) * Pn * 3);
    memset(dL_dopacity, 0, sizeof(real) * Pn);
    memset(dL_dmeans3D, 0, sizeof(real) * Pn * 3);
    memset(dL_dcov3D, 0, sizeof(real) * Pn * 6);
    if (dL_dsh) memset(dL_dsh, 0, sizeof(real) * Pn * (size_t)c->n_coeffs * 3);
    memset(dL_dscales, 0, sizeof(real) * Pn * 3);
    memset(dL_drots, 0, sizeof(real) * Pn * 4);
    real *dL_dconic = (real *)calloc(Pn * 3, sizeof(real));
    real *dL_ddepth = (real *)calloc(Pn, sizeof(real));

    /* ---- K7 composite backward ---- */
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        int tx0 = (int)(tile % gx) * BLOCK, ty0 = (int)(tile / gx) * BLOCK;
        int64_t s = c->ranges[2 * tile];
        for (int py = ty0; py < ty0 + BLOCK && py < H; ++py)
            for (int px = tx0; px < tx0 + BLOCK && px < W; ++px) {
                int64_t pix = (int64_t)py * W + px;
                const real T_final = c->final_T[pix];
                real T = T_final;
                const int last = c->n_contrib[pix];
                real accum_rec[NCH] = {0, 0, 0}, last_color[NCH] = {0, 0, 0}, dpix[NCH];
                real accum_drec = 0, last_depth = 0, last_alpha = 0;
                for (int ch = 0; ch < NCH; ++ch) dpix[ch] = dL_dcolor[(int64_t)ch * H * W + pix];
                const real dpix_d = dL_ddepth_img ? dL_ddepth_img[pix] : 0;
                const real ddelx_dx = (real)0.5 * W, ddely_dy = (real)0.5 * H;
                for (int64_t k = s + last - 1; k >= s; --k) {
                    uint32_t id = c->point_list[k];
                    real dx = c->xy[2 * id] - (real)px, dy = c->xy[2 * id + 1] - (real)py;
                    const real *con = c->conic + 3 * id;
                    real o = c->opacities[id];
                    real power = (real)-0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy;
                    if (power > 0) continue;
                    real G = R_EXP(power);
                    real alpha = fmin((real)0.99, o * G);
                    if (alpha < (real)(1.0 / 255.0)) continue;
                    T = T / (1 - alpha);
                    real dchannel_dcolor = alpha * T;
                    real dL_dalpha = 0;
                    for (int ch = 0; ch < NCH; ++ch) {
                        real col = c->rgb[3 * id + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1 - last_alpha) * accum_rec[ch];
                        last_color[ch] = col;
                        dL_dalpha += (col - accum_rec[ch]) * dpix[ch];
                        atomic_add(&dL_dcolors[3 * id + ch], dchannel_dcolor * dpix[ch]);
                    }
                    {
                        real dep = c->depth[id];
                        accum_drec = last_alpha * last_depth + (1 - last_alpha) * accum_drec;
                        last_depth = dep;
                        dL_dalpha += (dep - accum_drec) * dpix_d;
                        atomic_add(&dL_ddepth[id], dchannel_dcolor * dpix_d);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    real bg_dot = 0;
                    for (int ch = 0; ch < NCH; ++ch) bg_dot += c->bg[ch] * dpix[ch];
                    dL_dalpha += (-T_final / (1 - alpha)) * bg_dot;
                    real dL_dG = o * dL_dalpha;
                    real gdx = G * dx, gdy = G * dy;
                    real dG_ddelx = -gdx * con[0] - gdy * con[1];
                    real dG_ddely = -gdy * con[2] - gdx * con[1];
                    atomic_add(&dL_dmeans2D[3 * id], dL_dG * dG_ddelx * ddelx_dx);
                    atomic_add(&dL_dmeans2D[3 * id + 1], dL_dG * dG_ddely * ddely_dy);
                    atomic_add(&dL_dconic[3 * id], (real)-0.5 * gdx * dx * dL_dG);
                    atomic_add(&dL_dconic[3 * id + 1], (real)-0.5 * gdx * dy * dL_dG);
                    atomic_add(&dL_dconic[3 * id + 2], (real)-0.5 * gdy * dy * dL_dG);
                    atomic_add(&dL_dopacity[id], G * dL_dalpha);
                }
            }
    }

    /* ---- K8 cov2D backward + K9 preprocess backward ---- */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        if (c->radii[i] <= 0) continue;
        const real *m = c->means3D + 3 * i;
        const real *c6 = c->cov3D + 6 * i;
        const real *V = c->view, *PM = c->proj;
        real fx = W / (2 * c->tanfovx), fy = H / (2 * c->tanfovy);
        real t[3], M[6], abc[3], xm, ym;
        cov2d_terms(c, m, c6, t, M, abc, &xm, &ym);
        real a = abc[0], b = abc[1], cc = abc[2];
        real denom = a * cc - b * b;
        real d2inv = 1 / (denom * denom + (real)0.0000001);
        real gxx = dL_dconic[3 * i], gxy = dL_dconic[3 * i + 1], gzz = dL_dconic[3 * i + 2];
        real dL_da = 0, dL_db = 0, dL_dc = 0;
        real *dcv = dL_dcov3D + 6 * i;
        if (d2inv != 0) {
            dL_da = d2inv * (-cc * cc * gxx + 2 * b * cc * gxy + (denom - a * cc) * gzz);
            dL_dc = d2inv * (-a * a * gzz + 2 * a * b * gxy + (denom - a * cc) * gxx);
            dL_db = d2inv * 2 * (b * cc * gxx - (denom + 2 * b * b) * gxy + a * b * gzz);
            /* dSigma (six stored entries; off-diagonals carry both symmetric halves) */
            dcv[0] = M[0] * M[0] * dL_da + M[0] * M[3] * dL_db + M[3] * M[3] * dL_dc;
            dcv[3] = M[1] * M[1] * dL_da + M[1] * M[4] * dL_db + M[4] * M[4] * dL_dc;
            dcv[5] = M[2] * M[2] * dL_da + M[2] * M[5] * dL_db + M[5] * M[5] * dL_dc;
            dcv[1] = 2 * M[0] * M[1] * dL_da + (M[0] * M[4] + M[1] * M[3]) * dL_db + 2 * M[3] * M[4] * dL_dc;
            dcv[2] = 2 * M[0] * M[2] * dL_da + (M[0] * M[5] + M[2] * M[3]) * dL_db + 2 * M[3] * M[5] * dL_dc;
            dcv[4] = 2 * M[2] * M[1] * dL_da + (M[1] * M[5] + M[2] * M[4]) * dL_db + 2 * M[4] * M[5] * dL_dc;
        }
        /* dL/dM (2x3): a = M0 S M0^T, b = M0 S M1^T, c = M1 S M1^T */
        real S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
        real SM0[3], SM1[3], dM[6];
        for (int j = 0; j < 3; ++j) {
            SM0[j] = S[j * 3] * M[0] + S[j * 3 + 1] * M[1] + S[j * 3 + 2] * M[2];
            SM1[j] = S[j * 3] * M[3] + S[j * 3 + 1] * M[4] + S[j * 3 + 2] * M[5];
        }
        for (int j = 0; j < 3; ++j) {
            dM[j] = 2 * SM0[j] * dL_da + SM1[j] * dL_db;
            dM[3 + j] = 2 * SM1[j] * dL_dc + SM0[j] * dL_db;
        }
        /* M = J W3: M[0][j] = J00 R[0][j] + J02 R[2][j];  M[1][j] = J11 R[1][j] + J12 R[2][j] */
        real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int j = 0; j < 3; ++j) {
            dJ00 += V[4 * j + 0] * dM[j];
            dJ02 += V[4 * j + 2] * dM[j];
            dJ11 += V[4 * j + 1] * dM[3 + j];
            dJ12 += V[4 * j + 2] * dM[3 + j];
        }
        real tz = 1 / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        real dtx = xm * -fx * tz2 * dJ02;
        real dty = ym * -fy * tz2 * dJ12;
        real dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
        real dmean[3] = {V[0] * dtx + V[1] * dty + V[2] * dtz, V[4] * dtx + V[5] * dty + V[6] * dtz,
                         V[8] * dtx + V[9] * dty + V[10] * dtz};
        /* K9: projected mean */
        real mh[4];
        xf44(m, PM, mh);
        real mw = 1 / (mh[3] + (real)0.0000001);
        real mul1 = (PM[0] * m[0] + PM[4] * m[1] + PM[8] * m[2] + PM[12]) * mw * mw;
        real mul2 = (PM[1] * m[0] + PM[5] * m[1] + PM[9] * m[2] + PM[13]) * mw * mw;
        real g2x = dL_dmeans2D[3 * i], g2y = dL_dmeans2D[3 * i + 1];
        dmean[0] += (PM[0] * mw - PM[3] * mul1) * g2x + (PM[1] * mw - PM[3] * mul2) * g2y;
        dmean[1] += (PM[4] * mw - PM[7] * mul1) * g2x + (PM[5] * mw - PM[7] * mul2) * g2y;
        dmean[2] += (PM[8] * mw - PM[11] * mul1) * g2x + (PM[9] * mw - PM[11] * mul2) * g2y;
        /* depth variant: d(p_view.z)/d(mean) */
        dmean[0] += V[2] * dL_ddepth[i]; dmean[1] += V[6] * dL_ddepth[i]; dmean[2] += V[10] * dL_ddepth[i];
        /* SH colour */
        if (c->shs) {
            real dir0[3] = {m[0] - c->campos[0], m[1] - c->campos[1], m[2] - c->campos[2]};
            real len2 = dir0[0] * dir0[0] + dir0[1] * dir0[1] + dir0[2] * dir0[2];
            real inv = 1 / R_SQRT(len2);
            real dir[3] = {dir0[0] * inv, dir0[1] * inv, dir0[2] * inv};
            real B[16], dB[16][3];
            sh_basis(c->sh_deg, dir, B);
            sh_basis_grad(c->sh_deg, dir, dB);
            int nb = (c->sh_deg + 1) * (c->sh_deg + 1);
            const real *sh = c->shs + (size_t)i * c->n_coeffs * 3;
            real *dsh = dL_dsh + (size_t)i * c->n_coeffs * 3;
            real ddir[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ++ch) {
                real g = c->clamped[3 * i + ch] ? 0 : dL_dcolors[3 * i + ch];
                for (int k = 0; k < nb; ++k) {
                    dsh[k * 3 + ch] = B[k] * g;
                    for (int ax = 0; ax < 3; ++ax) ddir[ax] += dB[k][ax] * sh[k * 3 + ch] * g;
                }
            }
            /* through v/|v| */
            real dot = dir[0] * ddir[0] + dir[1] * ddir[1] + dir[2] * ddir[2];
            for (int ax = 0; ax < 3; ++ax) dmean[ax] += (ddir[ax] - dir[ax] * dot) * inv;
        }
        for (int ax = 0; ax < 3; ++ax) dL_dmeans3D[3 * i + ax] = dmean[ax];
        /* Sigma -> scale, quaternion */
        if (!c->cov3D_precomp) {
            real Rm[9];
            const real *q = c->rotations + 4 * i;
            quat_R(q, Rm);
            real s[3] = {c->scale_modifier * c->scales[3 * i], c->scale_modifier * c->scales[3 * i + 1],
                         c->scale_modifier * c->scales[3 * i + 2]};
            real Gs[9] = {dcv[0], (real)0.5 * dcv[1], (real)0.5 * dcv[2], (real)0.5 * dcv[1], dcv[3],
                          (real)0.5 * dcv[4], (real)0.5 * dcv[2], (real)0.5 * dcv[4], dcv[5]};
            real GR[9];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k)
                    GR[r * 3 + k] = Gs[r * 3] * Rm[k] + Gs[r * 3 + 1] * Rm[3 + k] + Gs[r * 3 + 2] * Rm[6 + k];
            real dR[9];
            for (int k = 0; k < 3; ++k) {
                real rgr = Rm[k] * GR[k] + Rm[3 + k] * GR[3 + k] + Rm[6 + k] * GR[6 + k];
                dL_dscales[3 * i + k] = 2 * s[k] * rgr * c->scale_modifier;
                for (int r = 0; r < 3; ++r) dR[r * 3 + k] = 2 * GR[r * 3 + k] * s[k] * s[k];
            }
            real r = q[0], x = q[1], y = q[2], z = q[3];
            dL_drots[4 * i + 0] = 2 * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
            dL_drots[4 * i + 1] = 2 * (y * dR[1] + z * dR[2] + y * dR[3] - 2 * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2 * x * dR[8]);
            dL_drots[4 * i + 2] = 2 * (-2 * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2 * y * dR[8]);
            dL_drots[4 * i + 3] = 2 * (-2 * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2 * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
        }
    }
    free(dL_dconic);
    free(dL_ddepth);
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU arm asks for all host threads explicitly */
void FN(fsgs_oracle_set_num_threads)(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int FN(fsgs_oracle_num_threads)(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
