// fsgs_math.cuh -- per-Gaussian arithmetic of the rasteriser (projection, EWA covariance, SH,
// and their closed-form backward), written as host/device inline functions so that the CUDA
// kernels use them on the GPU and tests/cpu_emul can compile the very same code with g++ to
// check it against the oracle in the build container (which has no GPU).
//
// Behavioural spec: SURVEY.md Appendix A (kernels K1, K8, K9), i.e. the arithmetic of the
// `diff_gaussian_rasterization` package that Free-SurGS calls at
// gaussian_renderer/__init__.py:68,69,131, plus the Python pre-processing of
// scene/gaussian_model.py:118-138,260-333, utils/sh_utils.py:57-112 and
// scene/pose_optimizer.py:960-989 that the fused path folds in.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FSGS_HD __host__ __device__ __forceinline__
#else
#define FSGS_HD inline
#endif

namespace fsgs {

// Pinned float32 operations: never contracted or re-associated by the compiler.  The forward projection is written
// with them because it is inlined into several kernels (k_preprocess_fused, k_preprocess_frozen, k_preprocess_api)
// that must produce the SAME bits from the same inputs, and ptxas picks which multiply of `a*b + c*d` joins the add
// per kernel (measured: 4 % of the conics differed in the last bit between two inlinings of the same source).
#if defined(__CUDA_ARCH__)
#define FSGS_MUL(a, b) __fmul_rn((a), (b))
#define FSGS_ADD(a, b) __fadd_rn((a), (b))
#define FSGS_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define FSGS_MUL(a, b) ((a) * (b))
#define FSGS_ADD(a, b) ((a) + (b))
#define FSGS_FMA(a, b, c) fmaf((a), (b), (c))
#endif

constexpr int TILE = 16;
constexpr float NEAR_CULL = 0.2f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_MIN = 0.0001f;
constexpr float COV_DILATION = 0.3f;

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2_0 = 1.0925484305920792f, SH_C2_1 = -1.0925484305920792f, SH_C2_2 = 0.31539156525252005f,
                SH_C2_3 = -1.0925484305920792f, SH_C2_4 = 0.5462742152960396f;
constexpr float SH_C3_0 = -0.5900435899266435f, SH_C3_1 = 2.890611442640554f, SH_C3_2 = -0.4570457994644658f,
                SH_C3_3 = 0.3731763325901154f, SH_C3_4 = -0.4570457994644658f, SH_C3_5 = 1.445305721320277f,
                SH_C3_6 = -0.5900435899266435f;

// Camera constants one launch needs, derived on the host from fsgs_settings.
struct CamConst {
    int W, H, gx, gy;
    float fx, fy;          // focal lengths in pixels: W/(2 tanfovx), H/(2 tanfovy)
    float limx, limy;      // 1.3 * tanfov
    float mod;             // scale modifier
    int sh_deg, n_coeffs;
};

// a0 b0 + a1 b1 + a2 b2 (+ c), fixed evaluation order
FSGS_HD float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return FSGS_FMA(a2, b2, FSGS_FMA(a1, b1, FSGS_MUL(a0, b0)));
}
FSGS_HD float dot3p(float a0, float b0, float a1, float b1, float a2, float b2, float c) {
    return FSGS_ADD(dot3(a0, b0, a1, b1, a2, b2), c);
}
// Result of projecting one Gaussian (K1).
struct Splat {
    float px, py;          // pixel-space centre
    float depth;           // view-space z
    float a, b, c;         // 2D covariance (with the +0.3 dilation)
    float conx, cony, conz;
    int radius;            // ceil(3 sqrt(lambda_max)), the reference's screen radius
    int rminx, rminy, rmaxx, rmaxy;   // the reference's tile rectangle [min,max)
};

FSGS_HD void xf43(const float *M, float x, float y, float z, float &ox, float &oy, float &oz) {
    ox = dot3p(M[0], x, M[4], y, M[8], z, M[12]);
    oy = dot3p(M[1], x, M[5], y, M[9], z, M[13]);
    oz = dot3p(M[2], x, M[6], y, M[10], z, M[14]);
}
// camera-frame mean: rows of the row-major 3x4 pose times (xyz, 1)   (transform_to_frame)
FSGS_HD void pose_mean(const float *pose, const float *xyz, float *mean) {
    for (int r = 0; r < 3; ++r)
        mean[r] = dot3p(pose[4 * r], xyz[0], pose[4 * r + 1], xyz[1], pose[4 * r + 2], xyz[2], pose[4 * r + 3]);
}

// Rotation matrix (row-major) of a quaternion (w,x,y,z) used as given.
FSGS_HD void quat_to_R(const float *q, float *R) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    // (pinned evaluation order, see FSGS_MUL above: this feeds Sigma_3D, which two kernels must agree on bit for bit)
    const float xx = FSGS_MUL(x, x), yy = FSGS_MUL(y, y), zz = FSGS_MUL(z, z);
    const float xy = FSGS_MUL(x, y), xz = FSGS_MUL(x, z), yz = FSGS_MUL(y, z);
    R[0] = FSGS_FMA(-2.f, FSGS_ADD(yy, zz), 1.f); R[1] = FSGS_MUL(2.f, FSGS_FMA(-r, z, xy)); R[2] = FSGS_MUL(2.f, FSGS_FMA(r, y, xz));
    R[3] = FSGS_MUL(2.f, FSGS_FMA(r, z, xy)); R[4] = FSGS_FMA(-2.f, FSGS_ADD(xx, zz), 1.f); R[5] = FSGS_MUL(2.f, FSGS_FMA(-r, x, yz));
    R[6] = FSGS_MUL(2.f, FSGS_FMA(-r, y, xz)); R[7] = FSGS_MUL(2.f, FSGS_FMA(r, x, yz)); R[8] = FSGS_FMA(-2.f, FSGS_ADD(xx, yy), 1.f);
}

// Sigma = R diag(s^2) R^T as (xx,xy,xz,yy,yz,zz); s already includes the scale modifier.
FSGS_HD void cov3d_from_scale_rot(const float *s, const float *q, float *c6) {
    float R[9];
    quat_to_R(q, R);
    const float s0 = FSGS_MUL(s[0], s[0]), s1 = FSGS_MUL(s[1], s[1]), s2 = FSGS_MUL(s[2], s[2]);
    const float a0 = FSGS_MUL(R[0], s0), a1 = FSGS_MUL(R[1], s1), a2 = FSGS_MUL(R[2], s2);
    const float b0 = FSGS_MUL(R[3], s0), b1 = FSGS_MUL(R[4], s1), b2 = FSGS_MUL(R[5], s2);
    const float d0 = FSGS_MUL(R[6], s0), d1 = FSGS_MUL(R[7], s1), d2 = FSGS_MUL(R[8], s2);
    c6[0] = dot3(R[0], a0, R[1], a1, R[2], a2);
    c6[1] = dot3(R[3], a0, R[4], a1, R[5], a2);
    c6[2] = dot3(R[6], a0, R[7], a1, R[8], a2);
    c6[3] = dot3(R[3], b0, R[4], b1, R[5], b2);
    c6[4] = dot3(R[6], b0, R[7], b1, R[8], b2);
    c6[5] = dot3(R[6], d0, R[7], d1, R[8], d2);
}

// View-space point with the reference's +-1.3 tanfov clamp, and M = J * W3 (2x3, row-major),
// W3[r][c] = V[4c + r].  xmask/ymask are 0 where the clamp is active.
FSGS_HD void ewa_M(const CamConst &cc, const float *V, const float *mean, float *t, float *M, float &xmask,
                   float &ymask) {
    xf43(V, mean[0], mean[1], mean[2], t[0], t[1], t[2]);
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    xmask = (txtz < -cc.limx || txtz > cc.limx) ? 0.f : 1.f;
    ymask = (tytz < -cc.limy || tytz > cc.limy) ? 0.f : 1.f;
    t[0] = FSGS_MUL(fminf(cc.limx, fmaxf(-cc.limx, txtz)), t[2]);
    t[1] = FSGS_MUL(fminf(cc.limy, fmaxf(-cc.limy, tytz)), t[2]);
    const float tz2 = FSGS_MUL(t[2], t[2]);
    const float J00 = cc.fx / t[2], J02 = -FSGS_MUL(cc.fx, t[0]) / tz2;
    const float J11 = cc.fy / t[2], J12 = -FSGS_MUL(cc.fy, t[1]) / tz2;
    M[0] = FSGS_FMA(J02, V[2], FSGS_MUL(J00, V[0])); M[1] = FSGS_FMA(J02, V[6], FSGS_MUL(J00, V[4]));
    M[2] = FSGS_FMA(J02, V[10], FSGS_MUL(J00, V[8]));
    M[3] = FSGS_FMA(J12, V[2], FSGS_MUL(J11, V[1])); M[4] = FSGS_FMA(J12, V[6], FSGS_MUL(J11, V[5]));
    M[5] = FSGS_FMA(J12, V[10], FSGS_MUL(J11, V[9]));
}

// (S M0, S M1) and a,b,c of M S M^T + 0.3 I.
FSGS_HD void ewa_abc(const float *M, const float *c6, float *SM0, float *SM1, float &a, float &b, float &c) {
    SM0[0] = dot3(c6[0], M[0], c6[1], M[1], c6[2], M[2]);
    SM0[1] = dot3(c6[1], M[0], c6[3], M[1], c6[4], M[2]);
    SM0[2] = dot3(c6[2], M[0], c6[4], M[1], c6[5], M[2]);
    SM1[0] = dot3(c6[0], M[3], c6[1], M[4], c6[2], M[5]);
    SM1[1] = dot3(c6[1], M[3], c6[3], M[4], c6[4], M[5]);
    SM1[2] = dot3(c6[2], M[3], c6[4], M[4], c6[5], M[5]);
    a = dot3p(M[0], SM0[0], M[1], SM0[1], M[2], SM0[2], COV_DILATION);
    b = dot3(M[0], SM1[0], M[1], SM1[1], M[2], SM1[2]);
    c = dot3p(M[3], SM1[0], M[4], SM1[1], M[5], SM1[2], COV_DILATION);
}

FSGS_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// K1 geometry of one Gaussian.  Returns false when the reference would skip it (radius 0).
FSGS_HD bool project_gaussian(const CamConst &cc, const float *V, const float *PM, const float *mean,
                              const float *c6, Splat &o) {
    float vx, vy, vz;
    xf43(V, mean[0], mean[1], mean[2], vx, vy, vz);
    if (vz <= NEAR_CULL) return false;
    const float hx = dot3p(PM[0], mean[0], PM[4], mean[1], PM[8], mean[2], PM[12]);
    const float hy = dot3p(PM[1], mean[0], PM[5], mean[1], PM[9], mean[2], PM[13]);
    const float hw = dot3p(PM[3], mean[0], PM[7], mean[1], PM[11], mean[2], PM[15]);
    const float pw = 1.0f / FSGS_ADD(hw, 0.0000001f);
    float t[3], M[6], xm, ym, SM0[3], SM1[3];
    ewa_M(cc, V, mean, t, M, xm, ym);
    ewa_abc(M, c6, SM0, SM1, o.a, o.b, o.c);
    const float det = FSGS_FMA(o.a, o.c, -FSGS_MUL(o.b, o.b));
    if (det == 0.0f) return false;
    const float det_inv = 1.f / det;
    o.conx = FSGS_MUL(o.c, det_inv); o.cony = FSGS_MUL(-o.b, det_inv); o.conz = FSGS_MUL(o.a, det_inv);
    const float mid = FSGS_MUL(0.5f, FSGS_ADD(o.a, o.c));
    const float lam = FSGS_ADD(mid, sqrtf(fmaxf(0.1f, FSGS_FMA(mid, mid, -det))));
    const float rad = ceilf(FSGS_MUL(3.f, sqrtf(lam)));
    o.px = FSGS_MUL(FSGS_FMA(FSGS_FMA(hx, pw, 1.0f), (float)cc.W, -1.0f), 0.5f);
    o.py = FSGS_MUL(FSGS_FMA(FSGS_FMA(hy, pw, 1.0f), (float)cc.H, -1.0f), 0.5f);
    o.rminx = clampi((int)((o.px - rad) / TILE), 0, cc.gx);
    o.rminy = clampi((int)((o.py - rad) / TILE), 0, cc.gy);
    o.rmaxx = clampi((int)((o.px + rad + TILE - 1) / TILE), 0, cc.gx);
    o.rmaxy = clampi((int)((o.py + rad + TILE - 1) / TILE), 0, cc.gy);
    if ((o.rmaxx - o.rminx) * (o.rmaxy - o.rminy) == 0) return false;
    o.depth = vz;
    o.radius = (int)rad;
    return true;
}

// ---- exact tile culling ---------------------------------------------------------------------
// A Gaussian can only change a pixel if alpha = opacity*exp(-q(d)) >= 1/255, i.e.
// q(d) = 0.5*(A dx^2 + C dy^2) + B dx dy <= tau = ln(255*opacity).  A tile of the reference's
// rectangle whose pixel square lies entirely outside that ellipse contributes nothing (every
// pixel takes the reference's `alpha < 1/255 -> continue` branch), so dropping the
// (tile, Gaussian) instance leaves the image and all gradients unchanged.  The test below is the
// exact minimum of the convex quadratic q over the (continuous) pixel square, padded so that
// float rounding can only keep extra tiles, never drop a contributing one.
struct CullEllipse {
    float px, py, A, B, C, tau;   // tau already padded; tau < 0 => Gaussian can never contribute
    float hx, hy;                 // half extents of the ellipse's axis-aligned bounding box
    float nBC, nBA;               // -B/C, -B/A: argmin slope of q along a vertical / horizontal line
};

// 1/x: MUFU.RCP + one Newton step (avoids the ~10-instruction IEEE division in hot loops)
FSGS_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return __fmaf_rn(r, __fmaf_rn(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}

FSGS_HD CullEllipse make_cull_ellipse(float px, float py, float conx, float cony, float conz, float opacity) {
    CullEllipse e;
    e.px = px; e.py = py; e.A = conx; e.B = cony; e.C = conz;
    const float t = logf(255.0f * opacity);
    e.tau = (opacity * 255.0f >= 0.999f) ? (t * 1.0005f + 0.01f) : -1.0f;
    // bbox of {q <= tau}: |dx| <= sqrt(2 tau * C / det(conic)), |dy| <= sqrt(2 tau * A / det(conic))
    const float det = fmaxf(conx * conz - cony * cony, 1e-30f);
    const float tt = fmaxf(e.tau, 0.f) * 2.0f;
    e.hx = sqrtf(tt * conz / det) * 1.0005f + 0.01f;
    e.hy = sqrtf(tt * conx / det) * 1.0005f + 0.01f;
    e.nBC = -cony * fast_rcp(conz);
    e.nBA = -cony * fast_rcp(conx);
    return e;
}

// min over dy in [lo,hi] of 0.5*(A dx^2 + C dy^2) + B dx dy for fixed dx
FSGS_HD float edge_min(float A, float B, float C, float dx, float lo, float hi) {
    float dy = -B * dx / C;
    dy = fminf(hi, fmaxf(lo, dy));
    return 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy;
}
// the same with the slope nBC = -B/C of the unconstrained minimiser precomputed (evaluating q a rounding
// error away from the exact minimiser changes it only to second order -- far inside the padding of tau)
FSGS_HD float edge_min_s(float A, float B, float C, float nBC, float dx, float lo, float hi) {
    const float dy = fminf(hi, fmaxf(lo, nBC * dx));
    return 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy;
}

// Does tile (tx,ty) contain a pixel centre (integer coordinates, as the compositor uses) with
// q <= tau?  Conservative (continuous relaxation + padding).
FSGS_HD bool tile_hit(const CullEllipse &e, int tx, int ty) {
    const float x0 = (float)(tx * TILE), x1 = x0 + (float)(TILE - 1);
    const float y0 = (float)(ty * TILE), y1 = y0 + (float)(TILE - 1);
    // d = centre - pixel, so dx ranges over [px - x1, px - x0]
    const float dxl = e.px - x1, dxh = e.px - x0, dyl = e.py - y1, dyh = e.py - y0;
    if (dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f) return true;   // centre inside the square
    float q = edge_min_s(e.A, e.B, e.C, e.nBC, dxl, dyl, dyh);
    q = fminf(q, edge_min_s(e.A, e.B, e.C, e.nBC, dxh, dyl, dyh));
    q = fminf(q, edge_min_s(e.C, e.B, e.A, e.nBA, dyl, dxl, dxh));
    q = fminf(q, edge_min_s(e.C, e.B, e.A, e.nBA, dyh, dxl, dxh));
    return q <= e.tau;
}

// Tile range = reference rectangle intersected with the ellipse's bounding box.
FSGS_HD void cull_rect(const CullEllipse &e, int rminx, int rminy, int rmaxx, int rmaxy, int &x0, int &y0,
                       int &x1, int &y1) {
    if (e.tau < 0.f) { x0 = x1 = y0 = y1 = 0; return; }
    // tile tx covers pixels [16 tx, 16 tx + 15]
    const int bx0 = (int)floorf((e.px - e.hx) / TILE), bx1 = (int)floorf((e.px + e.hx) / TILE) + 1;
    const int by0 = (int)floorf((e.py - e.hy) / TILE), by1 = (int)floorf((e.py + e.hy) / TILE) + 1;
    x0 = rminx > bx0 ? rminx : bx0; x1 = rmaxx < bx1 ? rmaxx : bx1;
    y0 = rminy > by0 ? rminy : by0; y1 = rmaxy < by1 ? rmaxy : by1;
    if (x1 < x0) x1 = x0;
    if (y1 < y0) y1 = y0;
}

// ---- spherical harmonics ----------------------------------------------------------------------
// basis values B[0..(deg+1)^2) at unit direction d (utils/sh_utils.py:74-100)
FSGS_HD void sh_basis(int deg, float x, float y, float z, float *B) {
    B[0] = SH_C0;
    if (deg > 0) {
        B[1] = -SH_C1 * y; B[2] = SH_C1 * z; B[3] = -SH_C1 * x;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            B[4] = SH_C2_0 * xy; B[5] = SH_C2_1 * yz; B[6] = SH_C2_2 * (2.f * zz - xx - yy);
            B[7] = SH_C2_3 * xz; B[8] = SH_C2_4 * (xx - yy);
            if (deg > 2) {
                B[9] = SH_C3_0 * y * (3.f * xx - yy); B[10] = SH_C3_1 * xy * z;
                B[11] = SH_C3_2 * y * (4.f * zz - xx - yy); B[12] = SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
                B[13] = SH_C3_4 * x * (4.f * zz - xx - yy); B[14] = SH_C3_5 * z * (xx - yy);
                B[15] = SH_C3_6 * x * (xx - 3.f * yy);
            }
        }
    }
}

// d(B[k])/d(dir) for k >= 1 (k = 0 is constant); dB is [16][3], entries not listed stay 0.
FSGS_HD void sh_basis_grad(int deg, float x, float y, float z, float (*dB)[3]) {
    for (int k = 0; k < 16; ++k) { dB[k][0] = 0.f; dB[k][1] = 0.f; dB[k][2] = 0.f; }
    if (deg > 0) {
        dB[1][1] = -SH_C1; dB[2][2] = SH_C1; dB[3][0] = -SH_C1;
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            dB[4][0] = SH_C2_0 * y; dB[4][1] = SH_C2_0 * x;
            dB[5][1] = SH_C2_1 * z; dB[5][2] = SH_C2_1 * y;
            dB[6][0] = SH_C2_2 * -2.f * x; dB[6][1] = SH_C2_2 * -2.f * y; dB[6][2] = SH_C2_2 * 4.f * z;
            dB[7][0] = SH_C2_3 * z; dB[7][2] = SH_C2_3 * x;
            dB[8][0] = SH_C2_4 * 2.f * x; dB[8][1] = SH_C2_4 * -2.f * y;
            if (deg > 2) {
                dB[9][0] = SH_C3_0 * 6.f * xy; dB[9][1] = SH_C3_0 * (3.f * xx - 3.f * yy);
                dB[10][0] = SH_C3_1 * yz; dB[10][1] = SH_C3_1 * xz; dB[10][2] = SH_C3_1 * xy;
                dB[11][0] = SH_C3_2 * -2.f * xy; dB[11][1] = SH_C3_2 * (4.f * zz - xx - 3.f * yy);
                dB[11][2] = SH_C3_2 * 8.f * yz;
                dB[12][0] = SH_C3_3 * -6.f * xz; dB[12][1] = SH_C3_3 * -6.f * yz;
                dB[12][2] = SH_C3_3 * (6.f * zz - 3.f * xx - 3.f * yy);
                dB[13][0] = SH_C3_4 * (4.f * zz - 3.f * xx - yy); dB[13][1] = SH_C3_4 * -2.f * xy;
                dB[13][2] = SH_C3_4 * 8.f * xz;
                dB[14][0] = SH_C3_5 * 2.f * xz; dB[14][1] = SH_C3_5 * -2.f * yz; dB[14][2] = SH_C3_5 * (xx - yy);
                dB[15][0] = SH_C3_6 * (3.f * xx - 3.f * yy); dB[15][1] = SH_C3_6 * -6.f * xy;
            }
        }
    }
}

// ---- backward of the projection (K8 + the mean part of K9) ---------------------------------------
// In : dconic = dL/d(conic.x, conic.y [half], conic.z) as accumulated by the compositor,
//      dmean2D = dL/d(NDC x,y) (already scaled by 0.5W / 0.5H), ddepth = dL/d(view z).
// Out: dmean[3] (gradient w.r.t. the point fed to the rasteriser), dc6[6] (w.r.t. Sigma entries).
FSGS_HD void project_backward(const CamConst &cc, const float *V, const float *PM, const float *mean,
                              const float *c6, float gconx, float gcony, float gconz, float g2x, float g2y,
                              float gdepth, float *dmean, float *dc6) {
    float t[3], M[6], xm, ym, SM0[3], SM1[3], a, b, c;
    ewa_M(cc, V, mean, t, M, xm, ym);
    ewa_abc(M, c6, SM0, SM1, a, b, c);
    const float denom = a * c - b * b;
    const float d2inv = 1.f / (denom * denom + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    if (d2inv != 0.f) {
        dL_da = d2inv * (-c * c * gconx + 2.f * b * c * gcony + (denom - a * c) * gconz);
        dL_dc = d2inv * (-a * a * gconz + 2.f * a * b * gcony + (denom - a * c) * gconx);
        dL_db = d2inv * 2.f * (b * c * gconx - (denom + 2.f * b * b) * gcony + a * b * gconz);
    }
    dc6[0] = M[0] * M[0] * dL_da + M[0] * M[3] * dL_db + M[3] * M[3] * dL_dc;
    dc6[3] = M[1] * M[1] * dL_da + M[1] * M[4] * dL_db + M[4] * M[4] * dL_dc;
    dc6[5] = M[2] * M[2] * dL_da + M[2] * M[5] * dL_db + M[5] * M[5] * dL_dc;
    dc6[1] = 2.f * M[0] * M[1] * dL_da + (M[0] * M[4] + M[1] * M[3]) * dL_db + 2.f * M[3] * M[4] * dL_dc;
    dc6[2] = 2.f * M[0] * M[2] * dL_da + (M[0] * M[5] + M[2] * M[3]) * dL_db + 2.f * M[3] * M[5] * dL_dc;
    dc6[4] = 2.f * M[2] * M[1] * dL_da + (M[1] * M[5] + M[2] * M[4]) * dL_db + 2.f * M[4] * M[5] * dL_dc;
    float dM[6];
    for (int j = 0; j < 3; ++j) {
        dM[j] = 2.f * SM0[j] * dL_da + SM1[j] * dL_db;
        dM[3 + j] = 2.f * SM1[j] * dL_dc + SM0[j] * dL_db;
    }
    const float dJ00 = V[0] * dM[0] + V[4] * dM[1] + V[8] * dM[2];
    const float dJ02 = V[2] * dM[0] + V[6] * dM[1] + V[10] * dM[2];
    const float dJ11 = V[1] * dM[3] + V[5] * dM[4] + V[9] * dM[5];
    const float dJ12 = V[2] * dM[3] + V[6] * dM[4] + V[10] * dM[5];
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xm * -cc.fx * tz2 * dJ02;
    const float dty = ym * -cc.fy * tz2 * dJ12;
    const float dtz = -cc.fx * tz2 * dJ00 - cc.fy * tz2 * dJ11 + (2.f * cc.fx * t[0]) * tz3 * dJ02 +
                      (2.f * cc.fy * t[1]) * tz3 * dJ12;
    dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
    dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
    dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    // projected centre (K9)
    const float hw = PM[3] * mean[0] + PM[7] * mean[1] + PM[11] * mean[2] + PM[15];
    const float mw = 1.0f / (hw + 0.0000001f);
    const float mul1 = (PM[0] * mean[0] + PM[4] * mean[1] + PM[8] * mean[2] + PM[12]) * mw * mw;
    const float mul2 = (PM[1] * mean[0] + PM[5] * mean[1] + PM[9] * mean[2] + PM[13]) * mw * mw;
    dmean[0] += (PM[0] * mw - PM[3] * mul1) * g2x + (PM[1] * mw - PM[3] * mul2) * g2y;
    dmean[1] += (PM[4] * mw - PM[7] * mul1) * g2x + (PM[5] * mw - PM[7] * mul2) * g2y;
    dmean[2] += (PM[8] * mw - PM[11] * mul1) * g2x + (PM[9] * mw - PM[11] * mul2) * g2y;
    // view-space depth
    dmean[0] += V[2] * gdepth; dmean[1] += V[6] * gdepth; dmean[2] += V[10] * gdepth;
}

// dL/dR (row-major 3x3) -> dL/dq for R = quat_to_R(q), q = (w,x,y,z) used as given.
FSGS_HD void quat_R_backward(const float *q, const float *dR, float *dq) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    dq[0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
    dq[1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] -
                   2.f * x * dR[8]);
    dq[2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] -
                   2.f * y * dR[8]);
    dq[3] = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] +
                   x * dR[6] + y * dR[7]);
}

// dSigma(6) -> d(scale) (w.r.t. the s fed to cov3d_from_scale_rot) and d(quaternion as given).
FSGS_HD void cov3d_backward(const float *s, const float *q, const float *dc6, float *ds, float *dq) {
    float R[9];
    quat_to_R(q, R);
    const float G[9] = {dc6[0], 0.5f * dc6[1], 0.5f * dc6[2], 0.5f * dc6[1], dc6[3],
                        0.5f * dc6[4], 0.5f * dc6[2], 0.5f * dc6[4], dc6[5]};
    float GR[9], dR[9];
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k)
            GR[r * 3 + k] = G[r * 3] * R[k] + G[r * 3 + 1] * R[3 + k] + G[r * 3 + 2] * R[6 + k];
    for (int k = 0; k < 3; ++k) {
        const float rgr = R[k] * GR[k] + R[3 + k] * GR[3 + k] + R[6 + k] * GR[6 + k];
        ds[k] = 2.f * s[k] * rgr;
        const float s2 = 2.f * s[k] * s[k];
        dR[k] = GR[k] * s2; dR[3 + k] = GR[3 + k] * s2; dR[6 + k] = GR[6 + k] * s2;
    }
    quat_R_backward(q, dR, dq);
}

// backward of v -> v/|v| : given dL/d(unit), unit vector n and 1/|v|, returns dL/dv
FSGS_HD void normalize_backward(const float *n, float inv_len, const float *dn, float *dv) {
    const float dot = n[0] * dn[0] + n[1] * dn[1] + n[2] * dn[2];
    dv[0] = (dn[0] - n[0] * dot) * inv_len;
    dv[1] = (dn[1] - n[1] * dot) * inv_len;
    dv[2] = (dn[2] - n[2] * dot) * inv_len;
}

// ---- per-(pixel, Gaussian) arithmetic of the compositor -----------------------------------------
// Exponent with a pinned operation order: the backward must reproduce the forward's per-pair
// decisions bit for bit, so no compiler re-association / contraction is allowed here.
FSGS_HD float gauss_power(float A, float B, float C, float dx, float dy) {
#if defined(__CUDA_ARCH__)
    const float s = __fmaf_rn(__fmul_rn(C, dy), dy, __fmul_rn(__fmul_rn(A, dx), dx));
    return __fmaf_rn(-0.5f, s, -__fmul_rn(__fmul_rn(B, dx), dy));
#else
    const float s = fmaf(C * dy, dy, (A * dx) * dx);
    return fmaf(-0.5f, s, -((B * dx) * dy));
#endif
}

// Per-pixel state of the back-to-front replay (K7).
struct BwdPixel {
    float T;
    // the colour seen BEHIND the current entry, already contracted with this pixel's upstream gradient:
    // ag_rgb = sum over the RGB planes of acc_plane * g_plane, ag_dep the same over depth | silhouette | depth^2.
    // (Everything the replay needs from the per-plane accumulators is linear in them, so two scalars replace
    // four to six accumulators and their subtract / multiply / update chains.)
    float ag_rgb, ag_dep;
};

// ---- whole-Gaussian forward / backward bodies (shared by the CUDA kernels and the CPU emulation) --
// SH -> RGB (+0.5, clamp at 0, remember which channels were clamped).  coef(k, ch) returns the
// k-th coefficient of channel ch.
template <typename Coef>
FSGS_HD void sh_to_rgb(int deg, const float *dir, Coef coef, float *rgb, uint8_t &clamp) {
    // B stays in registers: fixed trip count, basis values above the active degree are exact zeros
    // (their products add +0 and leave the sums bit-identical to a loop over the active coefficients only)
    float B[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; ++k) B[k] = 0.f;
    sh_basis(deg, dir[0], dir[1], dir[2], B);
    rgb[0] = rgb[1] = rgb[2] = 0.f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; ++k) {
        rgb[0] += B[k] * coef(k, 0); rgb[1] += B[k] * coef(k, 1); rgb[2] += B[k] * coef(k, 2);
    }
    clamp = 0;
    for (int ch = 0; ch < 3; ++ch) {
        rgb[ch] += 0.5f;
        if (rgb[ch] < 0.f) { clamp |= (uint8_t)(1u << ch); rgb[ch] = 0.f; }
    }
}

// Backward of sh_to_rgb + the direction normalisation: writes dcoef(k, ch, value) for every stored
// coefficient (zeros beyond the active degree) and returns dL/d(un-normalised direction) in dv.
template <typename Coef, typename DCoef>
FSGS_HD void sh_to_rgb_backward(int deg, int n_coeffs, const float *dir, float inv_len, Coef coef, uint8_t clamp,
                                const float *grgb, DCoef dcoef, float *dv) {
    float B[16], dB[16][3];
    sh_basis(deg, dir[0], dir[1], dir[2], B);
    sh_basis_grad(deg, dir[0], dir[1], dir[2], dB);
    const int nb = (deg + 1) * (deg + 1);
    const float gc[3] = {(clamp & 1) ? 0.f : grgb[0], (clamp & 2) ? 0.f : grgb[1], (clamp & 4) ? 0.f : grgb[2]};
    float ddir[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < n_coeffs; ++k) {
        if (k < nb) {
            // read the coefficients BEFORE writing their gradients: the fused backward kernel computes
            // in place (gradient overwrites the coefficient in its shared-memory staging buffer)
            const float t = coef(k, 0) * gc[0] + coef(k, 1) * gc[1] + coef(k, 2) * gc[2];
            dcoef(k, 0, B[k] * gc[0]); dcoef(k, 1, B[k] * gc[1]); dcoef(k, 2, B[k] * gc[2]);
            ddir[0] += dB[k][0] * t; ddir[1] += dB[k][1] * t; ddir[2] += dB[k][2] * t;
        } else {
            dcoef(k, 0, 0.f); dcoef(k, 1, 0.f); dcoef(k, 2, 0.f);
        }
    }
    normalize_backward(dir, inv_len, ddir, dv);
}

FSGS_HD float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// Fused flavour, forward: transform_to_frame + activations + projection + SH for Gaussian i.
// pose: row-major 3x4 (first 12 floats of LearnPose.forward's output).  Returns false if skipped.
// Pose-independent geometry of Gaussian i: Sigma_3D from the raw scale / rotation parameters.  Its inputs and outputs
// pass through opaque register barriers on the device, so that the compiler forms the same multiply-adds here
// whatever kernel the function is inlined into (it cannot contract across the barriers): k_preprocess_fused and
// k_freeze_model must produce the SAME bits (the frozen-model forward is tested bit-identical).
FSGS_HD void frozen_sigma(const CamConst &cc, const float *sc_raw, const float *rot_raw, float *c6) {
    float s[3] = {cc.mod * expf(sc_raw[0]), cc.mod * expf(sc_raw[1]), cc.mod * expf(sc_raw[2])};
    const float qq = FSGS_FMA(rot_raw[3], rot_raw[3], dot3(rot_raw[0], rot_raw[0], rot_raw[1], rot_raw[1], rot_raw[2], rot_raw[2]));
    const float qn = fmaxf(sqrtf(qq), 1e-12f);
    const float qi = 1.0f / qn;
    float q[4] = {rot_raw[0] * qi, rot_raw[1] * qi, rot_raw[2] * qi, rot_raw[3] * qi};
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int k = 0; k < 3; ++k) asm volatile("" : "+f"(s[k]));
#pragma unroll
    for (int k = 0; k < 4; ++k) asm volatile("" : "+f"(q[k]));
#endif
    cov3d_from_scale_rot(s, q, c6);
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int k = 0; k < 6; ++k) asm volatile("" : "+f"(c6[k]));
#endif
}
// ... and its colour: SH view direction = WORLD position minus the frozen camera centre (reference quirk iii)
FSGS_HD void frozen_colour(const CamConst &cc, const float *cam_center, const float *xyz, const float *dc,
                           const float *rest, float *rgb, uint8_t &clamp) {
    float d[3] = {xyz[0] - cam_center[0], xyz[1] - cam_center[1], xyz[2] - cam_center[2]};
    const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] *= inv; d[1] *= inv; d[2] *= inv;
    sh_to_rgb(cc.sh_deg, d, [&](int k, int ch) { return k == 0 ? dc[ch] : rest[3 * (k - 1) + ch]; }, rgb, clamp);
}

FSGS_HD bool fused_forward_one(const CamConst &cc, const float *V, const float *PM, const float *pose,
                               const float *cam_center, const float *xyz, const float *dc, const float *rest,
                               float op_raw, const float *sc_raw, const float *rot_raw, Splat &sp, float &opacity,
                               float *rgb, uint8_t &clamp) {
    float mean[3];
    pose_mean(pose, xyz, mean);
    float c6[6];
    frozen_sigma(cc, sc_raw, rot_raw, c6);
    if (!project_gaussian(cc, V, PM, mean, c6, sp)) return false;
    opacity = sigmoidf(op_raw);
    frozen_colour(cc, cam_center, xyz, dc, rest, rgb, clamp);
    return true;
}

// The same forward split at the pose: everything of Gaussian i that does not depend on the pose (activated
// opacity, Sigma_3D, SH colour + clamp mask -- reference quirks ii / iii) ...
FSGS_HD void fused_frozen_one(const CamConst &cc, const float *cam_center, const float *xyz, const float *dc,
                              const float *rest, float op_raw, const float *sc_raw, const float *rot_raw, float *c6,
                              float &opacity, float *rgb, uint8_t &clamp) {
    frozen_sigma(cc, sc_raw, rot_raw, c6);
    opacity = sigmoidf(op_raw);
    frozen_colour(cc, cam_center, xyz, dc, rest, rgb, clamp);
}
// ... and the pose-dependent rest: camera-frame mean + projection.  Returns false if skipped.
FSGS_HD bool fused_forward_frozen_one(const CamConst &cc, const float *V, const float *PM, const float *pose,
                                      const float *xyz, const float *c6, Splat &sp) {
    float mean[3];
    pose_mean(pose, xyz, mean);
    return project_gaussian(cc, V, PM, mean, c6, sp);
}

// Geometry half of the fused backward, shared by the full and the pose-only bodies: camera-frame mean,
// activated scale / rotation, then dL/d(camera-frame mean) and dL/d(Sigma) from the accumulator row.
FSGS_HD void fused_backward_geometry(const CamConst &cc, const float *V, const float *PM, const float *pose,
                                     const float *xyz, const float *sc_raw, const float *rot_raw, const float *acc,
                                     float *s, float *q, float &qi, float *dmean, float *dc6) {
    float mean[3];
    pose_mean(pose, xyz, mean);
    s[0] = cc.mod * expf(sc_raw[0]); s[1] = cc.mod * expf(sc_raw[1]); s[2] = cc.mod * expf(sc_raw[2]);
    const float qn = fmaxf(sqrtf(rot_raw[0] * rot_raw[0] + rot_raw[1] * rot_raw[1] + rot_raw[2] * rot_raw[2] +
                                 rot_raw[3] * rot_raw[3]), 1e-12f);
    qi = 1.0f / qn;
    q[0] = rot_raw[0] * qi; q[1] = rot_raw[1] * qi; q[2] = rot_raw[2] * qi; q[3] = rot_raw[3] * qi;
    float c6[6];
    cov3d_from_scale_rot(s, q, c6);
    // acc[9] = dL/d(view z) through the depth / depth^2 colour planes
    project_backward(cc, V, PM, mean, c6, acc[2], acc[3], acc[4], acc[0], acc[1], acc[9], dmean, dc6);
}

// Pose-only backward for Gaussian i (radius > 0): pg[12] = this Gaussian's contribution to dL/d(pose[:3,:]),
// i.e. the transform_to_frame matmul backward (reference scene/pose_optimizer.py:987) for one row.  The pose reaches
// a splat only through its camera-frame mean (colours use the world position and a frozen camera centre, quirk iii),
// so the accumulator's colour / opacity columns and the SH coefficients are not read.
FSGS_HD void fused_backward_pose_one(const CamConst &cc, const float *V, const float *PM, const float *pose,
                                     const float *xyz, const float *sc_raw, const float *rot_raw, const float *acc,
                                     float *pg) {
    float s[3], q[4], qi, dmean[3], dc6[6];
    fused_backward_geometry(cc, V, PM, pose, xyz, sc_raw, rot_raw, acc, s, q, qi, dmean, dc6);
    for (int r = 0; r < 3; ++r) {
        pg[4 * r] = dmean[r] * xyz[0]; pg[4 * r + 1] = dmean[r] * xyz[1]; pg[4 * r + 2] = dmean[r] * xyz[2];
        pg[4 * r + 3] = dmean[r];
    }
}

// Fused flavour, backward for Gaussian i (radius > 0).  acc = the compositor's 12-float row.
// Outputs: dxyz[3], dfdc[3], drest[45] (may be null), dop_raw, ds_raw[3], dq_raw[4],
// pg[12] = this Gaussian's contribution to dL/d(pose[:3,:]) (zeros unless cam_grad), m2d[2].
FSGS_HD void fused_backward_one(const CamConst &cc, const float *V, const float *PM, const float *pose,
                                const float *cam_center, const float *xyz, const float *dc, const float *rest,
                                float op_raw, const float *sc_raw, const float *rot_raw, uint8_t clamp,
                                const float *acc, int gs_grad, int cam_grad, float *dxyz, float *dfdc, float *drest,
                                float &dop_raw, float *ds_raw, float *dq_raw, float *pg, float *m2d, float *gc = nullptr) {
    float s[3], q[4], qi, dmean[3], dc6[6], ds[3], dq[4];
    fused_backward_geometry(cc, V, PM, pose, xyz, sc_raw, rot_raw, acc, s, q, qi, dmean, dc6);
    cov3d_backward(s, q, dc6, ds, dq);
    // exp / normalize / sigmoid
    ds_raw[0] = ds[0] * s[0]; ds_raw[1] = ds[1] * s[1]; ds_raw[2] = ds[2] * s[2];
    const float qdot = q[0] * dq[0] + q[1] * dq[1] + q[2] * dq[2] + q[3] * dq[3];
    for (int k = 0; k < 4; ++k) dq_raw[k] = (dq[k] - q[k] * qdot) * qi;
    const float o = sigmoidf(op_raw);
    dop_raw = acc[5] * o * (1.f - o);
    // SH (world position, frozen camera centre): gradient goes straight to xyz, not through the pose
    float d[3] = {xyz[0] - cam_center[0], xyz[1] - cam_center[1], xyz[2] - cam_center[2]};
    const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] *= inv; d[1] *= inv; d[2] *= inv;
    const float grgb[3] = {acc[6], acc[7], acc[8]};
    // gc = the colour gradient after the clamp mask: every SH-coefficient gradient is basis(dir) x gc, and
    // dir / the mask do not depend on the frame (quirk iii) -- what the frame-parallel exchange reduces
    if (gc) { gc[0] = (clamp & 1) ? 0.f : grgb[0]; gc[1] = (clamp & 2) ? 0.f : grgb[1]; gc[2] = (clamp & 4) ? 0.f : grgb[2]; }
    sh_to_rgb_backward(
        cc.sh_deg, 16, d, inv, [&](int k, int ch) { return k == 0 ? dc[ch] : rest[3 * (k - 1) + ch]; }, clamp, grgb,
        [&](int k, int ch, float v) {
            if (k == 0) dfdc[ch] = v;
            else if (drest) drest[3 * (k - 1) + ch] = v;
        },
        dxyz);
    // transform_to_frame backward
    if (gs_grad) {
        dxyz[0] += pose[0] * dmean[0] + pose[4] * dmean[1] + pose[8] * dmean[2];
        dxyz[1] += pose[1] * dmean[0] + pose[5] * dmean[1] + pose[9] * dmean[2];
        dxyz[2] += pose[2] * dmean[0] + pose[6] * dmean[1] + pose[10] * dmean[2];
    }
    for (int k = 0; k < 12; ++k) pg[k] = 0.f;
    if (cam_grad) {
        for (int r = 0; r < 3; ++r) {
            pg[4 * r] = dmean[r] * xyz[0]; pg[4 * r + 1] = dmean[r] * xyz[1]; pg[4 * r + 2] = dmean[r] * xyz[2];
            pg[4 * r + 3] = dmean[r];
        }
    }
    m2d[0] = acc[10]; m2d[1] = acc[11];
}

// API flavour, forward.  Exactly one of (colors_precomp | shs) and one of (scales+rots | cov3D).
FSGS_HD bool api_forward_one(const CamConst &cc, const float *V, const float *PM, const float *campos,
                             const float *mean, const float *color_i, const float *sh_i, const float *scale_i,
                             const float *rot_i, const float *cov_i, Splat &sp, float *rgb, uint8_t &clamp) {
    float c6[6];
    if (cov_i) {
        for (int k = 0; k < 6; ++k) c6[k] = cov_i[k];
    } else {
        const float s[3] = {cc.mod * scale_i[0], cc.mod * scale_i[1], cc.mod * scale_i[2]};
        cov3d_from_scale_rot(s, rot_i, c6);
    }
    if (!project_gaussian(cc, V, PM, mean, c6, sp)) return false;
    clamp = 0;
    if (sh_i) {
        float d[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
        const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        d[0] *= inv; d[1] *= inv; d[2] *= inv;
        // (sh_to_rgb visits all 16 basis slots; a tensor with fewer stored coefficients reads as zero there)
        sh_to_rgb(cc.sh_deg, d, [&](int k, int ch) { return k < cc.n_coeffs ? sh_i[3 * k + ch] : 0.f; }, rgb, clamp);
    } else {
        rgb[0] = color_i[0]; rgb[1] = color_i[1]; rgb[2] = color_i[2];
    }
    return true;
}

// API flavour, backward (radius > 0).  dsh_i may be null.
FSGS_HD void api_backward_one(const CamConst &cc, const float *V, const float *PM, const float *campos,
                              const float *mean, const float *sh_i, const float *scale_i, const float *rot_i,
                              const float *cov_i, uint8_t clamp, const float *acc, float *dmean, float *dc6,
                              float *dsh_i, float *ds, float *dq) {
    float c6[6], s[3] = {0.f, 0.f, 0.f};
    if (cov_i) {
        for (int k = 0; k < 6; ++k) c6[k] = cov_i[k];
    } else {
        s[0] = cc.mod * scale_i[0]; s[1] = cc.mod * scale_i[1]; s[2] = cc.mod * scale_i[2];
        cov3d_from_scale_rot(s, rot_i, c6);
    }
    project_backward(cc, V, PM, mean, c6, acc[2], acc[3], acc[4], acc[0], acc[1], acc[9], dmean, dc6);
    if (sh_i && dsh_i) {
        float d[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
        const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        d[0] *= inv; d[1] *= inv; d[2] *= inv;
        const float grgb[3] = {acc[6], acc[7], acc[8]};
        float dv[3];
        sh_to_rgb_backward(
            cc.sh_deg, cc.n_coeffs, d, inv, [&](int k, int ch) { return sh_i[3 * k + ch]; }, clamp, grgb,
            [&](int k, int ch, float v) { dsh_i[3 * k + ch] = v; }, dv);
        dmean[0] += dv[0]; dmean[1] += dv[1]; dmean[2] += dv[2];
    }
    ds[0] = ds[1] = ds[2] = 0.f;
    dq[0] = dq[1] = dq[2] = dq[3] = 0.f;
    if (!cov_i) {
        cov3d_backward(s, rot_i, dc6, ds, dq);
        ds[0] *= cc.mod; ds[1] *= cc.mod; ds[2] *= cc.mod;
    }
}

// ---- per-frame pose: LearnPose.forward (scene/pose_optimizer.py:822-877) -------------------------
// r_raw (w,x,y,z) -> F.normalize (eps 1e-12) -> q2rot (normalises once more) -> Rt = [[R, t],[0 0 0 1]]
FSGS_HD void pose_forward(const float *r_raw, const float *t, float *Rt) {
    const float n1 = fmaxf(sqrtf(r_raw[0] * r_raw[0] + r_raw[1] * r_raw[1] + r_raw[2] * r_raw[2] + r_raw[3] * r_raw[3]), 1e-12f);
    const float q1[4] = {r_raw[0] / n1, r_raw[1] / n1, r_raw[2] / n1, r_raw[3] / n1};
    const float n2 = sqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
    const float q2[4] = {q1[0] / n2, q1[1] / n2, q1[2] / n2, q1[3] / n2};
    float R[9];
    quat_to_R(q2, R);
    for (int r = 0; r < 3; ++r) {
        Rt[4 * r] = R[3 * r]; Rt[4 * r + 1] = R[3 * r + 1]; Rt[4 * r + 2] = R[3 * r + 2]; Rt[4 * r + 3] = t[r];
    }
    Rt[12] = 0.f; Rt[13] = 0.f; Rt[14] = 0.f; Rt[15] = 1.f;
}

// dL/dRt [4,4] -> dL/dr_raw [4], dL/dt [3]
FSGS_HD void pose_backward(const float *r_raw, const float *dRt, float *dr, float *dt) {
    const float n1 = fmaxf(sqrtf(r_raw[0] * r_raw[0] + r_raw[1] * r_raw[1] + r_raw[2] * r_raw[2] + r_raw[3] * r_raw[3]), 1e-12f);
    const float q1[4] = {r_raw[0] / n1, r_raw[1] / n1, r_raw[2] / n1, r_raw[3] / n1};
    const float n2 = sqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
    const float q2[4] = {q1[0] / n2, q1[1] / n2, q1[2] / n2, q1[3] / n2};
    const float dR[9] = {dRt[0], dRt[1], dRt[2], dRt[4], dRt[5], dRt[6], dRt[8], dRt[9], dRt[10]};
    float dq2[4], dq1[4];
    quat_R_backward(q2, dR, dq2);
    float dot = q2[0] * dq2[0] + q2[1] * dq2[1] + q2[2] * dq2[2] + q2[3] * dq2[3];
    for (int k = 0; k < 4; ++k) dq1[k] = (dq2[k] - q2[k] * dot) / n2;
    dot = q1[0] * dq1[0] + q1[1] * dq1[1] + q1[2] * dq1[2] + q1[3] * dq1[3];
    for (int k = 0; k < 4; ++k) dr[k] = (dq1[k] - q1[k] * dot) / n1;
    dt[0] = dRt[3]; dt[1] = dRt[7]; dt[2] = dRt[11];
}

// ---- compositor arithmetic, second generation ---------------------------------------------------
// The depth-sorted per-instance records carry the conic pre-scaled for a base-2 exponent:
//   a2 = -0.5*log2(e)*A,  b2 = -log2(e)*B,  c2 = -0.5*log2(e)*C
// so that  log2(G) = dx*(a2*dx + b2*dy) + c2*dy*dy  is 2 FMUL + ... 5 instructions, and
// G = ex2(log2 G) is one MUFU.  Forward and backward evaluate the SAME pinned expression on the
// SAME records, so their per-pair decisions agree bit for bit.
constexpr float LOG2E = 1.4426950408889634f;

FSGS_HD void scale_conic(float A, float B, float C, float &a2, float &b2, float &c2) {
    a2 = A * (-0.5f * LOG2E); b2 = B * (-LOG2E); c2 = C * (-0.5f * LOG2E);
}
FSGS_HD void unscale_conic(float a2, float b2, float c2, float &A, float &B, float &C) {
    A = a2 * (-2.0f / LOG2E); B = b2 * (-1.0f / LOG2E); C = c2 * (-2.0f / LOG2E);
}
FSGS_HD float gauss_power2(float a2, float b2, float c2, float dx, float dy) {
#if defined(__CUDA_ARCH__)
    const float t = __fmaf_rn(b2, dy, __fmul_rn(a2, dx));
    return __fmaf_rn(dx, t, __fmul_rn(__fmul_rn(c2, dy), dy));
#else
    const float t = fmaf(b2, dy, a2 * dx);
    return fmaf(dx, t, (c2 * dy) * dy);
#endif
}
FSGS_HD float fast_exp2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}
// Conservative test: can the rectangle of pixel centres [x0,x1]x[y0,y1] contain a pixel with
// q <= tau?  (tile_hit generalised; used for the per-instance 8x4-block masks.)
FSGS_HD bool rect_hit(float px, float py, float A, float B, float C, float tau, float x0, float x1, float y0,
                      float y1) {
    const float dxl = px - x1, dxh = px - x0, dyl = py - y1, dyh = py - y0;
    if (dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f) return true;
    float q = edge_min(A, B, C, dxl, dyl, dyh);
    q = fminf(q, edge_min(A, B, C, dxh, dyl, dyh));
    q = fminf(q, edge_min(C, B, A, dyl, dxl, dxh));
    q = fminf(q, edge_min(C, B, A, dyh, dxl, dxh));
    return q <= tau;
}
// Bit w of the mask = the 8x4 pixel block of warp w ((w&1)*8, (w>>1)*4 inside the 16x16 tile) may
// hold a pixel this splat changes.  tile_x0/tile_y0 = pixel coordinates of the tile's corner.
// Same test as rect_hit for each of the eight blocks (minimum of the quadratic over the block's four
// edges, or the centre inside the block), with the per-line terms shared: the blocks have only 4 distinct
// vertical and 8 distinct horizontal edge lines, and along a line x = const the quadratic is
//   q(dy) = 0.5 A dx^2 + dy (0.5 C dy + B dx),   minimised at dy = -B dx / C  (clamped to the edge),
// so each edge costs one clamp and two FMAs once -B/C, -B/A and the line terms are known.
FSGS_HD unsigned block_mask(float px, float py, float A, float B, float C, float opacity, int tile_x0, int tile_y0) {
    const float t = logf(255.0f * opacity);
    const float tau = (opacity * 255.0f >= 0.999f) ? (t * 1.0005f + 0.01f) : -1.0f;   // as make_cull_ellipse
    const float hA = 0.5f * A, hC = 0.5f * C;
    const float nBC = -B * fast_rcp(C), nBA = -B * fast_rcp(A);
    float dxv[4], tx[4], bx[4], sx[4];               // vertical lines x = tile_x0 + {0, 7, 8, 15}
    for (int k = 0; k < 4; ++k) {
        const float dx = px - (float)(tile_x0 + (k >> 1) * 8 + (k & 1) * 7);
        dxv[k] = dx; tx[k] = nBC * dx; bx[k] = hA * dx * dx; sx[k] = B * dx;
    }
    unsigned m = 0;
    for (int r = 0; r < 4; ++r) {
        // horizontal lines y = tile_y0 + 4 r + {0, 3};  d = centre - pixel, so dy runs over [dyl, dyh]
        const float dyh = py - (float)(tile_y0 + 4 * r), dyl = dyh - 3.0f;
        const float tyh = nBA * dyh, tyl = nBA * dyl, byh = hC * dyh * dyh, byl = hC * dyl * dyl;
        const float syh = B * dyh, syl = B * dyl;
        for (int c = 0; c < 2; ++c) {
            const float dxh = dxv[2 * c], dxl = dxv[2 * c + 1];
            bool hit = dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f;        // centre inside the block
            float d, q;
            d = fminf(dyh, fmaxf(dyl, tx[2 * c]));     q = bx[2 * c] + d * (hC * d + sx[2 * c]);
            d = fminf(dyh, fmaxf(dyl, tx[2 * c + 1])); q = fminf(q, bx[2 * c + 1] + d * (hC * d + sx[2 * c + 1]));
            d = fminf(dxh, fmaxf(dxl, tyh));           q = fminf(q, byh + d * (hA * d + syh));
            d = fminf(dxh, fmaxf(dxl, tyl));           q = fminf(q, byl + d * (hA * d + syl));
            if (hit || q <= tau) m |= 1u << (2 * r + c);
        }
    }
    return m;
}

// One contributing pair of the backward compositor, "moment" form: instead of the final
// per-Gaussian gradients each pixel adds the moments of q = G*o*dL/dalpha over d = centre - pixel;
// bwd_finalize turns the summed moments into the accumulator row once per (tile, Gaussian).
//   v = [Sx, Sy, Sxx, Sxy, Syy, S0 | d r, d g, d b, d z | Sx_rgb, Sy_rgb]
// LEVEL says which upstream gradients are non-zero for the whole warp (decided once per warp, so the
// dispatch is uniform): 0 = colour planes only, 1 = + the depth plane, 2 = + silhouette / depth^2
// (fused flavour only).  Skipped planes contribute exact zeros, so the result is unchanged.
//
// The pair is split in two halves so that the CUDA kernel can run them on different thread
// layouts (fsgs_kernels_composite.cuh):
//   bwd_pair_weights : the per-PIXEL half -- advances the pixel's replay state and returns the three
//                      scalars everything else is linear in:  q = G*o*dL/dalpha,  w = alpha*T,
//                      q_rgb = G*o*(dL/dalpha of the RGB planes only).
//                      `Go` = G*opacity and `alpha` may both be passed as 0 for a pair that does not
//                      contribute: the state then advances by an exact no-op (T*1, acc' folded with
//                      weight 0 at the next step) and q = w = q_rgb = 0.
//   bwd_pair_moments : the per-GAUSSIAN half -- the 12 moments of one pair from (q, w, q_rgb).
template <bool FUSED, int LEVEL>
FSGS_HD void bwd_pair_weights(BwdPixel &s, float Go, float alpha, float cr, float cg, float cb, float z,
                              const float *g, float T_final, float bgdot_rgb, float bgdot_dep, float &q, float &w,
                              float &q_rgb) {
    // upstream: accum_rec = last_alpha * last_color + (1 - last_alpha) * accum_rec per plane, then
    // dL/dalpha += (c - accum_rec) * dL/dC.  With e = c - acc the update for the next (nearer) entry is
    // acc + alpha * e; contracting with g first gives  da = c.g - ag,  ag <- ag + alpha * da.
    const float inv = fast_rcp(1.f - alpha);
    s.T = s.T * inv;
    w = alpha * s.T;
    float da_rgb = fmaf(cb, g[2], fmaf(cg, g[1], cr * g[0])) - s.ag_rgb;
    s.ag_rgb = fmaf(alpha, da_rgb, s.ag_rgb);
    float da_dep = 0.f;
    if (LEVEL >= 1) {
        float cgd = z * g[3];
        if (FUSED && LEVEL >= 2) cgd = fmaf(z * z, g[5], cgd + g[4]);
        da_dep = cgd - s.ag_dep;
        s.ag_dep = fmaf(alpha, da_dep, s.ag_dep);
    }
    const float tf = -T_final * inv;
    da_rgb = da_rgb * s.T + tf * bgdot_rgb;
    if (LEVEL >= 1) {
        da_dep = da_dep * s.T;
        if (FUSED) da_dep += tf * bgdot_dep;
    }
    q = Go * (da_rgb + da_dep);
    q_rgb = Go * da_rgb;
}

template <bool FUSED, int LEVEL>
FSGS_HD void bwd_pair_moments(float q, float w, float q_rgb, float z, float dx, float dy, const float *g, float *v) {
    const float qx = q * dx, qy = q * dy;
    v[0] = qx; v[1] = qy; v[2] = qx * dx; v[3] = qx * dy; v[4] = qy * dy; v[5] = q;
    v[6] = w * g[0]; v[7] = w * g[1]; v[8] = w * g[2];
    float dz = 0.f;
    if (LEVEL >= 1) {
        dz = w * g[3];
        if (FUSED && LEVEL >= 2) dz += w * 2.f * z * g[5];
    }
    v[9] = dz;
    if (FUSED) {
        v[10] = q_rgb * dx; v[11] = q_rgb * dy;
    } else {
        v[10] = 0.f; v[11] = 0.f;
    }
}

template <bool FUSED, int LEVEL>
FSGS_HD void bwd_pair2(BwdPixel &s, float opacity, float cr, float cg, float cb, float z, float dx, float dy,
                       float G, float alpha, const float *g, float T_final, float bgdot_rgb, float bgdot_dep,
                       float *v) {
    float q, w, q_rgb;
    bwd_pair_weights<FUSED, LEVEL>(s, G * opacity, alpha, cr, cg, cb, z, g, T_final, bgdot_rgb, bgdot_dep, q, w, q_rgb);
    bwd_pair_moments<FUSED, LEVEL>(q, w, q_rgb, z, dx, dy, g, v);
}

// Summed moments of one (tile, Gaussian) -> accumulator row (layout in fsgs_device.cuh).
// kx = 0.5 W, ky = 0.5 H (the reference returns the mean2D gradient in NDC units).
FSGS_HD void bwd_finalize(const float *m, float a2, float b2, float c2, float opacity, float kx, float ky, bool fused,
                          float *out) {
    float A, B, C;
    unscale_conic(a2, b2, c2, A, B, C);
    out[0] = -kx * (A * m[0] + B * m[1]);
    out[1] = -ky * (C * m[1] + B * m[0]);
    out[2] = -0.5f * m[2]; out[3] = -0.5f * m[3]; out[4] = -0.5f * m[4];
    out[5] = m[5] / opacity;
    out[6] = m[6]; out[7] = m[7]; out[8] = m[8]; out[9] = m[9];
    if (fused) {
        out[10] = -kx * (A * m[10] + B * m[11]);
        out[11] = -ky * (C * m[11] + B * m[10]);
    } else {
        out[10] = out[0]; out[11] = out[1];
    }
}

}  // namespace fsgs
