// tests/cpu_emul/emul.cpp -- TEST INFRASTRUCTURE ONLY (never shipped, never imported by the package).
//
// Sequential CPU emulation of the CUDA pipeline, compiled with g++ from the very same
// host/device header the kernels use (free-surgs_b200/csrc/fsgs_math.cuh).  It lets the build
// container -- which has no GPU -- check the kernels' arithmetic (projection, exact tile culling,
// SH, per-pair compositing backward, activation / pose backward) against the oracle before any
// GPU time is spent.  The parallel mechanics (TMA staging, warp reductions, atomics, the sorting
// network) are NOT emulated; those are covered by the `-m gpu` tests.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../free-surgs_b200/csrc/fsgs_math.cuh"

using namespace fsgs;

namespace {

struct Rec {
    float x, y, A, B, C, o, r, g, b, z;
    int radius, tiles;
};

CamConst make_cc(int W, int H, float tanfovx, float tanfovy, float mod, int sh_deg, int n_coeffs) {
    CamConst cc;
    cc.W = W; cc.H = H; cc.gx = (W + TILE - 1) / TILE; cc.gy = (H + TILE - 1) / TILE;
    cc.fx = W / (2.0f * tanfovx); cc.fy = H / (2.0f * tanfovy);
    cc.limx = 1.3f * tanfovx; cc.limy = 1.3f * tanfovy;
    cc.mod = mod; cc.sh_deg = sh_deg; cc.n_coeffs = n_coeffs;
    return cc;
}

// mirrors tile_range_ni + for_each_tile of fsgs_kernels_pre.cuh
template <typename F>
int for_each_tile(const CamConst &cc, const Rec &r, bool no_cull, F &&f) {
    const float rad = (float)r.radius;
    const int rminx = clampi((int)((r.x - rad) / TILE), 0, cc.gx), rminy = clampi((int)((r.y - rad) / TILE), 0, cc.gy);
    const int rmaxx = clampi((int)((r.x + rad + TILE - 1) / TILE), 0, cc.gx);
    const int rmaxy = clampi((int)((r.y + rad + TILE - 1) / TILE), 0, cc.gy);
    int x0 = rminx, y0 = rminy, x1 = rmaxx, y1 = rmaxy;
    CullEllipse e = make_cull_ellipse(r.x, r.y, r.A, r.B, r.C, r.o);
    if (no_cull) e.tau = 3.0e38f;
    else cull_rect(e, rminx, rminy, rmaxx, rmaxy, x0, y0, x1, y1);
    int n = 0;
    for (int ty = y0; ty < y1; ++ty)
        for (int tx = x0; tx < x1; ++tx)
            if (tile_hit(e, tx, ty)) { f(ty * cc.gx + tx); ++n; }
    return n;
}

struct Lists {
    std::vector<std::vector<uint64_t>> per_tile;   // (depth bits << 32 | id), sorted
    int64_t R = 0, Rrect = 0;
};

Lists bin_and_sort(const CamConst &cc, const std::vector<Rec> &recs, bool no_cull) {
    Lists L;
    L.per_tile.resize((size_t)cc.gx * cc.gy);
    for (size_t i = 0; i < recs.size(); ++i) {
        const Rec &r = recs[i];
        if (r.radius <= 0) continue;
        uint32_t db;
        std::memcpy(&db, &r.z, 4);
        const uint64_t key = ((uint64_t)db << 32) | (uint32_t)i;
        L.R += for_each_tile(cc, r, no_cull, [&](int t) { L.per_tile[t].push_back(key); });
        Rec full = r;
        L.Rrect += for_each_tile(cc, full, true, [&](int) {});
    }
    for (auto &v : L.per_tile) std::sort(v.begin(), v.end());
    return L;
}

// What k_tile_sort's epilogue writes per (tile, Gaussian) instance.
struct SRec {
    float x, y, a2, b2, c2, o, r, g, b, z;
    unsigned mask;
    uint32_t id;
};
std::vector<std::vector<SRec>> sorted_records(const CamConst &cc, const std::vector<Rec> &recs, const Lists &L, bool no_cull) {
    std::vector<std::vector<SRec>> out(L.per_tile.size());
    for (size_t t = 0; t < L.per_tile.size(); ++t) {
        const int tx0 = (int)(t % cc.gx) * TILE, ty0 = (int)(t / cc.gx) * TILE;
        for (uint64_t key : L.per_tile[t]) {
            const Rec &r = recs[(uint32_t)key];
            SRec s;
            s.x = r.x; s.y = r.y; s.o = r.o; s.r = r.r; s.g = r.g; s.b = r.b; s.z = r.z; s.id = (uint32_t)key;
            scale_conic(r.A, r.B, r.C, s.a2, s.b2, s.c2);
            s.mask = no_cull ? 0xffu : block_mask(r.x, r.y, r.A, r.B, r.C, r.o, tx0, ty0);
            out[t].push_back(s);
        }
    }
    return out;
}
inline unsigned warp_bit_of(int px, int py) { return 1u << ((((py % TILE) / 4) << 1) | ((px % TILE) / 8)); }

// composite forward for all pixels; planes = FUSED ? 6 : 3 (+ depth plane for the API flavour)
template <bool FUSED>
void composite_fwd(const CamConst &cc, const std::vector<std::vector<SRec>> &SR, const float *bg, float *planes,
                   float *depth, std::vector<float> &final_T, std::vector<int> &n_contrib) {
    const size_t HW = (size_t)cc.W * cc.H;
    final_T.assign(HW, 1.f);
    n_contrib.assign(HW, 0);
    for (int py = 0; py < cc.H; ++py)
        for (int px = 0; px < cc.W; ++px) {
            const auto &lst = SR[(size_t)(py / TILE) * cc.gx + px / TILE];
            const unsigned wb = warp_bit_of(px, py);
            float T = 1.f, C0 = 0, C1 = 0, C2 = 0, D = 0, S = 0, D2 = 0;
            int last = 0;
            for (size_t j = 0; j < lst.size(); ++j) {
                const SRec &r = lst[j];
                if (!(r.mask & wb)) continue;
                const float dx = r.x - (float)px, dy = r.y - (float)py;
                const float p2 = gauss_power2(r.a2, r.b2, r.c2, dx, dy);
                const float alpha = fminf(ALPHA_MAX, r.o * fast_exp2(p2));
                if (!(p2 <= 0.f && alpha >= ALPHA_MIN)) continue;
                const float test_T = T * (1.f - alpha);
                if (test_T < T_MIN) break;
                const float w = alpha * T;
                C0 += r.r * w; C1 += r.g * w; C2 += r.b * w; D += r.z * w;
                if (FUSED) { S += w; D2 += r.z * r.z * w; }
                T = test_T;
                last = (int)j + 1;
            }
            const size_t p = (size_t)py * cc.W + px;
            final_T[p] = T; n_contrib[p] = last;
            planes[p] = C0 + T * bg[0]; planes[HW + p] = C1 + T * bg[1]; planes[2 * HW + p] = C2 + T * bg[2];
            if (FUSED) {
                planes[3 * HW + p] = D + T * bg[0]; planes[4 * HW + p] = S + T * bg[1]; planes[5 * HW + p] = D2 + T * bg[2];
            } else {
                depth[p] = D;
            }
        }
}

template <bool FUSED>
void composite_bwd(const CamConst &cc, size_t P, const std::vector<std::vector<SRec>> &SR, const float *bg,
                   const std::vector<float> &final_T, const std::vector<int> &n_contrib, const float *dplanes,
                   const float *ddepth, std::vector<float> &acc) {
    const size_t HW = (size_t)cc.W * cc.H;
    acc.assign(P * 12, 0.f);
    std::vector<double> acc64(P * 12, 0.0);   // order-independent accumulation of the finalised rows
    const float kx = 0.5f * cc.W, ky = 0.5f * cc.H;
    // per 8x4 warp block: which upstream planes are non-zero (the kernel's warp-uniform `level`)
    const int bw = (cc.W + 7) / 8, bh = (cc.H + 3) / 4;
    std::vector<int> block_level((size_t)bw * bh, 0);
    for (int py = 0; py < cc.H; ++py)
        for (int px = 0; px < cc.W; ++px) {
            const size_t p = (size_t)py * cc.W + px;
            int lv = 0;
            if (FUSED) {
                if (dplanes[3 * HW + p] != 0.f) lv = 1;
                if (dplanes[4 * HW + p] != 0.f || dplanes[5 * HW + p] != 0.f) lv = 2;
            } else if (ddepth && ddepth[p] != 0.f) {
                lv = 1;
            }
            int &b = block_level[(size_t)(py / 4) * bw + px / 8];
            b = std::max(b, lv);
        }
    for (size_t t = 0; t < SR.size(); ++t) {
        const auto &lst = SR[t];
        if (lst.empty()) continue;
        std::vector<double> mom(lst.size() * 12, 0.0);      // per (tile, entry) moments, as s_acc
        const int tx0 = (int)(t % cc.gx) * TILE, ty0 = (int)(t / cc.gx) * TILE;
        for (int py = ty0; py < ty0 + TILE && py < cc.H; ++py)
            for (int px = tx0; px < tx0 + TILE && px < cc.W; ++px) {
                const size_t p = (size_t)py * cc.W + px;
                const unsigned wb = warp_bit_of(px, py);
                float g[6] = {dplanes[p], dplanes[HW + p], dplanes[2 * HW + p], 0, 0, 0};
                float bgdot_dep = 0.f;
                if (FUSED) {
                    g[3] = dplanes[3 * HW + p]; g[4] = dplanes[4 * HW + p]; g[5] = dplanes[5 * HW + p];
                    bgdot_dep = bg[0] * g[3] + bg[1] * g[4] + bg[2] * g[5];
                } else {
                    g[3] = ddepth ? ddepth[p] : 0.f;
                }
                const float bgdot_rgb = bg[0] * g[0] + bg[1] * g[1] + bg[2] * g[2];
                BwdPixel s;
                std::memset(&s, 0, sizeof(s));
                s.T = final_T[p];
                for (int j = n_contrib[p] - 1; j >= 0; --j) {
                    const SRec &r = lst[j];
                    if (!(r.mask & wb)) continue;
                    const float dx = r.x - (float)px, dy = r.y - (float)py;
                    const float p2 = gauss_power2(r.a2, r.b2, r.c2, dx, dy);
                    const float G = fast_exp2(p2);
                    const float alpha = fminf(ALPHA_MAX, r.o * G);
                    if (!(p2 <= 0.f && alpha >= ALPHA_MIN)) continue;
                    float v[12];
                    const int level = block_level[(size_t)(py / 4) * ((cc.W + 7) / 8) + px / 8];
                    if (level == 0)
                        bwd_pair2<FUSED, 0>(s, r.o, r.r, r.g, r.b, r.z, dx, dy, G, alpha, g, final_T[p], bgdot_rgb, bgdot_dep, v);
                    else if (!FUSED || level == 1)
                        bwd_pair2<FUSED, 1>(s, r.o, r.r, r.g, r.b, r.z, dx, dy, G, alpha, g, final_T[p], bgdot_rgb, bgdot_dep, v);
                    else
                        bwd_pair2<FUSED, 2>(s, r.o, r.r, r.g, r.b, r.z, dx, dy, G, alpha, g, final_T[p], bgdot_rgb, bgdot_dep, v);
                    for (int k = 0; k < 12; ++k) mom[(size_t)j * 12 + k] += v[k];
                }
            }
        for (size_t j = 0; j < lst.size(); ++j) {
            float m[12], o[12];
            bool nz = false;
            for (int k = 0; k < 12; ++k) { m[k] = (float)mom[j * 12 + k]; nz |= m[k] != 0.f; }
            if (!nz) continue;
            bwd_finalize(m, lst[j].a2, lst[j].b2, lst[j].c2, lst[j].o, kx, ky, FUSED, o);
            for (int k = 0; k < 12; ++k) acc64[(size_t)lst[j].id * 12 + k] += o[k];
        }
    }
    for (size_t k = 0; k < acc.size(); ++k) acc[k] = (float)acc64[k];
}

}  // namespace

extern "C" {

// Fused render: forward (+ backward when dplanes != null).  Pointer semantics as fsgs_render_*.
int emul_render_fused(int P, int W, int H, float tanfovx, float tanfovy, float mod, int sh_deg, const float *bg,
                      const float *xyz, const float *f_dc, const float *f_rest, const float *op_raw,
                      const float *sc_raw, const float *rot_raw, const float *pose, const float *cam_center,
                      const float *V, const float *PM, int no_cull, float *planes, int *radii, int64_t *R,
                      int64_t *Rrect, const float *dplanes, int gs_grad, int cam_grad, float *dxyz, float *dfdc,
                      float *dfrest, float *dop, float *dsc, float *drot, float *dpose, float *dm2d,
                      float *dpose_only) {
    const CamConst cc = make_cc(W, H, tanfovx, tanfovy, mod, sh_deg, 16);
    std::vector<Rec> recs((size_t)P);
    std::vector<uint8_t> clamp((size_t)P, 0);
    for (int i = 0; i < P; ++i) {
        Splat sp;
        float rgb[3], opacity = 0.f;
        Rec &r = recs[i];
        std::memset(&r, 0, sizeof(r));
        if (fused_forward_one(cc, V, PM, pose, cam_center, xyz + 3 * i, f_dc + 3 * i, f_rest + 45 * i, op_raw[i],
                              sc_raw + 3 * i, rot_raw + 4 * i, sp, opacity, rgb, clamp[i])) {
            r = Rec{sp.px, sp.py, sp.conx, sp.cony, sp.conz, opacity, rgb[0], rgb[1], rgb[2], sp.depth, sp.radius, 0};
        }
        radii[i] = r.radius;
    }
    const Lists L = bin_and_sort(cc, recs, no_cull != 0);
    *R = L.R; *Rrect = L.Rrect;
    std::vector<float> final_T;
    std::vector<int> n_contrib;
    const auto SR = sorted_records(cc, recs, L, no_cull != 0);
    composite_fwd<true>(cc, SR, bg, planes, nullptr, final_T, n_contrib);
    if (!dplanes) return 0;
    std::vector<float> acc;
    composite_bwd<true>(cc, (size_t)P, SR, bg, final_T, n_contrib, dplanes, nullptr, acc);
    double pose_acc[12] = {0}, pose_only_acc[12] = {0};
    for (int i = 0; i < P; ++i) {
        if (dpose_only && cam_grad && recs[i].radius > 0) {
            // what the POSE_ONLY compositor leaves in the row: mean2D, conic and view-depth columns only
            float a[12] = {0}, pg1[12] = {0};
            const float *full = acc.data() + 12 * (size_t)i;
            a[0] = full[0]; a[1] = full[1]; a[2] = full[2]; a[3] = full[3]; a[4] = full[4]; a[9] = full[9];
            fused_backward_pose_one(cc, V, PM, pose, xyz + 3 * i, sc_raw + 3 * i, rot_raw + 4 * i, a, pg1);
            for (int k = 0; k < 12; ++k) pose_only_acc[k] += pg1[k];
        }
        float dx3[3] = {0, 0, 0}, dd[3] = {0, 0, 0}, ds3[3] = {0, 0, 0}, dq4[4] = {0, 0, 0, 0}, pg[12] = {0}, m2[2] = {0, 0};
        float dopv = 0.f;
        float *drest = dfrest + 45 * i;
        for (int k = 0; k < 45; ++k) drest[k] = 0.f;
        if (recs[i].radius > 0)
            fused_backward_one(cc, V, PM, pose, cam_center, xyz + 3 * i, f_dc + 3 * i, f_rest + 45 * i, op_raw[i],
                               sc_raw + 3 * i, rot_raw + 4 * i, clamp[i], acc.data() + 12 * (size_t)i, gs_grad, cam_grad,
                               dx3, dd, drest, dopv, ds3, dq4, pg, m2);
        for (int k = 0; k < 3; ++k) { dxyz[3 * i + k] = dx3[k]; dfdc[3 * i + k] = dd[k]; dsc[3 * i + k] = ds3[k]; }
        for (int k = 0; k < 4; ++k) drot[4 * i + k] = dq4[k];
        dop[i] = dopv;
        dm2d[3 * i] = m2[0]; dm2d[3 * i + 1] = m2[1]; dm2d[3 * i + 2] = 0.f;
        for (int k = 0; k < 12; ++k) pose_acc[k] += pg[k];
    }
    for (int k = 0; k < 16; ++k) dpose[k] = k < 12 ? (float)pose_acc[k] : 0.f;
    if (dpose_only) for (int k = 0; k < 16; ++k) dpose_only[k] = k < 12 ? (float)pose_only_acc[k] : 0.f;
    return 0;
}

// API rasteriser: forward (+ backward when dcolor != null).
int emul_rasterize_api(int P, int W, int H, float tanfovx, float tanfovy, float mod, int sh_deg, int n_coeffs,
                       const float *bg, const float *means3D, const float *colors, const float *shs,
                       const float *opacities, const float *scales, const float *rots, const float *cov3D,
                       const float *V, const float *PM, const float *campos, int no_cull, float *out_color,
                       float *out_depth, int *radii, int64_t *R, int64_t *Rrect, const float *dcolor,
                       const float *ddepth, float *dm2d, float *dcolors, float *dopac, float *dmeans3D, float *dcov,
                       float *dsh, float *dscales, float *drots) {
    const CamConst cc = make_cc(W, H, tanfovx, tanfovy, mod, sh_deg, n_coeffs);
    std::vector<Rec> recs((size_t)P);
    std::vector<uint8_t> clamp((size_t)P, 0);
    const float zero3[3] = {0, 0, 0};
    for (int i = 0; i < P; ++i) {
        Splat sp;
        float rgb[3];
        Rec &r = recs[i];
        std::memset(&r, 0, sizeof(r));
        if (api_forward_one(cc, V, PM, campos ? campos : zero3, means3D + 3 * i, colors ? colors + 3 * i : nullptr,
                            shs ? shs + (size_t)i * n_coeffs * 3 : nullptr, scales ? scales + 3 * i : nullptr,
                            rots ? rots + 4 * i : nullptr, cov3D ? cov3D + 6 * i : nullptr, sp, rgb, clamp[i])) {
            r = Rec{sp.px, sp.py, sp.conx, sp.cony, sp.conz, opacities[i], rgb[0], rgb[1], rgb[2], sp.depth, sp.radius, 0};
        }
        radii[i] = r.radius;
    }
    const Lists L = bin_and_sort(cc, recs, no_cull != 0);
    *R = L.R; *Rrect = L.Rrect;
    std::vector<float> final_T;
    std::vector<int> n_contrib;
    const auto SR = sorted_records(cc, recs, L, no_cull != 0);
    composite_fwd<false>(cc, SR, bg, out_color, out_depth, final_T, n_contrib);
    if (!dcolor) return 0;
    std::vector<float> acc;
    composite_bwd<false>(cc, (size_t)P, SR, bg, final_T, n_contrib, dcolor, ddepth, acc);
    for (int i = 0; i < P; ++i) {
        float dmean[3] = {0, 0, 0}, dc6[6] = {0, 0, 0, 0, 0, 0}, ds[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0};
        const float *a = acc.data() + 12 * (size_t)i;
        float *dsh_i = (shs && dsh) ? dsh + (size_t)i * n_coeffs * 3 : nullptr;
        if (dsh_i) for (int k = 0; k < n_coeffs * 3; ++k) dsh_i[k] = 0.f;
        if (recs[i].radius > 0)
            api_backward_one(cc, V, PM, campos ? campos : zero3, means3D + 3 * i, shs ? shs + (size_t)i * n_coeffs * 3 : nullptr,
                             scales ? scales + 3 * i : nullptr, rots ? rots + 4 * i : nullptr,
                             cov3D ? cov3D + 6 * i : nullptr, clamp[i], a, dmean, dc6, dsh_i, ds, dq);
        const bool vis = recs[i].radius > 0;
        dm2d[3 * i] = vis ? a[0] : 0.f; dm2d[3 * i + 1] = vis ? a[1] : 0.f; dm2d[3 * i + 2] = 0.f;
        for (int k = 0; k < 3; ++k) { dcolors[3 * i + k] = vis ? a[6 + k] : 0.f; dmeans3D[3 * i + k] = dmean[k]; dscales[3 * i + k] = ds[k]; }
        dopac[i] = vis ? a[5] : 0.f;
        for (int k = 0; k < 6; ++k) dcov[6 * i + k] = dc6[k];
        for (int k = 0; k < 4; ++k) drots[4 * i + k] = dq[k];
    }
    return 0;
}

}  // extern "C"

extern "C" void emul_pose(const float *r_raw, const float *t, float *Rt, const float *dRt, float *dr, float *dt) {
    fsgs::pose_forward(r_raw, t, Rt);
    if (dRt) fsgs::pose_backward(r_raw, dRt, dr, dt);
}
