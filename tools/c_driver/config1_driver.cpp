// config1_driver.cpp -- the "CPU-side driver" of BASELINE.json configs[0]: one fused render
// forward + backward through the C ABI of libfsgs_raster.so with NO Python and NO torch in the process --
// plain cudaMalloc'd buffers, malloc-style allocation callbacks, one stream -- checked against expected results
// read from a flat binary file (written by tests/test_gpu_c_driver.py from the float64 oracle).
//
//   g++ -O2 -std=c++17 config1_driver.cpp -I../../include -I/usr/local/cuda/include -L<dir of libfsgs_raster.so>
//       -lfsgs_raster -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN' -o fsgs_config1_driver
//   ./fsgs_config1_driver case.bin            -> one JSON line on stdout, exit code 0 iff every gate holds
//
// File layout (little endian): char magic[8] = "FSGSC1\0\0"; int32 P, W, H, sh_degree; float tanfovx, tanfovy;
// then float32 arrays in this order:
//   bg[3] r[4] t[3] cam_center[3] viewmatrix[16] projmatrix[16]
//   xyz[3P] f_dc[3P] f_rest[45P] opacity[P] scaling[3P] rotation[4P]
//   G[4HW]                      upstream gradient of (rgb[3], depth)
//   expected: planes[6HW]  mask[HW] (1 = fragile pixel, not compared / carries no upstream gradient)
//             g_xyz[3P] g_f_dc[3P] g_f_rest[45P] g_opacity[P] g_scaling[3P] g_rotation[4P] g_r[4] g_t[3]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fsgs_raster.h"

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                      \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)
#define FS(call)                                                                              \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ != FSGS_OK) {                                                                 \
            fprintf(stderr, "%s: %d %s\n", #call, rc_, fsgs_error_string(rc_));              \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)

static std::vector<void *> g_allocs;
static void *dev_alloc(void *, size_t bytes) {          // the allocation callback: scratch owned by this driver
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 256) != cudaSuccess) return nullptr;
    g_allocs.push_back(p);
    return p;
}

struct Reader {
    FILE *f;
    bool ok = true;
    std::vector<float> arr(size_t n) {
        std::vector<float> v(n);
        if (fread(v.data(), sizeof(float), n, f) != n) ok = false;
        return v;
    }
};

static float *upload(const std::vector<float> &h) {
    float *d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(float) + 256) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
    return d;
}
static std::vector<float> download(const float *d, size_t n) {
    std::vector<float> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost);
    return h;
}
static double rel_err(const std::vector<float> &a, const std::vector<float> &b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        const double d = (double)a[i] - (double)b[i];
        num += d * d;
        den += (double)b[i] * (double)b[i];
    }
    return std::sqrt(num) / std::sqrt(den > 1e-60 ? den : 1e-60);
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s case.bin\n", argv[0]); return 64; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 65; }
    char magic[8];
    int32_t hdr[4];
    float fov[2];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "FSGSC1\0\0", 8) != 0 || fread(hdr, 4, 4, f) != 4 ||
        fread(fov, 4, 2, f) != 2) { fprintf(stderr, "bad header\n"); return 66; }
    const int P = hdr[0], W = hdr[1], H = hdr[2], deg = hdr[3];
    const size_t HW = (size_t)W * H, Ps = (size_t)P;
    Reader rd{f};
    auto bg = rd.arr(3), r = rd.arr(4), t = rd.arr(3), cc = rd.arr(3), view = rd.arr(16), proj = rd.arr(16);
    auto xyz = rd.arr(3 * Ps), fdc = rd.arr(3 * Ps), frest = rd.arr(45 * Ps), opa = rd.arr(Ps), sca = rd.arr(3 * Ps),
         rot = rd.arr(4 * Ps), G = rd.arr(4 * HW);
    auto e_planes = rd.arr(6 * HW), e_mask = rd.arr(HW);
    auto e_xyz = rd.arr(3 * Ps), e_fdc = rd.arr(3 * Ps), e_frest = rd.arr(45 * Ps), e_opa = rd.arr(Ps),
         e_sca = rd.arr(3 * Ps), e_rot = rd.arr(4 * Ps), e_r = rd.arr(4), e_t = rd.arr(3);
    fclose(f);
    if (!rd.ok) { fprintf(stderr, "truncated case file\n"); return 67; }

    if (fsgs_abi_version() != FSGS_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 68; }
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));
    float *d_bg = upload(bg), *d_r = upload(r), *d_t = upload(t), *d_cc = upload(cc), *d_view = upload(view),
          *d_proj = upload(proj), *d_xyz = upload(xyz), *d_fdc = upload(fdc), *d_frest = upload(frest), *d_opa = upload(opa),
          *d_sca = upload(sca), *d_rot = upload(rot), *d_G = upload(G);
    float *d_Rt, *d_planes, *d_gpose, *d_dr, *d_dt, *d_gx, *d_gfdc, *d_gfrest, *d_gopa, *d_gsca, *d_grot, *d_m2d;
    int32_t *d_radii;
    void *d_scratch;
    CK(cudaMalloc(&d_Rt, 64)); CK(cudaMalloc(&d_planes, 6 * HW * 4)); CK(cudaMalloc(&d_gpose, 64));
    CK(cudaMalloc(&d_dr, 16)); CK(cudaMalloc(&d_dt, 12)); CK(cudaMalloc(&d_radii, Ps * 4 + 256));
    CK(cudaMalloc(&d_gx, 3 * Ps * 4 + 256)); CK(cudaMalloc(&d_gfdc, 3 * Ps * 4 + 256)); CK(cudaMalloc(&d_gfrest, 45 * Ps * 4 + 256));
    CK(cudaMalloc(&d_gopa, Ps * 4 + 256)); CK(cudaMalloc(&d_gsca, 3 * Ps * 4 + 256)); CK(cudaMalloc(&d_grot, 4 * Ps * 4 + 256));
    CK(cudaMalloc(&d_m2d, 3 * Ps * 4 + 256)); CK(cudaMalloc(&d_scratch, fsgs_grad_scratch_bytes(P)));

    fsgs_settings st{};
    st.image_height = H; st.image_width = W; st.tanfovx = fov[0]; st.tanfovy = fov[1]; st.scale_modifier = 1.0f;
    st.sh_degree = deg; st.n_coeffs = 16; st.debug = 1; st.flags = 0;

    // LearnPose.forward -> render forward -> render backward -> LearnPose backward: the whole hot path
    FS(fsgs_pose_forward(d_r, d_t, 0, 1, d_Rt, stream));
    int64_t n_inst = 0, n_rect = 0;
    const size_t before = g_allocs.size();
    FS(fsgs_render_forward(&st, P, d_bg, d_xyz, d_fdc, d_frest, d_opa, d_sca, d_rot, d_Rt, d_cc, d_view, d_proj, dev_alloc,
                           nullptr, dev_alloc, nullptr, dev_alloc, nullptr, d_planes, d_radii, &n_inst, &n_rect, stream));
    if (g_allocs.size() < before + 3) { fprintf(stderr, "allocation callbacks not used\n"); return 69; }
    // the callbacks are invoked in the order geometry, image state, binning (binning possibly twice: optimistic + exact)
    void *geom = g_allocs[before], *img = g_allocs[before + 1], *binning = g_allocs.back();
    FS(fsgs_render_backward_ex(&st, P, n_inst, d_bg, d_xyz, d_fdc, d_frest, d_opa, d_sca, d_rot, d_Rt, d_cc, d_view, d_proj,
                               geom, binning, img, d_G, d_G + 3 * HW, nullptr, nullptr, d_scratch, 1, 1, d_gx, d_gfdc,
                               d_gfrest, d_gopa, d_gsca, d_grot, d_gpose, d_m2d, nullptr, stream));
    FS(fsgs_pose_backward(d_r, 0, 1, d_gpose, d_dr, d_dt, stream));
    CK(cudaStreamSynchronize(stream));
    if (fsgs_watchdog_flag(0, 0) != 0) { fprintf(stderr, "device watchdog fired\n"); return 70; }

    // ---- gates: <= 1e-5 abs on the planes outside the fragile mask (2e-5 depth / silhouette, 4e-5 depth^2),
    //            <= 1e-4 relative on every gradient
    auto planes = download(d_planes, 6 * HW);
    const double tol[6] = {1e-5, 1e-5, 1e-5, 2e-5, 2e-5, 4e-5};
    double worst = 0;
    long n_bad = 0, n_masked = 0;
    for (size_t p = 0; p < HW; ++p) {
        if (e_mask[p] != 0.f) { ++n_masked; continue; }
        for (int c = 0; c < 6; ++c) {
            const double e = std::fabs((double)planes[c * HW + p] - (double)e_planes[c * HW + p]) / tol[c];
            if (e > worst) worst = e;
            if (e > 1.0) ++n_bad;
        }
    }
    struct { const char *name; double err; } gr[] = {
        {"xyz", rel_err(download(d_gx, 3 * Ps), e_xyz)}, {"f_dc", rel_err(download(d_gfdc, 3 * Ps), e_fdc)},
        {"f_rest", rel_err(download(d_gfrest, 45 * Ps), e_frest)}, {"opacity", rel_err(download(d_gopa, Ps), e_opa)},
        {"scaling", rel_err(download(d_gsca, 3 * Ps), e_sca)}, {"rotation", rel_err(download(d_grot, 4 * Ps), e_rot)},
        {"dL/dr", rel_err(download(d_dr, 4), e_r)}, {"dL/dt", rel_err(download(d_dt, 3), e_t)}};
    bool ok = n_bad == 0;
    std::string js = "{\"driver\": \"config1 C driver (no Python, no torch)\", \"P\": " + std::to_string(P) + ", \"W\": " +
                     std::to_string(W) + ", \"H\": " + std::to_string(H) + ", \"tile_instances\": " + std::to_string(n_inst) +
                     ", \"tile_instances_rect\": " + std::to_string(n_rect) + ", \"masked_pixels\": " + std::to_string(n_masked) +
                     ", \"pixels_above_gate\": " + std::to_string(n_bad) + ", \"worst_plane_error_in_gates\": " + std::to_string(worst) +
                     ", \"grad_rel_err\": {";
    for (size_t k = 0; k < sizeof(gr) / sizeof(gr[0]); ++k) {
        char buf[96];
        snprintf(buf, sizeof(buf), "%s\"%s\": %.3e", k ? ", " : "", gr[k].name, gr[k].err);
        js += buf;
        if (!(gr[k].err <= 1e-4)) ok = false;
    }
    js += std::string("}, \"ok\": ") + (ok ? "true" : "false") + "}";
    printf("%s\n", js.c_str());
    for (void *p : g_allocs) cudaFree(p);
    return ok ? 0 : 1;
}
