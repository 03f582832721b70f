// Microbenchmark: issue rate of scalar FFMA/FMUL/FADD/FSEL vs packed fma/mul/add.f32x2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
// Phase-B-like mix of the backward compositor: per step 2 x LDS.128 (distinct per-lane addresses), 8 multiply-adds of the
// loaded values into accumulators, 4 integer instructions.  MODE 0: scalar FFMA; MODE 1: packed fma.rn.f32x2 on the
// register pairs the 128-bit loads deliver.
template <int MODE>
__global__ void kb(float *out, float a, int stride) {
    __shared__ float4 sm[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_float4(a + i, a * i, a - i, a);
    __syncthreads();
    float x0 = 0, x1 = 0, x2 = 0, x3 = 0, x4 = 0, x5 = 0, x6 = 0, x7 = 0;
    float y0 = 0, y1 = 0, y2 = 0, y3 = 0, y4 = 0, y5 = 0, y6 = 0, y7 = 0;
    const float c0 = a, c1 = a + 1.f, c2 = a + 2.f, c3 = a + 3.f;
    unsigned long long p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0, k01, k23, k10, k32;
    asm("mov.b64 %0, {%1,%2};" : "=l"(k01) : "f"(c0), "f"(c1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(k23) : "f"(c2), "f"(c3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(k10) : "f"(c1), "f"(c0));
    asm("mov.b64 %0, {%1,%2};" : "=l"(k32) : "f"(c3), "f"(c2));
    int idx = threadIdx.x, acc = 0;
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // 1 LDS.128 per 16 multiply-adds and 6 integer instructions: the ratio of the real phase B
            const float4 v = sm[(idx + 32 * u) & 511];      // lane-linear: conflict-free, 4 wavefronts
            idx += stride; acc ^= idx; acc += u; acc = (acc << 1) ^ (idx >> 2); acc += idx & 7; acc ^= acc >> 3;
            if (MODE == 0) {
                x0 = fmaf(v.x, c0, x0); x1 = fmaf(v.y, c1, x1); x2 = fmaf(v.z, c2, x2); x3 = fmaf(v.w, c3, x3);
                x4 = fmaf(v.x, c1, x4); x5 = fmaf(v.y, c0, x5); x6 = fmaf(v.z, c3, x6); x7 = fmaf(v.w, c2, x7);
                y0 = fmaf(v.x, c2, y0); y1 = fmaf(v.y, c3, y1); y2 = fmaf(v.z, c0, y2); y3 = fmaf(v.w, c1, y3);
                y4 = fmaf(v.x, c3, y4); y5 = fmaf(v.y, c2, y5); y6 = fmaf(v.z, c1, y6); y7 = fmaf(v.w, c0, y7);
            } else {
                unsigned long long v01, v23;
                asm("mov.b64 %0, {%1,%2};" : "=l"(v01) : "f"(v.x), "f"(v.y));
                asm("mov.b64 %0, {%1,%2};" : "=l"(v23) : "f"(v.z), "f"(v.w));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p0) : "l"(v01), "l"(k01));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p1) : "l"(v23), "l"(k23));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p2) : "l"(v01), "l"(k10));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p3) : "l"(v23), "l"(k32));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p4) : "l"(v01), "l"(k23));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p5) : "l"(v23), "l"(k01));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p6) : "l"(v01), "l"(k32));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p7) : "l"(v23), "l"(k10));
            }
        }
    }
    float z0, z1, s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7 + (float)acc;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p0)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p1)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p2)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p3)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p4)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p5)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p6)); s += z0 + z1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(z0), "=f"(z1) : "l"(p7)); s += z0 + z1;
    if (s == 12345.678f) out[0] = s;
}
template <int MODE>
void runb(const char *name) {
    float *d; cudaMalloc(&d, 4);
    int blocks = 148 * 4, threads = 256;
    kb<MODE><<<blocks, threads>>>(d, 1.0001f, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kb<MODE><<<blocks, threads>>>(d, 1.0001f, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // clocks per unrolled step per SMSP: 8 warps per SMSP resident (4 CTAs x 8 warps / 4)
    double steps = (double)blocks * threads / 32 * ITER * 8.0;
    printf("%-44s %8.3f ms  %6.2f clk per step per SMSP (1965 MHz)\n", name, ms, ms * 1e-3 * 1.965e9 * 148 * 4 / steps);
    cudaFree(d);
}

template <int MODE>
__global__ void k(float *out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned long long p0, p1, p2, p3, pa, pb;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(x2), "f"(x3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(x4), "f"(x5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(x6), "f"(x7));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) {   // 8 scalar FFMA (3-reg)
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            } else if (MODE == 1) {   // 4 packed FFMA2 = 8 FMAs
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            } else if (MODE == 2) {   // 8 scalar FMUL
                x0 = x0 * a; x1 = x1 * a; x2 = x2 * a; x3 = x3 * a; x4 = x4 * a; x5 = x5 * a; x6 = x6 * a; x7 = x7 * a;
            } else if (MODE == 3) {   // 4 packed FMUL2
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pa));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pa));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
            } else if (MODE == 4) {   // 8 scalar FADD
                x0 = x0 + a; x1 = x1 + a; x2 = x2 + a; x3 = x3 + a; x4 = x4 + a; x5 = x5 + a; x6 = x6 + a; x7 = x7 + a;
            } else if (MODE == 5) {   // 4 packed FADD2
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pa));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pa));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
            } else if (MODE == 6) {   // 4 FFMA + 4 FSEL-ish (alu pipe: FMNMX)
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = fminf(x4, x0); x5 = fminf(x5, x1); x6 = fminf(x6, x2); x7 = fminf(x7, x3);
            } else if (MODE == 7) {   // 2 FFMA2 + 4 FMNMX (operands taken from the packed results: nothing to hoist or merge)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                float u0, u1, u2, u3;
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(u0), "=f"(u1) : "l"(p0));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(u2), "=f"(u3) : "l"(p1));
                x4 = fminf(x4, u0); x5 = fmaxf(x5, u1); x6 = fminf(x6, u2); x7 = fmaxf(x7, u3);
            } else if (MODE == 10) {  // 2 FFMA2 + 2 FSETP/FSEL pairs + 2 MUFU.EX2 (the compositor's mix)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
                float u0, u1, u2, u3;
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(u0), "=f"(u1) : "l"(p0));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(u2), "=f"(u3) : "l"(p1));
                x4 = (u0 > x5) ? u1 : x4; x5 = (u2 > x4) ? u3 : x5;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x6) : "f"(u0 + x6));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x7) : "f"(u2 + x7));
            } else if (MODE == 11) {  // the same work with scalar FFMA: 4 FFMA + 2 select pairs + 2 adds + 2 MUFU.EX2
                x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
                x4 = (x0 > x5) ? x1 : x4; x5 = (x2 > x4) ? x3 : x5;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x6) : "f"(x0 + x6));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x7) : "f"(x2 + x7));
            } else if (MODE == 8) {   // 8 SHFL
                x0 = __shfl_xor_sync(~0u, x0, 1); x1 = __shfl_xor_sync(~0u, x1, 2); x2 = __shfl_xor_sync(~0u, x2, 4);
                x3 = __shfl_xor_sync(~0u, x3, 8); x4 = __shfl_xor_sync(~0u, x4, 16); x5 = __shfl_xor_sync(~0u, x5, 1);
                x6 = __shfl_xor_sync(~0u, x6, 2); x7 = __shfl_xor_sync(~0u, x7, 4);
            } else if (MODE == 9) {   // 4 SHFL + 4 FFMA
                x0 = __shfl_xor_sync(~0u, x0, 1); x1 = __shfl_xor_sync(~0u, x1, 2); x2 = __shfl_xor_sync(~0u, x2, 4);
                x3 = __shfl_xor_sync(~0u, x3, 8);
                x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            }
        }
    }
    float y0, y1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(y0), "=f"(y1) : "l"(p0));
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + y0 + y1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(y0), "=f"(y1) : "l"(p1)); s += y0 + y1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(y0), "=f"(y1) : "l"(p2)); s += y0 + y1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(y0), "=f"(y1) : "l"(p3)); s += y0 + y1;
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char *name, int ops_per_unroll) {
    float *d; cudaMalloc(&d, 4);
    int blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(d, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = (double)blocks * threads / 32 * ITER * 8.0 * ops_per_unroll;
    // per SMSP per clock at 1.965 GHz
    double rate = warp_inst / (ms * 1e-3) / (148.0 * 4) / 1.965e9;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SMSP (assuming 1965 MHz)\n", name, ms, rate);
    cudaFree(d);
}

int main() {
    run<0>("8x FFMA", 8);
    run<1>("4x FFMA2", 4);
    run<2>("8x FMUL", 8);
    run<3>("4x FMUL2", 4);
    run<4>("8x FADD", 8);
    run<5>("4x FADD2", 4);
    run<6>("4 FFMA + 4 FMNMX", 8);
    run<7>("2 FFMA2 + 4 FMNMX", 6);
    run<10>("2 FFMA2 + 2 sel + 2 add + 2 EX2", 10);
    run<11>("4 FFMA  + 2 sel + 2 add + 2 EX2", 12);
    run<8>("8x SHFL", 8);
    run<9>("4 SHFL + 4 FFMA", 8);
    runb<0>("phase-B mix: 1 LDS.128 + 16 FFMA + 6 int");
    runb<1>("phase-B mix: 1 LDS.128 + 8 FFMA2 + 6 int");
    return 0;
}
