// Mixed-pipe microbenchmark: packed FFMA2 + MUFU (+ ALU) at the ratio of the pair kernel, independent chains.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
constexpr int ITERS = 2048;
template <int NF2, int NMUFU, int NALU, int NF1, int NMUFU_PURE = 0, int NLOP = 0>
__global__ void __launch_bounds__(128) mix(float* out, float seed) {
    u64 p[4]; float m[4], al[4], f1[4];
    for (int c = 0; c < 4; ++c) { float lo = seed + c, hi = seed * 0.5f + threadIdx.x * 1e-3f; asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p[c]) : "f"(lo), "f"(hi)); m[c] = seed + 0.25f * c; al[c] = seed * c; f1[c] = seed - c; }
    u64 k1, k2; float a = 0.9999f + seed * 1e-7f, b = 1e-4f; asm volatile("mov.b64 %0, {%1,%1};" : "=l"(k1) : "f"(a)); asm volatile("mov.b64 %0, {%1,%1};" : "=l"(k2) : "f"(b));
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < NF2; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(k1), "l"(k2));
#pragma unroll
            for (int k = 0; k < NMUFU; ++k) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c])); asm volatile("add.f32 %0, %0, %1;" : "+f"(m[c]) : "f"(-a)); }
#pragma unroll
            for (int k = 0; k < NALU; ++k) { if (k & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(al[c]) : "f"(m[c])); else asm volatile("min.f32 %0, %0, %1;" : "+f"(al[c]) : "f"(b)); }
#pragma unroll
            for (int k = 0; k < NF1; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f1[c]) : "f"(a), "f"(b));
#pragma unroll
            for (int k = 0; k < NMUFU_PURE; ++k) { if (k & 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[c])); else asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m[c])); }
#pragma unroll
            for (int k = 0; k < NLOP; ++k) { unsigned u = __float_as_uint(al[c]); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u) : "r"(0x80000000u + k)); al[c] = __uint_as_float(u); }
        }
    }
    float acc = 0; for (int c = 0; c < 4; ++c) { float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[c])); acc += lo + hi + m[c] + al[c] + f1[c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NF2, int NMUFU, int NALU, int NF1, int NMUFU_PURE = 0, int NLOP = 0>
void run(const char* name, float* out, int ctas_per_sm) {
    int grid = 148 * ctas_per_sm;
    mix<NF2, NMUFU, NALU, NF1, NMUFU_PURE, NLOP><<<grid, 128>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) mix<NF2, NMUFU, NALU, NF1, NMUFU_PURE, NLOP><<<grid, 128>>>(out, 1.0f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double groups = (double)ITERS * 4 * (128 / 32) * ctas_per_sm / 4.0;   // warp-groups per SMSP
    double cyc = ms * 1e-3 * 1.965e9 / groups;
    printf("%-34s ctas/SM=%d: %.3f ms -> %.1f cycles per group per SMSP (FMA-pipe %d, XU %d, issue %d)\n", name, ctas_per_sm, ms, cyc,
           NF2 * 2 + NF1 + NMUFU, (NMUFU + NMUFU_PURE) * 8, NF2 + NMUFU * 2 + NALU + NF1 + NMUFU_PURE + NLOP);
}
int main() {
    float* out; cudaMalloc(&out, sizeof(float) * 128 * 148 * 16);
    for (int occ : {8}) {
        run<9, 0, 0, 0>("9 FFMA2", out, occ);
        run<9, 0, 0, 0, 2, 0>("9 FFMA2 + 2 MUFU pure", out, occ);
        run<9, 0, 0, 0, 1, 0>("9 FFMA2 + 1 MUFU pure", out, occ);
        run<9, 0, 0, 2, 0, 0>("9 FFMA2 + 2 FFMA scalar", out, occ);
        run<9, 0, 0, 1, 0, 0>("9 FFMA2 + 1 FFMA scalar", out, occ);
        run<9, 0, 3, 0, 0, 0>("9 FFMA2 + 3 FMNMX", out, occ);
        run<9, 0, 0, 0, 0, 3>("9 FFMA2 + 3 LOP3", out, occ);
        run<9, 0, 3, 0, 0, 3>("9 FFMA2 + 3 FMNMX + 3 LOP3", out, occ);
        run<9, 0, 3, 0, 2, 3>("9 FFMA2 + 2 MUFU pure + 3 FMNMX + 3 LOP3", out, occ);
        run<0, 0, 0, 0, 2, 0>("2 MUFU pure", out, occ);
        run<0, 2, 0, 0>("2 MUFU(+2 FADD)", out, occ);
        run<9, 2, 0, 0>("9 FFMA2 + 2 MUFU(+2 FADD)", out, occ);
        run<9, 2, 3, 0>("9 FFMA2 + 2 MUFU(+2 FADD) + 3 ALU", out, occ);
        run<9, 1, 3, 0>("9 FFMA2 + 1 MUFU(+1 FADD) + 3 ALU", out, occ);
        run<9, 2, 3, 1>("9 FFMA2 + 2 MUFU + 3 ALU + 1 FFMA", out, occ);
        run<0, 0, 3, 9>("9 FFMA + 3 ALU", out, occ);
        run<0, 2, 3, 18>("18 FFMA + 2 MUFU + 3 ALU", out, occ);
    }
    return 0;
}
