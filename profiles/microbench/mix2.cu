// Instruction-mix floor of the symmetric pair kernel's planar inner loop (profiles/microbench, not product code).
// One "group" = one packed pair_terms2 call (two pair terms per lane): NF2 packed FP32x2 ops (FFMA2/FMUL2/FADD2 at the
// kernel's 17:16:8 ratio), NF1 predicated scalar FADDs, NMU MUFUs, NLOP LOP3s, NSETP FSETPs, NMNMX FMNMXs and an LDS.128
// every fourth group -- all on independent chains, so the result is the issue / pipe floor of the mix without any of
// the kernel's data dependencies.  Ablations show what each instruction class costs next to the packed stream.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
constexpr int ITERS = 1024;
template <int NF2, int NF1, int NMU, int NLOP, int NSETP, int NMNMX, int LDS>
__global__ void __launch_bounds__(128) mix2(float* out, float seed) {
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    sm[threadIdx.x + 128] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    u64 p[8]; float m[4], al[4], f1[4];
    for (int c = 0; c < 4; ++c) {
        float lo = seed + c, hi = seed * 0.5f + threadIdx.x * 1e-3f;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p[c]) : "f"(lo), "f"(hi));
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p[c + 4]) : "f"(hi), "f"(lo));
        m[c] = seed + 0.25f * c; al[c] = seed * c; f1[c] = seed - c;
    }
    u64 k1, k2; float a = 0.9999f + seed * 1e-7f, b = 1e-4f;
    asm volatile("mov.b64 %0, {%1,%1};" : "=l"(k1) : "f"(a)); asm volatile("mov.b64 %0, {%1,%1};" : "=l"(k2) : "f"(b));
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < NF2; ++k) {
                const int kind = (k * 41 / (NF2 > 0 ? NF2 : 1)) ;      // 0..40 -> 17 fma, 16 mul, 8 add
                if (kind < 17) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c + 4 * (k & 1)]) : "l"(k1), "l"(k2));
                else if (kind < 33) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c + 4 * (k & 1)]) : "l"(k1));
                else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c + 4 * (k & 1)]) : "l"(k2));
            }
#pragma unroll
            for (int k = 0; k < NMU; ++k) {
                if (k & 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
                else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
            }
#pragma unroll
            for (int k = 0; k < NLOP; ++k) { unsigned u = __float_as_uint(al[c]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(0x80000000u + k), "r"(__float_as_uint(f1[c]))); al[c] = __uint_as_float(u); }
#pragma unroll
            for (int k = 0; k < NSETP; ++k) {
                if (NF1 > 0 && k < NF1)
                    asm volatile("{.reg .pred pp; setp.gt.f32 pp, %1, %2; @pp add.f32 %0, %0, %3;}" : "+f"(f1[c]) : "f"(al[c]), "f"(m[(c + k) & 3]), "f"(a));
                else
                    asm volatile("{.reg .pred pp; setp.gt.f32 pp, %1, %2; @pp mov.b32 %0, %3;}" : "+f"(f1[c]) : "f"(al[c]), "f"(m[(c + k) & 3]), "f"(a));
            }
            if (NSETP == 0) {
#pragma unroll
                for (int k = 0; k < NF1; ++k) asm volatile("add.f32 %0, %0, %1;" : "+f"(f1[c]) : "f"(a));
            }
#pragma unroll
            for (int k = 0; k < NMNMX; ++k) asm volatile("min.f32 %0, %0, %1;" : "+f"(al[c]) : "f"(m[c]));
        }
        if (LDS) {
            float4 v;
            const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + it) & 255]);
#pragma unroll
            for (int k = 0; k < LDS; ++k) {
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr ^ (k * 16)));
                al[k & 3] += v.x * 0.0f;
            }
        }
    }
    float acc = 0;
    for (int c = 0; c < 4; ++c) {
        float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[c])); acc += lo + hi;
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[c + 4])); acc += lo + hi + m[c] + al[c] + f1[c];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NF2, int NF1, int NMU, int NLOP, int NSETP, int NMNMX, int LDS>
void run(const char* name, float* out, int ctas_per_sm) {
    int grid = 148 * ctas_per_sm;
    mix2<NF2, NF1, NMU, NLOP, NSETP, NMNMX, LDS><<<grid, 128>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) mix2<NF2, NF1, NMU, NLOP, NSETP, NMNMX, LDS><<<grid, 128>>>(out, 1.0f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double groups = (double)ITERS * 4 * (128 / 32) * ctas_per_sm / 4.0;   // groups per SMSP
    double cyc = ms * 1e-3 * 1.965e9 / groups;
    printf("%-52s ctas/SM=%d: %.3f ms -> %6.1f cycles per group per SMSP (FMA-pipe %d, XU %d, issue %.1f)\n", name, ctas_per_sm, ms, cyc,
           NF2 * 2 + NF1, NMU * 8, NF2 + NF1 + NMU + NLOP + NSETP + NMNMX + LDS / 4.0);
}
int main() {
    float* out; cudaMalloc(&out, sizeof(float) * 128 * 148 * 16);
    for (int occ : {5}) {
        //   F2  F1 MU LOP SETP MNMX LDS
        run<41, 4, 8, 8, 4, 2, 8>("kernel mix (asin6): 41 F2 4 FADD 8 MUFU 8 LOP 4 SETP 2 MNMX", out, occ);
        run<41, 4, 8, 8, 4, 2, 0>("  - LDS", out, occ);
        run<41, 0, 8, 8, 4, 2, 0>("  - LDS - FADD", out, occ);
        run<41, 4, 8, 0, 4, 2, 0>("  - LDS - LOP3", out, occ);
        run<41, 4, 8, 8, 0, 2, 0>("  - LDS - FSETP (FADD unpredicated)", out, occ);
        run<41, 4, 0, 8, 4, 2, 0>("  - LDS - MUFU", out, occ);
        run<41, 4, 4, 8, 4, 2, 0>("  - LDS - 4 MUFU", out, occ);
        run<41, 0, 8, 0, 0, 0, 0>("41 F2 + 8 MUFU", out, occ);
        run<41, 0, 0, 0, 0, 0, 0>("41 F2", out, occ);
        run<41, 0, 0, 8, 4, 2, 0>("41 F2 + ALU only (8 LOP 4 SETP 2 MNMX)", out, occ);
        run<43, 4, 10, 8, 4, 4, 8>("previous mix (rcp): 43 F2 4 FADD 10 MUFU 8 LOP 4 SETP 4 MNMX", out, occ);
        run<37, 0, 8, 6, 2, 2, 8>("hypothetical: 37 F2 0 FADD 8 MUFU 6 LOP 2 SETP 2 MNMX", out, occ);
    }
    return 0;
}
