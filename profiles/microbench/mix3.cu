// Round-2 microbenchmark: does a packed FP32x2 stream leave FMA capacity that scalar FP32 can use (heavy / lite
// sub-pipes), and what do operand patterns cost?  Rates are reported as lane-operations per clock per SM at the SM clock
// (clock64 inside the kernel), 128 = the nominal FP32 rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix3 mix3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
constexpr int ITERS = 2048;
constexpr int CH = 6;
enum Kind { P_SHARED, P_DISTINCT, S_SHARED, S_DISTINCT, MIX_1P_1S, MIX_2P_1S, MIX_1P_2S, PMUL2, PADD2, MIX_1P_1ALU, MIX_2P_1MUFU, K_COUNT };
static const char* names[] = {"FFMA2 shared operands", "FFMA2 3 distinct operands", "FFMA shared operands", "FFMA 3 distinct operands",
  "FFMA2 + FFMA 1:1", "FFMA2 + FFMA 2:1", "FFMA2 + FFMA 1:2", "FMUL2 (2 operands)", "FADD2 (2 operands)", "FFMA2 + LOP3 1:1", "FFMA2 + MUFU.EX2 2:1"};
// lane-ops per inner step per chain: packed = 2, scalar = 1 (ALU / MUFU counted as 1)
static const double ops[] = {2, 2, 1, 1, 3, 5, 4, 2, 2, 3, 5};
static const double fp32ops[] = {2, 2, 1, 1, 3, 5, 4, 2, 2, 2, 4};

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, unsigned long long* cycles, float seed) {
    unsigned long long pa[CH], pb[CH], pc[CH];
    float a[CH], b[CH], c[CH];
    unsigned int ia[CH];
    for (int k = 0; k < CH; ++k) {
        a[k] = seed + k + threadIdx.x * 1e-3f; b[k] = 0.999f + k * 1e-6f; c[k] = 1e-3f + k * 1e-6f; ia[k] = k + threadIdx.x;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pa[k]) : "f"(a[k]), "f"(a[k] + 1.f));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pb[k]) : "f"(b[k]));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pc[k]) : "f"(c[k]));
    }
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int k1 = (k + 1) % CH, k2 = (k + 2) % CH;
                if (KIND == P_SHARED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[0]), "l"(pc[0]));
                if (KIND == P_DISTINCT) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[k]), "l"(pc[k1]));
                if (KIND == S_SHARED) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[0]), "f"(c[0]));
                if (KIND == S_DISTINCT) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c[k1]));
                if (KIND == MIX_1P_1S) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[k]), "l"(pc[k1]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c[k2]));
                }
                if (KIND == MIX_2P_1S) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[k]), "l"(pc[k1]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c[k2]));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pc[k]) : "l"(pb[k1]), "l"(pa[k1]));
                }
                if (KIND == MIX_1P_2S) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[k]), "l"(pc[k1]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b[k]), "f"(c[k2]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(c[k]) : "f"(b[k1]), "f"(a[k1]));
                }
                if (KIND == PMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(pa[k]) : "l"(pb[k]));
                if (KIND == PADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[k]) : "l"(pc[k]));
                if (KIND == MIX_1P_1ALU) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[k]), "l"(pc[k1]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[k]) : "r"(ia[k1]), "r"(ia[k2]));
                }
                if (KIND == MIX_2P_1MUFU) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[k]) : "l"(pb[k]), "l"(pc[k1]));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pc[k]) : "l"(pb[k1]), "l"(pa[k1]));
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    float acc = 0.f;
    for (int k = 0; k < CH; ++k) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pa[k]));
        acc += a[k] + b[k] + c[k] + lo + hi + (float)ia[k];
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pc[k]));
        acc += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(int sms, int ctas_per_sm, float* out, unsigned long long* cyc) {
    int grid = sms * ctas_per_sm;
    bench<KIND><<<grid, 256>>>(out, cyc, 1.0f);
    CHECK(cudaDeviceSynchronize());
    bench<KIND><<<grid, 256>>>(out, cyc, 1.0f);
    CHECK(cudaDeviceSynchronize());
    static unsigned long long h[8192];
    CHECK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    double steps = (double)ITERS * 4 * CH;
    double per_sm = steps * 256.0 * ctas_per_sm / mean;             // chain steps per clock per SM
    printf("%-28s ctas/SM=%d  lane-ops/clk/SM: all %.1f  fp32 %.1f   (warp-instr/clk/SM %.2f)\n", names[KIND], ctas_per_sm,
           per_sm * ops[KIND], per_sm * fp32ops[KIND],
           per_sm / 32.0 * (KIND == MIX_1P_1S || KIND == MIX_1P_1ALU ? 2 : (KIND == MIX_2P_1S || KIND == MIX_1P_2S || KIND == MIX_2P_1MUFU ? 3 : 1)));
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs=%d\n", p.name, p.multiProcessorCount);
    int sms = p.multiProcessorCount;
    float* out; unsigned long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(float) * 256 * sms * 8));
    CHECK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms * 8));
    for (int occ : {2, 4}) {
        run<P_SHARED>(sms, occ, out, cyc); run<P_DISTINCT>(sms, occ, out, cyc); run<S_SHARED>(sms, occ, out, cyc);
        run<S_DISTINCT>(sms, occ, out, cyc); run<MIX_1P_1S>(sms, occ, out, cyc); run<MIX_2P_1S>(sms, occ, out, cyc);
        run<MIX_1P_2S>(sms, occ, out, cyc); run<PMUL2>(sms, occ, out, cyc); run<PADD2>(sms, occ, out, cyc);
        run<MIX_1P_1ALU>(sms, occ, out, cyc); run<MIX_2P_1MUFU>(sms, occ, out, cyc);
    }
    return 0;
}
