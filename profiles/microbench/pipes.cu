// Pipe-throughput microbenchmark for sm_100a: measures issue rates that bound the
// all-pairs pedestrian-force kernel (FP32 FMA pipe, packed f32x2, ALU pipe, MUFU, FP64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;     // independent chains per thread

enum Kind { K_FFMA, K_FFMA2, K_FMUL, K_FADD, K_FMNMX, K_FFMA_FMNMX, K_EX2, K_RSQ, K_RCP, K_DFMA, K_DADD, K_DMUL,
            K_FFMA_EX2_4to1, K_FFMA2_FMNMX, K_FSEL, K_LDS128, K_FFMA_LDS, K_COUNT };
static const char* names[] = {"FFMA", "FFMA2(f32x2)", "FMUL", "FADD", "FMNMX", "FFMA+FMNMX 1:1", "EX2+FFMA 1:1", "RSQ+FFMA 1:1",
  "RCP+FFMA 1:1", "DFMA", "DADD", "DMUL", "FFMA+EX2 4:1", "FFMA2+FMNMX 1:1", "FSETP+FSEL", "LDS.128 bcast", "FFMA+LDS128 8:1"};
// lane-level "ops" per inner step per chain (for reporting instructions, not flops)
static const double instr_per_step[] = {1, 1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 1, 5, 2, 2, 1, 9};

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, unsigned long long* cycles, float seed) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
    float a[CH], b[CH];
    double da[CH];
    unsigned long long pa[CH];
    for (int c = 0; c < CH; ++c) {
        a[c] = seed + c + threadIdx.x * 1e-3f; b[c] = seed * 0.5f + c; da[c] = a[c];
        float lo = a[c], hi = b[c];
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pa[c]) : "f"(lo), "f"(hi));
    }
    float m1 = 0.999f + seed * 1e-6f, m2 = 1e-3f + seed * 1e-6f;
    double dm1 = m1, dm2 = m2;
    unsigned long long pm1, pm2;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pm1) : "f"(m1));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pm2) : "f"(m2));
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                if (KIND == K_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[c]) : "l"(pm1), "l"(pm2));
                if (KIND == K_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m1));
                if (KIND == K_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m2));
                if (KIND == K_FMNMX) { if (u & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c])); else asm volatile("min.f32 %1, %0, %1;" : "+f"(a[c]), "+f"(b[c])); }
                if (KIND == K_FFMA_FMNMX) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                    if (u & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(b[c]) : "f"(a[c])); else asm volatile("min.f32 %0, %0, %1;" : "+f"(b[c]) : "f"(m2));
                }
                if (KIND == K_EX2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[c])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m2), "f"(m1)); }
                if (KIND == K_RSQ) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[c])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m2), "f"(m1)); }
                if (KIND == K_RCP) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[c])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m2), "f"(m1)); }
                if (KIND == K_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(da[c]) : "d"(dm1), "d"(dm2));
                if (KIND == K_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(da[c]) : "d"(dm2));
                if (KIND == K_DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(da[c]) : "d"(dm1));
                if (KIND == K_FFMA_EX2_4to1) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b[c]) : "f"(a[c]));
                }
                if (KIND == K_FFMA2_FMNMX) {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[c]) : "l"(pm1), "l"(pm2));
                    if (u & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(b[c]) : "f"(a[c])); else asm volatile("min.f32 %0, %0, %1;" : "+f"(b[c]) : "f"(m2));
                }
                if (KIND == K_FSEL) {
                    asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p; }" : "+f"(a[c]) : "f"(b[c]), "f"(m2));
                }
                if (KIND == K_LDS128) {
                    float4 v; unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(it + c) & 63]);
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                    b[c] = v.x + v.w;  // keeps the load live
                }
                if (KIND == K_FFMA_LDS) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m1), "f"(m2));
                    float4 v; unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(it + c) & 63]);
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                    b[c] = v.x;
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    float acc = 0.f;
    for (int c = 0; c < CH; ++c) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pa[c]));
        acc += a[c] + b[c] + (float)da[c] + lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(int sms, int ctas_per_sm, float* out, unsigned long long* cyc, double mhz) {
    int grid = sms * ctas_per_sm;
    bench<KIND><<<grid, 256>>>(out, cyc, 1.0f);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KIND><<<grid, 256>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[4096];
    CHECK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    double steps = (double)ITERS * 4 * CH;                       // per thread
    double warp_instr = steps * instr_per_step[KIND] * 8 * ctas_per_sm;  // per SM (8 warps per CTA)
    // per-SM warp-instructions per cycle (clock64-based, CTA-resident time)
    double ipc_sm = warp_instr / mean;
    double total_lane = steps * instr_per_step[KIND] * 256.0 * grid;
    printf("%-18s ctas/SM=%d  cyc=%.0f  warp-instr/clk/SM=%.3f (lanes/clk/SM=%.1f)  event: %.3f ms -> %.2f Tlane-instr/s (eff clk %.0f MHz)\n",
           names[KIND], ctas_per_sm, mean, ipc_sm, ipc_sm * 32, ms, total_lane / (ms * 1e-3) / 1e12,
           total_lane / (ms * 1e-3) / (ipc_sm * 32 * sms) / 1e6);
    (void)mhz;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs=%d  clock=%d kHz  cc=%d.%d\n", p.name, p.multiProcessorCount, p.clockRate, p.major, p.minor);
    int sms = p.multiProcessorCount;
    float* out; unsigned long long* cyc;
    CHECK(cudaMalloc(&out, sizeof(float) * 256 * sms * 8));
    CHECK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms * 8));
    for (int occ : {2, 4}) {
        run<K_FFMA>(sms, occ, out, cyc, 0); run<K_FFMA2>(sms, occ, out, cyc, 0); run<K_FMUL>(sms, occ, out, cyc, 0);
        run<K_FADD>(sms, occ, out, cyc, 0); run<K_FMNMX>(sms, occ, out, cyc, 0); run<K_FFMA_FMNMX>(sms, occ, out, cyc, 0);
        run<K_FFMA2_FMNMX>(sms, occ, out, cyc, 0); run<K_FSEL>(sms, occ, out, cyc, 0);
        run<K_EX2>(sms, occ, out, cyc, 0); run<K_RSQ>(sms, occ, out, cyc, 0); run<K_RCP>(sms, occ, out, cyc, 0);
        run<K_FFMA_EX2_4to1>(sms, occ, out, cyc, 0);
        run<K_DFMA>(sms, occ, out, cyc, 0); run<K_DADD>(sms, occ, out, cyc, 0); run<K_DMUL>(sms, occ, out, cyc, 0);
        run<K_LDS128>(sms, occ, out, cyc, 0); run<K_FFMA_LDS>(sms, occ, out, cyc, 0);
    }
    return 0;
}
