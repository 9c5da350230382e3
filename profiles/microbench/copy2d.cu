// D2H of one narrow column into a strided host table (the drop-in's PedState records, 132-byte stride): 2-D DMA copy vs
// a packed copy + host scatter loop (1 and 4 threads).  Build: nvcc -O3 -o copy2d copy2d.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t n = 65536, stride = 132;
    char* table = (char*)malloc(n * stride + 64);
    memset(table, 1, n * stride + 64);
    CHECK(cudaHostRegister(table, n * stride, cudaHostRegisterPortable));
    double4* dev; CHECK(cudaMalloc(&dev, n * sizeof(double4)));
    CHECK(cudaMemset(dev, 0, n * sizeof(double4)));
    char* devtab; CHECK(cudaMalloc(&devtab, n * stride));
    double4* pinned; CHECK(cudaMallocHost(&pinned, n * sizeof(double4)));
    cudaStream_t st; CHECK(cudaStreamCreate(&st));
    for (int rep = 0; rep < 3; ++rep) {
        double t0 = now();
        for (int k = 0; k < 20; ++k) {
            CHECK(cudaMemcpy2DAsync(table + 60, stride, dev, 32, 24, n, cudaMemcpyDeviceToHost, st));
            CHECK(cudaMemcpy2DAsync(table + 124, stride, (char*)dev + 24, 32, 8, n, cudaMemcpyDeviceToHost, st));
            CHECK(cudaStreamSynchronize(st));
        }
        printf("2-D DMA (24 B + 8 B columns, dpitch 132): %.3f ms\n", (now() - t0) / 20 * 1e3);
        t0 = now();
        for (int k = 0; k < 20; ++k) {
            CHECK(cudaMemcpyAsync(table, devtab, n * stride, cudaMemcpyDeviceToHost, st));
            CHECK(cudaStreamSynchronize(st));
        }
        printf("whole table D2H (8.65 MB): %.3f ms\n", (now() - t0) / 20 * 1e3);
        t0 = now();
        for (int k = 0; k < 20; ++k) {
            CHECK(cudaMemcpyAsync(devtab, table, n * stride, cudaMemcpyHostToDevice, st));
            CHECK(cudaStreamSynchronize(st));
        }
        printf("whole table H2D (8.65 MB): %.3f ms\n", (now() - t0) / 20 * 1e3);
        for (int threads : {1, 2, 4, 8}) {
            t0 = now();
            for (int k = 0; k < 20; ++k) {
                CHECK(cudaMemcpyAsync(pinned, dev, n * sizeof(double4), cudaMemcpyDeviceToHost, st));
                CHECK(cudaStreamSynchronize(st));
                auto body = [&](size_t a, size_t b) {
                    for (size_t i = a; i < b; ++i) { memcpy(table + i * stride + 60, &pinned[i], 24); memcpy(table + i * stride + 124, &pinned[i].w, 8); }
                };
                std::vector<std::thread> th;
                for (int t = 1; t < threads; ++t) th.emplace_back(body, n * t / threads, n * (t + 1) / threads);
                body(0, n / threads);
                for (auto& x : th) x.join();
            }
            printf("packed D2H (2 MB) + host scatter, %d thread(s) spawned per call: %.3f ms\n", threads, (now() - t0) / 20 * 1e3);
        }
    }
    return 0;
}
