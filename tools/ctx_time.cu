// How long does a process wait for its CUDA context?  (tools/make_image_demo.sh: the floor under a one-shot `make image`.)
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    int n = 0;
    cudaGetDeviceCount(&n);
    auto t1 = clk::now();
    cudaSetDevice(0);
    cudaFree(0);
    auto t2 = clk::now();
    void *p = nullptr;
    cudaMalloc(&p, 1 << 20);
    cudaStream_t s;
    cudaStreamCreate(&s);
    auto t3 = clk::now();
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    printf("devices %d: driver init %.1f ms, context %.1f ms, first malloc + stream %.1f ms\n", n, ms(t0, t1), ms(t1, t2), ms(t2, t3));
    return 0;
}
