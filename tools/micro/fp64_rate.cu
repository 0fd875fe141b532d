// Micro-benchmark: throughput of plain (non-tensor) FP64 instructions per SM on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void chain(T* out, int iters, T a, T b) {
    T x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = (T)threadIdx.x + (T)i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + b;
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T, int ILP> void run(const char* name, int warps) {
    T* out; cudaMalloc(&out, 148 * 1024 * sizeof(T));
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<T, ILP><<<148, warps * 32>>>(out, 16, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e0);
    chain<T, ILP><<<148, warps * 32>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)iters * ILP * warps;          // warp-instructions per SM
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%s ILP=%d warps/SM=%2d: %.2f cycles per warp-FMA per SM (%.1f lane-FMA/clk/SM)\n", name, ILP, warps, cycles / instr, 32.0 * instr / cycles);
    cudaFree(out);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) run<double, 4>("fp64", w);
    for (int w : {4, 16}) run<float, 4>("fp32", w);
    return 0;
}
