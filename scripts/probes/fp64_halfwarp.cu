// Probe: does the FP64 pipe (16 lanes per SM sub-partition, 2 clk per full warp instruction) take 1 clk for a warp
// instruction whose upper (or lower) 16 lanes are inactive?  mode 0: all 32 lanes; 1: lanes 0-15; 2: lanes 16-31;
// 3: even lanes only (16 active lanes spread over both halves).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, int iters, double a, double b, int mode) {
    const int lane = threadIdx.x & 31;
    const bool on = mode == 0 || (mode == 1 && lane < 16) || (mode == 2 && lane >= 16) || (mode == 3 && (lane & 1) == 0);
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i + lane;
    __syncthreads();
    const long long t0 = clock64();
    if (on) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 8; ++rep)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
    }
    __syncwarp();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0 && blockIdx.x == 0) clk[threadIdx.x >> 5] = t1 - t0;
}
int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 32 * 8);
    const int iters = 4000;
    for (int nw : {4, 8}) for (int mode = 0; mode < 4; ++mode) {
        k<<<148, 32 * nw>>>(d, dc, iters, 1.0000001, 1e-9, mode);
        cudaDeviceSynchronize();
        long long h[32]; cudaMemcpy(h, dc, nw * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("warps %d mode %d: %.2f clk per DFMA warp-instruction per sub-partition\n", nw, mode, mx / (iters * 64.0 * (nw / 4)));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
