// Probe: FP64 pipe of one SM sub-partition -- cycles per warp instruction for DFMA / DADD / mixed
// streams as a function of warps per sub-partition and independent chains per warp (ILP).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int ILP, int MODE>
__global__ void k(double* out, long long* clk, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                else if (MODE == 1) x[i] = x[i] + b;
                else if (MODE == 2) x[i] = (rep & 1) ? fma(x[i], a, b) : x[i] + b;
                else x[i] = x[i] * a;
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int ILP, int MODE>
void run(int warps, double* d, long long* dc) {
    int iters = 2000;
    k<ILP, MODE><<<148, 32 * warps>>>(d, dc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double inst_per_warp = (double)iters * 8 * ILP;
    double wps = warps / 4.0;   // warps per sub-partition
    printf("mode %d ILP %d warps/SM %2d: %.2f clk per warp-instr per warp, %.2f clk per instr per sub-partition\n", MODE, ILP, warps,
           c / inst_per_warp, c / (inst_per_warp * (wps < 1 ? 1 : wps)));
}
int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int w : {1, 4, 8, 12, 16}) {
        run<1, 0>(w, d, dc); run<2, 0>(w, d, dc); run<4, 0>(w, d, dc); run<8, 0>(w, d, dc); run<16, 0>(w, d, dc);
    }
    for (int w : {4, 8, 12}) { run<4, 1>(w, d, dc); run<8, 1>(w, d, dc); run<16, 1>(w, d, dc); }
    for (int w : {4, 8, 12}) { run<4, 2>(w, d, dc); run<8, 2>(w, d, dc); run<16, 2>(w, d, dc); }
    for (int w : {4, 8, 12}) { run<8, 3>(w, d, dc); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
