// Probe: one radix-8 pass (8 LDS.128, DFT-8, 7 twiddle LDS.128 + complex multiplies, 8 STS.128) in a barrier-free loop on
// warp-private shared memory.  mode 0: full pass; 1: loads / stores only (no arithmetic); 2: arithmetic only (data stays in
// registers); 3: full pass with the twiddles held in registers (no twiddle loads).
// Question: does a warp mix that is half FP64 and half LSU overlap the two pipes when nothing synchronises the warps?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../transport_analysis_b200/csrc/ta_common.cuh"
#include "../../transport_analysis_b200/csrc/dft_regs.cuh"
using namespace ta;

template <int MODE>
__global__ void k(double* out, long long* clk, int iters) {
    extern __shared__ __align__(16) unsigned char raw[];
    cd* sm = reinterpret_cast<cd*>(raw);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    cd* buf = sm + w * 288;                 // 8 x 36 elements per warp (lane + 36 q: conflict free)
    cd* tw = sm + nw * 288 + lane;          // 7 x 32 twiddles, shared by all warps
    for (int i = threadIdx.x; i < nw * 288 + 7 * 32; i += blockDim.x) sm[i] = cmake<double>(1.0 + 1e-9 * i, 1e-9 * i);
    __syncthreads();
    cd x[8], t[7];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = buf[lane + 36 * q];
#pragma unroll
    for (int q = 0; q < 7; ++q) t[q] = tw[32 * q];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE != 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = buf[lane + 36 * q];
        }
        if (MODE != 1) {
            Dft<8, -1>::run(x);
            if (MODE == 0) {
#pragma unroll
                for (int q = 1; q < 8; ++q) x[q] = cmul(x[q], tw[32 * (q - 1)]);
            } else {
#pragma unroll
                for (int q = 1; q < 8; ++q) x[q] = cmul(x[q], t[q - 1]);
            }
        }
        if (MODE != 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) buf[lane + 36 * q] = x[q];
            __syncwarp();
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q].x + x[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0 && blockIdx.x == 0) clk[w] = t1 - t0;
}

template <int MODE>
void run(const char* name, int nw, double* d, long long* dc) {
    const int iters = 4000;
    const size_t smem = (size_t)(nw * 288 + 7 * 32) * 16;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<148, 32 * nw, smem>>>(d, dc, iters);
    cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, dc, nw * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-28s warps %2d: %8.1f clk per pass-iteration of the SM (%6.1f clk per warp-butterfly)\n", name, nw, (double)mx / iters,
           (double)mx / iters / nw);
}

int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 32 * 8);
    for (int nw : {4, 8, 12, 20, 32}) {
        run<0>("full (twiddles from smem)", nw, d, dc);
        run<3>("full (twiddles in registers)", nw, d, dc);
        run<1>("loads/stores only", nw, d, dc);
        run<2>("arithmetic only", nw, d, dc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
