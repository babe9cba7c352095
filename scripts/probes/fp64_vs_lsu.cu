// Probe: do FP64 math and shared-memory traffic overlap on one SM?  Warps [0, nfp) run DFMA chains
// (ILP 8), warps [nfp, nfp + nls) stream LDS.128 / STS.128 on a conflict-free layout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, int nfp, int nls, int iters, double a, double b, int mode) {
    extern __shared__ double2 sm[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    long long t0 = clock64();
    double s = 0;
    if (w < nfp) {
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = a + i + lane;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 8; ++rep)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[i];
    } else if (w < nfp + nls) {
        double2* p = sm + (w - nfp) * 512 + lane;
        double2 v[8];
        for (int it = 0; it < iters; ++it) {
            // 32 LDS.128 + 32 STS.128 per iteration = same instruction count as the 64 DFMAs
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    unsigned addr = (unsigned)__cvta_generic_to_shared(p + 32 * i + 256 * (rep & 1));
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[i].x), "=d"(v[i].y) : "r"(addr));
                }
                if (mode == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        unsigned addr = (unsigned)__cvta_generic_to_shared(p + 32 * i + 256 * ((rep + 1) & 1));
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(v[i].x), "d"(v[i].y) : "memory");
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;   // loads only (+16 DADD)
                }
            }
        }
        s += v[0].x;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0 && blockIdx.x == 0) clk[w] = t1 - t0;
}
int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 32 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
    const int iters = 2000;
    int cfg[][2] = {{4, 0}, {0, 4}, {4, 4}, {8, 0}, {0, 8}, {8, 8}, {4, 8}, {8, 4}, {0, 2}, {4, 2}, {0, 1}, {4, 1}};
    for (auto& c : cfg) {
        int nfp = c[0], nls = c[1], nw = nfp + nls;
        k<<<148, 32 * nw, 8192 * 16>>>(d, dc, nfp, nls, iters, 1.0000001, 1e-9, 1);
        cudaDeviceSynchronize();
        long long h[32]; cudaMemcpy(h, dc, nw * 8, cudaMemcpyDeviceToHost);
        long long fmax = 0, lmax = 0;
        for (int w = 0; w < nfp; ++w) fmax = h[w] > fmax ? h[w] : fmax;
        for (int w = nfp; w < nw; ++w) lmax = h[w] > lmax ? h[w] : lmax;
        printf("fp warps %d, lds/sts warps %d: ", nfp, nls);
        if (nfp) printf("FP64 %.2f clk/DFMA/SMSP  ", fmax / (iters * 64.0 * ((nfp + 3) / 4)));
        if (nls) printf("smem %.2f clk/wavefront/SM (%d warps x %d x 4 wavefronts)", lmax / (iters * 64.0 * 4 * nls), nls, iters * 64);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
