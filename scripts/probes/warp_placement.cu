// Probe: which hardware warp slots (%warpid; SM sub-partition = %warpid % 4) do the warps of
// co-resident CTAs get?  Usage: warp_placement <threads per CTA> <smem bytes per CTA>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void probe(int* out, int wpc) {
    extern __shared__ char sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) {}   // stay resident so that CTAs overlap
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        int* o = out + ((size_t)blockIdx.x * wpc + w) * 2;
        o[0] = smid; o[1] = warpid;
    }
}
int main(int argc, char** argv) {
    int nthr = argc > 1 ? atoi(argv[1]) : 160, smem = argc > 2 ? atoi(argv[2]) : 95000;
    int wpc = nthr / 32, grid = 296;
    int* d; cudaMalloc(&d, grid * wpc * 2 * sizeof(int));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<grid, nthr, smem>>>(d, wpc);
    int* h = (int*)malloc(grid * wpc * 2 * sizeof(int));
    cudaMemcpy(h, d, grid * wpc * 2 * sizeof(int), cudaMemcpyDeviceToHost);
    printf("threads/CTA=%d smem=%d err=%s\n", nthr, smem, cudaGetErrorString(cudaGetLastError()));
    for (int sm = 0; sm < 3; ++sm) {
        printf("SM %d:", sm);
        for (int b = 0; b < grid; ++b)
            if (h[(b * wpc) * 2] == sm) { printf("  cta %d warpids:", b); for (int w = 0; w < wpc; ++w) printf(" %d", h[(b * wpc + w) * 2 + 1]); }
        printf("\n");
    }
    return 0;
}
