// Library comparator for K1 (a probe, not a product path): what the same autocorrelation costs when the transforms are
// cuFFT calls.  Per particle: D = 3 zero-padded real-to-complex transforms of length L (cuFFT needs the padded input in
// HBM), a power-spectrum pass, one complex-to-real inverse.  L = 20,480 (the length K1 uses for T = 10,000) and 32,768
// (tidynamics' own power-of-two choice).  Times only the cuFFT executions; the padding, |X|^2 and normalisation kernels a
// real cuFFT pipeline also needs are NOT included, so this is a lower bound for the library route.
#include <cstdio>
#include <cuda_runtime.h>
#include <cufft.h>
#define CK(x) do { auto e_ = (x); if (e_ != 0) { printf("error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); return 1; } } while (0)
int main() {
    const int atoms = 20000, D = 3;
    for (int L : {20480, 32768}) {
        const long long nc = L / 2 + 1;
        double* in; cufftDoubleComplex* spec; double* out;
        CK(cudaMalloc(&in, sizeof(double) * (size_t)L * atoms * D));
        CK(cudaMalloc(&spec, sizeof(cufftDoubleComplex) * (size_t)nc * atoms * D));
        CK(cudaMalloc(&out, sizeof(double) * (size_t)L * atoms));
        CK(cudaMemset(in, 0, sizeof(double) * (size_t)L * atoms * D));
        cufftHandle fwd, inv;
        int n[1] = {L};
        CK(cufftPlanMany(&fwd, 1, n, nullptr, 1, L, nullptr, 1, (int)nc, CUFFT_D2Z, atoms * D));
        CK(cufftPlanMany(&inv, 1, n, nullptr, 1, (int)nc, nullptr, 1, L, CUFFT_Z2D, atoms));
        cudaEvent_t a, b, c; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(a);
            CK(cufftExecD2Z(fwd, in, spec));
            cudaEventRecord(b);
            CK(cufftExecZ2D(inv, spec, out));
            cudaEventRecord(c);
            CK(cudaDeviceSynchronize());
            float f, i; cudaEventElapsedTime(&f, a, b); cudaEventElapsedTime(&i, b, c);
            if (it == 2)
                printf("L = %d, %d particles x %d dims: D2Z %.3f ms + Z2D %.3f ms = %.3f ms  ->  %.1f ms per 100,000 particles (cuFFT executions only)\n",
                       L, atoms, D, f, i, f + i, (f + i) * 100000.0 / atoms);
        }
        cufftDestroy(fwd); cufftDestroy(inv); cudaFree(in); cudaFree(spec); cudaFree(out);
    }
    return 0;
}
