// issue_slots.cu -- how many issue cycles does a non-FP64 instruction cost next to a saturated
// FP64 pipe on sm_100a?  Each warp runs a loop of 64 independent-chain DFMAs plus K extra
// instructions of one kind; the slope of cycles/iteration over K is the cost of that kind.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_slots.bin tools/issue_slots.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum { K_NONE, K_MUFU, K_LDS128, K_LDS64, K_STS128, K_IADD, K_MUFU32, K_DADD, K_DMUL };

template <int KIND, int K>
__global__ void __launch_bounds__(1024, 1) probe(double *out, long long *cyc, int iters, double a, double b)
{
    __shared__ double2 sh[2048];
    const int tid = threadIdx.x;
    sh[tid] = make_double2(b, b);
    sh[tid + 1024] = make_double2(b, b);
    __syncthreads();
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = a + i + tid;
    double feed[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) feed[i] = b * (i + 1);
    int ia = tid;
    float fa = 1.0f + tid;
    unsigned sa = (unsigned)__cvta_generic_to_shared(&sh[tid]);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, feed[i]);
            // K extras spread over the 8 rounds, every one with its own operands
#pragma unroll
            for (int e = 0; e < (K + 7 - r) / 8; ++e) {
                const int u = (r * 4 + e) & 31;
                if (KIND == K_MUFU) {
                    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(feed[(r + e) & 7]) : "d"(acc[(r + 3 * e + 1) & 7]));
                } else if (KIND == K_MUFU32) {
                    asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(fa));
                } else if (KIND == K_LDS128) {
                    double y;
                    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(feed[(r + e) & 7]), "=d"(y) : "r"(sa + 512 * u));
                } else if (KIND == K_LDS64) {
                    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(feed[(r + e) & 7]) : "r"(sa + 512 * u));
                } else if (KIND == K_STS128) {
                    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(sa + 512 * u), "d"(acc[(r + e) & 7]), "d"(acc[(r + e + 1) & 7]) : "memory");
                } else if (KIND == K_IADD) {
                    asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(ia) : "r"(it));
                } else if (KIND == K_DADD) {
                    asm volatile("add.f64 %0, %0, %1;" : "+d"(acc[(r + e) & 7]) : "d"(feed[e & 7]));
                } else if (KIND == K_DMUL) {
                    asm volatile("mul.f64 %0, %0, %1;" : "+d"(acc[(r + e) & 7]) : "d"(a));
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + tid] = s + ia + feed[0] + feed[3] + fa;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int K>
double run(int threads, double *out, long long *cyc, int iters)
{
    probe<KIND, K><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    probe<KIND, K><<<148, threads>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += (double)h[i];
    return m / 148 / iters;
}

#define ROW(KIND, NAME)                                                                           \
    {                                                                                             \
        double c0 = run<K_NONE, 0>(threads, out, cyc, iters), c8 = run<KIND, 8>(threads, out, cyc, iters), \
               c16 = run<KIND, 16>(threads, out, cyc, iters), c32 = run<KIND, 32>(threads, out, cyc, iters); \
        printf("%-8s warps/SMSP=%d  cycles/iter: K=0 %.1f  K=8 %.1f  K=16 %.1f  K=32 %.1f   per extra (per SMSP issue "  \
               "cycles): %.2f %.2f %.2f\n", NAME, wps, c0, c8, c16, c32,                           \
               (c8 - c0) / 8 / wps, (c16 - c0) / 16 / wps, (c32 - c0) / 32 / wps);                 \
    }

int main()
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int threads = wps * 128;
        double c0 = run<K_NONE, 0>(threads, out, cyc, iters);
        printf("baseline warps/SMSP=%d: %.1f cycles per 64 DFMA per warp -> %.3f cycles per DFMA per SMSP\n", wps,
               c0, c0 / 64 / wps);
        ROW(K_MUFU, "MUFU64H")
        ROW(K_MUFU32, "MUFU32")
        ROW(K_LDS128, "LDS.128")
        ROW(K_LDS64, "LDS.64")
        ROW(K_STS128, "STS.128")
        ROW(K_IADD, "IMAD")
        ROW(K_DADD, "DADD")
        ROW(K_DMUL, "DMUL")
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
