// fp64_stream.cu -- sustained issue rate of pure FP64 instruction streams on sm_100a (cycles per
// warp instruction per SMSP), for different operand patterns.  Every kernel runs NCH independent
// chains, UNROLL x NCH instructions per loop trip, with inline PTX so that the counts are exact.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_stream.bin tools/fp64_stream.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int NCH, int UNROLL>
__global__ void __launch_bounds__(1024, 1) stream(double *out, long long *cyc, const double *in, int iters)
{
    double acc[NCH], x[NCH], y[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        acc[i] = in[i] + threadIdx.x;
        x[i] = in[32 + i];
        y[i] = in[64 + i];
    }
    const double a = in[100], b = in[101];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < UNROLL; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
                if (OP == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(x[i]), "d"(y[i]));
                if (OP == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[i]) : "d"(x[i]), "d"(y[i]));
                if (OP == 3) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(x[i]));
                if (OP == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(y[i]));
                if (OP == 5) asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+d"(acc[i]) : "d"(x[i]));
                if (OP == 6) asm volatile("fma.rn.f64 %0, %0, %0, %1;" : "+d"(acc[i]) : "d"(y[i]));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int NCH, int UNROLL>
double run(int threads, double *out, long long *cyc, const double *in)
{
    const int iters = 4096 / UNROLL;
    for (int rep = 0; rep < 2; ++rep) {
        stream<OP, NCH, UNROLL><<<148, threads>>>(out, cyc, in, iters);
        cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += (double)h[i];
    return m / 148 / iters / (UNROLL * NCH) / (threads / 128);
}

#define ROW(OP, NAME)                                                                                 \
    printf("%-40s 8 chains x8: %.3f  x32: %.3f | 16 chains x4: %.3f  x16: %.3f   (warps/SMSP=%d)\n", NAME, \
           run<OP, 8, 8>(threads, out, cyc, in), run<OP, 8, 32>(threads, out, cyc, in),               \
           run<OP, 16, 4>(threads, out, cyc, in), run<OP, 16, 16>(threads, out, cyc, in), threads / 128);

int main()
{
    double *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&in, 128 * sizeof(double));
    double h[128];
    for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-9 * i;
    h[100] = 0.999999;
    h[101] = 1e-9;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("SMSP cycles per warp instruction, pure FP64 streams\n");
    for (int threads = 256; threads <= 1024; threads *= 2) {
        ROW(0, "DFMA acc = acc*a + b (a, b invariant)")
        ROW(1, "DFMA acc = acc*x_i + y_i")
        ROW(2, "DFMA acc = x_i*y_i + acc")
        ROW(5, "DFMA acc = x_i*x_i + acc")
        ROW(6, "DFMA acc = acc*acc + y_i")
        ROW(3, "DMUL acc = acc*x_i")
        ROW(4, "DADD acc = acc + y_i")
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
