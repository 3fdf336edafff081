// FP64 latency / issue microbenchmark for sm_100a: dependent chains of DADD, DMUL, division, sqrt with 1..8 independent
// chains per thread and 1..16 warps per SM.  nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double* out, const double* in, int iters, long long* cyc)
{
    double x[ILP];
    double a = in[0], b = in[1];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = in[2 + i] + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) x[i] = x[i] + a;
                if (OP == 1) x[i] = x[i] * b;
                if (OP == 2) x[i] = a / x[i];
                if (OP == 3) x[i] = sqrt(x[i] + a);
                if (OP == 4) x[i] = x[i] * b + a;  // -fmad=false: DMUL then DADD
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
void run(const char* name, int warps)
{
    double *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 128); cudaMalloc(&cyc, 8);
    double h[16]; for (int i = 0; i < 16; i++) h[i] = 1.0 + 1e-3 * i; h[1] = 1.0000001;
    cudaMemcpy(in, h, 128, cudaMemcpyHostToDevice);
    const int iters = 256;
    k<OP, ILP><<<148, 32 * warps>>>(out, in, iters, cyc);
    k<OP, ILP><<<148, 32 * warps>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)iters * 8 * ILP;
    printf("%-10s ilp %d warps/SM %2d : %7.2f cycles per op per thread-chain step (%.2f cycles per warp-instruction slot per SM: %.3f)\n", name, ILP, warps,
           (double)c / (iters * 8), (double)c / ops, (double)c / (ops * warps));
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}
int main()
{
    run<0, 1>("DADD", 1); run<0, 2>("DADD", 1); run<0, 4>("DADD", 1); run<0, 8>("DADD", 1);
    run<0, 1>("DADD", 4); run<0, 1>("DADD", 8); run<0, 1>("DADD", 16); run<0, 4>("DADD", 16); run<0, 4>("DADD", 32);
    run<1, 1>("DMUL", 1); run<1, 4>("DMUL", 1);
    run<4, 1>("DMUL+DADD", 1); run<4, 4>("DMUL+DADD", 1);
    run<2, 1>("DIV", 1); run<2, 2>("DIV", 1); run<2, 4>("DIV", 1); run<2, 1>("DIV", 8); run<2, 1>("DIV", 16);
    run<3, 1>("SQRT", 1); run<3, 4>("SQRT", 1); run<3, 1>("SQRT", 16);
    return 0;
}
