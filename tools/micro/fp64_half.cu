// Does an FP64 warp instruction with only 16 (or 8) active lanes cost less issue time than one with 32?  ILP 8 per thread
// (pipe-bound), 4 and 8 warps per SM (1 and 2 per scheduler), lanes >= ACTIVE idle in a branch.
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o fp64_half fp64_half.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, const double* in, int iters, long long* cyc, int active)
{
    double x[ILP];
    double a = in[0];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = in[2 + i] + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    if ((threadIdx.x & 31) < active) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int i = 0; i < ILP; i++) x[i] = x[i] + a;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 256); cudaMalloc(&cyc, 8);
    double h[32]; for (int i = 0; i < 32; i++) h[i] = 1.0 + 1e-3 * i;
    cudaMemcpy(in, h, 256, cudaMemcpyHostToDevice);
    const int iters = 512;
    for (int warps : {4, 8, 16})
        for (int active : {32, 16, 8, 1}) {
            k<8><<<148, 32 * warps>>>(out, in, iters, cyc, active);
            k<8><<<148, 32 * warps>>>(out, in, iters, cyc, active);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d active lanes %2d : %.3f cycles per warp-instruction per scheduler\n", warps, active, (double)c / (iters * 64.0) / (warps / 4.0));
        }
    return 0;
}
