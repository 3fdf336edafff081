// Latency of ONE thread's manifold solve (the chain a round of the batch-wide step waits for): cycles per contact-iteration
// of the literal solve (resolve_t) and of the straight-line solve (resolve_s), for 1 / 4 / 8 contacts, with 1, 4 and 8
// warps per SM.  A box resting on a static ground, friction on, 10 iterations.
//   nvcc -O3 -std=c++17 -fmad=false -gencode arch=compute_100a,code=sm_100a -o solve_lat solve_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../physicssimulator_b200/csrc/ps_physics.cuh"
using namespace psp;

__device__ void make_case(int t, int nc, Dyn& b1, Par& p1, Dyn& b2, Par& p2, Contact* cs, bool ground)
{
    const double j = 1e-3 * (t % 32);
    p1.inv_mass = ground ? 0.0 : 0.01; p1.iinv[0] = p1.iinv[1] = p1.iinv[2] = ground ? 0.0 : 0.06;
    p1.F = mk(0, 0, 0); p1.T = mk(0, 0, 0); p1.e = 0.3; p1.mu = 0.5; p1.dims[0] = 40; p1.dims[1] = 40; p1.dims[2] = 1; p1.shape = 1; p1.is_static = ground;
    p2.inv_mass = 0.008; p2.iinv[0] = 0.05; p2.iinv[1] = 0.07; p2.iinv[2] = 0.06;
    p2.F = mk(0, 0, -9.81 / 0.008); p2.T = mk(0, 0, 0); p2.e = 0.3; p2.mu = 0.5; p2.dims[0] = 0.8; p2.dims[1] = 0.6; p2.dims[2] = 0.5; p2.shape = 1; p2.is_static = false;
    b1.x = mk(0, 0, -0.5); b1.v = mk(0, 0, 0); b1.L = mk(0, 0, 0);
    for (int k = 0; k < 9; k++) { b1.R.a[k] = (k % 4 == 0) ? 1.0 : 0.0; b2.R.a[k] = (k % 4 == 0) ? 1.0 : 0.0; }
    b2.R.a[0] = 0.8; b2.R.a[1] = -0.6; b2.R.a[3] = 0.6; b2.R.a[4] = 0.8;
    b2.x = mk(0.3 + j, -0.2, 0.249); b2.v = mk(0.01, -0.02 + j, -0.098); b2.L = mk(0.001, 0.002, -0.001);
    if (!ground) { b1.v = mk(0.02, 0.01, 0.05); b1.L = mk(0.002, -0.001, 0.003); }
    for (int k = 0; k < nc; k++) {
        cs[k].n = mk(0, 0, 1);
        cs[k].p = mk(b2.x.x + ((k & 1) ? 0.4 : -0.4) + 0.01 * k, b2.x.y + ((k & 2) ? 0.3 : -0.3), 0.0);
        cs[k].pen = -0.001 - 1e-4 * k; cs[k].feat = 0;
    }
}
template <int MODE>
__global__ void k(int nc, int ground, double* out, long long* cyc)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    Dyn b1, b2; Par p1, p2; Contact cs[8];
    make_case(t, nc, b1, p1, b2, p2, cs, ground != 0);
    StepParams sp; sp.eps = 1e-4; sp.baumgarte = 0.2; sp.slop = 0.001; sp.max_corr = 1.0; sp.dt = 0.01; sp.iterations = 10; sp.friction = 1; sp.integrator = 0;
    __syncthreads();
    long long t0 = clock64();
    bool ok = true;
    if (MODE == 0) { CPData cp[8]; resolve_t<0>(b1, p1, b2, p2, cs, nc, cp, sp); }
    else if (MODE == 1) { CPFast cp[8]; ok = resolve_s<0, false>(b1, p1, b2, p2, cs, nc, cp, sp); }
    else { CPFast cp[8]; ok = resolve_s<0, true>(b1, p1, b2, p2, cs, nc, cp, sp); }
    long long t1 = clock64();
    out[t] = b2.v.x + b2.v.y + b2.v.z + b2.L.x + b2.L.y + b2.L.z + b1.L.x + (ok ? 0.0 : 1e9);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const char* names[3] = {"literal resolve_t", "straight-line resolve_s", "straight-line, static body 1"};
    for (int mode = 0; mode < 3; mode++)
        for (int nc : {1, 4, 8})
            for (int warps : {1, 4, 8}) {
                const int ground = mode == 2;
                for (int rep = 0; rep < 2; rep++) {
                    if (mode == 0) k<0><<<148, warps * 32>>>(nc, ground, out, cyc);
                    if (mode == 1) k<1><<<148, warps * 32>>>(nc, ground, out, cyc);
                    if (mode == 2) k<2><<<148, warps * 32>>>(nc, ground, out, cyc);
                    cudaDeviceSynchronize();
                }
                long long c; double o0;
                cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&o0, out, 8, cudaMemcpyDeviceToHost);
                printf("%-30s contacts %d warps/SM %d : %8lld cycles = %6.0f per contact-iteration (incl. setup)  check %.17g  %s\n", names[mode], nc, warps, c,
                       (double)c / (10.0 * nc), o0, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
