// ps_team.cuh — a TEAM of L lanes that works on ONE candidate pair (L = 2, 4, 8, 16, 32 consecutive lanes of a warp).
//
// The per-pair routines of the step are long chains of dependent FP64 operations; a round of the batch-wide step lasts
// as long as its slowest pair.  What surrounds the strictly sequential impulse loop (CollisionResponse.fs:141-143) can be
// spread over lanes: the 15 SAT axes (SAT.fs:228-238), the edges of the clipped polygon (GraphicsUtils.fs:69-107), the
// contacts of a manifold (CollisionDetection.fs:63-67).  Teams exchange values with warp shuffles / ballots restricted
// to their own lanes and keep small tables (box normals, axes, the polygon) in shared memory.
//
// The same source compiles for the host (tests/host_harness, test-only): there a team is L std::threads that meet at a
// barrier for every exchange, so the cooperative routines are checked against the oracle bit for bit without a GPU.
#pragma once
#include "ps_math.cuh"

#if !defined(__CUDACC__)
#include <atomic>
#include <condition_variable>
#include <mutex>
#endif

namespace pstm {

#if defined(__CUDACC__)

template <int L>
struct Team {
    unsigned mask;  // the team's lanes inside the warp
    int lane;       // 0..L-1
    int shift;      // first lane of the team inside the warp
    __device__ __forceinline__ Team()
    {
        const int wl = (int)(threadIdx.x & 31u);
        lane = wl % L;
        shift = wl - lane;
        mask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << shift);
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(mask, v, src, L); }
    __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(mask, v, src, L); }
    __device__ __forceinline__ double shfl_xor(double v, int m) const { return __shfl_xor_sync(mask, v, m, L); }
    __device__ __forceinline__ int shfl_xor(int v, int m) const { return __shfl_xor_sync(mask, v, m, L); }
    __device__ __forceinline__ int shfl_up(int v, int d) const { return __shfl_up_sync(mask, v, d, L); }
    // bit l = predicate of lane l of the team
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & ((L == 32) ? 0xffffffffu : ((1u << L) - 1u)); }
    __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0u; }
};

#else  // host emulation: L threads, one barrier

template <int L>
struct TeamShared {
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long long gen = 0;
    double dslot[L];
    int islot[L];
    unsigned bal = 0;
    void barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g = gen;
        if (++waiting == L) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

template <int L>
struct Team {
    TeamShared<L>* sh;
    int lane;
    Team(TeamShared<L>* s, int l) : sh(s), lane(l) {}
    void sync() const { sh->barrier(); }
    double shfl(double v, int src) const
    {
        sh->dslot[lane] = v; sh->barrier();
        double r = sh->dslot[src & (L - 1)]; sh->barrier();
        return r;
    }
    int shfl(int v, int src) const
    {
        sh->islot[lane] = v; sh->barrier();
        int r = sh->islot[src & (L - 1)]; sh->barrier();
        return r;
    }
    double shfl_xor(double v, int m) const { return shfl(v, lane ^ m); }
    int shfl_xor(int v, int m) const { return shfl(v, lane ^ m); }
    int shfl_up(int v, int d) const
    {
        sh->islot[lane] = v; sh->barrier();
        int r = lane >= d ? sh->islot[lane - d] : v; sh->barrier();
        return r;
    }
    unsigned ballot(bool p) const
    {
        sh->islot[lane] = p ? 1 : 0; sh->barrier();
        unsigned r = 0;
        for (int l = 0; l < L; l++) r |= (unsigned)(sh->islot[l] != 0) << l;
        sh->barrier();
        return r;
    }
    bool any(bool p) const { return ballot(p) != 0u; }
};

#endif

// inclusive prefix sum of a small non-negative count over the team (Hillis-Steele, log2 L shuffles)
template <int L>
PSD int team_inclusive_scan(const Team<L>& tm, int v)
{
#pragma unroll
    for (int d = 1; d < L; d <<= 1) {
        const int o = tm.shfl_up(v, d);
        if (tm.lane >= d) v += o;
    }
    return v;
}

}  // namespace pstm
