// ps_wide.cuh — the batch-wide ("breadth-first") form of the step: the same per-body and per-pair routines as
// the fused per-world kernel (ps_step_kernel.cuh), but every stage is its own dense kernel over ALL worlds of the
// batch, and the wavefront levels become rounds of the whole batch:
//
//   wide_predict      thread per body : predicted copy + cull sphere                                   (phase A)
//   wide_near_count / wide_near_scan / wide_near_fill  : ordered near list per world                  (phase B)
//   wide_levels       thread per world: level of every near pair, bucket = (level, pair kind)          (phase C)
//   wide_bucket_scan, wide_scatter    : all near pairs of the batch sorted by (level, kind)
//   per round r = 1..max level, over the pairs of that level in ALL worlds:
//     wide_integrate    thread per (pair, side): re-integrate a body an earlier round changed
//     wide_detect<kind> thread per pair, one launch per kind (sphere-sphere, sphere-box, box-box): narrow phase;
//                       colliding pairs are appended to the round's hit list, contacts go to a pool
//     wide_resolve      thread per colliding pair: N-iteration solve + write-back
//   wide_static_fold  thread per static body : angular-momentum fold in pair order                    (phase E)
//   wide_final        thread per body : final integration + publish                                   (phase F)
//
// Why: a pair is a long chain of dependent FP64 operations and a world offers ~10 independent pairs at a time, so
// inside one CTA most lanes idle or diverge (ncu, fused kernel: 9 of 32 threads per instruction).  Across the
// batch a round holds 10^5 pairs; sorted by kind, every warp runs 32 pairs of the same kind through the same code.
// The price is one pass through L2/HBM per stage instead of registers — this is the part of the path where the
// HBM roofline starts to mean something.
//
// Exactness is the fused kernel's: pairs of one level touch disjoint dynamic bodies (so a round may run them in
// any order), the cull is conservative and verified on every re-integration.  A world that violates its margin,
// whose static bodies do not sit at a fixed point of integrate(), or that overflows a table is flagged and redone
// from the published state by the fused kernel (which owns the 8x-margin and exact-serial fallbacks).
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include "ps_step_kernel.cuh"
#include "ps_coop.cuh"
#include "ps_tma.cuh"

namespace psw {
using namespace psk;

constexpr int kWideMaxLevels = 1022;
constexpr int kBuckets = (kWideMaxLevels + 2) * 4;
constexpr int kHitSlots = 8;  // hit counters per level (6 used)

struct __align__(16) Entry {  // one near pair of the batch, sorted by (level, kind); written and read as one uint4
    int w;           // world
    unsigned ij;     // i | j<<16 (ids inside the world)
    unsigned knc;    // index in the world's near list (for the hit flag, 24 bits) | contacts found by the narrow phase << 24
    unsigned cfirst; // first contact in the pool
};

struct WArgs {
    int n_worlds, n_bodies, max_n;
    int w_begin, w_end;          // the worlds this pass steps (the whole batch, or one chunk of ps_batch_step_host) ...
    int g_begin, g_end;          // ... and their slots; every table below is indexed by absolute world / slot
    const int* off;              // per world: first SLOT (capacity prefix: a world owns spare slots for AddObject)
    const int* cnt;              // per world: bodies
    const int* body_world;       // [NB] world of every slot
    double* dyn; double* cur; double* pred; const double* par; double* S; double* cull;
    unsigned char* pvalid; unsigned char* nhit; unsigned char* nprev; const unsigned char* isstat; const unsigned char* isbox;  // [NB]
    int* row;                    // per world N+1 at off+w
    unsigned char* lists; const long long* lists_off;
    int* K; int* maxl;           // per world
    unsigned* wflags;            // per world: != 0 -> redo with the fused kernel
    const unsigned char* always_fused;  // per world
    unsigned* bucket;            // [kBuckets+1] counts, then exclusive starts
    unsigned* bucket_cur;        // [kBuckets]
    Entry* entries; unsigned entries_cap;
    unsigned* hit_cnt;           // [level][kHitSlots]: sphere-sphere, sphere-box, box manifolds by size class A, B, C, D
    unsigned* hitlist;           // [entries_cap], a bucket's hits are listed from the bucket's start (box: class A; class D from its end)
    unsigned* hitlist2;          // [entries_cap], box buckets only: class B from the start, class C from the end
    Contact* cpool; unsigned cpool_cap; unsigned* cpool_cur;  // cursor per level
    unsigned* surv;              // [entries_cap] box-box pairs the quick test could not separate (listed from the bucket's start)
    unsigned* surv_cnt;          // per level
    int* nflag;                  // [1] worlds the wide pass gave up on this step (wide_account)
    int* nlev;                   // [1] highest level that holds a pair (wide_bucket_scan); the device-driven rounds stop there
    unsigned team_max;           // a round with at most this many box-box survivors gives every survivor a team of 8 lanes
    WorldStats* stats;
    unsigned* tmp;               // per world x 6 words of THIS step's wide pass: [0-1] tested | colliding | contacts packed in 64 bits (count_pair), [3] overflow, [4] ref_exc, [5] bodies that left their margin
    unsigned long long* tmp_pen; // per world: max |penetration| bits of this step
    RecPair* rec_pairs; RecContact* rec_contacts; int* rec_counts; int rec_pair_cap, rec_contact_cap;
    StepParams sp;
    int margin_ncap;
    double margin_scale;         // safety factor on the motion margin (a violated margin costs a whole fused redo)
};

enum { WF_MARGIN = 1, WF_STATIC = 2, WF_CAPACITY = 4 };
#ifndef PS_WIDE_DETECT_MINB
#define PS_WIDE_DETECT_MINB 4  // stage 1 (sphere narrow phase + box-box quick test): 4 CTAs of 128 threads per SM, up to 128 registers
#endif
#ifndef PS_WIDE_MINB
#define PS_WIDE_MINB 2  // resident 128-thread CTAs per SM the box-box and resolve kernels are compiled for (2 -> up to 255 registers)
#endif

// Every pair of a round reserves its slot in the round's lists from the SAME counter.  One atomic per thread on one
// address serialises in L2 (a round of 25 K box pairs spent most of its 140 us there, round 1 with 10^6 pairs
// milliseconds); the lanes of a warp that arrive together reserve together and share one atomic.
__device__ __forceinline__ unsigned agg_reserve(unsigned* counter, unsigned n)
{
    namespace cg = cooperative_groups;
    cg::coalesced_group g = cg::coalesced_threads();
    const unsigned pre = cg::exclusive_scan(g, n);
    const unsigned total = g.shfl(pre + n, g.size() - 1);
    unsigned base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, total);
    return g.shfl(base, 0) + pre;
}
// per-world counters of this step's wide pass: tested (20 bits) | colliding (20 bits) | contacts (24 bits) in one
// 64-bit add per pair (a world has at most 2^19 pairs)
__device__ __forceinline__ void count_pair(const WArgs& a, int w, unsigned hit, unsigned contacts)
{
    atomicAdd(reinterpret_cast<unsigned long long*>(a.tmp + 6 * (size_t)w),
              1ull | ((unsigned long long)hit << 20) | ((unsigned long long)contacts << 40));
}

PSD unsigned* w_near(const WArgs& a, int w) { return reinterpret_cast<unsigned*>(a.lists + a.lists_off[w]); }
PSD size_t w_np(const WArgs& a, int w) { size_t N = (size_t)a.cnt[w]; return N * (N > 0 ? N - 1 : 0) / 2; }
PSD unsigned short* w_level(const WArgs& a, int w) { return reinterpret_cast<unsigned short*>(a.lists + a.lists_off[w] + 4 * w_np(a, w)); }
PSD unsigned char* w_hit(const WArgs& a, int w) { return a.lists + a.lists_off[w] + 8 * w_np(a, w); }

// ---- per step reset ------------------------------------------------------------------------------------
__global__ void wide_reset(WArgs a)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= kBuckets) a.bucket[t] = 0;
    if (t < kBuckets) a.bucket_cur[t] = 0;
    for (int q = t; q < (kWideMaxLevels + 2) * kHitSlots; q += gridDim.x * blockDim.x) a.hit_cnt[q] = 0;
    if (t < kWideMaxLevels + 2) { a.cpool_cur[t] = 0; a.surv_cnt[t] = 0; }
    if (t == 0) { a.nlev[0] = 0; a.nflag[0] = 0; }
    const int w = a.w_begin + t;
    if (w < a.w_end) {
        a.wflags[w] = a.always_fused[w] ? WF_STATIC : 0u;
        a.K[w] = 0; a.maxl[w] = 0;
        a.stats[w].nan_flag = 0;
        for (int k = 0; k < 6; k++) a.tmp[6 * w + k] = 0;
        a.tmp_pen[w] = 0ull;
        if (a.rec_counts) { a.rec_counts[2 * w] = 0; a.rec_counts[2 * w + 1] = 0; }
    }
}

// ---- A: predicted copies + cull spheres (thread per body, tile of 128 bodies per CTA) ---------------------------------
// The CTA's tile of the published state (128 x 144 B, contiguous) comes in with ONE bulk copy (ps_tma.cuh); it leaves
// again as the working copy `cur` straight from shared memory — no thread touches it — and the predicted records are
// written to a second shared tile and leave with one bulk copy as well.  Parameters are read with __ldg (128 B per
// thread = one full line each; at a 128-byte stride a shared tile of them would put a quarter warp on one bank group).
constexpr int kTile = 128;
constexpr size_t kTileBytes = (size_t)kTile * kDynF * sizeof(double);  // 18 432 B
__host__ __device__ inline size_t predict_smem_bytes() { return 2 * kTileBytes; }
__global__ void __launch_bounds__(kTile) wide_predict(WArgs a)
{
    extern __shared__ __align__(128) unsigned char tile_raw[];
    __shared__ __align__(8) uint64_t bar;
    double* s_in = reinterpret_cast<double*>(tile_raw);
    double* s_pred = reinterpret_cast<double*>(tile_raw + kTileBytes);
    const int g0 = a.g_begin + blockIdx.x * kTile;
    const int n = min(kTile, a.g_end - g0);
    const unsigned bytes = (unsigned)n * kDynF * (unsigned)sizeof(double);
    if (threadIdx.x == 0) pst_tma::mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        pst_tma::mbar_arrive_expect_tx(&bar, bytes);
        pst_tma::bulk_g2s(s_in, a.dyn + (size_t)g0 * kDynF, bytes, &bar);   // the published state is the step's input
    }
    const int t = threadIdx.x, g = g0 + t;
    Par p;
    if (t < n) ld_par(a.par, 0, g, p);                                      // (overlaps the bulk copy)
    pst_tma::mbar_wait(&bar, 0);
    if (threadIdx.x == 0) {                                                  // working copy = the tile as it came in
        pst_tma::bulk_s2g(a.cur + (size_t)g0 * kDynF, s_in, bytes);
        pst_tma::bulk_commit();
    }
    const int wq = t < n ? a.body_world[g] : 0;
    if (t < n && g - a.off[wq] < a.cnt[wq]) {  // (a spare slot holds no body)
        const int w = wq, o = a.off[w], N = a.cnt[w], b = g - o;
        Dyn cur, pr;
        ld_dyn(s_in, 0, t, cur);
        integrate(cur, p, a.sp, pr);
        st_dyn(s_pred, 0, t, pr);
        a.pvalid[g] = 1; a.nhit[g] = 0;
        if (p.is_static && !pose_fixed(cur, pr)) {  // the PREDICTED pose has to be the fixed point (see ps_step_kernel.cuh, phase A)
            Dyn pr2;
            integrate_slow(pr, p, a.sp, pr2);
            if (!pose_fixed(pr, pr2)) atomicOr(&a.wflags[w], (unsigned)WF_STATIC);
        }
        double* cu = a.cull + (size_t)o * kCullF;
        cu[b] = pr.x.x; cu[N + b] = pr.x.y; cu[2 * N + b] = pr.x.z;
        double speed = sqrt(pr.v.x * pr.v.x + pr.v.y * pr.v.y + pr.v.z * pr.v.z);
        double acc = p.inv_mass * sqrt(p.F.x * p.F.x + p.F.y * p.F.y + p.F.z * p.F.z);
        int np_ = a.nprev[g];
        double nn = (double)(np_ < a.margin_ncap ? np_ : a.margin_ncap) + 2.0;
        const double mscale = a.margin_scale * (a.stats[w].boost > 0 ? 2.0 : 1.0);
        double m = p.is_static ? 0.0 : mscale * (0.01 + nn * a.sp.dt * (speed + nn * acc * a.sp.dt + 0.5));
        if (!(m == m)) m = 0.0;
        double rs, rb;
        if (p.shape == 0) { rs = p.dims[0]; rb = 1.7320508075688774 * p.dims[0]; }
        else { double hx = 0.5 * p.dims[0], hy = 0.5 * p.dims[1], hz = 0.5 * p.dims[2]; rs = rb = sqrt(hx * hx + hy * hy + hz * hz); }
        const double slack = 1.0 + 1e-6;
        cu[3 * N + b] = rs * slack + 1e-9 + m;
        cu[4 * N + b] = rb * slack + 1e-9 + m;
        cu[5 * N + b] = m * m;
        cu[6 * N + b] = 0.0;
    }
    pst_tma::fence_proxy_async();  // the predicted tile was written by threads: make it visible to the copy engine
    __syncthreads();
    if (threadIdx.x == 0) {
        pst_tma::bulk_s2g(a.pred + (size_t)g0 * kDynF, s_pred, bytes);
        pst_tma::bulk_commit();
        pst_tma::bulk_wait_read();  // both stores have read their tiles: the CTA may go
    }
}

// ---- B: near list ---------------------------------------------------------------------------------------
PSD bool wide_near_test(const WArgs& a, const double* cu, int o, int N, int i, int j)
{
    if (a.isstat[o + i] && a.isstat[o + j]) return false;
    double dx = cu[i] - cu[j], dy = cu[N + i] - cu[N + j], dz = cu[2 * N + i] - cu[2 * N + j];
    double dd = dx * dx + dy * dy + dz * dz;
    double ri = a.isbox[o + j] ? cu[4 * N + i] : cu[3 * N + i];
    double rj = a.isbox[o + i] ? cu[4 * N + j] : cu[3 * N + j];
    double r = ri + rj;
    return !(dd > r * r);
}
// One CTA per world: the cull spheres of the world are staged in shared memory once; a thread takes rows i and N-1-i (so
// that every thread tests about N pairs), keeps the near partners j > i of a row as a bit mask in shared memory, a block
// scan over the row counts gives the row starts, and the rows are written from their masks — the list comes out in the
// Dummy order (i ascending, j ascending) with every pair tested once.  (First version: count / scan / fill over global
// memory, three launches, every pair tested twice: 0.87 ms per C3 step.)  Worlds of more than kNearMaskN bodies keep the
// masks out of shared memory and test twice.
constexpr int kNearT = 64;
constexpr int kNearMaskN = 256;
__host__ __device__ inline size_t near_smem_bytes(int max_n)
{
    const size_t words = max_n <= kNearMaskN ? (size_t)max_n * ((max_n + 31) / 32) : 0;
    return (size_t)max_n * (5 * sizeof(double) + 2 + sizeof(int)) + words * sizeof(unsigned) + 32;
}
__global__ void __launch_bounds__(kNearT) wide_near_list(WArgs a)
{
    extern __shared__ __align__(16) unsigned char near_raw[];
    __shared__ int wsum[kNearT / 32];
    __shared__ int carry;
    const int w = a.w_begin + blockIdx.x;
    const int o = a.off[w], N = a.cnt[w];
    double* sx = reinterpret_cast<double*>(near_raw);          // centre x | y | z | radius vs sphere | radius vs box
    int* rowcnt = reinterpret_cast<int*>(near_raw + (size_t)a.max_n * 5 * sizeof(double));
    unsigned* mask = reinterpret_cast<unsigned*>(rowcnt + a.max_n);
    const bool use_mask = a.max_n <= kNearMaskN;
    const int mw = (a.max_n + 31) / 32;  // mask words per row
    unsigned char* sbox = reinterpret_cast<unsigned char*>(mask + (use_mask ? (size_t)a.max_n * mw : 0));
    unsigned char* sstat = sbox + a.max_n;
    const double* cu = a.cull + (size_t)o * kCullF;
    for (int b = threadIdx.x; b < N; b += kNearT) {
        sx[b] = cu[b]; sx[N + b] = cu[N + b]; sx[2 * N + b] = cu[2 * N + b]; sx[3 * N + b] = cu[3 * N + b]; sx[4 * N + b] = cu[4 * N + b];
        sbox[b] = a.isbox[o + b]; sstat[b] = a.isstat[o + b];
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    auto near_ij = [&](int i, int j) {  // wide_near_test on the staged copy (same operations)
        if (sstat[i] && sstat[j]) return false;
        const double dx = sx[i] - sx[j], dy = sx[N + i] - sx[N + j], dz = sx[2 * N + i] - sx[2 * N + j];
        const double dd = dx * dx + dy * dy + dz * dz;
        const double ri = sbox[j] ? sx[4 * N + i] : sx[3 * N + i];
        const double rj = sbox[i] ? sx[4 * N + j] : sx[3 * N + j];
        const double r = ri + rj;
        return !(dd > r * r);
    };
    // rows of this thread: t, N-1-t, t + T, N-1-(t + T), ... (the first half of the rows paired with the second)
    const int half = (N + 1) / 2;
    for (int t = threadIdx.x; t < half; t += kNearT)
        for (int side = 0; side < 2; side++) {
            const int i = side ? N - 1 - t : t;
            if (side && i == t) break;
            int cn = 0;
            if (use_mask) {
                unsigned* m = mask + (size_t)i * mw;
                // near_ij with the row's own values in registers (the mask stores in between made the compiler re-read them
                // from shared memory for every pair: 8 of the 11 loads of a test) and the "both static" case as a predicate
                const double xi = sx[i], yi = sx[N + i], zi = sx[2 * N + i], r3i = sx[3 * N + i], r4i = sx[4 * N + i];
                const bool si = sstat[i] != 0;
                const double* rjv = sx + (sbox[i] ? 4 * N : 3 * N);
                for (int q = 0; q < mw; q++) {
                    unsigned bits = 0;
                    const int j0 = q * 32;
                    if (j0 + 31 > i) {
                        const int jb = j0 > i + 1 ? j0 : i + 1, je = N < j0 + 32 ? N : j0 + 32;
#pragma unroll 4
                        for (int j = jb; j < je; j++) {
                            const double dx = xi - sx[j], dy = yi - sx[N + j], dz = zi - sx[2 * N + j];
                            const double dd = dx * dx + dy * dy + dz * dz;
                            const double ri = sbox[j] ? r4i : r3i;
                            const double r = ri + rjv[j];
                            const bool nr = !(dd > r * r) && !(si && sstat[j]);
                            bits |= (nr ? 1u : 0u) << (j - j0);
                        }
                    }
                    m[q] = bits;
                    cn += __popc(bits);
                }
            } else {
                for (int j = i + 1; j < N; j++) cn += near_ij(i, j) ? 1 : 0;
            }
            rowcnt[i] = cn;
        }
    __syncthreads();
    // exclusive scan of the row counts, in row order
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i0 = 0; i0 < N; i0 += kNearT) {
        const int i = i0 + (int)threadIdx.x;
        const int cn = i < N ? rowcnt[i] : 0;
        int x = cn;
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        int base = carry;
        for (int q = 0; q < wid; q++) base += wsum[q];
        if (i < N) rowcnt[i] = base + x - cn;
        __syncthreads();
        if (threadIdx.x == kNearT - 1) carry = base + x;
        __syncthreads();
    }
    unsigned* near = w_near(a, w);
    unsigned char* hit = w_hit(a, w);
    for (int t = threadIdx.x; t < half; t += kNearT)
        for (int side = 0; side < 2; side++) {
            const int i = side ? N - 1 - t : t;
            if (side && i == t) break;
            int q = rowcnt[i];
            if (use_mask) {
                const unsigned* m = mask + (size_t)i * mw;
                for (int qw = 0; qw < mw; qw++) {
                    unsigned bits = m[qw];
                    while (bits) {
                        const int j = qw * 32 + __ffs(bits) - 1;
                        bits &= bits - 1;
                        near[q] = (unsigned)i | ((unsigned)j << 16);
                        hit[q] = 0;
                        q++;
                    }
                }
            } else {
                for (int j = i + 1; j < N; j++)
                    if (near_ij(i, j)) {
                        near[q] = (unsigned)i | ((unsigned)j << 16);
                        hit[q] = 0;
                        q++;
                    }
            }
        }
    if (threadIdx.x == 0) a.K[w] = carry;
}

// ---- C: levels + buckets ---------------------------------------------------------------------------------
PSD int pair_kind(const WArgs& a, int o, int i, int j)  // 0 sphere-sphere, 1 sphere/box mixed, 2 box-box
{
    return (int)a.isbox[o + i] + (int)a.isbox[o + j];
}
// A CTA owns kLvlWorlds consecutive worlds.  The level of a near pair depends on the previous near pair of either
// body, so one thread walks a world's list in order — with the per-body "last level" table in shared memory, so
// that the dependent chain never waits on global memory (with the table in global memory this kernel took 2 ms per
// C3 step).  Everything that is not a chain — the (level, kind) histogram here, the scatter below — is done by
// the whole CTA with coalesced reads of the lists and warp-aggregated shared-memory counters.
constexpr int kLvlWorlds = 32;
constexpr int kLvlThreads = 128;
__host__ __device__ inline int lvl_stride(int max_n) { return 2 * ((((max_n + 1) / 2)) | 1); }  // unsigned shorts per world; odd word count
__host__ __device__ inline size_t lvl_smem_bytes(int max_n) { return (size_t)kLvlWorlds * lvl_stride(max_n) * sizeof(unsigned short); }

// one pass of the CTA over the near lists of its worlds: f(bucket) per pair, all lanes of a warp call f together
template <class F>
__device__ __forceinline__ void for_cta_pairs(const WArgs& a, int w0, int wend, F f)
{
    for (int w = w0; w < wend; w++) {
        if (__ldcg(&a.wflags[w])) continue;  // CTA-uniform (read past L1: a thread of this CTA may just have set it)
        const int o = a.off[w], K = a.K[w];
        const unsigned* near = w_near(a, w);
        const unsigned short* level = w_level(a, w);
        for (int k0 = 0; k0 < K; k0 += kLvlThreads) {  // warp-uniform trip count
            const int k = k0 + (int)threadIdx.x;
            const bool valid = k < K;
            unsigned ij = 0; int b = -1;
            if (valid) {
                ij = near[k];
                b = (int)level[k] * 4 + pair_kind(a, o, ij & 0xffff, ij >> 16);
            }
            f(w, k, ij, b, valid);
        }
    }
}
// add 1 to counter[b] for every valid lane; returns the lane's slot (old value + rank among the lanes of the same b)
__device__ __forceinline__ unsigned warp_agg_inc(unsigned* counter, int b, bool valid)
{
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? b : -1 - (int)(threadIdx.x & 31));
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (valid && lane == leader) base = atomicAdd(&counter[b], (unsigned)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned)__popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kLvlThreads) wide_levels(WArgs a)
{
    extern __shared__ __align__(16) unsigned char lv_raw[];
    __shared__ unsigned hist[kBuckets];
    unsigned short* lastS = reinterpret_cast<unsigned short*>(lv_raw);  // per body: bit15 static | last level
    const int stride = lvl_stride(a.max_n);
    const int w0 = a.w_begin + blockIdx.x * kLvlWorlds, wend = min(w0 + kLvlWorlds, a.w_end);
    for (int b = threadIdx.x; b < kBuckets; b += kLvlThreads) hist[b] = 0;
    for (int g = a.off[w0] + (int)threadIdx.x; g < a.off[wend]; g += kLvlThreads) {
        const int w = a.body_world[g];
        lastS[(w - w0) * stride + (g - a.off[w])] = a.isstat[g] ? 0x8000 : 0;
    }
    __syncthreads();
    const int w = w0 + (int)threadIdx.x;
    if (threadIdx.x < kLvlWorlds && w < wend && !a.wflags[w]) {
        const int K = a.K[w];
        const unsigned* __restrict__ near = w_near(a, w);
        unsigned short* __restrict__ level = w_level(a, w);
        unsigned short* L = lastS + threadIdx.x * stride;
        int maxl = 0;
        bool over = false;
        for (int k0 = 0; k0 < K && !over; k0 += 4) {
            unsigned q[4];
            if (k0 + 4 <= K) {
                uint4 t = *reinterpret_cast<const uint4*>(near + k0);  // the list starts 16-byte aligned
                q[0] = t.x; q[1] = t.y; q[2] = t.z; q[3] = t.w;
            } else {
                for (int u = 0; u < 4; u++) q[u] = (k0 + u < K) ? near[k0 + u] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k0 + u >= K) break;
                const int i = q[u] & 0xffff, j = q[u] >> 16;
                const unsigned vi = L[i], vj = L[j];
                const int li = (vi & 0x8000u) ? 0 : (int)vi, lj = (vj & 0x8000u) ? 0 : (int)vj;
                const int l = (li > lj ? li : lj) + 1;
                if (l > kWideMaxLevels) { over = true; break; }
                level[k0 + u] = (unsigned short)l;
                if (!(vi & 0x8000u)) L[i] = (unsigned short)l;
                if (!(vj & 0x8000u)) L[j] = (unsigned short)l;
                if (l > maxl) maxl = l;
            }
        }
        if (over) atomicOr(&a.wflags[w], (unsigned)WF_CAPACITY);
        else a.maxl[w] = maxl;
    }
    __syncthreads();  // level[] of the CTA's worlds is complete (and wflags settled for them)
    __threadfence_block();
    for_cta_pairs(a, w0, wend, [&](int, int, unsigned, int b, bool valid) { warp_agg_inc(hist, b, valid); });
    __syncthreads();
    for (int b = threadIdx.x; b < kBuckets; b += kLvlThreads)
        if (hist[b]) atomicAdd(&a.bucket[b], hist[b]);
}
__global__ void wide_bucket_scan(WArgs a)  // one warp: kBuckets is ~4 K
{
    __shared__ unsigned part[32];
    const int lane = threadIdx.x, per = kBuckets / 32;
    unsigned s = 0;
    for (int b = lane * per; b < (lane + 1) * per; b++) s += a.bucket[b];
    part[lane] = s;
    __syncwarp();
    unsigned base = 0;
    for (int l = 0; l < lane; l++) base += part[l];
    int top = 0;
    for (int b = lane * per; b < (lane + 1) * per; b++) { unsigned t = a.bucket[b]; a.bucket[b] = base; base += t; if (t) top = b / 4; }
    if (lane == 31) a.bucket[kBuckets] = base;
    if (top) atomicMax(a.nlev, top);
}
__global__ void __launch_bounds__(kLvlThreads) wide_scatter(WArgs a)  // same CTA -> worlds map as wide_levels
{
    __shared__ unsigned cnt[kBuckets];   // pairs of this CTA per bucket, then the CTA's base inside the bucket
    __shared__ unsigned cur[kBuckets];
    for (int b = threadIdx.x; b < kBuckets; b += kLvlThreads) { cnt[b] = 0; cur[b] = 0; }
    const int w0 = a.w_begin + blockIdx.x * kLvlWorlds, wend = min(w0 + kLvlWorlds, a.w_end);
    if (a.bucket[kBuckets] > a.entries_cap) {  // the sorted list does not fit: the fused kernel takes the step
        for (int w = w0 + (int)threadIdx.x; w < wend; w += kLvlThreads) atomicOr(&a.wflags[w], (unsigned)WF_CAPACITY);
        return;
    }
    __syncthreads();
    for_cta_pairs(a, w0, wend, [&](int, int, unsigned, int b, bool valid) { warp_agg_inc(cnt, b, valid); });
    __syncthreads();
    for (int b = threadIdx.x; b < kBuckets; b += kLvlThreads)
        if (cnt[b]) cnt[b] = a.bucket[b] + atomicAdd(&a.bucket_cur[b], cnt[b]);
    __syncthreads();
    for_cta_pairs(a, w0, wend, [&](int w, int k, unsigned ij, int b, bool valid) {
        const unsigned slot = warp_agg_inc(cur, b, valid);
        if (valid) *reinterpret_cast<uint4*>(&a.entries[cnt[b] + slot]) = make_uint4((unsigned)w, ij, (unsigned)k, 0u);
    });
}

// ---- rounds ---------------------------------------------------------------------------------------------
// box-box, stage 1: the two most promising axes.  Most near box pairs are separated and end here; the rest are
// compacted so that stage 2 (full SAT + contact generation, wide_detect<2>) runs on dense warps
// (ncu before the split: 5 of 32 threads per instruction in the box-box kernel).
__device__ __forceinline__ void wide_bb_quick_body(const WArgs& a, unsigned first, unsigned count, int level, unsigned t)
{
    if (t >= count) return;
    const Entry& e = a.entries[first + t];
    const int w = e.w;
    if (a.wflags[w]) return;
    const int o = a.off[w];
    const int i = e.ij & 0xffff, j = e.ij >> 16;
    Par pi, pj; Dyn di, dj;
    ld_par(a.par, 0, o + i, pi);
    ld_par(a.par, 0, o + j, pj);
    ld_dyn(a.pred, 0, o + i, di);
    ld_dyn(a.pred, 0, o + j, dj);
    OBox o1, o2;
    make_obox_i(di, pi, o1);
    make_obox_i(dj, pj, o2);
    if (bb_quick_separated_t<true>(di, dj, o1, o2)) {
        count_pair(a, w, 0u, 0u);  // tested, no collision
        return;
    }
    unsigned slot = agg_reserve(&a.surv_cnt[level], 1u);
    a.surv[first + slot] = first + t;
}

// a colliding pair whose nc contacts sit at cpool[cf ..): entry, the round's hit list, max penetration, recording
__device__ __forceinline__ void wide_publish_hit(const WArgs& a, Entry& e, unsigned ei, unsigned first, unsigned bucket_count, int level, int kind, int w,
                                                 int i, int j, int nc, unsigned cf, double mp)
{
    e.knc |= (unsigned)nc << 24; e.cfirst = cf;
    // Box pairs: a warp of the solve walks as many contacts as its LARGEST manifold holds (the walk is warp-uniform), so the
    // manifolds are listed by size class — A: 1-2 contacts (edge contacts, corners), B: 3-4 (a face), C: 5-6, D: 7 and more —
    // each class in its own stretch: A from the front and D from the back of the box bucket's region of the hit list, B
    // and C the same way in a second list.  (ncu before any sorting: 16 of 32 threads per instruction in the solve; with
    // two classes every "large" warp still ran the 8-contact walk.)
    if (kind == 2) {
        const int cls = nc <= 2 ? 0 : (nc <= 4 ? 1 : (nc <= 6 ? 2 : 3));
        unsigned slot = 0;
#pragma unroll
        for (int c = 0; c < 4; c++)  // (agg_reserve pools the lanes that arrive together on ONE counter: one class at a time)
            if (cls == c) slot = agg_reserve(&a.hit_cnt[level * kHitSlots + 2 + c], 1u);
        unsigned* list = (cls == 0 || cls == 3) ? a.hitlist : a.hitlist2;
        list[(cls >= 2) ? first + bucket_count - 1u - slot : first + slot] = ei;
    } else {
        const unsigned slot = agg_reserve(&a.hit_cnt[level * kHitSlots + kind], 1u);
        a.hitlist[first + slot] = ei;
    }
    const unsigned long long mpb = (unsigned long long)__double_as_longlong(mp);
    if (mpb > __ldcg(&a.tmp_pen[w])) atomicMax(&a.tmp_pen[w], mpb);  // the maximum rarely moves
    if (a.rec_pairs) {
        int* cnt = a.rec_counts + 2 * w;
        int slot2 = atomicAdd(&cnt[0], 1);
        int f2 = atomicAdd(&cnt[1], nc);
        if (slot2 < a.rec_pair_cap && f2 + nc <= a.rec_contact_cap) {
            RecPair rp; rp.i = i; rp.j = j; rp.first = f2; rp.count = nc;
            a.rec_pairs[(size_t)w * a.rec_pair_cap + slot2] = rp;
            RecContact* out = a.rec_contacts + (size_t)w * a.rec_contact_cap + f2;
            for (int k = 0; k < nc; k++) {
                const Contact c = a.cpool[cf + k];
                RecContact r;
                r.n[0] = c.n.x; r.n[1] = c.n.y; r.n[2] = c.n.z;
                r.p[0] = c.p.x; r.p[1] = c.p.y; r.p[2] = c.p.z;
                r.pen = c.pen; r.feat = c.feat; r.pad = 0;
                out[k] = r;
            }
        }
    }
}

// narrow phase of one kind: thread per entry of bucket (level, kind)
template <int KIND>
__device__ __forceinline__ void wide_detect_body(const WArgs& a, unsigned first, unsigned count, int level, unsigned t)
{
    PS_CLK_BEGIN(KIND == 2 && level >= 3)
    unsigned ei = first + t;
    if (KIND == 2) {  // box-box: only the pairs stage 1 could not separate
        if (t >= a.surv_cnt[level]) return;
        ei = a.surv[first + t];
    } else if (t >= count) return;
    Entry& e = a.entries[ei];
    const int w = e.w;
    if (a.wflags[w]) return;
    const int o = a.off[w];
    const int i = e.ij & 0xffff, j = e.ij >> 16;
    Par pi, pj; Dyn di, dj;
    ld_par(a.par, 0, o + i, pi);
    ld_par(a.par, 0, o + j, pj);
    ld_dyn(a.pred, 0, o + i, di);
    ld_dyn(a.pred, 0, o + j, dj);
    // the inlined, unrolled forms of the narrow phase (ps_physics.cuh, F = true): a round lasts as long as its slowest
    // pair, so the dependent chain of ONE pair is what counts here
    PairScratch scb;   // box-box only: clip buffers, manifold
    Contact cs1[1];    // a pair with a sphere has one contact
    Contact* const cs = (KIND == 2) ? scb.cs : cs1;
    int err = 0, ovf = 0, nc;
    if (KIND == 0) {
        nc = detect_ss_t<true>(di, pi, dj, pj, cs);
    } else if (KIND == 2) {
        PS_CLK(8)
        nc = detect_bb_t<true>(di, pi, dj, pj, a.sp.eps, cs, err, ovf, scb, true, level >= 3);  // the quick test has already failed to separate
        PS_CLK(9)
    } else {
        OBox ob;
        nc = detect_sphere_pair_t<true>(di, pi, dj, pj, a.sp.eps, cs, err, ob);
    }
    unsigned* tm = a.tmp + 6 * (size_t)w;
    if (err) atomicAdd(&tm[4], 1u);
    if (ovf) atomicAdd(&tm[3], 1u);
    count_pair(a, w, nc > 0 ? 1u : 0u, nc > 0 ? (unsigned)nc : 0u);
    if (nc <= 0) return;
    unsigned cf = agg_reserve(&a.cpool_cur[level], (unsigned)nc);
    if (cf + (unsigned)nc > a.cpool_cap) { atomicOr(&a.wflags[w], (unsigned)WF_CAPACITY); return; }
    for (int k = 0; k < nc; k++) a.cpool[cf + k] = cs[k];
    double mp = 0.0;
    for (int k = 0; k < nc; k++) mp = fmax(mp, fabs(cs[k].pen));
    wide_publish_hit(a, e, ei, first, count, level, KIND, w, i, j, nc, cf, mp);
}

// the literal box-box routine on ONE survivor entry (fallback of the team form); a real function
static __device__ __noinline__ void wide_bb_literal(const WArgs& a, unsigned ei, unsigned first, unsigned c2, int level)
{
    Entry& e = a.entries[ei];
    const int w = e.w, o = a.off[w];
    const int i = e.ij & 0xffff, j = e.ij >> 16;
    Par pi, pj; Dyn di, dj;
    ld_par(a.par, 0, o + i, pi);
    ld_par(a.par, 0, o + j, pj);
    ld_dyn(a.pred, 0, o + i, di);
    ld_dyn(a.pred, 0, o + j, dj);
    PairScratch scb;
    int err = 0, ovf = 0;
    const int nc = detect_bb(di, pi, dj, pj, a.sp.eps, scb.cs, err, ovf, scb, false);
    unsigned* tm = a.tmp + 6 * (size_t)w;
    if (err) atomicAdd(&tm[4], 1u);
    if (ovf) atomicAdd(&tm[3], 1u);
    count_pair(a, w, nc > 0 ? 1u : 0u, nc > 0 ? (unsigned)nc : 0u);
    if (nc <= 0) return;
    unsigned cf = atomicAdd(&a.cpool_cur[level], (unsigned)nc);
    if (cf + (unsigned)nc > a.cpool_cap) { atomicOr(&a.wflags[w], (unsigned)WF_CAPACITY); return; }
    double mp = 0.0;
    for (int k = 0; k < nc; k++) { a.cpool[cf + k] = scb.cs[k]; mp = fmax(mp, fabs(scb.cs[k].pen)); }
    wide_publish_hit(a, e, ei, first, c2, level, 2, w, i, j, nc, cf, mp);
}

// the literal solve of one pair from the predicted records (fallback of the straight-line solve; a real function)
static __device__ __noinline__ void wide_resolve_literal(const WArgs& a, int gi, int gj, bool si, bool sj, const Contact* cs, int nc, Dyn& di, Dyn& dj)
{
    Par pi, pj;
    ld_par(a.par, 0, gi, pi);
    ld_par(a.par, 0, gj, pj);
    ld_dyn(a.pred, 0, gi, di);
    ld_dyn(a.pred, 0, gj, dj);
    if (si) di.L = mk(0.0, 0.0, 0.0);
    if (sj) dj.L = mk(0.0, 0.0, 0.0);
    CPData cp[PS_MAX_CONTACTS];
    resolve(di, pi, dj, pj, cs, nc, cp, a.sp);
}

// solve + write-back: thread per colliding pair of bucket (level, kind).  A body the solve changed is integrated
// again right here (its next use — the next pair or the final sweep — needs exactly this predicted copy), which
// keeps that work in a dense kernel; its margin only matters if a later near pair will test it.
// entry of hit t of slot KIND of a round: 0 sphere-sphere, 1 sphere-box, 2..5 box manifolds of size class A..D
// (wide_publish_hit); false past the end
__device__ __forceinline__ bool wide_pick_hit(const WArgs& a, int kind, unsigned first, unsigned c2, int level, unsigned t, unsigned& ei)
{
    if (t >= a.hit_cnt[level * kHitSlots + kind]) return false;
    const unsigned* list = (kind == 3 || kind == 4) ? a.hitlist2 : a.hitlist;
    ei = list[kind >= 4 ? first + c2 - 1u - t : first + t];
    return true;
}
template <bool BOX>
__device__ __forceinline__ void wide_resolve_entry(const WArgs& a, unsigned ei, int level, int only_side = -1, unsigned act = 0u);
template <int KIND>
__device__ __forceinline__ void wide_resolve_body(const WArgs& a, unsigned first, unsigned c2, int level, unsigned t)
{
    unsigned ei;
    if (wide_pick_hit(a, KIND, first, c2, level, t, ei)) wide_resolve_entry<(KIND >= 2)>(a, ei, level);
}
// only_side >= 0: TWO neighbouring lanes hold this pair and run the same solve on the same inputs; each writes back and
// re-integrates ONE of the two bodies (measured on the manifolds of 7 and more contacts that end a round: loads 9 K cycles,
// contact setup 18 K, iterations 148 K, write-back + two re-integrations 32 K)
template <bool BOX>
__device__ __forceinline__ void wide_resolve_entry(const WArgs& a, unsigned ei, int level, int only_side, unsigned act)
{
    constexpr int KIND = BOX ? 2 : 0;
    PS_CLK_BEGIN(KIND >= 2 && level >= 3)
    const Entry& e = a.entries[ei];
    const int w = e.w;
    const bool gone = a.wflags[w] != 0;
    if (BOX && only_side >= 0) act = __ballot_sync(act, !gone);  // (the two lanes of a pair read the same flag)
    if (gone) return;
    const int o = a.off[w], N = a.cnt[w];
    const int i = e.ij & 0xffff, j = e.ij >> 16;
    Par pi, pj; Dyn di, dj;
    ld_par(a.par, 0, o + i, pi);
    ld_par(a.par, 0, o + j, pj);
    ld_dyn(a.pred, 0, o + i, di);
    ld_dyn(a.pred, 0, o + j, dj);
    const bool si = pi.is_static, sj = pj.is_static;
    if (si) di.L = mk(0.0, 0.0, 0.0);
    if (sj) dj.L = mk(0.0, 0.0, 0.0);
    PS_CLK(12)
    // the straight-line solve (ps_physics.cuh, resolve_s); S1: body i is a static body at rest — the ground pairs of round 1
    const bool s1 = si && !sj && bits_of(di.v.x) == 0 && bits_of(di.v.y) == 0 && bits_of(di.v.z) == 0;
    bool ok;
    if (KIND >= 2) {
        const int nc = (int)(e.knc >> 24);
        CPFast cp[PS_MAX_CONTACTS];
        PS_CLK(13)
        const Contact* cs = a.cpool + e.cfirst;  // the manifold is read where the narrow phase left it
        if (only_side >= 0) {  // two lanes per manifold: each carries one body through the iterations (ps_physics.cuh, resolve_s2)
            const unsigned act2 = __ballot_sync(act, !s1);
            if (s1) ok = resolve_s<0, true>(di, pi, dj, pj, cs, nc, cp, a.sp);
            else ok = resolve_s2(act2, only_side != 0, di, pi, dj, pj, cs, nc, reinterpret_cast<CPFast2*>(cp), a.sp);
        } else
        ok = s1 ? resolve_s<0, true>(di, pi, dj, pj, cs, nc, cp, a.sp) : resolve_s<0, false>(di, pi, dj, pj, cs, nc, cp, a.sp);
        PS_CLK(14)
    } else {  // sphere pairs always carry exactly one contact
        Contact cs[1];
        cs[0] = a.cpool[e.cfirst];
        ok = s1 ? resolve_s<1, true>(di, pi, dj, pj, cs, 1, nullptr, a.sp) : resolve_s<1, false>(di, pi, dj, pj, cs, 1, nullptr, a.sp);
    }
    if (!ok) {  // a quotient outside the short division's range, a non-finite accumulator: the literal routine, same inputs
        PS_COLD_PATH;
        if (only_side <= 0) atomicAdd(&a.tmp[6 * (size_t)w + 2], 1u);  // (counted: PS_B200_WIDE_PROF prints the share of pairs that took this route)
        wide_resolve_literal(a, o + i, o + j, si, sj, a.cpool + e.cfirst, (int)(e.knc >> 24), di, dj);
    }
    w_hit(a, w)[e.knc & 0xffffffu] = 1;
    double* Sw = a.S + (size_t)o * 3 * kMaxStatics;
    double* cu = a.cull + (size_t)o * kCullF;
    // (measured: the two re-integrations as one two-lane straight-line computation (integrate_2) spill in this kernel —
    //  round 1 of C3 1.88 -> 2.01 ms, the other rounds unchanged — so they stay the literal routine, one after the other)
#pragma unroll 1
    for (int pass = 0; pass < (only_side >= 0 ? 1 : 2); pass++) {
        const int side = only_side >= 0 ? only_side : pass;  // (two lanes: both in the same pass, each with its own body)
        const int b = side ? j : i, other = side ? i : j;
        const bool stat = side ? sj : si;
        Dyn& d = side ? dj : di;
        const Par& p = side ? pj : pi;
        if (stat) {
            double* S = Sw + 3 * ((a.isstat[o + b] - 1) * N + other);
            S[0] = d.L.x; S[1] = d.L.y; S[2] = d.L.z;
            continue;
        }
        st_dyn(a.cur, 0, o + b, d);  // write back (CollisionResolver.fs:27-31)
        bump(&a.nhit[o + b]);
        Dyn pr;
        integrate(d, p, a.sp, pr);
        st_dyn(a.pred, 0, o + b, pr);  // pvalid stays 1: the predicted copy is current again
        // pr.x is where the reference tests the body in every pair that comes up before its next write-back, culled
        // pairs included: outside the margin the largest deviation is kept for wide_verify (this thread owns the body)
        d3 dd = pr.x - mk(cu[b], cu[N + b], cu[2 * N + b]);
        double q = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
        if (!(q <= cu[5 * N + b]) && !(q <= cu[6 * N + b])) cu[6 * N + b] = q;  // a NaN sticks
    }
    PS_CLK(15)
}

// One launch per stage and round: the CTA index selects the kind (CTA-uniform branch), so a round costs three
// launches instead of seven.
// stage 1: sphere-sphere + sphere-box narrow phase + box-box quick test
__global__ void __launch_bounds__(128) wide_round_detect(WArgs a, unsigned first, unsigned c0, unsigned c1, unsigned c2, int level)
{
    const unsigned b0 = (c0 + 127) / 128, b1 = (c1 + 127) / 128;
    const unsigned blk = blockIdx.x;
    if (blk < b0) wide_detect_body<0>(a, first, c0, level, blk * 128 + threadIdx.x);
    else if (blk < b0 + b1) wide_detect_body<1>(a, first + c0, c1, level, (blk - b0) * 128 + threadIdx.x);
    else wide_bb_quick_body(a, first + c0 + c1, c2, level, (blk - b0 - b1) * 128 + threadIdx.x);
}
// stage 2: full box-box SAT + contact generation on the survivors of the quick test
__global__ void __launch_bounds__(128, PS_WIDE_MINB) wide_round_bb(WArgs a, unsigned first_bb, unsigned c2, int level)
{
    wide_detect_body<2>(a, first_bb, c2, level, blockIdx.x * 128 + threadIdx.x);
}
// stage 3: solve + write-back of every colliding pair of the round
__global__ void __launch_bounds__(128, PS_WIDE_MINB) wide_round_resolve(WArgs a, unsigned first, unsigned c0, unsigned c1, unsigned c2, int level)
{
    const unsigned b0 = (c0 + 127) / 128, b1 = (c1 + 127) / 128;
    const unsigned blk = blockIdx.x;
    // (host-driven form: thread t of the box bucket takes manifold t of every size class)
    if (blk < b0) wide_resolve_body<0>(a, first, 0u, level, blk * 128 + threadIdx.x);
    else if (blk < b0 + b1) wide_resolve_body<1>(a, first + c0, 0u, level, (blk - b0) * 128 + threadIdx.x);
    else {
        wide_resolve_body<5>(a, first + c0 + c1, c2, level, (blk - b0 - b1) * 128 + threadIdx.x);
        wide_resolve_body<4>(a, first + c0 + c1, c2, level, (blk - b0 - b1) * 128 + threadIdx.x);
        wide_resolve_body<3>(a, first + c0 + c1, c2, level, (blk - b0 - b1) * 128 + threadIdx.x);
        wide_resolve_body<2>(a, first + c0 + c1, c2, level, (blk - b0 - b1) * 128 + threadIdx.x);
    }
}

// ---- box-box stage by TEAMS: 8 lanes per surviving pair (ps_coop.cuh) ----------------------------------------------
// The 15 SAT axes, the clipped polygon's edges and the manifold's contacts are spread over the team's lanes; a warp holds
// four pairs.  Used for the rounds whose survivors leave lanes idle anyway (a.team_max): there a pair's latency is the
// round's duration.  The contacts go to the pool straight from the lanes that made them; lane 0 does the bookkeeping.
__device__ __forceinline__ void wide_bb_team_body(const WArgs& a, unsigned first, unsigned c2, int level, unsigned slot, unsigned ns, psc::TeamScratch& ts)
{
    using namespace psc;
    if (slot >= ns) return;  // (the whole team)
    const pstm::Team<kTeam> tm;
    const unsigned ei = a.surv[first + slot];
    Entry& e = a.entries[ei];
    const int w = e.w;
    if (a.wflags[w]) return;
    const int o = a.off[w];
    const int i = e.ij & 0xffff, j = e.ij >> 16;
    Par pi, pj; Dyn di, dj;
    ld_par(a.par, 0, o + i, pi);
    ld_par(a.par, 0, o + j, pj);
    ld_dyn(a.pred, 0, o + i, di);
    ld_dyn(a.pred, 0, o + j, dj);
    TeamContacts out;
    int err = 0;
    bool lit = false;
    int nc = detect_bb_team(tm, di, pi, dj, pj, a.sp.eps, ts, out, err, lit);
    if (lit) {  // NaN penetrations / a clipped polygon of more than 16 vertices: lane 0 runs the literal routine
        PS_COLD_PATH;
        if (tm.lane == 0) wide_bb_literal(a, ei, first, c2, level);
        return;
    }
    unsigned cf = 0;
    if (tm.lane == 0) {
        if (err) atomicAdd(&a.tmp[6 * (size_t)w + 4], 1u);
        count_pair(a, w, nc > 0 ? 1u : 0u, nc > 0 ? (unsigned)nc : 0u);
        if (nc > 0) cf = agg_reserve(&a.cpool_cur[level], (unsigned)nc);
    }
    if (nc <= 0) return;
    cf = (unsigned)tm.shfl((int)cf, 0);
    if (cf + (unsigned)nc > a.cpool_cap) {
        if (tm.lane == 0) atomicOr(&a.wflags[w], (unsigned)WF_CAPACITY);
        return;
    }
    double mp = 0.0;
#pragma unroll
    for (int t = 0; t < 2; t++)
        if (out.ord[t] >= 0) {
            a.cpool[cf + out.ord[t]] = out.c[t];
            mp = fmax(mp, fabs(out.c[t].pen));
        }
#pragma unroll
    for (int m = kTeam / 2; m; m >>= 1) mp = fmax(mp, tm.shfl_xor(mp, m));
    tm.sync();  // the manifold is in the pool (lane 0 may read it back for the recording)
    if (tm.lane == 0) wide_publish_hit(a, e, ei, first, c2, level, 2, w, i, j, nc, cf, mp);
}

// ---- device-counted round kernels ------------------------------------------------------------------------------------
// One launch per stage and round as above, but the COUNTS are read from the bucket table in device memory and a CTA
// walks its chunks with a grid stride: the host sizes the grid from what the same round held a step or two ago (an
// estimate read back asynchronously, never waited for), and whatever the estimate, every pair of the round is
// processed.  A round the step does not have costs an empty launch.  No host round trip inside a step.
struct RoundCounts { unsigned first, c0, c1, c2; };
__device__ __forceinline__ RoundCounts round_counts(const WArgs& a, int r)
{
    RoundCounts q;
    q.first = a.bucket[r * 4];
    q.c0 = a.bucket[r * 4 + 1] - q.first; q.c1 = a.bucket[r * 4 + 2] - a.bucket[r * 4 + 1]; q.c2 = a.bucket[r * 4 + 3] - a.bucket[r * 4 + 2];
    return q;
}
__global__ void __launch_bounds__(128, PS_WIDE_DETECT_MINB) wide_round_detect_d(WArgs a, int r)
{
    if (r > a.nlev[0] || a.bucket[kBuckets] > a.entries_cap) return;
    const RoundCounts q = round_counts(a, r);
    const unsigned b0 = (q.c0 + 127) / 128, b1 = (q.c1 + 127) / 128, b2 = (q.c2 + 127) / 128;
    for (unsigned blk = blockIdx.x; blk < b0 + b1 + b2; blk += gridDim.x) {
        if (blk < b0) wide_detect_body<0>(a, q.first, q.c0, r, blk * 128 + threadIdx.x);
        else if (blk < b0 + b1) wide_detect_body<1>(a, q.first + q.c0, q.c1, r, (blk - b0) * 128 + threadIdx.x);
        else wide_bb_quick_body(a, q.first + q.c0 + q.c1, q.c2, r, (blk - b0 - b1) * 128 + threadIdx.x);
    }
}
__global__ void __launch_bounds__(128, PS_WIDE_MINB) wide_round_bb_d(WArgs a, int r)
{
    if (r > a.nlev[0] || a.bucket[kBuckets] > a.entries_cap) return;
    const RoundCounts q = round_counts(a, r);
    const unsigned ns = a.surv_cnt[r];
    for (unsigned blk = blockIdx.x; blk < (ns + 127) / 128; blk += gridDim.x)
        wide_detect_body<2>(a, q.first + q.c0 + q.c1, q.c2, r, blk * 128 + threadIdx.x);
}
// the same stage by teams of 8 lanes (16 pairs per CTA), for rounds whose survivors leave lanes idle anyway
__global__ void __launch_bounds__(128, PS_WIDE_MINB) wide_round_bb_team_d(WArgs a, int r)
{
    __shared__ psc::TeamScratch ts_all[16];
    if (r > a.nlev[0] || a.bucket[kBuckets] > a.entries_cap) return;
    const RoundCounts q = round_counts(a, r);
    const unsigned ns = a.surv_cnt[r];
    for (unsigned blk = blockIdx.x; blk < (ns + 15) / 16; blk += gridDim.x)
        wide_bb_team_body(a, q.first + q.c0 + q.c1, q.c2, r, blk * 16 + threadIdx.x / psc::kTeam, ns, ts_all[threadIdx.x / psc::kTeam]);
}
// MINB = resident CTAs per SM the register allocation allows: 2 (255 registers) for the rounds that wait for their longest
// manifold, 3 (168 registers, 12 warps per SM) for a round that is throughput bound — round 1, every body against the ground:
// 1.84 -> 1.68 ms; the other rounds are slower with it (110 -> 140 us)
// PAIRED (the rounds that wait for their slowest thread): a box manifold gets two neighbouring lanes.  Each lane carries one of
// the two bodies through the iterations (resolve_s2) and writes back and re-integrates that body.  Measured per phase on the
// manifolds of 7 and more contacts, one lane each: loads 9 K cycles, contact setup 18 K, iterations 148 K, write-back + two
// re-integrations 32 K.  kPairedFrom = first hit slot that is paired; C3, ms per step: nothing paired 11.86, classes C and D (4)
// 11.64, B too (3) 11.66, every box manifold (2) 11.49, sphere pairs too (0) 11.64.
#ifndef PS_PAIRED_FROM
#define PS_PAIRED_FROM 2
#endif
constexpr int kPairedFrom = PS_PAIRED_FROM;
template <int MINB, bool PAIRED>
__global__ void __launch_bounds__(128, MINB) wide_round_resolve_d(WArgs a, int r)
{
    if (r > a.nlev[0] || a.bucket[kBuckets] > a.entries_cap) return;
    const RoundCounts q = round_counts(a, r);
    const unsigned* hc = a.hit_cnt + r * kHitSlots;
    unsigned nb[6], tot = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) { nb[k] = (hc[k] * ((PAIRED && k >= kPairedFrom) ? 2u : 1u) + 127) / 128; tot += nb[k]; }
    const unsigned fbb = q.first + q.c0 + q.c1;
    for (unsigned blk = blockIdx.x; blk < tot; blk += gridDim.x) {  // the largest manifolds first: they are the long ones
        unsigned b = blk;
        int kind = 5;
        while (kind > 0 && b >= nb[kind]) { b -= nb[kind]; kind--; }
        const unsigned fk = kind >= 2 ? fbb : (kind == 1 ? q.first + q.c0 : q.first);
        unsigned ei, t = b * 128 + threadIdx.x;
        int only = -1;
        if (PAIRED && kind >= kPairedFrom) { only = (int)(t & 1u); t >>= 1; }
        const bool have = wide_pick_hit(a, kind, fk, q.c2, r, t, ei);
        const unsigned act = __ballot_sync(0xffffffffu, have);  // (the CTA is converged here: kind and b are CTA-uniform)
        if (!have) continue;
        if (kind >= 2) wide_resolve_entry<true>(a, ei, r, only, act);   // (one copy of each solver in the kernel)
        else wide_resolve_entry<false>(a, ei, r, only);
    }
}

// ---- the rounds from r_begin on in ONE cooperative launch -------------------------------------------------------------
// The catch-all behind the per-round launches above: the host enqueues launches for as many rounds as the previous steps
// had (plus a few); should this step have more, this kernel runs the rest — as many CTAs as are resident at once, a
// round is three stages separated by grid barriers, counts from the bucket table.  (Measured as the ONLY form, C3: 22.9 ms
// per step against 16.1 with a launch per stage: one register allocation for all stages, static chunk assignment.)
// A stage hands chunk c (32 consecutive pairs of one kind, or 4 team pairs) to warp  c mod #warps  in an order that
// walks the CTAs first (warp 0 of every CTA, then warp 1 ...): a small round spreads over all SMs.
__global__ void __launch_bounds__(128, PS_WIDE_MINB) wide_rounds(WArgs a, int r_begin)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ psc::TeamScratch ts_all[16];
    if (a.bucket[kBuckets] > a.entries_cap) return;  // (grid-uniform) wide_scatter flagged every world: the fused kernel takes the step
    const int nlev = a.nlev[0];
    const unsigned lane = threadIdx.x & 31u, wic = threadIdx.x >> 5;
    const unsigned gw = wic * gridDim.x + blockIdx.x, nw = gridDim.x * (blockDim.x >> 5);
    for (int r = r_begin; r <= nlev; r++) {
        const unsigned first = a.bucket[r * 4];
        const unsigned c0 = a.bucket[r * 4 + 1] - first, c1 = a.bucket[r * 4 + 2] - a.bucket[r * 4 + 1], c2 = a.bucket[r * 4 + 3] - a.bucket[r * 4 + 2];
        if (c0 + c1 + c2 == 0) continue;
        {   // stage 1
            const unsigned ch0 = (c0 + 31) / 32, ch1 = (c1 + 31) / 32, ch2 = (c2 + 31) / 32;
            for (unsigned c = gw; c < ch0 + ch1 + ch2; c += nw) {
                if (c < ch0) wide_detect_body<0>(a, first, c0, r, c * 32 + lane);
                else if (c < ch0 + ch1) wide_detect_body<1>(a, first + c0, c1, r, (c - ch0) * 32 + lane);
                else wide_bb_quick_body(a, first + c0 + c1, c2, r, (c - ch0 - ch1) * 32 + lane);
            }
        }
        grid.sync();
        if (c2) {  // stage 2
            const unsigned ns = __ldcg(&a.surv_cnt[r]);
            const unsigned fbb = first + c0 + c1;
            if (ns <= a.team_max) {
                for (unsigned c = gw; c < (ns + 3) / 4; c += nw)
                    wide_bb_team_body(a, fbb, c2, r, c * 4 + lane / psc::kTeam, ns, ts_all[wic * 4 + lane / psc::kTeam]);
            } else {
                for (unsigned c = gw; c < (ns + 31) / 32; c += nw) wide_detect_body<2>(a, fbb, c2, r, c * 32 + lane);
            }
            grid.sync();
        }
        {   // stage 3
            unsigned nch[6], tot = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) { nch[k] = (__ldcg(&a.hit_cnt[r * kHitSlots + k]) + 31) / 32; tot += nch[k]; }
            const unsigned fbb = first + c0 + c1;
            for (unsigned c = gw; c < tot; c += nw) {  // the largest manifolds first
                unsigned b = c;
                int kind = 5;
                while (kind > 0 && b >= nch[kind]) { b -= nch[kind]; kind--; }
                const unsigned fk = kind >= 2 ? fbb : (kind == 1 ? first + c0 : first);
                unsigned ei;
                if (!wide_pick_hit(a, kind, fk, c2, r, b * 32 + lane, ei)) continue;
                if (kind >= 2) wide_resolve_entry<true>(a, ei, r);
                else wide_resolve_entry<false>(a, ei, r);
            }
        }
        grid.sync();
    }
}

// ---- V: bodies that left their margin (thread per body).  A culled pair (b, x) stays impossible if the spheres grown
// to the largest deviations still do not touch; only a pair that was culled and now touches sends the world to the
// fused kernel's wider-margin retry.
__global__ void __launch_bounds__(128) wide_verify(WArgs a)
{
    int g = a.g_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.g_end) return;
    const int w = a.body_world[g];
    const int o = a.off[w], N = a.cnt[w], b = g - o;
    if (b >= N) return;
    const double* cu = a.cull + (size_t)o * kCullF;
    const double d2 = cu[6 * N + b], m2 = cu[5 * N + b];
    if (d2 <= m2) return;
    if (a.wflags[w]) return;
    atomicAdd(&a.tmp[6 * (size_t)w + 5], 1u);
    if (!(d2 == d2)) { atomicOr(&a.wflags[w], (unsigned)WF_MARGIN); return; }
    const double gb = cull_growth(d2, m2);
    for (int x = 0; x < N; x++) {
        if (x == b) continue;
        if (wide_near_test(a, cu, o, N, b, x)) continue;  // on the near list: tested where it stood
        const double gx = cull_growth(cu[6 * N + x], cu[5 * N + x]);
        double dx = cu[b] - cu[x], dy = cu[N + b] - cu[N + x], dz = cu[2 * N + b] - cu[2 * N + x];
        double dd = dx * dx + dy * dy + dz * dz;
        double rb = (a.isbox[o + x] ? cu[4 * N + b] : cu[3 * N + b]) + gb;
        double rx = (a.isbox[o + b] ? cu[4 * N + x] : cu[3 * N + x]) + gx;
        double r = rb + rx;
        if (!(dd > r * r)) { atomicOr(&a.wflags[w], (unsigned)WF_MARGIN); return; }
    }
}

// ---- E: static bodies: L = fold over their colliding pairs in pair order (WARP per static body) --------------------------
// The warp reads the world's near list 32 entries at a time, a ballot picks the colliding pairs of this body in list order,
// and lane 0 adds their per-pair sums in that order (quirk Q2: a static body's angular momentum is a running sum).  One
// thread walking the list alone (first version) was 0.30 ms per C3 step for 16 384 grounds.
__global__ void __launch_bounds__(128) wide_static_fold(WArgs a, const int* __restrict__ stat_slot, int n_stat)
{
    const int sidx = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (sidx >= n_stat) return;
    const int g = stat_slot[sidx];
    const int w = a.body_world[g];
    if (a.wflags[w]) return;
    const int o = a.off[w], N = a.cnt[w], b = g - o, K = a.K[w];
    if (b >= N || !a.isstat[g]) return;
    Par p; ld_par(a.par, 0, g, p);
    double* cb = a.cur + (size_t)g * kDynF;
    double* pb = a.pred + (size_t)g * kDynF;
    d3 L = mk(cb[15], cb[16], cb[17]);
    const d3 dL = scale(a.sp.dt, p.T);
    const double* Sb = a.S + (size_t)o * 3 * kMaxStatics + 3 * (a.isstat[g] - 1) * N;
    const unsigned* near = w_near(a, w);
    const unsigned char* hit = w_hit(a, w);
    // The list is sorted by (i, j): nothing behind the rows 0 .. b involves body b.  One 32-way probe bounds the walk (the
    // ground is body 0 of its world: its pairs are the first ~N of ~3 N entries).
    int Kb = K;
    if (K > 64) {
        const int step = (K + 31) / 32, kp = lane * step;
        const bool behind = kp < K ? (int)(near[kp] & 0xffff) > b : true;
        const unsigned mb = __ballot_sync(0xffffffffu, behind);
        if (mb) Kb = min(K, (__ffs(mb) - 1) * step);
    }
    for (int k0 = 0; k0 < Kb; k0 += 32) {
        const int k = k0 + lane;
        int other = -1;
        if (k < Kb && hit[k]) {
            const unsigned ij = near[k];
            const int i = ij & 0xffff, j = ij >> 16;
            if (i == b) other = j; else if (j == b) other = i;
        }
        unsigned m = __ballot_sync(0xffffffffu, other >= 0);
        while (m) {  // in list order
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int ot = __shfl_sync(0xffffffffu, other, src);
            if (lane == 0) {
                L = L + dL;
                L = L + mk(Sb[3 * ot], Sb[3 * ot + 1], Sb[3 * ot + 2]);
            }
        }
    }
    if (lane == 0) {
        cb[15] = L.x; cb[16] = L.y; cb[17] = L.z;
        const d3 Lf = L + dL;
        pb[15] = Lf.x; pb[16] = Lf.y; pb[17] = Lf.z;
    }
}

// ---- F: publish (tile of 128 bodies per CTA) --------------------------------------------------------------------------
// On this path every body's predicted copy is current when the rounds are over (the solve re-integrates what it changes),
// so SimulatorState.update's integration (SimulatorState.fs:169-173) is that copy: the tile of `pred` comes in with one
// bulk copy, is checked for non-finite values from shared memory, and goes out to the published state with one bulk copy.
// Worlds the pass gave up on (wflags) keep their published state: their records are put back before the tile leaves.
__global__ void __launch_bounds__(kTile) wide_final(WArgs a)
{
    extern __shared__ __align__(128) unsigned char tile_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int any_flagged;
    double* s_t = reinterpret_cast<double*>(tile_raw);
    const int g0 = a.g_begin + blockIdx.x * kTile;
    const int n = min(kTile, a.g_end - g0);
    const unsigned bytes = (unsigned)n * kDynF * (unsigned)sizeof(double);
    if (threadIdx.x == 0) { pst_tma::mbar_init(&bar, 1); any_flagged = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        pst_tma::mbar_arrive_expect_tx(&bar, bytes);
        pst_tma::bulk_g2s(s_t, a.pred + (size_t)g0 * kDynF, bytes, &bar);
    }
    const int t = threadIdx.x, g = g0 + t;
    int w = 0;
    bool flagged = false;
    bool body = false;
    if (t < n) {
        w = a.body_world[g];
        body = g - a.off[w] < a.cnt[w];  // (a spare slot holds no body: its record travels with the tile, unread)
        flagged = a.wflags[w] != 0;  // the fused kernel redoes this world from the published state
        if (flagged) any_flagged = 1;
        else if (body) a.nprev[g] = a.nhit[g];
    }
    pst_tma::mbar_wait(&bar, 0);
    if (body && !flagged) {
        const double* q = s_t + (size_t)t * kDynF;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < kDynF; k++) bad = bad || !isfinite(q[k]);
        if (bad) a.stats[w].nan_flag = 1;
    }
    __syncthreads();
    if (any_flagged) {  // rare: keep the published records of flagged worlds
        if (t < n && flagged) {
            Dyn keep;
            ld_dyn(a.dyn, 0, g, keep);
            st_dyn(s_t, 0, t, keep);
        }
        pst_tma::fence_proxy_async();
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pst_tma::bulk_s2g(a.dyn + (size_t)g0 * kDynF, s_t, bytes);
        pst_tma::bulk_commit();
        pst_tma::bulk_wait_read();
    }
}

// account the worlds handed over to the fused kernel (thread per world)
__global__ void wide_account(WArgs a)
{
    int w = a.w_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.w_end) return;
    unsigned f = a.wflags[w];
    WorldStats& st = a.stats[w];
    const unsigned* tm = a.tmp + 6 * (size_t)w;
    if (!a.always_fused[w]) {  // doubled margins for a while after a body left its margin (the fused kernel keeps its own count)
        if (tm[5] || (f & WF_MARGIN)) st.boost = kBoostSteps;
        else if (!f && st.boost > 0) st.boost--;
    }
    if (!f) {  // the wide pass stands: commit its counters
        const unsigned long long pk = *reinterpret_cast<const unsigned long long*>(tm);
        st.steps++; st.tested += pk & 0xfffffull; st.hits += (pk >> 20) & 0xfffffull; st.contacts += pk >> 40;
        st.overflow += tm[3]; st.ref_exc += tm[4];
        if (tm[5]) st.verified++;
        st.max_pen = fmax(st.max_pen, __longlong_as_double((long long)a.tmp_pen[w]));
        return;
    }
    atomicAdd(a.nflag, 1);
    if (a.always_fused[w]) return;
    st.fallback++;
    if (f & WF_STATIC) st.fb_static++;
    else if (f & WF_MARGIN) st.fb_margin++;
    else st.fb_capacity++;
}

}  // namespace psw
