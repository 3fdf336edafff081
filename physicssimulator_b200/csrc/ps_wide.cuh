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

namespace psw {
using namespace psk;

constexpr int kWideMaxLevels = 1022;
constexpr int kBuckets = (kWideMaxLevels + 2) * 4;

struct __align__(16) Entry {  // one near pair of the batch, sorted by (level, kind); written and read as one uint4
    int w;           // world
    unsigned ij;     // i | j<<16 (ids inside the world)
    unsigned knc;    // index in the world's near list (for the hit flag, 24 bits) | contacts found by the narrow phase << 24
    unsigned cfirst; // first contact in the pool
};

struct WArgs {
    int n_worlds, n_bodies, max_n;
    const int* off;
    const int* body_world;       // [NB]
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
    unsigned* hit_cnt;           // per (level, kind) bucket
    unsigned* hitlist;           // [entries_cap], a bucket's hits are listed from the bucket's start
    Contact* cpool; unsigned cpool_cap; unsigned* cpool_cur;  // cursor per level
    unsigned* surv;              // [entries_cap] box-box pairs the quick test could not separate (listed from the bucket's start)
    unsigned* surv_cnt;          // per level
    WorldStats* stats;
    unsigned* tmp;               // per world x 6 words of THIS step's wide pass: [0-1] tested | colliding | contacts packed in 64 bits (count_pair), [3] overflow, [4] ref_exc, [5] bodies that left their margin
    unsigned long long* tmp_pen; // per world: max |penetration| bits of this step
    RecPair* rec_pairs; RecContact* rec_contacts; int* rec_counts; int rec_pair_cap, rec_contact_cap;
    StepParams sp;
    int margin_ncap;
    double margin_scale;         // safety factor on the motion margin (a violated margin costs a whole fused redo)
};

enum { WF_MARGIN = 1, WF_STATIC = 2, WF_CAPACITY = 4 };
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
PSD size_t w_np(const WArgs& a, int w) { size_t N = (size_t)(a.off[w + 1] - a.off[w]); return N * (N > 0 ? N - 1 : 0) / 2; }
PSD unsigned short* w_level(const WArgs& a, int w) { return reinterpret_cast<unsigned short*>(a.lists + a.lists_off[w] + 4 * w_np(a, w)); }
PSD unsigned char* w_hit(const WArgs& a, int w) { return a.lists + a.lists_off[w] + 8 * w_np(a, w); }

// ---- per step reset ------------------------------------------------------------------------------------
__global__ void wide_reset(WArgs a)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= kBuckets) a.bucket[t] = 0;
    if (t < kBuckets) a.bucket_cur[t] = 0;
    if (t < kBuckets) a.hit_cnt[t] = 0;
    if (t < kWideMaxLevels + 2) { a.cpool_cur[t] = 0; a.surv_cnt[t] = 0; }
    if (t < a.n_worlds) {
        a.wflags[t] = a.always_fused[t] ? WF_STATIC : 0u;
        a.K[t] = 0; a.maxl[t] = 0;
        a.stats[t].nan_flag = 0;
        for (int k = 0; k < 6; k++) a.tmp[6 * t + k] = 0;
        a.tmp_pen[t] = 0ull;
        if (a.rec_counts) { a.rec_counts[2 * t] = 0; a.rec_counts[2 * t + 1] = 0; }
    }
}

// ---- A: predicted copies + cull spheres (thread per body) -------------------------------------------------
__global__ void __launch_bounds__(128) wide_predict(WArgs a)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_bodies) return;
    const int w = a.body_world[g], o = a.off[w], N = a.off[w + 1] - o, b = g - o;
    Par p; Dyn cur, pr;
    ld_par(a.par, 0, g, p);
    ld_dyn(a.dyn, 0, g, cur);   // the published state is the step's input
    st_dyn(a.cur, 0, g, cur);
    integrate(cur, p, a.sp, pr);
    st_dyn(a.pred, 0, g, pr);
    a.pvalid[g] = 1; a.nhit[g] = 0;
    if (p.is_static && !pose_fixed(cur, pr)) atomicOr(&a.wflags[w], (unsigned)WF_STATIC);
    double* cu = a.cull + (size_t)o * kCullF;
    cu[b] = pr.x.x; cu[N + b] = pr.x.y; cu[2 * N + b] = pr.x.z;
    double speed = sqrt(pr.v.x * pr.v.x + pr.v.y * pr.v.y + pr.v.z * pr.v.z);
    double acc = p.inv_mass * sqrt(p.F.x * p.F.x + p.F.y * p.F.y + p.F.z * p.F.z);
    int np_ = a.nprev[g];
    double n = (double)(np_ < a.margin_ncap ? np_ : a.margin_ncap) + 2.0;
    const double mscale = a.margin_scale * (a.stats[w].boost > 0 ? 2.0 : 1.0);
    double m = p.is_static ? 0.0 : mscale * (0.01 + n * a.sp.dt * (speed + n * acc * a.sp.dt + 0.5));
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
__global__ void __launch_bounds__(128) wide_near_count(WArgs a)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_bodies) return;
    const int w = a.body_world[g], o = a.off[w], N = a.off[w + 1] - o, i = g - o;
    const double* cu = a.cull + (size_t)o * kCullF;
    int cn = 0;
    for (int j = i + 1; j < N; j++) cn += wide_near_test(a, cu, o, N, i, j) ? 1 : 0;
    a.row[o + w + i] = cn;
}
__global__ void wide_near_scan(WArgs a)  // thread per world
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.n_worlds) return;
    const int o = a.off[w], N = a.off[w + 1] - o;
    int* row = a.row + o + w;
    int acc = 0;
    for (int i = 0; i < N; i++) { int t = row[i]; row[i] = acc; acc += t; }
    row[N] = acc;
    a.K[w] = acc;
}
__global__ void __launch_bounds__(128) wide_near_fill(WArgs a)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_bodies) return;
    const int w = a.body_world[g], o = a.off[w], N = a.off[w + 1] - o, i = g - o;
    const double* cu = a.cull + (size_t)o * kCullF;
    unsigned* near = w_near(a, w);
    unsigned char* hit = w_hit(a, w);
    int q = a.row[o + w + i];
    for (int j = i + 1; j < N; j++)
        if (wide_near_test(a, cu, o, N, i, j)) {
            near[q] = (unsigned)i | ((unsigned)j << 16);
            hit[q] = 0;
            q++;
        }
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
    const int w0 = blockIdx.x * kLvlWorlds, wend = min(w0 + kLvlWorlds, a.n_worlds);
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
    for (int b = lane * per; b < (lane + 1) * per; b++) { unsigned t = a.bucket[b]; a.bucket[b] = base; base += t; }
    if (lane == 31) a.bucket[kBuckets] = base;
}
__global__ void __launch_bounds__(kLvlThreads) wide_scatter(WArgs a)  // same CTA -> worlds map as wide_levels
{
    __shared__ unsigned cnt[kBuckets];   // pairs of this CTA per bucket, then the CTA's base inside the bucket
    __shared__ unsigned cur[kBuckets];
    for (int b = threadIdx.x; b < kBuckets; b += kLvlThreads) { cnt[b] = 0; cur[b] = 0; }
    const int w0 = blockIdx.x * kLvlWorlds, wend = min(w0 + kLvlWorlds, a.n_worlds);
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
    e.knc |= (unsigned)nc << 24; e.cfirst = cf;
    unsigned slot = agg_reserve(&a.hit_cnt[level * 4 + KIND], 1u);
    a.hitlist[first + slot] = ei;
    double mp = 0.0;
    for (int k = 0; k < nc; k++) mp = fmax(mp, fabs(cs[k].pen));
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
                RecContact r;
                r.n[0] = cs[k].n.x; r.n[1] = cs[k].n.y; r.n[2] = cs[k].n.z;
                r.p[0] = cs[k].p.x; r.p[1] = cs[k].p.y; r.p[2] = cs[k].p.z;
                r.pen = cs[k].pen; r.feat = cs[k].feat; r.pad = 0;
                out[k] = r;
            }
        }
    }
}

// solve + write-back: thread per colliding pair of bucket (level, kind).  A body the solve changed is integrated
// again right here (its next use — the next pair or the final sweep — needs exactly this predicted copy), which
// keeps that work in a dense kernel; its margin only matters if a later near pair will test it.
template <int KIND>
__device__ __forceinline__ void wide_resolve_body(const WArgs& a, unsigned first, int level, unsigned t)
{
    PS_CLK_BEGIN(KIND == 2 && level >= 3)
    if (t >= a.hit_cnt[level * 4 + KIND]) return;
    const Entry& e = a.entries[a.hitlist[first + t]];
    const int w = e.w;
    if (a.wflags[w]) return;
    const int o = a.off[w], N = a.off[w + 1] - o;
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
    if (KIND == 2) {
        const int nc = (int)(e.knc >> 24);
        CPData cp[PS_MAX_CONTACTS];
        PS_CLK(13)
        resolve_t<0>(di, pi, dj, pj, a.cpool + e.cfirst, nc, cp, a.sp);  // the manifold is read where the narrow phase left it
        PS_CLK(14)
    } else {  // sphere pairs always carry exactly one contact
        Contact cs[1];
        CPData cp[1];
        cs[0] = a.cpool[e.cfirst];
        resolve_t<1>(di, pi, dj, pj, cs, 1, cp, a.sp);
    }
    w_hit(a, w)[e.knc & 0xffffffu] = 1;
    double* Sw = a.S + (size_t)o * 3 * kMaxStatics;
    double* cu = a.cull + (size_t)o * kCullF;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
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
    if (blk < b0) wide_resolve_body<0>(a, first, level, blk * 128 + threadIdx.x);
    else if (blk < b0 + b1) wide_resolve_body<1>(a, first + c0, level, (blk - b0) * 128 + threadIdx.x);
    else wide_resolve_body<2>(a, first + c0 + c1, level, (blk - b0 - b1) * 128 + threadIdx.x);
}

// ---- V: bodies that left their margin (thread per body).  A culled pair (b, x) stays impossible if the spheres grown
// to the largest deviations still do not touch; only a pair that was culled and now touches sends the world to the
// fused kernel's wider-margin retry.
__global__ void __launch_bounds__(128) wide_verify(WArgs a)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_bodies) return;
    const int w = a.body_world[g];
    const int o = a.off[w], N = a.off[w + 1] - o, b = g - o;
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

// ---- E: static bodies: L = fold over their colliding pairs in pair order (thread per body, statics only) -----
__global__ void __launch_bounds__(128) wide_static_fold(WArgs a)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_bodies) return;
    if (!a.isstat[g]) return;
    const int w = a.body_world[g];
    if (a.wflags[w]) return;
    const int o = a.off[w], N = a.off[w + 1] - o, b = g - o, K = a.K[w];
    Par p; ld_par(a.par, 0, g, p);
    double* cb = a.cur + (size_t)g * kDynF;
    double* pb = a.pred + (size_t)g * kDynF;
    d3 L = mk(cb[15], cb[16], cb[17]);
    d3 dL = scale(a.sp.dt, p.T);
    const double* Sb = a.S + (size_t)o * 3 * kMaxStatics + 3 * (a.isstat[g] - 1) * N;
    const unsigned* near = w_near(a, w);
    const unsigned char* hit = w_hit(a, w);
    for (int k = 0; k < K; k++) {
        if (!hit[k]) continue;
        unsigned ij = near[k];
        int i = ij & 0xffff, j = ij >> 16;
        int other;
        if (i == b) other = j; else if (j == b) other = i; else continue;
        L = L + dL;
        L = L + mk(Sb[3 * other], Sb[3 * other + 1], Sb[3 * other + 2]);
    }
    cb[15] = L.x; cb[16] = L.y; cb[17] = L.z;
    d3 Lf = L + dL;
    pb[15] = Lf.x; pb[16] = Lf.y; pb[17] = Lf.z;
}

// ---- F: final integration + publish (thread per body) ---------------------------------------------------------
__global__ void __launch_bounds__(128) wide_final(WArgs a)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_bodies) return;
    const int w = a.body_world[g];
    if (a.wflags[w]) return;  // the fused kernel redoes this world from the published state
    Dyn out;
    if (a.pvalid[g]) {
        ld_dyn(a.pred, 0, g, out);
    } else {
        Par p; ld_par(a.par, 0, g, p);
        Dyn cur; ld_dyn(a.cur, 0, g, cur);
        integrate(cur, p, a.sp, out);
    }
    st_dyn(a.dyn, 0, g, out);
    a.nprev[g] = a.nhit[g];
    bool bad = !(isfinite(out.x.x) && isfinite(out.x.y) && isfinite(out.x.z) && isfinite(out.v.x) && isfinite(out.v.y) && isfinite(out.v.z) &&
                 isfinite(out.L.x) && isfinite(out.L.y) && isfinite(out.L.z));
    for (int k = 0; k < 9; k++) bad = bad || !isfinite(out.R.a[k]);
    if (bad) a.stats[w].nan_flag = 1;

}

// account the worlds handed over to the fused kernel (thread per world)
__global__ void wide_account(WArgs a)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.n_worlds) return;
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
    if (a.always_fused[w]) return;
    st.fallback++;
    if (f & WF_STATIC) st.fb_static++;
    else if (f & WF_MARGIN) st.fb_margin++;
    else st.fb_capacity++;
}

}  // namespace psw
