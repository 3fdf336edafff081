// ps_flow.cuh — the dependency-driven form of the step ("flow"): the batch-wide path without its rounds.
//
// The batch-wide path (ps_wide.cuh) runs a wavefront level as a round of the whole batch and every round waits for
// its slowest pair — ncu shows the SMs busy for 28 K of the 184 K cycles of a late resolve launch, the rest is a
// handful of 8-contact manifolds.  But a near pair only has to wait for the PREVIOUS near pair of each of its two
// dynamic bodies (that is all the ordered fold of CollisionResolver.fs:66-68 makes it depend on, given the cull).
// So: every near pair of the batch gets a record with a count of unfinished predecessors (0, 1 or 2) and the ids of
// its two successors; pairs with count 0 are queued; ONE persistent kernel drains the queues — a warp claims up to
// 32 tasks of one type, runs them, and a finished pair releases its successors, which are queued when their count
// reaches 0.  Task types keep the lanes of a warp on one code path: narrow phase per pair kind (sphere-sphere,
// sphere-box, box-box quick test, box-box full), solve with one contact, solve of a manifold.  No level tables, no
// sort, no launches per round; a straggler only delays the pairs that really depend on it.
//
// Queues: one array of pair ids per task type, a `tail` the producers reserve from and a `head` the consumers claim
// from (compare-and-swap, never beyond the reserved tail, so a claimed slot is at worst being written).  A slot holds
// kEmpty until its producer stores the id; the consumer puts kEmpty back, so the arrays are clean again when the step
// is drained.  State a task reads was written by a task on another SM: producers fence before they release, consumers
// read the state past L1 (ld.global.cg).  A watchdog bounds the kernel: on a timeout every world is flagged and the
// fused kernel redoes the step.
//
// Exactness is the batch-wide path's: same per-pair routines, same write-back, same margin bookkeeping (wide_verify),
// same static fold; only the schedule differs.
#pragma once
#include "ps_wide.cuh"

namespace psf {
using namespace psw;

constexpr int kQ = 6;  // 0 detect sphere-sphere, 1 detect sphere-box, 2 box-box quick test, 3 box-box full, 4 solve 1 contact, 5 solve manifold
constexpr unsigned kEmpty = 0xffffffffu;
constexpr int kQStride = 32;  // head/tail counters on their own 128-byte lines
constexpr int kFlowThreads = 128;

struct FArgs {
    WArgs w;            // buffers of the batch-wide path; w.entries[g] is the record of near pair g
    unsigned* pbase;    // [n_worlds+1] first pair id of every world (exclusive prefix of K; 0 pairs for flagged worlds)
    int2* link;         // [cap] the next near pair of body i / body j of this pair (pair id, -1 = none)
    unsigned* dep;      // [cap] unfinished predecessors
    unsigned* q;        // [kQ][cap] queued pair ids
    unsigned* qhead;    // [kQ * kQStride]
    unsigned* qtail;    // [kQ * kQStride]
    unsigned* ctl;      // [0] pairs completed, [1] abort, [2] pairs of this step
    unsigned cap;
    long long watchdog;  // SM cycles
};

// ---- loads past L1 (state written by another SM inside the same kernel) -----------------------------------------------
__device__ __forceinline__ void ld_dyn_cg(const double* a, size_t b, Dyn& d)
{
    const double2* q = reinterpret_cast<const double2*>(a + b * kDynF);
    double2 t0 = __ldcg(q), t1 = __ldcg(q + 1), t2 = __ldcg(q + 2), t3 = __ldcg(q + 3), t4 = __ldcg(q + 4), t5 = __ldcg(q + 5),
            t6 = __ldcg(q + 6), t7 = __ldcg(q + 7), t8 = __ldcg(q + 8);
    d.x = mk(t0.x, t0.y, t1.x);
    d.v = mk(t1.y, t2.x, t2.y);
    d.R.a[0] = t3.x; d.R.a[1] = t3.y; d.R.a[2] = t4.x; d.R.a[3] = t4.y; d.R.a[4] = t5.x; d.R.a[5] = t5.y;
    d.R.a[6] = t6.x; d.R.a[7] = t6.y; d.R.a[8] = t7.x;
    d.L = mk(t7.y, t8.x, t8.y);
}
__device__ __forceinline__ Entry ld_entry_cg(const Entry* e)
{
    uint4 t = __ldcg(reinterpret_cast<const uint4*>(e));
    Entry r; r.w = (int)t.x; r.ij = t.y; r.knc = t.z; r.cfirst = t.w;
    return r;
}

// ---- queues -------------------------------------------------------------------------------------------------------------
// append g to queue qi: the lanes of the warp that arrive together share one reservation
__device__ __forceinline__ void q_push(const FArgs& f, int qi, unsigned g)
{
    const unsigned slot = agg_reserve(f.qtail + qi * kQStride, 1u);
    if (slot >= f.cap) { atomicExch(&f.ctl[1], 1u); return; }  // cannot happen (a pair enters a queue once); abort rather than corrupt
    __stcg(f.q + (size_t)qi * f.cap + slot, g);
}
// every lane may push to a different queue: one converged group per queue
__device__ __forceinline__ void q_push_any(const FArgs& f, int qi, unsigned g)
{
#pragma unroll 1
    for (int k = 0; k < kQ; k++)
        if (qi == k) q_push(f, k, g);
}
__device__ __forceinline__ int detect_queue(const WArgs& a, int w, unsigned ij)
{
    const int o = a.off[w];
    return pair_kind(a, o, ij & 0xffff, ij >> 16);  // 0, 1, 2 = the three detection queues
}
// pair g is finished (its write-back, if any, is done): release the next near pair of each of its bodies
__device__ __forceinline__ void complete(const FArgs& f, unsigned g)
{
    __threadfence();  // the state this pair wrote is visible before a successor can start
    const int2 nx = __ldcg(f.link + g);
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const int s = side ? nx.y : nx.x;
        if (s < 0) continue;
        if (atomicSub(&f.dep[s], 1u) == 1u) {
            const Entry e = ld_entry_cg(f.w.entries + s);
            q_push_any(f, detect_queue(f.w, e.w, e.ij), (unsigned)s);
        }
    }
    const unsigned one = agg_reserve(&f.ctl[0], 1u);
    (void)one;
}

// ---- setup ---------------------------------------------------------------------------------------------------------------
// first pair id of every world; one CTA
__global__ void __launch_bounds__(1024) flow_scan(FArgs f)
{
    __shared__ unsigned part[1024];
    const WArgs& a = f.w;
    const int t = threadIdx.x, per = (a.n_worlds + 1023) / 1024;
    const int w0 = t * per, w1 = min(w0 + per, a.n_worlds);
    unsigned s = 0;
    for (int w = w0; w < w1; w++) s += a.wflags[w] ? 0u : (unsigned)a.K[w];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan of the partial sums
        unsigned v = (t >= o) ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned base = t ? part[t - 1] : 0u;
    for (int w = w0; w < w1; w++) {
        f.pbase[w] = base;
        base += a.wflags[w] ? 0u : (unsigned)a.K[w];
    }
    if (t == 1023) {
        const unsigned total = part[1023];
        f.pbase[a.n_worlds] = total;
        f.ctl[0] = 0u; f.ctl[1] = 0u;
        f.ctl[2] = total > f.cap ? 0u : total;  // too many pairs: flow_build flags every world instead
    }
    if (t < kQ) { f.qhead[t * kQStride] = 0u; f.qtail[t * kQStride] = 0u; }
}

// records, predecessor counts and successor links of every near pair; pairs without predecessors are queued.
// A CTA owns kLvlWorlds worlds, one thread walks a world's ordered list (the previous near pair of every body sits in
// shared memory), like wide_levels.
__host__ __device__ inline size_t flow_smem_bytes(int max_n) { return (size_t)kLvlWorlds * (size_t)(max_n + 1) * sizeof(unsigned); }
__global__ void __launch_bounds__(32) flow_build(FArgs f)
{
    extern __shared__ __align__(16) unsigned char fb_raw[];
    unsigned* lastS = reinterpret_cast<unsigned*>(fb_raw);  // per body: (previous near pair's id << 1 | side) + 1, 0 = none
    const WArgs& a = f.w;
    const int stride = a.max_n + 1;
    const int w = blockIdx.x * kLvlWorlds + (int)threadIdx.x;
    const bool too_many = f.pbase[a.n_worlds] > f.cap;
    const bool live = w < a.n_worlds && !a.wflags[w];
    if (live && too_many) atomicOr(&a.wflags[w], (unsigned)WF_CAPACITY);
    const bool on = live && !too_many;
    const int o = on ? a.off[w] : 0, N = on ? a.off[w + 1] - o : 0, K = on ? a.K[w] : 0;
    unsigned* L = lastS + threadIdx.x * stride;
    for (int b = 0; b < N; b++) L[b] = 0u;
    const unsigned* near = on ? w_near(a, w) : nullptr;
    const unsigned base = on ? f.pbase[w] : 0u;
    for (int k = 0; k < K; k++) {
        const unsigned ij = near[k];
        const int i = ij & 0xffff, j = ij >> 16;
        const unsigned g = base + (unsigned)k;
        unsigned deps = 0;
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
            const int b = side ? j : i;
            if (a.isstat[o + b]) continue;  // a static body does not order its pairs
            const unsigned prev = L[b];
            if (prev) {
                const unsigned pg = (prev - 1u) >> 1, pside = (prev - 1u) & 1u;
                int* slot = reinterpret_cast<int*>(f.link + pg) + pside;
                *slot = (int)g;
                deps++;
            }
            L[b] = ((g << 1) | (unsigned)side) + 1u;
        }
        *reinterpret_cast<uint4*>(&a.entries[g]) = make_uint4((unsigned)w, ij, (unsigned)k, 0u);
        f.link[g] = make_int2(-1, -1);
        f.dep[g] = deps;
    }
    // queue the pairs that can start right away (warp-uniform trip count: q_push is a warp-level reservation)
    int Kw = K;
#pragma unroll
    for (int s = 16; s; s >>= 1) Kw = max(Kw, __shfl_xor_sync(0xffffffffu, Kw, s));
    for (int k = 0; k < Kw; k++) {
        if (k < K && f.dep[base + (unsigned)k] == 0u) q_push_any(f, detect_queue(a, w, near[k]), base + (unsigned)k);
    }
}

// ---- tasks (one lane = one pair) ------------------------------------------------------------------------------------------
struct PairIn {
    int w, o, N, i, j;
    Par pi, pj;
    Dyn di, dj;
};
__device__ __forceinline__ void load_pair(const WArgs& a, const Entry& e, PairIn& p)
{
    p.w = e.w;
    p.o = a.off[p.w];
    p.N = a.off[p.w + 1] - p.o;
    p.i = e.ij & 0xffff; p.j = e.ij >> 16;
    ld_par(a.par, 0, p.o + p.i, p.pi);
    ld_par(a.par, 0, p.o + p.j, p.pj);
    ld_dyn_cg(a.pred, (size_t)(p.o + p.i), p.di);
    ld_dyn_cg(a.pred, (size_t)(p.o + p.j), p.dj);
}
__device__ __forceinline__ void record_hit(const WArgs& a, const PairIn& p, const Contact* cs, int nc)
{
    double mp = 0.0;
    for (int k = 0; k < nc; k++) mp = fmax(mp, fabs(cs[k].pen));
    const unsigned long long mpb = (unsigned long long)__double_as_longlong(mp);
    if (mpb > __ldcg(&a.tmp_pen[p.w])) atomicMax(&a.tmp_pen[p.w], mpb);
    if (a.rec_pairs) {
        int* cnt = a.rec_counts + 2 * p.w;
        int slot2 = atomicAdd(&cnt[0], 1);
        int f2 = atomicAdd(&cnt[1], nc);
        if (slot2 < a.rec_pair_cap && f2 + nc <= a.rec_contact_cap) {
            RecPair rp; rp.i = p.i; rp.j = p.j; rp.first = f2; rp.count = nc;
            a.rec_pairs[(size_t)p.w * a.rec_pair_cap + slot2] = rp;
            RecContact* out = a.rec_contacts + (size_t)p.w * a.rec_contact_cap + f2;
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
// narrow phase of pair g; KIND 0 sphere-sphere, 1 sphere-box, 2 box-box (the quick test has failed to separate)
template <int KIND>
__device__ __forceinline__ void task_detect(const FArgs& f, unsigned g)
{
    const WArgs& a = f.w;
    const Entry e = ld_entry_cg(a.entries + g);
    PairIn p;
    load_pair(a, e, p);
    PairScratch scb;
    Contact cs1[1];
    Contact* const cs = (KIND == 2) ? scb.cs : cs1;
    int err = 0, ovf = 0, nc;
    if (KIND == 0) nc = detect_ss_t<true>(p.di, p.pi, p.dj, p.pj, cs);
    else if (KIND == 2) nc = detect_bb_t<true>(p.di, p.pi, p.dj, p.pj, a.sp.eps, cs, err, ovf, scb, true);
    else { OBox ob; nc = detect_sphere_pair_t<true>(p.di, p.pi, p.dj, p.pj, a.sp.eps, cs, err, ob); }
    unsigned* tm = a.tmp + 6 * (size_t)p.w;
    if (err) atomicAdd(&tm[4], 1u);
    if (ovf) atomicAdd(&tm[3], 1u);
    count_pair(a, p.w, nc > 0 ? 1u : 0u, nc > 0 ? (unsigned)nc : 0u);
    if (nc <= 0) { complete(f, g); return; }
    const unsigned cf = agg_reserve(&a.cpool_cur[0], (unsigned)nc);
    if (cf + (unsigned)nc > a.cpool_cap) { atomicOr(&a.wflags[p.w], (unsigned)WF_CAPACITY); complete(f, g); return; }
    for (int k = 0; k < nc; k++) a.cpool[cf + k] = cs[k];
    *reinterpret_cast<uint4*>(&a.entries[g]) = make_uint4((unsigned)e.w, e.ij, (e.knc & 0xffffffu) | ((unsigned)nc << 24), cf);
    record_hit(a, p, cs, nc);
    __threadfence();  // contacts and record before the task is visible
    q_push(f, KIND == 2 ? 5 : 4, g);
}
__device__ __forceinline__ void task_bb_quick(const FArgs& f, unsigned g)
{
    const WArgs& a = f.w;
    const Entry e = ld_entry_cg(a.entries + g);
    PairIn p;
    load_pair(a, e, p);
    OBox o1, o2;
    make_obox_i(p.di, p.pi, o1);
    make_obox_i(p.dj, p.pj, o2);
    if (bb_quick_separated_t<true>(p.di, p.dj, o1, o2)) {
        count_pair(a, p.w, 0u, 0u);
        complete(f, g);
        return;
    }
    q_push(f, 3, g);
}
// solve + write-back + the next prediction of both bodies (as wide_resolve_body), then release the successors
template <int KIND>
__device__ __forceinline__ void task_resolve(const FArgs& f, unsigned g)
{
    const WArgs& a = f.w;
    const Entry e = ld_entry_cg(a.entries + g);
    PairIn p;
    load_pair(a, e, p);
    const bool si = p.pi.is_static, sj = p.pj.is_static;
    if (si) p.di.L = mk(0.0, 0.0, 0.0);
    if (sj) p.dj.L = mk(0.0, 0.0, 0.0);
    if (KIND == 2) {
        CPData cp[PS_MAX_CONTACTS];
        resolve_t<0>(p.di, p.pi, p.dj, p.pj, a.cpool + e.cfirst, (int)(e.knc >> 24), cp, a.sp);
    } else {
        Contact cs[1];
        CPData cp[1];
        cs[0] = a.cpool[e.cfirst];
        resolve_t<1>(p.di, p.pi, p.dj, p.pj, cs, 1, cp, a.sp);
    }
    w_hit(a, p.w)[e.knc & 0xffffffu] = 1;
    double* Sw = a.S + (size_t)p.o * 3 * kMaxStatics;
    double* cu = a.cull + (size_t)p.o * kCullF;
    const int N = p.N;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const int b = side ? p.j : p.i, other = side ? p.i : p.j;
        const bool stat = side ? sj : si;
        Dyn& d = side ? p.dj : p.di;
        const Par& pp = side ? p.pj : p.pi;
        if (stat) {
            double* S = Sw + 3 * ((a.isstat[p.o + b] - 1) * N + other);
            S[0] = d.L.x; S[1] = d.L.y; S[2] = d.L.z;
            continue;
        }
        st_dyn(a.cur, 0, p.o + b, d);  // write back (CollisionResolver.fs:27-31)
        {
            unsigned char* nh = a.nhit + p.o + b;
            const unsigned v = __ldcg(nh);
            if (v < 255u) *nh = (unsigned char)(v + 1u);
        }
        Dyn pr;
        integrate(d, pp, a.sp, pr);
        st_dyn(a.pred, 0, p.o + b, pr);
        d3 dd = pr.x - mk(cu[b], cu[N + b], cu[2 * N + b]);
        double q = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
        if (!(q <= cu[5 * N + b]) && !(q <= __ldcg(&cu[6 * N + b]))) cu[6 * N + b] = q;  // a NaN sticks
    }
    complete(f, g);
}

// ---- the executor: persistent warps drain the queues -----------------------------------------------------------------------
__global__ void __launch_bounds__(kFlowThreads, 2) flow_exec(FArgs f)
{
    const int lane = threadIdx.x & 31;
    const long long t_start = clock64();
    const unsigned total = __ldcg(&f.ctl[2]);
    volatile unsigned* ctl = f.ctl;
    volatile unsigned* qhead = f.qhead;
    volatile unsigned* qtail = f.qtail;
    unsigned idle = 0;
    for (;;) {
        // lane 0 looks for work: solves first (they are what successors wait for), then the narrow phase; a queue that
        // fills a warp wins over a fuller-but-later one, otherwise whatever is there
        unsigned code = 0, h = 0, n = 0;
        int qi = -1;
        if (lane == 0) {
            if (ctl[0] >= total || ctl[1]) {
                code = 2;  // drained (or aborted)
            } else {
                unsigned best_avail = 0, best_h = 0;
                const int order[kQ] = {5, 4, 3, 2, 1, 0};
#pragma unroll
                for (int r = 0; r < kQ; r++) {
                    const int k = order[r];
                    const unsigned hh = qhead[k * kQStride], tt = qtail[k * kQStride];
                    const unsigned avail = tt > hh ? tt - hh : 0u;
                    if (avail >= 32u && best_avail < 32u) { qi = k; best_avail = avail; best_h = hh; }
                    else if (avail > best_avail && best_avail < 32u) { qi = k; best_avail = avail; best_h = hh; }
                }
                if (qi >= 0 && best_avail > 0) {
                    n = best_avail < 32u ? best_avail : 32u;
                    if (atomicCAS(f.qhead + qi * kQStride, best_h, best_h + n) == best_h) { h = best_h; code = 1; }
                }
                if (code == 0 && clock64() - t_start > f.watchdog) { atomicExch(f.ctl + 1, 1u); code = 2; }
            }
        }
        code = __shfl_sync(0xffffffffu, code, 0);
        if (code == 2) break;
        if (code == 0) {
            idle++;
            __nanosleep(idle < 8 ? 100 : 400);
            continue;
        }
        idle = 0;
        qi = __shfl_sync(0xffffffffu, qi, 0);
        h = __shfl_sync(0xffffffffu, h, 0);
        n = __shfl_sync(0xffffffffu, n, 0);
        unsigned g = kEmpty;
        if ((unsigned)lane < n) {
            volatile unsigned* slot = f.q + (size_t)qi * f.cap + h + lane;
            while ((g = *slot) == kEmpty) {}  // reserved but not yet stored: its producer is between the two
            *slot = kEmpty;                   // the arrays are clean again when the step is drained
        }
        __syncwarp();
        __threadfence();  // what the producer wrote before it queued the task
        if ((unsigned)lane < n) {
            switch (qi) {  // warp-uniform
            case 0: task_detect<0>(f, g); break;
            case 1: task_detect<1>(f, g); break;
            case 2: task_bb_quick(f, g); break;
            case 3: task_detect<2>(f, g); break;
            case 4: task_resolve<1>(f, g); break;
            default: task_resolve<2>(f, g); break;
            }
        }
        __syncwarp();
    }
}

// the executor gave up (watchdog): every world goes to the fused kernel, the queues are wiped
__global__ void flow_abort_flag(FArgs f)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < f.w.n_worlds) atomicOr(&f.w.wflags[w], (unsigned)WF_CAPACITY);
}

}  // namespace psf
