// ps_step_kernel.cuh — the fused per-world step kernel (one CTA per world).
//
// A TILE of TS lanes (8, 16 or 32) owns one world for n_steps steps; a one-warp CTA hosts 32/TS worlds.
// A wavefront level of a 128-body pile holds ~10 pairs, so a whole warp per world would leave two thirds of
// the lanes idle; with several worlds per warp the tiles walk their levels in lockstep (same loop, same
// program counter) and the lanes of one warp instruction work on pairs of DIFFERENT worlds.  Tiles only ever
// synchronise among themselves (__syncwarp with the tile mask).
//
// Memory: the bodies live in a per-world scratch record in global memory (body-major, 144 B of state +
// 128 B of parameters per body, read with 16-byte loads; the working set of the resident tiles sits in
// L2/L1), and so do the per-step lists (cull spheres, near list, levels, order).  Shared memory only holds
// ~2 KB of flags and level offsets per world, so that residency is bounded by registers, not by shared
// memory: the step is a chain of dependent FP64 operations per pair and needs warps, not bandwidth.  The
// state is published to the batch buffer once per step, which doubles as the restart point of the fallback.
//
// Per step (reference order: Simulator.fs:52-56 = resolveAll (CollisionResolver.fs:49-69) then
// SimulatorState.update (SimulatorState.fs:169-173)):
//   A  per body : predicted copy = integrate(body) (CollisionResolver.fs:42-43); cull sphere =
//                 bounding radius + motion margin around the PREDICTED position
//   B  per row  : conservative cull of the Dummy-order pair list (BroadPhase.fs:33-39) -> ordered
//                 near list in shared memory (row scan keeps the lexicographic order)
//   C  thread 0 : wavefront level of every near pair (a pair waits for the previous near pair of
//                 either of its dynamic bodies), then a counting sort by level
//   D  per level: one thread per pair: (re-integrate a body an earlier pair changed) -> narrow
//                 phase -> N-iteration solve -> write back (CollisionResolver.fs:13-38)
//   E  per static body: fold the angular-momentum sums of its pairs in pair order
//   J  joints (off by default, Simulator.fs:55)
//   F  per body : final integration (a body no pair changed reuses its predicted copy — same value)
//
// Exactness: pairs in one level touch disjoint dynamic bodies, so the fold over the ordered pair
// list (CollisionResolver.fs:66-68) commutes inside a level.  A culled pair can never collide as
// long as every position used for detection stays within the margin of the sphere centre; that is
// checked on every re-integration.  A violated margin restarts the step from the published state
// with an 8x margin; a full list, too many levels or a static body whose pose is not a fixed point
// of integrate() sends the step to the exact serial walk over ALL pairs.
// Static bodies do not serialise their pairs when their pose is a fixed point of integrate()
// (checked every step); their angular momentum (Q2) is then summed per pair and folded in pair
// order — equal to the reference up to floating-point association; the serial
// walk reproduces it bit for bit.
#pragma once
#include "ps_physics.cuh"

#ifndef PS_FUSED_STRAIGHT_SOLVE
#define PS_FUSED_STRAIGHT_SOLVE 0  // 1: sphere pairs of the fused kernel take the straight-line solve (resolve_s).  Measured, ms per step: C2 4096 worlds
                                   // 1.70 against 1.51, C4 7.89 against 7.76, C3 2048 worlds 8.68 against 8.49 — two more inlined copies of the solver in a
                                   // kernel that lives on its instruction cache; the batch-wide kernels, one routine per launch, do profit from it
#endif
namespace psk {
using namespace psp;

constexpr int kDynF = 18;       // doubles of dynamic state per body
constexpr int kParF = 16;       // doubles of parameters per body
constexpr int kMaxLevels = 1022;
#ifndef PS_PREFETCH
#define PS_PREFETCH 0  // 1: process_pair requests both body records before it consumes either (experiment)
#endif
constexpr int kMaxStatics = 4;  // static bodies per world whose pairs run in parallel (more -> serial walk)
constexpr int kMaxThreads = 128;
constexpr int kBoostSteps = 8;  // steps of doubled margins after a world had to verify (a restart costs a whole world-step)
constexpr int kCullF = 7;       // doubles of cull-sphere data per body: centre 3, radius vs sphere, radius vs box, margin^2, max deviation^2

struct WorldStats {  // per world, accumulated over steps
    unsigned long long steps, tested, hits, contacts, fallback, overflow, ref_exc;
    unsigned long long fb_margin, fb_capacity, fb_static;
    unsigned long long verified;  // world-steps in which a body left its margin and the culled pairs were re-examined
    double max_pen;
    int nan_flag;
    int boost;  // > 0: a body of this world left its margin recently -> margins doubled for that many more steps
    // SM cycles of thread 0 per phase, summed over steps: predict+cull spheres, near list, levels+sort,
    // wavefronts (or serial walk), static fold + joints + final sweep + publish; and the slowest single step
    unsigned long long cyc[5], cyc_max_step;
    unsigned long long sum_K, sum_levels;
};

struct RecPair {  // one entry of the last step's Collisions map
    int i, j, first, count;
};
struct RecContact {
    double n[3], p[3], pen;
    unsigned feat, pad;
};

struct KArgs {
    double* dyn;          // [total_bodies*18], body-major: dyn[(off+b)*18 + f]  f: x 0-2, v 3-5, R 6-14, L 15-17
    const double* par;    // [total_bodies*16], body-major
    const int* off; const int* cnt;       // n_worlds+1
    WorldStats* stats;    // n_worlds
    unsigned char* hitcnt;  // [total_bodies] colliding pairs each body was in during its last step (margin model)
    double* cur;          // [total_bodies*18] working copy of the state during a step
    double* pred;         // [total_bodies*18] predicted copies
    double* S;            // [total_bodies*3*kMaxStatics] angular-momentum sums of static-body pairs
    int n_worlds, n_steps;
    int world_base;           // first world of this launch (n_worlds is the END of the range); a multiple of 32 / T
    StepParams sp;
    int exact_all_pairs;
    int restore_joints;
    int margin_ncap;      // cap of the hit count that scales a body's motion margin
    double margin_scale;  // safety factor on the motion margin (a violated margin costs the world a second attempt)
    // joints, grouped per world
    const int* joff;      // n_worlds+1 (nullptr when no joints)
    const int* ja;
    const int* jb;
    const double* jla;    // 3/joint
    const double* jlb;
    const unsigned char* jlast;  // bit0: last writer of a, bit1: last writer of b
    // recording of the last step's contacts (nullptr when off)
    RecPair* rec_pairs;       // n_worlds * rec_pair_cap
    RecContact* rec_contacts; // n_worlds * rec_contact_cap
    int* rec_counts;          // n_worlds * 2 : pairs, contacts
    int rec_pair_cap, rec_contact_cap;
    int max_n;                // largest world (sizes the shared-memory carve-up)
    // per-step lists of every world in global memory
    double* cull;             // [total_bodies*kCullF]: per world sx(3N) | rad_s(N) | rad_b(N) | marg2(N) | dev2(N)
    int* row;                 // [total_bodies + n_worlds]: per world N+1 row offsets of the near list
    int* lvl;                 // [n_worlds * (kMaxLevels+2)] level offsets
    unsigned short* last;     // [total_bodies] level of the last near pair of each body
    unsigned char* lists;     // per world: n_pairs * 9 bytes (near u32 | level u16 | order u16 | hit u8)
    const long long* lists_off; // n_worlds+1 byte offsets (16-byte aligned)
    // when set: only worlds with a non-zero flag are stepped (the batch-wide path hands its failed worlds over);
    // bit0 margin violated -> start at the 8x margin, bit1 static pose / too many statics -> exact serial walk
    const unsigned* only_flagged;
    // the redo launch of the batch-wide path comes in two tile widths; each runs only when the number of flagged worlds
    // (counted on the device) is in its range: few worlds -> all 32 lanes on each, many -> the batch's own tile
    const int* redo_count; int redo_lo, redo_hi;
    const int* perm;              // slot -> world inside each segment (nullptr: identity)
    unsigned short* cost;         // per world: wavefront passes (+ serially walked pairs) of the last step, saturating
    // SpatialTree candidate order (ps_step_config.broadphase_kind 1): the ordered candidate list of every world for THIS
    // step, i | j << 16 with i < j, static-static pairs already dropped (csrc/ps_octree.hpp builds it on the host);
    // nullptr = Dummy order, all (i<j) lexicographic
    const unsigned* cand;
    const int* cand_off;      // n_worlds+1
};

// ---- per-world working set: pointers into global scratch + a small shared-memory block -------------
struct Smem {
    double* dyn;     // GLOBAL scratch: 18*N current bodies (body-major)
    double* pred;    // GLOBAL scratch: 18*N predicted bodies
    const double* par;  // GLOBAL: 16*N parameters
    double* S;       // GLOBAL scratch: 3*kMaxStatics*N, [static ordinal][partner]
    double* sx;      // GLOBAL scratch: 3*N cull-sphere centres (predicted positions)
    double* rad_s;   // GLOBAL: N   cull radius against a sphere (margin included)
    double* rad_b;   // GLOBAL: N   cull radius against a box
    double* marg2;   // GLOBAL: N   squared margin
    double* dev2;    // GLOBAL: N   largest squared distance of a predicted position from the sphere centre, once it exceeds marg2
    unsigned* near;  // GLOBAL: n_pairs  i | j<<16
    unsigned short* level;  // GLOBAL: n_pairs
    unsigned short* order;  // GLOBAL: n_pairs
    unsigned char* hit;     // GLOBAL: n_pairs
    int* row;        // GLOBAL: N+1
    int* lvl_start;  // GLOBAL: kMaxLevels+2
    unsigned short* last;   // GLOBAL: N
    unsigned char* pvalid;  // shared: N
    unsigned char* isstat;  // shared: N   0 dynamic, 1+ordinal static
    unsigned char* isbox;   // shared: N
    unsigned char* nhit;    // shared: N   colliding pairs of the body this step (saturating)
    unsigned char* nprev;   // shared: N   ... during the previous step
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// shared memory of ONE world (a CTA hosts 32/TS of them)
__host__ __device__ inline size_t smem_bytes(int N)
{
    return align_up((size_t)(5 * N), 16);
}

__device__ inline void carve(unsigned char* base, int N, Smem& s)
{
    unsigned char* c = base;
    s.pvalid = c; c += N;
    s.isstat = c; c += N;
    s.isbox = c;  c += N;
    s.nhit = c;   c += N;
    s.nprev = c;
}

// ---- body-major accessors (records are 16-byte aligned: 144 B and 128 B per body) ------------------
PSD void ld_dyn(const double* a, int N, int b, Dyn& d)
{
    (void)N;
    const double2* q = reinterpret_cast<const double2*>(a + (size_t)b * kDynF);
    double2 t0 = q[0], t1 = q[1], t2 = q[2], t3 = q[3], t4 = q[4], t5 = q[5], t6 = q[6], t7 = q[7], t8 = q[8];
    d.x = mk(t0.x, t0.y, t1.x);
    d.v = mk(t1.y, t2.x, t2.y);
    d.R.a[0] = t3.x; d.R.a[1] = t3.y; d.R.a[2] = t4.x; d.R.a[3] = t4.y; d.R.a[4] = t5.x; d.R.a[5] = t5.y;
    d.R.a[6] = t6.x; d.R.a[7] = t6.y; d.R.a[8] = t7.x;
    d.L = mk(t7.y, t8.x, t8.y);
}
PSD void st_dyn(double* a, int N, int b, const Dyn& d)
{
    (void)N;
    double2* q = reinterpret_cast<double2*>(a + (size_t)b * kDynF);
    q[0] = make_double2(d.x.x, d.x.y); q[1] = make_double2(d.x.z, d.v.x); q[2] = make_double2(d.v.y, d.v.z);
    q[3] = make_double2(d.R.a[0], d.R.a[1]); q[4] = make_double2(d.R.a[2], d.R.a[3]); q[5] = make_double2(d.R.a[4], d.R.a[5]);
    q[6] = make_double2(d.R.a[6], d.R.a[7]); q[7] = make_double2(d.R.a[8], d.L.x); q[8] = make_double2(d.L.y, d.L.z);
}
// parameter record: 0 1/mass (host-evaluated 1.0/mass, what Vector3D `/` multiplies by), 1-3 iinv,
// 4-6 F, 7-9 T, 10 e, 11 mu_k, 12-14 dims, 15 shape
PSD void ld_par(const double* a, int N, int b, Par& p)
{
    (void)N;
    const double2* q = reinterpret_cast<const double2*>(a + (size_t)b * kParF);
    double2 t0 = __ldg(q), t1 = __ldg(q + 1), t2 = __ldg(q + 2), t3 = __ldg(q + 3), t4 = __ldg(q + 4), t5 = __ldg(q + 5),
            t6 = __ldg(q + 6), t7 = __ldg(q + 7);
    p.inv_mass = t0.x;
    p.is_static = (p.inv_mass == 0.0);  // 1/mass == 0 <=> mass is +-inf (Mass.Infinite)
    p.iinv[0] = t0.y; p.iinv[1] = t1.x; p.iinv[2] = t1.y;
    p.F = mk(t2.x, t2.y, t3.x);
    p.T = mk(t3.y, t4.x, t4.y);
    p.e = t5.x;
    p.mu = t5.y;
    p.dims[0] = t6.x; p.dims[1] = t6.y; p.dims[2] = t7.x;
    p.shape = (int)t7.y;
}

PSD bool same_bits(double a, double b) { return bits_of(a) == bits_of(b); }
PSD bool pose_fixed(const Dyn& a, const Dyn& b)
{
    bool ok = same_bits(a.x.x, b.x.x) && same_bits(a.x.y, b.x.y) && same_bits(a.x.z, b.x.z) &&
              same_bits(a.v.x, b.v.x) && same_bits(a.v.y, b.v.y) && same_bits(a.v.z, b.v.z);
#pragma unroll
    for (int k = 0; k < 9; k++) ok = ok && same_bits(a.R.a[k], b.R.a[k]);
    return ok;
}

// block-wide flags / counters in static shared memory
struct BlockShared {
    int violate;      // a body left its margin -> its culled pairs are re-examined after the wavefronts (phase V)
    int redo;         // ... and one of them can no longer be ruled out -> restart with a wider margin
    int need_serial;  // a static pose moves / list full / too many levels -> exact serial walk
    int K, maxl, nstat;
    unsigned tested, hits, contacts, overflow, ref_exc;
    unsigned long long max_pen_bits;
};

struct Ctx {
    Smem s;
    int N, w;
    const KArgs* a;
    BlockShared* bs;
    // per-thread counters of the current attempt
    unsigned tested, hits, contacts, overflow, ref_exc;
    double max_pen;
};

// record one colliding pair for ps_batch_get_contacts
PSD void record_pair(const Ctx& c, int i, int j, const Contact* cs, int nc)
{
    const KArgs& a = *c.a;
    if (!a.rec_pairs) return;
    int* cnt = a.rec_counts + 2 * c.w;
    int slot = atomicAdd(&cnt[0], 1);
    int first = atomicAdd(&cnt[1], nc);
    if (slot >= a.rec_pair_cap || first + nc > a.rec_contact_cap) return;  // counted; host reports capacity
    RecPair rp; rp.i = i; rp.j = j; rp.first = first; rp.count = nc;
    a.rec_pairs[(size_t)c.w * a.rec_pair_cap + slot] = rp;
    RecContact* out = a.rec_contacts + (size_t)c.w * a.rec_contact_cap + first;
    for (int k = 0; k < nc; k++) {
        RecContact r;
        r.n[0] = cs[k].n.x; r.n[1] = cs[k].n.y; r.n[2] = cs[k].n.z;
        r.p[0] = cs[k].p.x; r.p[1] = cs[k].p.y; r.p[2] = cs[k].p.z;
        r.pen = cs[k].pen; r.feat = cs[k].feat; r.pad = 0;
        out[k] = r;
    }
}

// A predicted position is what the reference tests in every pair of the body that comes up before its next
// write-back, culled pairs included.  Inside the margin the cull stands; outside, the largest deviation is kept and
// phase V re-examines the body's culled pairs with the sphere grown to it.  (The owner of the body at this level is
// the only writer.)
PSD void track_margin(Ctx& c, int b, d3 x)
{
    d3 d = x - mk(c.s.sx[b], c.s.sx[c.N + b], c.s.sx[2 * c.N + b]);
    double dd = d.x * d.x + d.y * d.y + d.z * d.z;
    if (!(dd <= c.s.marg2[b])) {
        c.bs->violate = 1;
        if (!(dd <= c.s.dev2[b])) c.s.dev2[b] = dd;  // a NaN sticks
    }
}

// predicted copy of body b, integrating it first if an earlier pair changed it (lazy refresh).
// check_margin: the position is about to be used for detection under a culled pair list.
PSI void get_pred(Ctx& c, int b, const Par& p, Dyn& out, bool check_margin)
{
    if (c.s.pvalid[b]) {
        ld_dyn(c.s.pred, c.N, b, out);
        return;
    }
    Dyn cur;
    ld_dyn(c.s.dyn, c.N, b, cur);
    integrate(cur, p, c.a->sp, out);
    st_dyn(c.s.pred, c.N, b, out);
    c.s.pvalid[b] = 1;
    if (check_margin) track_margin(c, b, out.x);
}

PSD void bump(unsigned char* p)  // saturating per-body hit counter (the body is owned by this thread at this level)
{
    unsigned v = *p;
    if (v < 255u) *p = (unsigned char)(v + 1u);
}

// one candidate pair of the ordered fold (CollisionResolver.fs:40-46, 13-38).
// fast: static bodies are shared (pose fixed): their angular-momentum sum goes to S[ordinal][partner].
// inlined into its two call sites: measured C2 2.54 -> 2.28 ms per step
PSD void process_pair(Ctx& c, int i, int j, int near_k, bool fast)
{
    Par pi, pj;
    ld_par(c.s.par, c.N, i, pi);
    ld_par(c.s.par, c.N, j, pj);
    Dyn di, dj;
#if PS_PREFETCH
    {   // both records are requested before either is consumed (the two L2 round trips overlap)
        const bool vi = c.s.pvalid[i] != 0, vj = c.s.pvalid[j] != 0;
        ld_dyn(vi ? c.s.pred : c.s.dyn, c.N, i, di);
        ld_dyn(vj ? c.s.pred : c.s.dyn, c.N, j, dj);
        if (!vi) {
            Dyn cur = di;
            integrate(cur, pi, c.a->sp, di);
            st_dyn(c.s.pred, c.N, i, di);
            c.s.pvalid[i] = 1;
            if (fast) track_margin(c, i, di.x);
        }
        if (!vj) {
            Dyn cur = dj;
            integrate(cur, pj, c.a->sp, dj);
            st_dyn(c.s.pred, c.N, j, dj);
            c.s.pvalid[j] = 1;
            if (fast) track_margin(c, j, dj.x);
        }
    }
#else
    get_pred(c, i, pi, di, fast);
    get_pred(c, j, pj, dj, fast);
#endif
    const bool bb = (pi.shape == 1) && (pj.shape == 1);
    int err = 0, ovf = 0, nc;
    c.tested++;
    bool si = fast && pi.is_static, sj = fast && pj.is_static;
    if (!bb) {
        // a pair with a sphere: one contact; everything stays in registers / a 240-byte box record
        Contact cs[1];
        CPData cp[1];
        OBox ob;
        nc = detect_sphere_pair(di, pi, dj, pj, c.a->sp.eps, cs, err, ob);
        if (err) c.ref_exc++;
        if (nc <= 0) return;
        c.hits++;
        c.contacts += 1;
        c.max_pen = fmax(c.max_pen, fabs(cs[0].pen));
        record_pair(c, i, j, cs, 1);
        if (si) di.L = mk(0.0, 0.0, 0.0);
        if (sj) dj.L = mk(0.0, 0.0, 0.0);
#if PS_FUSED_STRAIGHT_SOLVE
        {   // the straight-line solve (ps_physics.cuh, resolve_s); the literal one where a quotient leaves its range
            const d3 v1 = di.v, L1 = di.L, v2 = dj.v, L2 = dj.L;
            const bool s1 = pi.is_static && !pj.is_static && bits_of(v1.x) == 0 && bits_of(v1.y) == 0 && bits_of(v1.z) == 0;
            const bool ok = s1 ? resolve_s<1, true>(di, pi, dj, pj, cs, 1, nullptr, c.a->sp) : resolve_s<1, false>(di, pi, dj, pj, cs, 1, nullptr, c.a->sp);
            if (!ok) {
                PS_COLD_PATH;
                di.v = v1; di.L = L1; dj.v = v2; dj.L = L2;
                resolve_t<1>(di, pi, dj, pj, cs, 1, cp, c.a->sp);
            }
        }
#else
        resolve_t<1>(di, pi, dj, pj, cs, 1, cp, c.a->sp);
#endif
    } else {
        PairScratch sc;  // clip buffers + manifold on the local stack (a per-thread global record was measured: slower)
        Contact* cs = sc.cs;
        nc = detect_bb(di, pi, dj, pj, c.a->sp.eps, cs, err, ovf, sc);
        if (err) c.ref_exc++;
        if (ovf) c.overflow++;
        if (nc <= 0) return;
        c.hits++;
        c.contacts += nc;
        for (int k = 0; k < nc; k++) c.max_pen = fmax(c.max_pen, fabs(cs[k].pen));
        record_pair(c, i, j, cs, nc);
        if (si) di.L = mk(0.0, 0.0, 0.0);
        if (sj) dj.L = mk(0.0, 0.0, 0.0);
        resolve(di, pi, dj, pj, cs, nc, sc.cp, c.a->sp);
    }
    if (fast) c.s.hit[near_k] = 1;
    if (si) {  // a static-static pair never reaches here (canIntersect)
        double* S = c.s.S + 3 * ((c.s.isstat[i] - 1) * c.N + j);
        S[0] = di.L.x; S[1] = di.L.y; S[2] = di.L.z;
    } else {
        st_dyn(c.s.dyn, c.N, i, di);  // write back (CollisionResolver.fs:27-31)
        c.s.pvalid[i] = 0;
        bump(&c.s.nhit[i]);
    }
    if (sj) {
        double* S = c.s.S + 3 * ((c.s.isstat[j] - 1) * c.N + i);
        S[0] = dj.L.x; S[1] = dj.L.y; S[2] = dj.L.z;
    } else {
        st_dyn(c.s.dyn, c.N, j, dj);
        c.s.pvalid[j] = 0;
        bump(&c.s.nhit[j]);
    }
}

// conservative "may these two ever collide during this step" test on the cull spheres
PSD bool near_test(const Smem& s, int N, int i, int j)
{
    if (s.isstat[i] && s.isstat[j]) return false;  // canIntersect (CollisionResolver.fs:58-60)
    double dx = s.sx[i] - s.sx[j], dy = s.sx[N + i] - s.sx[N + j], dz = s.sx[2 * N + i] - s.sx[2 * N + j];
    double dd = dx * dx + dy * dy + dz * dz;
    // radius of i depends on what j is (a sphere against a box can report a hit up to sqrt(3) r from a corner)
    double ri = s.isbox[j] ? s.rad_b[i] : s.rad_s[i];
    double rj = s.isbox[i] ? s.rad_b[j] : s.rad_s[j];
    double r = ri + rj;
    return !(dd > r * r);  // NaN -> near
}

// phase V: how far the cull sphere of body y has to grow to hold every position y was predicted at
PSD double cull_growth(double dev2, double marg2)
{
    if (dev2 <= marg2) return 0.0;
    double g = sqrt(dev2) * (1.0 + 1e-6) + 1e-9 - sqrt(marg2) * (1.0 - 1e-6);  // rounded outwards
    return g > 0.0 ? g : 0.0;
}
PSD bool near_test_grown(const Smem& s, int N, int i, int j, double gi, double gj)
{
    double dx = s.sx[i] - s.sx[j], dy = s.sx[N + i] - s.sx[N + j], dz = s.sx[2 * N + i] - s.sx[2 * N + j];
    double dd = dx * dx + dy * dy + dz * dz;
    double ri = (s.isbox[j] ? s.rad_b[i] : s.rad_s[i]) + gi;
    double rj = (s.isbox[i] ? s.rad_b[j] : s.rad_s[j]) + gj;
    double r = ri + rj;
    return !(dd > r * r);
}

#ifndef PS_MINB32
#define PS_MINB32 16  // resident one-warp CTAs per SM the register allocation must allow (65536 / (32 * regs))
#endif

PSD int warp_max(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// T = lanes per world (tile); the CTA is one warp and hosts 32/T worlds.
//
// Control flow is WARP-uniform: every lane of the warp walks the same phases and the same loops (trip counts are
// the maximum over the warp's tiles, lanes past their tile's bound are predicated off), so that the heavy routines
// (integrate, process_pair) are entered by lanes of different worlds at the same program counter and share the
// warp's instruction issue.  No lane returns early; a tile without a world has N = 0.
//
// MINB = one-warp CTAs per SM the register allocation has to allow: 16 -> 128 registers (batches that need several
// waves are throughput bound and want the residency), 8 -> 255 registers (a batch whose CTAs are all resident at 8 per
// SM is bound by the latency of each warp, and the 128-register build spills the two body records of a pair: measured
// C2 4096 worlds 1.82 -> 1.62 ms, C4 8192 worlds 3.23 -> 2.32 ms, C2 1024 worlds 1.00 -> 0.77 ms per step; C3 on the
// fused kernel alone, 16384 worlds, 57.6 -> 64.7 ms: many waves, there the residency wins).
template <int T, int MINB>
__global__ void __launch_bounds__(32, MINB) step_kernel(KArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BlockShared bs_all[32 / T];
    constexpr unsigned FULL = 0xffffffffu;
    if (a.redo_count) {  // (grid-uniform)
        const int nf = *a.redo_count;
        if (nf < a.redo_lo || nf >= a.redo_hi) return;
    }

    const int tile = threadIdx.x / T;
    const int tid = threadIdx.x % T;
    // slot -> world.  world_base: a launch over a sub-range of the batch.  perm: the worlds of a segment re-ordered by
    // the wavefront passes their last step took, so that the tiles of a warp finish together (a warp runs as many
    // passes as its slowest tile: measured 39 per warp against 24 per tile with the worlds in creation order).
    const int slot = a.world_base + blockIdx.x * (32 / T) + tile;
    const bool inrange = slot < a.n_worlds;
    const int wraw = inrange ? (a.perm ? a.perm[slot] : slot) : 0;
    const unsigned wflag = (inrange && a.only_flagged) ? a.only_flagged[wraw] : 0u;
    const bool live = inrange && (!a.only_flagged || wflag != 0u);
    const int w = live ? wraw : 0;
    BlockShared& bs = bs_all[tile];
    const int off = a.off[w];
    const int N = live ? a.cnt[w] : 0;  // bodies of the world; off[] is the CAPACITY prefix (spare slots for AddObject)
    Ctx c;
    c.N = N; c.w = w; c.a = &a; c.bs = &bs;
    carve(smem_raw + (size_t)tile * smem_bytes(a.max_n), N, c.s);
    Smem& s = c.s;
    {
        double* cu = a.cull + (size_t)off * kCullF;
        s.sx = cu; s.rad_s = cu + 3 * N; s.rad_b = cu + 4 * N; s.marg2 = cu + 5 * N; s.dev2 = cu + 6 * N;
        s.row = a.row + off + w;
        s.lvl_start = a.lvl + (size_t)w * (kMaxLevels + 2);
        s.last = a.last + off;
        unsigned char* o = a.lists + a.lists_off[w];
        const size_t np = (size_t)N * (size_t)(N > 0 ? N - 1 : 0) / 2;
        s.near = reinterpret_cast<unsigned*>(o);
        s.level = reinterpret_cast<unsigned short*>(o + 4 * np);
        s.order = reinterpret_cast<unsigned short*>(o + 6 * np);
        s.hit = o + 8 * np;
    }
    double* gdyn = a.dyn + (size_t)off * kDynF;
    s.dyn = a.cur + (size_t)off * kDynF;
    s.pred = a.pred + (size_t)off * kDynF;
    s.par = a.par + (size_t)off * kParF;
    s.S = a.S + (size_t)off * 3 * kMaxStatics;
    const StepParams sp = a.sp;
    const unsigned* cl = a.cand ? a.cand + a.cand_off[w] : nullptr;
    const int ncand = (a.cand && live) ? a.cand_off[w + 1] - a.cand_off[w] : 0;
    const int Nw = warp_max(N);  // loop bound of the per-body phases (warp-uniform)

    // working copy of the world (coalesced, contiguous per world)
    for (int k = tid; k < kDynF * N; k += T) s.dyn[k] = gdyn[k];
    for (int b = tid; b < N; b += T) {
        s.nprev[b] = a.hitcnt[off + b];
        s.isbox[b] = (__ldg(s.par + (size_t)b * kParF + 15) != 0.0) ? 1 : 0;
        s.isstat[b] = (__ldg(s.par + (size_t)b * kParF) == 0.0) ? 1 : 0;
    }
    __syncwarp();
    if (tid == 0) {  // static ordinals (1-based); more than kMaxStatics -> always the serial walk
        int ns = 0;
        for (int b = 0; b < N; b++)
            if (s.isstat[b]) { ns++; s.isstat[b] = (unsigned char)(ns <= kMaxStatics ? ns : kMaxStatics); }
        bs.nstat = ns;
    }
    __syncwarp();

    unsigned long long t_steps = 0, t_tested = 0, t_hits = 0, t_contacts = 0, t_fallback = 0, t_overflow = 0, t_exc = 0;
    unsigned long long t_fbm = 0, t_fbc = 0, t_fbs = 0, t_ver = 0;
    double t_maxpen = 0.0;
    unsigned long long cyc[5] = {0, 0, 0, 0, 0}, cyc_max = 0, sumK = 0, sumL = 0;
    int boost = live ? a.stats[w].boost : 0;  // per lane, kept equal across the tile
    int npass = 0;                             // passes of the current step (lane 0 of the tile counts for the world)

    for (int step = 0; step < a.n_steps; step++) {
        long long tk0 = clock64(), tk = tk0;
        // per tile: attempt 0 = margin x1, 1 = margin x8, 2 = exact serial walk over all pairs
        int attempt = (a.exact_all_pairs || bs.nstat > kMaxStatics || (wflag & 2u)) ? 2 : ((wflag & 1u) ? 1 : 0);
        int fb_reason = 0;  // 1 margin, 2 levels, 3 static pose
        bool finished = !live;
        npass = 0;
        c.tested = c.hits = c.contacts = c.overflow = c.ref_exc = 0;
        c.max_pen = 0.0;

        for (int pass = 0; pass < 3; pass++) {  // a tile needs at most three attempts
            if (!__any_sync(FULL, !finished)) break;
            const bool run = !finished;
            bool folding = false;  // this pass ends with the static fold (phase E)
            bool fast = run && attempt < 2;
            const bool serial = run && attempt == 2;
            if (run) {
                if (tid == 0) {
                    bs.violate = 0; bs.redo = 0; bs.need_serial = 0; bs.K = 0; bs.maxl = 0;
                    if (a.rec_counts) { a.rec_counts[2 * w] = 0; a.rec_counts[2 * w + 1] = 0; }
                }
                c.tested = c.hits = c.contacts = c.overflow = c.ref_exc = 0;
                c.max_pen = 0.0;
                for (int b = tid; b < N; b += T) { s.pvalid[b] = 0; s.nhit[b] = 0; }
            }
            __syncwarp();

            // ---- A: predicted copy of every body + cull spheres ---------------------------------------
            // The sphere is centred on the PREDICTED position: a body no pair has changed is tested exactly
            // there (deviation 0).  A write-back keeps the position, and each further prediction moves it by
            // dt * (its new velocity): the margin covers n such moves, n estimated from the number of
            // colliding pairs the body was in during the previous step.
            {
                const double mscale = (attempt == 0 ? (boost > 0 ? 2.0 : 1.0) : 8.0) * a.margin_scale;
                for (int b = tid; b < Nw; b += T) {
                    const bool on = fast && b < N;
                    Par p;
                    Dyn cur, pr;
                    if (on) {
                        ld_par(s.par, N, b, p);
                        ld_dyn(s.dyn, N, b, cur);
                    }
                    if (on) integrate(cur, p, sp, pr);  // lanes of all the warp's worlds enter together
                    if (on) {
                        st_dyn(s.pred, N, b, pr);
                        s.pvalid[b] = 1;
                        // A static body must not move under integrate() for its pairs to commute.  The reference tests it at
                        // integrate(current) in EVERY pair and writes that copy back on a hit, so what has to be a fixed point is
                        // the PREDICTED pose: pose -> P1 = integrate(pose) -> integrate(P1) == P1.  (First step after creation: the
                        // -0.0 entries of a yaw = pi ground become +0.0 once; every pair of that step already sees P1, and the
                        // final sweep publishes P1 whether or not the body was hit.)
                        if (p.is_static && !pose_fixed(cur, pr)) {
                            Dyn pr2;
                            integrate_slow(pr, p, sp, pr2);
                            if (!pose_fixed(pr, pr2)) bs.need_serial = 1;
                        }
                        s.sx[b] = pr.x.x; s.sx[N + b] = pr.x.y; s.sx[2 * N + b] = pr.x.z;
                        double speed = sqrt(pr.v.x * pr.v.x + pr.v.y * pr.v.y + pr.v.z * pr.v.z);
                        double acc = p.inv_mass * sqrt(p.F.x * p.F.x + p.F.y * p.F.y + p.F.z * p.F.z);
                        double n = (double)(s.nprev[b] < a.margin_ncap ? s.nprev[b] : a.margin_ncap) + 2.0;
                        double m = p.is_static ? 0.0 : mscale * (0.01 + n * sp.dt * (speed + n * acc * sp.dt + 0.5));
                        if (!(m == m)) m = 0.0;  // NaN state: the world is flagged anyway
                        double rs, rb;
                        if (p.shape == 0) {
                            rs = p.dims[0]; rb = 1.7320508075688774 * p.dims[0];
                        } else {
                            double hx = 0.5 * p.dims[0], hy = 0.5 * p.dims[1], hz = 0.5 * p.dims[2];
                            rs = rb = sqrt(hx * hx + hy * hy + hz * hz);
                        }
                        const double slack = 1.0 + 1e-6;
                        s.rad_s[b] = rs * slack + 1e-9 + m;
                        s.rad_b[b] = rb * slack + 1e-9 + m;
                        s.marg2[b] = m * m;
                        s.dev2[b] = 0.0;
                        s.last[b] = 0;
                    }
                }
            }
            __syncwarp();
            if (fast && bs.need_serial) {  // a static pose still moves under integrate(): exact serial walk
                fast = false;              // (state untouched so far: no reload needed)
                attempt = 2; fb_reason = 3;
            }
            if (tid == 0 && live) { long long t = clock64(); cyc[0] += t - tk; tk = t; }
            // ---- B: near list ---------------------------------------------------------------------------------
            // SpatialTree order: the near list is the candidate list minus the pairs the cull rules out, order kept by
            // a ballot scan over chunks of T candidates (warp-uniform trip count)
            if (a.cand) {
                const unsigned tmask_ = (T == 32) ? FULL : (((1u << T) - 1u) << (tile * T));
                const int nc_ = fast ? ncand : 0;
                const int ncw = warp_max(nc_);
                int kacc = 0;
                for (int base = 0; base < ncw; base += T) {
                    const int q = base + tid;
                    unsigned ij = 0;
                    bool pred = false;
                    if (q < nc_) {
                        ij = cl[q];
                        pred = near_test(s, N, (int)(ij & 0xffff), (int)(ij >> 16));
                    }
                    const unsigned mine = (__ballot_sync(FULL, pred) & tmask_) >> (tile * T);
                    if (pred) {
                        const int pos = kacc + __popc(mine & ((1u << tid) - 1u));
                        s.near[pos] = ij;
                        s.hit[pos] = 0;
                    }
                    kacc += __popc(mine);
                }
                if (fast && tid == 0) bs.K = kacc;
            }
            // Dummy order: lexicographic order kept by a row scan
            if (fast && !a.cand) {
                for (int i = tid; i < N; i += T) {
                    int cn = 0;
                    for (int j = i + 1; j < N; j++) cn += near_test(s, N, i, j) ? 1 : 0;
                    s.row[i] = cn;
                }
            }
            __syncwarp();
            if (fast && !a.cand && tid == 0) {
                int acc = 0;
                for (int i = 0; i < N; i++) { int t = s.row[i]; s.row[i] = acc; acc += t; }
                s.row[N] = acc;
                bs.K = acc;
            }
            __syncwarp();
            const int K = fast ? bs.K : 0;
            if (fast && !a.cand) {
                for (int i = tid; i < N; i += T) {
                    int o = s.row[i];
                    for (int j = i + 1; j < N; j++)
                        if (near_test(s, N, i, j)) {
                            s.near[o] = (unsigned)i | ((unsigned)j << 16);
                            s.hit[o] = 0;
                            o++;
                        }
                }
            }
            __syncwarp();
            if (tid == 0 && live) { long long t = clock64(); cyc[1] += t - tk; tk = t; }
            // ---- C: wavefront levels (serial scan of the ordered list by one lane per world) ------------------
            if (fast && tid == 0) {
                int maxl = 0;
                for (int k = 0; k < K; k++) {
                    unsigned ij = s.near[k];
                    int i = ij & 0xffff, j = ij >> 16;
                    int li = s.isstat[i] ? 0 : s.last[i];
                    int lj = s.isstat[j] ? 0 : s.last[j];
                    int l = (li > lj ? li : lj) + 1;
                    s.level[k] = (unsigned short)l;
                    if (!s.isstat[i]) s.last[i] = (unsigned short)l;
                    if (!s.isstat[j]) s.last[j] = (unsigned short)l;
                    if (l > maxl) maxl = l;
                }
                bs.maxl = maxl;
                if (maxl > kMaxLevels) bs.need_serial = 1;
            }
            __syncwarp();
            if (fast && bs.need_serial) {  // more levels than the table holds: exact serial walk next pass
                fast = false;
                attempt = 2; fb_reason = 2;
            }
            // counting sort of the near list by level (order inside a level is free)
            const int maxl = fast ? bs.maxl : 0;
            if (fast)
                for (int l = tid; l <= maxl + 1; l += T) s.lvl_start[l] = 0;
            __syncwarp();
            if (fast)
                for (int k = tid; k < K; k += T) atomicAdd(&s.lvl_start[s.level[k]], 1);
            __syncwarp();
            if (fast && tid == 0) {
                int acc = 0;
                for (int l = 0; l <= maxl + 1; l++) { int t = s.lvl_start[l]; s.lvl_start[l] = acc; acc += t; }
            }
            __syncwarp();
            // scatter; lvl_start[l] is advanced to the END of level l, i.e. the start of l+1
            if (fast)
                for (int k = tid; k < K; k += T) {
                    int pos = atomicAdd(&s.lvl_start[s.level[k]], 1);
                    s.order[pos] = (unsigned short)k;
                }
            __syncwarp();
            if (tid == 0 && live) { long long t = clock64(); cyc[2] += t - tk; tk = t; if (fast) { sumK += K; sumL += maxl; } }

            // ---- D: wavefronts.  Every tile walks ITS levels, T pairs per round; the rounds are warp-uniform ----
            {
                int lvl = 1, pos = 0;
                bool dact = fast && maxl >= 1;
                for (;;) {
                    if (!__any_sync(FULL, dact)) break;
                    if (dact) npass++;
                    int pi_ = 0, pj_ = 0, pk_ = 0;
                    bool valid = false;
                    int end = 0, beg = 0;
                    if (dact) {
                        beg = s.lvl_start[lvl - 1]; end = s.lvl_start[lvl];
                        int q = beg + pos + tid;
                        if (q < end) {
                            pk_ = s.order[q];
                            unsigned ij = s.near[pk_];
                            pi_ = ij & 0xffff; pj_ = ij >> 16;
                            valid = true;
                        }
                    }
                    if (valid) process_pair(c, pi_, pj_, pk_, true);
                    __syncwarp();
                    if (dact) {
                        pos += T;
                        if (beg + pos >= end) {
                            lvl++; pos = 0;
                            if (lvl > maxl) dact = false;
                        }
                    }
                }
            }
            // ---- exact serial walk over ALL candidates in order, one lane per world: Dummy order (BroadPhase.fs:33-39)
            // or the SpatialTree list ---------------------------------------------------------------------------------
            {
                int si = 0, sj = 1, sk = 0;
                bool sact = serial && tid == 0 && (a.cand ? ncand > 0 : N >= 2);
                for (;;) {
                    if (!__any_sync(FULL, sact)) break;
                    if (sact) npass++;
                    if (sact && a.cand) { const unsigned ij = cl[sk]; si = (int)(ij & 0xffff); sj = (int)(ij >> 16); }
                    bool valid = sact && !(s.isstat[si] && s.isstat[sj]);
                    if (valid) process_pair(c, si, sj, 0, false);
                    if (sact) {
                        if (a.cand) {
                            if (++sk >= ncand) sact = false;
                        } else {
                            sj++;
                            if (sj >= N) { si++; sj = si + 1; if (sj >= N) sact = false; }
                        }
                    }
                }
            }
            __syncwarp();
            if (tid == 0 && live) { long long t = clock64(); cyc[3] += t - tk; tk = t; }

            // ---- R: predicted copy of every body a pair changed after its last prediction: the position the reference
            // tests in the body's remaining (culled) pairs, so it is held to the margin as well.  The final sweep (F)
            // reuses the copy.
            for (int b = tid; b < Nw; b += T) {
                const bool need = fast && b < N && !s.pvalid[b];
                Par p;
                Dyn cur, pr;
                if (need) {
                    ld_par(s.par, N, b, p);
                    ld_dyn(s.dyn, N, b, cur);
                }
                if (need) integrate(cur, p, sp, pr);
                if (need) {
                    st_dyn(s.pred, N, b, pr);
                    s.pvalid[b] = 1;
                    track_margin(c, b, pr.x);
                }
            }
            __syncwarp();
            // ---- V: a body left its margin.  A culled pair (b, x) stays impossible if the spheres grown to the largest
            // deviations still do not touch; only a pair that was culled and now touches voids the step.
            if (fast && bs.violate) {
                for (int b = tid; b < N; b += T) {
                    const double d2 = s.dev2[b], m2 = s.marg2[b];
                    if (d2 <= m2) continue;
                    if (!(d2 == d2)) { bs.redo = 1; continue; }
                    const double gb = cull_growth(d2, m2);
                    for (int x = 0; x < N; x++) {
                        if (x == b || (s.isstat[b] && s.isstat[x])) continue;
                        if (near_test(s, N, b, x)) continue;  // on the near list: tested where it stood
                        if (near_test_grown(s, N, b, x, gb, cull_growth(s.dev2[x], s.marg2[x]))) { bs.redo = 1; break; }
                    }
                }
                if (tid == 0 && live) t_ver++;
            }
            __syncwarp();  // (all lanes of the warp: tiles diverge above)
            if (fast) {
                if (bs.redo) {  // restart from the published state with the wider margin (or serially)
                    for (int k = tid; k < kDynF * N; k += T) s.dyn[k] = gdyn[k];
                    fb_reason = 1;
                    attempt++;
                } else {
                    finished = true;
                    folding = true;
                }
            } else if (serial) {
                finished = true;
            }
            // ---- E: static bodies: L = fold over their colliding pairs in pair order.  The tile first compacts, in
            // order, the partners of the colliding pairs that involve a static body (ballot scan over chunks of T
            // near pairs, warp-uniform trip count; the level-order table is free by now and takes the list), then one
            // lane per static body folds its entries: the dependent additions of a ~20-entry list instead of a
            // latency-bound walk over the whole near list by that one lane.
            {
                const unsigned tmask_ = (T == 32) ? FULL : (((1u << T) - 1u) << (tile * T));
                const int Kf = folding ? K : 0;
                const int Kw = warp_max(Kf);
                int cnt = 0;
                for (int base = 0; base < Kw; base += T) {
                    const int k = base + tid;
                    bool pred = false;
                    unsigned short ent = 0;
                    if (k < Kf && s.hit[k]) {
                        const unsigned ij = s.near[k];
                        const int i = ij & 0xffff, j = ij >> 16;
                        // at most one body of a near pair is static (canIntersect); entry = partner | ordinal << 12
                        if (s.isstat[i]) { pred = true; ent = (unsigned short)(j | ((s.isstat[i] - 1) << 12)); }
                        else if (s.isstat[j]) { pred = true; ent = (unsigned short)(i | ((s.isstat[j] - 1) << 12)); }
                    }
                    const unsigned mine = (__ballot_sync(FULL, pred) & tmask_) >> (tile * T);
                    if (pred) s.order[cnt + __popc(mine & ((1u << tid) - 1u))] = ent;
                    cnt += __popc(mine);
                }
                __syncwarp();
                if (folding)
                    for (int b = tid; b < N; b += T) {
                        if (!s.isstat[b]) continue;
                        Par p; ld_par(s.par, N, b, p);
                        double* cb = s.dyn + (size_t)b * kDynF;
                        double* pb = s.pred + (size_t)b * kDynF;
                        d3 L = mk(cb[15], cb[16], cb[17]);
                        d3 dL = scale(sp.dt, p.T);
                        const unsigned ord = (unsigned)(s.isstat[b] - 1);
                        const double* Sb = s.S + 3 * ord * N;
                        for (int q = 0; q < cnt; q++) {
                            const unsigned ent = s.order[q];
                            if ((ent >> 12) != ord) continue;
                            const int other = ent & 0xfff;
                            L = L + dL;  // the predicted copy's integration
                            L = L + mk(Sb[3 * other], Sb[3 * other + 1], Sb[3 * other + 2]);
                        }
                        cb[15] = L.x; cb[16] = L.y; cb[17] = L.z;
                        // pose is a fixed point: the final sweep only has to add dt*T to L
                        d3 Lf = L + dL;
                        pb[15] = Lf.x; pb[16] = Lf.y; pb[17] = Lf.z;
                    }
            }
            // (a tile that switched to the serial walk above has run == true, fast == false, serial == false:
            //  it goes round again with attempt == 2)
            __syncwarp();
        }
        {   // a world that had to verify (or restart) keeps doubled margins for a while; otherwise the boost runs out
            const unsigned tm_ = (T == 32) ? FULL : (((1u << T) - 1u) << (tile * T));
            const bool hot = (__ballot_sync(FULL, live && (bs.violate || fb_reason == 1)) & tm_) != 0u;
            boost = hot ? kBoostSteps : (boost > 0 ? boost - 1 : 0);
        }
        if (tid == 0 && live && fb_reason != 0 && !a.exact_all_pairs && !a.only_flagged) {
            t_fallback++;
            if (fb_reason == 1) t_fbm++; else if (fb_reason == 2) t_fbc++; else t_fbs++;
        }

        // ---- J: joints (JointsRestorer.fs:9-26, Joints.fs:26-94) -----------------------------------------------
        // Every joint reads the state at ENTRY (Jacobi) and the results are written back in joint-list order, the
        // last writer of a body wins (the host marks it).  Results are staged in the predicted-copy area of the
        // bodies they replace and committed after all joints have read their inputs.
        if (a.restore_joints && a.joff) {
            const int j0 = a.joff[w], nj = live ? a.joff[w + 1] - j0 : 0;
            for (int b = tid; b < N; b += T) s.last[b] = 0;
            __syncwarp();
            for (int base = 0; base < nj; base += T) {
                if (base + tid < nj) {
                    const int q = j0 + base + tid;
                    const int ba = a.ja[q], bb = a.jb[q];
                    const unsigned char lastf = a.jlast[q];
                    Par p1, p2;
                    ld_par(s.par, N, ba, p1);
                    ld_par(s.par, N, bb, p2);
                    Dyn o1, o2;
                    ld_dyn(s.dyn, N, ba, o1);
                    ld_dyn(s.dyn, N, bb, o2);
                    d3 la = mk(a.jla[3 * q], a.jla[3 * q + 1], a.jla[3 * q + 2]);
                    d3 lb = mk(a.jlb[3 * q], a.jlb[3 * q + 1], a.jlb[3 * q + 2]);
                    m33 Iw1 = inertia_inv_world(o1.R, p1.iinv), Iw2 = inertia_inv_world(o2.R, p2.iinv);
                    d3 aw1 = mulMV(o1.R, la), aw2 = mulMV(o2.R, lb);  // JointImpulseData.init (Joints.fs:26-40)
                    m33 K;
#pragma unroll
                    for (int k = 0; k < 9; k++) K.a[k] = 0.0;
                    add_K(K, p1, Iw1, aw1);
                    add_K(K, p2, Iw2, aw2);
                    d3 vrel = velocity_at(o2, Iw2, lb) - velocity_at(o1, Iw1, la);  // local offsets (Q14)
                    d3 imp = mulMV(K, mk(0.0, 0.0, 0.0) - vrel);
                    for (int it = 0; it < sp.iterations; it++) {
                        apply_impulse(o1, p1, la, imp);
                        apply_impulse(o2, p2, lb, -imp);
                    }
                    // changePhysicalObjects (SimulatorState.fs:149-162) pairs the objects in ASCENDING ID order with the
                    // results in joint order (quirk Q15): the first result goes to the lower id, the second to the higher
                    // one — for a joint given as (a > b) the two bodies swap states; a self joint keeps the first result
                    const int t1 = ba < bb ? ba : bb, t2 = ba < bb ? bb : ba;
                    if (lastf & 1) { st_dyn(s.pred, N, t1, o1); s.last[t1] = 1; }
                    if ((lastf & 2) && ba != bb) { st_dyn(s.pred, N, t2, o2); s.last[t2] = 1; }
                }
            }
            __syncwarp();
            for (int b = tid; b < N; b += T)
                if (s.last[b]) {
                    Dyn d;
                    ld_dyn(s.pred, N, b, d);
                    st_dyn(s.dyn, N, b, d);
                    s.pvalid[b] = 0;
                }
            __syncwarp();
        }

        // ---- F: final integration of every body (SimulatorState.fs:169-173) -----------------------------------
        for (int b = tid; b < Nw; b += T) {
            const bool on = b < N;
            const bool need = on && !s.pvalid[b];
            Dyn out, cur;
            Par p;
            if (need) {
                ld_par(s.par, N, b, p);
                ld_dyn(s.dyn, N, b, cur);
            }
            if (need) integrate(cur, p, sp, out);
            if (on) {
                if (!need) ld_dyn(s.pred, N, b, out);
                st_dyn(s.dyn, N, b, out);
                st_dyn(gdyn, N, b, out);  // publish the step (also the restart point of the next step's fallback)
                s.nprev[b] = s.nhit[b];
            }
        }
        __syncwarp();
        if (tid == 0 && live) { long long t = clock64(); cyc[4] += t - tk; if ((unsigned long long)(t - tk0) > cyc_max) cyc_max = t - tk0; }

        // fold per-lane counters into the tile's block
        if (tid == 0) { bs.tested = bs.hits = bs.contacts = bs.overflow = bs.ref_exc = 0; bs.max_pen_bits = 0ull; }
        __syncwarp();
        if (live) {
            atomicAdd(&bs.tested, c.tested); atomicAdd(&bs.hits, c.hits); atomicAdd(&bs.contacts, c.contacts);
            atomicAdd(&bs.overflow, c.overflow); atomicAdd(&bs.ref_exc, c.ref_exc);
            atomicMax(&bs.max_pen_bits, (unsigned long long)__double_as_longlong(c.max_pen));
        }
        __syncwarp();
        if (tid == 0 && live) {
            t_steps++; t_tested += bs.tested; t_hits += bs.hits; t_contacts += bs.contacts;
            t_overflow += bs.overflow; t_exc += bs.ref_exc;
            t_maxpen = fmax(t_maxpen, __longlong_as_double((long long)bs.max_pen_bits));
        }
        __syncwarp();
    }

    // NaN/Inf scan + stats write-out
    int bad = 0;
    for (int k = tid; k < kDynF * N; k += T) bad |= !isfinite(s.dyn[k]);
    const unsigned tmask = (T == 32) ? FULL : (((1u << T) - 1u) << (tile * T));
    bad = (__ballot_sync(FULL, bad) & tmask) != 0u;
    for (int b = tid; b < N; b += T) a.hitcnt[off + b] = s.nprev[b];
    if (tid == 0 && live && a.cost) a.cost[w] = (unsigned short)(npass < 65535 ? npass : 65535);
    if (tid == 0 && live) {
        WorldStats& st = a.stats[w];
        st.steps += t_steps; st.tested += t_tested; st.hits += t_hits; st.contacts += t_contacts;
        st.fallback += t_fallback; st.overflow += t_overflow; st.ref_exc += t_exc;
        st.fb_margin += t_fbm; st.fb_capacity += t_fbc; st.fb_static += t_fbs; st.verified += t_ver;
        st.max_pen = fmax(st.max_pen, t_maxpen);
        st.nan_flag = bad ? 1 : 0;
        st.boost = boost;
        for (int k = 0; k < 5; k++) st.cyc[k] += cyc[k];
        if (cyc_max > st.cyc_max_step) st.cyc_max_step = cyc_max;
        st.sum_K += sumK; st.sum_levels += sumL;
    }
}

}  // namespace psk
