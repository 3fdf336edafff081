// ps_capi.cu — the C ABI of libps_b200.so (include/ps_b200.h): handle management, host<->device
// marshalling between the ABI layout (per-field, xyz-interleaved) and the per-world planar layout
// the step kernel stages into shared memory, and the kernel launches.
//
// There is no CPU path in this file: every entry point that computes runs a CUDA kernel or fails
// with PS_ERR_CUDA.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ps_b200.h"
#include "ps_step_kernel.cuh"
#include "ps_wide.cuh"
#include "ps_octree.hpp"

using namespace psk;

namespace {

thread_local std::string g_err;
thread_local int g_device = -1;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(x)                                                                                              \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) return fail(PS_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// host-side description of the batch in ABI layout (kept for add_body / reset / candidates)
struct HostBodies {
    std::vector<int32_t> off;
    std::vector<int32_t> shape;
    std::vector<double> dims, mass, iinv, pos, vel, rot, L, force, torque, e, muk, mus;
    int n() const { return (int)mass.size(); }
};
struct HostJoints {
    std::vector<int32_t> world, a, b;
    std::vector<double> la, lb;
};

// ps_batch_step_host on the batch-wide path: a chunk of worlds runs its own chain of rounds on its own stream, so that the
// copies of one chunk and the (latency-bound) rounds of the others overlap.  A chunk owns the small tables of a pass
// (buckets, counters, the grid estimates) and a SLICE of the batch's pair / hit / contact lists; everything indexed by world
// or slot is the batch's.
struct WideChunk {
    int w0 = 0, w1 = 0, g0 = 0, g1 = 0;
    unsigned* d_small = nullptr;  // bucket | bucket_cur | hit_cnt | cpool_cur | surv_cnt | nlev | nflag
    size_t e_off = 0, c_off = 0;  // first entry / contact of the slice
    unsigned e_cap = 0, c_cap = 0;
    int stat_first = 0, stat_count = 0;  // its static bodies in d_stat_slots
    unsigned* h_est = nullptr;  // pinned
    cudaEvent_t est_ready = nullptr;
    bool est_pending = false, est_valid = false;
    std::vector<unsigned> est;
};

struct Batch {
    int device = 0;
    ps_step_config cfg{};
    HostBodies init, cur;  // `cur`: parameters of the bodies as of the last upload (compact ABI layout); dynamic state lives on the device
    // Device layout: world w owns the slots [cap_off[w], cap_off[w + 1]) of every per-body array and fills the first
    // cnt[w] of them; the spare slots are what makes Simulator.AddObject a record write instead of a re-upload
    // (SimulatorState.fs:89-107).  abi_off = prefix of cnt = the caller's compact body index.
    std::vector<int> cap_off, cnt, abi_off;
    std::vector<double> slot_mass;   // per slot (static tests on the host: candidates, trees)
    std::vector<HostBodies> added;   // bodies added since the last upload, in order (spliced into `cur` when a re-layout is needed)
    std::vector<int> added_world;
    int n_slots = 0;
    int *d_cnt = nullptr, *d_abi_off = nullptr;
    HostJoints joints;
    int n_worlds = 0, n_bodies = 0, n_dynamic = 0, max_n = 0;
    int threads = 8;  // lanes per world
    int minb = 16;    // register build of the step kernel: one-warp CTAs per SM it allows (16 -> 128 regs, 8 -> 255)
    size_t smem = 0;
    // device
    double* d_dyn = nullptr;
    double* d_par = nullptr;
    double* d_stage = nullptr;  // 18 doubles/body staging in ABI layout: pos | vel | rot | L
    int* d_off = nullptr;
    WorldStats* d_stats = nullptr;
    unsigned char* d_hitcnt = nullptr;
    unsigned short* d_cost = nullptr;  // per world: passes of its last step (step kernel)
    int* d_perm = nullptr;             // slot -> world, sorted by cost inside each segment
    int seg = 0;                       // worlds per segment (= chunk of ps_batch_step_host); 0: no re-ordering
    bool cost_valid = false, perm_valid = false;
    int steps_since_sort = 1 << 20;
    double *d_cur = nullptr, *d_pred = nullptr, *d_S = nullptr;  // per-world scratch of the step kernel
    double* d_cull = nullptr;         // cull spheres, 6 doubles per body
    int* d_row = nullptr;             // near-list row offsets
    unsigned char* d_lists = nullptr; // near list | level | order | hit of every world (9 B per pair of the world)
    long long* d_lists_off = nullptr;
    int* d_lvl = nullptr;
    unsigned short* d_last = nullptr;
    // batch-wide path (ps_wide.cuh)
    bool broken = false;  // a failed re-upload left no device state: entry points refuse until ps_batch_reset succeeds
    bool wide = false;
    bool wide_device_rounds = false;  // rounds driven from the device (no host read-back inside a step)
    int* d_body_world = nullptr;
    unsigned char *d_pvalid = nullptr, *d_nhit = nullptr, *d_isstat = nullptr, *d_isbox = nullptr, *d_always_fused = nullptr;
    int *d_K = nullptr, *d_maxl = nullptr;
    unsigned *d_wflags = nullptr, *d_bucket = nullptr, *d_bucket_cur = nullptr, *d_hit_cnt = nullptr, *d_hitlist = nullptr, *d_hitlist2 = nullptr, *d_cpool_cur = nullptr;
    psw::Entry* d_entries = nullptr;
    Contact* d_cpool = nullptr;
    unsigned entries_cap = 0, cpool_cap = 0;
    unsigned* h_bucket = nullptr;  // pinned
    unsigned* d_tmp = nullptr;
    unsigned *d_surv = nullptr, *d_surv_cnt = nullptr;
    int* d_nlev = nullptr;
    int* d_nflag = nullptr;
    int* d_stat_slots = nullptr;   // slots of the static bodies
    int n_stat_slots = 0;
    // what the rounds of an earlier step held, read back asynchronously and never waited for: sizes the grids of the
    // device-counted round kernels (ps_wide.cuh).  [0..kBuckets] bucket starts, then hit counts, then survivor counts
    unsigned* h_est = nullptr;       // pinned
    cudaEvent_t est_ready = nullptr;
    bool est_pending = false, est_valid = false;
    std::vector<unsigned> est;       // host copy the launches are sized from
    int rounds_grid = 0;       // CTAs of the cooperative rounds kernel (all resident at once)
    int sm_count = 148;
    unsigned team_max = 0;     // box-box survivors per round up to which every survivor gets a team of 8 lanes
    unsigned long long* d_tmp_pen = nullptr;
    double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // PS_B200_WIDE_PROF: setup, integrate, ss, sb, bb, resolve, tail, fused
    long long prof_steps = 0, prof_rounds = 0;
    std::vector<float> prof_round;  // last profiled step: per round c0, c1, c2, detect ms, bb ms, resolve ms
    int *d_joff = nullptr, *d_ja = nullptr, *d_jb = nullptr;
    double *d_jla = nullptr, *d_jlb = nullptr;
    unsigned char* d_jlast = nullptr;
    RecPair* d_rec_pairs = nullptr;
    RecContact* d_rec_contacts = nullptr;
    int* d_rec_counts = nullptr;
    int rec_pair_cap = 0, rec_contact_cap = 0;
    ps_stats* d_stats_sum = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_own = nullptr, ev_foreign = nullptr;  // ordering between the handle's stream and a caller's stream
    std::vector<WideChunk> wchunks;  // ps_batch_step_host on the batch-wide path (built at the first such call)
    std::vector<int> h_stat_slots;   // host copy of d_stat_slots (ascending)
    std::vector<cudaStream_t> chunk_streams;  // ps_batch_step_host: one stream per world chunk
    std::vector<cudaEvent_t> chunk_done;
    cudaEvent_t chunk_start = nullptr;
    long long launches = 0;
    // SpatialTree candidate order (broadphase_kind 1): one persistent tree per world on the host (the reference keeps it
    // in SimulatorState, BroadPhase.fs:100-127), extents come from the device, the ordered list goes back
    bool tree_mode = false;
    std::vector<pst::Tree<3>> trees;
    double* d_ext = nullptr;     // 6 doubles per body: AABB size xyz, min corner xyz
    double* h_ext = nullptr;     // pinned
    unsigned* d_cand = nullptr;  // candidate lists of all worlds for the coming step
    int* d_cand_off = nullptr;
    std::vector<uint32_t> h_cand;
    std::vector<int> h_cand_off;
};

// ---- layout kernels ------------------------------------------------------------------------------
// ABI (pos[3n] vel[3n] rot[9n] L[3n]) <-> per-world planar dyn
// one thread per body of the worlds [w0, w1) in the caller's compact order; abi = prefix of the body counts, off = first
// slot of every world on the device
__device__ __forceinline__ int slot_of(const int* __restrict__ abi, const int* __restrict__ off, int w0, int w1, int s)
{
    int lo = w0, hi = w1;  // world of compact index s: abi[lo] <= s < abi[lo + 1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (abi[mid] <= s) lo = mid; else hi = mid;
    }
    return off[lo] + (s - abi[lo]);
}
__global__ void pack_dyn_kernel(const int* __restrict__ abi, const int* __restrict__ off, int w0, int w1, const double* __restrict__ pos,
                                const double* __restrict__ vel, const double* __restrict__ rot,
                                const double* __restrict__ L, double* __restrict__ dyn)
{
    const int body0 = abi[w0];
    int c = blockIdx.x * blockDim.x + threadIdx.x + body0;
    if (c >= abi[w1]) return;
    double* d = dyn + (size_t)slot_of(abi, off, w0, w1, c) * kDynF;  // body-major record
    int s = c - body0;  // index into the caller arrays (which start at world w0)
    for (int k = 0; k < 3; k++) {
        d[k] = pos[3 * s + k];
        d[3 + k] = vel[3 * s + k];
        d[15 + k] = L[3 * s + k];
    }
    for (int k = 0; k < 9; k++) d[6 + k] = rot[9 * s + k];
}
__global__ void unpack_dyn_kernel(const int* __restrict__ abi, const int* __restrict__ off, int w0, int w1, double* __restrict__ pos,
                                  double* __restrict__ vel, double* __restrict__ rot, double* __restrict__ L,
                                  const double* __restrict__ dyn)
{
    const int body0 = abi[w0];
    int c = blockIdx.x * blockDim.x + threadIdx.x + body0;
    if (c >= abi[w1]) return;
    const double* d = dyn + (size_t)slot_of(abi, off, w0, w1, c) * kDynF;
    int s = c - body0;
    for (int k = 0; k < 3; k++) {
        if (pos) pos[3 * s + k] = d[k];
        if (vel) vel[3 * s + k] = d[3 + k];
        if (L) L[3 * s + k] = d[15 + k];
    }
    if (rot)
        for (int k = 0; k < 9; k++) rot[9 * s + k] = d[6 + k];
}

// Simulator.ApplyImpulse -> RigidBodyMotion.applyImpulse (RigidBody.fs:194-197)
__global__ void apply_impulse_kernel(double* dyn, const double* par, const int* off, int w, int b, double jx, double jy,
                                     double jz, double ox, double oy, double oz)
{
    int o = off[w], N = 0;
    Dyn d;
    Par p;
    ld_dyn(dyn + (size_t)o * kDynF, N, b, d);
    ld_par(par + (size_t)o * kParF, N, b, p);
    apply_impulse(d, p, mk(ox, oy, oz), mk(jx, jy, jz));
    st_dyn(dyn + (size_t)o * kDynF, N, b, d);
}

// K6: per-world counters -> one ps_stats (integer sums + max)
__global__ void stats_reduce_kernel(const WorldStats* ws, const int* off, const int* cnt, const double* par, int n_worlds, ps_stats* out)
{
    unsigned long long steps = 0, tested = 0, hits = 0, contacts = 0, fallback = 0, overflow = 0, exc = 0, nanw = 0, bsteps = 0;
    unsigned long long fbm = 0, fbc = 0, fbs = 0, ver = 0;
    double mp = 0.0;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_worlds; w += gridDim.x * blockDim.x) {
        const WorldStats& s = ws[w];
        steps += s.steps; tested += s.tested; hits += s.hits; contacts += s.contacts;
        fallback += s.fallback; overflow += s.overflow; exc += s.ref_exc; nanw += s.nan_flag ? 1 : 0;
        fbm += s.fb_margin; fbc += s.fb_capacity; fbs += s.fb_static; ver += s.verified;
        mp = fmax(mp, s.max_pen);
        int o = off[w], N = cnt[w], dynb = 0;
        const double* p = par + (size_t)o * kParF;
        for (int b = 0; b < N; b++) dynb += (p[(size_t)b * kParF] != 0.0);
        bsteps += s.steps * (unsigned long long)dynb;
    }
    atomicAdd((unsigned long long*)&out->steps, steps);
    atomicAdd((unsigned long long*)&out->body_steps, bsteps);
    atomicAdd((unsigned long long*)&out->pairs_tested, tested);
    atomicAdd((unsigned long long*)&out->pairs_colliding, hits);
    atomicAdd((unsigned long long*)&out->contacts, contacts);
    atomicAdd((unsigned long long*)&out->fallback_steps, fallback);
    atomicAdd((unsigned long long*)&out->nan_worlds, nanw);
    atomicAdd((unsigned long long*)&out->contact_overflow, overflow);
    atomicAdd((unsigned long long*)&out->reference_exceptions, exc);
    atomicAdd((unsigned long long*)&out->fallback_margin_steps, fbm);
    atomicAdd((unsigned long long*)&out->fallback_capacity_steps, fbc);
    atomicAdd((unsigned long long*)&out->fallback_static_steps, fbs);
    atomicAdd((unsigned long long*)&out->margin_verified_steps, ver);
    atomicMax((unsigned long long*)&out->max_penetration, (unsigned long long)__double_as_longlong(mp));
}

// SimulatorObject.getAABoundingBox (SimulatorObject.fs:58-71: a sphere's box has side = radius, quirk Q5;
// Box.worldAxisAlignedSize, Box.fs:69-85) as objExtentProvider hands it to the tree (BroadPhase.fs:21-31: size and
// MIN corner = centre - size / 2), from the body-major records
// (host + device: ps_batch_add_body checks a new body against the tree's space before it changes anything)
__host__ __device__ inline void body_extent6(const double* pos, const double* rot, const double* dims, bool sphere, int aabb_mode, double* o)
{
    // scalar on purpose (psm:: routines are device functions): mulMV's sums written out, `+ 0.0` = the zero position of
    // toWorldCoordinates with the orientation only (Box.fs:69-85), scale(1/2, size) for the min corner (BroadPhase.fs:21-31)
    double sx, sy, sz;
    if (sphere) {
        sx = sy = sz = (aabb_mode == 1) ? 2.0 * dims[0] : dims[0];  // Box.createCube radius (Q5); 1 = conservative, side 2r
    } else {
        const double h[3] = {dims[0] / 2.0, dims[1] / 2.0, dims[2] / 2.0};
        double s[3];
        for (int r = 0; r < 3; r++) {
            double acc = 0.0;
            for (int c = 0; c < 3; c++) {  // column c of R scaled by the half size: R * (h_c e_c) + 0
                const double vx = c == 0 ? h[0] : 0.0, vy = c == 1 ? h[1] : 0.0, vz = c == 2 ? h[2] : 0.0;
                const double e = (((0.0 + rot[3 * r] * vx) + rot[3 * r + 1] * vy) + rot[3 * r + 2] * vz) + 0.0;
                acc = c == 0 ? fabs(e) : acc + fabs(e);
            }
            s[r] = acc * 2.0;
        }
        sx = s[0]; sy = s[1]; sz = s[2];
    }
    const double half = 1.0 / 2.0;
    o[0] = sx; o[1] = sy; o[2] = sz;
    o[3] = pos[0] - half * sx; o[4] = pos[1] - half * sy; o[5] = pos[2] - half * sz;
}
__global__ void extent_kernel(int n, const double* __restrict__ dyn, const double* __restrict__ par, double* __restrict__ ext, int aabb_mode)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const double* d = dyn + (size_t)g * kDynF;
    const double* p = par + (size_t)g * kParF;
    body_extent6(d, d + 6, p + 12, p[15] == 0.0, aabb_mode, ext + (size_t)g * 6);
}

// ---- helpers ---------------------------------------------------------------------------------------
template <class T>
void dfree(T*& p)
{
    if (p) cudaFree(p);
    p = nullptr;
}
void free_device(Batch* B)
{
    dfree(B->d_dyn); dfree(B->d_par); dfree(B->d_stage); dfree(B->d_off); dfree(B->d_stats); dfree(B->d_hitcnt); dfree(B->d_cost); dfree(B->d_perm); dfree(B->d_cur); dfree(B->d_pred); dfree(B->d_S); dfree(B->d_cull); dfree(B->d_row); dfree(B->d_lists); dfree(B->d_lists_off); dfree(B->d_lvl); dfree(B->d_last);
    dfree(B->d_body_world); dfree(B->d_pvalid); dfree(B->d_nhit); dfree(B->d_isstat); dfree(B->d_isbox); dfree(B->d_always_fused);
    dfree(B->d_K); dfree(B->d_maxl); dfree(B->d_wflags); dfree(B->d_bucket); dfree(B->d_bucket_cur); dfree(B->d_hit_cnt);
    dfree(B->d_hitlist); dfree(B->d_hitlist2); dfree(B->d_cpool_cur); dfree(B->d_entries); dfree(B->d_cpool); dfree(B->d_tmp); dfree(B->d_tmp_pen); dfree(B->d_surv); dfree(B->d_surv_cnt); dfree(B->d_nlev); dfree(B->d_nflag); dfree(B->d_stat_slots); dfree(B->d_cnt); dfree(B->d_abi_off);
    for (WideChunk& c : B->wchunks) {
        dfree(c.d_small);
        if (c.h_est) cudaFreeHost(c.h_est);
        if (c.est_ready) cudaEventDestroy(c.est_ready);
    }
    B->wchunks.clear();
    if (B->h_bucket) { cudaFreeHost(B->h_bucket); B->h_bucket = nullptr; }
    if (B->h_est) { cudaFreeHost(B->h_est); B->h_est = nullptr; }
    if (B->est_ready) { cudaEventDestroy(B->est_ready); B->est_ready = nullptr; }
    B->est_pending = B->est_valid = false;
    dfree(B->d_joff); dfree(B->d_ja); dfree(B->d_jb); dfree(B->d_jla); dfree(B->d_jlb); dfree(B->d_jlast);
    dfree(B->d_rec_pairs); dfree(B->d_rec_contacts); dfree(B->d_rec_counts); dfree(B->d_stats_sum);
    dfree(B->d_ext); dfree(B->d_cand); dfree(B->d_cand_off);
    if (B->h_ext) { cudaFreeHost(B->h_ext); B->h_ext = nullptr; }
}

int validate_bodies(const HostBodies& h)
{
    int nw = (int)h.off.size() - 1;
    for (int w = 0; w < nw; w++) {
        int N = h.off[w + 1] - h.off[w];
        if (N < 0) return fail(PS_ERR_INVALID_ARGUMENT, "world_offset must be non-decreasing (world %d)", w);
        if (N > 1024) return fail(PS_ERR_INVALID_ARGUMENT, "world %d has %d bodies; the per-world stepper takes at most 1024", w, N);
    }
    for (int i = 0; i < h.n(); i++) {
        if (h.shape[i] != 0 && h.shape[i] != 1) return fail(PS_ERR_INVALID_ARGUMENT, "body %d: shape must be 0 (sphere) or 1 (box)", i);
        for (int k = 0; k < 3; k++)
            if (h.dims[3 * i + k] < 0.0) return fail(PS_ERR_INVALID_ARGUMENT, "body %d: Must be positive (size)", i);  // throwIfNegative
        if (!(h.mass[i] > 0.0)) return fail(PS_ERR_INVALID_ARGUMENT, "body %d: Must be positive (mass)", i);  // Mass.Value rejects <= 0 (Base.fs:13-20); NaN too
        // Mass.Infinite: the reference inverts an infinite inertia to exactly 0 (RigidBody.fs:76); the parallel paths fold a
        // static body's angular momentum per pair, which is only the reference's value with that zero
        if (std::isinf(h.mass[i]) && (h.iinv[3 * i] != 0.0 || h.iinv[3 * i + 1] != 0.0 || h.iinv[3 * i + 2] != 0.0))
            return fail(PS_ERR_INVALID_ARGUMENT, "body %d: infinite mass needs inertia_inv = 0", i);
        if (h.e[i] < 0.0 || h.e[i] > 1.0) return fail(PS_ERR_INVALID_ARGUMENT, "body %d: elasticityCoeff Invalid value", i);
        if (h.muk[i] < 0.0 || h.mus[i] < 0.0) return fail(PS_ERR_INVALID_ARGUMENT, "body %d: staticFrictionCoeff Invalid value", i);
    }
    return PS_OK;
}

// the 16-double parameter record of body c of a compact HostBodies (slot 0 = 1.0/mass, IEEE division as on the reference's `v / mass`)
static void par_record(const HostBodies& h, int c, double* p)
{
    for (int k = 0; k < kParF; k++) p[k] = 0.0;
    p[0] = 1.0 / h.mass[c];
    for (int k = 0; k < 3; k++) {
        p[1 + k] = h.iinv[3 * c + k];
        p[4 + k] = h.force[3 * c + k];
        p[7 + k] = h.torque[3 * c + k];
        p[12 + k] = h.dims[3 * c + k];
    }
    p[10] = h.e[c];
    p[11] = h.muk[c];
    p[15] = (double)h.shape[c];
}

// upload parameters + (optionally) dynamic state of B->cur, (re)allocating device buffers
int upload(Batch* B, bool with_state)
{
    const HostBodies& h = B->cur;
    CK(cudaSetDevice(B->device));
    free_device(B);
    B->n_worlds = (int)h.off.size() - 1;
    B->n_bodies = h.n();
    B->max_n = 0;
    B->n_dynamic = 0;
    B->added.clear(); B->added_world.clear();
    {   // slots: every world gets room for a few more bodies than it has (PS_B200_SPARE overrides the rule below)
        int spare_min = 4;
        if (const char* e = getenv("PS_B200_SPARE")) spare_min = std::max(0, atoi(e));
        B->cap_off.assign((size_t)B->n_worlds + 1, 0);
        B->cnt.assign((size_t)B->n_worlds, 0);
        B->abi_off = h.off;
        for (int w = 0; w < B->n_worlds; w++) {
            const int N = h.off[w + 1] - h.off[w];
            const int cap = std::min(1024, N + std::max(spare_min, N / 16));
            B->cnt[w] = N;
            B->cap_off[w + 1] = B->cap_off[w] + std::max(cap, N);
            B->max_n = std::max(B->max_n, std::max(cap, N));
        }
        B->n_slots = B->cap_off[B->n_worlds];
        B->slot_mass.assign((size_t)std::max(1, B->n_slots), 1.0);
        for (int w = 0; w < B->n_worlds; w++)
            for (int b = 0; b < B->cnt[w]; b++) B->slot_mass[B->cap_off[w] + b] = h.mass[h.off[w] + b];
    }
    for (int i = 0; i < B->n_bodies; i++) B->n_dynamic += std::isinf(h.mass[i]) ? 0 : 1;
    // lanes per world (tile): a wavefront level of a 128-body pile holds ~10 pairs, so 8 lanes per world and 4
    // worlds per warp keep the lanes busy; with few worlds wider tiles give the SMs more warps to overlap
    B->threads = 8;
    B->minb = PS_MINB32;
    if ((long long)B->n_worlds * 4 / 32 >= 148 * 12) B->threads = 4;
    {   // Latency regime: every one-warp CTA of the batch is resident at 8 per SM -> the 255-register build, and the
        // widest tile (up to 16 lanes per world) that still keeps the batch in that single wave.  Measured (B200, ms per
        // step, 255 registers): C2 4096 worlds T=4 2.38, T=8 1.62, T=16 (two waves) 1.78; C4 8192 worlds T=4 2.32, T=8
        // (two waves) 2.65; C2 1024 worlds T=16 0.77, T=32 0.80.
        const long long wave = 148LL * 8;
        int t = 0;
        if ((long long)B->n_worlds * 16 / 32 <= wave) t = 16;
        else if ((long long)B->n_worlds * 8 / 32 <= wave) t = 8;
        else if ((long long)B->n_worlds * 4 / 32 <= wave) t = 4;
        if (t) { B->threads = t; B->minb = 8; }
        if ((long long)B->n_worlds * 16 / 32 < 148) B->threads = 32;  // a handful of worlds: all lanes to each of them
    }
    if (const char* e = getenv("PS_B200_TILE")) { int t = atoi(e); if (t == 4 || t == 8 || t == 16 || t == 32) B->threads = t; }  // tuning knob
    if (const char* e = getenv("PS_B200_MINB")) { int m = atoi(e); if (m == 8 || m == PS_MINB32) B->minb = m; }              // tuning knob
    B->smem = smem_bytes(B->max_n) * (size_t)(32 / B->threads);
    int dev_max = 0;
    CK(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, B->device));
    if ((long long)B->smem > dev_max)
        return fail(PS_ERR_INVALID_ARGUMENT, "largest world (%d bodies) needs %zu B of shared memory; device allows %d", B->max_n, B->smem, dev_max);
    {   // the kernel needs little shared memory: reserve what PS_RESIDENT CTAs per SM need and leave the rest of the
        // unified array to L1 (local stack of the per-pair routines)
        int resident = B->minb == 8 ? 8 : 12;
        if (const char* e = getenv("PS_B200_RESIDENT")) resident = std::max(1, atoi(e));
        size_t need = (size_t)resident * (B->smem + 1024 + 256);
        int pct = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024) + 1);
        if (const char* e = getenv("PS_B200_CARVEOUT")) pct = atoi(e);
        const int dyn = (int)std::max<size_t>(B->smem, 1024);
#define PS_SET_ATTRS(T_, M_)                                                                                  \
        CK(cudaFuncSetAttribute(step_kernel<T_, M_>, cudaFuncAttributePreferredSharedMemoryCarveout, pct)); \
        CK(cudaFuncSetAttribute(step_kernel<T_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        PS_SET_ATTRS(4, PS_MINB32) PS_SET_ATTRS(8, PS_MINB32) PS_SET_ATTRS(16, PS_MINB32) PS_SET_ATTRS(32, PS_MINB32)
        PS_SET_ATTRS(4, 8) PS_SET_ATTRS(8, 8) PS_SET_ATTRS(16, 8) PS_SET_ATTRS(32, 8)
#undef PS_SET_ATTRS
    }

    size_t nb = (size_t)std::max(1, B->n_slots);  // per-body device arrays are per SLOT
    CK(cudaMalloc(&B->d_dyn, nb * kDynF * sizeof(double)));
    CK(cudaMemsetAsync(B->d_dyn, 0, nb * kDynF * sizeof(double), B->stream));
    CK(cudaMalloc(&B->d_cnt, (size_t)std::max(1, B->n_worlds) * sizeof(int)));
    CK(cudaMalloc(&B->d_abi_off, (size_t)(B->n_worlds + 1) * sizeof(int)));
    CK(cudaMalloc(&B->d_par, nb * kParF * sizeof(double)));
    CK(cudaMalloc(&B->d_stage, nb * kDynF * sizeof(double)));
    CK(cudaMalloc(&B->d_off, (size_t)(B->n_worlds + 1) * sizeof(int)));
    CK(cudaMalloc(&B->d_stats, (size_t)std::max(1, B->n_worlds) * sizeof(WorldStats)));
    CK(cudaMalloc(&B->d_stats_sum, sizeof(ps_stats)));
    CK(cudaMalloc(&B->d_cur, nb * kDynF * sizeof(double)));
    CK(cudaMalloc(&B->d_pred, nb * kDynF * sizeof(double)));
    CK(cudaMalloc(&B->d_S, nb * 3 * kMaxStatics * sizeof(double)));
    {   // per-step lists of every world: 9 bytes per pair of the world
        std::vector<long long> loff(B->n_worlds + 1, 0);
        for (int w = 0; w < B->n_worlds; w++) {
            long long N = B->cap_off[w + 1] - B->cap_off[w], np = N * (N - 1) / 2;  // room for the world at its capacity
            loff[w + 1] = loff[w] + ((np * 9 + 15) / 16) * 16;
        }
        CK(cudaMalloc(&B->d_lists, (size_t)std::max<long long>(loff[B->n_worlds], 16)));
        CK(cudaMalloc(&B->d_lists_off, loff.size() * sizeof(long long)));
        CK(cudaMemcpy(B->d_lists_off, loff.data(), loff.size() * sizeof(long long), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&B->d_cull, nb * kCullF * sizeof(double)));
        CK(cudaMalloc(&B->d_row, (nb + (size_t)B->n_worlds + 1) * sizeof(int)));
        CK(cudaMalloc(&B->d_lvl, (size_t)std::max(1, B->n_worlds) * (kMaxLevels + 2) * sizeof(int)));
        CK(cudaMalloc(&B->d_last, nb * sizeof(unsigned short)));
    }
    {   // batch-wide path: per-body flags, per-world tables, the (level, kind)-sorted pair list
        std::vector<int> bw(nb, 0);
        std::vector<unsigned char> isstat(nb, 0), isbox(nb, 0), af(std::max(1, B->n_worlds), 0);
        for (int w = 0; w < B->n_worlds; w++) {
            int ns = 0;
            for (int g = B->cap_off[w]; g < B->cap_off[w + 1]; g++) bw[g] = w;  // spare slots know their world too
            for (int b = 0; b < B->cnt[w]; b++) {
                const int g = B->cap_off[w] + b, c = h.off[w] + b;
                isbox[g] = h.shape[c] != 0;
                if (std::isinf(h.mass[c])) { ns++; isstat[g] = (unsigned char)std::min(ns, kMaxStatics); }
            }
            af[w] = ns > kMaxStatics;
        }
        size_t nw = (size_t)std::max(1, B->n_worlds);
        CK(cudaMalloc(&B->d_body_world, nb * sizeof(int)));
        CK(cudaMalloc(&B->d_pvalid, nb)); CK(cudaMalloc(&B->d_nhit, nb)); CK(cudaMalloc(&B->d_isstat, nb)); CK(cudaMalloc(&B->d_isbox, nb));
        CK(cudaMalloc(&B->d_always_fused, nw));
        CK(cudaMalloc(&B->d_K, nw * sizeof(int))); CK(cudaMalloc(&B->d_maxl, nw * sizeof(int))); CK(cudaMalloc(&B->d_wflags, nw * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_bucket, (psw::kBuckets + 1) * sizeof(unsigned))); CK(cudaMalloc(&B->d_bucket_cur, psw::kBuckets * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_hit_cnt, (size_t)(psw::kWideMaxLevels + 2) * psw::kHitSlots * sizeof(unsigned))); CK(cudaMalloc(&B->d_cpool_cur, (psw::kWideMaxLevels + 2) * sizeof(unsigned)));
        B->entries_cap = (unsigned)std::min<size_t>(16 * nb + 1024, 400u << 20);
        B->cpool_cap = (unsigned)std::min<size_t>(4 * nb + 4096, 200u << 20);
        CK(cudaMalloc(&B->d_entries, (size_t)B->entries_cap * sizeof(psw::Entry)));
        CK(cudaMalloc(&B->d_hitlist, (size_t)B->entries_cap * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_hitlist2, (size_t)B->entries_cap * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_cpool, (size_t)B->cpool_cap * sizeof(Contact)));
        CK(cudaMallocHost(&B->h_bucket, (psw::kBuckets + 1) * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_tmp, nw * 6 * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_surv, (size_t)B->entries_cap * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_surv_cnt, (psw::kWideMaxLevels + 2) * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_tmp_pen, nw * sizeof(unsigned long long)));
        CK(cudaMalloc(&B->d_nlev, sizeof(int)));
        CK(cudaMalloc(&B->d_nflag, sizeof(int)));
        CK(cudaMallocHost(&B->h_est, (size_t)(4 * psw::kBuckets + 8) * sizeof(unsigned)));
        CK(cudaEventCreateWithFlags(&B->est_ready, cudaEventDisableTiming));
        {   // the rounds of a step run in one cooperative launch whose CTAs must all be resident
            int per_sm = 0, sms = 0, coop = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, psw::wide_rounds, 128, 0));
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, B->device));
            CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, B->device));
            B->sm_count = sms;
            B->rounds_grid = per_sm * sms;
            B->wide_device_rounds = coop != 0 && B->rounds_grid > 0;
            if (const char* e = getenv("PS_B200_WIDE_HOST_ROUNDS")) { if (atoi(e) != 0) B->wide_device_rounds = false; }  // tuning/profiling: one launch per stage and round, driven by the host
            // teams of 8 lanes where a thread per pair would leave lanes idle anyway: survivors x 8 <= resident lanes
            B->team_max = (unsigned)((long long)B->rounds_grid * 128 / 8);
            if (const char* e = getenv("PS_B200_TEAM_MAX")) B->team_max = (unsigned)std::max(0, atoi(e));  // tuning knob
        }
        CK(cudaMemcpy(B->d_body_world, bw.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_isstat, isstat.data(), nb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_isbox, isbox.data(), nb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_always_fused, af.data(), nw, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(psw::wide_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psw::lvl_smem_bytes(B->max_n)));
        CK(cudaFuncSetAttribute(psw::wide_near_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psw::near_smem_bytes(B->max_n)));
        {   // the slots the static fold visits (a static body added later goes through a re-layout, which rebuilds this list)
            std::vector<int> ss;
            for (int w = 0; w < B->n_worlds; w++)
                for (int b = 0; b < B->cnt[w]; b++)
                    if (std::isinf(B->slot_mass[B->cap_off[w] + b])) ss.push_back(B->cap_off[w] + b);
            B->n_stat_slots = (int)ss.size();
            B->h_stat_slots = ss;
            CK(cudaMalloc(&B->d_stat_slots, std::max<size_t>(ss.size(), 1) * sizeof(int)));
            if (!ss.empty()) CK(cudaMemcpy(B->d_stat_slots, ss.data(), ss.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        CK(cudaFuncSetAttribute(psw::wide_predict, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psw::predict_smem_bytes()));
        CK(cudaFuncSetAttribute(psw::wide_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psw::kTileBytes));
        // the wide path pays one pass through memory per stage and ~3 launches per level: it wins when the batch fills the
        // GPU per round, and earlier with boxes (a box pair costs the fused kernel a whole tile's pass).  Measured, ms per
        // settled step fused / wide: C2 (spheres) 8192 worlds 2.87 / 3.21, 16384 worlds 5.79 / 4.52; C3 (65 boxes + 62
        // spheres per world) 1024 worlds 4.35 / 5.38, 2048 worlds 8.35 / 5.94, 4096 worlds 13.87 / 7.28.
        size_t n_box = 0, n_live = 0;
        for (int w = 0; w < B->n_worlds; w++)
            for (int b = 0; b < B->cnt[w]; b++) { n_live++; n_box += isbox[B->cap_off[w] + b]; }
        B->wide = !B->cfg.exact_all_pairs && !(B->cfg.restore_joints && !B->joints.a.empty()) && (n_live - n_box) + 6 * n_box >= 700000;
        if (const char* e = getenv("PS_B200_PATH")) { if (!strcmp(e, "fused")) B->wide = false; else if (!strcmp(e, "wide")) B->wide = !B->cfg.exact_all_pairs && !(B->cfg.restore_joints && !B->joints.a.empty()); }
    }
    CK(cudaMalloc(&B->d_hitcnt, nb));
    CK(cudaMemsetAsync(B->d_hitcnt, 0, nb, B->stream));
    {   // segments of worlds: the chunks ps_batch_step_host pipelines, and the unit inside which worlds are re-ordered
        int chunks = 8;  // measured, C2 4096 worlds, e2e body-steps/s: 1 chunk 0.875e8, 2 1.00e8, 4 1.04e8, 8 1.08e8
        if (const char* e = getenv("PS_B200_HOST_CHUNKS")) chunks = std::max(1, std::min(64, atoi(e)));  // tuning knob
        const int wpc = 32 / B->threads;
        int per = (std::max(1, B->n_worlds) + chunks - 1) / chunks;
        B->seg = (per + wpc - 1) / wpc * wpc;
        CK(cudaMalloc(&B->d_cost, (size_t)std::max(1, B->n_worlds) * sizeof(unsigned short)));
        CK(cudaMalloc(&B->d_perm, (size_t)std::max(1, B->n_worlds) * sizeof(int)));
        CK(cudaMemsetAsync(B->d_cost, 0, (size_t)std::max(1, B->n_worlds) * sizeof(unsigned short), B->stream));
        B->cost_valid = B->perm_valid = false;
        B->steps_since_sort = 1 << 20;
    }
    CK(cudaMemsetAsync(B->d_stats, 0, (size_t)std::max(1, B->n_worlds) * sizeof(WorldStats), B->stream));
    CK(cudaMemcpyAsync(B->d_off, B->cap_off.data(), (size_t)(B->n_worlds + 1) * sizeof(int), cudaMemcpyHostToDevice, B->stream));
    CK(cudaMemcpyAsync(B->d_abi_off, B->abi_off.data(), (size_t)(B->n_worlds + 1) * sizeof(int), cudaMemcpyHostToDevice, B->stream));
    if (B->n_worlds > 0) CK(cudaMemcpyAsync(B->d_cnt, B->cnt.data(), (size_t)B->n_worlds * sizeof(int), cudaMemcpyHostToDevice, B->stream));

    // parameters -> body-major records (slot 0 = 1.0/mass, IEEE division as on the reference's `v / mass`)
    std::vector<double> par(nb * kParF, 0.0);
    for (int w = 0; w < B->n_worlds; w++)
        for (int b = 0; b < B->cnt[w]; b++) par_record(h, h.off[w] + b, par.data() + (size_t)(B->cap_off[w] + b) * kParF);
    CK(cudaMemcpyAsync(B->d_par, par.data(), nb * kParF * sizeof(double), cudaMemcpyHostToDevice, B->stream));
    CK(cudaStreamSynchronize(B->stream));  // `par` is a local

    // joints grouped per world
    const HostJoints& J = B->joints;
    int nj = (int)J.a.size();
    if (nj > 0) {
        std::vector<int> order(nj);
        for (int q = 0; q < nj; q++) order[q] = q;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return J.world[x] < J.world[y]; });
        std::vector<int> joff(B->n_worlds + 1, 0), ja(nj), jb(nj);
        std::vector<double> la(3 * nj), lb(3 * nj);
        std::vector<unsigned char> last(nj, 0);
        for (int q = 0; q < nj; q++) joff[J.world[order[q]] + 1]++;
        for (int w = 0; w < B->n_worlds; w++) {
            joff[w + 1] += joff[w];
        }
        for (int q = 0; q < nj; q++) {
            int s = order[q];
            ja[q] = J.a[s]; jb[q] = J.b[s];
            for (int k = 0; k < 3; k++) { la[3 * q + k] = J.la[3 * s + k]; lb[3 * q + k] = J.lb[3 * s + k]; }
        }
        // write-back is in joint-list order, the last writer wins (JointsRestorer.fs:18-26)
        for (int w = 0; w < B->n_worlds; w++) {
            std::vector<int> seen(h.off[w + 1] - h.off[w], 0);
            for (int q = joff[w + 1] - 1; q >= joff[w]; q--) {
                // targets of the two results: lower id, higher id (Q15); within a joint the second write comes last
                const int t1 = std::min(ja[q], jb[q]), t2 = std::max(ja[q], jb[q]);
                if (t1 != t2 && !seen[t2]) { last[q] |= 2; seen[t2] = 1; }
                if (!seen[t1]) { last[q] |= 1; seen[t1] = 1; }
            }
        }
        CK(cudaMalloc(&B->d_joff, joff.size() * sizeof(int)));
        CK(cudaMalloc(&B->d_ja, nj * sizeof(int)));
        CK(cudaMalloc(&B->d_jb, nj * sizeof(int)));
        CK(cudaMalloc(&B->d_jla, 3 * nj * sizeof(double)));
        CK(cudaMalloc(&B->d_jlb, 3 * nj * sizeof(double)));
        CK(cudaMalloc(&B->d_jlast, nj));
        CK(cudaMemcpy(B->d_joff, joff.data(), joff.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_ja, ja.data(), nj * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_jb, jb.data(), nj * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_jla, la.data(), 3 * nj * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_jlb, lb.data(), 3 * nj * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B->d_jlast, last.data(), nj, cudaMemcpyHostToDevice));
    }
    if (B->cfg.record_contacts) {
        B->rec_pair_cap = std::max(16, 6 * B->max_n);
        B->rec_contact_cap = std::max(64, 24 * B->max_n);
        size_t nw = (size_t)std::max(1, B->n_worlds);
        CK(cudaMalloc(&B->d_rec_pairs, nw * B->rec_pair_cap * sizeof(RecPair)));
        CK(cudaMalloc(&B->d_rec_contacts, nw * B->rec_contact_cap * sizeof(RecContact)));
        CK(cudaMalloc(&B->d_rec_counts, nw * 2 * sizeof(int)));
        CK(cudaMemset(B->d_rec_counts, 0, nw * 2 * sizeof(int)));
    }
    B->tree_mode = B->cfg.broadphase_kind == 1;
    if (B->tree_mode) {
        B->wide = false;  // the batch-wide path builds its lists from the Dummy order
        size_t np = 0;
        for (int w = 0; w < B->n_worlds; w++) { size_t N = (size_t)(B->cap_off[w + 1] - B->cap_off[w]); np += N * (N > 0 ? N - 1 : 0) / 2; }
        CK(cudaMalloc(&B->d_ext, nb * 6 * sizeof(double)));
        CK(cudaMallocHost(&B->h_ext, nb * 6 * sizeof(double)));
        CK(cudaMalloc(&B->d_cand, std::max<size_t>(np, 1) * sizeof(unsigned)));
        CK(cudaMalloc(&B->d_cand_off, (size_t)(B->n_worlds + 1) * sizeof(int)));
    }
    if (with_state && B->n_bodies > 0) {
        int rc = ps_batch_set_state(B, 0, B->n_worlds, h.pos.data(), h.vel.data(), h.rot.data(), h.L.data());
        if (rc) return rc;
    }
    return PS_OK;
}

// ---- SpatialTree candidate order ------------------------------------------------------------------------------------
// extents of every body at the state the device holds now -> B->h_ext
int tree_pull_extents(Batch* B, cudaStream_t st)
{
    if (B->n_slots == 0) return PS_OK;
    extent_kernel<<<(B->n_slots + 127) / 128, 128, 0, st>>>(B->n_slots, B->d_dyn, B->d_par, B->d_ext, B->cfg.aabb_mode);  // (spare slots: unread)
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(B->h_ext, B->d_ext, (size_t)B->n_slots * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PS_OK;
}
static pst::Ext<3> ext_of(const Batch* B, int g)
{
    pst::Ext<3> e;
    const double* q = B->h_ext + (size_t)g * 6;
    for (int k = 0; k < 3; k++) { e.size[k] = q[k]; e.pos[k] = q[3 + k]; }
    return e;
}
// BroadPhase.init (BroadPhase.fs:72-81, 41-60): every id of the world, ascending; an object outside the tree's space
// makes the reference throw ArgumentException (SpatialTree.fs:171-172)
int tree_init_world(Batch* B, int w)
{
    const int o = B->cap_off[w], N = B->cnt[w];
    pst::Tree<3>& t = B->trees[w];
    if (!t.init(B->cfg.tree_leaf_capacity, B->cfg.tree_max_depth, B->cfg.tree_min, B->cfg.tree_max))
        return fail(PS_ERR_INVALID_ARGUMENT, "SpatialTree: must be positive (maxLeafObjects / maxDepth)");
    for (int id = 0; id < N; id++)
        if (!t.insert(id, [&](int x) { return ext_of(B, o + x); }))
            return fail(PS_ERR_INVALID_ARGUMENT, "world %d: object %d out of space bounds", w, id);
    return PS_OK;
}
int tree_init_all(Batch* B)
{
    int rc = tree_pull_extents(B, B->stream);
    if (rc) return rc;
    B->trees.assign((size_t)B->n_worlds, pst::Tree<3>());
    for (int w = 0; w < B->n_worlds; w++) {
        rc = tree_init_world(B, w);
        if (rc) return rc;
    }
    B->h_cand.clear();
    B->h_cand_off.assign((size_t)B->n_worlds + 1, 0);
    return PS_OK;
}
// BroadPhase.update (BroadPhase.fs:100-127) for every world at the step-start state: dynamic ids leave every leaf and
// come back in ascending order (tryInsert: an object outside the space is silently absent this step), then the
// candidate list in leaf order; uploaded for the step kernel
int tree_update_and_upload(Batch* B, cudaStream_t st)
{
    int rc = tree_pull_extents(B, st);
    if (rc) return rc;
    B->h_cand.clear();
    B->h_cand_off.assign((size_t)B->n_worlds + 1, 0);
    std::vector<uint32_t> one;
    std::vector<unsigned char> stat;
    for (int w = 0; w < B->n_worlds; w++) {
        const int o = B->cap_off[w], N = B->cnt[w];
        stat.assign((size_t)N, 0);
        for (int b = 0; b < N; b++) stat[b] = std::isinf(B->slot_mass[o + b]) ? 1 : 0;
        pst::Tree<3>& t = B->trees[w];
        t.remove_if([&](int id) { return !stat[id]; });
        for (int id = 0; id < N; id++)
            if (!stat[id]) t.insert(id, [&](int x) { return ext_of(B, o + x); });
        pst::candidates(t, stat.data(), one);
        B->h_cand.insert(B->h_cand.end(), one.begin(), one.end());
        B->h_cand_off[w + 1] = (int)B->h_cand.size();
    }
    if (!B->h_cand.empty())
        CK(cudaMemcpyAsync(B->d_cand, B->h_cand.data(), B->h_cand.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B->d_cand_off, B->h_cand_off.data(), B->h_cand_off.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // the host vectors are reused by the next step
    return PS_OK;
}

// compact description of the batch as it stands NOW: `cur` (as of the last upload) + every body added since, spliced in
// at the end of its world, + the live dynamic state from the device
int merged_description(Batch* B, HostBodies& out)
{
    const HostBodies& h = B->cur;
    const int nw = B->n_worlds;
    std::vector<std::vector<int>> add_of((size_t)nw);
    for (size_t q = 0; q < B->added.size(); q++) add_of[(size_t)B->added_world[q]].push_back((int)q);
    out = HostBodies();
    out.off.assign((size_t)nw + 1, 0);
    auto app = [](std::vector<double>& dst, const std::vector<double>& src, size_t first, size_t count, int per) {
        dst.insert(dst.end(), src.begin() + (ptrdiff_t)(first * per), src.begin() + (ptrdiff_t)((first + count) * per));
    };
    for (int w = 0; w < nw; w++) {
        const size_t f = (size_t)h.off[w], c = (size_t)(h.off[w + 1] - h.off[w]);
        out.shape.insert(out.shape.end(), h.shape.begin() + (ptrdiff_t)f, h.shape.begin() + (ptrdiff_t)(f + c));
        app(out.dims, h.dims, f, c, 3); app(out.mass, h.mass, f, c, 1); app(out.iinv, h.iinv, f, c, 3); app(out.force, h.force, f, c, 3);
        app(out.torque, h.torque, f, c, 3); app(out.e, h.e, f, c, 1); app(out.muk, h.muk, f, c, 1); app(out.mus, h.mus, f, c, 1);
        for (int q : add_of[(size_t)w]) {
            const HostBodies& a = B->added[(size_t)q];
            out.shape.push_back(a.shape[0]);
            app(out.dims, a.dims, 0, 1, 3); app(out.mass, a.mass, 0, 1, 1); app(out.iinv, a.iinv, 0, 1, 3); app(out.force, a.force, 0, 1, 3);
            app(out.torque, a.torque, 0, 1, 3); app(out.e, a.e, 0, 1, 1); app(out.muk, a.muk, 0, 1, 1); app(out.mus, a.mus, 0, 1, 1);
        }
        out.off[(size_t)w + 1] = (int32_t)out.mass.size();
    }
    const size_t n = out.mass.size();
    if ((int)n != B->n_bodies) return fail(PS_ERR_INVALID_OPERATION, "internal: %zu merged bodies, %d on the device", n, B->n_bodies);
    out.pos.assign(3 * n, 0.0); out.vel.assign(3 * n, 0.0); out.rot.assign(9 * n, 0.0); out.L.assign(3 * n, 0.0);
    if (n == 0) return PS_OK;
    return ps_batch_get_state(B, 0, nw, out.pos.data(), out.vel.data(), out.rot.data(), out.L.data());
}

void copy_view(HostBodies& h, const ps_bodies_view* v)
{
    int nw = v->n_worlds;
    h.off.assign(v->world_offset, v->world_offset + nw + 1);
    int n = h.off[nw];
    auto cp = [&](std::vector<double>& dst, const double* src, int per) { dst.assign(src, src + (size_t)per * n); };
    h.shape.assign(v->shape, v->shape + n);
    cp(h.dims, v->dims, 3); cp(h.mass, v->mass, 1); cp(h.iinv, v->inertia_inv, 3);
    cp(h.pos, v->pos, 3); cp(h.vel, v->vel, 3); cp(h.rot, v->rot, 9); cp(h.L, v->ang_mom, 3);
    cp(h.force, v->force, 3); cp(h.torque, v->torque, 3); cp(h.e, v->elasticity, 1); cp(h.muk, v->mu_kinetic, 1);
    if (v->mu_static) cp(h.mus, v->mu_static, 1); else h.mus.assign(n, 0.0);
}

int check_view(const ps_bodies_view* v)
{
    if (!v) return fail(PS_ERR_INVALID_ARGUMENT, "bodies view is null");
    if (v->n_worlds < 0 || !v->world_offset) return fail(PS_ERR_INVALID_ARGUMENT, "n_worlds/world_offset invalid");
    if (v->world_offset[0] != 0) return fail(PS_ERR_INVALID_ARGUMENT, "world_offset[0] must be 0");
    for (int w = 0; w < v->n_worlds; w++)
        if (v->world_offset[w + 1] < v->world_offset[w]) return fail(PS_ERR_INVALID_ARGUMENT, "world_offset must be non-decreasing");
    if (v->world_offset[v->n_worlds] > 0 &&
        (!v->shape || !v->dims || !v->mass || !v->inertia_inv || !v->pos || !v->vel || !v->rot || !v->ang_mom || !v->force ||
         !v->torque || !v->elasticity || !v->mu_kinetic))
        return fail(PS_ERR_INVALID_ARGUMENT, "a body array is null");
    return PS_OK;
}

}  // namespace

// ====================================================================================================
extern "C" {

#define PS_REFUSE_IF_BROKEN(B) do { if ((B)->broken) return fail(PS_ERR_INVALID_OPERATION, "the batch lost its device state in a failed ps_batch_add_body; ps_batch_reset rebuilds it"); } while (0)

int ps_version(void) { return 100; }
const char* ps_last_error(void) { return g_err.c_str(); }

int ps_init(int32_t device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(PS_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (device >= n) return fail(PS_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", device, n);
    if (device >= 0) {
        CK(cudaSetDevice(device));
        g_device = device;
    } else {
        CK(cudaGetDevice(&g_device));
    }
    return PS_OK;
}

int ps_batch_create(void** handle, const ps_bodies_view* bodies, const ps_joints_view* joints, const ps_step_config* config)
{
    if (!handle || !config) return fail(PS_ERR_INVALID_ARGUMENT, "handle/config is null");
    *handle = nullptr;
    int rc = check_view(bodies);
    if (rc) return rc;
    if (config->solver_iterations < 0) return fail(PS_ERR_INVALID_ARGUMENT, "solver_iterations must be >= 0");
    if (config->broadphase_kind != 0 && config->broadphase_kind != 1) return fail(PS_ERR_INVALID_ARGUMENT, "broadphase_kind must be 0 (Dummy) or 1 (SpatialTree)");
    if (config->aabb_mode != 0 && config->aabb_mode != 1) return fail(PS_ERR_INVALID_ARGUMENT, "aabb_mode must be 0 (reference) or 1 (conservative)");
    if (!(config->dt == config->dt)) return fail(PS_ERR_INVALID_ARGUMENT, "dt is NaN");
    if (g_device < 0) {
        rc = ps_init(-1);
        if (rc) return rc;
    }
    Batch* B = new Batch();
    B->device = g_device;
    B->cfg = *config;
    copy_view(B->init, bodies);
    rc = validate_bodies(B->init);
    if (rc) { delete B; return rc; }
    if (joints && joints->n_joints > 0) {
        int nj = joints->n_joints;
        B->joints.world.assign(joints->world, joints->world + nj);
        B->joints.a.assign(joints->body_a, joints->body_a + nj);
        B->joints.b.assign(joints->body_b, joints->body_b + nj);
        B->joints.la.assign(joints->anchor_a, joints->anchor_a + 3 * nj);
        B->joints.lb.assign(joints->anchor_b, joints->anchor_b + 3 * nj);
        for (int q = 0; q < nj; q++) {
            int w = B->joints.world[q];
            if (w < 0 || w >= bodies->n_worlds) { delete B; return fail(PS_ERR_OUT_OF_RANGE, "joint %d: world %d", q, w); }
            int N = B->init.off[w + 1] - B->init.off[w];
            int a = B->joints.a[q], b = B->joints.b[q];
            if (a < 0 || b < 0 || a >= N || b >= N) { delete B; return fail(PS_ERR_OUT_OF_RANGE, "joint %d: body id out of range", q); }
            // a > b is taken as given: changePhysicalObjects' pairing of id-sorted objects with joint-ordered results (quirk
            // Q15: the two bodies swap states) is reproduced by the step kernel.  a == b: the reference zips ONE object
            // with TWO results and List.zip throws ArgumentException (SimulatorState.fs:153-160).
            if (a == b) { delete B; return fail(PS_ERR_INVALID_ARGUMENT, "joint %d: a body cannot be joined to itself (List.zip: lists of different lengths)", q); }
        }
    }
    B->cur = B->init;
    if (cudaSetDevice(B->device) != cudaSuccess || cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete B;
        return fail(PS_ERR_CUDA, "cannot create a stream on device %d", g_device);
    }
    rc = upload(B, true);
    if (!rc && B->tree_mode) rc = tree_init_all(B);
    if (rc) {
        free_device(B);
        cudaStreamDestroy(B->stream);
        delete B;
        return rc;
    }
    *handle = B;
    return PS_OK;
}

int ps_batch_destroy(void* handle)
{
    if (!handle) return PS_OK;
    Batch* B = (Batch*)handle;
    cudaSetDevice(B->device);
    cudaStreamSynchronize(B->stream);
    if (getenv("PS_B200_WIDE_PROF") && B->prof_steps > 0)
        fprintf(stderr, "[ps_b200 wide profile] steps %lld rounds/step %.1f | ms/step: setup %.3f (-) %.3f detect_stage1 %.3f (-) %.3f bb_full %.3f resolve %.3f tail %.3f fused %.3f\n",
                B->prof_steps, (double)B->prof_rounds / B->prof_steps, B->prof_ms[0] / B->prof_steps, B->prof_ms[1] / B->prof_steps, B->prof_ms[2] / B->prof_steps,
                B->prof_ms[3] / B->prof_steps, B->prof_ms[4] / B->prof_steps, B->prof_ms[5] / B->prof_steps, B->prof_ms[6] / B->prof_steps, B->prof_ms[7] / B->prof_steps);
#ifdef PS_PROF_CLK
    {
        unsigned long long sm[32], mx[32], cn[32];
        cudaMemcpyFromSymbol(sm, psp::g_clk_sum, sizeof(sm)); cudaMemcpyFromSymbol(mx, psp::g_clk_max, sizeof(mx)); cudaMemcpyFromSymbol(cn, psp::g_clk_cnt, sizeof(cn));
        const char* nm[32] = {"bb: boxes+face axes", "bb: edge axes", "bb: face setup", "bb: clipping", "bb: contacts", "bb: edge contact", "", "", "bbkernel: loads", "bbkernel: detect_bb total", "", "",
                              "resolve: loads", "resolve: contact loads", "resolve: solve", "resolve: write-back + 2 integrations"};
        static unsigned long long hs[32][32];
        cudaMemcpyFromSymbol(hs, psp::g_clk_hist, sizeof(hs));
        for (int i = 0; i < 16; i++)
            if (cn[i]) {
                fprintf(stderr, "[clk] %-38s n=%9llu mean %8.0f max %8llu cycles | log2 histogram:", nm[i], cn[i], (double)sm[i] / cn[i], mx[i]);
                for (int b = 8; b < 24; b++) fprintf(stderr, " 2^%d:%llu", b, hs[i][b]);
                fprintf(stderr, "\n");
            }
    }
#endif
    if (getenv("PS_B200_WIDE_PROF") && B->wide && B->d_tmp) {
        std::vector<unsigned> tm((size_t)B->n_worlds * 6);
        cudaMemcpy(tm.data(), B->d_tmp, tm.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
        unsigned long long lit = 0, hits = 0;
        for (int w = 0; w < B->n_worlds; w++) { lit += tm[(size_t)w * 6 + 2]; hits += ((*(unsigned long long*)&tm[(size_t)w * 6]) >> 20) & 0xfffffull; }
        fprintf(stderr, "[ps_b200 wide profile] last step: %llu of %llu colliding pairs were solved again by the literal routine\n", lit, hits);
    }
    if (getenv("PS_B200_WIDE_PROF") && !B->prof_round.empty()) {
        std::vector<unsigned> hc((size_t)(psw::kWideMaxLevels + 2) * psw::kHitSlots), sc(psw::kWideMaxLevels + 2);
        cudaMemcpy(hc.data(), B->d_hit_cnt, hc.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
        cudaMemcpy(sc.data(), B->d_surv_cnt, sc.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ps_b200 wide profile] last step, per round: pairs ss/sb/bb | bb survivors | hits ss/sb/bb | ms detect, bb, resolve\n");
        for (size_t r = 0; r * 6 < B->prof_round.size(); r++) {
            const float* q = &B->prof_round[r * 6];
            fprintf(stderr, "  round %3zu: %7.0f %7.0f %7.0f | %7u | %7u %7u %7u | %.3f %.3f %.3f\n", r + 1, q[0], q[1], q[2], sc[r + 1],
                    hc[(r + 1) * 8], hc[(r + 1) * 8 + 1], hc[(r + 1) * 8 + 2] + hc[(r + 1) * 8 + 3] + hc[(r + 1) * 8 + 4] + hc[(r + 1) * 8 + 5], q[3], q[4], q[5]);
        }
    }
    free_device(B);
    for (cudaStream_t s : B->chunk_streams) cudaStreamDestroy(s);
    for (cudaEvent_t e : B->chunk_done) cudaEventDestroy(e);
    if (B->chunk_start) cudaEventDestroy(B->chunk_start);
    if (B->ev_own) cudaEventDestroy(B->ev_own);
    if (B->ev_foreign) cudaEventDestroy(B->ev_foreign);
    cudaStreamDestroy(B->stream);
    delete B;
    return PS_OK;
}

static void fill_kargs(Batch* B, KArgs& a, int n_steps)
{
    a.dyn = B->d_dyn; a.par = B->d_par; a.off = B->d_off; a.cnt = B->d_cnt; a.stats = B->d_stats;
    a.n_worlds = B->n_worlds; a.n_steps = n_steps;
    a.sp.eps = B->cfg.epsilon; a.sp.baumgarte = B->cfg.baumgarte; a.sp.slop = B->cfg.allowed_penetration;
    a.sp.max_corr = B->cfg.max_correction_velocity; a.sp.dt = B->cfg.dt;
    a.sp.iterations = B->cfg.solver_iterations; a.sp.friction = B->cfg.enable_friction; a.sp.integrator = B->cfg.integrator_kind;
    a.exact_all_pairs = B->cfg.exact_all_pairs;
    a.restore_joints = B->cfg.restore_joints;
    a.margin_ncap = 16;
    a.margin_scale = 1.0;  // with margins verified after the fact (phase V): C2 2.0 -> 2.25 ms, 1.0 -> 2.11 ms, 0.7 -> 3.36 ms (10 % of world-steps restart); C3 20.1, 18.8, 29.7 (0.5)
    if (const char* e = getenv("PS_B200_MSCALE")) a.margin_scale = atof(e);  // tuning knob
    if (const char* e = getenv("PS_B200_NCAP")) a.margin_ncap = std::max(0, std::min(255, atoi(e)));  // tuning knob
    a.joff = B->d_joff; a.ja = B->d_ja; a.jb = B->d_jb; a.jla = B->d_jla; a.jlb = B->d_jlb; a.jlast = B->d_jlast;
    a.rec_pairs = B->d_rec_pairs; a.rec_contacts = B->d_rec_contacts; a.rec_counts = B->d_rec_counts;
    a.rec_pair_cap = B->rec_pair_cap; a.rec_contact_cap = B->rec_contact_cap;
    a.max_n = B->max_n;
    a.hitcnt = B->d_hitcnt; a.cur = B->d_cur; a.pred = B->d_pred; a.S = B->d_S;
    a.cull = B->d_cull; a.row = B->d_row; a.lists = B->d_lists; a.lists_off = B->d_lists_off; a.lvl = B->d_lvl; a.last = B->d_last;
    a.only_flagged = nullptr;
    a.redo_count = nullptr; a.redo_lo = 0; a.redo_hi = 0;
    a.world_base = 0;
    a.perm = nullptr;
    a.cost = B->d_cost;
    a.cand = B->tree_mode ? B->d_cand : nullptr;
    a.cand_off = B->tree_mode ? B->d_cand_off : nullptr;
}

// One block per segment: bitonic sort of (cost << 16 | index in segment) in shared memory -> perm.
__global__ void perm_sort_kernel(const unsigned short* cost, int* perm, int seg, int n_worlds, int npow2)
{
    extern __shared__ unsigned keys[];
    const int base = blockIdx.x * seg;
    const int cnt = min(seg, n_worlds - base);
    for (int k = threadIdx.x; k < npow2; k += blockDim.x) keys[k] = k < cnt ? ((unsigned)cost[base + k] << 16) | (unsigned)k : 0xffffffffu;
    __syncthreads();
    for (int size = 2; size <= npow2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int k = threadIdx.x; k < npow2; k += blockDim.x) {
                const int p = k ^ stride;
                if (p > k) {
                    const bool up = (k & size) == 0;
                    const unsigned x = keys[k], y = keys[p];
                    if ((x > y) == up) { keys[k] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) perm[base + k] = base + (int)(keys[k] & 0xffffu);
}

// Re-order the worlds inside their segments by the passes of their last step, every kResortSteps steps.
// Which warp a world runs in changes nothing about its results.
static const int* current_perm(Batch* B, int n_steps, cudaStream_t st)
{
    const int kResortSteps = 16;
    const bool kPermDefault = false;  // measured: C2 1.636 ms re-ordered vs 1.595 ms in creation order -> off
    const char* pe = getenv("PS_B200_PERM");
    if (!(pe ? atoi(pe) != 0 : kPermDefault)) return nullptr;
    if (B->wide || B->tree_mode || B->seg <= 0 || B->seg > 65536) return nullptr;
    int npow2 = 1;
    while (npow2 < B->seg) npow2 <<= 1;
    if ((size_t)npow2 * sizeof(unsigned) > 48 * 1024) return nullptr;
    if (B->cost_valid && B->steps_since_sort >= kResortSteps) {
        const int nseg = (B->n_worlds + B->seg - 1) / B->seg;
        perm_sort_kernel<<<nseg, 256, (size_t)npow2 * sizeof(unsigned), st>>>(B->d_cost, B->d_perm, B->seg, B->n_worlds, npow2);
        B->perm_valid = true;
        B->steps_since_sort = 0;
    }
    B->steps_since_sort += n_steps;
    B->cost_valid = true;  // (after the launch this call is about to make)
    return B->perm_valid ? B->d_perm : nullptr;
}

static void launch_fused(Batch* B, const KArgs& a, cudaStream_t st)
{
    const int wpc = 32 / B->threads;  // worlds per one-warp CTA
    const int grid = (a.n_worlds - a.world_base + wpc - 1) / wpc;  // (a.n_worlds is the end of the launched range)
    if (B->minb == 8) {
        if (B->threads == 4) step_kernel<4, 8><<<grid, 32, B->smem, st>>>(a);
        else if (B->threads == 8) step_kernel<8, 8><<<grid, 32, B->smem, st>>>(a);
        else if (B->threads == 16) step_kernel<16, 8><<<grid, 32, B->smem, st>>>(a);
        else step_kernel<32, 8><<<grid, 32, B->smem, st>>>(a);
    } else {
        if (B->threads == 4) step_kernel<4, PS_MINB32><<<grid, 32, B->smem, st>>>(a);
        else if (B->threads == 8) step_kernel<8, PS_MINB32><<<grid, 32, B->smem, st>>>(a);
        else if (B->threads == 16) step_kernel<16, PS_MINB32><<<grid, 32, B->smem, st>>>(a);
        else step_kernel<32, PS_MINB32><<<grid, 32, B->smem, st>>>(a);
    }
    B->launches++;
}

// one step of the batch-wide path (ps_wide.cuh)
static int launch_wide_step(Batch* B, const KArgs& ka, cudaStream_t st, WideChunk* ck = nullptr)
{
    using namespace psw;
    WArgs a{};
    a.n_worlds = B->n_worlds; a.n_bodies = B->n_slots; a.max_n = B->max_n; a.off = B->d_off; a.cnt = B->d_cnt; a.body_world = B->d_body_world;
    a.dyn = B->d_dyn; a.cur = B->d_cur; a.pred = B->d_pred; a.par = B->d_par; a.S = B->d_S; a.cull = B->d_cull;
    a.pvalid = B->d_pvalid; a.nhit = B->d_nhit; a.nprev = B->d_hitcnt; a.isstat = B->d_isstat; a.isbox = B->d_isbox;
    a.row = B->d_row; a.lists = B->d_lists; a.lists_off = B->d_lists_off;
    a.K = B->d_K; a.maxl = B->d_maxl; a.wflags = B->d_wflags; a.always_fused = B->d_always_fused;
    a.bucket = B->d_bucket; a.bucket_cur = B->d_bucket_cur; a.entries = B->d_entries; a.entries_cap = B->entries_cap;
    a.hit_cnt = B->d_hit_cnt; a.hitlist = B->d_hitlist; a.hitlist2 = B->d_hitlist2; a.cpool = B->d_cpool; a.cpool_cap = B->cpool_cap; a.cpool_cur = B->d_cpool_cur;
    a.stats = B->d_stats; a.tmp = B->d_tmp; a.tmp_pen = B->d_tmp_pen; a.surv = B->d_surv; a.surv_cnt = B->d_surv_cnt;
    a.rec_pairs = B->d_rec_pairs; a.rec_contacts = B->d_rec_contacts; a.rec_counts = B->d_rec_counts;
    a.rec_pair_cap = B->rec_pair_cap; a.rec_contact_cap = B->rec_contact_cap;
    a.nlev = B->d_nlev; a.nflag = B->d_nflag; a.team_max = B->team_max;
    a.sp = ka.sp; a.margin_ncap = ka.margin_ncap;
    a.margin_scale = ka.margin_scale;
    // the whole batch, or one chunk of ps_batch_step_host with its own small tables and its slice of the lists
    unsigned* h_est = B->h_est;
    cudaEvent_t est_ready = B->est_ready;
    bool *est_pending = &B->est_pending, *est_valid = &B->est_valid;
    std::vector<unsigned>* est = &B->est;
    const int* d_stat = B->d_stat_slots;
    int n_stat = B->n_stat_slots;
    a.w_begin = 0; a.w_end = B->n_worlds; a.g_begin = 0; a.g_end = B->n_slots;
    if (ck) {
        unsigned* p = ck->d_small;
        a.bucket = p; p += kBuckets + 1;
        a.bucket_cur = p; p += kBuckets;
        a.hit_cnt = p; p += (kWideMaxLevels + 2) * kHitSlots;
        a.cpool_cur = p; p += kWideMaxLevels + 2;
        a.surv_cnt = p; p += kWideMaxLevels + 2;
        a.nlev = reinterpret_cast<int*>(p); p += 1;
        a.nflag = reinterpret_cast<int*>(p);
        a.entries = B->d_entries + ck->e_off; a.entries_cap = ck->e_cap;
        a.hitlist = B->d_hitlist + ck->e_off; a.hitlist2 = B->d_hitlist2 + ck->e_off; a.surv = B->d_surv + ck->e_off;
        a.cpool = B->d_cpool + ck->c_off; a.cpool_cap = ck->c_cap;
        a.w_begin = ck->w0; a.w_end = ck->w1; a.g_begin = ck->g0; a.g_end = ck->g1;
        h_est = ck->h_est; est_ready = ck->est_ready; est_pending = &ck->est_pending; est_valid = &ck->est_valid; est = &ck->est;
        d_stat = B->d_stat_slots + ck->stat_first; n_stat = ck->stat_count;
    }
    const int nb = a.g_end - a.g_begin, nw = a.w_end - a.w_begin;
    const int gb = (nb + 127) / 128, gw = (nw + 63) / 64, gl = (nw + psw::kLvlWorlds - 1) / psw::kLvlWorlds;
    const bool prof = getenv("PS_B200_WIDE_PROF") != nullptr && !ck;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (prof) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); cudaEventRecord(pe0, st); }
    float lap_ms = 0.f;
    auto lap = [&](int cat) {  // profiling only: serialises the stream
        if (!prof) return;
        cudaEventRecord(pe1, st); cudaEventSynchronize(pe1);
        float ms = 0.f; cudaEventElapsedTime(&ms, pe0, pe1);
        B->prof_ms[cat] += ms;
        lap_ms = ms;
        cudaEventRecord(pe0, st);
    };
    if (prof) B->prof_round.clear();
    const int greset = (std::max(nw, kBuckets + 1) + 255) / 256;
    wide_reset<<<greset, 256, 0, st>>>(a);
    wide_predict<<<gb, psw::kTile, psw::predict_smem_bytes(), st>>>(a);
    wide_near_list<<<nw, psw::kNearT, psw::near_smem_bytes(B->max_n), st>>>(a);
    wide_levels<<<gl, psw::kLvlThreads, psw::lvl_smem_bytes(B->max_n), st>>>(a);
    wide_bucket_scan<<<1, 32, 0, st>>>(a);
    wide_scatter<<<gl, psw::kLvlThreads, 0, st>>>(a);
    B->launches += 6;
    if (B->wide_device_rounds && !prof) {
        // Device-counted rounds: no host round trip.  Grids are sized from the counts an earlier step left (if its
        // read-back has arrived; the host never waits for it), every kernel reads the real counts on the device and walks
        // its chunks with a grid stride; rounds beyond the ones enqueued here go to the cooperative catch-all.
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        CK(cudaStreamIsCapturing(st, &cap));
        const bool capturing = cap != cudaStreamCaptureStatusNone;
        if (!capturing && *est_pending && cudaEventQuery(est_ready) == cudaSuccess) {  // (an event query would invalidate a capture)
            est->assign(h_est, h_est + 4 * kBuckets + 8);
            *est_pending = false;
            *est_valid = true;
        }
        const unsigned dflt = (unsigned)std::max(1, B->sm_count) * 8;   // no estimate yet: a grid that fills the device, strides do the rest
        auto blocks = [&](unsigned n, unsigned per) { return (unsigned)(((unsigned long long)n * 9 / 8 + per - 1) / per) + 4u; };  // estimate + 12 %
        int r_launch = 48;
        static const unsigned big_g3 = getenv("PS_B200_RESOLVE_BIG_CTAS") ? (unsigned)atoi(getenv("PS_B200_RESOLVE_BIG_CTAS")) : 1200u;  // tuning knob (measured, C3: 4000 11.48, 1500 11.46, 900 11.45, 500 11.57 ms per step): CTAs from which a round's solve is the 168-register one-lane build
        static const bool paired = !(getenv("PS_B200_RESOLVE_PAIRED") && atoi(getenv("PS_B200_RESOLVE_PAIRED")) == 0);  // tuning knob
        const unsigned* hb = *est_valid ? est->data() : nullptr;
        const unsigned* hh = hb ? hb + kBuckets + 1 : nullptr;      // hit counts per (level, kind)
        const unsigned* hs = hb ? hb + 3 * kBuckets + 1 : nullptr;  // survivors per level
        if (hb) {
            int top = 0;
            for (int r = 1; r <= kWideMaxLevels; r++) if (hb[(r + 1) * 4] > hb[r * 4]) top = r;
            r_launch = std::min(kWideMaxLevels, top + 3);
        }
        if (const char* e = getenv("PS_B200_COOP_ALL")) { if (atoi(e) != 0) r_launch = 0; }  // tuning: every round in the cooperative kernel
        for (int r = 1; r <= r_launch; r++) {
            unsigned g1 = dflt, g2 = dflt, g2t = dflt, g3 = dflt, g3p = dflt, ns_est = ~0u;
            if (hb) {
                const unsigned c0 = hb[r * 4 + 1] - hb[r * 4], c1 = hb[r * 4 + 2] - hb[r * 4 + 1], c2 = hb[r * 4 + 3] - hb[r * 4 + 2];
                g1 = blocks(c0, 128) + blocks(c1, 128) + blocks(c2, 128);
                ns_est = hs[r];
                g2 = blocks(ns_est, 128); g2t = blocks(ns_est, 16);
                g3 = 0;
                for (int k = 0; k < 6; k++) g3 += blocks(hh[r * psw::kHitSlots + k], 128);
                g3p = g3;
                for (int k = psw::kPairedFrom; k < 6; k++) g3p += blocks(hh[r * psw::kHitSlots + k], 128);  // the large classes on two lanes each
            }
            wide_round_detect_d<<<g1, 128, 0, st>>>(a, r);
            if (ns_est <= B->team_max) wide_round_bb_team_d<<<g2t, 128, 0, st>>>(a, r);
            else wide_round_bb_d<<<g2, 128, 0, st>>>(a, r);
            if (hb && g3 > big_g3) wide_round_resolve_d<3, false><<<g3, 128, 0, st>>>(a, r);  // many colliding pairs: throughput bound
            else if (paired) wide_round_resolve_d<PS_WIDE_MINB, true><<<g3p, 128, 0, st>>>(a, r);
            else wide_round_resolve_d<PS_WIDE_MINB, false><<<g3, 128, 0, st>>>(a, r);
        }
        B->launches += 3 * (long long)r_launch;
        if (r_launch < kWideMaxLevels) {
            int r_begin = r_launch + 1;
            void* args[] = {(void*)&a, (void*)&r_begin};
            CK(cudaLaunchCooperativeKernel((const void*)wide_rounds, dim3((unsigned)B->rounds_grid), dim3(128), args, 0, st));
            B->launches += 1;
        }
        if (!capturing && !*est_pending) {  // this step's counts, for the grids of a later step
            CK(cudaMemcpyAsync(h_est, a.bucket, (kBuckets + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h_est + kBuckets + 1, a.hit_cnt, (size_t)(kWideMaxLevels + 2) * kHitSlots * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(h_est + 3 * kBuckets + 1, a.surv_cnt, (kWideMaxLevels + 2) * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(est_ready, st));
            *est_pending = true;
        }
    } else {
    CK(cudaMemcpyAsync(B->h_bucket, a.bucket, (kBuckets + 1) * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    lap(0);
    const unsigned* hb = B->h_bucket;
    const bool overflow = hb[kBuckets] > a.entries_cap;  // wide_scatter flagged every world: the fused kernel takes the step
    for (int r = 1; r <= kWideMaxLevels && !overflow; r++) {
        unsigned first = hb[r * 4], total = hb[(r + 1) * 4] - first;
        if (first >= hb[kBuckets]) break;  // no pair at this or any later level
        if (total == 0) continue;
        unsigned c0 = hb[r * 4 + 1] - hb[r * 4], c1 = hb[r * 4 + 2] - hb[r * 4 + 1], c2 = hb[r * 4 + 3] - hb[r * 4 + 2];
        const unsigned nblk = (c0 + 127) / 128 + (c1 + 127) / 128 + (c2 + 127) / 128;
        wide_round_detect<<<nblk, 128, 0, st>>>(a, first, c0, c1, c2, r);
        lap(2);
        if (prof) { B->prof_round.push_back((float)c0); B->prof_round.push_back((float)c1); B->prof_round.push_back((float)c2); B->prof_round.push_back(lap_ms); }
        if (c2) wide_round_bb<<<(c2 + 127) / 128, 128, 0, st>>>(a, first + c0 + c1, c2, r);  // threads past the survivor count exit
        lap(4);
        if (prof) B->prof_round.push_back(c2 ? lap_ms : 0.f);
        wide_round_resolve<<<nblk, 128, 0, st>>>(a, first, c0, c1, c2, r);
        lap(5);
        if (prof) B->prof_round.push_back(lap_ms);
        B->prof_rounds++;
        B->launches += 2 + (c2 != 0);
    }
    }
    wide_verify<<<gb, 128, 0, st>>>(a);
    if (n_stat > 0) wide_static_fold<<<(n_stat * 32 + 127) / 128, 128, 0, st>>>(a, d_stat, n_stat);
    wide_final<<<gb, psw::kTile, psw::kTileBytes, st>>>(a);
    wide_account<<<gw, 64, 0, st>>>(a);
    B->launches += 4;
    lap(6);
    // worlds the wide pass gave up on: redo from the published state with the fused kernel
    // Few flagged worlds: one world per warp (32 lanes each) — a handful of worlds on the batch's narrow tiles would hold the
    // whole step for as long as ONE world takes on 4 lanes.  Many: the batch's own tile.  The count lives on the device; each
    // of the two launches returns at once when it is not in its range.
    KArgs kf = ka;
    kf.n_steps = 1;
    kf.only_flagged = B->d_wflags;
    kf.redo_count = a.nflag;
    kf.world_base = a.w_begin; kf.n_worlds = a.w_end;  // (the end of the launched range)
    const int few = 2048;
    if (B->threads != 32) {
        kf.redo_lo = 1; kf.redo_hi = few;
        const int keep_t = B->threads, keep_m = B->minb;
        const size_t keep_s = B->smem;
        B->threads = 32; B->minb = 8; B->smem = smem_bytes(B->max_n);
        launch_fused(B, kf, st);
        B->threads = keep_t; B->minb = keep_m; B->smem = keep_s;
        kf.redo_lo = few; kf.redo_hi = 1 << 30;
    } else {
        kf.redo_lo = 1; kf.redo_hi = 1 << 30;
    }
    launch_fused(B, kf, st);
    lap(7);
    B->prof_steps++;
    if (prof) { cudaEventDestroy(pe0); cudaEventDestroy(pe1); }
    CK(cudaGetLastError());
    return PS_OK;
}

// Work launched on a caller's stream touches the same device buffers as everything the handle does on its own
// stream (get/set_state, apply_impulse, step_host, stats ...).  enter: the caller's stream waits for what the handle's
// stream holds; leave: the handle's stream (and so ps_batch_synchronize and every later call) waits for the caller's
// work.  While the caller's stream is being captured into a CUDA graph neither is possible (an event recorded outside
// a capture cannot be waited on inside it): the capture then owns the ordering, as the header says.
static int foreign_enter(Batch* B, cudaStream_t st, bool* capturing)
{
    *capturing = false;
    if (st == B->stream) return PS_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    CK(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone) { *capturing = true; return PS_OK; }
    if (!B->ev_own) CK(cudaEventCreateWithFlags(&B->ev_own, cudaEventDisableTiming));
    if (!B->ev_foreign) CK(cudaEventCreateWithFlags(&B->ev_foreign, cudaEventDisableTiming));
    CK(cudaEventRecord(B->ev_own, B->stream));
    CK(cudaStreamWaitEvent(st, B->ev_own, 0));
    return PS_OK;
}
static int foreign_leave(Batch* B, cudaStream_t st, bool capturing)
{
    if (st == B->stream || capturing) return PS_OK;
    CK(cudaEventRecord(B->ev_foreign, st));
    CK(cudaStreamWaitEvent(B->stream, B->ev_foreign, 0));
    return PS_OK;
}

static bool step_needs_host(const Batch* B) { return B->tree_mode || (B->wide && !B->wide_device_rounds); }

static int launch_step(Batch* B, int n_steps, cudaStream_t st)
{
    if (n_steps < 0) return fail(PS_ERR_INVALID_ARGUMENT, "n_steps must be >= 0");
    if (n_steps == 0 || B->n_worlds == 0) return PS_OK;
    KArgs a{};
    fill_kargs(B, a, n_steps);
    if (B->wide) {
        for (int s = 0; s < n_steps; s++) {
            int rc = launch_wide_step(B, a, st);
            if (rc) return rc;
        }
        return PS_OK;
    }
    if (B->tree_mode) {  // the candidate list of a step depends on the state the previous step left: one launch per step
        a.n_steps = 1;
        for (int s = 0; s < n_steps; s++) {
            int rc = tree_update_and_upload(B, st);
            if (rc) return rc;
            launch_fused(B, a, st);
            CK(cudaGetLastError());
        }
        return PS_OK;
    }
    a.perm = current_perm(B, n_steps, st);
    launch_fused(B, a, st);
    CK(cudaGetLastError());
    return PS_OK;
}

int ps_batch_step(void* handle, int32_t n_steps)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    CK(cudaSetDevice(B->device));
    int rc = launch_step(B, n_steps, B->stream);
    if (rc) return rc;
    CK(cudaStreamSynchronize(B->stream));
    return PS_OK;
}
int ps_batch_step_async(void* handle, int32_t n_steps, void* cuda_stream)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    CK(cudaSetDevice(B->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : B->stream;
    bool capturing = false;
    int rc = foreign_enter(B, st, &capturing);
    if (rc) return rc;
    if (capturing && step_needs_host(B))
        return fail(PS_ERR_INVALID_OPERATION, "this batch steps with a host round trip per step (SpatialTree order, or the batch-wide path's per-step bucket read-back): it cannot be captured into a CUDA graph");
    rc = launch_step(B, n_steps, st);
    if (rc) return rc;
    return foreign_leave(B, st, capturing);
}
int ps_batch_synchronize(void* handle)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    return PS_OK;
}

static int check_range(Batch* B, int w0, int w1)
{
    if (w0 < 0 || w1 > B->n_worlds || w0 > w1) return fail(PS_ERR_OUT_OF_RANGE, "world range [%d,%d) outside [0,%d)", w0, w1, B->n_worlds);
    return PS_OK;
}

int ps_batch_get_state_device(void* handle, int32_t w0, int32_t w1, double* d_pos, double* d_vel, double* d_rot,
                              double* d_ang_mom, void* cuda_stream)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    int rc = check_range(B, w0, w1);
    if (rc) return rc;
    CK(cudaSetDevice(B->device));
    int b0 = B->abi_off[w0], n = B->abi_off[w1] - b0;
    if (n == 0) return PS_OK;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : B->stream;
    bool capturing = false;
    rc = foreign_enter(B, st, &capturing);
    if (rc) return rc;
    unpack_dyn_kernel<<<(n + 255) / 256, 256, 0, st>>>(B->d_abi_off, B->d_off, w0, w1, d_pos, d_vel, d_rot, d_ang_mom, B->d_dyn);
    CK(cudaGetLastError());
    return foreign_leave(B, st, capturing);
}

int ps_batch_get_state(void* handle, int32_t w0, int32_t w1, double* pos, double* vel, double* rot, double* ang_mom)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    int rc = check_range(B, w0, w1);
    if (rc) return rc;
    CK(cudaSetDevice(B->device));
    int b0 = B->abi_off[w0], n = B->abi_off[w1] - b0;
    if (n == 0) return PS_OK;
    double* sp = B->d_stage + (size_t)b0 * kDynF;  // this range's staging slice: pos | vel | rot | L
    double *dp = sp, *dv = sp + 3 * (size_t)n, *dr = sp + 6 * (size_t)n, *dl = sp + 15 * (size_t)n;
    unpack_dyn_kernel<<<(n + 255) / 256, 256, 0, B->stream>>>(B->d_abi_off, B->d_off, w0, w1, pos ? dp : nullptr, vel ? dv : nullptr,
                                                               rot ? dr : nullptr, ang_mom ? dl : nullptr, B->d_dyn);
    CK(cudaGetLastError());
    if (pos) CK(cudaMemcpyAsync(pos, dp, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, B->stream));
    if (vel) CK(cudaMemcpyAsync(vel, dv, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, B->stream));
    if (rot) CK(cudaMemcpyAsync(rot, dr, 9 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, B->stream));
    if (ang_mom) CK(cudaMemcpyAsync(ang_mom, dl, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return PS_OK;
}

int ps_batch_set_state(void* handle, int32_t w0, int32_t w1, const double* pos, const double* vel, const double* rot,
                       const double* ang_mom)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    int rc = check_range(B, w0, w1);
    if (rc) return rc;
    if (!pos || !vel || !rot || !ang_mom) return fail(PS_ERR_INVALID_ARGUMENT, "state array is null");
    CK(cudaSetDevice(B->device));
    int b0 = B->abi_off[w0], n = B->abi_off[w1] - b0;
    if (n == 0) return PS_OK;
    double* sp = B->d_stage + (size_t)b0 * kDynF;
    double *dp = sp, *dv = sp + 3 * (size_t)n, *dr = sp + 6 * (size_t)n, *dl = sp + 15 * (size_t)n;
    CK(cudaMemcpyAsync(dp, pos, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, B->stream));
    CK(cudaMemcpyAsync(dv, vel, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, B->stream));
    CK(cudaMemcpyAsync(dr, rot, 9 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, B->stream));
    CK(cudaMemcpyAsync(dl, ang_mom, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, B->stream));
    pack_dyn_kernel<<<(n + 255) / 256, 256, 0, B->stream>>>(B->d_abi_off, B->d_off, w0, w1, dp, dv, dr, dl, B->d_dyn);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(B->stream));
    return PS_OK;
}

// marshal -> step -> unmarshal in one call (what Simulator.updateSimulation does around the native step, Simulator.fs:47-62),
// pipelined over chunks of worlds: chunk k's host->device copy, its step kernel and its device->host copy run on their own
// stream, so the copies of one chunk hide behind the step of the others (worlds are independent, the step kernel takes a
// world range).  The batch-wide and SpatialTree paths step the whole batch at once: there the three stages run in sequence.
}  // extern "C"

// The chunks of ps_batch_step_host on the batch-wide path (see WideChunk).  Measured, C3 with 16 384 worlds, pinned host state
// in -> one step -> pinned host state out: 1 chunk 24.0 ms, 2 chunks 19.0, 4 chunks 16.3, 8 chunks 18.6.
static int ensure_wide_chunks(Batch* B)
{
    if (!B->wchunks.empty()) return PS_OK;
    int K = B->n_slots >= 200000 ? 4 : 1;  // (a small batch is one latency-bound chain either way, and its copies are short)
    if (const char* e = getenv("PS_B200_WIDE_HOST_CHUNKS")) K = std::max(1, std::min(16, atoi(e)));  // tuning knob
    K = std::min(K, std::max(1, B->n_worlds / 32));
    const size_t small = (size_t)(psw::kBuckets + 1) + psw::kBuckets + (size_t)(psw::kWideMaxLevels + 2) * (psw::kHitSlots + 2) + 2;
    std::vector<double> cut((size_t)K + 1, 0.0);  // chunk borders as fractions of the worlds
    for (int k = 0; k <= K; k++) cut[k] = (double)k / K;
    if (const char* e = getenv("PS_B200_WIDE_CHUNK_SHAPE")) {  // tuning knob: relative chunk sizes, e.g. "40,30,20,10"
        std::vector<double> wts;
        for (const char* q = e; *q;) { wts.push_back(std::max(1e-3, atof(q))); while (*q && *q != ',') q++; if (*q) q++; }
        if ((int)wts.size() == K) {
            double tot = 0, acc = 0;
            for (double x : wts) tot += x;
            for (int k = 0; k < K; k++) { acc += wts[k]; cut[k + 1] = acc / tot; }
        }
    }
    std::vector<WideChunk> cs((size_t)K);
    for (int k = 0; k < K; k++) {
        WideChunk& c = cs[k];
        c.w0 = k == 0 ? 0 : (int)(B->n_worlds * cut[k]) / 32 * 32;
        c.w1 = k == K - 1 ? B->n_worlds : (int)(B->n_worlds * cut[k + 1]) / 32 * 32;
        c.g0 = B->cap_off[c.w0]; c.g1 = B->cap_off[c.w1];
        const size_t ns = (size_t)std::max(1, B->n_slots);
        c.e_off = (size_t)B->entries_cap * (size_t)c.g0 / ns;
        c.e_cap = (unsigned)((size_t)B->entries_cap * (size_t)c.g1 / ns - c.e_off);
        c.c_off = (size_t)B->cpool_cap * (size_t)c.g0 / ns;
        c.c_cap = (unsigned)((size_t)B->cpool_cap * (size_t)c.g1 / ns - c.c_off);
        const auto lo = std::lower_bound(B->h_stat_slots.begin(), B->h_stat_slots.end(), c.g0);
        const auto hi = std::lower_bound(B->h_stat_slots.begin(), B->h_stat_slots.end(), c.g1);
        c.stat_first = (int)(lo - B->h_stat_slots.begin()); c.stat_count = (int)(hi - lo);
    }
    for (WideChunk& c : cs) {
        CK(cudaMalloc(&c.d_small, small * sizeof(unsigned)));
        CK(cudaMemset(c.d_small, 0, small * sizeof(unsigned)));
        CK(cudaMallocHost(&c.h_est, (size_t)(4 * psw::kBuckets + 8) * sizeof(unsigned)));
        CK(cudaEventCreateWithFlags(&c.est_ready, cudaEventDisableTiming));
    }
    B->wchunks = std::move(cs);
    return PS_OK;
}

extern "C" {

int ps_batch_step_host(void* handle, int32_t n_steps, const double* pos_in, const double* vel_in, const double* rot_in,
                       const double* ang_mom_in, double* pos_out, double* vel_out, double* rot_out, double* ang_mom_out)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    if (n_steps < 0) return fail(PS_ERR_INVALID_ARGUMENT, "n_steps must be >= 0");
    if (!pos_in || !vel_in || !rot_in || !ang_mom_in || !pos_out || !vel_out || !rot_out || !ang_mom_out)
        return fail(PS_ERR_INVALID_ARGUMENT, "state array is null");
    if (B->n_worlds == 0) return PS_OK;
    const auto t_call = std::chrono::steady_clock::now();
    const int per = B->seg;  // chunk = segment (upload(): PS_B200_HOST_CHUNKS, default 8, borders on CTA borders)
    int chunks = per > 0 ? (B->n_worlds + per - 1) / per : 1;
    const bool wide_chunked = B->wide && B->wide_device_rounds && !B->tree_mode && n_steps > 0 && !getenv("PS_B200_WIDE_PROF");
    if (wide_chunked) {
        CK(cudaSetDevice(B->device));
        int rc = ensure_wide_chunks(B);
        if (rc) return rc;
        chunks = (int)B->wchunks.size();
    }
    if ((B->wide && !wide_chunked) || B->tree_mode || chunks <= 1 || n_steps == 0) {
        int rc = ps_batch_set_state(handle, 0, B->n_worlds, pos_in, vel_in, rot_in, ang_mom_in);
        if (!rc) rc = ps_batch_step(handle, n_steps);
        if (!rc) rc = ps_batch_get_state(handle, 0, B->n_worlds, pos_out, vel_out, rot_out, ang_mom_out);
        return rc;
    }
    CK(cudaSetDevice(B->device));
    while ((int)B->chunk_streams.size() < chunks) {
        cudaStream_t s; cudaEvent_t e;
        // an earlier chunk is further along its chain of small latency-bound rounds while a later one is still in its bulk
        // stages: the earlier one goes first whenever an SM has room
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        int prio = std::min(least, greatest + (int)B->chunk_streams.size());
        if (const char* e = getenv("PS_B200_CHUNK_PRIO")) { if (atoi(e) == 0) prio = least; }  // tuning knob
        CK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio));
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        B->chunk_streams.push_back(s); B->chunk_done.push_back(e);
    }
    if (!B->chunk_start) CK(cudaEventCreateWithFlags(&B->chunk_start, cudaEventDisableTiming));
    KArgs a{};
    fill_kargs(B, a, n_steps);
    a.perm = wide_chunked ? nullptr : current_perm(B, n_steps, B->stream);  // (a re-sort runs on the handle's stream, before the chunks start)
    CK(cudaEventRecord(B->chunk_start, B->stream));  // after whatever the handle's own stream still holds
    static const bool trace = getenv("PS_B200_HOST_TRACE") != nullptr;  // tuning: the timeline of the chunks, host time of the enqueue
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize((size_t)3 * chunks + 1);
        for (cudaEvent_t& e : tev) CK(cudaEventCreate(&e));
        CK(cudaEventRecord(tev[3 * chunks], B->stream));
    }
    for (int k = 0; k < chunks; k++) {
        const int w0 = wide_chunked ? B->wchunks[k].w0 : k * per, w1 = wide_chunked ? B->wchunks[k].w1 : std::min(B->n_worlds, w0 + per);
        const int b0 = B->abi_off[w0], n = B->abi_off[w1] - b0;
        cudaStream_t st = B->chunk_streams[k];
        CK(cudaStreamWaitEvent(st, B->chunk_start, 0));
        if (n > 0) {
            double* sp = B->d_stage + (size_t)b0 * kDynF;
            double *dp = sp, *dv = sp + 3 * (size_t)n, *dr = sp + 6 * (size_t)n, *dl = sp + 15 * (size_t)n;
            CK(cudaMemcpyAsync(dp, pos_in + 3 * (size_t)b0, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(dv, vel_in + 3 * (size_t)b0, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(dr, rot_in + 9 * (size_t)b0, 9 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(dl, ang_mom_in + 3 * (size_t)b0, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
            pack_dyn_kernel<<<(n + 255) / 256, 256, 0, st>>>(B->d_abi_off, B->d_off, w0, w1, dp, dv, dr, dl, B->d_dyn);
            if (trace) CK(cudaEventRecord(tev[3 * k], st));
            if (wide_chunked) {
                for (int s = 0; s < n_steps; s++) {
                    int rc = launch_wide_step(B, a, st, &B->wchunks[k]);
                    if (rc) return rc;
                }
            } else {
                KArgs ak = a;
                ak.world_base = w0; ak.n_worlds = w1;
                launch_fused(B, ak, st);
            }
            if (trace) CK(cudaEventRecord(tev[3 * k + 1], st));
            unpack_dyn_kernel<<<(n + 255) / 256, 256, 0, st>>>(B->d_abi_off, B->d_off, w0, w1, dp, dv, dr, dl, B->d_dyn);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(pos_out + 3 * (size_t)b0, dp, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(vel_out + 3 * (size_t)b0, dv, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(rot_out + 9 * (size_t)b0, dr, 9 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ang_mom_out + 3 * (size_t)b0, dl, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        }
        if (trace) CK(cudaEventRecord(tev[3 * k + 2], st));
        CK(cudaEventRecord(B->chunk_done[k], st));
        CK(cudaStreamWaitEvent(B->stream, B->chunk_done[k], 0));
    }
    const auto t_enq = std::chrono::steady_clock::now();
    CK(cudaStreamSynchronize(B->stream));
    if (trace) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "[ps_b200 step_host] %d chunks: enqueue %.3f ms, call %.3f ms\n", chunks,
                std::chrono::duration<double, std::milli>(t_enq - t_call).count(), std::chrono::duration<double, std::milli>(t_end - t_call).count());
        for (int k = 0; k < chunks; k++) {
            float t0 = 0, t1 = 0, t2 = 0;
            if (cudaEventElapsedTime(&t0, tev[3 * chunks], tev[3 * k]) == cudaSuccess && cudaEventElapsedTime(&t1, tev[3 * chunks], tev[3 * k + 1]) == cudaSuccess &&
                cudaEventElapsedTime(&t2, tev[3 * chunks], tev[3 * k + 2]) == cudaSuccess)
                fprintf(stderr, "[ps_b200 step_host]   chunk %d: state on the device at %.2f ms, stepped at %.2f, back on the host at %.2f\n", k, t0, t1, t2);
        }
        (void)cudaGetLastError();
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return PS_OK;
}

int ps_batch_get_contacts(void* handle, int32_t world, int32_t* n_pairs, int32_t* pair_ids, int32_t* first_contact,
                          double* normal, double* position, double* penetration, uint32_t* feature_id, int32_t pair_capacity,
                          int32_t contact_capacity)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    if (world < 0 || world >= B->n_worlds) return fail(PS_ERR_OUT_OF_RANGE, "world %d outside [0,%d)", world, B->n_worlds);
    if (!B->d_rec_pairs) return fail(PS_ERR_INVALID_OPERATION, "batch was created with record_contacts = 0");
    if (!n_pairs) return fail(PS_ERR_INVALID_ARGUMENT, "n_pairs is null");
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    int cnt[2];
    CK(cudaMemcpy(cnt, B->d_rec_counts + 2 * world, sizeof cnt, cudaMemcpyDeviceToHost));
    if (cnt[0] > B->rec_pair_cap || cnt[1] > B->rec_contact_cap)
        return fail(PS_ERR_CAPACITY, "world %d: %d pairs / %d contacts exceed the recording capacity (%d / %d)", world, cnt[0], cnt[1],
                    B->rec_pair_cap, B->rec_contact_cap);
    *n_pairs = cnt[0];
    if (cnt[0] > pair_capacity || cnt[1] > contact_capacity)
        return fail(PS_ERR_CAPACITY, "need room for %d pairs and %d contacts", cnt[0], cnt[1]);
    std::vector<RecPair> rp(cnt[0]);
    std::vector<RecContact> rcs(cnt[1]);
    if (cnt[0]) CK(cudaMemcpy(rp.data(), B->d_rec_pairs + (size_t)world * B->rec_pair_cap, cnt[0] * sizeof(RecPair), cudaMemcpyDeviceToHost));
    if (cnt[1]) CK(cudaMemcpy(rcs.data(), B->d_rec_contacts + (size_t)world * B->rec_contact_cap, cnt[1] * sizeof(RecContact), cudaMemcpyDeviceToHost));
    // Collisions is a Map keyed by the id set: ascending (i,j) (CollisionResolver.fs:33-34)
    std::sort(rp.begin(), rp.end(), [](const RecPair& x, const RecPair& y) { return x.i != y.i ? x.i < y.i : x.j < y.j; });
    int c = 0;
    for (int k = 0; k < cnt[0]; k++) {
        if (pair_ids) { pair_ids[2 * k] = rp[k].i; pair_ids[2 * k + 1] = rp[k].j; }
        if (first_contact) first_contact[k] = c;
        for (int q = 0; q < rp[k].count; q++, c++) {
            const RecContact& r = rcs[rp[k].first + q];
            if (normal) for (int t = 0; t < 3; t++) normal[3 * c + t] = r.n[t];
            if (position) for (int t = 0; t < 3; t++) position[3 * c + t] = r.p[t];
            if (penetration) penetration[c] = r.pen;
            if (feature_id) feature_id[c] = r.feat;
        }
    }
    if (first_contact) first_contact[cnt[0]] = c;
    return PS_OK;
}

int ps_batch_get_candidates(void* handle, int32_t world, int32_t* n_pairs, int32_t* pair_ids, int32_t pair_capacity)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    if (world < 0 || world >= B->n_worlds) return fail(PS_ERR_OUT_OF_RANGE, "world %d outside [0,%d)", world, B->n_worlds);
    if (B->tree_mode) {  // the list the last step walked (empty before the first step)
        const int c0 = B->h_cand_off[world], c1 = B->h_cand_off[world + 1];
        for (int k = c0; k < c1; k++)
            if (pair_ids && k - c0 < pair_capacity) { pair_ids[2 * (k - c0)] = (int32_t)(B->h_cand[k] & 0xffff); pair_ids[2 * (k - c0) + 1] = (int32_t)(B->h_cand[k] >> 16); }
        if (n_pairs) *n_pairs = c1 - c0;
        if (pair_ids && c1 - c0 > pair_capacity) return fail(PS_ERR_CAPACITY, "need room for %d pairs", c1 - c0);
        return PS_OK;
    }
    // Dummy order (BroadPhase.fs:33-39) filtered by canIntersect (CollisionResolver.fs:58-60): a function of
    // the body set only, so it is produced here from the static flags.
    int o = B->cap_off[world], N = B->cnt[world];
    int n = 0;
    for (int i = 0; i < N; i++)
        for (int j = i + 1; j < N; j++) {
            if (std::isinf(B->slot_mass[o + i]) && std::isinf(B->slot_mass[o + j])) continue;
            if (pair_ids && n < pair_capacity) { pair_ids[2 * n] = i; pair_ids[2 * n + 1] = j; }
            n++;
        }
    if (n_pairs) *n_pairs = n;
    if (pair_ids && n > pair_capacity) return fail(PS_ERR_CAPACITY, "need room for %d pairs", n);
    return PS_OK;
}

int ps_batch_apply_impulse(void* handle, int32_t world, int32_t body, const double impulse[3], const double offset[3])
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    if (world < 0 || world >= B->n_worlds) return fail(PS_ERR_OUT_OF_RANGE, "world %d outside [0,%d)", world, B->n_worlds);
    int N = B->cnt[world];
    if (body < 0 || body >= N) return fail(PS_ERR_OUT_OF_RANGE, "body %d outside [0,%d)", body, N);  // KeyNotFoundException
    if (!impulse || !offset) return fail(PS_ERR_INVALID_ARGUMENT, "impulse/offset is null");
    CK(cudaSetDevice(B->device));
    apply_impulse_kernel<<<1, 1, 0, B->stream>>>(B->d_dyn, B->d_par, B->d_off, world, body, impulse[0], impulse[1], impulse[2],
                                                 offset[0], offset[1], offset[2]);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(B->stream));
    return PS_OK;
}

int ps_batch_add_body(void* handle, int32_t world, const ps_bodies_view* one)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    if (world < 0 || world >= B->n_worlds) return fail(PS_ERR_OUT_OF_RANGE, "world %d outside [0,%d)", world, B->n_worlds);
    int rc = check_view(one);
    if (rc) return rc;
    if (one->n_worlds != 1 || one->world_offset[1] != 1) return fail(PS_ERR_INVALID_ARGUMENT, "add_body takes exactly one body");
    HostBodies nb;
    copy_view(nb, one);
    rc = validate_bodies(nb);
    if (rc) return rc;
    // Everything that can refuse the body is checked BEFORE anything changes (a failed call leaves the batch as it was):
    // the per-world size limit, and in SpatialTree mode the tree's space (BroadPhase.onNewObjectAdded throws
    // "out of space bounds", BroadPhase.fs:83-98 / SpatialTree.fs:171-172).
    const int N0 = B->cnt[world];
    if (N0 + 1 > 1024)
        return fail(PS_ERR_INVALID_ARGUMENT, "world %d would have %d bodies; the per-world stepper takes at most 1024", world, N0 + 1);
    double e6[6];
    if (B->tree_mode) {
        body_extent6(nb.pos.data(), nb.rot.data(), nb.dims.data(), nb.shape[0] == 0, B->cfg.aabb_mode, e6);
        pst::Ext<3> oe;
        for (int k = 0; k < 3; k++) { oe.size[k] = e6[k]; oe.pos[k] = e6[3 + k]; }
        if (!B->trees[world].in_space(oe))
            return fail(PS_ERR_INVALID_ARGUMENT, "world %d: object %d out of space bounds", world, N0);
    }
    CK(cudaSetDevice(B->device));
    const bool is_static = std::isinf(nb.mass[0]);
    if (N0 < B->cap_off[world + 1] - B->cap_off[world] && !is_static) {  // (a static body: re-layout, the static-fold list is rebuilt)
        // ---- the world has a spare slot: the new body (id = max id + 1, SimulatorState.fs:90-92) is a record write ------
        const int g = B->cap_off[world] + N0;
        int nstat = 0;
        for (int q = 0; q < N0; q++) nstat += std::isinf(B->slot_mass[B->cap_off[world] + q]) ? 1 : 0;
        double rec[kParF + kDynF];
        par_record(nb, 0, rec);
        double* d = rec + kParF;
        for (int k = 0; k < 3; k++) { d[k] = nb.pos[k]; d[3 + k] = nb.vel[k]; d[15 + k] = nb.L[k]; }
        for (int k = 0; k < 9; k++) d[6 + k] = nb.rot[k];
        const unsigned char flags[2] = {(unsigned char)(is_static ? std::min(nstat + 1, kMaxStatics) : 0), (unsigned char)(nb.shape[0] != 0)};
        const unsigned char af = (unsigned char)((nstat + (is_static ? 1 : 0)) > kMaxStatics);
        const int newcnt = N0 + 1;
        for (int w = world + 1; w <= B->n_worlds; w++) B->abi_off[w]++;
        cudaStream_t st = B->stream;
        CK(cudaMemcpyAsync(B->d_par + (size_t)g * kParF, rec, kParF * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B->d_dyn + (size_t)g * kDynF, d, kDynF * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B->d_isstat + g, &flags[0], 1, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B->d_isbox + g, &flags[1], 1, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(B->d_hitcnt + g, 0, 1, st));
        CK(cudaMemcpyAsync(B->d_always_fused + world, &af, 1, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B->d_cnt + world, &newcnt, sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(B->d_abi_off + world + 1, B->abi_off.data() + world + 1, (size_t)(B->n_worlds - world) * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        B->cnt[world] = newcnt;
        B->slot_mass[g] = nb.mass[0];
        B->n_bodies++;
        if (!is_static) B->n_dynamic++;
        B->added.push_back(nb);
        B->added_world.push_back(world);
        for (WideChunk& c : B->wchunks) c.est_valid = false;
        B->est_valid = false;  // (grid estimates of the device-counted rounds: harmless if stale, but a changed batch starts afresh)
        if (B->tree_mode) {  // BroadPhase.onNewObjectAdded (BroadPhase.fs:83-98): the new id joins the existing tree
            rc = tree_pull_extents(B, B->stream);  // a split re-files the leaf's objects by their CURRENT extents
            if (rc) return rc;
            const int o = B->cap_off[world];
            if (!B->trees[world].insert(N0, [&](int x) { return ext_of(B, o + x); }))
                return fail(PS_ERR_INVALID_ARGUMENT, "world %d: object %d out of space bounds", world, N0);  // (ruled out above: same extent, same test)
            B->h_cand_off.assign((size_t)B->n_worlds + 1, 0);
            B->h_cand.clear();
        }
        return PS_OK;
    }
    // ---- the world is full: pull the live state, splice every body added so far and this one into the compact
    //      description, and lay the batch out again (with fresh spare slots) --------------------------------------------
    HostBodies merged;
    rc = merged_description(B, merged);
    if (rc) return rc;
    std::vector<WorldStats> keep(B->n_worlds);
    CK(cudaMemcpy(keep.data(), B->d_stats, keep.size() * sizeof(WorldStats), cudaMemcpyDeviceToHost));
    std::vector<pst::Tree<3>> trees_keep;
    if (B->tree_mode) trees_keep = B->trees;
    const HostBodies before = merged;
    {
        HostBodies& h = merged;
        const int at = h.off[world + 1];
        auto ins = [&](std::vector<double>& dst, const std::vector<double>& src, int per) { dst.insert(dst.begin() + (size_t)per * at, src.begin(), src.end()); };
        h.shape.insert(h.shape.begin() + at, nb.shape[0]);
        ins(h.dims, nb.dims, 3); ins(h.mass, nb.mass, 1); ins(h.iinv, nb.iinv, 3); ins(h.pos, nb.pos, 3); ins(h.vel, nb.vel, 3);
        ins(h.rot, nb.rot, 9); ins(h.L, nb.L, 3); ins(h.force, nb.force, 3); ins(h.torque, nb.torque, 3); ins(h.e, nb.e, 1);
        ins(h.muk, nb.muk, 1); ins(h.mus, nb.mus, 1);
        for (int w = world + 1; w <= B->n_worlds; w++) h.off[w]++;
    }
    B->cur = merged;
    rc = upload(B, true);
    if (rc) {  // (shared-memory limit of a grown world, out of device memory): back to the batch as it was
        const std::string why = g_err;
        B->cur = before;
        if (upload(B, true) != PS_OK) {
            B->broken = true;  // no device state left: every entry point but reset/destroy refuses from here on
            return fail(PS_ERR_CUDA, "add_body failed (%s) and the previous batch could not be restored: call ps_batch_reset", why.c_str());
        }
        cudaMemcpy(B->d_stats, keep.data(), keep.size() * sizeof(WorldStats), cudaMemcpyHostToDevice);
        return fail(rc, "%s", why.c_str());
    }
    CK(cudaMemcpy(B->d_stats, keep.data(), keep.size() * sizeof(WorldStats), cudaMemcpyHostToDevice));
    if (B->tree_mode) {
        rc = tree_pull_extents(B, B->stream);
        if (rc) return rc;
        const int o = B->cap_off[world], id = B->cnt[world] - 1;
        if (!B->trees[world].insert(id, [&](int x) { return ext_of(B, o + x); }))
            return fail(PS_ERR_INVALID_ARGUMENT, "world %d: object %d out of space bounds", world, id);  // (ruled out above)
        B->h_cand_off.assign((size_t)B->n_worlds + 1, 0);
        B->h_cand.clear();
    }
    return PS_OK;
}

int ps_batch_reset(void* handle)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    B->cur = B->init;  // bodies added later are dropped (Simulator.fs:36-40,113-121)
    B->broken = false;
    int rc = upload(B, true);
    if (!rc && B->tree_mode) rc = tree_init_all(B);  // initSimulationState builds the tree anew
    B->broken = rc != PS_OK;  // the initial batch uploaded at create; if it does not now, the device is gone
    return rc;
}

int ps_batch_stats(void* handle, ps_stats* out)
{
    if (!handle || !out) return fail(PS_ERR_INVALID_ARGUMENT, "handle/out is null");
    Batch* B = (Batch*)handle;
    PS_REFUSE_IF_BROKEN(B);
    CK(cudaSetDevice(B->device));
    CK(cudaMemsetAsync(B->d_stats_sum, 0, sizeof(ps_stats), B->stream));
    if (B->n_worlds > 0) {
        int blocks = std::min(64, (B->n_worlds + 127) / 128);
        stats_reduce_kernel<<<blocks, 128, 0, B->stream>>>(B->d_stats, B->d_off, B->d_cnt, B->d_par, B->n_worlds, B->d_stats_sum);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(out, B->d_stats_sum, sizeof(ps_stats), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return PS_OK;
}

int ps_batch_schedule(void* handle, int32_t* path, int32_t* lanes_per_world, int64_t* kernel_launches)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    if (path) *path = B->wide ? 1 : 0;
    if (lanes_per_world) *lanes_per_world = B->threads;
    if (kernel_launches) *kernel_launches = B->launches;
    return PS_OK;
}

// debug/profiling: per-world phase cycles (8 x u64 per world: cyc[0..4], slowest step, sum K, sum levels)
int ps_batch_debug_cycles(void* handle, unsigned long long* out)
{
    if (!handle || !out) return fail(PS_ERR_INVALID_ARGUMENT, "handle/out is null");
    Batch* B = (Batch*)handle;
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    std::vector<WorldStats> ws(B->n_worlds);
    CK(cudaMemcpy(ws.data(), B->d_stats, ws.size() * sizeof(WorldStats), cudaMemcpyDeviceToHost));
    for (int w = 0; w < B->n_worlds; w++) {
        for (int k = 0; k < 5; k++) out[8 * w + k] = ws[w].cyc[k];
        out[8 * w + 5] = ws[w].cyc_max_step; out[8 * w + 6] = ws[w].sum_K; out[8 * w + 7] = ws[w].sum_levels;
    }
    return PS_OK;
}

int ps_batch_num_bodies(void* handle, int32_t* n_bodies_total, int32_t* n_dynamic_total)
{
    if (!handle) return fail(PS_ERR_INVALID_ARGUMENT, "handle is null");
    Batch* B = (Batch*)handle;
    if (n_bodies_total) *n_bodies_total = B->n_bodies;
    if (n_dynamic_total) *n_dynamic_total = B->n_dynamic;
    return PS_OK;
}

}  // extern "C"
