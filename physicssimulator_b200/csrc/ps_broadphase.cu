// ps_broadphase.cu — sort-based uniform-grid broad phase for one large world (BASELINE config 5).
//
// Output = OVERLAP(aabb_mode) of SURVEY.md §8(a'): the sorted list of pairs (i<j), at least one body dynamic
// (canIntersect, CollisionResolver.fs:58-60), whose axis-aligned boxes overlap STRICTLY on all three axes with
// the predicate of SpatialTree.areIntersecting (Utilities/SpatialTree.fs:39-41)
//        (sizeA + sizeB) / 2.0 > abs(centreB - centreA)          centre = minCorner + size / 2.0   (:26-28)
// applied to SimulatorObject.getAABoundingBox (Entities/SimulatorObject.fs:58-71; box: Box.worldAxisAlignedSize,
// Entities/Box.fs:69-85; sphere: cube of side = radius, quirk Q5) and objExtentProvider (BroadPhase.fs:21-31).
// Every floating-point expression of that predicate is evaluated in the reference's order, so the set is
// bit-exact against the CPU restatement; the grid only decides which pairs get TESTED.
//
// Pipeline (all kernels hand-written, no CUB/Thrust):
//   aabb_kernel        per body: AABB size + centre, max side                       (HBM-bound, 128 B/body read)
//   classify_kernel    bodies much larger than the mean (the ground) go to a "big" list; cell size = largest
//                      remaining side, grid origin/extent = bounds of the remaining centres
//   key_kernel         64-bit key = cell id << 32 | body id
//   radix sort         LSD, 8 bits per pass, stable (match.any ranking), over the cell bits only
//   cell_range_kernel  [start,end) of every cell in the sorted order
//   pair_kernel        thread per sorted body: the 27 neighbour cell ranges are staged in shared memory,
//                      then walked; a pair is emitted by its lower id
//   big_pair_kernel    each big body against everything
//   radix sort         of the pair keys (i<<32|j) -> ascending (i,j)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/ps_b200.h"
#include "ps_math.cuh"

using namespace psm;
typedef unsigned long long u64;

namespace {

#define BCK(x)                                                     \
    do {                                                           \
        cudaError_t e_ = (x);                                      \
        if (e_ != cudaSuccess) {                                   \
            fprintf(stderr, "ps_broadphase: %s failed: %s\n", #x, cudaGetErrorString(e_)); \
            return PS_ERR_CUDA;                                    \
        }                                                          \
    } while (0)

struct Aabb {       // 56 B per body
    double size[3];
    double cen[3];  // minCorner + size/2 (withPositionCentered)
    unsigned flags; // bit0 static, bit1 big
    unsigned pad;
};

struct GridInfo {
    double sum_side;            // sum of max sides (for the mean)
    unsigned long long hmax;    // ordered-bits max of the max side among non-big bodies
    unsigned long long lo[3];   // ordered-bits min of centres (non-big)
    unsigned long long hi[3];   // ordered-bits max of centres
    unsigned n_big;
    unsigned pad;
};

__device__ __forceinline__ unsigned long long ord(double v)  // order-preserving map double -> u64
{
    long long b = __double_as_longlong(v);
    return (unsigned long long)(b < 0 ? ~b : (b | (1ll << 63)));
}
__host__ __device__ inline double unord(unsigned long long u)
{
    long long b = (u >> 63) ? (long long)(u & ~(1ull << 63)) : ~(long long)u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// SimulatorObject.getAABoundingBox + objExtentProvider, literal operation order
__global__ void aabb_kernel(int n, const int* __restrict__ shape, const double* __restrict__ dims,
                            const double* __restrict__ mass, const double* __restrict__ pos,
                            const double* __restrict__ rot, int aabb_mode, Aabb* __restrict__ out, double* __restrict__ side,
                            GridInfo* gi)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ms = 0.0;
    if (i < n) {
        double sx, sy, sz;
        if (shape[i] == 0) {
            double r = dims[3 * i];
            double s = (aabb_mode == 1) ? 2.0 * r : r;  // Box.createCube radius (Q5) | conservative
            sx = sy = sz = s;
        } else {
            m33 R;
#pragma unroll
            for (int k = 0; k < 9; k++) R.a[k] = rot[9 * (size_t)i + k];
            double hx = dims[3 * i] / 2.0, hy = dims[3 * i + 1] / 2.0, hz = dims[3 * i + 2] / 2.0;
            d3 z = mk(0.0, 0.0, 0.0);
            d3 ex = mulMV(R, mk(hx, 0.0, 0.0)) + z;  // toWorldCoordinates orientation zero
            d3 ey = mulMV(R, mk(0.0, hy, 0.0)) + z;
            d3 ez = mulMV(R, mk(0.0, 0.0, hz)) + z;
            double wx = fabs(ex.x) + fabs(ey.x) + fabs(ez.x);
            double wy = fabs(ex.y) + fabs(ey.y) + fabs(ez.y);
            double wz = fabs(ex.z) + fabs(ey.z) + fabs(ez.z);
            sx = wx * 2.0; sy = wy * 2.0; sz = wz * 2.0;
        }
        d3 c = mk(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2]);
        d3 mn = c - scale(1.0 / 2.0, mk(sx, sy, sz));  // BroadPhase.fs:24
        Aabb a;
        a.size[0] = sx; a.size[1] = sy; a.size[2] = sz;
        a.cen[0] = mn.x + sx / 2.0; a.cen[1] = mn.y + sy / 2.0; a.cen[2] = mn.z + sz / 2.0;
        a.flags = isinf(mass[i]) ? 1u : 0u;
        a.pad = 0;
        out[i] = a;
        ms = fmax(sx, fmax(sy, sz));
        side[i] = ms;
    }
    // block sum of the max side -> mean (only steers the big/small split, never the result)
    __shared__ double red[8];
    double v = (ms == ms && !isinf(ms)) ? ms : 0.0;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) s += red[k];
        atomicAdd(&gi->sum_side, s);
    }
}

// big/small split + bounds of the small bodies; reduced per CTA before touching the 7 global words
// (1.1 M threads hammering them directly took milliseconds)
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v)
{
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ unsigned long long warp_max64(unsigned long long v)
{
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
__global__ void classify_kernel(int n, Aabb* __restrict__ aabb, const double* __restrict__ side, GridInfo* gi,
                                int* __restrict__ big_list, int big_cap)
{
    // reduced per warp with shuffles, then per CTA in shared memory, before touching the 7 global words
    // (first version: every thread on the global words, milliseconds; second: every thread on 7 shared words, 37 us)
    __shared__ unsigned long long s_h, s_lo[3], s_hi[3];
    if (threadIdx.x == 0) {
        s_h = 0ull;
        for (int k = 0; k < 3; k++) { s_lo[k] = ~0ull; s_hi[k] = 0ull; }
    }
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = 0ull, lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
    if (i < n) {
        double mean = gi->sum_side / (double)n;
        double s = side[i];
        const double c0 = aabb[i].cen[0], c1 = aabb[i].cen[1], c2 = aabb[i].cen[2];
        bool big = !(s <= 8.0 * mean) || !isfinite(c0) || !isfinite(c1) || !isfinite(c2);
        if (big) {
            unsigned k = atomicAdd(&gi->n_big, 1u);
            if ((int)k < big_cap) big_list[k] = i;
            aabb[i].flags |= 2u;
        } else {
            h = ord(s);
            lo[0] = hi[0] = ord(c0); lo[1] = hi[1] = ord(c1); lo[2] = hi[2] = ord(c2);
        }
    }
    h = warp_max64(h);
    for (int k = 0; k < 3; k++) { lo[k] = warp_min64(lo[k]); hi[k] = warp_max64(hi[k]); }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&s_h, h);
        for (int k = 0; k < 3; k++) { atomicMin(&s_lo[k], lo[k]); atomicMax(&s_hi[k], hi[k]); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicMax(&gi->hmax, s_h);
        for (int k = 0; k < 3; k++) {
            atomicMin(&gi->lo[k], s_lo[k]);
            atomicMax(&gi->hi[k], s_hi[k]);
        }
    }
}

struct Grid {
    double org[3], inv_h;
    int dim[3];
    unsigned ncells;
};

__device__ __forceinline__ int cell_coord(double c, double org, double inv_h, int dim)
{
    int v = (int)floor((c - org) * inv_h);
    return v < 0 ? 0 : (v >= dim ? dim - 1 : v);
}

__global__ void key_kernel(int n, const Aabb* __restrict__ aabb, Grid g, u64* __restrict__ keys)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Aabb& a = aabb[i];
    unsigned cell;
    if (a.flags & 2u) {
        cell = g.ncells;  // big bodies sort to the tail
    } else {
        int ix = cell_coord(a.cen[0], g.org[0], g.inv_h, g.dim[0]);
        int iy = cell_coord(a.cen[1], g.org[1], g.inv_h, g.dim[1]);
        int iz = cell_coord(a.cen[2], g.org[2], g.inv_h, g.dim[2]);
        cell = (unsigned)ix + (unsigned)g.dim[0] * ((unsigned)iy + (unsigned)g.dim[1] * (unsigned)iz);
    }
    keys[i] = ((u64)cell << 32) | (unsigned)i;
}

// ---- exclusive scan of unsigned (block = 512 threads x 8 items, recursive over block sums) ---------------------------
constexpr int SCAN_T = 512, SCAN_I = 8, SCAN_TILE = SCAN_T * SCAN_I;
__global__ void __launch_bounds__(SCAN_T) scan_block_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out, int n, unsigned* __restrict__ bsum)
{
    __shared__ unsigned wsum[32];
    const int base = blockIdx.x * SCAN_TILE + SCAN_I * threadIdx.x;
    unsigned v[SCAN_I];
    if (base + SCAN_I <= n && (((size_t)in & 15) == 0)) {
        const uint4 a = *reinterpret_cast<const uint4*>(in + base), b = *reinterpret_cast<const uint4*>(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_I; k++) v[k] = base + k < n ? in[base + k] : 0u;
    }
    unsigned tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_I; k++) { unsigned t = v[k]; v[k] = tsum; tsum += t; }  // exclusive inside the thread
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned x = tsum;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned t = lane < SCAN_T / 32 ? wsum[lane] : 0u;
        unsigned sc = t;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o) sc += y;
        }
        if (lane < SCAN_T / 32) wsum[lane] = sc - t;  // exclusive warp offsets
        if (lane == 31 && bsum) bsum[blockIdx.x] = sc;
    }
    __syncthreads();
    const unsigned excl = x - tsum + wsum[wid];
    if (base + SCAN_I <= n && (((size_t)out & 15) == 0)) {
        *reinterpret_cast<uint4*>(out + base) = make_uint4(v[0] + excl, v[1] + excl, v[2] + excl, v[3] + excl);
        *reinterpret_cast<uint4*>(out + base + 4) = make_uint4(v[4] + excl, v[5] + excl, v[6] + excl, v[7] + excl);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_I; k++)
            if (base + k < n) out[base + k] = v[k] + excl;
    }
}
__global__ void __launch_bounds__(SCAN_T) scan_add_kernel(unsigned* __restrict__ out, int n, const unsigned* __restrict__ boff)
{
    const int base = blockIdx.x * SCAN_TILE + SCAN_I * threadIdx.x;
    const unsigned o = boff[blockIdx.x];
    if (o == 0u) return;
    if (base + SCAN_I <= n && (((size_t)out & 15) == 0)) {
        uint4 a = *reinterpret_cast<uint4*>(out + base), b = *reinterpret_cast<uint4*>(out + base + 4);
        a.x += o; a.y += o; a.z += o; a.w += o; b.x += o; b.y += o; b.z += o; b.w += o;
        *reinterpret_cast<uint4*>(out + base) = a;
        *reinterpret_cast<uint4*>(out + base + 4) = b;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_I; k++)
            if (base + k < n) out[base + k] += o;
    }
}
int exclusive_scan(unsigned* d_in_out, int n, unsigned* scratch /* >= n/4096 + n/4096^2 + 4 */, cudaStream_t st)
{
    int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_block_kernel<<<nb, SCAN_T, 0, st>>>(d_in_out, d_in_out, n, nb > 1 ? scratch : nullptr);
    if (nb > 1) {
        int rc = exclusive_scan(scratch, nb, scratch + ((nb + 3) & ~3), st);
        if (rc) return rc;
        scan_add_kernel<<<nb, SCAN_T, 0, st>>>(d_in_out, n, scratch);
    }
    return 0;
}

// ---- LSD radix sort of u64 keys, 8 bits per pass, stable -----------------------------------------------------
constexpr int RS_T = 256, RS_ITEMS = 8, RS_TILE = RS_T * RS_ITEMS;
__global__ void rs_hist_kernel(const u64* __restrict__ keys, int n, int shift, unsigned* __restrict__ hist, int nblocks)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
    for (int c = 0; c < RS_ITEMS; c++) {
        int i = base + c * RS_T + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}
__global__ void rs_scatter_kernel(const u64* __restrict__ in, u64* __restrict__ out, int n, int shift,
                                  const unsigned* __restrict__ hist, int nblocks)
{
    __shared__ unsigned base[256];
    __shared__ unsigned wcnt[RS_T / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    base[tid] = hist[tid * nblocks + blockIdx.x];
    for (int w = 0; w < RS_T / 32; w++) wcnt[w][tid] = 0;
    __syncthreads();
    int tile = blockIdx.x * RS_TILE;
    for (int c = 0; c < RS_ITEMS; c++) {
        int i = tile + c * RS_T + tid;
        bool valid = i < n;
        u64 key = valid ? in[i] : 0ull;
        unsigned d = valid ? ((unsigned)(key >> shift) & 255u) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[wid][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned pre = 0;
            for (int w = 0; w < wid; w++) pre += wcnt[w][d];
            out[base[d] + pre + rank] = key;
        }
        __syncthreads();
        unsigned tot = 0;
        for (int w = 0; w < RS_T / 32; w++) {
            tot += wcnt[w][tid];
            wcnt[w][tid] = 0;
        }
        base[tid] += tot;
        __syncthreads();
    }
}
// sorts bits [lo, hi) of the keys; result ends in *a (ping-pong with *b)
int radix_sort(u64** a, u64** b, int n, int lo, int hi, unsigned* hist, unsigned* scratch, cudaStream_t st)
{
    if (n <= 1) return 0;
    int nblocks = (n + RS_TILE - 1) / RS_TILE;
    for (int shift = lo; shift < hi; shift += 8) {
        rs_hist_kernel<<<nblocks, RS_T, 0, st>>>(*a, n, shift, hist, nblocks);
        int rc = exclusive_scan(hist, 256 * nblocks, scratch, st);
        if (rc) return rc;
        rs_scatter_kernel<<<nblocks, RS_T, 0, st>>>(*a, *b, n, shift, hist, nblocks);
        std::swap(*a, *b);
    }
    return 0;
}

__global__ void cell_range_kernel(int n_small, const u64* __restrict__ keys, unsigned* __restrict__ cstart, unsigned* __restrict__ cend)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_small) return;
    unsigned c = (unsigned)(keys[s] >> 32);
    if (s == 0 || (unsigned)(keys[s - 1] >> 32) != c) cstart[c] = (unsigned)s;
    if (s == n_small - 1 || (unsigned)(keys[s + 1] >> 32) != c) cend[c] = (unsigned)s + 1u;
}

// SpatialTree.areIntersecting on two centred extents, all three axes
__device__ __forceinline__ bool overlap(const Aabb& A, const Aabb& B)
{
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (!((A.size[k] + B.size[k]) / 2.0 > fabs(B.cen[k] - A.cen[k]))) return false;
    return true;
}

__device__ __forceinline__ void emit(u64* pairs, unsigned long long* count, long long cap, unsigned i, unsigned j)
{
    unsigned long long k = atomicAdd(count, 1ull);
    if ((long long)k < cap) pairs[k] = ((u64)i << 32) | j;
}

constexpr int PK_T = 128;
__global__ void __launch_bounds__(PK_T) pair_kernel(int n_small, const u64* __restrict__ keys, const Aabb* __restrict__ aabb, Grid g,
                                                    const unsigned* __restrict__ cstart, const unsigned* __restrict__ cend,
                                                    u64* __restrict__ pairs, unsigned long long* count, long long cap)
{
    // the 27 neighbour cell ranges of each thread's body, staged in shared memory (struct of arrays: no bank conflicts)
    __shared__ unsigned rs[27][PK_T], re[27][PK_T];
    int s = blockIdx.x * PK_T + threadIdx.x;
    bool act = s < n_small;
    unsigned i = 0;
    Aabb A;
    if (act) {
        u64 key = keys[s];
        i = (unsigned)key;
        A = aabb[i];
        unsigned c = (unsigned)(key >> 32);
        int ix = c % g.dim[0], iy = (c / g.dim[0]) % g.dim[1], iz = c / (g.dim[0] * g.dim[1]);
        int q = 0;
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++, q++) {
                    int x = ix + dx, y = iy + dy, z = iz + dz;
                    unsigned b = 0, e = 0;
                    if (x >= 0 && y >= 0 && z >= 0 && x < g.dim[0] && y < g.dim[1] && z < g.dim[2]) {
                        unsigned cc = (unsigned)x + (unsigned)g.dim[0] * ((unsigned)y + (unsigned)g.dim[1] * (unsigned)z);
                        b = cstart[cc];
                        e = cend[cc];
                    }
                    rs[q][threadIdx.x] = b;
                    re[q][threadIdx.x] = e;
                }
    }
    __syncthreads();
    if (!act) return;
    for (int q = 0; q < 27; q++) {
        unsigned b = rs[q][threadIdx.x], e = re[q][threadIdx.x];
        for (unsigned t = b; t < e; t++) {
            unsigned j = (unsigned)keys[t];
            if (j <= i) continue;  // the lower id emits
            const Aabb& Bb = aabb[j];
            if ((A.flags & Bb.flags & 1u)) continue;  // static-static
            if (overlap(A, Bb)) emit(pairs, count, cap, i, j);
        }
    }
}

__global__ void big_pair_kernel(int n, int g_id, const Aabb* __restrict__ aabb, u64* __restrict__ pairs, unsigned long long* count, long long cap)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || j == g_id) return;
    const Aabb& G = aabb[g_id];
    const Aabb& B = aabb[j];
    if ((B.flags & 2u) && j < g_id) return;  // big-big pairs once
    if (G.flags & B.flags & 1u) return;
    unsigned lo = (unsigned)min(j, g_id), hi = (unsigned)max(j, g_id);
    if (overlap(aabb[lo], aabb[hi])) emit(pairs, count, cap, lo, hi);
}

__global__ void split_pairs_kernel(long long n, const u64* __restrict__ keys, int* __restrict__ out)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[2 * k] = (int)(keys[k] >> 32);
    out[2 * k + 1] = (int)(keys[k] & 0xffffffffu);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fast path (round 2): counting sort by cell, cell-sorted COPIES of the boxes, ownership by the lower id.
//   cell_kernel      per body: cell id + its rank inside the cell (the histogram's atomicAdd returns it)
//   scan             cell histogram -> start of every cell in the sorted order (no key sort, no cell_range pass)
//   scatter_kernel   48-byte box + id + static flag + cell to their sorted slot: the neighbours of a body are then
//                    CONTIGUOUS — the three cells x-1, x, x+1 of a grid row are one span of the sorted arrays, so a body
//                    walks 9 spans instead of 27 cells and never gathers a box by id
//   own_pairs_kernel thread per sorted body i: partners j > i among the 9 spans and among the big bodies, kept as a short
//                    SORTED list in registers (insertion), count to cnt[i] — every pair has one owner, no atomics
//   big_owner_*      a big body as the owner: ordered compaction over all j > g (flags + scan)
//   scan cnt         -> first pair of every body in the output: the list comes out sorted by (i, j) by construction,
//                    the pair keys are never sorted
//   write_pairs_kernel
// A body with more partners than the list holds (dense clusters) or more than a few big bodies send the call to the
// general path below (global emission + radix sort of the pair keys); results are identical either way.
// ---------------------------------------------------------------------------------------------------------------------
struct SBox {  // 48 B, cell-sorted copy of an Aabb's numbers
    double size[3], cen[3];
};
constexpr int kListCap = 12;
constexpr int kMaxBigFast = 4;

__global__ void cell_kernel(int n, const Aabb* __restrict__ aabb, Grid g, unsigned* __restrict__ cellid, unsigned* __restrict__ rank,
                            unsigned* __restrict__ hist)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Aabb& a = aabb[i];
    if (a.flags & 2u) { cellid[i] = 0xffffffffu; return; }
    int ix = cell_coord(a.cen[0], g.org[0], g.inv_h, g.dim[0]);
    int iy = cell_coord(a.cen[1], g.org[1], g.inv_h, g.dim[1]);
    int iz = cell_coord(a.cen[2], g.org[2], g.inv_h, g.dim[2]);
    unsigned c = (unsigned)ix + (unsigned)g.dim[0] * ((unsigned)iy + (unsigned)g.dim[1] * (unsigned)iz);
    cellid[i] = c;
    rank[i] = atomicAdd(&hist[c], 1u);
}
__global__ void scatter_kernel(int n, const Aabb* __restrict__ aabb, const unsigned* __restrict__ cellid, const unsigned* __restrict__ rank,
                               const unsigned* __restrict__ cstart, SBox* __restrict__ sbox, unsigned* __restrict__ sid,
                               unsigned* __restrict__ scell, unsigned char* __restrict__ sstat)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned c = cellid[i];
    if (c == 0xffffffffu) return;
    const unsigned pos = cstart[c] + rank[i];
    const Aabb a = aabb[i];
    double2* q = reinterpret_cast<double2*>(sbox + pos);  // 48-byte records: three 16-byte stores
    q[0] = make_double2(a.size[0], a.size[1]);
    q[1] = make_double2(a.size[2], a.cen[0]);
    q[2] = make_double2(a.cen[1], a.cen[2]);
    sid[pos] = (unsigned)i;
    scell[pos] = c;
    sstat[pos] = (unsigned char)(a.flags & 1u);
}
__device__ __forceinline__ bool overlap_s(const SBox& A, const SBox& B)
{
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (!((A.size[k] + B.size[k]) / 2.0 > fabs(B.cen[k] - A.cen[k]))) return false;
    return true;
}
constexpr int OP_T = 128;
__global__ void __launch_bounds__(OP_T) own_pairs_kernel(int n_small, Grid g, const unsigned* __restrict__ cstart, const SBox* __restrict__ sbox,
                                                         const unsigned* __restrict__ sid, const unsigned* __restrict__ scell,
                                                         const unsigned char* __restrict__ sstat, const Aabb* __restrict__ aabb,
                                                         const int* __restrict__ big_list, int n_big, unsigned* __restrict__ cnt,
                                                         unsigned* __restrict__ lists, unsigned* __restrict__ overflow)
{
    __shared__ SBox bigb[kMaxBigFast];
    __shared__ unsigned bigid[kMaxBigFast], bigst[kMaxBigFast];
    if ((int)threadIdx.x < n_big) {
        const Aabb a = aabb[big_list[threadIdx.x]];
        for (int k = 0; k < 3; k++) { bigb[threadIdx.x].size[k] = a.size[k]; bigb[threadIdx.x].cen[k] = a.cen[k]; }
        bigid[threadIdx.x] = (unsigned)big_list[threadIdx.x];
        bigst[threadIdx.x] = a.flags & 1u;
    }
    __syncthreads();
    const int s = blockIdx.x * OP_T + threadIdx.x;
    if (s >= n_small) return;
    const SBox A = sbox[s];
    const unsigned i = sid[s], c = scell[s], ast = sstat[s];
    const int ix = (int)(c % (unsigned)g.dim[0]), iy = (int)((c / (unsigned)g.dim[0]) % (unsigned)g.dim[1]), iz = (int)(c / ((unsigned)g.dim[0] * (unsigned)g.dim[1]));
    unsigned list[kListCap] = {};
    int m = 0;
    bool over = false;
    auto add = [&](unsigned j) {  // sorted insertion; partners are few
        if (m == kListCap) { over = true; return; }
        int q = m++;
#pragma unroll 1
        while (q > 0 && list[q - 1] > j) { list[q] = list[q - 1]; q--; }
        list[q] = j;
    };
    const int x0 = max(ix - 1, 0), x1 = min(ix + 1, g.dim[0] - 1);
#pragma unroll 1
    for (int dz = -1; dz <= 1; dz++) {
        const int z = iz + dz;
        if (z < 0 || z >= g.dim[2]) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            const int y = iy + dy;
            if (y < 0 || y >= g.dim[1]) continue;
            const unsigned row = (unsigned)g.dim[0] * ((unsigned)y + (unsigned)g.dim[1] * (unsigned)z);
            const unsigned b = cstart[row + (unsigned)x0], e = cstart[row + (unsigned)x1 + 1u];  // one span: cells x0..x1 of the row
#pragma unroll 1
            for (unsigned t = b; t < e; t++) {
                const unsigned j = sid[t];
                if (j <= i) continue;  // the lower id owns the pair
                if (ast & sstat[t]) continue;  // static-static (canIntersect)
                if (overlap_s(A, sbox[t])) add(j);
            }
        }
    }
    for (int q = 0; q < n_big; q++) {
        const unsigned gidx = bigid[q];
        if (gidx <= i || (ast & bigst[q])) continue;
        if (overlap_s(A, bigb[q])) add(gidx);
    }
    cnt[i] = (unsigned)m;
    if (over) atomicOr(overflow, 1u);
    unsigned* L = lists + (size_t)i * kListCap;
    for (int q = 0; q < m; q++) L[q] = list[q];
}
// a big body g as the owner: flag of every j > g that overlaps it (1 / 0), to be scanned into ordered slots
__global__ void big_owner_flag_kernel(int n, int gidx, const Aabb* __restrict__ aabb, unsigned* __restrict__ flags)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    unsigned f = 0;
    if (j > gidx) {
        const Aabb& G = aabb[gidx];
        const Aabb& B = aabb[j];
        if (!(G.flags & B.flags & 1u) && overlap(G, B)) f = 1u;
    }
    flags[j] = f;
}
// cnt[g] = total of the scanned flags (slot of the last body + its own flag)
__global__ void big_owner_count_kernel(int n, int gidx, const Aabb* __restrict__ aabb, const unsigned* __restrict__ slots, unsigned* __restrict__ cnt)
{
    const int j = n - 1;
    unsigned f = 0;
    if (j > gidx) {
        const Aabb& G = aabb[gidx];
        const Aabb& B = aabb[j];
        if (!(G.flags & B.flags & 1u) && overlap(G, B)) f = 1u;
    }
    cnt[gidx] = slots[j] + f;
}
__global__ void big_owner_write_kernel(int n, int gidx, const Aabb* __restrict__ aabb, const unsigned* __restrict__ slots,
                                       const unsigned* __restrict__ off, int* __restrict__ out, long long cap)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || j <= gidx) return;
    const Aabb& G = aabb[gidx];
    const Aabb& B = aabb[j];
    if ((G.flags & B.flags & 1u) || !overlap(G, B)) return;
    const long long k = (long long)off[gidx] + slots[j];
    if (k < cap) { out[2 * k] = gidx; out[2 * k + 1] = j; }
}
__global__ void write_pairs_kernel(int n, const Aabb* __restrict__ aabb, const unsigned* __restrict__ cnt, const unsigned* __restrict__ off,
                                   const unsigned* __restrict__ lists, int* __restrict__ out, long long cap)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (aabb[i].flags & 2u) return;  // big owners write their own
    const unsigned m = cnt[i];
    if (!m) return;
    const unsigned* L = lists + (size_t)i * kListCap;
    long long k = off[i];
    for (unsigned q = 0; q < m && q < (unsigned)kListCap; q++, k++)
        if (k < cap) { out[2 * k] = i; out[2 * k + 1] = (int)L[q]; }
}

template <class T>
struct DBuf {
    T* p = nullptr;
    ~DBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

}  // namespace

extern "C" int ps_broadphase_pairs(int32_t n, const int32_t* shape, const double* dims, const double* mass, const double* pos,
                                   const double* rot, int32_t aabb_mode, int64_t* n_pairs, int32_t* pair_ids, int64_t pair_capacity,
                                   double* kernel_ms)
{
    if (n < 0 || !n_pairs || (n > 0 && (!shape || !dims || !mass || !pos || !rot))) return PS_ERR_INVALID_ARGUMENT;
    if (aabb_mode != 0 && aabb_mode != 1) return PS_ERR_INVALID_ARGUMENT;
    *n_pairs = 0;
    if (kernel_ms) *kernel_ms = 0.0;
    if (n < 2) return PS_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PS_ERR_CUDA;  // no CPU path

    cudaStream_t st = nullptr;
    DBuf<int> d_shape, d_big, d_out;
    DBuf<double> d_dims, d_mass, d_pos, d_rot, d_side;
    DBuf<Aabb> d_aabb;
    DBuf<GridInfo> d_gi;
    DBuf<u64> d_k0, d_k1, d_p0, d_p1;
    DBuf<unsigned> d_hist, d_scr, d_cs, d_ce;
    DBuf<unsigned long long> d_count;
    DBuf<unsigned> d_cellid, d_rank, d_sid, d_scell, d_cnt, d_off, d_lists, d_flags, d_over;
    DBuf<SBox> d_sbox;
    DBuf<unsigned char> d_sstat;
    const int big_cap = 4096;
    long long cap = std::max<long long>(pair_capacity, 1);
    BCK(d_shape.alloc(n)); BCK(d_dims.alloc(3 * (size_t)n)); BCK(d_mass.alloc(n)); BCK(d_pos.alloc(3 * (size_t)n));
    BCK(d_rot.alloc(9 * (size_t)n)); BCK(d_side.alloc(n)); BCK(d_aabb.alloc(n)); BCK(d_gi.alloc(1)); BCK(d_big.alloc(big_cap));
    BCK(d_k0.alloc(n)); BCK(d_k1.alloc(n)); BCK(d_p0.alloc(cap)); BCK(d_p1.alloc(cap)); BCK(d_count.alloc(1));
    long long nsort = std::max<long long>(n, cap);
    int nblocks_max = (int)((nsort + RS_TILE - 1) / RS_TILE);
    BCK(d_hist.alloc(256 * (size_t)nblocks_max));
    BCK(d_scr.alloc((size_t)(256 * (size_t)nblocks_max) / SCAN_TILE + 67108864u / SCAN_TILE + 8192));
    // cell tables for the largest grid the sizing loop below may pick, and the output: no allocation inside the timed region
    const size_t max_cells = 67108864u;
    BCK(d_cs.alloc(max_cells + 2)); BCK(d_ce.alloc(max_cells + 1));
    BCK(d_out.alloc(2 * (size_t)cap));
    BCK(d_cellid.alloc(n)); BCK(d_rank.alloc(n)); BCK(d_sid.alloc(n)); BCK(d_scell.alloc(n)); BCK(d_sstat.alloc(n)); BCK(d_sbox.alloc(n));
    BCK(d_cnt.alloc((size_t)n + 1)); BCK(d_off.alloc((size_t)n + 1)); BCK(d_lists.alloc((size_t)n * kListCap));
    BCK(d_flags.alloc((size_t)n * kMaxBigFast)); BCK(d_over.alloc(1));
    BCK(cudaMemcpy(d_shape.p, shape, n * sizeof(int), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(d_dims.p, dims, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(d_mass.p, mass, n * sizeof(double), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(d_pos.p, pos, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(d_rot.p, rot, 9 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice));

    cudaEvent_t e0, e1;
    BCK(cudaEventCreate(&e0)); BCK(cudaEventCreate(&e1));
    BCK(cudaEventRecord(e0, st));

    GridInfo gi{};
    gi.sum_side = 0.0; gi.hmax = 0ull; gi.n_big = 0;
    for (int k = 0; k < 3; k++) { gi.lo[k] = ~0ull; gi.hi[k] = 0ull; }
    BCK(cudaMemcpyAsync(d_gi.p, &gi, sizeof gi, cudaMemcpyHostToDevice, st));
    BCK(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), st));
    int nb = (n + 255) / 256;
    aabb_kernel<<<nb, 256, 0, st>>>(n, d_shape.p, d_dims.p, d_mass.p, d_pos.p, d_rot.p, aabb_mode, d_aabb.p, d_side.p, d_gi.p);
    classify_kernel<<<nb, 256, 0, st>>>(n, d_aabb.p, d_side.p, d_gi.p, d_big.p, big_cap);
    int big_head[kMaxBigFast] = {0, 0, 0, 0};  // (the first few big ids travel with the grid numbers: one host round trip)
    BCK(cudaMemcpyAsync(&gi, d_gi.p, sizeof gi, cudaMemcpyDeviceToHost, st));
    BCK(cudaMemcpyAsync(big_head, d_big.p, sizeof big_head, cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    if ((int)gi.n_big > big_cap) return PS_ERR_CAPACITY;
    int n_small = n - (int)gi.n_big;

    Grid g{};
    if (n_small > 0) {
        double h = unord(gi.hmax);
        if (!(h > 0.0)) h = 1.0;
        h *= 1.0 + 1e-9;  // cell >= every small side, so overlapping small boxes sit in adjacent cells
        double lo[3], ext[3];
        for (int k = 0; k < 3; k++) { lo[k] = unord(gi.lo[k]); ext[k] = unord(gi.hi[k]) - lo[k]; }
        for (;;) {  // keep the dense cell table below 2^26 cells
            double cells = 1.0;
            for (int k = 0; k < 3; k++) { g.dim[k] = (int)std::min(2.0e9, std::floor(ext[k] / h)) + 1; cells *= g.dim[k]; }
            if (cells <= 67108864.0) break;
            h *= 1.26;
        }
        for (int k = 0; k < 3; k++) g.org[k] = lo[k];
        g.inv_h = 1.0 / h;
        g.ncells = (unsigned)g.dim[0] * (unsigned)g.dim[1] * (unsigned)g.dim[2];
    } else {
        g.dim[0] = g.dim[1] = g.dim[2] = 1; g.ncells = 1; g.inv_h = 1.0;
    }
    // ---- fast path: counting sort by cell, owned pair lists, output sorted by construction (see the kernels) ----------
    bool done = false;
    unsigned long long cnt = 0;
    if ((int)gi.n_big <= kMaxBigFast && !getenv("PS_B200_BROADPHASE_GENERAL")) {
        const int nbig = (int)gi.n_big;
        BCK(cudaMemsetAsync(d_cs.p, 0, ((size_t)g.ncells + 2) * sizeof(unsigned), st));
        BCK(cudaMemsetAsync(d_cnt.p, 0, ((size_t)n + 1) * sizeof(unsigned), st));
        BCK(cudaMemsetAsync(d_over.p, 0, sizeof(unsigned), st));
        if (n_small > 0) {
            cell_kernel<<<nb, 256, 0, st>>>(n, d_aabb.p, g, d_cellid.p, d_rank.p, d_cs.p);
            if (exclusive_scan(d_cs.p, (int)g.ncells + 1, d_scr.p, st)) return PS_ERR_CUDA;   // d_cs: start of every cell; [ncells] = n_small
            scatter_kernel<<<nb, 256, 0, st>>>(n, d_aabb.p, d_cellid.p, d_rank.p, d_cs.p, d_sbox.p, d_sid.p, d_scell.p, d_sstat.p);
            own_pairs_kernel<<<(n_small + OP_T - 1) / OP_T, OP_T, 0, st>>>(n_small, g, d_cs.p, d_sbox.p, d_sid.p, d_scell.p, d_sstat.p, d_aabb.p, d_big.p, nbig,
                                                                            d_cnt.p, d_lists.p, d_over.p);
        }
        for (int q = 0; q < nbig; q++) {  // the big bodies as owners: ordered slots of their partners j > g
            unsigned* fl = d_flags.p + (size_t)q * n;
            big_owner_flag_kernel<<<nb, 256, 0, st>>>(n, big_head[q], d_aabb.p, fl);
            if (exclusive_scan(fl, n, d_scr.p, st)) return PS_ERR_CUDA;
            big_owner_count_kernel<<<1, 1, 0, st>>>(n, big_head[q], d_aabb.p, fl, d_cnt.p);
        }
        BCK(cudaMemcpyAsync(d_off.p, d_cnt.p, ((size_t)n + 1) * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
        if (exclusive_scan(d_off.p, n + 1, d_scr.p, st)) return PS_ERR_CUDA;                   // first pair of every body; [n] = number of pairs
        unsigned tot_over[2] = {0, 0};
        BCK(cudaMemcpyAsync(&tot_over[0], d_off.p + n, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(&tot_over[1], d_over.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (!tot_over[1]) {
            cnt = tot_over[0];
            *n_pairs = (int64_t)cnt;
            if ((long long)cnt > pair_capacity) {
                cudaEventDestroy(e0); cudaEventDestroy(e1);
                return PS_ERR_CAPACITY;
            }
            if (cnt > 0) {
                write_pairs_kernel<<<nb, 256, 0, st>>>(n, d_aabb.p, d_cnt.p, d_off.p, d_lists.p, d_out.p, cap);
                for (int q = 0; q < nbig; q++)
                    big_owner_write_kernel<<<nb, 256, 0, st>>>(n, big_head[q], d_aabb.p, d_flags.p + (size_t)q * n, d_off.p, d_out.p, cap);
            }
            done = true;
        }
    }
    if (!done) {
    // ---- general path: any number of big bodies, any cluster density ---------------------------------------------------
    BCK(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned long long), st));
    BCK(cudaMemsetAsync(d_cs.p, 0, ((size_t)g.ncells + 1) * sizeof(unsigned), st));
    BCK(cudaMemsetAsync(d_ce.p, 0, ((size_t)g.ncells + 1) * sizeof(unsigned), st));
    key_kernel<<<nb, 256, 0, st>>>(n, d_aabb.p, g, d_k0.p);
    int cell_bits = 1;
    while ((1ull << cell_bits) <= (unsigned long long)g.ncells) cell_bits++;
    u64 *ka = d_k0.p, *kb = d_k1.p;
    if (radix_sort(&ka, &kb, n, 32, 32 + cell_bits, d_hist.p, d_scr.p, st)) return PS_ERR_CUDA;
    if (n_small > 0) {
        cell_range_kernel<<<(n_small + 255) / 256, 256, 0, st>>>(n_small, ka, d_cs.p, d_ce.p);
        pair_kernel<<<(n_small + PK_T - 1) / PK_T, PK_T, 0, st>>>(n_small, ka, d_aabb.p, g, d_cs.p, d_ce.p, d_p0.p, d_count.p, cap);
    }
    if (gi.n_big > 0) {
        std::vector<int> big(gi.n_big);
        BCK(cudaMemcpyAsync(big.data(), d_big.p, gi.n_big * sizeof(int), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        for (int gidx : big) big_pair_kernel<<<nb, 256, 0, st>>>(n, gidx, d_aabb.p, d_p0.p, d_count.p, cap);
    }
    BCK(cudaMemcpyAsync(&cnt, d_count.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    *n_pairs = (int64_t)cnt;
    if ((long long)cnt > pair_capacity) {
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return PS_ERR_CAPACITY;
    }
    int id_bits = 1;
    while ((1ull << id_bits) < (unsigned long long)n) id_bits++;
    u64 *pa = d_p0.p, *pb = d_p1.p;
    if (cnt > 1) {
        if (radix_sort(&pa, &pb, (int)cnt, 0, id_bits, d_hist.p, d_scr.p, st)) return PS_ERR_CUDA;
        if (radix_sort(&pa, &pb, (int)cnt, 32, 32 + id_bits, d_hist.p, d_scr.p, st)) return PS_ERR_CUDA;
    }
    if (cnt > 0) {
        split_pairs_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>((long long)cnt, pa, d_out.p);
    }
    }  // general path
    BCK(cudaEventRecord(e1, st));
    BCK(cudaEventSynchronize(e1));
    float ms = 0.f;
    BCK(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = (double)ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    BCK(cudaGetLastError());
    if (cnt > 0 && pair_ids) BCK(cudaMemcpy(pair_ids, d_out.p, 2 * (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
    return PS_OK;
}
