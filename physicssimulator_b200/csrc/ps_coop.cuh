// ps_coop.cuh — box-box narrow phase + contact generation by a TEAM of 8 lanes per pair (ps_team.cuh).
//
// detectBoxBoxCollision (CollisionDetection.fs:29-164) over SAT.tryFindSeparatingAxis (SAT.fs:206-240), the same
// operations on the same operands as psp::detect_bb_t, spread over lanes:
//   phase 0  the 12 face normals and the 6 RigidBody.Axes (RigidBody.fs:34-41) — one or two per lane, into shared memory
//   phase 1  the 15 SAT "axis jobs" (3 +- pairs of face normals per box, 9 edge cross axes), two per lane: projection of
//            both boxes, overlap test; arg-min over the team on (penetration, candidate index), lower index wins
//            (the reference keeps the FIRST minimum: SAT.fs:185-198)
//   phase 2  face case: Sutherland-Hodgman clipping (GraphicsUtils.fs:69-107) one polygon EDGE per lane, output order kept
//            by a team prefix sum; then one clipped vertex per lane: List.distinct, below-the-reference-plane filter,
//            closest point on the reference quad, feature id
//            edge case (3 % of the manifolds): lane 0 runs the literal routine (psp::bb_edge_contact)
// Everything whose order matters (candidate order, polygon order, contact ordinals) is decided by index, never by
// arrival, so results are the literal routine's bit for bit.  Inputs the team form does not cover — a NaN penetration
// (the sequential "first minimum" scan is not a reduction then) or a clipped polygon of more than 16 vertices — are
// reported back (need_literal) and the caller runs psp::detect_bb on them.
#pragma once
#include "ps_physics.cuh"
#include "ps_team.cuh"

namespace psc {
using namespace psp;
using pstm::Team;

PSD int popc(unsigned v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
constexpr int kTeam = 8;      // lanes per pair
constexpr int kPolyCap = 16;  // clipped polygon vertices the team form holds (two per lane)

struct TeamScratch {  // shared memory of one team
    d3 nrm[12];       // world face normals of box 1 (0..5) and box 2 (6..11), Box.fs:42-49 order
    d3 ax[6];         // RigidBody.Axes of box 1 (0..2) and box 2 (3..5)
    PolyV poly[2][kPolyCap];
};

// the box without its normals (they live in the team's shared memory)
PSD void make_obox_core(const Dyn& b, const Par& p, OBox& ob)
{
    double hx = p.dims[0] / 2.0, hy = p.dims[1] / 2.0, hz = p.dims[2] / 2.0;
    ob.A = mk(b.R.a[0] * hx, b.R.a[3] * hx, b.R.a[6] * hx);
    ob.B = mk(b.R.a[1] * hy, b.R.a[4] * hy, b.R.a[7] * hy);
    ob.C = mk(b.R.a[2] * hz, b.R.a[5] * hz, b.R.a[8] * hz);
    ob.x = b.x;
}
PSD d3 box_normal(const m33& R, int f)  // as make_obox_i: R * (+-e_axis) + 0
{
    const int ax = kFaceAxis[f];
    const double sg = (double)kFaceSign[f];
    d3 local = mk(ax == 0 ? sg : 0.0, ax == 1 ? sg : 0.0, ax == 2 ? sg : 0.0);
    return mulMV(R, local) + mk(0.0, 0.0, 0.0);
}
// vertex k of face f without the constant tables being indexed twice
PSD d3 face_vertex_c(const OBox& ob, int f, int k) { return box_vertex(ob, kFaceVert[f][k]); }

// What one lane holds when the team is done: up to two contacts (polygon vertices lane and lane + 8) and their
// ordinals in the manifold (-1: none).  nc = contacts of the manifold (same on every lane).
struct TeamContacts {
    Contact c[2];
    int ord[2];
};

// returns the number of contacts (0 = no collision); every lane of the team calls it with the same arguments
template <class TM>
PSD int detect_bb_team(const TM& tm, const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, TeamScratch& ts,
                       TeamContacts& out, int& err, bool& need_literal)
{
    const int lane = tm.lane;
    need_literal = false;
    out.ord[0] = out.ord[1] = -1;
    OBox o1, o2;
    make_obox_core(b1, p1, o1);
    make_obox_core(b2, p2, o2);
    // ---- phase 0: normals and axes ------------------------------------------------------------------------------------
#pragma unroll 1
    for (int q = lane; q < 12; q += kTeam) ts.nrm[q] = box_normal(q < 6 ? b1.R : b2.R, q < 6 ? q : q - 6);
    if (lane < 6) {
        const m33& R = lane < 3 ? b1.R : b2.R;
        const int k = lane < 3 ? lane : lane - 3;
        ts.ax[lane] = normalize_t<true>(mk(R.a[k], R.a[3 + k], R.a[6 + k]) + mk(0.0, 0.0, 0.0));
    }
    tm.sync();
    const d3 fromFirst = b2.x - b1.x;

    // ---- phase 1: SAT.  job 0..2: box 1 normal pairs, 3..5: box 2 normal pairs (boxes flipped), 6..14: edge axes ------
    // candidate index (the reference's axis order, SAT.fs:228-238): face job j -> 2j, 2j+1; edge job j -> j + 6
    double bpen = 0.0;
    int bidx = -1;
    d3 bn = mk(0.0, 0.0, 0.0);  // the winning axis, already flipped
    bool sep = false, bad = false;
#pragma unroll 1
    for (int t = 0; t < 2; t++) {
        const int job = lane + kTeam * t;
        if (job >= 15) break;
        d3 axis;
        bool valid = true;
        const bool face = job < 6;
        if (face) {
            axis = ts.nrm[(job / 3) * 6 + 2 * (job % 3)];
        } else {
            const int i = (job - 6) / 3, j = (job - 6) % 3;
            d3 cr = cross(ts.ax[3 + j], ts.ax[i]);
            double cn = norm_t<true>(cr);
            if (cn == 0.0) {
                valid = false;
                axis = cr;
            } else {
                d3 e = scale(1.0 / cn, cr);
                if (dot(e, fromFirst) < 0.0) e = -e;
                axis = e;
            }
        }
        if (!valid) continue;
        double a1, b1_, a2, b2_;
        project_box_i(o1, axis, a1, b1_);
        project_box_i(o2, axis, a2, b2_);
        if (face) {
            const bool flipped = job >= 3;  // Face2: T = box 2, O = box 1
            const double a0 = flipped ? a2 : a1, b0 = flipped ? b2_ : b1_, c0 = flipped ? a1 : a2, d0 = flipped ? b1_ : b2_;
            const int nbase = (job / 3) * 6 + 2 * (job % 3);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double a = h ? 0.0 - b0 : a0, b = h ? 0.0 - a0 : b0;
                double c = h ? 0.0 - d0 : c0, d = h ? 0.0 - c0 : d0;
                double pen;
                bool flip;
                if (!sat_ranges(a, b, c, d, pen, flip)) { sep = true; continue; }
                if (pen != pen) bad = true;
                if (bidx < 0 || pen < bpen) {
                    const d3 ax = ts.nrm[nbase + h];
                    bidx = 2 * job + h;
                    bpen = pen;
                    bn = flip ? -ax : ax;
                }
            }
        } else {
            double pen;
            bool flip;
            if (!sat_ranges(a1, b1_, a2, b2_, pen, flip)) { sep = true; continue; }
            if (pen != pen) bad = true;
            if (bidx < 0 || pen < bpen) {
                bidx = job + 6;
                bpen = pen;
                bn = flip ? -axis : axis;
            }
        }
    }
    if (tm.any(sep)) return 0;  // one separated axis: None, whichever axis it is (SAT.fs:185-198 `sequence`)
    if (tm.any(bad)) { need_literal = true; return 0; }
    // arg-min over the team on (penetration, candidate index); the lower index wins ties = the first minimum
    {
        double rp = bpen;
        int ri = bidx;
#pragma unroll
        for (int o = kTeam / 2; o; o >>= 1) {
            const double op = tm.shfl_xor(rp, o);
            const int oi = tm.shfl_xor(ri, o);
            const bool take = (oi >= 0) && (ri < 0 || op < rp || (op == rp && oi < ri));
            if (take) { rp = op; ri = oi; }
        }
        const int job = ri < 12 ? ri / 2 : ri - 6;
        const int wl = job & (kTeam - 1);  // the lane that evaluated it; its local best IS the winner
        bn.x = tm.shfl(bn.x, wl); bn.y = tm.shfl(bn.y, wl); bn.z = tm.shfl(bn.z, wl);
        bpen = rp;
        bidx = ri;
    }
    const int origin = bidx < 6 ? 0 : (bidx < 12 ? 1 : 2);
    const d3 normal = bn;

    if (origin == 2) {
        // generateEdgeContactPoints (CollisionDetection.fs:75-132): lane 0, literal
        const int bi = (bidx - 12) / 3, bj = (bidx - 12) % 3;
        int nc = 0, e = 0;
        if (lane == 0) {
            nc = bb_edge_contact<false>(o1, o2, ts.ax[bi], ts.ax[3 + bj], normal, bpen, bi, bj, eps, out.c, e);
            if (nc) out.ord[0] = 0;
        }
        nc = tm.shfl(nc, 0);
        e = tm.shfl(e, 0);
        if (e) err = 1;
        return nc;
    }

    // ---- phase 2: generateFaceContactPoints (CollisionDetection.fs:30-73) ------------------------------------------------
    OBox ref = o1, oth = o2;
    if (origin != 0) { ref = o2; oth = o1; }
    const d3* refn = ts.nrm + (origin == 0 ? 0 : 6);
    const d3* othn = ts.nrm + (origin == 0 ? 6 : 0);
    int rf = -1;
    d3 rn = mk(0.0, 0.0, 0.0);
#pragma unroll 1
    for (int f = 0; f < 6; f++)
        if (rf < 0 && veq(refn[f], normal)) {  // Box.findFaceByNormal (Box.fs:113-114): the first match
            rf = f;
            rn = refn[f];
        }
    if (rf < 0) {
        err = 1;
        return 0;
    }
    int inc = 0;
    double incv = dot(rn, othn[0]);
#pragma unroll 1
    for (int f = 1; f < 6; f++) {  // findIncidentFace: first min (:35-37)
        double d = dot(rn, othn[f]);
        if (d < incv) {
            incv = d;
            inc = f;
        }
    }
    PolyV* cur = ts.poly[0];
    PolyV* nxt = ts.poly[1];
    int n = 4;
    if (lane < 4) {
        cur[lane].p = face_vertex_c(oth, inc, lane);
        cur[lane].tag = (unsigned)lane;
    }
    tm.sync();
    int pi = 0;
    bool too_big = false;
#pragma unroll 1
    for (int f = 0; f < 6; f++) {
        double ad = fabs(dot(rn, refn[f]));
        if (near_eq(eps, 1.0, ad)) continue;  // not adjacent
        if (n == 0) break;                     // "clipping error": the reference prints and returns the empty list
        const d3 pn = -refn[f];                                        // Plane.inverted
        const double pd = -dot(refn[f], face_vertex_c(ref, f, 0));     // Face.toPlane then inverted
        // edge k = (cur[k-1], cur[k]) emits 0, 1 or 2 vertices; lanes take edges lane and lane + 8
        PolyV o_[2][2];
        int cnt[2] = {0, 0};
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int k = lane + kTeam * t;
            if (k >= n) continue;
            const PolyV s = (k == 0) ? cur[n - 1] : cur[k - 1];
            const PolyV e = cur[k];
            const bool sin_ = in_plane(pn, pd, s.p), ein = in_plane(pn, pd, e.p);
            const unsigned xtag = 0x80u | ((unsigned)(pi & 3) << 4) | (unsigned)(k & 15);
            d3 x;
            int m = 0;
            if (sin_ && ein) {
                o_[t][m++] = e;
            } else if (sin_ && !ein) {
                if (plane_edge_hit_i(eps, pn, pd, s.p, e.p, x)) { o_[t][m].p = x; o_[t][m].tag = xtag; m++; }
            } else if (!sin_ && ein) {
                o_[t][m++] = e;
                if (plane_edge_hit_i(eps, pn, pd, s.p, e.p, x)) { o_[t][m].p = x; o_[t][m].tag = xtag; m++; }
            }
            cnt[t] = m;
        }
        // output order = edge order: all edges k < 8 (in lane order), then the edges k >= 8
        const int inc0 = pstm::team_inclusive_scan<kTeam>(tm, cnt[0]);
        const int tot0 = tm.shfl(inc0, kTeam - 1);
        const int inc1 = pstm::team_inclusive_scan<kTeam>(tm, cnt[1]);
        const int tot1 = tm.shfl(inc1, kTeam - 1);
        const int total = tot0 + tot1;
        if (total > kPolyCap) { too_big = true; break; }
        int w0 = inc0 - cnt[0], w1 = tot0 + inc1 - cnt[1];
        for (int q = 0; q < cnt[0]; q++) nxt[w0 + q] = o_[0][q];
        for (int q = 0; q < cnt[1]; q++) nxt[w1 + q] = o_[1][q];
        tm.sync();
        PolyV* sw = cur; cur = nxt; nxt = sw;
        n = total;
        pi++;
    }
    if (too_big) { need_literal = true; return 0; }

    // List.distinct, then keep points on/below the reference plane (:52-54), then contacts (:63-67); vertex per lane
    d3 refq[4];
#pragma unroll
    for (int k = 0; k < 4; k++) refq[k] = face_vertex_c(ref, rf, k);
    const d3 rpn = -rn;
    const double rpd = -dot(rn, refq[0]);
    bool keep[2] = {false, false};
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int k = lane + kTeam * t;
        if (k >= n) continue;
        const d3 pk = cur[k].p;
        bool dup = false;
        for (int q = 0; q < k; q++)
            if (veq(cur[q].p, pk)) { dup = true; break; }
        keep[t] = !dup && in_plane(rpn, rpd, pk);
    }
    const unsigned k0 = tm.ballot(keep[0]), k1 = tm.ballot(keep[1]);
    const int nc_all = popc(k0) + popc(k1);
#pragma unroll
    for (int t = 0; t < 2; t++) {
        if (!keep[t]) continue;
        const int k = lane + kTeam * t;
        const int ord = t == 0 ? popc(k0 & ((1u << lane) - 1u)) : popc(k0) + popc(k1 & ((1u << lane) - 1u));
        if (ord >= PS_MAX_CONTACTS) continue;  // (cannot happen with 16 vertices; the literal routine truncates there)
        const d3 pk = cur[k].p;
        const d3 cl = closest_on_quad_i(pk, refq);
        Contact& c = out.c[t];
        c.pen = dot(normal, pk - cl);
        c.n = (origin == 1) ? -normal : normal;  // WithInvertedNormals (:157)
        c.p = pk;
        c.feat = 3u | ((unsigned)origin << 2) | ((unsigned)rf << 4) | ((unsigned)inc << 7) | ((unsigned)(ord & 31) << 10) | (cur[k].tag << 15);
        out.ord[t] = ord;
    }
    tm.sync();  // the polygon buffers are free again
    return nc_all < PS_MAX_CONTACTS ? nc_all : PS_MAX_CONTACTS;
}

}  // namespace psc
