// ps_physics.cuh — per-body and per-pair device routines of the step:
// integration, narrow phase (sphere/sphere, sphere/box, box/box SAT + clipping), and the
// per-pair sequential-impulse solve with friction.  One thread runs one body or one pair.
//
// These follow the reference's discrete decisions and floating-point operation order exactly
// (first-wins arg-min/arg-max, reverse contact order, MathNet scalar special cases ...), so that
// candidate sets, contact features and states match the CPU restatement bit for bit.
//
// Citations are relative to /root/reference/PhysicsSimulatorCore/.
#pragma once
#include "../../include/ps_b200.h"
#include "ps_math.cuh"

namespace psp {
using namespace psm;

// -DPS_PROF_CLK: SM-cycle laps inside the pair routines of the batch-wide kernels (tools only; off in the product)
#if defined(PS_PROF_CLK) && defined(__CUDACC__)
__device__ unsigned long long g_clk_sum[32], g_clk_max[32], g_clk_cnt[32], g_clk_hist[32][32];  // hist: log2 buckets of the lap
#endif
#if defined(PS_PROF_CLK) && defined(__CUDA_ARCH__)
#define PS_CLK_BEGIN(on) long long clk_t0_ = clock64(); const bool clk_on_ = (on);
#define PS_CLK(i) { if (clk_on_) { long long t_ = clock64(); unsigned long long d_ = (unsigned long long)(t_ - clk_t0_); atomicAdd(&g_clk_sum[i], d_); atomicMax(&g_clk_max[i], d_); atomicAdd(&g_clk_cnt[i], 1ull); atomicAdd(&g_clk_hist[i][d_ ? 63 - __clzll((long long)d_) : 0], 1ull); clk_t0_ = clock64(); } }
#else
#define PS_CLK_BEGIN(on)
#define PS_CLK(i)
#endif

// step parameters in kernel-argument form
struct StepParams {
    double eps, baumgarte, slop, max_corr, dt;
    int iterations, friction, integrator;
};

// dynamic state of one body: 18 doubles (Entities/RigidBody.fs:6-8, Entities/Particle.fs:8-10)
struct Dyn {
    d3 x, v;
    m33 R;
    d3 L;
};
// constant parameters of one body (16 doubles in HBM)
struct Par {
    double inv_mass;  // 1.0 / mass (0 for +inf): what `v / mass` multiplies by (Vector3D.fs:34)
    double iinv[3];
    d3 F, T;
    double e, mu;
    double dims[3];
    int shape;  // 0 sphere, 1 box
    bool is_static;
};

struct Contact {
    d3 n, p;
    double pen;
    uint32_t feat;
};
// per-contact solver data (ContactPointImpulseData.fs:9-21)
struct CPData {
    d3 r1, r2;
    double bias, mN, mT0, mT1, accN, accT0, accT1;
};

// Box.fs:23-49 tables.  Vertex k = (sx, sy, sz) signs; kept as bit masks: bit2 = +x, bit1 = +y, bit0 = +z.
// V = (---)(-+-)(++-)(+--)(--+)(-++)(+++)(+-+)
PS_TABLE unsigned char kVertSign[8] = {0, 2, 6, 4, 1, 3, 7, 5};
PS_TABLE unsigned char kFaceVert[6][4] = {{0, 1, 2, 3}, {7, 6, 5, 4}, {5, 6, 2, 1},
                                                               {0, 3, 7, 4}, {6, 7, 3, 2}, {4, 5, 1, 0}};
// distinct-vertex order of OrientedBox.Vertices (OrientedBox.fs:25): faces 0 and 1 already cover all 8
PS_TABLE unsigned char kDistinctVert[8] = {0, 1, 2, 3, 7, 6, 5, 4};
// face normal k = sign * e_axis
PS_TABLE signed char kFaceAxis[6] = {2, 2, 1, 1, 0, 0};
PS_TABLE signed char kFaceSign[6] = {-1, 1, 1, -1, 1, -1};

// ----------------------------------------------------------------------------------------
// Integration (RigidBody.fs:202-220 -> Particle.fs:25-30 + RigidBody.fs:142-190)
// ----------------------------------------------------------------------------------------
PSI void integrate(const Dyn& b, const Par& p, const StepParams& sp, Dyn& o)
{
    const double dt = sp.dt;
    d3 acc = scale(p.inv_mass, p.F);       // RigidBody.fs:210  totalForce / mass
    d3 nv = b.v + scale(dt, acc);          // Particle.fs:27
    o.x = b.x + scale(dt, nv);             // Particle.fs:29
    o.v = nv;
    m33 Iw = inertia_inv_world(b.R, p.iinv);  // RigidBody.fs:216 (old orientation)
    d3 dL = scale(dt, p.T);
    d3 axis;
    if (sp.integrator == 0) {              // firstOrder :142-160
        d3 nL = b.L + dL;
        axis = scale(dt, mulMV(Iw, nL));
        o.L = nL;
    } else {                               // augmentedSecondOrder :162-190
        d3 w = mulMV(Iw, b.L);
        d3 vv = cross(w, b.L);             // L |> crossProduct w  = w x L
        d3 deriv = mulMV(Iw, p.T - vv);
        d3 comp1 = cross(deriv, w);        // w |> crossProduct deriv = deriv x w
        d3 avg = (w + scale(dt * 0.5, deriv)) + scale(dt * dt / 12.0, comp1);
        axis = scale(dt, avg);
        o.L = b.L + dL;
    }
    double angle = norm(axis);
    // normalize(axis) = Normalize(2.0): same norm as `angle` (same routine, same input): reuse it
    d3 u = (angle == 0.0) ? axis : scale(1.0 / angle, axis);
    m33 rot = from_axis_angle(u, angle);
    o.R = orthonormalize(mulMM(rot, b.R));
}

// The literal form as a real function: the fallback of integrate_f, entered when a lane's operands leave the range
// the straight-line arithmetic covers (a zero vector to normalise, a denormal quotient, an angle beyond pi/4, NaN).
PSN void integrate_slow(const Dyn& b, const Par& p, const StepParams& sp, Dyn& o) { integrate(b, p, sp, o); }
// Straight-line integrate (ps_math.cuh, "_f" forms): identical results wherever ok stays true.  Two halves, so that a
// caller with two bodies can put the two rotation halves — all the divisions and square roots — side by side in one
// basic block (the integrator switch of the first half is a branch, uniform but a block boundary all the same).
PSD d3 integrate_lin(const Dyn& b, const Par& p, const StepParams& sp, Dyn& o)  // x, v, L; returns the rotation vector
{
    const double dt = sp.dt;
    d3 acc = scale(p.inv_mass, p.F);
    d3 nv = b.v + scale(dt, acc);
    o.x = b.x + scale(dt, nv);
    o.v = nv;
    m33 Iw = inertia_inv_world(b.R, p.iinv);
    d3 dL = scale(dt, p.T);
    d3 axis;
    if (sp.integrator == 0) {
        d3 nL = b.L + dL;
        axis = scale(dt, mulMV(Iw, nL));
        o.L = nL;
    } else {
        d3 w = mulMV(Iw, b.L);
        d3 vv = cross(w, b.L);
        d3 deriv = mulMV(Iw, p.T - vv);
        d3 comp1 = cross(deriv, w);
        d3 avg = (w + scale(dt * 0.5, deriv)) + scale(dt * dt / 12.0, comp1);
        axis = scale(dt, avg);
        o.L = b.L + dL;
    }
    return axis;
}
// one body: straight-line first, literal on the rare miss
PSD void integrate_a(const Dyn& b, const Par& p, const StepParams& sp, Dyn& o)
{
    bool ok = true;
    const d3 axis = integrate_lin(b, p, sp, o);
    m3<double> R;
#pragma unroll
    for (int k = 0; k < 9; k++) R.a[k] = b.R.a[k];
    const m3<double> Rn = rotate_f(R, mkv<double>(axis.x, axis.y, axis.z), ok);
#pragma unroll
    for (int k = 0; k < 9; k++) o.R.a[k] = Rn.a[k];
    if (!ok) integrate_slow(b, p, sp, o);
}
// the two bodies of a pair: the rotation halves run as ONE two-lane computation (ps_math.cuh, T = dd)
PSD void integrate_2(const Dyn& b1, const Par& p1, Dyn& o1, const Dyn& b2, const Par& p2, Dyn& o2, const StepParams& sp)
{
    const d3 axis1 = integrate_lin(b1, p1, sp, o1);
    const d3 axis2 = integrate_lin(b2, p2, sp, o2);
    m3<dd> R;
#pragma unroll
    for (int k = 0; k < 9; k++) { R.a[k].a = b1.R.a[k]; R.a[k].b = b2.R.a[k]; }
    v3<dd> ax;
    ax.x.a = axis1.x; ax.y.a = axis1.y; ax.z.a = axis1.z;
    ax.x.b = axis2.x; ax.y.b = axis2.y; ax.z.b = axis2.z;
    bb ok; ok.a = true; ok.b = true;
    const m3<dd> Rn = rotate_f(R, ax, ok);
#pragma unroll
    for (int k = 0; k < 9; k++) { o1.R.a[k] = Rn.a[k].a; o2.R.a[k] = Rn.a[k].b; }
    if (!ok.a) integrate_slow(b1, p1, sp, o1);
    if (!ok.b) integrate_slow(b2, p2, sp, o2);
}

// ----------------------------------------------------------------------------------------
// Oriented box in world space (OrientedBox.fs:10-27 over Box.fs:100-111)
// ----------------------------------------------------------------------------------------
// The 8 world vertices are never stored: vertex k = R*(+-hx,+-hy,+-hz) + x evaluated as MathNet does,
//   v_i = (((0.0 + R_i0*lx) + R_i1*ly) + R_i2*lz) + x_i ,   R_i0*(+-hx) = +-(R_i0*hx) exactly,
// so the box keeps the nine products A = R col0 * hx, B = R col1 * hy, C = R col2 * hz and the centre, and a
// vertex costs four additions per component.  (The stored form was 2 x 336 B of local memory per box-box test and
// every projection re-read all of it.)
struct OBox {
    d3 A, B, C, x;
    d3 n[6];  // world face normals (Box.fs:42-49 order)
};
PSD void make_obox_i(const Dyn& b, const Par& p, OBox& ob)
{
    double hx = p.dims[0] / 2.0, hy = p.dims[1] / 2.0, hz = p.dims[2] / 2.0;
    ob.A = mk(b.R.a[0] * hx, b.R.a[3] * hx, b.R.a[6] * hx);
    ob.B = mk(b.R.a[1] * hy, b.R.a[4] * hy, b.R.a[7] * hy);
    ob.C = mk(b.R.a[2] * hz, b.R.a[5] * hz, b.R.a[8] * hz);
    ob.x = b.x;
#pragma unroll
    for (int f = 0; f < 6; f++) {
        int ax = kFaceAxis[f];
        double sg = (double)kFaceSign[f];
        d3 local = mk(ax == 0 ? sg : 0.0, ax == 1 ? sg : 0.0, ax == 2 ? sg : 0.0);
        ob.n[f] = mulMV(b.R, local) + mk(0.0, 0.0, 0.0);  // toWorldCoordinates orientation zero (OrientedBox.fs:18-21)
    }
}
PSN void make_obox(const Dyn& b, const Par& p, OBox& ob) { make_obox_i(b, p, ob); }
// world vertex k of Box.fs:23-32 (PointwiseMultiply by the half sizes, then toWorldCoordinates)
PSD d3 box_vertex(const OBox& ob, int k)
{
    unsigned s = kVertSign[k];
    d3 a = (s & 4) ? ob.A : -ob.A, b = (s & 2) ? ob.B : -ob.B, c = (s & 1) ? ob.C : -ob.C;
    return mk((((0.0 + a.x) + b.x) + c.x) + ob.x.x, (((0.0 + a.y) + b.y) + c.y) + ob.x.y, (((0.0 + a.z) + b.z) + c.z) + ob.x.z);
}
PSD d3 face_vertex(const OBox& ob, int f, int k) { return box_vertex(ob, kFaceVert[f][k]); }

struct PolyV {
    d3 p;
    unsigned tag;
};
// Working arrays of ONE pair test.  On the device they live in a per-thread record in GLOBAL memory,
// contiguous per thread: the compiler's local-memory stack is interleaved across the 32 lanes of a warp,
// so with a handful of active lanes every 8-byte access drags a 32-byte sector through the caches for 4
// useful bytes (ncu: 2.7 B per sector, DRAM traffic 44x the algorithmic bytes).
struct PairScratch {
    Contact cs[PS_MAX_CONTACTS];
    CPData cp[PS_MAX_CONTACTS];
    OBox o1, o2;
    PolyV bufA[PS_MAX_CONTACTS + 4], bufB[PS_MAX_CONTACTS + 4];
};

// min/max of the 8 vertex projections (OrientedBox.Vertices order is irrelevant: only the two values are used).
// The partial sums of the vertex formula are shared down the sign tree: 2 + 4 + 8 additions per component.
PSD void project_box_i(const OBox& ob, d3 axis, double& mn, double& mx)
{
    const d3 A = ob.A, B = ob.B, C = ob.C, x = ob.x;
    bool first = true;
    double lo = 0.0, hi = 0.0;
#pragma unroll
    for (int sa = 0; sa < 2; sa++) {
        d3 p0 = sa ? mk(0.0 + A.x, 0.0 + A.y, 0.0 + A.z) : mk(0.0 + -A.x, 0.0 + -A.y, 0.0 + -A.z);
#pragma unroll
        for (int sb = 0; sb < 2; sb++) {
            d3 p1 = sb ? p0 + B : p0 + (-B);
#pragma unroll
            for (int sc = 0; sc < 2; sc++) {
                d3 v = (sc ? p1 + C : p1 + (-C)) + x;
                double p = dot(axis, v);
                if (first) { lo = p; hi = p; first = false; }
                else {
                    if (p < lo) lo = p;
                    if (p > hi) hi = p;
                }
            }
        }
    }
    mn = lo;
    mx = hi;
}

PSN void project_box(const OBox& ob, d3 axis, double& mn, double& mx) { project_box_i(ob, axis, mn, mx); }
template <bool F> PSD void pbox(const OBox& ob, d3 axis, double& mn, double& mx)
{
    if (F) project_box_i(ob, axis, mn, mx);
    else project_box(ob, axis, mn, mx);
}

// ----------------------------------------------------------------------------------------
// GraphicsUtils (Utilities/GraphicsUtils.fs)
// ----------------------------------------------------------------------------------------
PSD d3 closest_on_edge(d3 a, d3 b, d3 pos)  // :9-24
{
    d3 ap = pos - a, ab = b - a;
    double t = dot(ab, ap) / dot(ab, ab);
    t = fmax_net(0.0, fmin_net(1.0, t));
    return a + scale(t, ab);
}
PSD d3 closest_on_quad_i(d3 pos, const d3 q[4])  // :26-37, edges (last,first),(0,1),(1,2),(2,3); first min
{
    d3 best = closest_on_edge(q[3], q[0], pos);
    d3 d = pos - best;
    double bd = dot(d, d);
#pragma unroll
    for (int k = 1; k < 4; k++) {
        d3 c = closest_on_edge(q[k - 1], q[k], pos);
        d3 e = pos - c;
        double dd = dot(e, e);
        if (dd < bd) {
            bd = dd;
            best = c;
        }
    }
    return best;
}
PSN d3 closest_on_quad(d3 pos, const d3 q[4]) { return closest_on_quad_i(pos, q); }
PSD bool plane_edge_hit_i(double eps, d3 pn, double pd, d3 s, d3 e, d3& out)  // :41-63
{
    d3 ab = e - s;
    double ab_p = dot(pn, ab);
    if (fabs(ab_p) > eps) {
        d3 p_co = scale(pd, pn);
        double fac = -(dot(pn, s - p_co)) / ab_p;
        fac = fmin_net(fmax_net(fac, 0.0), 1.0);
        out = s + scale(fac, ab);
        return true;
    }
    return false;
}
PSN bool plane_edge_hit(double eps, d3 pn, double pd, d3 s, d3 e, d3& out) { return plane_edge_hit_i(eps, pn, pd, s, e, out); }
template <bool F> PSD bool pehit(double eps, d3 pn, double pd, d3 s, d3 e, d3& out)
{
    return F ? plane_edge_hit_i(eps, pn, pd, s, e, out) : plane_edge_hit(eps, pn, pd, s, e, out);
}
template <bool F> PSD d3 cquad(d3 pos, const d3 q[4]) { return F ? closest_on_quad_i(pos, q) : closest_on_quad(pos, q); }
PSD bool in_plane(d3 pn, double pd, d3 p) { return (dot(pn, p) - pd) >= 0.0; }  // :65-66
PSD void tangents(d3 n, d3& t1, d3& t2)  // :112-123
{
    double v = sqrt(1.0 / 3.0);
    d3 t = (fabs(n.x) >= v) ? mk(n.y, n.x, 0.0) : mk(0.0, n.z, -n.y);
    t1 = normalize(t);
    t2 = cross(n, t1);
}

// ----------------------------------------------------------------------------------------
// Narrow phase
// ----------------------------------------------------------------------------------------
// detectSphereSphereCollision (CollisionDetection.fs:186-205)
template <bool F>
PSD int detect_ss_t(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, Contact* out)
{
    d3 d = b1.x - b2.x;
    double dist = norm_t<F>(d);
    double r1 = p1.dims[0], r2 = p2.dims[0];
    if (dist > r1 + r2) return 0;
    // normalize(d) = Normalize(2.0): same norm as `dist` (same routine, same input): reuse it
    d3 n = (dist == 0.0) ? d : scale(1.0 / dist, d);
    d3 cp1 = b1.x + scale(r1, -n);
    d3 cp2 = b2.x + scale(r2, n);
    d3 cp = cp1 + scale(1.0 / 2.0, cp2 - cp1);
    out[0].n = n;
    out[0].p = cp;
    out[0].pen = norm_t<F>(cp1 - cp2);
    out[0].feat = 0u;
    return 1;
}
PSN int detect_ss(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, Contact* out)
{
    return detect_ss_t<false>(b1, p1, b2, p2, out);
}

// SAT.SphereBox.tryFindSeparatingAxis (SAT.fs:108-151) + detectedSphereBoxCollision
// (CollisionDetection.fs:166-184). Normal returned "from box". err set where the reference throws.
template <bool F>
PSD int detect_sphere_box_t(const Dyn& sb, const Par& sp_, const Dyn& bb, const Par& bp, double eps, Contact* out,
                            int& face, int& err, OBox& ob)
{
    if (F) make_obox_i(bb, bp, ob);
    else make_obox(bb, bp, ob);
    double r = sp_.dims[0];
    double best_pen = 0.0;
    int best = -1;
    d3 nb = mk(0.0, 0.0, 0.0);
    // The six face normals come in +- pairs (f, f+1) and n[f+1] == 0.0 - n[f] bit for bit (Box.fs:42-49, both are
    // R*(+-e_k) + 0).  A dot product against the negated axis is 0.0 - (the dot product): every product flips
    // sign and a zero-started sum is never -0.0.  So one projection pass serves both axes of a pair.
#pragma unroll (F ? 3 : 1)
    for (int fp = 0; fp < 6; fp += 2) {
        d3 axis = ob.n[fp];
        double a0, b0;
        pbox<F>(ob, axis, a0, b0);
        double c0 = dot(axis, sb.x);
#pragma unroll (F ? 2 : 1)
        for (int h = 0; h < 2; h++) {
            double a = h ? 0.0 - b0 : a0, b = h ? 0.0 - a0 : b0, c = h ? 0.0 - c0 : c0;
            double cp = c + r, cm = c - r;
            double c1 = (cm < cp) ? cm : cp;  // List.min [c+r; c-r]
            double c2 = (cm > cp) ? cm : cp;  // List.max
            double pen;
            if (c1 >= a && c1 <= b)
                pen = b - c1;
            else if (c1 <= a && c2 >= a)
                pen = c2 - a;
            else
                return 0;
            if (best < 0 || pen < best_pen) {
                best = fp + h;
                best_pen = pen;
                nb = ob.n[fp + h];
            }
        }
    }
    d3 endP = sb.x - scale(r, nb);
    double pd = dot(nb, face_vertex(ob, best, 0));  // Face.toPlane (Entities/Base.fs:27-29)
    d3 pos;
    if (!pehit<F>(eps, nb, pd, sb.x, endP, pos)) {
        err = 1;  // Option.get of None
        return 0;
    }
    out[0].n = nb;
    out[0].p = pos;
    out[0].pen = best_pen;
    face = best;
    return 1;
}
PSN int detect_sphere_box(const Dyn& sb, const Par& sp_, const Dyn& bb, const Par& bp, double eps, Contact* out,
                          int& face, int& err, OBox& ob)
{
    return detect_sphere_box_t<false>(sb, sp_, bb, bp, eps, out, face, err, ob);
}

// SAT.tryGetCollisionDataForAxis (SAT.fs:155-183) on precomputed projection ranges
PSD bool sat_ranges(double a, double b, double c, double d, double& pen, bool& flip)
{
    if (a <= c && b >= c) {
        pen = b - c;
        flip = false;
        return true;
    }
    if (c <= a && d >= a) {
        pen = d - a;
        flip = true;
        return true;
    }
    return false;
}


// Early "no collision" for box-box: any ONE separated axis of the reference's list makes the result None
// (SAT.fs:185-198 `sequence`), whichever it is.  Tries the face axis of each box that looks most along the centre
// offset, with the reference's own arithmetic for that axis.  true = proven separated (exactly what the full walk
// would conclude); false = undecided.
template <bool F>
PSD bool bb_quick_separated_t(const Dyn& b1, const Dyn& b2, const OBox& o1, const OBox& o2)
{
    d3 fromFirst = b2.x - b1.x;
#pragma unroll (F ? 2 : 1)
    for (int side = 0; side < 2; side++) {
        const Dyn& bb = side == 0 ? b1 : b2;
        int k = 0;
        double bestd = -1.0;
#pragma unroll (F ? 3 : 1)
        for (int q = 0; q < 3; q++) {
            double dq = fabs(bb.R.a[q] * fromFirst.x + bb.R.a[3 + q] * fromFirst.y + bb.R.a[6 + q] * fromFirst.z);
            if (dq > bestd) { bestd = dq; k = q; }
        }
        const OBox& T = side == 0 ? o1 : o2;
        const OBox& O = side == 0 ? o2 : o1;
        const d3 axis = (k == 0) ? T.n[4] : (k == 1 ? T.n[2] : T.n[1]);  // the +e_k face (Box.fs:42-49)
        double ta, tb, oc, od, pen;
        bool flip;
        pbox<F>(T, axis, ta, tb);
        pbox<F>(O, axis, oc, od);
        if (!sat_ranges(ta, tb, oc, od, pen, flip)) return true;
    }
    return false;
}
PSN bool bb_quick_separated(const Dyn& b1, const Dyn& b2, const OBox& o1, const OBox& o2)
{
    return bb_quick_separated_t<false>(b1, b2, o1, o2);
}

// generateEdgeContactPoints (CollisionDetection.fs:75-132): the supporting edge of each box along the SAT edge axis
// (sel1 / sel2 = RigidBody.Axes entries of the winning pair of edges), the midpoint of their closest points.
// returns 1 (one contact) or 0 with err set (List.maxBy on an empty list)
template <bool F>
PSD int bb_edge_contact(const OBox& o1, const OBox& o2, d3 sel1, d3 sel2, d3 normal, double best_pen, int best_i, int best_j,
                        double eps, Contact* out, int& err)
{
    d3 ea[2], eb[2];
    int eidx[2];
#pragma unroll (F ? 2 : 1)
    for (int side = 0; side < 2; side++) {
        const OBox& ob = side == 0 ? o1 : o2;
        d3 axis = side == 0 ? sel1 : sel2;
        d3 nrm = side == 0 ? normal : -normal;
        bool have = false;
        double bestv = 0.0;
        int idx = 0;
        d3 bea = mk(0.0, 0.0, 0.0), beb = mk(0.0, 0.0, 0.0);
        // OrientedBox.getEdges: faces 0..5, per face (v3,v0),(v0,v1),(v1,v2),(v2,v3); List.distinct on
        // ordered pairs — every directed edge of a non-degenerate box appears once, so index = 4*f+k.
        d3 vv[8];  // the box's vertices once (every one is an end of six of the 24 directed edges)
#pragma unroll (F ? 8 : 1)
        for (int q = 0; q < 8; q++) vv[q] = box_vertex(ob, q);
#pragma unroll 1
        for (int f = 0; f < 6; f++)
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                d3 a = vv[kFaceVert[f][(k + 3) & 3]], b = vv[kFaceVert[f][k]];
                d3 dd = b - a;
                d3 cr = cross(dd, axis);  // areParallel: axis |> crossProduct d = d x axis, isZero eps of its norm
                // Exact shortcut: Hypotenuse never returns less than its larger argument (|x| * sqrt(1 + r*r) >= |x|), so
                // the norm is >= every |component|; one component above eps decides "not parallel" without the two
                // divisions and square roots (20 of the 24 edges end here; a NaN norm compares false either way).
                if (fabs(cr.x) > eps || fabs(cr.y) > eps || fabs(cr.z) > eps) continue;
                // and the other way round: with every |component| <= eps / 2 the norm is at most sqrt(3) / 2 eps (1 + 8 ulp) < eps,
                // "parallel" without the norm as well.  What is left (a component in (eps / 2, eps]) does not occur for a box
                // edge against a box axis (the cross product is ~1e-16 or an edge length); without this a warp of the batch-wide
                // kernel ran the two divisions and square roots once per edge for ONE lane each (ncu: 60 % of this routine).
                const double heps = 0.5 * eps;  // (exact above the subnormals)
                const bool tiny = eps > 1e-290 && fabs(cr.x) <= heps && fabs(cr.y) <= heps && fabs(cr.z) <= heps;
                if (!tiny && !near_eq(eps, 0.0, norm_t<F>(cr))) continue;
                double pa = dot(nrm, a), pb = dot(nrm, b);
                double key = (pb > pa) ? pb : pa;
                if (!have || key > bestv) {
                    have = true;
                    bestv = key;
                    idx = 4 * f + k;
                    bea = a;
                    beb = b;
                }
            }
        if (!have) {
            err = 1;  // List.maxBy on an empty list
            return 0;
        }
        eidx[side] = idx;
        ea[side] = bea;
        eb[side] = beb;
    }
    d3 DA = eb[0] - ea[0], DB = eb[1] - ea[1];
    d3 r = ea[0] - ea[1];
    double a = dot(DA, DA), e = dot(DB, DB), f = dot(DB, r), c = dot(DA, r), b = dot(DA, DB);
    double denom = a * e - b * b;
    double TA = (b * f - c * e) / denom;
    double TB = (b * TA + f) / e;
    d3 CA = ea[0] + scale(TA, DA);
    d3 CB = ea[1] + scale(TB, DB);
    out[0].n = normal;
    out[0].p = scale(0.5, CA) + scale(0.5, CB);
    out[0].pen = -best_pen;
    out[0].feat = 3u | (2u << 2) | ((unsigned)best_i << 4) | ((unsigned)best_j << 7) | ((unsigned)eidx[0] << 10) |
                  ((unsigned)eidx[1] << 15);
    return 1;
}

// detectBoxBoxCollision (CollisionDetection.fs:29-164) over SAT.tryFindSeparatingAxis (SAT.fs:206-240)
// returns #contacts (0 = none); overflow set when the manifold would exceed PS_MAX_CONTACTS
// boxes_ready: sc.o1/sc.o2 are already built and the quick test has already failed to separate them
template <bool F>
PSD int detect_bb_t(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, Contact* out, int& err,
                    int& overflow, PairScratch& sc, bool boxes_ready, bool prof = false)
{
    PS_CLK_BEGIN(prof)
    // F: the boxes are values of this frame (registers once everything below is unrolled to static indices)
    OBox l1, l2;
    OBox& o1 = F ? l1 : sc.o1;
    OBox& o2 = F ? l2 : sc.o2;
    if (F) {
        make_obox_i(b1, p1, o1);
        make_obox_i(b2, p2, o2);
    } else if (!boxes_ready) {
        make_obox(b1, p1, o1);
        make_obox(b2, p2, o2);
    }

    int best_origin = -1, best_i = 0, best_j = 0;
    double best_pen = 0.0;
    d3 best_n = mk(0, 0, 0);
    d3 fromFirst = b2.x - b1.x;

    if (!boxes_ready && bb_quick_separated_t<F>(b1, b2, o1, o2)) return 0;

    // Face1 then Face2 (SAT.fs:230-234).  Normals come in +- pairs (f, f+1) with n[f+1] == 0.0 - n[f] bit for
    // bit, and a projection onto the negated axis is 0.0 - projection (see detect_sphere_box): one pass per pair.
#pragma unroll (F ? 2 : 1)
    for (int origin = 0; origin < 2; origin++) {
        const OBox& T = origin == 0 ? o1 : o2;  // Face2: boxes flipped
        const OBox& O = origin == 0 ? o2 : o1;
#pragma unroll (F ? 3 : 1)
        for (int fp = 0; fp < 6; fp += 2) {
            d3 axis0 = T.n[fp];
            double a0, b0, c0, d0;
            pbox<F>(T, axis0, a0, b0);
            pbox<F>(O, axis0, c0, d0);
#pragma unroll (F ? 2 : 1)
            for (int h = 0; h < 2; h++) {
                double a = h ? 0.0 - b0 : a0, b = h ? 0.0 - a0 : b0;
                double c = h ? 0.0 - d0 : c0, d = h ? 0.0 - c0 : d0;
                double pen;
                bool flip;
                if (!sat_ranges(a, b, c, d, pen, flip)) return 0;
                if (best_origin < 0 || pen < best_pen) {
                    d3 axis = T.n[fp + h];
                    best_origin = origin;
                    best_pen = pen;
                    best_n = flip ? -axis : axis;
                }
            }
        }
    }
    PS_CLK(0)
    // Edge axes (:218-226): RigidBody.Axes = normalize(R e_k + 0) (RigidBody.fs:34-41)
    d3 ax1[3], ax2[3];
#pragma unroll (F ? 3 : 1)
    for (int k = 0; k < 3; k++) {
        ax1[k] = normalize_t<F>(mk(b1.R.a[k], b1.R.a[3 + k], b1.R.a[6 + k]) + mk(0.0, 0.0, 0.0));
        ax2[k] = normalize_t<F>(mk(b2.R.a[k], b2.R.a[3 + k], b2.R.a[6 + k]) + mk(0.0, 0.0, 0.0));
    }
#pragma unroll (F ? 3 : 1)
    for (int i = 0; i < 3; i++)
#pragma unroll (F ? 3 : 1)
        for (int j = 0; j < 3; j++) {
            // normalized(second x first), dropped iff isZero 0.01 of the NORMALISED vector (:223-225): that norm is
            // ~1 unless the cross product is exactly zero (then Normalize returns the zero vector), so the test is
            // `norm(cross) == 0` (finite inputs)
            d3 cr = cross(ax2[j], ax1[i]);
            double cn = norm_t<F>(cr);
            if (cn == 0.0) continue;
            d3 e = scale(1.0 / cn, cr);
            if (dot(e, fromFirst) < 0.0) e = -e;
            double a, b, c, d, pen;
            bool flip;
            pbox<F>(o1, e, a, b);
            pbox<F>(o2, e, c, d);
            if (!sat_ranges(a, b, c, d, pen, flip)) return 0;
            if (pen < best_pen) {
                best_origin = 2;
                best_pen = pen;
                best_n = flip ? -e : e;
                best_i = i;
                best_j = j;
            }
        }

    PS_CLK(1)
    if (best_origin < 2) {
        // generateFaceContactPoints (CollisionDetection.fs:30-73)
        OBox rv, ov;  // F: reference / other box by value (a select per field instead of a pointer into memory)
        if (F) {
            rv = o1; ov = o2;
            if (best_origin != 0) { rv = o2; ov = o1; }
        }
        const OBox& ref = F ? rv : ((best_origin == 0) ? o1 : o2);
        const OBox& oth = F ? ov : ((best_origin == 0) ? o2 : o1);
        d3 normal = best_n;
        int rf = -1;
        d3 rn = mk(0.0, 0.0, 0.0);
#pragma unroll (F ? 6 : 1)
        for (int f = 0; f < 6; f++)
            if (rf < 0 && veq(ref.n[f], normal)) {  // Box.findFaceByNormal (Box.fs:113-114): the first match
                rf = f;
                rn = ref.n[f];
            }
        if (rf < 0) {
            err = 1;
            return 0;
        }
        int inc = 0;
        double incv = dot(rn, oth.n[0]);
#pragma unroll (F ? 5 : 1)
        for (int f = 1; f < 6; f++) {  // findIncidentFace: first min (:35-37)
            double d = dot(rn, oth.n[f]);
            if (d < incv) {
                incv = d;
                inc = f;
            }
        }
        // clip the incident quad by the inverted planes of the adjacent faces (Box.fs:116-124, :56-61)
        PolyV* cur = sc.bufA;
        PolyV* nxt = sc.bufB;
        int n = 4;
#pragma unroll (F ? 4 : 1)
        for (int k = 0; k < 4; k++) {
            cur[k].p = face_vertex(oth, inc, k);
            cur[k].tag = (unsigned)k;
        }
        int pi = 0;
        PS_CLK(2)
#pragma unroll (F ? 6 : 1)
        for (int f = 0; f < 6; f++) {
            double ad = fabs(dot(rn, ref.n[f]));
            if (near_eq(eps, 1.0, ad)) continue;  // not adjacent
            if (n == 0) break;  // "clipping error": the reference prints and returns the empty list
            d3 pn = -ref.n[f];                                   // Plane.inverted
            double pd = -dot(ref.n[f], face_vertex(ref, f, 0));  // Face.toPlane then inverted
            int m = 0;
#pragma unroll 1
            for (int k = 0; k < n; k++) {
                const PolyV& s = (k == 0) ? cur[n - 1] : cur[k - 1];
                const PolyV& e = cur[k];
                bool sin_ = in_plane(pn, pd, s.p), ein = in_plane(pn, pd, e.p);
                unsigned xtag = 0x80u | ((unsigned)(pi & 3) << 4) | (unsigned)(k & 15);
                d3 x;
                if (sin_ && ein) {
                    if (m < PS_MAX_CONTACTS + 4) nxt[m++] = e; else overflow = 1;
                } else if (sin_ && !ein) {
                    if (pehit<F>(eps, pn, pd, s.p, e.p, x)) {
                        if (m < PS_MAX_CONTACTS + 4) { nxt[m].p = x; nxt[m].tag = xtag; m++; } else overflow = 1;
                    }
                } else if (!sin_ && ein) {
                    if (m < PS_MAX_CONTACTS + 4) nxt[m++] = e; else overflow = 1;
                    if (pehit<F>(eps, pn, pd, s.p, e.p, x)) {
                        if (m < PS_MAX_CONTACTS + 4) { nxt[m].p = x; nxt[m].tag = xtag; m++; } else overflow = 1;
                    }
                }
            }
            PolyV* t = cur; cur = nxt; nxt = t;
            n = m;
            pi++;
        }
        PS_CLK(3)
        // List.distinct, then keep points on/below the reference plane (:52-54), then contacts (:63-67)
        d3 refq[4];
#pragma unroll (F ? 4 : 1)
        for (int k = 0; k < 4; k++) refq[k] = face_vertex(ref, rf, k);
        d3 rpn = -rn;
        double rpd = -dot(rn, refq[0]);
        int nc = 0;
#pragma unroll 1
        for (int k = 0; k < n; k++) {
            bool dup = false;
#pragma unroll 1
            for (int q = 0; q < k; q++)
                if (veq(cur[q].p, cur[k].p)) {
                    dup = true;
                    break;
                }
            if (dup) continue;
            if (!in_plane(rpn, rpd, cur[k].p)) continue;
            if (nc >= PS_MAX_CONTACTS) {
                overflow = 1;
                break;
            }
            d3 cl = cquad<F>(cur[k].p, refq);
            out[nc].pen = dot(normal, cur[k].p - cl);
            out[nc].n = (best_origin == 1) ? -normal : normal;  // WithInvertedNormals (:157)
            out[nc].p = cur[k].p;
            out[nc].feat = 3u | ((unsigned)best_origin << 2) | ((unsigned)rf << 4) | ((unsigned)inc << 7) |
                           ((unsigned)(nc & 31) << 10) | (cur[k].tag << 15);
            nc++;
        }
        PS_CLK(4)
        return nc;
    }

    // generateEdgeContactPoints (CollisionDetection.fs:75-132)
    const d3 sel1 = best_i == 0 ? ax1[0] : (best_i == 1 ? ax1[1] : ax1[2]);
    const d3 sel2 = best_j == 0 ? ax2[0] : (best_j == 1 ? ax2[1] : ax2[2]);
    const int ne = bb_edge_contact<F>(o1, o2, sel1, sel2, best_n, best_pen, best_i, best_j, eps, out, err);
    PS_CLK(5)
    return ne;
}
PSN int detect_bb(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, Contact* out, int& err,
                  int& overflow, PairScratch& sc, bool boxes_ready = false)
{
    return detect_bb_t<false>(b1, p1, b2, p2, eps, out, err, overflow, sc, boxes_ready);
}

// areColliding for a pair with at least one sphere: exactly one contact, no clip buffers needed
template <bool F>
PSD int detect_sphere_pair_t(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, Contact* out, int& err,
                             OBox& ob)
{
    if (p1.shape == 0 && p2.shape == 0) return F ? detect_ss_t<true>(b1, p1, b2, p2, out) : detect_ss(b1, p1, b2, p2, out);
    int face = 0;
    if (p1.shape == 0) {
        int n = F ? detect_sphere_box_t<true>(b1, p1, b2, p2, eps, out, face, err, ob)
                  : detect_sphere_box(b1, p1, b2, p2, eps, out, face, err, ob);
        if (n) {
            out[0].n = -out[0].n;  // WithInvertedNormals (CollisionDetection.fs:217-219)
            out[0].feat = 1u | ((unsigned)face << 4);
        }
        return n;
    }
    int n = F ? detect_sphere_box_t<true>(b2, p2, b1, p1, eps, out, face, err, ob)
              : detect_sphere_box(b2, p2, b1, p1, eps, out, face, err, ob);
    if (n) out[0].feat = 2u | ((unsigned)face << 4);
    return n;
}
PSD int detect_sphere_pair(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, Contact* out, int& err,
                           OBox& ob)
{
    return detect_sphere_pair_t<false>(b1, p1, b2, p2, eps, out, err, ob);
}

// CollisionDetection.areColliding (CollisionDetection.fs:208-220)
template <bool F>
PSD int detect_t(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, Contact* out, int& err,
                 int& overflow, PairScratch& sc)
{
    if (p1.shape == 1 && p2.shape == 1)
        return F ? detect_bb_t<true>(b1, p1, b2, p2, eps, out, err, overflow, sc, false)
                 : detect_bb(b1, p1, b2, p2, eps, out, err, overflow, sc);
    return detect_sphere_pair_t<F>(b1, p1, b2, p2, eps, out, err, sc.o1);
}
PSD int detect(const Dyn& b1, const Par& p1, const Dyn& b2, const Par& p2, double eps, Contact* out, int& err,
               int& overflow, PairScratch& sc)
{
    return detect_t<false>(b1, p1, b2, p2, eps, out, err, overflow, sc);
}

// ----------------------------------------------------------------------------------------
// Per-pair solve (CollisionResponse.fs:73-146, ContactPointImpulseData.fs:23-66, Friction.fs:39-68)
// ----------------------------------------------------------------------------------------

// hat(r)^T * Iw * hat(r), products evaluated as dense 3x3 (RigidBody.fs:225-229)
PSD m33 hat(d3 v)
{
    m33 o;
    o.a[0] = 0.0;  o.a[1] = -v.z; o.a[2] = v.y;
    o.a[3] = v.z;  o.a[4] = 0.0;  o.a[5] = -v.x;
    o.a[6] = -v.y; o.a[7] = v.x;  o.a[8] = 0.0;
    return o;
}
PSD m33 transpose(const m33& m)
{
    m33 o;
    o.a[0] = m.a[0]; o.a[1] = m.a[3]; o.a[2] = m.a[6];
    o.a[3] = m.a[1]; o.a[4] = m.a[4]; o.a[5] = m.a[7];
    o.a[6] = m.a[2]; o.a[7] = m.a[5]; o.a[8] = m.a[8];
    return o;
}
PSD void add_K(m33& s, const Par& p, const m33& Iw, d3 r)  // inlined: measured C2 2.71 -> 2.53 ms
{
    if (p.is_static) {  // Mass.Infinite -> Matrix3.zero: s + 0
#pragma unroll
        for (int k = 0; k < 9; k++) s.a[k] = s.a[k] + 0.0;
        return;
    }
    m33 h = hat(r);
    m33 t = mulMM(mulMM(transpose(h), Iw), h);
    t.a[0] = t.a[0] + p.inv_mass;
    t.a[4] = t.a[4] + p.inv_mass;
    t.a[8] = t.a[8] + p.inv_mass;
#pragma unroll
    for (int k = 0; k < 9; k++) s.a[k] = s.a[k] + t.a[k];
}
// RigidBodyMotion.applyImpulse (RigidBody.fs:194-197)
PSD void apply_impulse(Dyn& b, const Par& p, d3 off, d3 P)
{
    b.v = b.v + scale(p.inv_mass, P);
    b.L = b.L + cross(off, P);
}
PSD d3 velocity_at(const Dyn& b, const m33& Iw, d3 off) { return b.v + cross(mulMV(Iw, b.L), off); }

// ContactPointImpulseData.init (ContactPointImpulseData.fs:23-52) + calculateBaumgarteBias (CollisionResponse.fs:35-41)
// for ONE contact of a manifold; t0/t1 are the tangents of the manifold's FIRST contact (CollisionResponse.fs:80-81)
PSD void contact_setup(const Dyn& b1, const Par& p1, const m33& Iw1, const Dyn& b2, const Par& p2, const m33& Iw2,
                       const Contact& c, d3 t0, d3 t1, const StepParams& sp, CPData& d)
{
    d.bias = fmin_net(sp.max_corr, -sp.baumgarte / sp.dt * fmin_net(0.0, c.pen + sp.slop));  // :35-41
    d.r1 = c.p - b1.x;
    d.r2 = c.p - b2.x;
    m33 K;
#pragma unroll
    for (int q = 0; q < 9; q++) K.a[q] = 0.0;
    add_K(K, p1, Iw1, d.r1);
    add_K(K, p2, Iw2, d.r2);
    d.mN = dot(c.n, mulMV(K, c.n));
    d.mT0 = dot(t0, mulMV(K, t0));
    d.mT1 = dot(t1, mulMV(K, t1));
    d.accN = 0.0;
    d.accT0 = 0.0;
    d.accT1 = 0.0;
}
// one contact of one solver iteration (CollisionResponse.fs:91-131: normal impulse, then friction, Friction.fs:39-68);
// the velocity state of the two bodies (v, L) and the contact's accumulators are updated in place
struct SolveConst {
    double e, mu, inf;
    int friction;
};
PSD void solve_contact(Dyn& b1, const Par& p1, const m33& Iw1, Dyn& b2, const Par& p2, const m33& Iw2, d3 n, d3 t0, d3 t1,
                       CPData& d, const SolveConst& sc)
{
    d3 P;
    if (d.mN == 0.0) {
        P = mk(0.0, 0.0, 0.0);
    } else {
        d3 vrel = velocity_at(b2, Iw2, d.r2) - velocity_at(b1, Iw1, d.r1);
        double vn = dot(vrel, n);
        double j = (-(sc.e + 1.0) * vn + d.bias) / d.mN;
        double old = d.accN;
        d.accN = clamp_net(0.0, sc.inf, old + j);
        P = scale(-(d.accN - old), n);
    }
    apply_impulse(b1, p1, d.r1, P);
    apply_impulse(b2, p2, d.r2, -P);
    if (sc.friction) {
        d3 vrel = velocity_at(b2, Iw2, d.r2) - velocity_at(b1, Iw1, d.r1);
        double maxF = sc.mu * d.accN;
        double j0 = dot(t0, vrel) / d.mT0;
        double j1 = dot(t1, vrel) / d.mT1;
        double o0 = d.accT0, o1 = d.accT1;
        d.accT0 = clamp_net(-maxF, maxF, o0 + j0);
        d.accT1 = clamp_net(-maxF, maxF, o1 + j1);
        d3 P0 = scale(d.accT0 - o0, t0);
        d3 P1 = scale(d.accT1 - o1, t1);
        apply_impulse(b1, p1, d.r1, P0);
        apply_impulse(b1, p1, d.r1, P1);
        apply_impulse(b2, p2, d.r2, -P0);
        apply_impulse(b2, p2, d.r2, -P1);
    }
}
PSD SolveConst solve_const(const Par& p1, const Par& p2, const StepParams& sp)
{
    SolveConst sc;
    sc.e = fmin_net(p1.e, p2.e);
    sc.mu = fmax_net(p2.mu, p1.mu);
    sc.inf = from_bits(0x7ff0000000000000LL);
    sc.friction = sp.friction;
    return sc;
}

// resolves one pair in place; b1/b2 are the predicted bodies
// FIX = 0: nc contacts at run time; FIX = 1: exactly one contact (every sphere pair), arrays collapse to registers
template <int FIX>
PSD void resolve_t(Dyn& b1, const Par& p1, Dyn& b2, const Par& p2, const Contact* cs, int nc_in, CPData* cp,
                   const StepParams& sp)
{
    const int nc = FIX ? FIX : nc_in;
    // orientation is constant during the solve, so R Ip^-1 R^T is evaluated once per body (exact)
    m33 Iw1 = inertia_inv_world(b1.R, p1.iinv);
    m33 Iw2 = inertia_inv_world(b2.R, p2.iinv);
    d3 t0, t1;
    tangents(cs[0].n, t0, t1);  // CollisionResponse.fs:80-81
#pragma unroll 1
    for (int k = 0; k < nc; k++) {
        CPData d;
        contact_setup(b1, p1, Iw1, b2, p2, Iw2, cs[k], t0, t1, sp, d);
        cp[k] = d;
    }
    const SolveConst sc = solve_const(p1, p2, sp);
#pragma unroll 1
    for (int it = 0; it < sp.iterations; it++) {
#pragma unroll 1
        for (int k = nc - 1; k >= 0; k--) {  // List.foldBack: reverse order
            CPData d = cp[k];
            solve_contact(b1, p1, Iw1, b2, p2, Iw2, cs[k].n, t0, t1, d, sc);
            cp[k].accN = d.accN;
            cp[k].accT0 = d.accT0;
            cp[k].accT1 = d.accT1;
        }
    }
}

// ----------------------------------------------------------------------------------------
// The solve of the batch-wide kernels: resolve_t's operations on resolve_t's operands in resolve_t's order, arranged for
// the latency of ONE thread (a round of the batch lasts as long as its slowest manifold; measured before: ~4 K cycles per
// contact-iteration, i.e. ~400 FP64 operations at their full dependent latency — the zero tests of scale(), the clamps
// and the division fall-backs cut the iteration into a dozen basic blocks, and the assembler interleaves independent
// operations only inside a block):
//   * ONE basic block per contact-iteration: every data-dependent decision is a select (scale_s, fmin_s/fmax_s, the
//     "mN == 0 -> no impulse" case), no call, no branch;
//   * the three divisions of a contact-iteration divide by the contact's constant masses: their reciprocals are refined
//     once per contact (recip_seed) and a division is the last three operations of the device's own inline sequence
//     (div_by_s); its acceptance test is folded into `ok` instead of branching to the complete IEEE routine;
//   * a contact's solver data (incl. its normal) sits in one record; the NEXT contact's record is loaded while this one
//     is being solved;
//   * S1 = body 1 is a static body at rest (1/m = 0, I^-1 = 0, v = +0: the ground of every scene).  Its velocity at
//     any point is +0 for as long as its angular-momentum accumulator and the contact offsets stay finite ((+-0) *
//     finite sums to +0, and x - (+0) = x for every x), so vRel = velocity_at(body 2); its v stays +0
//     (v + scale(0, P) = v); only the accumulator L1 (quirk Q2) still takes every impulse, in order.
// Returns false when a quotient left the range the short division covers, or (S1) L1 / r1 left the finite range: the
// results are then NOT the reference's, and the caller redoes the pair with the literal routine from the same inputs.
// On the host (tests/host_harness) div_by_s is a plain division and the selects are the same selects.
// ----------------------------------------------------------------------------------------
struct CPFast {
    d3 n, r1, r2;
    double bias, mN, mT0, mT1, yN, yT0, yT1, accN, accT0, accT1;
};
PSD d3 scale_s(double s, d3 v)  // scale() with its zero case as selects
{
    const bool z = (s == 0.0);
    const double px = s * v.x, py = s * v.y, pz = s * v.z;
    return mk(z ? 0.0 : px, z ? 0.0 : py, z ? 0.0 : pz);
}
PSD double fmin_s(double a, double b)  // fmin_net
{
    const bool c = (a != b) && (a == a);
    const double r1 = a < b ? a : b, r2 = is_neg(a) ? a : b;
    return c ? r1 : r2;
}
PSD double fmax_s(double a, double b)  // fmax_net
{
    const double r1 = (a == a) ? (b < a ? a : b) : a, r2 = is_neg(b) ? a : b;
    return (a != b) ? r1 : r2;
}
PSD double div_by_s(double a, double b, double y, bool& ok)
{
#if defined(__CUDA_ARCH__)
    double q = __dmul_rn(a, y);
    const double r = __fma_rn(-b, q, a);
    q = __fma_rn(y, r, q);
    // a numerator that is exactly zero (bodies at rest: a relative velocity of exactly 0 along a tangent) is the one
    // frequent case the short sequence does not cover; its quotient is a zero with the sign of a xor the sign of b
    // for every b that is neither zero nor NaN
    const bool zero_num = (a == 0.0) && (b != 0.0) && (b == b);
    const double qz = from_bits((bits_of(a) ^ bits_of(b)) & (int64_t)0x8000000000000000ULL);
    ok = ok && (zero_num || div_ok(a, b, q));
    return zero_num ? qz : q;
#else
    (void)y; (void)ok;
    return a / b;
#endif
}
struct SolveS {
    d3 t0, t1;
    double im1, im2, e1, mu, inf;  // e1 = -(min e + 1)
    bool friction;                 // (uniform over the batch: the one branch of a contact-iteration)
};
template <bool S1>
PSD void solve_contact_s(d3& v1, d3& L1, d3& v2, d3& L2, const m33& Iw1, const m33& Iw2, const SolveS& c, CPFast& d, bool& ok)
{
    const bool live = !(d.mN == 0.0);  // CollisionResponse.fs:45-49: mN = 0 -> no normal impulse
    d3 vrel = v2 + cross(mulMV(Iw2, L2), d.r2);
    if (!S1) vrel = vrel - (v1 + cross(mulMV(Iw1, L1), d.r1));
    const double vn = dot(vrel, d.n);
    bool okq = true;
    const double j = div_by_s(c.e1 * vn + d.bias, d.mN, d.yN, okq);
    ok = ok && (okq || !live);
    const double old = d.accN;
    const double acc = fmax_s(0.0, fmin_s(old + j, c.inf));
    d3 P = scale_s(-(acc - old), d.n);
    d.accN = live ? acc : old;
    P = mk(live ? P.x : 0.0, live ? P.y : 0.0, live ? P.z : 0.0);
    if (!S1) v1 = v1 + scale_s(c.im1, P);
    L1 = L1 + cross(d.r1, P);
    d3 nP = -P;
    v2 = v2 + scale_s(c.im2, nP);
    L2 = L2 + cross(d.r2, nP);
    if (c.friction) {
        d3 vr = v2 + cross(mulMV(Iw2, L2), d.r2);
        if (!S1) vr = vr - (v1 + cross(mulMV(Iw1, L1), d.r1));
        const double maxF = c.mu * d.accN;
        const double j0 = div_by_s(dot(c.t0, vr), d.mT0, d.yT0, ok);
        const double j1 = div_by_s(dot(c.t1, vr), d.mT1, d.yT1, ok);
        const double o0 = d.accT0, o1 = d.accT1;
        d.accT0 = fmax_s(-maxF, fmin_s(o0 + j0, maxF));
        d.accT1 = fmax_s(-maxF, fmin_s(o1 + j1, maxF));
        const d3 P0 = scale_s(d.accT0 - o0, c.t0);
        const d3 P1 = scale_s(d.accT1 - o1, c.t1);
        if (!S1) {
            v1 = v1 + scale_s(c.im1, P0);
            v1 = v1 + scale_s(c.im1, P1);
        }
        L1 = L1 + cross(d.r1, P0);
        L1 = L1 + cross(d.r1, P1);
        const d3 n0 = -P0, n1 = -P1;
        v2 = v2 + scale_s(c.im2, n0);
        L2 = L2 + cross(d.r2, n0);
        v2 = v2 + scale_s(c.im2, n1);
        L2 = L2 + cross(d.r2, n1);
    }
}
PSD bool finite3(d3 v)
{
    const double big = 1.7976931348623157e308;
    return fabs(v.x) <= big && fabs(v.y) <= big && fabs(v.z) <= big;
}
// FIX = 1: exactly one contact, no array
template <int FIX, bool S1>
PSD bool resolve_s(Dyn& b1, const Par& p1, Dyn& b2, const Par& p2, const Contact* cs, int nc_in, CPFast* cp, const StepParams& sp)
{
    const int nc = FIX ? FIX : nc_in;
    const m33 Iw1 = inertia_inv_world(b1.R, p1.iinv);
    const m33 Iw2 = inertia_inv_world(b2.R, p2.iinv);
    SolveS c;
    tangents(cs[0].n, c.t0, c.t1);  // CollisionResponse.fs:80-81
    {
        const SolveConst sc = solve_const(p1, p2, sp);
        c.im1 = p1.inv_mass; c.im2 = p2.inv_mass; c.e1 = -(sc.e + 1.0); c.mu = sc.mu; c.inf = sc.inf; c.friction = sc.friction != 0;
    }
    bool ok = true;
    CPFast d;
#pragma unroll 1
    for (int k = 0; k < nc; k++) {
        CPData q;
        const Contact ck = cs[k];
        contact_setup(b1, p1, Iw1, b2, p2, Iw2, ck, c.t0, c.t1, sp, q);
        d.n = ck.n; d.r1 = q.r1; d.r2 = q.r2;
        d.bias = q.bias; d.mN = q.mN; d.mT0 = q.mT0; d.mT1 = q.mT1;
        d.yN = recip_seed(q.mN); d.yT0 = recip_seed(q.mT0); d.yT1 = recip_seed(q.mT1);
        d.accN = 0.0; d.accT0 = 0.0; d.accT1 = 0.0;
        if (S1) ok = ok && finite3(d.r1);
        if (!FIX) cp[k] = d;
    }
    d3 v1 = b1.v, L1 = b1.L, v2 = b2.v, L2 = b2.L;
    if (FIX == 1) {  // d is the contact
#pragma unroll 1
        for (int it = 0; it < sp.iterations; it++) solve_contact_s<S1>(v1, L1, v2, L2, Iw1, Iw2, c, d, ok);
    } else {
        // The lanes of a warp hold manifolds of different sizes.  The record array lives in local memory, which is
        // interleaved across the lanes: the SAME index in every lane is one contiguous access, a different index per lane
        // is 32 scattered sectors (ncu: 4.7 useful bytes per 32-byte sector with `k = nc - 1 ... 0` per lane).  So the
        // walk is over kk = (largest manifold of the warp) - 1 ... 0 for everybody and a lane takes part while kk < nc:
        // its own contacts still come up in List.foldBack order nc - 1 ... 0.
#if defined(__CUDA_ARCH__)
        const int ncmax = __reduce_max_sync(__activemask(), nc);
#else
        const int ncmax = nc;
#endif
        d = cp[ncmax - 1];
#pragma unroll 1
        for (int it = 0; it < sp.iterations; it++) {
#pragma unroll 1
            for (int kk = ncmax - 1; kk >= 0; kk--) {
                const CPFast nx = cp[kk > 0 ? kk - 1 : ncmax - 1];  // the next record of the walk, loaded while this one is solved
                if (kk < nc) {
                    solve_contact_s<S1>(v1, L1, v2, L2, Iw1, Iw2, c, d, ok);
                    cp[kk].accN = d.accN;
                    cp[kk].accT0 = d.accT0;
                    cp[kk].accT1 = d.accT1;
                }
                if (ncmax > 1) d = nx;  // (a single record stays where it is, accumulators included: nx was read before the stores)
            }
        }
    }
    if (S1) ok = ok && finite3(L1);
    if (!S1) b1.v = v1;
    b1.L = L1; b2.v = v2; b2.L = L2;
    return ok;
}

#if defined(__CUDACC__)
// ---- the same solve on TWO neighbouring lanes of a warp: lane `second == false` carries body 1, the other body 2 ------------
// What a round of the batch-wide step waits for is ONE thread's walk through the iterations of its largest manifold (measured,
// manifolds of 7 and more contacts: 148 K cycles of iterations, ~430 instructions per contact-iteration, an FP64 instruction
// costs a warp two issue cycles whatever its lanes).  Three quarters of a contact-iteration is the same code on the data of
// body 1 and of body 2 (velocity at the contact point, applying the impulse): each lane does it for its own body, the two
// point velocities are exchanged by shuffles (vRel = u2 - u1 on both lanes), the impulse arithmetic in between is done by both
// lanes on identical operands.  Every operation is an operation of resolve_s on the same operands: same bits.
struct CPFast2 {
    d3 n, r;  // r: the contact offset of the lane's own body
    double bias, mN, mT0, mT1, yN, yT0, yT1, accN, accT0, accT1;
};
PSD d3 shfl_pair(unsigned m, d3 v) { return mk(__shfl_xor_sync(m, v.x, 1), __shfl_xor_sync(m, v.y, 1), __shfl_xor_sync(m, v.z, 1)); }
PSD d3 sel3(bool c, d3 a, d3 b) { return mk(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }
PSD void solve_contact_s2(unsigned m, bool second, d3& v, d3& L, const m33& Iw, double im, const SolveS& c, CPFast2& d, bool& ok)
{
    const bool live = !(d.mN == 0.0);
    const d3 u = v + cross(mulMV(Iw, L), d.r);
    const d3 uo = shfl_pair(m, u);
    const d3 vrel = sel3(second, u, uo) - sel3(second, uo, u);  // velocity_at(body 2) - velocity_at(body 1)
    const double vn = dot(vrel, d.n);
    bool okq = true;
    const double j = div_by_s(c.e1 * vn + d.bias, d.mN, d.yN, okq);
    ok = ok && (okq || !live);
    const double old = d.accN;
    const double acc = fmax_s(0.0, fmin_s(old + j, c.inf));
    d3 P = scale_s(-(acc - old), d.n);
    d.accN = live ? acc : old;
    P = mk(live ? P.x : 0.0, live ? P.y : 0.0, live ? P.z : 0.0);
    const d3 Pl = sel3(second, -P, P);  // body 1 takes P, body 2 takes -P
    v = v + scale_s(im, Pl);
    L = L + cross(d.r, Pl);
    if (c.friction) {
        const d3 w = v + cross(mulMV(Iw, L), d.r);
        const d3 wo = shfl_pair(m, w);
        const d3 vr = sel3(second, w, wo) - sel3(second, wo, w);
        const double maxF = c.mu * d.accN;
        const double j0 = div_by_s(dot(c.t0, vr), d.mT0, d.yT0, ok);
        const double j1 = div_by_s(dot(c.t1, vr), d.mT1, d.yT1, ok);
        const double o0 = d.accT0, o1 = d.accT1;
        d.accT0 = fmax_s(-maxF, fmin_s(o0 + j0, maxF));
        d.accT1 = fmax_s(-maxF, fmin_s(o1 + j1, maxF));
        const d3 P0 = scale_s(d.accT0 - o0, c.t0);
        const d3 P1 = scale_s(d.accT1 - o1, c.t1);
        const d3 Q0 = sel3(second, -P0, P0), Q1 = sel3(second, -P1, P1);
        v = v + scale_s(im, Q0);
        v = v + scale_s(im, Q1);
        L = L + cross(d.r, Q0);
        L = L + cross(d.r, Q1);
    }
}
// act: the lanes of the warp that are in this call (pairs of neighbours, all converged at the call)
PSD bool resolve_s2(unsigned act, bool second, Dyn& b1, const Par& p1, Dyn& b2, const Par& p2, const Contact* cs, int nc, CPFast2* cp,
                    const StepParams& sp)
{
    const m33 Iw1 = inertia_inv_world(b1.R, p1.iinv);
    const m33 Iw2 = inertia_inv_world(b2.R, p2.iinv);
    SolveS c;
    tangents(cs[0].n, c.t0, c.t1);
    {
        const SolveConst sc = solve_const(p1, p2, sp);
        c.im1 = p1.inv_mass; c.im2 = p2.inv_mass; c.e1 = -(sc.e + 1.0); c.mu = sc.mu; c.inf = sc.inf; c.friction = sc.friction != 0;
    }
    bool ok = true;
    CPFast2 d;
#pragma unroll 1
    for (int k = 0; k < nc; k++) {  // (both lanes set up every contact: the effective masses need both bodies)
        CPData q;
        const Contact ck = cs[k];
        contact_setup(b1, p1, Iw1, b2, p2, Iw2, ck, c.t0, c.t1, sp, q);
        d.n = ck.n; d.r = sel3(second, q.r2, q.r1);
        d.bias = q.bias; d.mN = q.mN; d.mT0 = q.mT0; d.mT1 = q.mT1;
        d.yN = recip_seed(q.mN); d.yT0 = recip_seed(q.mT0); d.yT1 = recip_seed(q.mT1);
        d.accN = 0.0; d.accT0 = 0.0; d.accT1 = 0.0;
        cp[k] = d;
    }
    m33 Iw;
#pragma unroll
    for (int q = 0; q < 9; q++) Iw.a[q] = second ? Iw2.a[q] : Iw1.a[q];
    const double im = second ? p2.inv_mass : p1.inv_mass;
    d3 v = sel3(second, b2.v, b1.v), L = sel3(second, b2.L, b1.L);
#if defined(__CUDA_ARCH__)
    const int ncmax = __reduce_max_sync(act, nc);  // the walk of resolve_s: one index for all lanes
#else
    const int ncmax = nc;
#endif
    d = cp[ncmax - 1];
#pragma unroll 1
    for (int it = 0; it < sp.iterations; it++) {
#pragma unroll 1
        for (int kk = ncmax - 1; kk >= 0; kk--) {
            const CPFast2 nx = cp[kk > 0 ? kk - 1 : ncmax - 1];
            const unsigned m = __ballot_sync(act, kk < nc);  // (the two lanes of a pair hold the same nc)
            if (kk < nc) {
                solve_contact_s2(m, second, v, L, Iw, im, c, d, ok);
                cp[kk].accN = d.accN;
                cp[kk].accT0 = d.accT0;
                cp[kk].accT1 = d.accT1;
            }
            if (ncmax > 1) d = nx;
        }
    }
    ok = ok && (__shfl_xor_sync(act, ok ? 1 : 0, 1) != 0);
    if (second) { b2.v = v; b2.L = L; }
    else { b1.v = v; b1.L = L; }
    return ok;
}
#endif

PSN void resolve(Dyn& b1, const Par& p1, Dyn& b2, const Par& p2, const Contact* cs, int nc, CPData* cp,
                 const StepParams& sp)
{
    resolve_t<0>(b1, p1, b2, p2, cs, nc, cp, sp);
}

}  // namespace psp
