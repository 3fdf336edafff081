// oracle/ps_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, scalar, one world at a time) of the per-step hot path of
// Mic43/PhysicsSimulator's PhysicsSimulatorCore.  It is the parity oracle for the CUDA
// stepper and the "port" CPU baseline of bench.py.  It is NOT part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.
//
// PARITY STATUS: the reference is F#/.NET and cannot be built or run in this image (no
// dotnet/mono).  Its own tests pin only the octree (PhysicsSimulator.Tests/
// UtilitiesTests.fs:159-276) — those five known-answer facts are checked against this
// file's SpatialTree in tests/test_oracle_octree.py.  Integration, narrow phase, contact
// generation, solver, friction and joints are PARITY UNPINNED: no golden vector exists in
// the reference for them; this file follows the source line by line instead.
//
// Every function cites the reference lines it follows, relative to
// /root/reference/PhysicsSimulatorCore/.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off; no fast-math).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../include/ps_b200.h"
#include "ref_math.hpp"

using namespace ref;

namespace {

// ------------------------------------------------------------------------------------
// World model (Entities/RigidBody.fs:6-17, Entities/Particle.fs:8-14, SimulatorObject.fs:41-44,
// SimulatorState.fs:11-19)
// ------------------------------------------------------------------------------------
struct Body {
    int shape;  // 0 sphere, 1 box   (Entities/Collider.fs:5-7)
    double dims[3];
    double mass;  // +inf == Mass.Infinite
    double iinv[3];
    V3 force, torque;
    double e, mu_k, mu_s;
    V3 x, v;
    M3 R;
    V3 L;
    bool is_static() const { return std::isinf(mass); }  // SimulatorObject.fs:33-36
};

struct Contact {
    V3 n, p;
    double pen;
    uint32_t feat;
};

struct Joint {
    int a, b;
    V3 la, lb;
};

struct Face {
    V3 v[4];
    V3 n;
};
struct OBox {
    Face f[6];
    std::vector<V3> verts;  // List.distinct of all face vertices, first-occurrence order
};

// Box.fs:23-49
const double kVerts[8][3] = {{-1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {1, -1, -1},
                             {-1, -1, 1},  {-1, 1, 1},  {1, 1, 1},  {1, -1, 1}};
const int kFaces[6][4] = {{0, 1, 2, 3}, {7, 6, 5, 4}, {5, 6, 2, 1}, {0, 3, 7, 4}, {6, 7, 3, 2}, {4, 5, 1, 0}};
const double kNormals[6][3] = {{0, 0, -1}, {0, 0, 1}, {0, 1, 0}, {0, -1, 0}, {1, 0, 0}, {-1, 0, 0}};

// GraphicsUtils.toWorldCoordinates (GraphicsUtils.fs:109-110): R*v then + t
inline V3 to_world(const M3& R, V3 t, V3 v) { return add(mulMV(R, v), t); }

// OrientedBox.create (OrientedBox.fs:10-27) over Box.getFaces (Box.fs:100-111)
OBox make_obox(const Body& b)
{
    OBox ob;
    V3 scale = v3(b.dims[0] / 2.0, b.dims[1] / 2.0, b.dims[2] / 2.0);
    for (int f = 0; f < 6; f++) {
        for (int k = 0; k < 4; k++) {
            const double* lv = kVerts[kFaces[f][k]];
            V3 local = v3(lv[0] * scale.x, lv[1] * scale.y, lv[2] * scale.z);  // PointwiseMultiply
            ob.f[f].v[k] = to_world(b.R, b.x, local);
        }
        ob.f[f].n = to_world(b.R, v3(0.0, 0.0, 0.0), v3(kNormals[f][0], kNormals[f][1], kNormals[f][2]));
    }
    for (int f = 0; f < 6; f++)
        for (int k = 0; k < 4; k++) {
            V3 p = ob.f[f].v[k];
            bool seen = false;
            for (auto& q : ob.verts)
                if (veq(p, q)) {
                    seen = true;
                    break;
                }
            if (!seen) ob.verts.push_back(p);
        }
    return ob;
}

// RigidBody.CalcRotationalInertiaInverse (RigidBody.fs:19-26): (R * Ip^-1) * R^T
M3 inertia_inv_world(const Body& b) { return mulMM(mulMDiag(b.R, b.iinv), transpose(b.R)); }
// RigidBody.CalcAngularVelocity (RigidBody.fs:31-32)
V3 angular_velocity(const Body& b) { return mulMV(inertia_inv_world(b), b.L); }
// RigidBody.Axes (RigidBody.fs:34-41)
void body_axes(const Body& b, V3 out[3])
{
    const V3 e[3] = {v3(1, 0, 0), v3(0, 1, 0), v3(0, 0, 1)};
    for (int k = 0; k < 3; k++) out[k] = normalize(to_world(b.R, v3(0, 0, 0), e[k]));
}

// ------------------------------------------------------------------------------------
// Integration: SimulatorObject.update (SimulatorObject.fs:82-97) -> RigidBodyMotion.update
// (RigidBody.fs:202-220) -> forwardEuler (Particle.fs:25-30) + firstOrder (RigidBody.fs:142-160)
// / augmentedSecondOrder (:162-190)
// ------------------------------------------------------------------------------------
Body integrate(const Body& b, double dt, int integrator_kind)
{
    Body o = b;
    V3 acc = sdiv(b.force, b.mass);                // RigidBody.fs:210
    V3 nv = add(b.v, smul(dt, acc));               // Particle.fs:27
    o.x = add(b.x, smul(dt, nv));                  // Particle.fs:29
    o.v = nv;
    M3 Iw = inertia_inv_world(b);                  // RigidBody.fs:216 (old orientation)
    V3 dL = smul(dt, b.torque);                    // :146 / :166
    V3 axis;
    if (integrator_kind == 0) {
        V3 nL = add(b.L, dL);                      // :148
        V3 w = mulMV(Iw, nL);                      // :150
        axis = smul(dt, w);                        // :152
        o.L = nL;
    } else {
        V3 w = mulMV(Iw, b.L);                     // :170
        V3 vv = cross(w, b.L);                     // :172  oldAngMomentum |> crossProduct angularVelocity
        V3 deriv = mulMV(Iw, sub(b.torque, vv));   // :174
        V3 comp1 = cross(deriv, w);                // :176  angularVelocity |> crossProduct deriv
        // :178-181  w + dt*0.5*deriv + (dt*dt/12)*comp1   (left-assoc: (dt*0.5)*deriv)
        V3 avg = add(add(w, smul(dt * 0.5, deriv)), smul(dt * dt / 12.0, comp1));
        axis = smul(dt, avg);                      // :183
        o.L = add(b.L, dL);                        // :190
    }
    double angle = norm(axis);                     // :153 / :184
    M3 rot = from_axis_angle(normalize(axis), angle);
    o.R = orthonormalize(mulMM(rot, b.R));         // :155-159
    return o;
}

// RigidBodyMotion.applyImpulse (RigidBody.fs:194-197) + ParticleMotion.applyImpulse (Particle.fs:34-36)
void apply_impulse(Body& b, V3 offset, V3 impulse)
{
    b.v = add(b.v, sdiv(impulse, b.mass));
    b.L = add(b.L, cross(offset, impulse));  // impulse |> crossProduct offset
}

// RigidBodyMotion.calculateTranslationConstraintMass (RigidBody.fs:222-234)
M3 constraint_mass(const Body& b1, const Body& b2, V3 r1, V3 r2)
{
    auto K = [](const Body& b, V3 r) -> M3 {
        if (b.is_static()) return zeroM();
        M3 h = hat(r);
        M3 t = mulMM(mulMM(transpose(h), inertia_inv_world(b)), h);
        double im = 1.0 / b.mass;  // GetInverseMassMatrix (Particle.fs:16-17), a DiagonalMatrix
        for (int i = 0; i < 3; i++) t.m[i][i] = t.m[i][i] + im;
        return t;
    };
    M3 s = zeroM();
    s = addMM(s, K(b1, r1));
    s = addMM(s, K(b2, r2));
    return s;
}
// calculateVelocityAtOffset / calculateRelativeVelocity (RigidBody.fs:236-248)
V3 velocity_at(const Body& b, V3 off) { return add(b.v, cross(angular_velocity(b), off)); }
V3 relative_velocity(const Body& b1, const Body& b2, V3 r1, V3 r2) { return sub(velocity_at(b2, r2), velocity_at(b1, r1)); }

// ------------------------------------------------------------------------------------
// GraphicsUtils (Utilities/GraphicsUtils.fs)
// ------------------------------------------------------------------------------------
struct Plane {
    V3 n;
    double d;
};
inline Plane face_plane(const Face& f) { return Plane{f.n, dot(f.n, f.v[0])}; }   // Entities/Base.fs:27-29
inline Plane plane_inverted(Plane p) { return Plane{neg(p.n), -p.d}; }            // Plane.fs:16-17

// getClosestPointToEdge (GraphicsUtils.fs:9-24)
V3 closest_point_to_edge(V3 a, V3 b, V3 pos)
{
    V3 ap = sub(pos, a);
    V3 ab = sub(b, a);
    double abap = dot(ab, ap);
    double mag = dot(ab, ab);
    double dist = abap / mag;
    dist = fs_max(0.0, fs_min(1.0, dist));  // (distance |> min 1.0) |> max 0.0
    return add(a, smul(dist, ab));
}
// getClosestPointToPoly (GraphicsUtils.fs:26-37): edges (last,first)::pairwise, first min wins
V3 closest_point_to_poly(V3 pos, const V3* poly, int n)
{
    double best = 0.0;
    V3 bestp = pos;
    for (int k = 0; k < n; k++) {
        V3 s = (k == 0) ? poly[n - 1] : poly[k - 1];
        V3 e = (k == 0) ? poly[0] : poly[k];
        V3 c = closest_point_to_edge(s, e, pos);
        V3 d = sub(pos, c);
        double dd = dot(d, d);
        if (k == 0 || dd < best) {
            best = dd;
            bestp = c;
        }
    }
    return bestp;
}
// tryGetPlaneEdgeIntersection (GraphicsUtils.fs:41-63)
bool plane_edge_intersection(double eps, Plane pl, V3 s, V3 e, V3* out)
{
    V3 ab = sub(e, s);
    double ab_p = dot(pl.n, ab);
    if (std::fabs(ab_p) > eps) {
        V3 p_co = smul(pl.d, pl.n);
        double fac = -(dot(pl.n, sub(s, p_co))) / ab_p;
        fac = fs_min(fs_max(fac, 0.0), 1.0);
        *out = add(s, smul(fac, ab));
        return true;
    }
    return false;
}
inline bool point_in_plane(Plane pl, V3 p) { return (dot(pl.n, p) - pl.d) >= 0.0; }  // :65-66

struct PV {  // polygon vertex with provenance tag (feature ids; not in the reference)
    V3 p;
    uint8_t tag;
};
// SutherlandHodgmanClipping (GraphicsUtils.fs:69-107)
std::vector<PV> sh_clip(double eps, const std::vector<Plane>& planes, std::vector<PV> poly, bool* clip_error)
{
    for (size_t pi = 0;; pi++) {
        if (poly.empty()) {
            *clip_error = true;  // printfn "clipping error"
            return poly;
        }
        if (pi == planes.size()) break;
        Plane pl = planes[pi];
        std::vector<PV> outp;
        int n = (int)poly.size();
        for (int k = 0; k < n; k++) {
            PV s = (k == 0) ? poly[n - 1] : poly[k - 1];
            PV e = (k == 0) ? poly[0] : poly[k];
            bool sin_ = point_in_plane(pl, s.p), ein = point_in_plane(pl, e.p);
            V3 x;
            uint8_t xtag = (uint8_t)(0x80 | ((pi & 3) << 4) | (k & 15));
            if (sin_ && ein) {
                outp.push_back(e);
            } else if (sin_ && !ein) {
                if (plane_edge_intersection(eps, pl, s.p, e.p, &x)) outp.push_back(PV{x, xtag});
            } else if (!sin_ && ein) {
                outp.push_back(e);  // endPoint :: intersection  (Q9)
                if (plane_edge_intersection(eps, pl, s.p, e.p, &x)) outp.push_back(PV{x, xtag});
            }
        }
        poly.swap(outp);
    }
    // List.distinct (position equality, first occurrence kept)
    std::vector<PV> d;
    for (auto& q : poly) {
        bool seen = false;
        for (auto& r : d)
            if (veq(q.p, r.p)) {
                seen = true;
                break;
            }
        if (!seen) d.push_back(q);
    }
    return d;
}
// computeTangentVectors (GraphicsUtils.fs:112-123)
void tangent_vectors(V3 n, V3* t1, V3* t2)
{
    double v = std::sqrt(1.0 / 3.0);
    V3 t = (std::fabs(n.x) >= v) ? v3(n.y, n.x, 0.0) : v3(0.0, n.z, -n.y);
    *t1 = normalize(t);
    *t2 = cross(n, *t1);  // tangent.Get |> crossProduct normal
}

// ------------------------------------------------------------------------------------
// Narrow phase
// ------------------------------------------------------------------------------------
// detectSphereSphereCollision (CollisionDetection.fs:186-205)
bool detect_ss(const Body& b1, const Body& b2, std::vector<Contact>* out)
{
    V3 nrm = sub(b1.x, b2.x);
    double dist = norm(nrm);
    double r1 = b1.dims[0], r2 = b2.dims[0];
    if (dist > r1 + r2) return false;
    V3 n = normalize(nrm);
    V3 cp1 = add(b1.x, smul(r1, neg(n)));
    V3 cp2 = add(b2.x, smul(r2, n));
    V3 cp = add(cp1, sdiv(sub(cp2, cp1), 2.0));
    double pen = norm(sub(cp1, cp2));
    out->push_back(Contact{n, cp, pen, 0u});
    return true;
}

// SAT.SphereBox.tryFindSeparatingAxis (SAT.fs:108-151) + detectedSphereBoxCollision
// (CollisionDetection.fs:166-184).  Returns the contact with the normal "from box".
bool detect_sphere_box(const Body& sph, const Body& box, double eps, Contact* out, int* face_idx, bool* error)
{
    OBox ob = make_obox(box);
    double r = sph.dims[0];
    double best_pen = 0.0;
    int best = -1;
    for (int f = 0; f < 6; f++) {
        V3 axis = ob.f[f].n;
        double a = 0, b = 0;
        for (size_t k = 0; k < ob.verts.size(); k++) {
            double p = dot(axis, ob.verts[k]);
            if (k == 0) {
                a = p;
                b = p;
            } else {
                if (p < a) a = p;  // List.min
                if (p > b) b = p;  // List.max
            }
        }
        double c = dot(axis, sph.x);
        double cp = c + r, cm = c - r;
        double c1 = (cm < cp) ? cm : cp;  // List.min [c+r; c-r]
        double c2 = (cm > cp) ? cm : cp;  // List.max
        double pen;
        if (c1 >= a && c1 <= b)
            pen = b - c1;
        else if (c1 <= a && c2 >= a)
            pen = c2 - a;
        else
            return false;  // `sequence` of options: any None -> None
        if (best < 0 || pen < best_pen) {  // Seq.minBy: first minimum
            best = f;
            best_pen = pen;
        }
    }
    V3 nbox = ob.f[best].n;
    V3 startP = sph.x;
    V3 endP = sub(sph.x, smul(r, nbox));
    V3 pos;
    if (!plane_edge_intersection(eps, face_plane(ob.f[best]), startP, endP, &pos)) {
        *error = true;  // Option.get on None -> exception in the reference
        return false;
    }
    *out = Contact{nbox, pos, best_pen, 0u};
    *face_idx = best;
    return true;
}

struct SatAxis {
    int origin;  // 0 Face1, 1 Face2, 2 Edge
    int ei, ej;  // edge axes indices
    V3 n;        // NormalFromTarget
    double pen;
};

// SAT.tryGetCollisionDataForAxis (SAT.fs:155-183): T = target polyhedron, O = other
bool sat_axis(const OBox& T, const OBox& O, V3 axis, V3* n, double* pen)
{
    auto minmax = [&](const OBox& ob, double* mn, double* mx) {
        for (size_t k = 0; k < ob.verts.size(); k++) {
            double p = dot(ob.verts[k], axis);  // v |> dotProduct axis.Get = axis . v (products commute)
            if (k == 0) {
                *mn = p;
                *mx = p;
            } else {
                if (p < *mn) *mn = p;  // Seq.minBy snd
                if (p > *mx) *mx = p;  // Seq.maxBy snd
            }
        }
    };
    double a = 0, b = 0, c = 0, d = 0;
    minmax(T, &a, &b);
    minmax(O, &c, &d);
    if (a <= c && b >= c) {
        *n = axis;
        *pen = b - c;
        return true;
    }
    if (c <= a && d >= a) {
        *n = neg(axis);
        *pen = d - a;
        return true;
    }
    return false;
}

// Face.getEdges / OrientedBox.getEdges (Entities/Base.fs:30-31, OrientedBox.fs:28-29)
struct Edge {
    V3 a, b;
};
std::vector<Edge> obox_edges(const OBox& ob)
{
    std::vector<Edge> all;
    for (int f = 0; f < 6; f++) {
        const Face& F = ob.f[f];
        Edge e4[4] = {{F.v[3], F.v[0]}, {F.v[0], F.v[1]}, {F.v[1], F.v[2]}, {F.v[2], F.v[3]}};
        for (auto& e : e4) {
            bool seen = false;
            for (auto& q : all)
                if (veq(q.a, e.a) && veq(q.b, e.b)) {  // Pair equality is ordered (Pair.fs:17-21)
                    seen = true;
                    break;
                }
            if (!seen) all.push_back(e);
        }
    }
    return all;
}

// detectBoxBoxCollision (CollisionDetection.fs:29-164) over SAT.tryFindSeparatingAxis (SAT.fs:206-240)
bool detect_bb(const Body& b1, const Body& b2, double eps, std::vector<Contact>* out, bool* error)
{
    OBox ob1 = make_obox(b1), ob2 = make_obox(b2);
    V3 fromFirst = sub(b2.x, b1.x);
    V3 ax1[3], ax2[3];
    body_axes(b1, ax1);
    body_axes(b2, ax2);

    bool have = false;
    SatAxis best{};
    auto consider = [&](int origin, int ei, int ej, const OBox& T, const OBox& O, V3 axis) -> bool {
        V3 n;
        double pen;
        if (!sat_axis(T, O, axis, &n, &pen)) return false;  // any None -> whole result None
        if (!have || pen < best.pen) {                       // Seq.minBy: first minimum (SAT.fs:190)
            best = SatAxis{origin, ei, ej, n, pen};
            have = true;
        }
        return true;
    };
    for (int f = 0; f < 6; f++)
        if (!consider(0, 0, 0, ob1, ob2, ob1.f[f].n)) return false;  // Face1 (SAT.fs:230-231)
    for (int f = 0; f < 6; f++)
        if (!consider(1, 0, 0, ob2, ob1, ob2.f[f].n)) return false;  // Face2, boxes flipped (:233-234)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            V3 e = normalize(cross(ax2[j], ax1[i]));  // firstAxis |> crossProduct secondAxis (:223)
            if (is_zero(0.01, e)) continue;           // :225
            if (dot(e, fromFirst) < 0.0) e = neg(e);  // :226
            if (!consider(2, i, j, ob1, ob2, e)) return false;
        }
    if (!have) return false;

    if (best.origin == 0 || best.origin == 1) {
        // generateFaceContactPoints (CollisionDetection.fs:30-73)
        const OBox& ref = (best.origin == 0) ? ob1 : ob2;
        const OBox& oth = (best.origin == 0) ? ob2 : ob1;
        V3 normal = best.n;
        int rf = -1;
        for (int f = 0; f < 6; f++)
            if (veq(ref.f[f].n, normal)) {  // Box.findFaceByNormal (Box.fs:113-114)
                rf = f;
                break;
            }
        if (rf < 0) {
            *error = true;  // Seq.find throws
            return false;
        }
        const Face& refFace = ref.f[rf];
        int inc = 0;
        double incv = 0;
        for (int f = 0; f < 6; f++) {  // findIncidentFace (:35-37) first min
            double d = dot(refFace.n, oth.f[f].n);
            if (f == 0 || d < incv) {
                inc = f;
                incv = d;
            }
        }
        std::vector<Plane> planes;
        for (int f = 0; f < 6; f++) {  // Box.findAdjacentFaces (Box.fs:116-124)
            double d = std::fabs(dot(refFace.n, ref.f[f].n));
            if (!fs_equals(eps, 1.0, d)) planes.push_back(plane_inverted(face_plane(ref.f[f])));
        }
        std::vector<PV> poly;
        for (int k = 0; k < 4; k++) poly.push_back(PV{oth.f[inc].v[k], (uint8_t)k});
        bool clip_error = false;
        std::vector<PV> clipped = sh_clip(eps, planes, poly, &clip_error);
        Plane refInv = plane_inverted(face_plane(refFace));  // :52-54
        int ordinal = 0;
        for (auto& q : clipped) {
            if (!point_in_plane(refInv, q.p)) continue;
            V3 closest = closest_point_to_poly(q.p, refFace.v, 4);
            double pen = dot(normal, sub(q.p, closest));  // :42-45
            V3 n_out = (best.origin == 1) ? neg(normal) : normal;  // :157 WithInvertedNormals
            uint32_t feat = 3u | ((uint32_t)best.origin << 2) | ((uint32_t)rf << 4) | ((uint32_t)inc << 7) |
                            ((uint32_t)(ordinal & 31) << 10) | ((uint32_t)q.tag << 15);
            out->push_back(Contact{n_out, q.p, pen, feat});
            ordinal++;
        }
        return !out->empty();
    }
    // generateEdgeContactPoints (CollisionDetection.fs:75-132)
    V3 normal = best.n;
    auto supporting = [&](const OBox& ob, V3 edgeAxis, V3 nrm, int* idx) -> bool {
        std::vector<Edge> edges = obox_edges(ob);
        bool have_e = false;
        double bestv = 0;
        for (size_t k = 0; k < edges.size(); k++) {
            V3 d = sub(edges[k].b, edges[k].a);
            if (!are_parallel(eps, edgeAxis, d)) continue;  // d |> areParallel eps axis
            double p0 = dot(nrm, edges[k].a), p1 = dot(nrm, edges[k].b);
            double key = (p1 > p0) ? p1 : p0;  // Pair.max = List.max
            if (!have_e || key > bestv) {      // List.maxBy: first maximum
                have_e = true;
                bestv = key;
                *idx = (int)k;
            }
        }
        return have_e;
    };
    int e1i = 0, e2i = 0;
    if (!supporting(ob1, ax1[best.ei], normal, &e1i) || !supporting(ob2, ax2[best.ej], neg(normal), &e2i)) {
        *error = true;  // List.maxBy on empty list throws
        return false;
    }
    Edge E1 = obox_edges(ob1)[e1i], E2 = obox_edges(ob2)[e2i];
    // getContactPoint (:92-111)
    V3 DA = sub(E1.b, E1.a), DB = sub(E2.b, E2.a);
    V3 r = sub(E1.a, E2.a);
    double a = dot(DA, DA), e = dot(DB, DB), f = dot(DB, r), c = dot(DA, r), b = dot(DA, DB);
    double denom = a * e - b * b;
    double TA = (b * f - c * e) / denom;
    double TB = (b * TA + f) / e;
    V3 CA = add(E1.a, smul(TA, DA));
    V3 CB = add(E2.a, smul(TB, DB));
    V3 cp = add(smul(0.5, CA), smul(0.5, CB));
    uint32_t feat = 3u | (2u << 2) | ((uint32_t)best.ei << 4) | ((uint32_t)best.ej << 7) | ((uint32_t)e1i << 10) |
                    ((uint32_t)e2i << 15);
    out->push_back(Contact{normal, cp, -best.pen, feat});
    return true;
}

// CollisionDetection.areColliding (CollisionDetection.fs:208-220)
bool detect(const Body& b1, const Body& b2, double eps, std::vector<Contact>* out, bool* error)
{
    out->clear();
    if (b1.shape == 0 && b2.shape == 0) return detect_ss(b1, b2, out);
    if (b1.shape == 1 && b2.shape == 1) return detect_bb(b1, b2, eps, out, error);
    Contact c;
    int face = 0;
    if (b1.shape == 0) {  // Sphere, Box -> inverted normals (:217-219)
        if (!detect_sphere_box(b1, b2, eps, &c, &face, error)) return false;
        c.n = neg(c.n);
        c.feat = 1u | ((uint32_t)face << 4);
    } else {  // Box, Sphere (:220)
        if (!detect_sphere_box(b2, b1, eps, &c, &face, error)) return false;
        c.feat = 2u | ((uint32_t)face << 4);
    }
    out->push_back(c);
    return true;
}

// ------------------------------------------------------------------------------------
// Response: CollisionResponse.resolveCollision (CollisionResponse.fs:73-146)
// ------------------------------------------------------------------------------------
struct CPData {  // ContactPointImpulseData (ContactPointImpulseData.fs:9-21)
    double bias, accN, accT[2];
    V3 r1, r2;
    double mT[2], mN;
};

void resolve(Body& b1, Body& b2, const std::vector<Contact>& cs, const ps_step_config& cfg)
{
    V3 t[2];
    tangent_vectors(cs[0].n, &t[0], &t[1]);  // :80-81 tangents from the FIRST contact
    std::vector<CPData> cp(cs.size());
    for (size_t k = 0; k < cs.size(); k++) {
        // calculateBaumgarteBias (:35-41)
        double bias = fs_min(cfg.max_correction_velocity,
                             -cfg.baumgarte / cfg.dt * fs_min(0.0, cs[k].pen + cfg.allowed_penetration));
        // ContactPointImpulseData.init (ContactPointImpulseData.fs:23-52)
        CPData d;
        d.r1 = sub(cs[k].p, b1.x);
        d.r2 = sub(cs[k].p, b2.x);
        M3 K = constraint_mass(b1, b2, d.r1, d.r2);
        d.mN = dot(cs[k].n, mulMV(K, cs[k].n));
        for (int q = 0; q < 2; q++) d.mT[q] = dot(t[q], mulMV(K, t[q]));
        d.bias = bias;
        d.accN = 0.0;
        d.accT[0] = d.accT[1] = 0.0;
        cp[k] = d;
    }
    double inf = std::numeric_limits<double>::infinity();
    for (int it = 0; it < cfg.solver_iterations; it++) {       // applyN (:141-143)
        for (int k = (int)cs.size() - 1; k >= 0; k--) {        // List.foldBack (:118-131)
            CPData& d = cp[k];
            V3 n = cs[k].n;
            // calculateNormalImpulse (:45-71)
            V3 P;
            if (d.mN == 0.0) {
                P = v3(0.0, 0.0, 0.0);
            } else {
                double e = fs_min(b1.e, b2.e);
                V3 vrel = relative_velocity(b1, b2, d.r1, d.r2);
                double vn = dot(vrel, n);
                double j = (-(e + 1.0) * vn + d.bias) / d.mN;
                double old = d.accN;  // clampImpulseValue 0 None (ContactPointImpulseData.fs:56-66)
                d.accN = fs_clamp(0.0, inf, old + j);
                double dj = d.accN - old;
                P = smul(-dj, n);
            }
            apply_impulse(b1, d.r1, P);        // :98-102
            apply_impulse(b2, d.r2, neg(P));
            if (cfg.enable_friction) {
                // Friction.calculateImpulse (Friction.fs:39-68)
                double mu = fs_max(b2.mu_k, b1.mu_k);  // target |> compound other = max other target
                V3 vrel = relative_velocity(b1, b2, d.r1, d.r2);
                double maxF = mu * d.accN;
                V3 Pt[2];
                for (int q = 0; q < 2; q++) {
                    double jt = dot(t[q], vrel) / d.mT[q];
                    double old = d.accT[q];
                    d.accT[q] = fs_clamp(-maxF, maxF, old + jt);
                    Pt[q] = smul(d.accT[q] - old, t[q]);
                }
                apply_impulse(b1, d.r1, Pt[0]);  // applyImpulses: fold in list order (:110-113)
                apply_impulse(b1, d.r1, Pt[1]);
                apply_impulse(b2, d.r2, neg(Pt[0]));
                apply_impulse(b2, d.r2, neg(Pt[1]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Joints (Joints.fs:26-94, JointsRestorer.fs:9-26) — as written, off by default
// ------------------------------------------------------------------------------------
void joint_restore(const Body& e1, const Body& e2, const Joint& j, int iterations, Body* o1, Body* o2)
{
    V3 aw1 = mulMV(e1.R, j.la), aw2 = mulMV(e2.R, j.lb);            // JointImpulseData.init :26-40
    M3 K = constraint_mass(e1, e2, aw1, aw2);
    V3 vrel = relative_velocity(e1, e2, j.la, j.lb);                // :57-59 local offsets, entry bodies
    V3 impulse = mulMV(K, sub(v3(0.0, 0.0, 0.0), vrel));            // :61-62
    Body b1 = e1, b2 = e2;
    for (int it = 0; it < iterations; it++) {                       // applyN (:91)
        apply_impulse(b1, j.la, impulse);                           // :70-75
        apply_impulse(b2, j.lb, neg(impulse));
    }
    *o1 = b1;
    *o2 = b2;
}

// ------------------------------------------------------------------------------------
// AABB + SpatialTree broad phase
// ------------------------------------------------------------------------------------
struct Extent {
    double size, pos;
};
// SimulatorObject.getAABoundingBox (SimulatorObject.fs:58-71) + objExtentProvider (BroadPhase.fs:21-31)
// aabb_mode 1 = conservative sphere (side = 2r) — not a reference behaviour (Q5).
void body_extent(const Body& b, int aabb_mode, Extent out[3])
{
    double sx, sy, sz;
    if (b.shape == 0) {
        double side = (aabb_mode == 1) ? 2.0 * b.dims[0] : b.dims[0];  // Box.createCube radius (Q5)
        sx = sy = sz = side;
    } else {
        // Box.worldAxisAlignedSize (Box.fs:69-85)
        double hx = b.dims[0] / 2.0, hy = b.dims[1] / 2.0, hz = b.dims[2] / 2.0;
        V3 z = v3(0, 0, 0);
        V3 ex = to_world(b.R, z, v3(hx, 0.0, 0.0));
        V3 ey = to_world(b.R, z, v3(0.0, hy, 0.0));
        V3 ez = to_world(b.R, z, v3(0.0, 0.0, hz));
        double wx = std::fabs(ex.x) + std::fabs(ey.x) + std::fabs(ez.x);
        double wy = std::fabs(ex.y) + std::fabs(ey.y) + std::fabs(ez.y);
        double wz = std::fabs(ex.z) + std::fabs(ey.z) + std::fabs(ez.z);
        sx = wx * 2.0;
        sy = wy * 2.0;
        sz = wz * 2.0;
    }
    V3 mn = sub(b.x, sdiv(v3(sx, sy, sz), 2.0));  // BroadPhase.fs:24
    out[0] = Extent{sx, mn.x};
    out[1] = Extent{sy, mn.y};
    out[2] = Extent{sz, mn.z};
}

// SpatialTree.areIntersecting (SpatialTree.fs:39-41) with withPositionCentered (:26-28)
bool extents_intersect(const std::vector<Extent>& A, const std::vector<Extent>& B)
{
    for (size_t k = 0; k < A.size(); k++) {
        double ca = A[k].pos + A[k].size / 2.0, cb = B[k].pos + B[k].size / 2.0;
        if (!((A[k].size + B[k].size) / 2.0 > std::fabs(cb - ca))) return false;
    }
    return true;
}

// n-D 2^n tree (Utilities/SpatialTree.fs)
struct TNode {
    bool leaf = true;
    std::set<int> objs;
    std::vector<TNode> kids;
};
struct Tree {
    TNode root;
    int max_leaf = 4, max_depth = 10;
    std::vector<std::pair<double, double>> bounds;
};
typedef std::vector<Extent> (*ExtentFn)(void* ctx, int id);

std::vector<Extent> tree_space(const Tree& t)  // SpaceExtent.init (:33-36)
{
    std::vector<Extent> s;
    for (auto& b : t.bounds) s.push_back(Extent{b.second - b.first, b.first});
    return s;
}
// getSubspaceByIndex / getNodePositionByIndex (:69-86): LAST dimension takes bit 0
std::vector<Extent> subspace(const std::vector<Extent>& parent, int index)
{
    size_t n = parent.size();
    std::vector<double> off(n);
    int rest = index;
    for (int k = (int)n - 1; k >= 0; k--) {
        off[k] = (rest % 2 == 0) ? 0.0 : (parent[k].size / 2.0);
        rest /= 2;
    }
    std::vector<Extent> s(n);
    for (size_t k = 0; k < n; k++) s[k] = Extent{parent[k].size / 2.0, parent[k].pos + off[k]};
    return s;
}
// insert (:166-208). Returns false where the reference throws ArgumentException (out of bounds).
bool tree_insert(Tree& t, ExtentFn fn, void* ctx, int id)
{
    std::vector<Extent> oe = fn(ctx, id);
    if (oe.size() != t.bounds.size()) return false;
    if (!extents_intersect(oe, tree_space(t))) return false;  // :171-172 (areIntersecting object space)
    struct Rec {
        Tree& t;
        ExtentFn fn;
        void* ctx;
        int id;
        const std::vector<Extent>& oe;
        void go(int depth, const std::vector<Extent>& ext, TNode& node)
        {
            if (node.leaf) {
                if ((int)node.objs.size() < t.max_leaf || depth >= t.max_depth) {
                    node.objs.insert(id);
                } else {
                    std::set<int> all = node.objs;
                    all.insert(id);
                    int nk = 1 << t.bounds.size();
                    node.kids.assign(nk, TNode{});
                    for (int s = 0; s < nk; s++) {
                        std::vector<Extent> sub = subspace(ext, s);
                        for (int o : all)
                            if (extents_intersect(fn(ctx, o), sub)) node.kids[s].objs.insert(o);  // :179-182
                    }
                    node.leaf = false;
                    node.objs.clear();
                }
            } else {
                for (size_t i = 0; i < node.kids.size(); i++) {
                    std::vector<Extent> sub = subspace(ext, (int)i);
                    if (extents_intersect(sub, oe)) go(depth + 1, sub, node.kids[i]);  // :194-203
                }
            }
        }
    } rec{t, fn, ctx, id, oe};
    rec.go(1, tree_space(t), t.root);
    return true;
}
void tree_remove_all(TNode& n, const std::set<int>& ids)  // removeAll (:122-135)
{
    if (n.leaf) {
        for (int i : ids) n.objs.erase(i);
    } else
        for (auto& k : n.kids) tree_remove_all(k, ids);
}
void tree_buckets(const TNode& n, std::vector<std::set<int>>* out)  // getObjectBuckets (:217-224)
{
    if (n.leaf) {
        if (!n.objs.empty()) out->push_back(n.objs);
    } else
        for (auto& k : n.kids) tree_buckets(k, out);
}
// serialise tree shape for the known-answer tests: "L{a,b}" / "N[...]"
void tree_dump(const TNode& n, std::string* s)
{
    if (n.leaf) {
        *s += "L{";
        bool first = true;
        for (int o : n.objs) {
            if (!first) *s += ",";
            *s += std::to_string(o);
            first = false;
        }
        *s += "}";
    } else {
        *s += "N[";
        for (size_t i = 0; i < n.kids.size(); i++) {
            if (i) *s += " ";
            tree_dump(n.kids[i], s);
        }
        *s += "]";
    }
}

// ------------------------------------------------------------------------------------
// World + step orchestration
// ------------------------------------------------------------------------------------
struct World {
    std::vector<Body> bodies;
    std::vector<Joint> joints;
    ps_step_config cfg{};
    // broad phase: 0 Dummy, 1 SpatialTree
    int bp_kind = 0;
    Tree tree;
    int aabb_mode = 0;
    // results of the last step
    std::vector<std::pair<int, int>> candidates;
    std::map<std::pair<int, int>, std::vector<Contact>> collisions;  // SimulatorState.Collisions
    int64_t n_tested = 0, n_hits = 0, n_contacts = 0, n_errors = 0;
    double max_pen = 0.0;
};

std::vector<Extent> world_extent_fn(void* ctx, int id)
{
    World* w = (World*)ctx;
    Extent e[3];
    body_extent(w->bodies[id], w->aabb_mode, e);
    return {e[0], e[1], e[2]};
}

// BroadPhase.init (BroadPhase.fs:72-81, :41-60): insert ALL ids in id order; throws when out of bounds
bool world_init_tree(World& w, int leaf_cap, int max_depth, const double mn[3], const double mx[3])
{
    w.bp_kind = 1;
    w.tree = Tree{};
    w.tree.max_leaf = leaf_cap;
    w.tree.max_depth = max_depth;
    w.tree.bounds = {{mn[0], mx[0]}, {mn[1], mx[1]}, {mn[2], mx[2]}};
    for (int i = 0; i < (int)w.bodies.size(); i++)
        if (!tree_insert(w.tree, world_extent_fn, &w, i)) return false;
    return true;
}

// BroadPhase.update (BroadPhase.fs:100-127) + getCollisionCandidates (:62-70) + dummy_init (:33-39)
std::vector<std::pair<int, int>> broadphase_update(World& w)
{
    std::vector<std::pair<int, int>> cand;
    int n = (int)w.bodies.size();
    if (w.bp_kind == 0) {
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) cand.push_back({i, j});
        return cand;
    }
    std::set<int> dyn;
    for (int i = 0; i < n; i++)
        if (!w.bodies[i].is_static()) dyn.insert(i);
    tree_remove_all(w.tree.root, dyn);
    for (int i : dyn) tree_insert(w.tree, world_extent_fn, &w, i);  // tryInsert: silently dropped (:117-120)
    std::vector<std::set<int>> buckets;
    tree_buckets(w.tree.root, &buckets);
    std::vector<std::set<int>> distinct;  // Seq.filter count>1 |> Seq.distinct
    for (auto& b : buckets) {
        if (b.size() <= 1) continue;
        if (std::find(distinct.begin(), distinct.end(), b) == distinct.end()) distinct.push_back(b);
    }
    std::set<std::pair<int, int>> seen;  // Seq.distinctBy unorderedKey
    for (auto& b : distinct) {
        std::vector<int> v(b.begin(), b.end());
        for (size_t i = 0; i < v.size(); i++)  // subSetsOf2Tail |> Set.toSeq: lexicographic
            for (size_t j = i + 1; j < v.size(); j++) {
                std::pair<int, int> p{v[i], v[j]};
                if (seen.insert(p).second) cand.push_back(p);
            }
    }
    return cand;
}

bool body_has_nan(const Body& b)
{
    const double* p[] = {&b.x.x, &b.v.x, &b.L.x};
    for (auto q : p)
        for (int k = 0; k < 3; k++)
            if (!std::isfinite(q[k])) return true;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            if (!std::isfinite(b.R.m[i][j])) return true;
    return false;
}

// Simulator.updateSimulation (Simulator.fs:52-56): resolveAll (CollisionResolver.fs:49-69) then
// [JointsRestorer.restoreAll] then SimulatorState.update (SimulatorState.fs:169-173)
void world_step(World& w)
{
    const ps_step_config& cfg = w.cfg;
    w.collisions.clear();                                    // CollisionResolver.fs:50-52
    std::vector<std::pair<int, int>> all = broadphase_update(w);  // :62-64 (step-start states)
    w.candidates.clear();
    for (auto& p : all)                                      // canIntersect (:58-60,67)
        if (!w.bodies[p.first].is_static() || !w.bodies[p.second].is_static()) w.candidates.push_back(p);
    std::vector<Contact> cs;
    for (auto& p : w.candidates) {                           // List.fold withCollisionResponse (:68)
        Body Pi = integrate(w.bodies[p.first], cfg.dt, cfg.integrator_kind);   // :42-43
        Body Pj = integrate(w.bodies[p.second], cfg.dt, cfg.integrator_kind);
        bool err = false;
        w.n_tested++;
        if (!detect(Pi, Pj, cfg.epsilon, &cs, &err)) {       // :16
            if (err) w.n_errors++;
            continue;                                        // :46 prediction discarded
        }
        w.collisions[p] = cs;                                // :33-34 (pre-resolution contacts, Q18)
        w.n_hits++;
        w.n_contacts += (int64_t)cs.size();
        for (auto& c : cs) w.max_pen = std::max(w.max_pen, std::fabs(c.pen));
        resolve(Pi, Pj, cs, cfg);                            // :21-25
        w.bodies[p.first] = Pi;                              // :27-31 write back predicted+resolved
        w.bodies[p.second] = Pj;
    }
    if (cfg.restore_joints && !w.joints.empty()) {           // Simulator.fs:55 (commented out there)
        std::vector<Body> entry = w.bodies;                  // every joint reads simState (JointsRestorer.fs:12-17)
        for (auto& j : w.joints) {
            Body o1, o2;
            joint_restore(entry[j.a], entry[j.b], j, cfg.solver_iterations, &o1, &o2);
            // changePhysicalObjects zips id-SORTED objects with pair-ordered results (Q15)
            int lo = std::min(j.a, j.b), hi = std::max(j.a, j.b);
            auto inject = [&](int id, const Body& src) {  // withPhysicalObject: physical state only
                Body& dst = w.bodies[id];
                int shape = dst.shape;
                double d0 = dst.dims[0], d1 = dst.dims[1], d2 = dst.dims[2];
                V3 F = dst.force, T = dst.torque;  // forces are keyed by id, collider stays with the object
                dst = src;
                dst.shape = shape;
                dst.dims[0] = d0; dst.dims[1] = d1; dst.dims[2] = d2;
                dst.force = F; dst.torque = T;
            };
            if (lo == hi) {
                inject(lo, o1);
            } else {
                inject(lo, o1);
                inject(hi, o2);
            }
        }
    }
    for (auto& b : w.bodies) b = integrate(b, cfg.dt, cfg.integrator_kind);  // SimulatorState.fs:169-173
}

// deep copy of a Body from flat arrays
void fill_body(Body& b, int i, const int32_t* shape, const double* dims, const double* mass, const double* iinv,
               const double* pos, const double* vel, const double* rot, const double* L, const double* force,
               const double* torque, const double* e, const double* muk, const double* mus)
{
    b.shape = shape[i];
    for (int k = 0; k < 3; k++) {
        b.dims[k] = dims[3 * i + k];
        b.iinv[k] = iinv[3 * i + k];
    }
    b.mass = mass[i];
    b.x = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    b.v = v3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) b.R.m[r][c] = rot[9 * i + 3 * r + c];
    b.L = v3(L[3 * i], L[3 * i + 1], L[3 * i + 2]);
    b.force = v3(force[3 * i], force[3 * i + 1], force[3 * i + 2]);
    b.torque = v3(torque[3 * i], torque[3 * i + 1], torque[3 * i + 2]);
    b.e = e[i];
    b.mu_k = muk[i];
    b.mu_s = mus ? mus[i] : 0.0;
}

}  // namespace

// ====================================================================================
// C API for ctypes (tests/, smoke(), bench.py cpu_baseline only)
// ====================================================================================
extern "C" {

void* orc_world_create(int n, const int32_t* shape, const double* dims, const double* mass, const double* iinv,
                       const double* pos, const double* vel, const double* rot, const double* L,
                       const double* force, const double* torque, const double* e, const double* muk,
                       const double* mus, const ps_step_config* cfg)
{
    World* w = new World();
    w->bodies.resize(n);
    for (int i = 0; i < n; i++) fill_body(w->bodies[i], i, shape, dims, mass, iinv, pos, vel, rot, L, force, torque, e, muk, mus);
    w->cfg = *cfg;
    return w;
}
void orc_world_destroy(void* h) { delete (World*)h; }

void orc_world_set_joints(void* h, int nj, const int32_t* a, const int32_t* b, const double* la, const double* lb)
{
    World* w = (World*)h;
    w->joints.clear();
    for (int k = 0; k < nj; k++)
        w->joints.push_back(Joint{a[k], b[k], v3(la[3 * k], la[3 * k + 1], la[3 * k + 2]), v3(lb[3 * k], lb[3 * k + 1], lb[3 * k + 2])});
}

// returns 0, or -1 where the reference throws (an object outside the bounds at init)
int orc_world_use_spatial_tree(void* h, int leaf_cap, int max_depth, const double* mn, const double* mx, int aabb_mode)
{
    World* w = (World*)h;
    w->aabb_mode = aabb_mode;
    return world_init_tree(*w, leaf_cap, max_depth, mn, mx) ? 0 : -1;
}

void orc_world_step(void* h, int nsteps)
{
    World* w = (World*)h;
    for (int s = 0; s < nsteps; s++) world_step(*w);
}

void orc_world_get_state(void* h, double* pos, double* vel, double* rot, double* L)
{
    World* w = (World*)h;
    for (size_t i = 0; i < w->bodies.size(); i++) {
        const Body& b = w->bodies[i];
        pos[3 * i] = b.x.x; pos[3 * i + 1] = b.x.y; pos[3 * i + 2] = b.x.z;
        vel[3 * i] = b.v.x; vel[3 * i + 1] = b.v.y; vel[3 * i + 2] = b.v.z;
        L[3 * i] = b.L.x; L[3 * i + 1] = b.L.y; L[3 * i + 2] = b.L.z;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) rot[9 * i + 3 * r + c] = b.R.m[r][c];
    }
}
void orc_world_set_state(void* h, const double* pos, const double* vel, const double* rot, const double* L)
{
    World* w = (World*)h;
    for (size_t i = 0; i < w->bodies.size(); i++) {
        Body& b = w->bodies[i];
        b.x = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        b.v = v3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        b.L = v3(L[3 * i], L[3 * i + 1], L[3 * i + 2]);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) b.R.m[r][c] = rot[9 * i + 3 * r + c];
    }
}
int orc_world_num_bodies(void* h) { return (int)((World*)h)->bodies.size(); }

// SimulatorState.applyImpulse (SimulatorState.fs:175-181)
void orc_world_apply_impulse(void* h, int body, const double* J, const double* off)
{
    World* w = (World*)h;
    apply_impulse(w->bodies[body], v3(off[0], off[1], off[2]), v3(J[0], J[1], J[2]));
}
// SimulatorStateBuilder.withPrototype (SimulatorState.fs:89-107): id = max id + 1; returns the id, or -1
// where the reference throws (octree insert out of bounds, BroadPhase.fs:83-98)
int orc_world_add_body(void* h, int shape, const double* dims, double mass, const double* iinv, const double* pos,
                       const double* vel, const double* rot, const double* L, const double* force,
                       const double* torque, double e, double muk, double mus)
{
    World* w = (World*)h;
    Body b;
    int32_t sh = shape;
    fill_body(b, 0, &sh, dims, &mass, iinv, pos, vel, rot, L, force, torque, &e, &muk, &mus);
    w->bodies.push_back(b);
    int id = (int)w->bodies.size() - 1;
    if (w->bp_kind == 1 && !tree_insert(w->tree, world_extent_fn, w, id)) {
        w->bodies.pop_back();
        return -1;
    }
    return id;
}

int orc_world_get_candidates(void* h, int32_t* ids, int cap)
{
    World* w = (World*)h;
    int n = (int)w->candidates.size();
    for (int k = 0; k < n && k < cap; k++) {
        ids[2 * k] = w->candidates[k].first;
        ids[2 * k + 1] = w->candidates[k].second;
    }
    return n;
}

// Collisions map of the last step, ascending (i,j) (Map<Set<Id>,_> order)
int orc_world_get_contacts(void* h, int32_t* n_pairs, int32_t* ids, int32_t* first, double* normal, double* position,
                           double* pen, uint32_t* feat, int pair_cap, int contact_cap)
{
    World* w = (World*)h;
    int np = 0, nc = 0;
    for (auto& kv : w->collisions) {
        if (np >= pair_cap) return -1;
        ids[2 * np] = kv.first.first;
        ids[2 * np + 1] = kv.first.second;
        first[np] = nc;
        for (auto& c : kv.second) {
            if (nc >= contact_cap) return -1;
            normal[3 * nc] = c.n.x; normal[3 * nc + 1] = c.n.y; normal[3 * nc + 2] = c.n.z;
            position[3 * nc] = c.p.x; position[3 * nc + 1] = c.p.y; position[3 * nc + 2] = c.p.z;
            pen[nc] = c.pen;
            feat[nc] = c.feat;
            nc++;
        }
        np++;
    }
    first[np] = nc;
    *n_pairs = np;
    return 0;
}

void orc_world_stats(void* h, int64_t* tested, int64_t* hits, int64_t* contacts, int64_t* errors, double* max_pen,
                     int32_t* nan_flag)
{
    World* w = (World*)h;
    *tested = w->n_tested;
    *hits = w->n_hits;
    *contacts = w->n_contacts;
    *errors = w->n_errors;
    *max_pen = w->max_pen;
    int nan = 0;
    for (auto& b : w->bodies)
        if (body_has_nan(b)) nan = 1;
    *nan_flag = nan;
}

// step many independent worlds on `nthreads` host threads (CPU baseline)
void orc_batch_step(void** hs, int n_worlds, int nsteps, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=]() {
            for (int wi = t; wi < n_worlds; wi += nthreads) {
                World* w = (World*)hs[wi];
                for (int s = 0; s < nsteps; s++) world_step(*w);
            }
        });
    for (auto& t : th) t.join();
}

// ---- unit-level entry points (tests T2-T6) ------------------------------------------
// one integrate() on every body of the world, no collisions (T1)
void orc_world_integrate_only(void* h, int nsteps)
{
    World* w = (World*)h;
    for (int s = 0; s < nsteps; s++)
        for (auto& b : w->bodies) b = integrate(b, w->cfg.dt, w->cfg.integrator_kind);
}
// narrow phase on the CURRENT (not predicted) poses of bodies i,j: returns #contacts or -1 on a reference exception
int orc_world_detect(void* h, int i, int j, double* normal, double* position, double* pen, uint32_t* feat, int cap)
{
    World* w = (World*)h;
    std::vector<Contact> cs;
    bool err = false;
    bool hit = detect(w->bodies[i], w->bodies[j], w->cfg.epsilon, &cs, &err);
    if (err) return -1;
    if (!hit) return 0;
    int n = (int)cs.size();
    for (int k = 0; k < n && k < cap; k++) {
        normal[3 * k] = cs[k].n.x; normal[3 * k + 1] = cs[k].n.y; normal[3 * k + 2] = cs[k].n.z;
        position[3 * k] = cs[k].p.x; position[3 * k + 1] = cs[k].p.y; position[3 * k + 2] = cs[k].p.z;
        pen[k] = cs[k].pen;
        feat[k] = cs[k].feat;
    }
    return n;
}
// Sutherland–Hodgman clip of a polygon by planes (n.p - d >= 0 kept): returns #vertices
int orc_clip(double eps, int n_planes, const double* plane_n, const double* plane_d, int n_poly, const double* poly,
             double* out, int cap)
{
    std::vector<Plane> pls;
    for (int k = 0; k < n_planes; k++) pls.push_back(Plane{v3(plane_n[3 * k], plane_n[3 * k + 1], plane_n[3 * k + 2]), plane_d[k]});
    std::vector<PV> p;
    for (int k = 0; k < n_poly; k++) p.push_back(PV{v3(poly[3 * k], poly[3 * k + 1], poly[3 * k + 2]), (uint8_t)k});
    bool ce = false;
    std::vector<PV> r = sh_clip(eps, pls, p, &ce);
    for (int k = 0; k < (int)r.size() && k < cap; k++) {
        out[3 * k] = r[k].p.x; out[3 * k + 1] = r[k].p.y; out[3 * k + 2] = r[k].p.z;
    }
    return (int)r.size();
}

// ---- octree known-answer interface (PhysicsSimulator.Tests/UtilitiesTests.fs) ---------
struct OrcTree {
    Tree t;
    std::map<int, std::vector<Extent>> ext;
};
static std::vector<Extent> orc_tree_extent(void* ctx, int id)
{
    OrcTree* ot = (OrcTree*)ctx;
    auto it = ot->ext.find(id);
    if (it == ot->ext.end()) return {};
    return it->second;
}
void* orc_tree_create(int max_leaf, int max_depth, int ndim, const double* bounds /*2*ndim*/)
{
    if (ndim < 1 || max_leaf < 1 || max_depth < 1) return nullptr;  // SpatialTree.init throws (:93-101)
    OrcTree* ot = new OrcTree();
    ot->t.max_leaf = max_leaf;
    ot->t.max_depth = max_depth;
    for (int k = 0; k < ndim; k++) ot->t.bounds.push_back({bounds[2 * k], bounds[2 * k + 1]});
    return ot;
}
void orc_tree_destroy(void* h) { delete (OrcTree*)h; }
void orc_tree_set_extent(void* h, int id, const double* size_pos /*2*ndim: size,pos per dim*/)
{
    OrcTree* ot = (OrcTree*)h;
    std::vector<Extent> e;
    for (size_t k = 0; k < ot->t.bounds.size(); k++) e.push_back(Extent{size_pos[2 * k], size_pos[2 * k + 1]});
    ot->ext[id] = e;
}
// replace the root by a leaf holding `ids` (tests build trees by hand: singleNodeTreeFromPosMap)
void orc_tree_set_root_leaf(void* h, int n, const int32_t* ids)
{
    OrcTree* ot = (OrcTree*)h;
    ot->t.root = TNode{};
    for (int k = 0; k < n; k++) ot->t.root.objs.insert(ids[k]);
}
// root := NonLeaf of leaves; `counts[k]` ids per child, concatenated in `ids`
void orc_tree_set_root_children(void* h, int n_children, const int32_t* counts, const int32_t* ids)
{
    OrcTree* ot = (OrcTree*)h;
    ot->t.root = TNode{};
    ot->t.root.leaf = false;
    ot->t.root.kids.assign(n_children, TNode{});
    int p = 0;
    for (int c = 0; c < n_children; c++)
        for (int k = 0; k < counts[c]; k++) ot->t.root.kids[c].objs.insert(ids[p++]);
}
int orc_tree_insert(void* h, int id)
{
    OrcTree* ot = (OrcTree*)h;
    return tree_insert(ot->t, orc_tree_extent, ot, id) ? 0 : -1;
}
void orc_tree_remove(void* h, int id)
{
    OrcTree* ot = (OrcTree*)h;
    tree_remove_all(ot->t.root, std::set<int>{id});
}
int orc_tree_dump(void* h, char* buf, int cap)
{
    OrcTree* ot = (OrcTree*)h;
    std::string s;
    tree_dump(ot->t.root, &s);
    if ((int)s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
int orc_tree_num_buckets(void* h)
{
    OrcTree* ot = (OrcTree*)h;
    std::vector<std::set<int>> b;
    tree_buckets(ot->t.root, &b);
    return (int)b.size();
}

// ---- OVERLAP(aabb_mode) pair set for config 5 (SURVEY.md §8(a')) ----------------------
// sorted (i<j), >= 1 dynamic, AABBs overlap strictly on all axes (SpatialTree.fs:39-41 predicate
// on SimulatorObject.getAABoundingBox).  method 0: all pairs O(n^2); 1: sort-and-sweep on x.
int64_t orc_overlap_pairs(int n, const int32_t* shape, const double* dims, const double* mass, const double* pos,
                          const double* rot, int aabb_mode, int method, int32_t* out_ids, int64_t cap)
{
    std::vector<std::vector<Extent>> ex(n);
    std::vector<char> stat(n);
    for (int i = 0; i < n; i++) {
        Body b{};
        b.shape = shape[i];
        for (int k = 0; k < 3; k++) b.dims[k] = dims[3 * i + k];
        b.mass = mass[i];
        b.x = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) b.R.m[r][c] = rot[9 * i + 3 * r + c];
        Extent e[3];
        body_extent(b, aabb_mode, e);
        ex[i] = {e[0], e[1], e[2]};
        stat[i] = b.is_static();
    }
    std::vector<std::pair<int, int>> pairs;
    if (method == 0) {
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++)
                if (!(stat[i] && stat[j]) && extents_intersect(ex[i], ex[j])) pairs.push_back({i, j});
    } else {
        std::vector<int> order(n);
        for (int i = 0; i < n; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return ex[a][0].pos < ex[b][0].pos; });
        for (int a = 0; a < n; a++) {
            int i = order[a];
            double hi = ex[i][0].pos + ex[i][0].size;
            for (int bq = a + 1; bq < n; bq++) {
                int j = order[bq];
                if (ex[j][0].pos > hi + 1e-9 * (1.0 + std::fabs(hi))) break;  // sweep bound is conservative; exact test below
                if (stat[i] && stat[j]) continue;
                int lo = std::min(i, j), hh = std::max(i, j);
                if (extents_intersect(ex[lo], ex[hh])) pairs.push_back({lo, hh});
            }
        }
        std::sort(pairs.begin(), pairs.end());
    }
    int64_t np = (int64_t)pairs.size();
    for (int64_t k = 0; k < np && k < cap; k++) {
        out_ids[2 * k] = pairs[k].first;
        out_ids[2 * k + 1] = pairs[k].second;
    }
    return np;
}

// AABB (size, min corner) of every body — lets tests check the device AABB kernel bit for bit
void orc_aabbs(int n, const int32_t* shape, const double* dims, const double* pos, const double* rot, int aabb_mode,
               double* size_out, double* min_out)
{
    for (int i = 0; i < n; i++) {
        Body b{};
        b.shape = shape[i];
        for (int k = 0; k < 3; k++) b.dims[k] = dims[3 * i + k];
        b.x = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) b.R.m[r][c] = rot[9 * i + 3 * r + c];
        Extent e[3];
        body_extent(b, aabb_mode, e);
        for (int k = 0; k < 3; k++) {
            size_out[3 * i + k] = e[k].size;
            min_out[3 * i + k] = e[k].pos;
        }
    }
}

}  // extern "C"
