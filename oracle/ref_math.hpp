// oracle/ref_math.hpp — TEST INFRASTRUCTURE ONLY (parity oracle, CPU).
//
// FP64 3-vector / 3x3-matrix arithmetic in the evaluation order the reference
// gets from MathNet.Numerics 5.0.0 (managed provider) and FSharp.Core.
// MathNet is a NuGet dependency (PhysicsSimulatorCore.fsproj:11-12) and is NOT
// vendored under /root/reference, so these semantics restate its published
// algorithms.  PARITY UNPINNED at this boundary: the
// reference has no test that fixes a numeric result of these operations.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
// anything under oracle/.  The product (physicssimulator_b200/) never does.
//
// Citations are relative to /root/reference/PhysicsSimulatorCore/.
#pragma once
#include <cmath>
#include <cstdint>

#include "../physicssimulator_b200/csrc/ps_sincos.h"  // the one shared polynomial (see its header)

namespace ref {

struct V3 {
    double x, y, z;
};
struct M3 {
    double m[3][3];  // row-major: m[row][col]
};

// System.Math.Min / Max (F# `min`/`max` on float dispatch to these): NaN
// propagates, -0.0 < +0.0.
inline double fs_min(double a, double b)
{
    if (a != b && !std::isnan(a)) return a < b ? a : b;
    return std::signbit(a) ? a : b;
}
inline double fs_max(double a, double b)
{
    if (a != b) {
        if (!std::isnan(a)) return b < a ? a : b;
        return a;
    }
    return std::signbit(b) ? a : b;
}
// Utilities/Utils.fs:48  clamp low high value = high |> min value |> max low
inline double fs_clamp(double low, double high, double value) { return fs_max(low, fs_min(value, high)); }
// Utilities/Utils.fs:46  equals epsilon a b = (b - a |> abs) <= epsilon
inline bool fs_equals(double eps, double a, double b) { return std::fabs(b - a) <= eps; }

inline V3 v3(double x, double y, double z) { return V3{x, y, z}; }
inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }   // Vector3D.fs:29
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }   // Vector3D.fs:30
inline V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }                        // Vector3D.fs:35

// scalar * Vector (Vector3D.fs:32-34). MathNet Vector<T>.Multiply(scalar):
// scalar == 1 -> clone, scalar == 0 -> zero vector, else element-wise.
inline V3 smul(double s, V3 v)
{
    if (s == 1.0) return v;
    if (s == 0.0) return V3{0.0, 0.0, 0.0};
    return V3{s * v.x, s * v.y, s * v.z};
}
// Vector3D `/` (Vector3D.fs:34): v * (1.0 / scalar)
inline V3 sdiv(V3 v, double s) { return smul(1.0 / s, v); }

// DotProduct (Vector3D.fs:31,75): sum = 0; sum += x[i]*y[i]
inline double dot(V3 a, V3 b)
{
    double s = 0.0;
    s += a.x * b.x;
    s += a.y * b.y;
    s += a.z * b.z;
    return s;
}
// Vector3D.fs:60-65
inline V3 cross(V3 a, V3 b)
{
    return V3{(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
// MathNet SpecialFunctions.Hypotenuse
inline double hyp(double a, double b)
{
    if (std::fabs(a) > std::fabs(b)) {
        double r = b / a;
        return std::fabs(a) * std::sqrt(1.0 + (r * r));
    }
    if (b != 0.0) {
        double r = a / b;
        return std::fabs(b) * std::sqrt(1.0 + (r * r));
    }
    return 0.0;
}
// L2Norm (Vector3D.fs:67): values.Aggregate(0, Hypotenuse)
inline double norm(V3 v) { return hyp(hyp(hyp(0.0, v.x), v.y), v.z); }
// Normalize(2.0) (Vector3D.fs:69-70): norm == 0 -> copy, else v * (1/norm)
inline V3 normalize(V3 v)
{
    double n = norm(v);
    if (n == 0.0) return v;
    return smul(1.0 / n, v);
}
inline bool is_zero(double eps, V3 v) { return fs_equals(eps, 0.0, norm(v)); }          // Vector3D.fs:76
inline bool are_parallel(double eps, V3 v1, V3 v2) { return is_zero(eps, cross(v2, v1)); }  // Vector3D.fs:77: v1 |> crossProduct v2
inline bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }  // structural equality (NaN never reaches it here)

// dense M * v (Matrix3.fs:17): out[i] = sum_k m[i][k]*v[k], ascending k from 0.0
inline V3 mulMV(const M3& a, V3 v)
{
    double o[3];
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
        s += a.m[i][0] * v.x;
        s += a.m[i][1] * v.y;
        s += a.m[i][2] * v.z;
        o[i] = s;
    }
    return V3{o[0], o[1], o[2]};
}
// dense M * N (Matrix3.fs:15)
inline M3 mulMM(const M3& a, const M3& b)
{
    M3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
            o.m[i][j] = s;
        }
    return o;
}
// dense M * DiagonalMatrix(d): element-wise m[i][j]*d[j] (MathNet special case)
inline M3 mulMDiag(const M3& a, const double d[3])
{
    M3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.m[i][j] = a.m[i][j] * d[j];
    return o;
}
inline M3 transpose(const M3& a)
{
    M3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.m[i][j] = a.m[j][i];
    return o;
}
inline M3 addMM(const M3& a, const M3& b)
{
    M3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.m[i][j] = a.m[i][j] + b.m[i][j];
    return o;
}
inline M3 zeroM()
{
    M3 o;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.m[i][j] = 0.0;
    return o;
}
// Matrix3.fs:32-37
inline M3 hat(V3 v)
{
    M3 o;
    o.m[0][0] = 0.0;  o.m[0][1] = -v.z; o.m[0][2] = v.y;
    o.m[1][0] = v.z;  o.m[1][1] = 0.0;  o.m[1][2] = -v.x;
    o.m[2][0] = -v.y; o.m[2][1] = v.x;  o.m[2][2] = 0.0;
    return o;
}
inline V3 col(const M3& a, int c) { return V3{a.m[0][c], a.m[1][c], a.m[2][c]}; }

// Matrix3.orthonormalize (Matrix3.fs:41-61)
inline M3 orthonormalize(const M3& a)
{
    auto deltaCol = [](V3 c1, V3 c2) { return smul(dot(c1, c2) / dot(c1, c1), c1); };  // :44-45
    V3 c0 = normalize(col(a, 0));
    V3 c1 = col(a, 1);
    V3 c2 = col(a, 2);
    c1 = sub(c1, deltaCol(c0, c1));
    c1 = sub(c1, deltaCol(c0, c1));
    c1 = normalize(c1);
    c2 = sub(c2, deltaCol(c0, c2));
    c2 = sub(c2, deltaCol(c1, c2));
    c2 = sub(c2, deltaCol(c0, c2));
    c2 = sub(c2, deltaCol(c1, c2));
    c2 = normalize(c2);
    M3 o;
    o.m[0][0] = c0.x; o.m[1][0] = c0.y; o.m[2][0] = c0.z;
    o.m[0][1] = c1.x; o.m[1][1] = c1.y; o.m[2][1] = c1.z;
    o.m[0][2] = c2.x; o.m[1][2] = c2.y; o.m[2][2] = c2.z;
    return o;
}

// RotationMatrix3D.fromAxisAndAngle (Matrix3.fs:67-88)
inline M3 from_axis_angle(V3 u, double angle)
{
    double sinA, cosA;
    ps_sincos(angle, &sinA, &cosA);
    double cosInv = 1.0 - cosA;
    double ux = u.x, uy = u.y, uz = u.z;
    M3 o;
    o.m[0][0] = cosA + ux * ux * cosInv;
    o.m[0][1] = ux * uy * cosInv - uz * sinA;
    o.m[0][2] = ux * uz * cosInv + uy * sinA;
    o.m[1][0] = uy * ux * cosInv + uz * sinA;
    o.m[1][1] = cosA + uy * uy * cosInv;
    o.m[1][2] = uy * uz * cosInv - ux * sinA;
    o.m[2][0] = uz * ux * cosInv - uy * sinA;
    o.m[2][1] = uz * uy * cosInv + ux * sinA;
    o.m[2][2] = cosA + uz * uz * cosInv;
    return o;
}

}  // namespace ref
