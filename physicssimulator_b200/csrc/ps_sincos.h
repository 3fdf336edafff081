/* ps_sincos.h — one sin/cos for host and device.
 *
 * The reference takes sin/cos from System.Math (platform CRT), which is not
 * bit-reproducible across platforms.  To make the CUDA
 * path and the CPU oracle agree bit for bit, both evaluate THIS polynomial
 * (fdlibm-style: Cody–Waite reduction by pi/2 in three pieces + minimax
 * kernels on [-pi/4, pi/4]), compiled without FMA contraction on both sides
 * (nvcc -fmad=false, gcc -ffp-contract=off).  Accuracy is < 1 ulp for
 * |x| < 2^20*pi/2, which covers every angle the step produces
 * (angle = |omega|*dt, RigidBody.fs:151-156).
 *
 * Used at: RotationMatrix3D.fromAxisAndAngle (Utilities/Matrix3.fs:67-88).
 */
#ifndef PS_SINCOS_H
#define PS_SINCOS_H

#include <math.h>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#else
#define PS_HD static inline
#endif

PS_HD double ps_ksin(double x, double y)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double w = z * z;
    double r = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    double v = z * x;
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

PS_HD double ps_kcos(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

/* sin and cos of x in one call */
PS_HD void ps_sincos(double x, double* s, double* c)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    const double pio2_2 = 6.07710050630396597660e-11;  /* next 33 bits */
    const double pio2_2t = 2.02226624879595063154e-21; /* pi/2 - (pio2_1+pio2_2) */
    double ax = fabs(x);
    if (!(ax <= 1.0e300)) { /* inf or nan */
        double q = x - x;
        *s = q;
        *c = q;
        return;
    }
    if (ax <= 0.78539816339744830962) {
        *s = ps_ksin(x, 0.0);
        *c = ps_kcos(x, 0.0);
        return;
    }
    double fn = rint(x * invpio2);
    double t = x - fn * pio2_1;
    double w = fn * pio2_2;
    double r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    double y0 = r - w;
    double y1 = (r - y0) - w;
    /* quadrant: fn may exceed int range for silly inputs; fold first */
    double q4 = fn - 4.0 * floor(fn * 0.25);
    int n = (int)q4 & 3;
    double ks = ps_ksin(y0, y1);
    double kc = ps_kcos(y0, y1);
    switch (n) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
    }
}

#endif /* PS_SINCOS_H */
