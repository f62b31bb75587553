// vsb_orbit_math.cuh -- scalar device routines shared by the vessel COI / event kernels
// (vsb_intersect.cu, vsb_events.cu).  Same operation order as the reference; the including
// files are compiled with -fmad=false.
#pragma once
#include "vsb_common.cuh"
#include "vsb_ddtrig.h"

namespace {

constexpr double PI_D = 3.141592653589793;

// "Exact trigonometry" mode of the COI-entry march (vsb_set_march_mode bit 1):
// sin / cos evaluated in double-double arithmetic and rounded to nearest (vsb_ddtrig.h) instead of
// CUDA's 1-2 ulp routines.  ~20x the cost of sin / cos; ON by default (vsb_set_march_mode).  One
// copy per including file (only vsb_intersect.cu sets it, before every launch that reads it).
__constant__ int pe_exact_trig = 0;

__device__ __forceinline__ double pe_sin(double x) { return pe_exact_trig ? vsb_dd_sin(x) : sin(x); }
__device__ __forceinline__ double pe_cos(double x) { return pe_exact_trig ? vsb_dd_cos(x) : cos(x); }
// the same for the cubic / quartic of the orbit-circle intersections (pow, acos, atan2: one Newton
// step in double-double from the library value, rounded to nearest)
__device__ __forceinline__ double pe_pow(double x, double y) { return pe_exact_trig ? vsb_dd_pow(x, y) : pow(x, y); }
__device__ __forceinline__ double pe_acos(double x) { return pe_exact_trig ? vsb_dd_acos(x) : acos(x); }
__device__ __forceinline__ double pe_atan2(double y, double x) { return pe_exact_trig ? vsb_dd_atan2(y, x) : atan2(y, x); }

// Python's float % (sign of the divisor)
__device__ __forceinline__ double py_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

// enhanced_kepler_solver.solve_kepler_ell, enhanced_kepler_solver.py:18-69
__device__ double pe_solve_kepler_ell(double ecc, double ma, double tol)
{
    double mr = ma, flip;
    if (mr > PI_D) {
        mr = 2 * PI_D - mr;
        flip = 1;
    } else {
        flip = -1;
    }
    if (ecc > 0.99 && mr < 0.0045) {
        double al = tol / 1e7, be = tol / 0.3, fp = 2.7 * mr, fpp = 0.301, f = 0.154;
        for (int i = 0; i < 2; ++i) {
            if ((fpp - fp) <= (al + be * f)) break;
            if ((f - ecc * pe_sin(f) - mr) > 0) fpp = f;
            else fp = f;
        }
        f = 0.5 * (fp + fpp);
        return ma + flip * (mr - f);
    }
    double tol2s = 2 * tol / (ecc + 2.2e-16);
    double eapp = mr + 0.999999 * mr * (PI_D - mr) /
                           (2. * mr + ecc - PI_D + 2.4674011002723395 / (ecc + 2.2e-16));
    double fpp = ecc * pe_sin(eapp), fppp = ecc * pe_cos(eapp);
    double fp = 1 - fppp, f = eapp - fpp - mr;
    double delta = -f / fp;
    double fp3 = fp * fp * fp, ffpfpp = f * fp * fpp, f2fppp = f * f * fppp;
    delta = delta * (fp3 - 0.5 * ffpfpp + f2fppp / 3) / (fp3 - ffpfpp + 0.5 * f2fppp);
    for (int i = 0; i < 30; ++i) {
        if (delta * delta < fp * tol2s) break;
        eapp = eapp + delta;
        fp = 1 - ecc * pe_cos(eapp);
        delta = (mr - eapp + ecc * pe_sin(eapp)) / fp;
    }
    eapp = eapp + delta;
    return ma + flip * (mr - eapp);
}

// phys_shared.newton_root_kepler_hyp, phys_shared.py:114-122
__device__ double pe_newton_hyp(double ecc, double ma, double ea)
{
    for (int i = 0; i < 50; ++i) {
        double dx = (ea - ecc * sinh(ea) - ma) / (1.0 - ecc * cosh(ea));
        ea -= dx;
        if (fabs(dx) < 1e-10) return ea;
    }
    return ea;
}

// phys_shared.point_between, phys_shared.py:67-80
__device__ __forceinline__ bool point_between(double n, double a, double b, double dir)
{
    n = py_mod(n, 2 * PI_D);
    a = py_mod(a, 2 * PI_D);
    b = py_mod(b, 2 * PI_D);
    if (dir < 0) {
        const double t = a;
        a = b;
        b = t;
    }
    if (a == b) return a == n;
    if (a < b) return a <= n && n <= b;
    return a <= n || n <= b;
}

// sort_intersect_indices(ea, pts, dir)[0], orbit_intersect.py:68-79: index of the nearest
// non-NaN point in orbit direction (stable), -1 when every point is NaN
__device__ __forceinline__ int first_intersect(double ea_vessel, const double *pts, int cnt,
                                               double dir)
{
    int best = -1;
    double best_ang = 0;
    for (int k = 0; k < cnt; ++k) {
        if (pts[k] != pts[k]) continue;
        double ang = fabs(ea_vessel - pts[k]);
        if (ea_vessel > pts[k]) ang = 2 * PI_D - ang;
        if (dir < 0) ang = PI_D * 2 - ang;
        if (best < 0 || ang < best_ang) {
            best = k;
            best_ang = ang;
        }
    }
    return best;
}

// phys_shared.orb2xy
__device__ void ps_orb2xy(double a, double b, double f, double ecc, double pea, double rx, double ry,
                          double ea, double *out)
{   // phys_shared.py:187-199; ea == 0 gives (0, 0) WITHOUT the reference position
    if (ea == 0) {
        out[0] = out[1] = 0;
        return;
    }
    double xn, yn;
    if (ecc < 1) {
        xn = a * cos(ea) - f;
        yn = b * sin(ea);
    } else {
        xn = a * cosh(ea) - f;
        yn = b * sinh(ea);
    }
    out[0] = (xn * cos(pea - PI_D) - yn * sin(pea - PI_D)) + rx;
    out[1] = (xn * sin(pea - PI_D) + yn * cos(pea - PI_D)) + ry;
}

// phys_shared.orbit_time_to, phys_shared.py:58-64 (Python's max(x, 0.0) keeps a NaN x)
__device__ __forceinline__ double orbit_time_to(double src, double dst, double ecc, double dr, double n)
{
    if (ecc >= 1) {
        const double t = -dr * (dst - src) / n;
        return 0.0 > t ? 0.0 : t;
    }
    if (dr < 0) return py_mod(src - dst, 2 * PI_D) / n;
    return py_mod(dst - src, 2 * PI_D) / n;
}

// COI of a body on its orbit around `parent` (simplified_orbit_coi, phys_editor.py:338-345)
__device__ __forceinline__ double soc_coi_of(double2 pb, double2 vb, double mb, double2 pp,
                                             double2 vp, double mp, double gc, double coi_coef)
{
    double rpx = pp.x - pb.x, rpy = pp.y - pb.y;
    double rvx = vp.x - vb.x, rvy = vp.y - vb.y;
    double u = gc * mp;
    double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / sqrt(rpx * rpx + rpy * rpy)));
    return a > 0 ? a * pow(mb / mp, coi_coef) : 0.0;
}

}  // namespace
