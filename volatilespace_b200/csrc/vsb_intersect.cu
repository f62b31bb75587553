// vsb_intersect.cu -- SURVEY 8f row N2: the numeric kernels of the vessel COI-crossing
// machinery, batched one thread per item (sm_100a):
//   solve_quartic / solve_cubic_one     quartic_solver.py:28-93   (modified Ferrari / Cardano)
//   ell_hyp_intersect_circle            orbit_intersect.py:17-52
//   predict_enter_coi (+ move, orb2xy)  orbit_intersect.py:95-174 (time march + 30-step bisection)
// The branchy per-vessel bookkeeping around them (phys_vessel.points / check_warp / cross_coi)
// stays host side.  Compiled with -fmad=false.
#include "vsb_common.cuh"

namespace {

constexpr double PI_D = 3.141592653589793;

__device__ __forceinline__ double py_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

__device__ double solve_cubic_one(double a, double b, double c)
{
    double q = b / 3 - a * a / 9;
    double r = (b * a - 3 * c) / 6 - a * a * a / 27;
    double rq = r * r + q * q * q;
    if (rq > 0) {
        double aa = pow(fabs(r) + sqrt(rq), 1.0 / 3);
        double t = r >= 0 ? aa - q / aa : q / aa - aa;
        return t - a / 3;
    }
    double theta = 0;
    if (q < 0) theta = acos(r / pow(-q, 3.0 / 2));
    return 2 * sqrt(-q) * cos(theta / 3) - a / 3;
}

// cmath.sqrt of a real argument
__device__ __forceinline__ double2 csqrt_real(double x)
{
    if (x != x) return make_double2(x, x);
    if (x >= 0) return make_double2(sqrt(x), 0.0);
    return make_double2(0.0, sqrt(-x));
}

__device__ void solve_quartic(double a, double b, double c, double d, double e, double2 *z)
{
    double a3 = b / a, a2 = c / a, a1 = d / a, a0 = e / a;
    double cc = a3 / 4, cc2 = cc * cc;
    double b2 = a2 - 6 * cc2;
    double b1 = a1 - 2 * a2 * cc + 8 * (cc * cc2);
    double b0 = a0 - a1 * cc + a2 * cc2 - 3 * (cc2 * cc2);
    double y = solve_cubic_one(b2, b2 * b2 / 4 - b0, -(b1 * b1) / 8);
    if (0 > y) y = 0;
    double s = y * y + b2 * y + b2 * b2 / 4 - b0;
    double r;
    if (s > 0) r = b1 < 0 ? -sqrt(s) : sqrt(s);
    else r = __longlong_as_double(0x7ff8000000000000ll);
    double2 h = csqrt_real(y / 2);
    double2 p = make_double2(h.x - cc, h.y), p1 = make_double2(-h.x - cc, -h.y);
    double q = -y / 2 - b2 / 2;
    double2 s1 = csqrt_real(q - r), s2 = csqrt_real(q + r);
    z[0] = make_double2(p.x + s1.x, p.y + s1.y);
    z[1] = make_double2(p.x - s1.x, p.y - s1.y);
    z[2] = make_double2(p1.x + s2.x, p1.y + s2.y);
    z[3] = make_double2(p1.x - s2.x, p1.y - s2.y);
}

__global__ void quartic_kernel(int64_t n, const double *__restrict__ coef /* (n,5) */,
                               double *__restrict__ roots /* (n,8) re,im */)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 z[4];
    solve_quartic(coef[5 * i], coef[5 * i + 1], coef[5 * i + 2], coef[5 * i + 3], coef[5 * i + 4], z);
    for (int k = 0; k < 4; ++k) {
        roots[8 * i + 2 * k] = z[k].x;
        roots[8 * i + 2 * k + 1] = z[k].y;
    }
}

// in6 = a b ecc c_x c_y r per item; ea (n,4) NaN padded, cnt (n) (0 = the reference's [nan])
__global__ void intersect_kernel(int64_t n, const double *__restrict__ in6,
                                 double *__restrict__ ea_out, int64_t *__restrict__ cnt_out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = in6[6 * i], b = in6[6 * i + 1], ecc = in6[6 * i + 2], c_x = in6[6 * i + 3],
                 c_y = in6[6 * i + 4], r = in6[6 * i + 5];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double2 z[4];
    double rr[4];
    int cnt = 0;
    bool any = false;
    double a_0, a_1, a_2, a_3, a_4;
    if (ecc < 1) {
        a_0 = a * a - 2 * a * c_x + c_x * c_x + c_y * c_y - r * r;
        a_1 = -4 * b * c_y;
        a_2 = -2 * (a * a) + 4 * (b * b) + 2 * (c_x * c_x) + 2 * (c_y * c_y) - 2 * (r * r);
        a_3 = -4 * b * c_y;
        a_4 = a * a + 2 * a * c_x + c_x * c_x + c_y * c_y - r * r;
    } else {
        a_0 = -(a * a) * (c_y * c_y + b * b) + b * b * ((c_x + r) * (c_x + r));
        a_1 = -4 * (a * a) * r * c_y;
        a_2 = -2 * (a * a * (c_y * c_y + b * b + 2 * (r * r)) + b * b * (r * r - c_x * c_x));
        a_3 = -4 * (a * a) * r * c_y;
        a_4 = -(a * a) * (c_y * c_y + b * b) + b * b * ((c_x - r) * (c_x - r));
    }
    solve_quartic(a_4, a_3, a_2, a_1, a_0, z);
    for (int k = 0; k < 4; ++k)
        if (z[k].y == 0.0) rr[cnt++] = z[k].x;
    for (int k = 0; k < cnt; ++k) any = any || rr[k] != 0.0;
    double out[4] = {nan, nan, nan, nan};
    if (!any) cnt = 0;
    for (int k = 0; k < cnt; ++k) {
        if (ecc < 1) {
            out[k] = py_mod(atan2(2 * rr[k], 1.0 - rr[k] * rr[k]), 2 * PI_D);
        } else {
            double x = c_x + r * (1 - rr[k] * rr[k]) / (1 + rr[k] * rr[k]);
            double y = c_y + r * 2 * rr[k] / (1 + rr[k] * rr[k]);
            double ta = atan2(y, x - (a * ecc));
            double ea = acosh((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
            out[k] = rr[k] < 0 ? -ea : ea;
        }
    }
    for (int k = 0; k < 4; ++k) ea_out[4 * i + k] = out[k];
    cnt_out[i] = cnt;
}

// --- predict_enter_coi ---------------------------------------------------------------------
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
            if ((f - ecc * sin(f) - mr) > 0) fpp = f;
            else fp = f;
        }
        f = 0.5 * (fp + fpp);
        return ma + flip * (mr - f);
    }
    double tol2s = 2 * tol / (ecc + 2.2e-16);
    double eapp = mr + 0.999999 * mr * (PI_D - mr) /
                           (2. * mr + ecc - PI_D + 2.4674011002723395 / (ecc + 2.2e-16));
    double fpp = ecc * sin(eapp), fppp = ecc * cos(eapp);
    double fp = 1 - fppp, f = eapp - fpp - mr;
    double delta = -f / fp;
    double fp3 = fp * fp * fp, ffpfpp = f * fp * fpp, f2fppp = f * f * fppp;
    delta = delta * (fp3 - 0.5 * ffpfpp + f2fppp / 3) / (fp3 - ffpfpp + 0.5 * f2fppp);
    for (int i = 0; i < 30; ++i) {
        if (delta * delta < fp * tol2s) break;
        eapp = eapp + delta;
        fp = 1 - ecc * cos(eapp);
        delta = (mr - eapp + ecc * sin(eapp)) / fp;
    }
    eapp = eapp + delta;
    return ma + flip * (mr - eapp);
}

__device__ double pe_newton_hyp(double ecc, double ma, double ea)
{
    for (int i = 0; i < 50; ++i) {
        double dx = (ea - ecc * sinh(ea) - ma) / (1.0 - ecc * cosh(ea));
        ea -= dx;
        if (fabs(dx) < 1e-10) return ea;
    }
    return ea;
}

struct orbit {
    double a, b, f, ecc, ma, ea, pea, n, dr, prev_ea;
};

__device__ __forceinline__ void oi_move(orbit &o, double dt)
{   // orbit_intersect.py:108-121 (its two wrap statements are no-ops)
    if (o.ecc < 1) {
        o.ma += o.dr * o.n * dt;
        o.ea = pe_solve_kepler_ell(o.ecc, o.ma, 1e-10);
    } else {
        o.ma += o.dr * o.n * -1 * dt;
        o.ea = pe_newton_hyp(o.ecc, o.ma, o.prev_ea);
        o.prev_ea = o.ea;
    }
}

__device__ __forceinline__ double2 oi_xy(const orbit &o)
{   // orbit_intersect.py:95-105
    double xn, yn;
    if (o.ecc < 1) {
        xn = o.a * cos(o.ea) - o.f;
        yn = o.b * sin(o.ea);
    } else {
        xn = o.a * cosh(o.ea) - o.f;
        yn = o.b * sinh(o.ea);
    }
    const double c = cos(o.pea - PI_D), s = sin(o.pea - PI_D);
    return make_double2(xn * c - yn * s, xn * s + yn * c);
}

__device__ __forceinline__ double oi_wrap(double angle)
{
    if (angle >= 2 * PI_D) return angle - 2 * PI_D;
    if (angle < 0) return angle + 2 * PI_D;
    return angle;
}

__device__ __forceinline__ double oi_dist(const orbit &v, const orbit &b)
{
    double2 pv = oi_xy(v), pb = oi_xy(b);
    double dx = pv.x - pb.x, dy = pv.y - pb.y;
    return pow(dx * dx + dy * dy, 0.5);
}

// vd (n,10) = a b f ecc ma ea pea period n dr ; bd (n,10) = a b f ecc ma ea pea n dr coi
__global__ void __launch_bounds__(128)
predict_enter_coi_kernel(int64_t n, const double *__restrict__ vd, const double *__restrict__ bd,
                         double tol, int steps_limit, double *__restrict__ out5)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *V = vd + 10 * i, *B = bd + 10 * i;
    orbit v{V[0], V[1], V[2], V[3], V[4], V[5], V[6], V[8], V[9], V[5]};
    orbit b{B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], B[8], B[5]};
    const double period = V[7], b_coi = B[9];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double *o = out5 + 5 * i;
    for (int k = 0; k < 5; ++k) o[k] = nan;
    if (period == 0 || b_coi == 0) return;
    double t = 0;
    const double c1 = period / b_coi * 5, c2 = period / 2, c3 = period / 30;
    const double m1 = c2 < c1 ? c2 : c1;
    const double m2 = c3 > m1 ? c3 : m1;
    long long steps = (long long)m2;
    if (steps_limit > steps) steps = steps_limit;
    double dt = period / (double)steps;
    while (t < period) {
        oi_move(v, dt);
        oi_move(b, dt);
        double d = oi_dist(v, b);
        if (d < b_coi) {
            for (int it = 0; it < 30; ++it) {
                dt = fabs(dt) / 2 * (d < b_coi ? -1 : 1);
                t += dt;
                oi_move(v, dt);
                oi_move(b, dt);
                d = oi_dist(v, b);
                if (fabs(d - b_coi) < tol) break;
            }
            o[0] = oi_wrap(v.ea);
            o[1] = oi_wrap(b.ea);
            o[2] = oi_wrap(v.ma);
            o[3] = oi_wrap(b.ma);
            o[4] = t;
            return;
        }
        t += dt;
    }
}

}  // namespace

int vsb_launch_quartic(vsb_ctx *c, int64_t n, const double *d_coef, double *d_roots)
{
    if (n <= 0) return VSB_OK;
    quartic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, d_coef, d_roots);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_intersect(vsb_ctx *c, int64_t n, const double *d_in6, double *d_ea, int64_t *d_cnt)
{
    if (n <= 0) return VSB_OK;
    intersect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, d_in6, d_ea, d_cnt);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_predict_enter_coi(vsb_ctx *c, int64_t n, const double *d_vd, const double *d_bd,
                                 double tol, int steps_limit, double *d_out5)
{
    if (n <= 0) return VSB_OK;
    predict_enter_coi_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(
        n, d_vd, d_bd, tol, steps_limit, d_out5);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
