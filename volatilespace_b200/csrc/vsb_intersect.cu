// vsb_intersect.cu -- SURVEY 8f row N2: the numeric kernels of the vessel COI-crossing
// machinery, batched one thread per item (sm_100a):
//   solve_quartic / solve_cubic_one     quartic_solver.py:28-93   (modified Ferrari / Cardano)
//   ell_hyp_intersect_circle            orbit_intersect.py:17-52
//   predict_enter_coi (+ move, orb2xy)  orbit_intersect.py:95-174 (time march + 30-step bisection)
//   phys_vessel.Physics.points          phys_vessel.py:372-474     (batched over vessels / pairs)
// predict_enter_coi is a serial time march in the reference (up to ~5e7 steps for one pair); here
// one CTA marches one (vessel, body) pair in parallel and still reproduces the sequential
// recurrences bit for bit: see "parallel march" below.  Compiled with -fmad=false.
#include "vsb_common.cuh"
#include "vsb_orbit_math.cuh"
#include <vector>

namespace {


__device__ double solve_cubic_one(double a, double b, double c)
{
    double q = b / 3 - a * a / 9;
    double r = (b * a - 3 * c) / 6 - a * a * a / 27;
    double rq = r * r + q * q * q;
    if (rq > 0) {
        double aa = pe_pow(fabs(r) + sqrt(rq), 1.0 / 3);
        double t = r >= 0 ? aa - q / aa : q / aa - aa;
        return t - a / 3;
    }
    double theta = 0;
    if (q < 0) theta = pe_acos(r / pe_pow(-q, 3.0 / 2));
    return 2 * sqrt(-q) * pe_cos(theta / 3) - a / 3;
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

// orbit_intersect.py:17-52: eccentric anomalies of the orbit/circle intersections into out[4]
// (NaN padded); returns their number (0 = the reference's [nan]).
__device__ int intersect_circle(double a, double b, double ecc, double c_x, double c_y, double r,
                                double *out)
{
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
    for (int k = 0; k < 4; ++k) out[k] = nan;
    if (!any) cnt = 0;
    for (int k = 0; k < cnt; ++k) {
        if (ecc < 1) {
            out[k] = py_mod(pe_atan2(2 * rr[k], 1.0 - rr[k] * rr[k]), 2 * PI_D);
        } else {
            double x = c_x + r * (1 - rr[k] * rr[k]) / (1 + rr[k] * rr[k]);
            double y = c_y + r * 2 * rr[k] / (1 + rr[k] * rr[k]);
            double ta = pe_atan2(y, x - (a * ecc));
            double ea = acosh((ecc + pe_cos(ta)) / (1 + (ecc * pe_cos(ta))));
            out[k] = rr[k] < 0 ? -ea : ea;
        }
    }
    return cnt;
}

// in6 = a b ecc c_x c_y r per item; ea (n,4) NaN padded, cnt (n) (0 = the reference's [nan])
__global__ void intersect_kernel(int64_t n, const double *__restrict__ in6,
                                 double *__restrict__ ea_out, int64_t *__restrict__ cnt_out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double out[4];
    const int cnt = intersect_circle(in6[6 * i], in6[6 * i + 1], in6[6 * i + 2], in6[6 * i + 3],
                                     in6[6 * i + 4], in6[6 * i + 5], out);
    for (int k = 0; k < 4; ++k) ea_out[4 * i + k] = out[k];
    cnt_out[i] = cnt;
}

// --- predict_enter_coi ---------------------------------------------------------------------
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
        xn = o.a * pe_cos(o.ea) - o.f;
        yn = o.b * pe_sin(o.ea);
    } else {
        xn = o.a * cosh(o.ea) - o.f;
        yn = o.b * sinh(o.ea);
    }
    const double c = pe_cos(o.pea - PI_D), s = pe_sin(o.pea - PI_D);
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

// A march of more than 2^32 steps (hours of serial work in the reference) is refused: the pair is
// answered "no entry" and counted, instead of occupying the GPU for minutes.
constexpr long long MARCH_MAX_STEPS = 1ll << 32;

// The reference's serial march, one thread per item (kept as the cross-check of the parallel
// march and as its fallback).  vd = a b f ecc ma ea pea period n dr; bd = a b f ecc ma ea pea n dr coi
__device__ void march_sequential(const double *V, const double *B, double tol, int steps_limit,
                                 double *o)
{
    orbit v{V[0], V[1], V[2], V[3], V[4], V[5], V[6], V[8], V[9], V[5]};
    orbit b{B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], B[8], B[5]};
    const double period = V[7], b_coi = B[9];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int k = 0; k < 5; ++k) o[k] = nan;
    if (period == 0 || b_coi == 0) return;
    double t = 0;
    const double c1 = period / b_coi * 5, c2 = period / 2, c3 = period / 30;
    const double m1 = c2 < c1 ? c2 : c1;
    const double m2 = c3 > m1 ? c3 : m1;
    long long steps = (long long)m2;
    if (steps_limit > steps) steps = steps_limit;
    if (steps > MARCH_MAX_STEPS) return;
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

__global__ void __launch_bounds__(128)
predict_enter_coi_seq_kernel(int64_t n, const double *__restrict__ vd, const double *__restrict__ bd,
                             const unsigned char *__restrict__ valid, double tol, int steps_limit,
                             double *__restrict__ out5)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (valid && !valid[i]) {
        for (int k = 0; k < 5; ++k) out5[5 * i + k] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    march_sequential(vd + 10 * i, bd + 10 * i, tol, steps_limit, out5 + 5 * i);
}

// --- parallel march ------------------------------------------------------------------------
// The march carries three running sums x_k = fl(x_{k-1} + c) (the time t and the two mean
// anomalies) and, for a hyperbolic orbit, a Newton chain ea_k = newton(ma_k, guess = ea_{k-1}).
//  * Running sums: while x stays inside one binade, fl(x + c) = x + cr with cr = c rounded to the
//    binade's ulp (no tie), so x_k is an exact closed form per binade.  march_walk() lists those
//    segments (a few per binade, O(log K) in all) and any thread can then evaluate x_k for any k
//    with the very bits the serial loop would hold.  Ties, zero crossings and binade changes are
//    single true additions.
//  * Newton chains: each thread marches a chunk of consecutive steps from a speculated start
//    (a short warm-up from a closed-form guess) and the CTA then checks, chunk by chunk, that the
//    speculated start equals the predecessor's final value bit for bit; a chunk whose start was
//    wrong is re-run from the true value until the whole wave is one consistent chain.
//  Elliptic orbits are stateless (ENRKE of ma_k): one step per thread, no speculation.
struct mseg {
    long long k0, m;   // steps k in (k0, k0+m]
    double xb, cr;     // x_k = xb + (k - k0 - 1) * cr   (exact)
};
constexpr int MSEG_CAP = 288;
constexpr int MARCH_T = 256;       // threads per pair
constexpr int MARCH_C = 16;        // chunk of consecutive steps per thread when a chain exists
constexpr int MARCH_W = 8;         // warm-up steps of a speculated chain

// Lists the segments of x_k = fl(x_{k-1} + c), x_0 = x0, for k = 1..K where K = kmax, or (bounded)
// the first k with !(x_k < bound) -- the trip count of `while t < t_end`.
// fail: 1 = too many segments, 2 = a bounded chain that never reaches the bound.
__host__ __device__ void march_walk(double x0, double c, long long kmax, bool bounded, double bound,
                           mseg *seg, int *nseg_out, long long *K_out, int *fail)
{
    const long long LO = (1ll << 52) + 2, HI = (1ll << 53) - 2;
    long long k = 0;
    double x = x0;
    int ns = 0;
    *fail = 0;
    for (;;) {
        if (bounded && !(x < bound)) break;
        if (k >= kmax) {
            if (bounded) *fail = 2;
            break;
        }
        if (ns >= MSEG_CAP) {
            *fail = 1;
            break;
        }
        bool single = true;
        long long m = 0;
        double cr = 0;
        if (x != 0 && isfinite(x) && isfinite(c) && c != 0) {
            const int e = ilogb(x);
            if (e > -900 && e < 900) {
                const double q = scalbn(c, 52 - e);
                if (fabs(q) < 4503599627370496.0) {
                    const double qr = rint(q);      // ties to even
                    const long long X = (long long)scalbn(x, 52 - e), CR = (long long)qr;
                    // a tie (c = (n + 1/2) ulp) rounds to the even neighbour: from an even X that is
                    // X + rint(q) again and X stays even; from an odd X take one true step first
                    if (fabs(q - qr) != 0.5 || !(X & 1)) {
                        const long long lo = X > 0 ? LO : -HI, hi = X > 0 ? HI : -LO;
                        if (CR == 0) {                 // |c| below half an ulp: x stays put
                            if (bounded) {
                                *fail = 2;
                                break;
                            }
                            m = kmax - k;
                            single = false;
                        } else if (X >= lo && X <= hi) {
                            m = CR > 0 ? (hi - X) / CR : (X - lo) / (-CR);
                            if (m >= 1) {
                                single = false;
                                cr = scalbn((double)CR, e - 52);
                            }
                        }
                    }
                }
            }
        }
        if (single) {
            const double x1 = x + c;
            seg[ns++] = mseg{k, 1, x1, 0.0};
            x = x1;
            k += 1;
            continue;
        }
        if (m > kmax - k) m = kmax - k;
        if (bounded && cr > 0) {
            const double jd = ceil((bound - x) / cr);   // estimate, corrected exactly below
            if (jd <= (double)m + 2) {
                long long j = jd < 1 ? 1 : (jd > (double)m ? m : (long long)jd);
                while (j > 1 && x + (double)(j - 1) * cr >= bound) --j;
                while (j < m && x + (double)j * cr < bound) ++j;
                m = j;
            }
        }
        seg[ns++] = mseg{k, m, x + cr, cr};
        x = x + (double)m * cr;
        k += m;
    }
    *nseg_out = ns;
    *K_out = k;
}

__host__ __device__ __forceinline__ double march_at(const mseg *seg, int &idx, long long k)
{   // k >= 1, inside the walked range; idx is a monotone cursor
    while (k > seg[idx].k0 + seg[idx].m) ++idx;
    return seg[idx].xb + (double)(k - seg[idx].k0 - 1) * seg[idx].cr;
}

__host__ __device__ inline int march_find(const mseg *seg, int ns, long long k)
{
    int lo = 0, hi = ns - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (k > seg[mid].k0 + seg[mid].m) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

struct morbit {   // loop invariants of one orbit
    double a, b, f, ecc, cp, sp;
    bool hyp;
};

__device__ __forceinline__ double m_solve(const morbit &o, double ma, double prev_ea)
{
    return o.hyp ? pe_newton_hyp(o.ecc, ma, prev_ea) : pe_solve_kepler_ell(o.ecc, ma, 1e-10);
}

__device__ __forceinline__ double2 m_xy(const morbit &o, double ea)
{
    double xn, yn;
    if (!o.hyp) {
        xn = o.a * pe_cos(ea) - o.f;
        yn = o.b * pe_sin(ea);
    } else {
        xn = o.a * cosh(ea) - o.f;
        yn = o.b * sinh(ea);
    }
    return make_double2(xn * o.cp - yn * o.sp, xn * o.sp + yn * o.cp);
}

struct mchunk {
    double v_end, b_end;            // chain values after the last step of the chunk
    long long hit;                  // first step with d < coi, or -1
    double hv_ma, hv_ea, hb_ma, hb_ea, hd;
};

__device__ void march_chunk(const morbit &v, const morbit &b, const mseg *sv, const mseg *sb,
                            int iv, int ib, long long s, long long e, double v_prev, double b_prev,
                            double b_coi, mchunk &r)
{
    r.hit = -1;
    for (long long k = s; k <= e; ++k) {
        const double vma = march_at(sv, iv, k), bma = march_at(sb, ib, k);
        const double vea = m_solve(v, vma, v_prev), bea = m_solve(b, bma, b_prev);
        if (v.hyp) v_prev = vea;
        if (b.hyp) b_prev = bea;
        if (r.hit < 0) {
            const double2 pv = m_xy(v, vea), pb = m_xy(b, bea);
            const double dx = pv.x - pb.x, dy = pv.y - pb.y;
            const double d = pow(dx * dx + dy * dy, 0.5);
            if (d < b_coi) {
                r.hit = k;
                r.hv_ma = vma; r.hv_ea = vea; r.hb_ma = bma; r.hb_ea = bea; r.hd = d;
                if (!v.hyp && !b.hyp) break;
            }
        }
    }
    r.v_end = v_prev;
    r.b_end = b_prev;
}

// closed-form starting guess of a speculated hyperbolic chain (root of ea - ecc*sinh(ea) - ma)
__device__ __forceinline__ double hyp_guess(double ecc, double ma)
{
    const double m = -ma;
    const double h = log(2 * fabs(m) / ecc + 1.8);
    return m < 0 ? -h : h;
}

// Per-pair constants of the march, computed identically by every thread of the CTA.
struct msetup {
    morbit v, b;
    double dt, period, b_coi;
    double v_ma0, v_ea0, v_pea, v_n, v_dr, b_ma0, b_ea0, b_pea, b_n, b_dr, cv, cb;
    bool chained;
};

__device__ __forceinline__ msetup march_setup(const double *V, const double *B, int steps_limit)
{
    msetup m;
    m.period = V[7];
    m.b_coi = B[9];
    const double c1 = m.period / m.b_coi * 5, c2 = m.period / 2, c3 = m.period / 30;
    const double m1 = c2 < c1 ? c2 : c1;
    const double m2 = c3 > m1 ? c3 : m1;
    long long steps = (long long)m2;
    if (steps_limit > steps) steps = steps_limit;
    m.dt = m.period / (double)steps;
    m.v = morbit{V[0], V[1], V[2], V[3], pe_cos(V[6] - PI_D), pe_sin(V[6] - PI_D), !(V[3] < 1)};
    m.b = morbit{B[0], B[1], B[2], B[3], pe_cos(B[6] - PI_D), pe_sin(B[6] - PI_D), !(B[3] < 1)};
    m.v_ma0 = V[4]; m.v_ea0 = V[5]; m.v_pea = V[6]; m.v_n = V[8]; m.v_dr = V[9];
    m.b_ma0 = B[4]; m.b_ea0 = B[5]; m.b_pea = B[6]; m.b_n = B[7]; m.b_dr = B[8];
    m.cv = m.v.hyp ? m.v_dr * m.v_n * -1 * m.dt : m.v_dr * m.v_n * m.dt;
    m.cb = m.b.hyp ? m.b_dr * m.b_n * -1 * m.dt : m.b_dr * m.b_n * m.dt;
    m.chained = m.v.hyp || m.b.hyp;
    return m;
}

struct mshared {
    mseg seg_t[MSEG_CAP], seg_v[MSEG_CAP], seg_b[MSEG_CAP];
    int ns[3], fail[3];
    long long K;
};

// CTA-collective: the three segment lists of one pair.  Returns 0 = ready, 1 = the reference
// would never return (t stops growing): answer NaN, 2 = no closed form: serial fallback.
__device__ int march_segments(const msetup &m, mshared &sh)
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        long long K;
        march_walk(0.0, m.dt, 1ll << 60, true, m.period, sh.seg_t, &sh.ns[0], &K, &sh.fail[0]);
        sh.K = K;
    }
    __syncthreads();
    const long long K = sh.K;
    if (tid == 32) {
        long long kk;
        march_walk(m.v_ma0, m.cv, K, false, 0.0, sh.seg_v, &sh.ns[1], &kk, &sh.fail[1]);
    }
    if (tid == 64) {
        long long kk;
        march_walk(m.b_ma0, m.cb, K, false, 0.0, sh.seg_b, &sh.ns[2], &kk, &sh.fail[2]);
    }
    __syncthreads();
    if (sh.fail[0] == 2) return 1;
    if (sh.fail[0] | sh.fail[1] | sh.fail[2]) return 2;
    return 0;
}

// the 30-step refinement of the reference (orbit_intersect.py:151-168) from the exact state at
// the first step inside the COI
__device__ void march_refine(const msetup &m, const mshared &sh, const mchunk &r, double tol,
                             double *o)
{
    orbit ov{m.v.a, m.v.b, m.v.f, m.v.ecc, r.hv_ma, r.hv_ea, m.v_pea, m.v_n, m.v_dr,
             m.v.hyp ? r.hv_ea : m.v_ea0};
    orbit ob{m.b.a, m.b.b, m.b.f, m.b.ecc, r.hb_ma, r.hb_ea, m.b_pea, m.b_n, m.b_dr,
             m.b.hyp ? r.hb_ea : m.b_ea0};
    double t = 0, d = r.hd, dtt = m.dt;
    if (r.hit > 1) {
        int it_ = march_find(sh.seg_t, sh.ns[0], r.hit - 1);
        t = march_at(sh.seg_t, it_, r.hit - 1);
    }
    for (int it = 0; it < 30; ++it) {
        dtt = fabs(dtt) / 2 * (d < m.b_coi ? -1 : 1);
        t += dtt;
        oi_move(ov, dtt);
        oi_move(ob, dtt);
        d = oi_dist(ov, ob);
        if (fabs(d - m.b_coi) < tol) break;
    }
    o[0] = oi_wrap(ov.ea);
    o[1] = oi_wrap(ob.ea);
    o[2] = oi_wrap(ov.ma);
    o[3] = oi_wrap(ob.ma);
    o[4] = t;
}

// Long elliptic-elliptic pairs are split over a team of CTAs (team kernel below).
constexpr long long TEAM_MIN_STEPS = 64 * MARCH_T;    // below: one CTA marches the whole pair
constexpr int TEAM_MAX = 256;                         // parts per pair
constexpr int TEAM_TASK_CAP = 1 << 16;
struct mtask {
    long long item;
    int part, parts;
};
struct mteam {          // device-global bookkeeping of one launch pair
    unsigned long long stats[4];      // chunks, re-runs, serial fallbacks, item counter
    unsigned long long ntasks, next_task;
    unsigned long long refused;       // pairs longer than MARCH_MAX_STEPS (answered NaN)
};


// stats[0] += chunks marched, stats[1] += chunks re-run after a failed speculation,
// stats[2] += items that fell back to the serial march, stats[3] = work counter (zero at launch)
__global__ void __launch_bounds__(MARCH_T)
enter_coi_march_kernel(int64_t n, const double *__restrict__ vd, const double *__restrict__ bd,
                       const unsigned char *__restrict__ valid, double tol, int steps_limit,
                       double *__restrict__ out5, mteam *__restrict__ team,
                       mtask *__restrict__ tasks, unsigned long long *__restrict__ best_hit,
                       unsigned int *__restrict__ parts_done)
{
    __shared__ mshared sh;
    __shared__ double end_v[MARCH_T], end_b[MARCH_T], carry[2];
    __shared__ unsigned long long hit_sh, item_sh, slot_sh;
    const int tid = threadIdx.x;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned long long *stats = team->stats;

    for (;;) {
        __syncthreads();                       // shared state of the previous item is dead
        if (tid == 0) item_sh = atomicAdd(&stats[3], 1ull);   // march lengths differ by 1e5x
        __syncthreads();
        const int64_t item = (int64_t)item_sh;
        if (item >= n) break;
        const double *V = vd + 10 * item, *B = bd + 10 * item;
        double *o = out5 + 5 * item;
        if ((valid && !valid[item]) || V[7] == 0 || B[9] == 0) {
            if (tid < 5) o[tid] = nan;
            continue;
        }
        const msetup m = march_setup(V, B, steps_limit);
        const int st = march_segments(m, sh);
        if (st == 1) {                          // CTA-uniform
            if (tid < 5) o[tid] = nan;
            continue;
        }
        if (st == 2) {
            if (tid == 0) {
                march_sequential(V, B, tol, steps_limit, o);
                atomicAdd(&stats[2], 1ull);
            }
            continue;
        }
        const long long K = sh.K;
        if (K > MARCH_MAX_STEPS) {
            if (tid < 5) o[tid] = nan;
            if (tid == 0) atomicAdd(&team->refused, 1ull);
            continue;
        }
        const morbit &v = m.v, &b = m.b;
        const double b_coi = m.b_coi;
        const bool chained = m.chained;
        if (!chained && K > TEAM_MIN_STEPS && tasks) {
            // hand the pair to a team: `parts` CTAs of the team kernel take its waves round-robin
            long long parts = (K + TEAM_MIN_STEPS - 1) / TEAM_MIN_STEPS;
            if (parts > TEAM_MAX) parts = TEAM_MAX;
            if (tid == 0) {
                unsigned long long slot = *(volatile unsigned long long *)&team->ntasks;
                for (;;) {                      // reserve `parts` consecutive task slots
                    if (slot + parts > TEAM_TASK_CAP) {
                        slot = ~0ull;
                        break;
                    }
                    const unsigned long long seen = atomicCAS(&team->ntasks, slot, slot + parts);
                    if (seen == slot) break;
                    slot = seen;
                }
                if (slot != ~0ull) {
                    best_hit[item] = ~0ull;
                    parts_done[item] = 0;
                    for (int j = 0; j < (int)parts; ++j)
                        tasks[slot + j] = mtask{(long long)item, j, (int)parts};
                }
                slot_sh = slot;
            }
            __syncthreads();
            if (slot_sh != ~0ull) continue;     // otherwise: no room, march it here
        }
        const int C = chained ? MARCH_C : 1;
        bool done = false;
        unsigned long long n_chunks = 0, n_rerun = 0;
        for (long long base = 0; base < K && !done; base += (long long)MARCH_T * C) {
            const long long s = base + (long long)tid * C + 1;
            const long long e = s + C - 1 < K ? s + C - 1 : K;
            const bool active = s <= K;
            double sv = m.v_ea0, sb = m.b_ea0;
            int iv = 0, ib = 0;
            mchunk r;
            r.hit = -1;
            r.v_end = sv;
            r.b_end = sb;
            if (tid == 0) hit_sh = ~0ull;
            if (active) {
                if (chained && s > 1) {
                    if (tid == 0) {
                        sv = carry[0];
                        sb = carry[1];
                    } else {
                        // speculate the chain value at step s-1: warm up from step w0
                        const long long w0 = s - MARCH_W > 1 ? s - MARCH_W : 1;
                        int jv = march_find(sh.seg_v, sh.ns[1], w0), jb = march_find(sh.seg_b, sh.ns[2], w0);
                        if (w0 > 1) {
                            if (v.hyp) sv = hyp_guess(v.ecc, march_at(sh.seg_v, jv, w0));
                            if (b.hyp) sb = hyp_guess(b.ecc, march_at(sh.seg_b, jb, w0));
                        }
                        for (long long k = w0; k < s; ++k) {
                            if (v.hyp) sv = pe_newton_hyp(v.ecc, march_at(sh.seg_v, jv, k), sv);
                            if (b.hyp) sb = pe_newton_hyp(b.ecc, march_at(sh.seg_b, jb, k), sb);
                        }
                    }
                }
                iv = march_find(sh.seg_v, sh.ns[1], s);
                ib = march_find(sh.seg_b, sh.ns[2], s);
                march_chunk(v, b, sh.seg_v, sh.seg_b, iv, ib, s, e, sv, sb, b_coi, r);
                ++n_chunks;
            }
            if (chained) {
                end_v[tid] = r.v_end;
                end_b[tid] = r.b_end;
                __syncthreads();
                for (;;) {
                    bool mism = false;
                    double tv = 0, tb = 0;
                    if (active && tid > 0) {
                        tv = end_v[tid - 1];
                        tb = end_b[tid - 1];
                        mism = (v.hyp && __double_as_longlong(tv) != __double_as_longlong(sv)) ||
                               (b.hyp && __double_as_longlong(tb) != __double_as_longlong(sb));
                    }
                    if (!__syncthreads_or(mism)) break;
                    if (mism) {
                        sv = tv;
                        sb = tb;
                        march_chunk(v, b, sh.seg_v, sh.seg_b, iv, ib, s, e, sv, sb, b_coi, r);
                        ++n_rerun;
                        end_v[tid] = r.v_end;
                        end_b[tid] = r.b_end;
                    }
                    __syncthreads();
                }
            } else {
                __syncthreads();
            }
            if (r.hit >= 0) atomicMin(&hit_sh, (unsigned long long)r.hit);
            __syncthreads();
            const unsigned long long hit = hit_sh;
            if (hit != ~0ull) {
                if (r.hit >= 0 && (unsigned long long)r.hit == hit) march_refine(m, sh, r, tol, o);
                done = true;
            } else if (chained && tid == MARCH_T - 1) {
                carry[0] = r.v_end;
                carry[1] = r.b_end;
            }
            __syncthreads();
        }
        if (!done && tid < 5) o[tid] = nan;
        if (n_chunks) atomicAdd(&stats[0], n_chunks);
        if (n_rerun) atomicAdd(&stats[1], n_rerun);
    }
}

// Team kernel: the long elliptic-elliptic pairs queued by the kernel above.  Part j of a pair
// marches the waves j, j+parts, ... in increasing order and stops as soon as its next wave lies
// beyond the best hit known so far (a device-wide atomicMin); the last part to finish recomputes
// the state at the winning step (stateless for elliptic orbits) and runs the refinement.
__global__ void __launch_bounds__(MARCH_T)
enter_coi_team_kernel(const double *__restrict__ vd, const double *__restrict__ bd, double tol,
                      int steps_limit, double *__restrict__ out5, mteam *__restrict__ team,
                      const mtask *__restrict__ tasks, unsigned long long *__restrict__ best_hit,
                      unsigned int *__restrict__ parts_done)
{
    __shared__ mshared sh;
    __shared__ unsigned long long hit_sh, task_sh;
    __shared__ unsigned int ticket_sh;
    const int tid = threadIdx.x;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const unsigned long long ntasks = team->ntasks;
    for (;;) {
        __syncthreads();
        if (tid == 0) task_sh = atomicAdd(&team->next_task, 1ull);
        __syncthreads();
        if (task_sh >= ntasks) break;
        const mtask tk = tasks[task_sh];
        const double *V = vd + 10 * tk.item, *B = bd + 10 * tk.item;
        double *o = out5 + 5 * tk.item;
        const msetup m = march_setup(V, B, steps_limit);
        march_segments(m, sh);                  // cannot fail: the first kernel walked the same sums
        const long long K = sh.K;
        unsigned long long n_chunks = 0;
        volatile unsigned long long *best = best_hit + tk.item;
        for (long long base = (long long)tk.part * MARCH_T; base < K;
             base += (long long)tk.parts * MARCH_T) {
            if (tid == 0) hit_sh = *best;
            __syncthreads();
            if (hit_sh < (unsigned long long)(base + 1)) break;     // CTA-uniform
            __syncthreads();
            if (tid == 0) hit_sh = ~0ull;
            __syncthreads();
            const long long s = base + tid + 1;
            if (s <= K) {
                mchunk r;
                const int iv = march_find(sh.seg_v, sh.ns[1], s), ib = march_find(sh.seg_b, sh.ns[2], s);
                march_chunk(m.v, m.b, sh.seg_v, sh.seg_b, iv, ib, s, s, 0.0, 0.0, m.b_coi, r);
                ++n_chunks;
                if (r.hit >= 0) atomicMin(&hit_sh, (unsigned long long)r.hit);
            }
            __syncthreads();
            if (hit_sh != ~0ull) {
                if (tid == 0) atomicMin(best_hit + tk.item, hit_sh);
                break;                                               // later waves are farther
            }
        }
        if (n_chunks) atomicAdd(&team->stats[0], n_chunks);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            ticket_sh = atomicAdd(parts_done + tk.item, 1u);
        }
        __syncthreads();
        if (ticket_sh == (unsigned)tk.parts - 1) {                   // last part of the pair
            __threadfence();
            const unsigned long long k = *best;
            if (k == ~0ull) {
                if (tid < 5) o[tid] = nan;
            } else if (tid == 0) {
                mchunk r;
                const int iv = march_find(sh.seg_v, sh.ns[1], (long long)k),
                          ib = march_find(sh.seg_b, sh.ns[2], (long long)k);
                march_chunk(m.v, m.b, sh.seg_v, sh.seg_b, iv, ib, (long long)k, (long long)k, 0.0, 0.0,
                            m.b_coi, r);
                march_refine(m, sh, r, tol, o);
            }
        }
    }
}

// --- phys_vessel.Physics.points, phys_vessel.py:372-474 --------------------------------------
// vessel12 = a b f ecc ma ea pea n dr pe_d ap_d period; body12 = a b f ecc ma ea pea n dr coi size atm
__device__ void next_point(double ea_vessel, const double *pts, int cnt, double dir, double *out2)
{   // orbit_intersect.py:55-65
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    out2[0] = out2[1] = nan;
    if (cnt == 0) return;
    double ang[4];
    int idx[4];
    for (int k = 0; k < cnt; ++k) {
        if (pts[k] != pts[k]) return;
        double a = fabs(ea_vessel - pts[k]);
        if (ea_vessel > pts[k]) a = 2 * PI_D - a;
        if (dir < 0) a = PI_D * 2 - a;
        ang[k] = a;
        idx[k] = k;
    }
    for (int i = 1; i < cnt; ++i) {     // stable insertion argsort
        const int t = idx[i];
        int j = i - 1;
        while (j >= 0 && ang[idx[j]] > ang[t]) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = t;
    }
    out2[0] = pts[idx[0]];
    out2[1] = pts[idx[cnt - 1]];
}

__device__ __forceinline__ double advance_ea(double ecc, double ma, double ea, double n, double dr)
{   // :411-418 / :432-439
    if (ecc < 1) return pe_solve_kepler_ell(ecc, ma + n * dr, 1e-10);
    return pe_newton_hyp(ecc, ma + n * dr * -1, ea);
}

__global__ void points_vessel_kernel(int64_t nsel, const int64_t *__restrict__ sel,
                                     const double *__restrict__ vessel12,
                                     const int64_t *__restrict__ v_ref,
                                     const double *__restrict__ body12, int64_t entered_coi,
                                     int64_t left_coi, int only_enter_coi,
                                     double *__restrict__ body_impact,
                                     double *__restrict__ body_enter_atm,
                                     double *__restrict__ coi_leave, double *__restrict__ ea_eff)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsel) return;
    const int64_t v = sel[s];
    const double *V = vessel12 + 12 * v;
    const double a = V[0], b = V[1], f = V[2], ecc = V[3], ma = V[4], ea = V[5], n = V[7], dr = V[8],
                 pe_d = V[9], ap_d = V[10];
    const bool ell = ecc < 1;
    const double *R = body12 + 12 * v_ref[v];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double pts[4];
    int cnt;
    if (!only_enter_coi) {
        cnt = (ell && pe_d > R[10]) ? 0 : intersect_circle(a, b, ecc, f, 0.0, R[10], pts);
        next_point(ea, pts, cnt, dr, body_impact + 2 * v);
        cnt = (ell && pe_d > R[10] + R[11]) ? 0
                  : intersect_circle(a, b, ecc, f, 0.0, R[10] + R[11], pts);
        next_point(ea, pts, cnt, dr, body_enter_atm + 2 * v);
        if (ell && ap_d < R[9]) {
            coi_leave[2 * v] = coi_leave[2 * v + 1] = nan;
        } else {
            cnt = intersect_circle(a, b, ecc, f, 0.0, R[9], pts);
            const double e2 = entered_coi == v ? advance_ea(ecc, ma, ea, n, dr) : ea;
            next_point(e2, pts, cnt, dr, coi_leave + 2 * v);
        }
    }
    ea_eff[s] = left_coi == v ? advance_ea(ecc, ma, ea, n, dr) : ea;
}

// one (vessel, candidate body) pair: the arguments of predict_enter_coi (:442-466)
__global__ void points_pair_kernel(int64_t npairs, const int64_t *__restrict__ pair_s,
                                   const int64_t *__restrict__ pair_body,
                                   const int64_t *__restrict__ sel,
                                   const double *__restrict__ vessel12,
                                   const int64_t *__restrict__ v_ref,
                                   const double *__restrict__ body12, int64_t left_coi,
                                   const double *__restrict__ coi_leave,
                                   const double *__restrict__ ea_eff, double *__restrict__ vd10,
                                   double *__restrict__ bd10, unsigned char *__restrict__ valid)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int64_t s = pair_s[p], v = sel[s], body = pair_body[p];
    const double *V = vessel12 + 12 * v, *B = body12 + 12 * body;
    const double a = V[0], b = V[1], f = V[2], ecc = V[3], ma = V[4], pea = V[6], n = V[7], dr = V[8];
    const double ea = ea_eff[s];
    const bool ell = ecc < 1;
    double b_ma = B[4];
    if (left_coi == v) {
        if (ell) b_ma += B[7] * B[8];
        else b_ma += B[7] * B[8] * -1;
    }
    double period;
    bool ok = true;
    if (ell) {
        period = V[11];
    } else {
        double ea_leave;
        if (v_ref[v] == 0) {
            double pts[4], lp[2];
            const int cnt = intersect_circle(a, b, ecc, f, 0.0, B[0] + B[9], pts);
            next_point(ea, pts, cnt, dr, lp);
            ea_leave = lp[0];
            ok = !(ea_leave != ea_leave);      // `continue`: the row is dropped
        } else {
            ea_leave = coi_leave[2 * v];
        }
        const double ma_leave = (ecc * sinh(ea_leave) - ea_leave) * -1;
        const double tt = -dr * (ma_leave - ma) / n;        // orbit_time_to, phys_shared.py:58-64
        period = (0.0 > tt) ? 0.0 : tt;
    }
    double *vd = vd10 + 10 * p, *bd = bd10 + 10 * p;
    vd[0] = a; vd[1] = b; vd[2] = f; vd[3] = ecc; vd[4] = ma; vd[5] = ea; vd[6] = pea;
    vd[7] = period; vd[8] = n; vd[9] = dr;
    bd[0] = B[0]; bd[1] = B[1]; bd[2] = B[2]; bd[3] = B[3]; bd[4] = b_ma; bd[5] = B[5];
    bd[6] = B[6]; bd[7] = B[7]; bd[8] = B[8]; bd[9] = B[9];
    valid[p] = ok ? 1 : 0;
}

// sort_intersect_indices(...)[0] over a vessel's rows + the np.append of :468-474
__global__ void points_select_kernel(int64_t nsel, const int64_t *__restrict__ sel,
                                     const int64_t *__restrict__ offsets,
                                     const int64_t *__restrict__ pair_body,
                                     const unsigned char *__restrict__ valid,
                                     const double *__restrict__ out5,
                                     const double *__restrict__ ea_eff,
                                     const double *__restrict__ vessel12,
                                     double *__restrict__ coi_enter)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsel) return;
    const int64_t v = sel[s], p0 = offsets[s], p1 = offsets[s + 1];
    const double ea = ea_eff[s], dr = vessel12[12 * v + 8];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    int64_t best = -1, best_p = -1, rows = 0;
    double best_ang = 0;
    for (int64_t p = p0; p < p1; ++p) {
        if (!valid[p]) continue;
        const double x = out5[5 * p];
        if (x == x) {
            double ang = fabs(ea - x);
            if (ea > x) ang = 2 * PI_D - ang;
            if (dr < 0) ang = PI_D * 2 - ang;
            if (best < 0 || ang < best_ang) {
                best = rows;
                best_p = p;
                best_ang = ang;
            }
        }
        ++rows;
    }
    double *ce = coi_enter + 6 * v;
    for (int k = 0; k < 6; ++k) ce[k] = nan;
    if (best >= 0) {
        // check_bodies[first_enter]: the index counts the rows that were kept, but is applied to
        // the full candidate list (reference behaviour, :470-472)
        ce[0] = (double)pair_body[p0 + best];
        for (int k = 0; k < 5; ++k) ce[1 + k] = out5[5 * best_p + k];
    }
}

}  // namespace

// pe_exact_trig of this file follows the context's mode (vsb_set_march_mode bit 1)
static int set_exact_mode(vsb_ctx *c)
{
    static int last_of[64];                     // per device: 0 unknown, 1 + mode otherwise
    const int exact = (c->march_sequential >> 1) & 1;
    int &last = last_of[c->device & 63];
    if (1 + exact != last) {
        VSB_CUDA(cudaMemcpyToSymbolAsync(pe_exact_trig, &exact, sizeof(int), 0, cudaMemcpyHostToDevice,
                                         c->stream));
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        last = 1 + exact;
    }
    return VSB_OK;
}

int vsb_launch_quartic(vsb_ctx *c, int64_t n, const double *d_coef, double *d_roots)
{
    VSB_TRY(set_exact_mode(c));
    if (n <= 0) return VSB_OK;
    quartic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, d_coef, d_roots);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_intersect(vsb_ctx *c, int64_t n, const double *d_in6, double *d_ea, int64_t *d_cnt)
{
    VSB_TRY(set_exact_mode(c));
    if (n <= 0) return VSB_OK;
    intersect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, d_in6, d_ea, d_cnt);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// d_work: mteam header | mtask[TEAM_TASK_CAP] | best_hit[n] (u64) | parts_done[n] (u32)
size_t vsb_march_work_bytes(int64_t n)
{
    return sizeof(mteam) + sizeof(mtask) * (size_t)TEAM_TASK_CAP + 8 * (size_t)n +
           4 * (size_t)((n + 1) / 2 * 2) + 64;
}

int vsb_launch_predict_enter_coi(vsb_ctx *c, int64_t n, const double *d_vd, const double *d_bd,
                                 const unsigned char *d_valid, double tol, int steps_limit,
                                 double *d_out5, void *d_work, int sequential)
{
    if (n <= 0) return VSB_OK;
    // bit 0: the serial loop, bit 1: exact trigonometry (vsb_set_march_mode)
    sequential &= 1;
    VSB_TRY(set_exact_mode(c));
    mteam *team = (mteam *)d_work;
    VSB_CUDA(cudaMemsetAsync(team, 0, sizeof(mteam), c->stream));
    if (sequential) {
        predict_enter_coi_seq_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(
            n, d_vd, d_bd, d_valid, tol, steps_limit, d_out5);
        c->launches += 1;
        VSB_CUDA(cudaGetLastError());
        return VSB_OK;
    }
    mtask *tasks = (mtask *)(team + 1);
    unsigned long long *best = (unsigned long long *)(tasks + TEAM_TASK_CAP);
    unsigned int *done = (unsigned int *)(best + n);
    const int64_t cap = (int64_t)c->sm_count * 2;
    enter_coi_march_kernel<<<(unsigned)(n < cap ? n : cap), MARCH_T, 0, c->stream>>>(
        n, d_vd, d_bd, d_valid, tol, steps_limit, d_out5, team, tasks, best, done);
    enter_coi_team_kernel<<<(unsigned)cap, MARCH_T, 0, c->stream>>>(
        d_vd, d_bd, tol, steps_limit, d_out5, team, tasks, best, done);
    c->launches += 2;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// Host-side evaluation of the closed-form running sum (the same march_walk the kernel runs):
// xs[i] = x_{ks[i]} of x_k = fl(x_{k-1} + c).  CPU tests pin it against the plain loop.
int vsb_march_walk_host(double x0, double c, int64_t kmax, int bounded, double bound, int64_t nq,
                        const int64_t *ks, double *xs, int64_t *K_out, int64_t *nseg_out)
{
    std::vector<mseg> seg(MSEG_CAP);
    int ns = 0, fail = 0;
    long long K = 0;
    march_walk(x0, c, kmax, bounded != 0, bound, seg.data(), &ns, &K, &fail);
    if (K_out) *K_out = K;
    if (nseg_out) *nseg_out = ns;
    if (fail) return fail;
    for (int64_t i = 0; i < nq; ++i) {
        if (ks[i] == 0) {
            xs[i] = x0;
            continue;
        }
        if (ks[i] < 0 || ks[i] > K) return -1;
        int idx = march_find(seg.data(), ns, ks[i]);
        xs[i] = march_at(seg.data(), idx, ks[i]);
    }
    return 0;
}

int vsb_launch_points_vessel(vsb_ctx *c, int64_t nsel, const int64_t *d_sel, const double *d_vessel12,
                             const int64_t *d_vref, const double *d_body12, int64_t entered_coi,
                             int64_t left_coi, int only_enter_coi, double *d_impact, double *d_atm,
                             double *d_leave, double *d_ea_eff)
{
    VSB_TRY(set_exact_mode(c));
    if (nsel <= 0) return VSB_OK;
    points_vessel_kernel<<<(unsigned)((nsel + 127) / 128), 128, 0, c->stream>>>(
        nsel, d_sel, d_vessel12, d_vref, d_body12, entered_coi, left_coi, only_enter_coi, d_impact,
        d_atm, d_leave, d_ea_eff);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_points_pairs(vsb_ctx *c, int64_t npairs, const int64_t *d_pair_s,
                            const int64_t *d_pair_body, const int64_t *d_sel,
                            const double *d_vessel12, const int64_t *d_vref, const double *d_body12,
                            int64_t left_coi, const double *d_leave, const double *d_ea_eff,
                            double *d_vd10, double *d_bd10, unsigned char *d_valid)
{
    VSB_TRY(set_exact_mode(c));
    if (npairs <= 0) return VSB_OK;
    points_pair_kernel<<<(unsigned)((npairs + 127) / 128), 128, 0, c->stream>>>(
        npairs, d_pair_s, d_pair_body, d_sel, d_vessel12, d_vref, d_body12, left_coi, d_leave,
        d_ea_eff, d_vd10, d_bd10, d_valid);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_points_select(vsb_ctx *c, int64_t nsel, const int64_t *d_sel, const int64_t *d_offsets,
                             const int64_t *d_pair_body, const unsigned char *d_valid,
                             const double *d_out5, const double *d_ea_eff, const double *d_vessel12,
                             double *d_enter)
{
    VSB_TRY(set_exact_mode(c));
    if (nsel <= 0) return VSB_OK;
    points_select_kernel<<<(unsigned)((nsel + 127) / 128), 128, 0, c->stream>>>(
        nsel, d_sel, d_offsets, d_pair_body, d_valid, d_out5, d_ea_eff, d_vessel12, d_enter);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
