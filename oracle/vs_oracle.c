/*
 * vs_oracle.c -- CPU restatement of VolatileSpace's physics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in volatilespace_b200/csrc.  It may be imported / linked / executed only by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, and there only as the checker or as the timed CPU
 * baseline -- never by the product path (volatilespace_b200/ has no CPU
 * fallback and fails loudly without its CUDA library).
 *
 * Every function cites the reference lines (relative to the reference repo,
 * mzivic7/VolatileSpace v0.5.2) whose arithmetic it restates.  The reference
 * runs these as numba @njit (fastmath=False: no FMA contraction, libm
 * transcendentals) or as plain numpy; this file is compiled with
 * -ffp-contract=off so every a*b+c is two IEEE roundings exactly as there.
 *
 * Pinning: the reference ships no tests or golden vectors.
 * The oracle is pinned against outputs of the reference itself, generated in
 * the dev container by oracle/gen_golden.py (imports /root/reference live)
 * and committed under tests/golden/; tests/test_oracle_golden.py checks every
 * function below against them.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>

/* "Exact trigonometry" experiment: sin / cos of the COI-entry march
 * evaluated in double-double arithmetic and rounded to nearest -- the SAME routine the device path
 * uses in its exact mode (generated header, plain C) -- instead of libm's.  Off by default: the
 * oracle then uses the reference's own libm and matches the live fixtures bit for bit. */
#include "../volatilespace_b200/csrc/vsb_ddtrig.h"
static int vso_exact_trig = 0;
void vso_set_exact_trig(int on) { vso_exact_trig = on; }
static double t_sin(double x) { return vso_exact_trig ? vsb_dd_sin(x) : sin(x); }
static double t_cos(double x) { return vso_exact_trig ? vsb_dd_cos(x) : cos(x); }
static double t_pow(double x, double y) { return vso_exact_trig ? vsb_dd_pow(x, y) : pow(x, y); }
static double t_acos(double x) { return vso_exact_trig ? vsb_dd_acos(x) : acos(x); }
static double t_atan2(double y, double x) { return vso_exact_trig ? vsb_dd_atan2(y, x) : atan2(y, x); }
#endif

#define VSO_PI 3.141592653589793 /* math.pi == np.pi */

int vso_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void vso_set_num_threads(int t)
{
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* Python float % for a positive divisor (numba lowers `%` on floats this way). */
static double py_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

/* ------------------------------------------------------------------ */
/* Part A -- body-sim step                                            */
/* ------------------------------------------------------------------ */

/* A1: find_parent_one, phys_editor.py:36-48, driven by find_parents :350-354.
 * order = argsort(mass)[::-1] is computed by the caller with numpy (mass-tie
 * order is whatever numpy's introsort gives). */
void vso_find_parents(int64_t n, const int64_t *order, const double *pos,
                      const double *coi, int64_t *parents)
{
    int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t k = 0; k < n; ++k) rank[order[k]] = k;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        int64_t parent = 0;
        if (i != 0) {
            for (int64_t l = 0; l < rank[i]; ++l) {
                int64_t p = order[l];
                double rx = pos[2 * i] - pos[2 * p];
                double ry = pos[2 * i + 1] - pos[2 * p + 1];
                if (rx * rx + ry * ry < coi[p] * coi[p]) parent = p;
            }
        }
        parents[i] = parent;
    }
    free(rank);
}

/* A2: gravity_one, phys_editor.py:51-60.  rel_pos = pos[parent] - pos[body]. */
void vso_gravity_one(int64_t body, int64_t parent, const double *mass, double rx,
                     double ry, double gc, double *out)
{
    if (body) {
        double distance = sqrt(rx * rx + ry * ry);
        double force = gc * mass[body] * mass[parent] / (distance * distance);
        double angle = atan2(ry, rx);
        double acc = force / mass[body];
        out[0] = acc * cos(angle);
        out[1] = acc * sin(angle);
    } else {
        out[0] = 0.0;
        out[1] = 0.0;
    }
}

/* A3: Physics.gravity, phys_editor.py:357-378 (one tick of the parent/COI
 * "simplified n-body" model, in place).  parents holds last tick's parents on
 * entry and this tick's on exit. */
void vso_gravity_ref(int64_t n, const int64_t *order, const double *mass,
                     const double *coi, double gc, double *pos, double *vel,
                     double *rel_vel, int64_t *parents)
{
    int64_t *parents_old = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    memcpy(parents_old, parents, sizeof(int64_t) * (size_t)n);
    vso_find_parents(n, order, pos, coi, parents);
    for (int64_t i = 0; i < n; ++i) {
        int64_t p = parents[i];
        double a[2];
        vso_gravity_one(i, p, mass, pos[2 * p] - pos[2 * i],
                        pos[2 * p + 1] - pos[2 * i + 1], gc, a);
        rel_vel[2 * i] += a[0];
        rel_vel[2 * i + 1] += a[1];
    }
    /* leaving / entering a COI, sequential in index order (:366-373) */
    for (int64_t b = 0; b < n; ++b) {
        if (parents[b] != parents_old[b] && b != 0) {
            int64_t parent = parents[b], old = parents_old[b];
            if (mass[parent] > mass[old]) {
                rel_vel[2 * b] += rel_vel[2 * old];
                rel_vel[2 * b + 1] += rel_vel[2 * old + 1];
            }
            if (mass[parent] < mass[old]) {
                rel_vel[2 * b] -= rel_vel[2 * parent];
                rel_vel[2 * b + 1] -= rel_vel[2 * parent + 1];
            }
        }
    }
    for (int64_t k = 1; k < n; ++k) { /* descending mass, root skipped (:375-377) */
        int64_t b = order[k], p = parents[b];
        vel[2 * b] = rel_vel[2 * b] + vel[2 * p];
        vel[2 * b + 1] = rel_vel[2 * b + 1] + vel[2 * p + 1];
    }
    for (int64_t i = 0; i < 2 * n; ++i) pos[i] += vel[i];
    free(parents_old);
}

/* A4: the documented all-pairs model, documentation/orbital-physics/
 * 01-complex-orbit-model.txt:4-12, evaluated with the reference's own pair
 * formula gravity_one (phys_editor.py:51-60; its result depends only on
 * mass[parent], rel_pos and gc when body != 0).  acc rows [row0,row1) only;
 * source order j ascending, j != i. */
void vso_allpairs_accel(int64_t n, const double *mass, const double *pos, double gc,
                        int64_t row0, int64_t row1, double *acc)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = row0; i < row1; ++i) {
        double ax = 0.0, ay = 0.0;
        double xi = pos[2 * i], yi = pos[2 * i + 1], mi = mass[i];
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            double rx = pos[2 * j] - xi, ry = pos[2 * j + 1] - yi;
            double distance = sqrt(rx * rx + ry * ry);
            double force = gc * mi * mass[j] / (distance * distance);
            double angle = atan2(ry, rx);
            double a = force / mi;
            ax += a * cos(angle);
            ay += a * sin(angle);
        }
        acc[2 * (i - row0)] = ax;
        acc[2 * (i - row0) + 1] = ay;
    }
}

/* A4 integrator: v = v + av ; P = P + v  (01-complex-orbit-model.txt:11-12). */
void vso_allpairs_step(int64_t n, const double *mass, double *pos, double *vel,
                       double gc, int64_t nsteps)
{
    double *acc = (double *)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
    for (int64_t s = 0; s < nsteps; ++s) {
        vso_allpairs_accel(n, mass, pos, gc, 0, n, acc);
        for (int64_t i = 0; i < 2 * n; ++i) {
            vel[i] += acc[i];
            pos[i] += vel[i];
        }
    }
    free(acc);
}

/* A6: kepler_basic_one, phys_editor.py:63-83, over all bodies (:381-384). */
void vso_kepler_basic(int64_t n, const int64_t *parents, const double *mass,
                      const double *pos, const double *vel, double gc, double coi_coef,
                      double *focus, double *ecc_v, double *semi_major,
                      double *semi_minor, double *periapsis_arg, double *coi)
{
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < n; ++b) {
        if (!b) {
            focus[0] = 0.0; ecc_v[0] = 0.0; ecc_v[1] = 0.0; semi_major[0] = 0.0;
            semi_minor[0] = 0.0; periapsis_arg[0] = 0.0; coi[0] = 0.0;
            continue;
        }
        int64_t p = parents[b];
        double rpx = pos[2 * b] - pos[2 * p], rpy = pos[2 * b + 1] - pos[2 * p + 1];
        double rvx = vel[2 * b] - vel[2 * p], rvy = vel[2 * b + 1] - vel[2 * p + 1];
        double u = gc * mass[p];
        double magp = sqrt(rpx * rpx + rpy * rpy);
        double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / magp));
        double momentum = rpx * rvy - rpy * rvx;
        double sx = rvy, sy = -rvx; /* swap axes and -y (:73) */
        double ex = (sx * momentum / u) - rpx / magp;
        double ey = (sy * momentum / u) - rpy / magp;
        double ecc = sqrt(ex * ex + ey * ey);
        double pea = py_mod((3 * VSO_PI / 2) + atan2(-ex, ey), 2 * VSO_PI);
        double f = a * ecc;
        double bm = sqrt(fabs(f * f - a * a));
        double c = 0.0;
        if (a > 0) c = a * pow(mass[b] / mass[p], coi_coef);
        focus[b] = f; ecc_v[2 * b] = ex; ecc_v[2 * b + 1] = ey; semi_major[b] = a;
        semi_minor[b] = bm; periapsis_arg[b] = pea; coi[b] = c;
    }
}

/* A7: simplified_orbit_coi, phys_editor.py:325-346.  order as for A1; returns
 * largest = argmax(mass) (first maximum). */
int64_t vso_simplified_orbit_coi(int64_t n, const int64_t *order, const double *mass,
                                 const double *pos, const double *vel, double gc,
                                 double coi_coef, double *coi)
{
    int64_t largest = 0;
    for (int64_t i = 1; i < n; ++i)
        if (mass[i] > mass[largest]) largest = i;
    for (int64_t i = 0; i < n; ++i) coi[i] = 0.0;
    for (int64_t num = 0; num + 1 < n; ++num) {
        int64_t body = order[num + 1];
        double body_mass = mass[body];
        int64_t parent = largest;
        for (int64_t l = 0; l <= num; ++l) {
            int64_t pot = order[l];
            double dx = pos[2 * body] - pos[2 * pot], dy = pos[2 * body + 1] - pos[2 * pot + 1];
            if (dx * dx + dy * dy < coi[pot] * coi[pot]) parent = pot;
        }
        double rpx = pos[2 * parent] - pos[2 * body], rpy = pos[2 * parent + 1] - pos[2 * body + 1];
        double rvx = vel[2 * parent] - vel[2 * body], rvy = vel[2 * parent + 1] - vel[2 * body + 1];
        double u = gc * mass[parent];
        double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / sqrt(rpx * rpx + rpy * rpy)));
        if (a > 0)
            coi[body] = a * pow(body_mass / mass[parent], coi_coef);
        else
            coi[body] = 0.0;
    }
    return largest;
}

/* ------------------------------------------------------------------ */
/* Part B -- collisions                                               */
/* ------------------------------------------------------------------ */

/* B1: check_collision_one (phys_editor.py:86-97) for every i, first non-None
 * result (:547-551) == lexicographically smallest colliding (i<j).  Writes
 * (body_del, body_add), or (-1,-1) for (None, None). */
void vso_check_collision(int64_t n, const double *mass, const double *pos,
                         const double *rad, int64_t *del, int64_t *add)
{
    int64_t best_i = -1, best_j = -1;
#pragma omp parallel
    {
        int64_t li = -1, lj = -1;
#pragma omp for schedule(dynamic, 64) nowait
        for (int64_t i = 0; i < n - 1; ++i) {
            if (li >= 0 && li < i) continue; /* a smaller i already hit */
            for (int64_t j = i + 1; j < n; ++j) {
                double rx = pos[2 * i] - pos[2 * j], ry = pos[2 * i + 1] - pos[2 * j + 1];
                double distance = sqrt(rx * rx + ry * ry);
                if (distance <= rad[i] + rad[j]) {
                    if (li < 0 || i < li) { li = i; lj = j; }
                    break;
                }
            }
        }
#pragma omp critical
        {
            if (li >= 0 && (best_i < 0 || li < best_i)) { best_i = li; best_j = lj; }
        }
    }
    if (best_i < 0) { *del = -1; *add = -1; return; }
    if (mass[best_i] <= mass[best_j]) { *del = best_i; *add = best_j; }
    else { *del = best_j; *add = best_i; }
}

/* Number of predicate tests the reference algorithm performs for a tick with
 * no hit (N(N-1)/2); helper for the CPU-baseline bookkeeping. */
int64_t vso_collision_tests(int64_t n) { return n * (n - 1) / 2; }

/* ------------------------------------------------------------------ */
/* Part C -- Keplerian propagation and orbit curves                   */
/* ------------------------------------------------------------------ */

/* C2: solve_kepler_ell, enhanced_kepler_solver.py:18-69 (ENRKE). */
double vso_solve_kepler_ell(double ecc, double ma, double tol)
{
    double mr = ma, flip;
    if (mr > VSO_PI) { mr = 2 * VSO_PI - mr; flip = 1; }
    else flip = -1;

    if (ecc > 0.99 && mr < 0.0045) {
        double al = tol / 1e7, be = tol / 0.3, fp = 2.7 * mr, fpp = 0.301, f = 0.154;
        for (int i = 0; i < 1000; ++i) { /* midpoint update is outside the loop (F8b) */
            if ((fpp - fp) <= (al + be * f)) break;
            if ((f - ecc * t_sin(f) - mr) > 0) fpp = f;
            else fp = f;
        }
        f = 0.5 * (fp + fpp);
        return ma + flip * (mr - f);
    }

    double tol2s = 2 * tol / (ecc + 2.2e-16);
    double eapp = mr + 0.999999 * mr * (VSO_PI - mr) /
                           (2. * mr + ecc - VSO_PI + 2.4674011002723395 / (ecc + 2.2e-16));
    double fpp = ecc * t_sin(eapp);
    double fppp = ecc * t_cos(eapp);
    double fp = 1 - fppp;
    double f = eapp - fpp - mr;
    double delta = -f / fp;
    double fp3 = fp * fp * fp;
    double ffpfpp = f * fp * fpp;
    double f2fppp = f * f * fppp;
    delta = delta * (fp3 - 0.5 * ffpfpp + f2fppp / 3) / (fp3 - ffpfpp + 0.5 * f2fppp);
    for (int i = 0; i < 30; ++i) {
        if (delta * delta < fp * tol2s) break;
        eapp = eapp + delta;
        fp = 1 - ecc * t_cos(eapp);
        delta = (mr - eapp + ecc * t_sin(eapp)) / fp;
    }
    eapp = eapp + delta;
    return ma + flip * (mr - eapp);
}

/* C3: newton_root_kepler_hyp, phys_shared.py:114-122 (negated convention). */
double vso_newton_root_kepler_hyp(double ecc, double ma, double ea_guess)
{
    double ea = ea_guess;
    for (int i = 0; i < 50; ++i) {
        double delta_x = (ea - ecc * sinh(ea) - ma) / (1.0 - ecc * cosh(ea));
        ea -= delta_x;
        if (fabs(delta_x) < 1e-10) return ea;
    }
    return ea;
}

/* newton_root_kepler_ell, phys_shared.py:103-111 (plain twin, off the hot path). */
double vso_newton_root_kepler_ell(double ecc, double ma, double ea_guess)
{
    double ea = ea_guess;
    for (int i = 0; i < 50; ++i) {
        double delta_x = (ea - ecc * sin(ea) - ma) / (1.0 - ecc * cos(ea));
        ea -= delta_x;
        if (fabs(delta_x) < 1e-10) return ea;
    }
    return ea;
}

void vso_solve_kepler_ell_batch(int64_t n, const double *ecc, const double *ma, double tol,
                                double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = vso_solve_kepler_ell(ecc[i], ma[i], tol);
}

void vso_newton_root_kepler_hyp_batch(int64_t n, const double *ecc, const double *ma,
                                      const double *guess, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        out[i] = vso_newton_root_kepler_hyp(ecc[i], ma[i], guess[i]);
}

/* C8: calc_orb_one, phys_body.py:27-60 (with_coi=1: bodies, root -> zeros) and
 * phys_vessel.py:82-110 (with_coi=0: vessels, no root special case).
 * out8 = b, f, coi, pe_d, ap_d, period, n, u. */
void vso_calc_orb_one(int with_coi, int64_t body, int64_t ref, const double *mass,
                      double body_mass, double gc, double coi_coef, double a, double ecc,
                      double *out8)
{
    if (with_coi && !body) {
        for (int k = 0; k < 8; ++k) out8[k] = 0.0;
        return;
    }
    double u = gc * mass[ref];
    double f = a * ecc;
    double b = sqrt(fabs(f * f - a * a));
    double coi = 0.0, pe_d, ap_d, period, n;
    if (with_coi && a > 0) coi = a * pow(body_mass / mass[ref], coi_coef);
    if (ecc != 0) {
        pe_d = a * (1 - ecc);
        if (ecc < 1) {
            period = 2 * VSO_PI * sqrt(pow(a, 3.0) / u);   /* a**3: libm pow, as CPython */
            ap_d = a * (1 + ecc);
            n = 2 * VSO_PI / period;
        } else {
            if (ecc > 1) pe_d = a * (1 - ecc);
            else pe_d = a;
            period = 0;
            ap_d = 0;
            n = sqrt(u / pow(fabs(a), 3.0));
        }
    } else {
        period = (2 * VSO_PI * sqrt(pow(a, 3.0) / u)) / 10;
        pe_d = a * (1 - ecc);
        ap_d = 0;
        n = 2 * VSO_PI / period;
    }
    out8[0] = b; out8[1] = f; out8[2] = coi; out8[3] = pe_d; out8[4] = ap_d;
    out8[5] = period; out8[6] = n; out8[7] = u;
}

/* C1: phys_body.Physics.move, phys_body.py:323-346.  order = argsort(coi)[::-1]
 * from the caller (numpy).  In place on ma, ea, pos. */
void vso_body_move(int64_t n, double warp, const int64_t *order, const int64_t *ref,
                   const double *a, const double *b, const double *f, const double *ecc,
                   const double *pea, const double *nmot, const double *dr, double *ma,
                   double *ea, double *pos)
{
    for (int64_t i = 0; i < n; ++i) {
        double m = ma[i] + dr[i] * nmot[i] * warp;
        if (m > 2 * VSO_PI) m = m - 2 * VSO_PI;
        if (m < 0) m = m + 2 * VSO_PI;
        ma[i] = m;
    }
    for (int64_t k = 0; k < n; ++k) {
        int64_t body = order[k];
        double e, prx, pry;
        if (ecc[body] < 1) {
            e = vso_solve_kepler_ell(ecc[body], ma[body], 1e-10);
            prx = a[body] * cos(e) - f[body];
            pry = b[body] * sin(e);
        } else {
            e = vso_newton_root_kepler_hyp(ecc[body], ma[body], ea[body]);
            prx = a[body] * cosh(e) - f[body];
            pry = b[body] * sinh(e);
        }
        double ang = pea[body] - VSO_PI;
        double x = prx * cos(ang) - pry * sin(ang);
        double y = prx * sin(ang) + pry * cos(ang);
        int64_t r = ref[body];
        double rx = pos[2 * r], ry = pos[2 * r + 1];
        pos[2 * body] = rx + x;
        pos[2 * body + 1] = ry + y;
        ea[body] = e;
    }
}

/* C4: phys_vessel.Physics.move (phys_vessel.py:477-494) with ea2coord (:247-255)
 * and orb2xy (phys_shared.py:187-199).  In place on ma, ea; writes pos. */
void vso_vessel_move(int64_t n, double warp, const int64_t *ref, const double *body_pos,
                     const double *a, const double *b, const double *f, const double *ecc,
                     const double *pea, const double *nmot, const double *dr, double *ma,
                     double *ea, double *pos)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double hyp = ecc[i] > 1 ? -1.0 : 1.0;
        double m = ma[i] + dr[i] * nmot[i] * warp * hyp;
        if (ecc[i] < 1 && m > 2 * VSO_PI) m = m - 2 * VSO_PI;
        if (ecc[i] < 1 && m < 0) m = m + 2 * VSO_PI;
        ma[i] = m;
        double e;
        if (ecc[i] < 1) e = vso_solve_kepler_ell(ecc[i], m, 1e-10);
        else e = vso_newton_root_kepler_hyp(ecc[i], m, ea[i]);
        double x = 0.0, y = 0.0;
        if (e != 0.0) {
            double xn, yn;
            if (ecc[i] < 1) { xn = a[i] * cos(e) - f[i]; yn = b[i] * sin(e); }
            else { xn = a[i] * cosh(e) - f[i]; yn = b[i] * sinh(e); }
            double ang = pea[i] - VSO_PI;
            int64_t r = ref[i];
            x = (xn * cos(ang) - yn * sin(ang)) + body_pos[2 * r];
            y = (xn * sin(ang) + yn * cos(ang)) + body_pos[2 * r + 1];
        }
        pos[2 * i] = x;
        pos[2 * i + 1] = y;
        ea[i] = e;
    }
}

/* C5: curve_points, phys_shared.py:202-216.  out is (P,2). */
void vso_curve_points(double ecc, double a, double b, double pea, const double *t,
                      int64_t P, double *out)
{
    double sin_p = sin(pea), cos_p = cos(pea);
    for (int64_t k = 0; k < P; ++k) {
        double x, y;
        if (ecc < 1) { x = a * cos(t[k]); y = b * sin(t[k]); }
        else { x = -a * cosh(t[k]); y = b * sinh(t[k]); }
        out[2 * k] = x * cos_p - y * sin_p;
        out[2 * k + 1] = y * cos_p + x * sin_p;
    }
}

/* C5 over all objects: Physics.curve(i) for every i (phys_body.py:315-320,
 * phys_vessel.py:364-369).  curves is (N,P,2). */
void vso_curves_all(int64_t n, const double *ecc, const double *a, const double *b,
                    const double *pea, const double *t, int64_t P, double *curves)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        vso_curve_points(ecc[i], a[i], b[i], pea[i], t, P, curves + 2 * P * i);
}

/* C6: curve_move_to, phys_shared.py:219-222.  (N,P,2) -> (N,P,2). */
void vso_curve_move_to(int64_t n, int64_t P, const double *curves, const double *body_pos,
                       const int64_t *ref, const double *f, const double *pea, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double fx = f[i] * cos(pea[i]), fy = f[i] * sin(pea[i]);
        double bx = body_pos[2 * ref[i]], by = body_pos[2 * ref[i] + 1];
        const double *c = curves + 2 * P * i;
        double *o = out + 2 * P * i;
        for (int64_t k = 0; k < P; ++k) {
            o[2 * k] = c[2 * k] + fx + bx;
            o[2 * k + 1] = c[2 * k + 1] + fy + by;
        }
    }
}

/* C7: editor Physics.curve, phys_editor.py:519-544.  out is (2,N,P).
 * ell_t = linspace(-pi,pi,P), par_t = linspace(-pi-1,pi+1,P) (:155-157). */
void vso_editor_curve(int64_t n, int64_t P, const double *ecc_v, const double *semi_major,
                      const double *semi_minor, const double *focus,
                      const double *periapsis_arg, const int64_t *parents,
                      const double *pos, const double *ell_t, const double *par_t,
                      double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double ex = ecc_v[2 * i], ey = ecc_v[2 * i + 1];
        double ecc = sqrt(ex * ex + ey * ey);
        double cp = cos(periapsis_arg[i]), sp = sin(periapsis_arg[i]);
        double fx = focus[i] * cp, fy = focus[i] * sp;
        double px = pos[2 * parents[i]], py = pos[2 * parents[i] + 1];
        double a = semi_major[i], b = semi_minor[i];
        for (int64_t k = 0; k < P; ++k) {
            double x = 0.0, y = 0.0;
            if (ecc < 1) { x = a * cos(ell_t[k]); y = b * sin(ell_t[k]); }
            else if (ecc == 1) { x = a * (par_t[k] * par_t[k]) - a; y = 2 * a * par_t[k]; }
            else if (ecc > 1) { x = -a * cosh(ell_t[k]); y = b * sinh(ell_t[k]); }
            double xr = cp * x + (-sp) * y;
            double yr = sp * x + cp * y;
            out[(0 * n + i) * P + k] = xr + fx + px;
            out[(1 * n + i) * P + k] = yr + fy + py;
        }
    }
}

/* ------------------------------------------------------------------ */
/* N1 -- Newton <-> Kepler conversion at load/save        */
/* ------------------------------------------------------------------ */

/* convert.to_kepler, convert.py:193-261, including its index quirks: the first loop walks the
 * bodies in INDEX order, and after the parent search it uses the body's mass RANK `body` as an
 * index into pos / vel / mass / coi (:219-225).  order = argsort(mass)[::-1] from the caller. */
void vso_to_kepler(int64_t n, const int64_t *order, const double *mass, const double *pos,
                   const double *vel, double gc, double coi_coef, int64_t *ref_l, double *a_l,
                   double *ecc_l, double *pe_l, double *ma_l, double *dir_l)
{
    int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    double *coi = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int64_t k = 0; k < n; ++k) rank[order[k]] = k;
    for (int64_t s = 0; s < n; ++s) {
        int64_t body = rank[s], ref = 0;
        for (int64_t l = 0; l < body; ++l) {
            int64_t pp = order[l];
            double rx = pos[2 * s] - pos[2 * pp], ry = pos[2 * s + 1] - pos[2 * pp + 1];
            if (rx * rx + ry * ry < coi[pp] * coi[pp]) ref = pp;
        }
        ref_l[s] = ref;
        double rpx = pos[2 * ref] - pos[2 * body], rpy = pos[2 * ref + 1] - pos[2 * body + 1];
        double rvx = vel[2 * ref] - vel[2 * body], rvy = vel[2 * ref + 1] - vel[2 * body + 1];
        double u = gc * mass[ref];
        double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / sqrt(rpx * rpx + rpy * rpy)));
        coi[body] = a > 0 ? a * pow(mass[body] / mass[ref], coi_coef) : 0.0;
    }
    for (int64_t b = 0; b < n; ++b) {
        a_l[b] = ecc_l[b] = pe_l[b] = ma_l[b] = dir_l[b] = 0.0;
        if (!b) continue;
        int64_t ref = ref_l[b];
        double rpx = pos[2 * b] - pos[2 * ref], rpy = pos[2 * b + 1] - pos[2 * ref + 1];
        double rvx = vel[2 * b] - vel[2 * ref], rvy = vel[2 * b + 1] - vel[2 * ref + 1];
        double u = gc * mass[ref];
        double magp = sqrt(rpx * rpx + rpy * rpy);
        double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / magp));
        double momentum = rpx * rvy - rpy * rvx;
        double sx = rvy, sy = -rvx;
        double ex = (sx * momentum / u) - rpx / magp, ey = (sy * momentum / u) - rpy / magp;
        double ecc = sqrt(ex * ex + ey * ey);
        double pe = py_mod((3 * VSO_PI / 2) + atan2(-ex, ey), 2 * VSO_PI);
        double dir = copysign(1.0, momentum);
        double ta = py_mod(pe - (atan2(rpy, rpx) - VSO_PI), 2 * VSO_PI);
        if (dir == -1) ta = 2 * VSO_PI - ta;
        double ea, ma;
        if (ecc < 1) {
            ea = acos((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
            if (VSO_PI < ta && ta < 2 * VSO_PI) ea = 2 * VSO_PI - ea;
            ma = py_mod(ea - ecc * sin(ea), 2 * VSO_PI);
        } else {
            ea = acosh((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
            ma = ecc * sinh(ea) - ea;
        }
        a_l[b] = a; ecc_l[b] = ecc; pe_l[b] = pe; ma_l[b] = ma; dir_l[b] = dir;
    }
    free(rank);
    free(coi);
}

/* rot_ellipse_by_y, phys_shared.py:133-144 */
static double rot_ellipse_by_y(double x, double a, double b, double p)
{
    double sin_p = sin(p), cos_p = cos(p);
    double a2 = a * a, b2 = b * b, x2 = x * x, sin_p2 = sin_p * sin_p, cos_p2 = cos_p * cos_p;
    double c4 = cos_p2 * cos_p2, s4 = sin_p2 * sin_p2; /* cos_p**4, sin_p**4 */
    return (a * b * sqrt(-x2 * (c4 + s4) - 2 * x2 * cos_p2 * sin_p2 + b2 * sin_p2 + a2 * cos_p2) +
            a2 * x * sin_p * cos_p - b2 * x * sin_p * cos_p) / (a2 * cos_p2 + b2 * sin_p2);
}

/* impl_derivative_rot_ell, phys_shared.py:160-171 */
static double impl_derivative_rot_ell(double px, double py, double a, double b, double p)
{
    double sin_p = sin(p), cos_p = cos(p);
    double a2 = a * a, b2 = b * b, sc = sin_p * cos_p, sin_p2 = sin_p * sin_p, cos_p2 = cos_p * cos_p;
    return atan(-(b2 * px * cos_p2 + a2 * px * sin_p2 + b2 * py * sc - a2 * py * sc) /
                (a2 * py * cos_p2 + b2 * py * sin_p2 + b2 * px * sc - a2 * px * sc));
}

/* convert.to_newton, convert.py:132-190 with kepler_to_velocity :20-64 (ellipse branch; the
 * reference raises ValueError at :170 for ecc > 1, reported here as return value -1).  Its
 * loop variable `body` is the position inside bodies_sorted, i.e. it visits the bodies in
 * INDEX order, whatever the COI sort says (:158-160). */
int vso_to_newton(int64_t n, const double *mass, const double *a_l, const double *ecc_l,
                  const double *pe_l, const double *ma_l, const int64_t *ref_l,
                  const double *dir_l, double gc, double *pos, double *vel)
{
    for (int64_t i = 0; i < 2 * n; ++i) pos[i] = vel[i] = 0.0;
    for (int64_t body = 1; body < n; ++body) {
        double a = a_l[body], ecc = ecc_l[body], pe = pe_l[body], ma = ma_l[body], dr = dir_l[body];
        int64_t ref = ref_l[body];
        double u = gc * mass[ref];
        if (dr > 0) ma = -ma;
        if (ecc == 0) ecc = 0.00001;
        if (ecc == 1) ecc = 1.00001;
        if (1 - ecc * ecc < 0) return -1;
        double b = a * sqrt(1 - ecc * ecc);
        double f = sqrt(a * a - b * b);
        double ea = vso_solve_kepler_ell(ecc, ma, 1e-10);
        double rx = a * cos(ea) - f, ry = b * sin(ea);
        double ang = pe - VSO_PI;
        double px = rx * cos(ang) - ry * sin(ang), py = rx * sin(ang) + ry * cos(ang);
        /* kepler_to_velocity(rel_pos, a, ecc, pe_arg, u, dr), ellipse */
        double frx = f * cos(pe), fry = f * sin(pe);
        double qx = px - f * cos(pe), qy = py - f * sin(pe);
        double angle = impl_derivative_rot_ell(qx, qy, a, b, pe);
        double arg = a * a * cos(2 * pe) + a * a - b * b * cos(2 * pe) + b * b;
        if (arg < 0) return -2;
        double x_max = sqrt(arg) / sqrt(2.0) - 1e-06;
        double y_max = rot_ellipse_by_y(x_max, a, b, pe);
        double x1 = x_max + frx, y1 = y_max + fry, x2 = -x_max + frx, y2 = -y_max + fry;
        if (((x2 - x1) * (py - y1)) - ((y2 - y1) * (px - x1)) < 0) angle += VSO_PI;
        angle = py_mod(angle, 2 * VSO_PI);
        double distance = sqrt(px * px + py * py);
        double speed = dr * sqrt(u * ((2 / distance) - (1 / a)));
        pos[2 * body] = pos[2 * ref] + px;
        pos[2 * body + 1] = pos[2 * ref + 1] + py;
        vel[2 * body] = vel[2 * ref] + speed * cos(angle);
        vel[2 * body + 1] = vel[2 * ref + 1] + speed * sin(angle);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* N2 -- numeric kernels of the vessel COI-crossing machinery          */
/* ------------------------------------------------------------------ */

/* solve_cubic_one, quartic_solver.py:28-54 */
static double solve_cubic_one(double a, double b, double c)
{
    double q = b / 3 - a * a / 9;
    double r = (b * a - 3 * c) / 6 - a * a * a / 27;
    double rq = r * r + q * q * q;
    double z1;
    if (rq > 0) {
        double aa = t_pow(fabs(r) + sqrt(rq), 1.0 / 3);
        double t = r >= 0 ? aa - q / aa : q / aa - aa;
        z1 = t - a / 3;
    } else {
        double theta = 0;
        if (q < 0) theta = t_acos(r / t_pow(-q, 3.0 / 2));
        double fi = theta / 3;
        z1 = 2 * sqrt(-q) * t_cos(fi) - a / 3;
    }
    return z1;
}

/* complex square root of a REAL argument (every cmath.sqrt in solve_quartic gets one) */
static void csqrt_real(double x, double *re, double *im)
{
    if (x != x) { *re = x; *im = x; }
    else if (x >= 0) { *re = sqrt(x); *im = 0.0; }
    else { *re = 0.0; *im = sqrt(-x); }
}

/* solve_quartic, quartic_solver.py:57-93: roots8 = re,im of z1..z4 */
void vso_solve_quartic(double a, double b, double c, double d, double e, double *roots8)
{
    double a3 = b / a, a2 = c / a, a1 = d / a, a0 = e / a;
    double cc = a3 / 4;
    double cc2 = cc * cc;
    double b2 = a2 - 6 * cc2;
    double b1 = a1 - 2 * a2 * cc + 8 * (cc * cc2);
    double b0 = a0 - a1 * cc + a2 * cc2 - 3 * (cc2 * cc2);
    double y = solve_cubic_one(b2, b2 * b2 / 4 - b0, -(b1 * b1) / 8);
    if (0 > y) y = 0;
    double s = y * y + b2 * y + b2 * b2 / 4 - b0;
    double r;
    if (s > 0) r = b1 < 0 ? -sqrt(s) : sqrt(s);
    else r = NAN;
    double pr, pi_, dummy;
    csqrt_real(y / 2, &pr, &pi_);
    double p_re = pr - cc, p_im = pi_;
    double p1_re = -pr - cc, p1_im = -pi_;
    (void)dummy;
    double q = -y / 2 - b2 / 2;
    double s1r, s1i, s2r, s2i;
    csqrt_real(q - r, &s1r, &s1i);
    csqrt_real(q + r, &s2r, &s2i);
    roots8[0] = p_re + s1r;  roots8[1] = p_im + s1i;
    roots8[2] = p_re - s1r;  roots8[3] = p_im - s1i;
    roots8[4] = p1_re + s2r; roots8[5] = p1_im + s2i;
    roots8[6] = p1_re - s2r; roots8[7] = p1_im - s2i;
}

/* ell_hyp_intersect_circle, orbit_intersect.py:17-52.  Returns the number of values written
 * to ea4 (0 stands for the reference's [nan]). */
int vso_ell_hyp_intersect_circle(double a, double b, double ecc, double c_x, double c_y, double r,
                                 double *ea4)
{
    double roots[8], rr[4];
    int cnt = 0, any = 0;
    if (ecc < 1) {
        double a_0 = a * a - 2 * a * c_x + c_x * c_x + c_y * c_y - r * r;
        double a_1 = -4 * b * c_y;
        double a_2 = -2 * (a * a) + 4 * (b * b) + 2 * (c_x * c_x) + 2 * (c_y * c_y) - 2 * (r * r);
        double a_3 = -4 * b * c_y;
        double a_4 = a * a + 2 * a * c_x + c_x * c_x + c_y * c_y - r * r;
        vso_solve_quartic(a_4, a_3, a_2, a_1, a_0, roots);
        for (int k = 0; k < 4; ++k)
            if (roots[2 * k + 1] == 0.0) { rr[cnt++] = roots[2 * k]; }
        for (int k = 0; k < cnt; ++k) if (rr[k] != 0.0) any = 1;
        if (!any) return 0;
        for (int k = 0; k < cnt; ++k)
            ea4[k] = py_mod(t_atan2(2 * rr[k], 1.0 - rr[k] * rr[k]), 2 * VSO_PI);
        return cnt;
    }
    double a_0 = -(a * a) * (c_y * c_y + b * b) + b * b * ((c_x + r) * (c_x + r));
    double a_1 = -4 * (a * a) * r * c_y;
    double a_2 = -2 * (a * a * (c_y * c_y + b * b + 2 * (r * r)) + b * b * (r * r - c_x * c_x));
    double a_3 = -4 * (a * a) * r * c_y;
    double a_4 = -(a * a) * (c_y * c_y + b * b) + b * b * ((c_x - r) * (c_x - r));
    vso_solve_quartic(a_4, a_3, a_2, a_1, a_0, roots);
    for (int k = 0; k < 4; ++k)
        if (roots[2 * k + 1] == 0.0) { rr[cnt++] = roots[2 * k]; }
    for (int k = 0; k < cnt; ++k) if (rr[k] != 0.0) any = 1;
    if (!any) return 0;
    for (int k = 0; k < cnt; ++k) {
        double x = c_x + r * (1 - rr[k] * rr[k]) / (1 + rr[k] * rr[k]);
        double y = c_y + r * 2 * rr[k] / (1 + rr[k] * rr[k]);
        double ta = t_atan2(y, x - (a * ecc));
        double ea = acosh((ecc + t_cos(ta)) / (1 + (ecc * t_cos(ta))));
        ea4[k] = rr[k] < 0 ? -ea : ea;
    }
    return cnt;
}

static void oi_orb2xy(double a, double b, double f, double ecc, double pea, double ea, double *x,
                      double *y)
{ /* orbit_intersect.py:95-105 */
    double xn, yn;
    if (ecc < 1) { xn = a * t_cos(ea) - f; yn = b * t_sin(ea); }
    else { xn = a * cosh(ea) - f; yn = b * sinh(ea); }
    *x = xn * t_cos(pea - VSO_PI) - yn * t_sin(pea - VSO_PI);
    *y = xn * t_sin(pea - VSO_PI) + yn * t_cos(pea - VSO_PI);
}

static void oi_move(double *ma, double ecc, double dr, double n, double *prev_ea, double dt,
                    double *ea)
{ /* orbit_intersect.py:108-121 (the two wrap statements there are no-ops) */
    if (ecc < 1) {
        *ma += dr * n * dt;
        *ea = vso_solve_kepler_ell(ecc, *ma, 1e-10);
    } else {
        *ma += dr * n * -1 * dt;
        *ea = vso_newton_root_kepler_hyp(ecc, *ma, *prev_ea);
        *prev_ea = *ea;
    }
}

static double oi_wrap(double angle)
{
    if (angle >= 2 * VSO_PI) return angle - 2 * VSO_PI;
    if (angle < 0) return angle + 2 * VSO_PI;
    return angle;
}

/* predict_enter_coi, orbit_intersect.py:125-174.  vd = a,b,f,ecc,ma,ea,pea,period,n,dr;
 * bd = a,b,f,ecc,ma,ea,pea,n,dr,coi; out5 = ea, b_ea, ma, b_ma, t (NaN x5 when nothing). */
void vso_predict_enter_coi(const double *vd, const double *bd, double tol, int steps_limit,
                           double *out5)
{
    double a = vd[0], b = vd[1], f = vd[2], ecc = vd[3], ma = vd[4], ea = vd[5], pea = vd[6],
           period = vd[7], n = vd[8], dr = vd[9];
    double b_a = bd[0], b_b = bd[1], b_f = bd[2], b_ecc = bd[3], b_ma = bd[4], b_ea = bd[5],
           b_pea = bd[6], b_n = bd[7], b_dr = bd[8], b_coi = bd[9];
    double prev_ea = ea, b_prev_ea = b_ea;
    for (int k = 0; k < 5; ++k) out5[k] = NAN;
    if (period == 0 || b_coi == 0) return;
    double t = 0;
    double c1 = period / b_coi * 5, c2 = period / 2, c3 = period / 30;
    double m1 = c2 < c1 ? c2 : c1;              /* min(c1, c2) */
    double m2 = c3 > m1 ? c3 : m1;              /* max(m1, c3) */
    long long steps = (long long)m2;
    if (steps_limit > steps) steps = steps_limit;
    double dt = period / (double)steps;
    double t_end = period;
    while (t < t_end) {
        oi_move(&ma, ecc, dr, n, &prev_ea, dt, &ea);
        oi_move(&b_ma, b_ecc, b_dr, b_n, &b_prev_ea, dt, &b_ea);
        double rvx, rvy, rbx, rby;
        oi_orb2xy(a, b, f, ecc, pea, ea, &rvx, &rvy);
        oi_orb2xy(b_a, b_b, b_f, b_ecc, b_pea, b_ea, &rbx, &rby);
        double dx = rvx - rbx, dy = rvy - rby;
        double d = pow(dx * dx + dy * dy, 0.5);
        if (d < b_coi) {
            for (int it = 0; it < 30; ++it) {
                dt = fabs(dt) / 2 * (d < b_coi ? -1 : 1);
                t += dt;
                oi_move(&ma, ecc, dr, n, &prev_ea, dt, &ea);
                oi_move(&b_ma, b_ecc, b_dr, b_n, &b_prev_ea, dt, &b_ea);
                oi_orb2xy(a, b, f, ecc, pea, ea, &rvx, &rvy);
                oi_orb2xy(b_a, b_b, b_f, b_ecc, b_pea, b_ea, &rbx, &rby);
                dx = rvx - rbx; dy = rvy - rby;
                d = pow(dx * dx + dy * dy, 0.5);
                if (fabs(d - b_coi) < tol) break;
            }
            out5[0] = oi_wrap(ea); out5[1] = oi_wrap(b_ea); out5[2] = oi_wrap(ma);
            out5[3] = oi_wrap(b_ma); out5[4] = t;
            return;
        }
        t += dt;
    }
}

/* ---------------------------------------------------------------------------------------------
 * N2 bookkeeping: phys_vessel.Physics.points, phys_vessel.py:372-474, with next_point
 * (orbit_intersect.py:55-65), sort_intersect_indices (:68-79) and orbit_time_to
 * (phys_shared.py:58-64).
 * vessel12 (nv,12) = a b f ecc ma ea pea n dr pe_d ap_d period; body12 (nb,12) = a b f ecc ma ea
 * pea n dr coi size atm.  sel lists the vessels to process (the reference calls points(vessel)
 * one vessel at a time); entered_coi / left_coi / left_coi_prev_ref are the reference's
 * attributes of the same names (-1 = None).  coi_leave is in/out (only_enter_coi reads the
 * stored value, :441). */
static void next_point(double ea_vessel, const double *pts, int cnt, double dir, double *out2)
{
    out2[0] = out2[1] = NAN;
    if (cnt == 0) return;                       /* the reference's [nan] */
    for (int k = 0; k < cnt; ++k) if (isnan(pts[k])) return;
    double ang[4];
    int idx[4];
    for (int k = 0; k < cnt; ++k) {
        double a = fabs(ea_vessel - pts[k]);
        if (ea_vessel > pts[k]) a = 2 * VSO_PI - a;
        if (dir < 0) a = VSO_PI * 2 - a;
        ang[k] = a;
        idx[k] = k;
    }
    for (int i = 1; i < cnt; ++i) {             /* stable insertion argsort */
        int t = idx[i], j = i - 1;
        while (j >= 0 && ang[idx[j]] > ang[t]) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = t;
    }
    out2[0] = pts[idx[0]];
    out2[1] = pts[idx[cnt - 1]];
}

static double orbit_time_to(double src, double dst, double ecc, double dr, double n)
{
    if (ecc >= 1) {
        double v = -dr * (dst - src) / n;
        return (0.0 > v) ? 0.0 : v;             /* python max(v, 0.0): NaN stays NaN */
    }
    if (dr < 0) return py_mod(src - dst, 2 * VSO_PI) / n;
    return py_mod(dst - src, 2 * VSO_PI) / n;
}

static void advance_ea(double ecc, double ma, double ea, double n, double dr, double *next_ea)
{ /* :411-418 / :432-439 */
    if (ecc < 1) *next_ea = vso_solve_kepler_ell(ecc, ma + n * dr, 1e-10);
    else *next_ea = vso_newton_root_kepler_hyp(ecc, ma + n * dr * -1, ea);
}

void vso_vessel_points(int64_t nsel, const int64_t *sel, int64_t nb, const double *vessel12,
                       const int64_t *v_ref, const double *body12, const int64_t *body_ref,
                       int64_t entered_coi, int64_t left_coi, int64_t left_coi_prev_ref,
                       int only_enter_coi, int steps_limit, double *body_impact,
                       double *body_enter_atm, double *coi_leave, double *coi_enter)
{
    for (int64_t s = 0; s < nsel; ++s) {
        const int64_t v = sel[s];
        const double *V = vessel12 + 12 * v;
        const double a = V[0], b = V[1], f = V[2], ecc = V[3], ma = V[4], pea = V[6], n = V[7],
                     dr = V[8], pe_d = V[9], ap_d = V[10];
        double ea = V[5];
        const int ell = ecc < 1;
        const int64_t ref = v_ref[v];
        const double *R = body12 + 12 * ref;
        const double xc = f, yc = 0;
        double pts[4];
        int cnt;
        if (!only_enter_coi) {
            cnt = (ell && pe_d > R[10]) ? 0 : vso_ell_hyp_intersect_circle(a, b, ecc, xc, yc, R[10], pts);
            next_point(ea, pts, cnt, dr, body_impact + 2 * v);
            cnt = (ell && pe_d > R[10] + R[11]) ? 0
                      : vso_ell_hyp_intersect_circle(a, b, ecc, xc, yc, R[10] + R[11], pts);
            next_point(ea, pts, cnt, dr, body_enter_atm + 2 * v);
            if (ell && ap_d < R[9]) {
                coi_leave[2 * v] = coi_leave[2 * v + 1] = NAN;
            } else {
                cnt = vso_ell_hyp_intersect_circle(a, b, ecc, xc, yc, R[9], pts);
                double e2 = ea;
                if (entered_coi == v) advance_ea(ecc, ma, ea, n, dr, &e2);
                next_point(e2, pts, cnt, dr, coi_leave + 2 * v);
            }
        }
        /* enter_coi */
        if (left_coi == v) advance_ea(ecc, ma, ea, n, dr, &ea);
        double best_ang = 0;
        int64_t best = -1, rows = 0;
        double best_row[5];
        for (int64_t body = 0; body < nb; ++body) {
            if (body_ref[body] != ref || body == 0 || body == left_coi_prev_ref) continue;
            const double *B = body12 + 12 * body;
            double b_ma = B[4];
            if (left_coi == v) {
                if (ell) b_ma += B[7] * B[8];
                else b_ma += +B[7] * B[8] * -1;
            }
            double period;
            if (ecc < 1) {
                period = V[11];
            } else {
                double ea_leave;
                if (ref == 0) {
                    double lp[2];
                    cnt = vso_ell_hyp_intersect_circle(a, b, ecc, xc, yc, B[0] + B[9], pts);
                    next_point(ea, pts, cnt, dr, lp);
                    ea_leave = lp[0];
                    if (isnan(ea_leave)) continue;      /* row dropped: see the index quirk below */
                } else {
                    ea_leave = coi_leave[2 * v];
                }
                double ma_leave = (ecc * sinh(ea_leave) - ea_leave) * -1;
                period = orbit_time_to(ma, ma_leave, ecc, dr, n);
            }
            double vd[10] = {a, b, f, ecc, ma, ea, pea, period, n, dr};
            double bd[10] = {B[0], B[1], B[2], B[3], b_ma, B[5], B[6], B[7], B[8], B[9]};
            double row[5];
            vso_predict_enter_coi(vd, bd, 1e-5, steps_limit, row);
            if (!isnan(row[0])) {
                double ang = fabs(ea - row[0]);
                if (ea > row[0]) ang = 2 * VSO_PI - ang;
                if (dr < 0) ang = VSO_PI * 2 - ang;
                if (best < 0 || ang < best_ang) {
                    best = rows;
                    best_ang = ang;
                    for (int k = 0; k < 5; ++k) best_row[k] = row[k];
                }
            }
            ++rows;
        }
        double *ce = coi_enter + 6 * v;
        for (int k = 0; k < 6; ++k) ce[k] = NAN;
        if (best >= 0) {
            /* check_bodies[first_enter]: first_enter indexes the rows that were NOT dropped by the
             * `continue` above, but is applied to the full candidate list (:470-472). */
            int64_t seen = 0, bid = -1;
            for (int64_t body = 0; body < nb; ++body) {
                if (body_ref[body] != ref || body == 0 || body == left_coi_prev_ref) continue;
                if (seen == best) { bid = body; break; }
                ++seen;
            }
            ce[0] = (double)bid;
            for (int k = 0; k < 5; ++k) ce[1 + k] = best_row[k];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * N2 per-tick vessel events: phys_vessel.Physics.check_warp (:506-560), predict_enter_coi_service
 * (:744-758, which vessels it re-predicts) and cross_coi (:563-632) with convert.kepler_to_velocity
 * (convert.py:20-64), convert.velocity_to_kepler (:67-131), phys_shared.point_between (:67-80),
 * compare_coord (:35-55) and orb2xy (:187-199). */
int vso_point_between(double n, double a, double b, double dir)
{
    n = py_mod(n, 2 * VSO_PI);
    a = py_mod(a, 2 * VSO_PI);
    b = py_mod(b, 2 * VSO_PI);
    if (dir < 0) { double t = a; a = b; b = t; }
    if (a == b) return a == n;
    if (a < b) return a <= n && n <= b;
    return a <= n || n <= b;
}

/* sort_intersect_indices(ea, pts, dir)[0]: index of the nearest non-NaN point, -1 when all NaN */
static int first_intersect(double ea_vessel, const double *pts, int cnt, double dir)
{
    int best = -1;
    double best_ang = 0;
    for (int k = 0; k < cnt; ++k) {
        if (isnan(pts[k])) continue;
        double ang = fabs(ea_vessel - pts[k]);
        if (ea_vessel > pts[k]) ang = 2 * VSO_PI - ang;
        if (dir < 0) ang = VSO_PI * 2 - ang;
        if (best < 0 || ang < best_ang) { best = k; best_ang = ang; }
    }
    return best;
}

/* check_warp: hold (nv bytes) is the membership mask of physical_hold, updated in place.
 * out3 = situation (-1 = None), alarm_vessel (-1 = None), len(physical_hold). */
void vso_check_warp(int64_t nv, const double *ecc, const double *dr, const double *ea,
                    const double *ma, const double *n, const double *body_impact,
                    const double *body_enter_atm, const double *coi_leave, const double *coi_enter,
                    double warp, unsigned char *hold, int64_t *out3)
{
    int64_t situation = -1, alarm = -1, cnt = 0;
    for (int64_t v = 0; v < nv; ++v) {
        const double all[5] = {body_impact[2 * v], body_enter_atm[2 * v], body_enter_atm[2 * v + 1],
                               coi_enter[6 * v + 1], coi_leave[2 * v]};
        const int nx = first_intersect(ea[v], all, 5, dr[v]);
        if (nx >= 0) {
            double ea2;
            if (ecc[v] < 1) ea2 = vso_solve_kepler_ell(ecc[v], ma[v] + n[v] * warp * dr[v] * 2, 1e-10);
            else ea2 = vso_newton_root_kepler_hyp(ecc[v], ma[v] + n[v] * warp * dr[v] * -1 * 2, ea[v]);
            if (vso_point_between(all[nx], ea[v], ea2, dr[v])) {
                alarm = v;
                if (nx <= 1) { hold[v] = 1; situation = nx; }
                else if (nx == 2) { hold[v] = 0; situation = 2; }
                else situation = 3;
            } else {
                hold[v] = 0;
            }
        } else {
            hold[v] = 0;
        }
        cnt += hold[v];
    }
    out3[0] = situation; out3[1] = alarm; out3[2] = cnt;
}

/* predict_enter_coi_service: flags[v] = 1 where points(v, True) is re-run */
void vso_enter_coi_service(int64_t nv, const double *dr, const double *prev_ea, const double *ea,
                           const double *coi_enter, int any_crossing_pending, unsigned char *flags)
{
    for (int64_t v = 0; v < nv; ++v) {
        flags[v] = 0;
        if (!isnan(coi_enter[6 * v]) || any_crossing_pending) continue;
        if (vso_point_between(0.0, prev_ea[v], ea[v], dr[v]) ||
            vso_point_between(VSO_PI, prev_ea[v], ea[v], dr[v])) flags[v] = 1;
    }
}

/* cross_coi, detection part: first vessel whose next point is a COI crossing passed this tick.
 * out2 = vessel (-1 = none), kind (1 = enter, 2 = leave). */
void vso_cross_scan(int64_t nv, const double *ecc, const double *dr, const double *prev_ea,
                    const double *ea, const double *body_impact, const double *coi_leave,
                    const double *coi_enter, int64_t *out2)
{
    out2[0] = -1; out2[1] = 0;
    for (int64_t v = 0; v < nv; ++v) {
        const int ell = ecc[v] < 1;
        const double all[3] = {body_impact[2 * v + (ell ? 0 : 1)], coi_enter[6 * v + 1],
                               coi_leave[2 * v]};
        const int nx = first_intersect(prev_ea[v], all, 3, dr[v]);
        if (nx == 2 && vso_point_between(all[2], prev_ea[v], ea[v], dr[v])) {
            out2[0] = v; out2[1] = 2; return;
        }
        if (nx == 1 && vso_point_between(all[1], prev_ea[v], ea[v], dr[v])) {
            out2[0] = v; out2[1] = 1; return;
        }
    }
}

static void ps_orb2xy(double a, double b, double f, double ecc, double pea, double rx, double ry,
                      double ea, double *out)
{ /* phys_shared.py:187-199; ea == 0 gives (0, 0) WITHOUT the reference position */
    if (ea == 0) { out[0] = 0; out[1] = 0; return; }
    double xn, yn;
    if (ecc < 1) { xn = a * cos(ea) - f; yn = b * sin(ea); }
    else { xn = a * cosh(ea) - f; yn = b * sinh(ea); }
    out[0] = (xn * cos(pea - VSO_PI) - yn * sin(pea - VSO_PI)) + rx;
    out[1] = (xn * sin(pea - VSO_PI) + yn * cos(pea - VSO_PI)) + ry;
}

static double ps_compare(double a, double b)
{
    if (a == b) return 0.0;
    if (a == 0 || b == 0) return 100.0;
    if ((a > 0 && b > 0) || (a < 0 && b < 0)) {
        a = fabs(a); b = fabs(b);
        return fabs(a - b) / (a > b ? a : b) * 100;
    }
    return 100;
}

void vso_kepler_to_velocity(const double *rel_pos, double a, double ecc, double pe_arg, double u,
                            double dr, double *out2)
{
    const double sin_p = sin(pe_arg), cos_p = cos(pe_arg), sc = sin_p * cos_p,
                 sin_p2 = sin_p * sin_p, cos_p2 = cos_p * cos_p;
    double angle;
    if (ecc < 1) {
        const double b = a * sqrt(1 - ecc * ecc), f = sqrt(a * a - b * b);
        const double a2 = a * a, b2 = b * b;
        const double frx = f * cos_p, fry = f * sin_p;
        const double px = rel_pos[0] - f * cos_p, py = rel_pos[1] - f * sin_p;
        angle = atan(-(b2 * px * cos_p2 + a2 * px * sin_p2 + b2 * py * sc - a2 * py * sc) /
                     (a2 * py * cos_p2 + b2 * py * sin_p2 + b2 * px * sc - a2 * px * sc));
        const double x_max = sqrt(a2 * cos(2 * pe_arg) + a2 - b2 * cos(2 * pe_arg) + b2) / sqrt(2.0) - 1e-06;
        const double x2 = x_max * x_max, c4 = cos_p2 * cos_p2, s4 = sin_p2 * sin_p2;
        const double y_max = (a * b * sqrt(-x2 * (c4 + s4) - 2 * x2 * cos_p2 * sin_p2 + b2 * sin_p2 + a2 * cos_p2) +
                              a2 * x_max * sin_p * cos_p - b2 * x_max * sin_p * cos_p) /
                             (a2 * cos_p2 + b2 * sin_p2);
        const double x1 = x_max + frx, y1 = y_max + fry, xx2 = -x_max + frx, y2 = -y_max + fry;
        if (((xx2 - x1) * (rel_pos[1] - y1)) - ((y2 - y1) * (rel_pos[0] - x1)) < 0) angle += VSO_PI;
    } else {
        const double f = a * ecc, b = sqrt(f * f - a * a), a2 = a * a, b2 = b * b;
        const double frx = f * cos_p, fry = f * sin_p;
        const double px = rel_pos[0] - f * cosh(pe_arg), py = rel_pos[1] - f * sinh(pe_arg);
        angle = atan(-(b2 * px * cos_p2 - a2 * px * sin_p2 + b2 * py * sc + a2 * py * sc) /
                     (-a2 * py * cos_p2 + b2 * py * sin_p2 + b2 * px * sc + a2 * px * sc));
        const double arg = a2 * cos(2 * pe_arg) + a2 + b2 * cos(2 * pe_arg) - b2;
        if (arg >= 0) {
            const double x_max = sqrt(arg) / sqrt(2.0) - 1e-06;
            const double x2 = x_max * x_max, c4 = cos_p2 * cos_p2, s4 = sin_p2 * sin_p2;
            const double y_max = -((a * b * sqrt(-x2 * (c4 + s4) + 2 * x2 * cos_p2 * sin_p2 + b2 * sin_p2 - a2 * cos_p2) +
                                    a2 * x_max * sin_p * cos_p + b2 * x_max * sin_p * cos_p) /
                                   (-a2 * cos_p2 + b2 * sin_p2));
            const double x1 = x_max + frx, y1 = y_max + fry, xx2 = -x_max + frx, y2 = -y_max + fry;
            if (((xx2 - x1) * (rel_pos[1] - y1)) - ((y2 - y1) * (rel_pos[0] - x1)) < 0) angle += VSO_PI;
        } else if (VSO_PI <= pe_arg && pe_arg < 2 * VSO_PI) {   /* except ValueError */
            angle += VSO_PI;
        }
    }
    angle = py_mod(angle, 2 * VSO_PI);
    const double distance = sqrt(rel_pos[0] * rel_pos[0] + rel_pos[1] * rel_pos[1]);
    const double speed = dr * sqrt(u * ((2 / distance) - (1 / a)));
    out2[0] = speed * cos(angle);
    out2[1] = speed * sin(angle);
}

/* out5 = a, ecc, pe_arg, ma, dr */
void vso_velocity_to_kepler(const double *rel_pos, const double *rel_vel_in, double u, int failsafe,
                            double *out5)
{
    const double px = rel_pos[0], py = rel_pos[1];
    double vx = rel_vel_in[0], vy = rel_vel_in[1];
    const double mp = sqrt(px * px + py * py);
    const double a = -1 * u / (2 * ((vx * vx + vy * vy) / 2 - u / mp));
    const double mom = px * vy - py * vx;
    { double t = vx; vx = vy; vy = -t; }                 /* rel_vel[0], rel_vel[1] = rel_vel[1], -rel_vel[0] */
    const double ex = (vx * mom / u) - px / mp, ey = (vy * mom / u) - py / mp;
    const double ecc = sqrt(ex * ex + ey * ey);
    double pe_arg = py_mod((3 * VSO_PI / 2) + atan2(-ex, ey), 2 * VSO_PI);
    const double dr = copysign(1.0, mom);
    double ta = py_mod(pe_arg - (atan2(py, px) - VSO_PI), 2 * VSO_PI);
    if (dr > 0) ta = 2 * VSO_PI - ta;
    double ea, ma, f = 0, b = 0, np_[2];
    if (ecc < 1) {
        ea = acos((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
        if (VSO_PI < ta && ta < 2 * VSO_PI) ea = 2 * VSO_PI - ea;
        ma = py_mod(ea - ecc * sin(ea), 2 * VSO_PI);
        if (failsafe) {
            const double new_ea = vso_solve_kepler_ell(ecc, ma, 1e-10);
            f = a * ecc;
            b = sqrt(fabs(f * f - a * a));
            ps_orb2xy(a, b, f, ecc, pe_arg, 0, 0, new_ea, np_);
            if ((ps_compare(px, np_[0]) + ps_compare(py, np_[1])) / 2 > 1) {
                ea = 2 * VSO_PI - ea;
                ma = py_mod(ea - ecc * sin(ea), 2 * VSO_PI);
            }
        }
    } else {
        ea = acosh((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
        ma = ecc * sinh(ea) - ea;
        if (failsafe) {
            const double new_ea = vso_newton_root_kepler_hyp(ecc, ma, ma);
            f = a * ecc;
            b = sqrt(fabs(f * f - a * a));
            ps_orb2xy(a, b, f, ecc, pe_arg, 0, 0, new_ea, np_);
            if ((ps_compare(px, np_[0]) + ps_compare(py, np_[1])) / 2 > 1) {
                ea = -ea;
                ma = -ma;
            }
        }
    }
    if (failsafe == 2) {
        { double t = vx; vx = -vy; vy = t; }             /* undo the swap */
        const double va = py_mod(atan2(vy, vx), 2 * VSO_PI), pa = py_mod(atan2(py, px), 2 * VSO_PI);
        if (sin(3 * VSO_PI / 2 + va - pa) < 0) {
            if (ecc < 1) {
                ea = 2 * VSO_PI - ea;
                ma = py_mod(ea - ecc * sin(ea), 2 * VSO_PI);
            } else {
                ea = -ea;
                ma = -ma;
            }
            pe_arg = py_mod(pe_arg + VSO_PI, 2 * VSO_PI);
            ps_orb2xy(a, b, f, ecc, pe_arg, 0, 0, ea, np_);
            const double npa = py_mod(atan2(np_[1], np_[0]), 2 * VSO_PI);
            pe_arg = pe_arg - (npa - pa);
        }
    }
    out5[0] = a; out5[1] = ecc; out5[2] = pe_arg; out5[3] = ma; out5[4] = dr;
}

/* cross_coi, the orbit change of one vessel (:585-629).  vessel8 = a b f ecc pea u dr ea_cross;
 * body rows of body7 = a ecc pea dr mass pos_x pos_y.  kind 2 = leave (new_ref = body_ref[ref]),
 * 1 = enter (new_ref given).  out6 = a, ecc, pea, ma, dr, new_ref; returns 0, or 1 when the
 * reference returns None (enter with ref == new_ref). */
int vso_cross_apply(int kind, const double *vessel8, int64_t ref, int64_t new_ref,
                    const double *body7, double gc, double *out6)
{
    const double a = vessel8[0], b = vessel8[1], f = vessel8[2], ecc = vessel8[3], pea = vessel8[4],
                 u = vessel8[5], dr = vessel8[6], eax = vessel8[7];
    const double *R = body7 + 7 * ref, *N = body7 + 7 * new_ref;
    if (kind == 1 && ref == new_ref) return 1;
    const double new_u = gc * N[4];
    double vabs[2], vrel[2], vv[2], rv[2], npos[2], nvel[2], rrel[2];
    ps_orb2xy(a, b, f, ecc, pea, R[5], R[6], eax, vabs);
    vrel[0] = vabs[0] - R[5]; vrel[1] = vabs[1] - R[6];
    if (kind == 2) {
        rrel[0] = R[5] - N[5]; rrel[1] = R[6] - N[6];
        vso_kepler_to_velocity(vrel, a, ecc, pea, u, dr, vv);
        vso_kepler_to_velocity(rrel, R[0], R[1], R[2], new_u, R[3], rv);
        npos[0] = vrel[0] + rrel[0]; npos[1] = vrel[1] + rrel[1];
        nvel[0] = vv[0] + rv[0]; nvel[1] = vv[1] + rv[1];
        vso_velocity_to_kepler(npos, nvel, new_u, 1, out6);
    } else {
        rrel[0] = N[5] - R[5]; rrel[1] = N[6] - R[6];
        vso_kepler_to_velocity(vrel, a, ecc, pea, u, dr, vv);
        vso_kepler_to_velocity(rrel, N[0], N[1], N[2], new_u, N[3], rv);
        npos[0] = vrel[0] - rrel[0]; npos[1] = vrel[1] - rrel[1];
        nvel[0] = vv[0] - rv[0]; nvel[1] = vv[1] - rv[1];
        vso_velocity_to_kepler(npos, nvel, new_u, 2, out6);
    }
    out6[5] = (double)new_ref;
    return 0;
}
