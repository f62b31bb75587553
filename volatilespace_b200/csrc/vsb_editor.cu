// vsb_editor.cu -- O(N) pieces of the editor / body-sim tick (sm_100a):
//   A2+A3  Physics.gravity            phys_editor.py:357-378 (gravity_one :51-60)
//   A6     Physics.kepler_basic       phys_editor.py:381-384 (kepler_basic_one :63-83)
//   A7     simplified_orbit_coi       phys_editor.py:325-346
//   B3     del_body / set_root / re-rooting moves  phys_editor.py:250-313, 844-857
//   C7     editor Physics.curve       phys_editor.py:519-544
// Compiled with -fmad=false: the reference runs un-fused (numba fastmath=False).
#include <stdlib.h>

#include "vsb_common.cuh"
#include "vsb_orbit_math.cuh"

namespace {

// ------------------------------------------------------------------ A2/A3
// gravity_one, phys_editor.py:51-60 (body != 0): acceleration of `pi` toward `pp`.
__device__ __forceinline__ double2 gravity_one_dev(double2 pi, double2 pp, double mb, double mp,
                                                   double gc)
{
    double rx = pp.x - pi.x, ry = pp.y - pi.y;
    double distance = sqrt(rx * rx + ry * ry);
    double force = gc * mb * mp / (distance * distance);
    double angle = atan2(ry, rx);
    double acc = force / mb;
    double sn, cs;
    sincos(angle, &sn, &cs);
    return make_double2(acc * cs, acc * sn);
}

// rel_vel += gravity_one(body, parents[body], ...) for every body (:361-363).
__global__ void ref_accel_kernel(int64_t n, const double *__restrict__ mass,
                                 const double2 *__restrict__ pos,
                                 const int64_t *__restrict__ parents, double gc,
                                 double2 *__restrict__ rel_vel)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;  // root: acceleration [0,0] (:60)
    int64_t p = parents[i];
    double2 a = gravity_one_dev(pos[i], pos[p], mass[i], mass[p], gc);
    double2 rv = rel_vel[i];
    rv.x += a.x;
    rv.y += a.y;
    rel_vel[i] = rv;
}

// COI enter/leave fix-up (:366-373): sequential in index order because a body may
// read the rel_vel of a body fixed earlier in the same loop.  A grid-wide pass flags the
// 1024-wide slabs that contain a parent change; one CTA then walks only those, in index order.
// slab_flag[s] = 1 if any body of the 1024-wide slab s changed its parent this tick
__global__ void __launch_bounds__(1024)
ref_changed_kernel(int64_t n, const int64_t *__restrict__ parents,
                   const int64_t *__restrict__ parents_old, unsigned char *__restrict__ slab_flag)
{
    const int64_t b = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int ch = (b < n && b != 0 && parents[b] != parents_old[b]) ? 1 : 0;
    const int any = __syncthreads_or(ch);
    if (threadIdx.x == 0) slab_flag[blockIdx.x] = any ? 1 : 0;
}

__global__ void __launch_bounds__(1024)
ref_fixup_kernel(int64_t n, const double *__restrict__ mass, const int64_t *__restrict__ parents,
                 const int64_t *__restrict__ parents_old, double2 *rel_vel,
                 const unsigned char *__restrict__ slab_flag, int64_t nslabs)
{
    __shared__ int changed[1024];
    __shared__ unsigned char flagged[1024];
    for (int64_t s0 = 0; s0 < nslabs; s0 += 1024) {
        const int f = (s0 + threadIdx.x < nslabs) ? slab_flag[s0 + threadIdx.x] : 0;
        if (!__syncthreads_or(f)) continue;
        flagged[threadIdx.x] = (unsigned char)f;
        __syncthreads();
        for (int sl = 0; sl < 1024; ++sl) {
            if (!flagged[sl]) continue;      // uniform: shared flag
            const int64_t base = (s0 + sl) * 1024;
            const int64_t b = base + threadIdx.x;
            changed[threadIdx.x] = (b < n && b != 0 && parents[b] != parents_old[b]) ? 1 : 0;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int t = 0; t < 1024; ++t) {
                    if (!changed[t]) continue;
                    int64_t body = base + t;
                    int64_t parent = parents[body], old = parents_old[body];
                    volatile double *rv = (volatile double *)rel_vel;
                    if (mass[parent] > mass[old]) {  // leaving orbit
                        rv[2 * body] += rv[2 * old];
                        rv[2 * body + 1] += rv[2 * old + 1];
                    }
                    if (mass[parent] < mass[old]) {  // entering orbit
                        rv[2 * body] -= rv[2 * parent];
                        rv[2 * body + 1] -= rv[2 * parent + 1];
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// vel compose in descending-mass order + pos += vel (:375-378), parallelised by
// letting each body fold its own ancestor chain top-down; the arithmetic per
// body is exactly vel[b] = rel_vel[b] + vel[parents[b]] with vel[parent] being
// the NEW value when the parent comes earlier in the order and last tick's
// otherwise.  order[0] is never recomputed (the loop starts at bodies_sorted[1]).
__global__ void ref_compose_kernel(int64_t n, const int64_t *__restrict__ rank,
                                   const int64_t *__restrict__ parents,
                                   const double2 *__restrict__ rel_vel,
                                   const double2 *__restrict__ vel_old, double2 *__restrict__ vel,
                                   double2 *__restrict__ pos, int *overflow)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    constexpr int MAXD = 48;
    int64_t chain[MAXD];
    int depth = 0;
    int64_t node = b;
    double2 v;
    for (;;) {
        if (rank[node] == 0) {  // heaviest body: keeps its old velocity
            v = vel_old[node];
            break;
        }
        if (depth == MAXD) {
            atomicExch(overflow, 1);
            return;
        }
        chain[depth++] = node;
        int64_t p = parents[node];
        if (rank[p] < rank[node]) {
            node = p;  // parent already recomputed this tick: descend into it
        } else {
            v = vel_old[p];  // parent comes later in the order (or is the body itself)
            break;
        }
    }
    for (int d = depth - 1; d >= 0; --d) {
        double2 r = rel_vel[chain[d]];
        v.x = r.x + v.x;
        v.y = r.y + v.y;
    }
    vel[b] = v;
    double2 p = pos[b];
    p.x += v.x;
    p.y += v.y;
    pos[b] = p;
}

// ------------------------------------------------------------------ A6
struct kb_out {
    double focus, semi_major, semi_minor, periapsis_arg, coi;
    double2 ecc_v;
};

// kepler_basic_one, phys_editor.py:63-83 (body != 0)
__device__ __forceinline__ kb_out kepler_basic_one_dev(double2 pb, double2 pp, double2 vb,
                                                       double2 vp, double mb, double mp, double gc,
                                                       double coi_coef)
{
    double rpx = pb.x - pp.x, rpy = pb.y - pp.y;
    double rvx = vb.x - vp.x, rvy = vb.y - vp.y;
    double u = gc * mp;
    double magp = sqrt(rpx * rpx + rpy * rpy);
    double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / magp));
    double momentum = rpx * rvy - rpy * rvx;
    double sx = rvy, sy = -rvx;
    double ex = (sx * momentum / u) - rpx / magp;
    double ey = (sy * momentum / u) - rpy / magp;
    double ecc = sqrt(ex * ex + ey * ey);
    kb_out o;
    o.periapsis_arg = py_mod((3 * PI_D / 2) + atan2(-ex, ey), 2 * PI_D);
    o.focus = a * ecc;
    o.semi_minor = sqrt(fabs(o.focus * o.focus - a * a));
    o.coi = a > 0 ? a * pow(mb / mp, coi_coef) : 0.0;
    o.semi_major = a;
    o.ecc_v = make_double2(ex, ey);
    return o;
}

__global__ void kepler_basic_kernel(int64_t n, const int64_t *__restrict__ parents,
                                    const double *__restrict__ mass,
                                    const double2 *__restrict__ pos,
                                    const double2 *__restrict__ vel, double gc, double coi_coef,
                                    double *__restrict__ focus, double2 *__restrict__ ecc_v,
                                    double *__restrict__ semi_major,
                                    double *__restrict__ semi_minor,
                                    double *__restrict__ periapsis_arg, double *__restrict__ coi)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    if (b == 0) {
        focus[0] = 0.0;
        ecc_v[0] = make_double2(0.0, 0.0);
        semi_major[0] = 0.0;
        semi_minor[0] = 0.0;
        periapsis_arg[0] = 0.0;
        coi[0] = 0.0;
        return;
    }
    kb_out o = kepler_basic_one_dev(pos[b], pos[parents[b]], vel[b], vel[parents[b]], mass[b],
                                    mass[parents[b]], gc, coi_coef);
    focus[b] = o.focus;
    ecc_v[b] = o.ecc_v;
    semi_major[b] = o.semi_major;
    semi_minor[b] = o.semi_minor;
    periapsis_arg[b] = o.periapsis_arg;
    coi[b] = o.coi;
}

// ------------------------------------------------------------------ A7
// Right-looking blocked sweep over ranks (descending mass).  For a slab of T
// consecutive ranks: soc_slab_kernel resolves the slab inside one CTA (rank s needs
// coi of every earlier rank: fixed-point rounds, see there), then soc_update_kernel tests all
// later ranks against the slab's T finished COIs (grid-wide, "last match wins" =
// max rank).  best[] holds the rank of the current parent candidate (-1: none).
constexpr int SOC_T = 1024;

__global__ void soc_init_kernel(int64_t npad, int *best, double *scoi)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npad) {
        best[k] = -1;
        scoi[k] = 0.0;
    }
}

// One slab of SOC_T consecutive ranks.  The reference's loop is sequential (rank s needs the COI
// of every earlier rank), but a COI only depends on the body's parent, and the parent only on the
// COIs of earlier ranks: iterate "all COIs from the current parents, all parents from the current
// COIs" to the fixed point.  After k rounds the first k ranks of every dependency chain inside
// the slab are final (induction over the rank), so the fixed point is the sequential answer --
// same expressions, bit-identical -- and it is reached after (longest chain inside the slab + 1)
// rounds, typically 2-3, instead of SOC_T barrier-separated steps.
__global__ void __launch_bounds__(SOC_T)
soc_slab_kernel(int64_t n, int64_t r0, const int64_t *__restrict__ order, int64_t largest,
                const double *__restrict__ mass, const double2 *__restrict__ pos,
                const double2 *__restrict__ vel, double gc, double coi_coef, int *best,
                double *scoi /* coi by rank */, double *__restrict__ coi /* by body */)
{
    __shared__ double2 s_pos[SOC_T];
    __shared__ double s_coi2[SOC_T];
    const int t = threadIdx.x;
    const int64_t r = r0 + t;
    const bool live = r < n;
    const int64_t body = live ? order[r] : 0;
    const double2 pb = live ? pos[body] : make_double2(VSB_FAR, VSB_FAR);
    const double2 vb = live ? vel[body] : make_double2(0.0, 0.0);
    const double mb = live ? mass[body] : 0.0;
    s_pos[t] = pb;
    const int outside = live ? best[r] : -1;   // parent candidate among the ranks before the slab
    int my_best = outside;
    double c = 0.0;
    for (;;) {
        c = 0.0;
        if (live && r > 0) {  // rank 0 keeps coi 0 (loop starts at bodies_sorted[1:], :332)
            const int64_t parent = my_best < 0 ? largest : order[my_best];
            c = soc_coi_of(pb, vb, mb, pos[parent], vel[parent], mass[parent], gc, coi_coef);
        }
        __syncthreads();          // everyone is done reading s_coi2 of the previous round
        s_coi2[t] = live ? c * c : -1.0;
        __syncthreads();
        int nb = outside;
        if (live) {
            for (int s = 0; s < t; ++s) {
                // (pos[body]-pos[pot])**2 summed < coi[pot]**2  (:337); last match wins
                double dx = pb.x - s_pos[s].x, dy = pb.y - s_pos[s].y;
                if (dx * dx + dy * dy < s_coi2[s]) nb = (int)(r0 + s);
            }
        }
        const int changed = nb != my_best;
        my_best = nb;
        if (!__syncthreads_or(changed)) break;
    }
    if (live) {
        coi[body] = c;
        scoi[r] = c;
        best[r] = my_best;
    }
}

__global__ void __launch_bounds__(256)
soc_update_kernel(int64_t n, int64_t r0, int cnt, const int64_t *__restrict__ order,
                  const double2 *__restrict__ pos, const double *__restrict__ scoi, int *best)
{
    __shared__ double2 s_pos[SOC_T];
    __shared__ double s_coi2[SOC_T];
    for (int s = threadIdx.x; s < cnt; s += blockDim.x) {
        s_pos[s] = pos[order[r0 + s]];
        double c = scoi[r0 + s];
        s_coi2[s] = c * c;
    }
    __syncthreads();
    int64_t r = r0 + cnt + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double2 pb = pos[order[r]];
    int found = -1;
    for (int s = 0; s < cnt; ++s) {
        double dx = pb.x - s_pos[s].x, dy = pb.y - s_pos[s].y;
        if (dx * dx + dy * dy < s_coi2[s]) found = s;
    }
    if (found >= 0) best[r] = (int)r0 + found;  // later slabs overwrite earlier ones: max rank
}

// first index of the maximum mass (np.argmax, :327)
__global__ void __launch_bounds__(1024) argmax_kernel(int64_t n, const double *__restrict__ mass,
                                                      int64_t *out)
{
    __shared__ double sv[1024];
    __shared__ int64_t si[1024];
    double bv = -1.0e308 * 10;  // -inf
    int64_t bi = 0x7fffffffffffffffll;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        double m = mass[i];
        if (m > bv) {
            bv = m;
            bi = i;
        }
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double v2 = sv[threadIdx.x + o];
            int64_t i2 = si[threadIdx.x + o];
            if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = si[0];
}

// ------------------------------------------------------------------ B3 structural
struct body_arrays {
    double *w1[8];    // mass den rad coi focus semi_major semi_minor periapsis_arg
    int64_t *i1;      // parents
    double2 *w2[5];   // pos vel rel_vel acc ecc_v
};

// np.delete(row): every row above `del` moves down by one, order preserved.  Each
// thread stages row r+1 of all arrays in registers, the CTA synchronises, then
// writes row r; the one row a CTA needs from its right neighbour was saved to
// `edge` by delete_edges_kernel beforehand, so no CTA reads what another writes.
constexpr int DEL_T = 256;

__global__ void delete_edges_kernel(body_arrays A, int64_t del, int64_t n, double *edge)
{
    // CTA k of the move kernel covers rows [del + k*T, del + (k+1)*T); it needs row del+(k+1)*T
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t src = del + (k + 1) * DEL_T;
    if (src >= n) return;
    double *e = edge + k * 19;
#pragma unroll
    for (int a = 0; a < 8; ++a) e[a] = A.w1[a][src];
    e[8] = __longlong_as_double(A.i1[src]);
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        double2 v = A.w2[a][src];
        e[9 + 2 * a] = v.x;
        e[10 + 2 * a] = v.y;
    }
}

__global__ void __launch_bounds__(DEL_T)
delete_move_kernel(body_arrays A, int64_t del, int64_t n, const double *__restrict__ edge)
{
    const int64_t dst = del + (int64_t)blockIdx.x * DEL_T + threadIdx.x;
    const int64_t src = dst + 1;
    const bool live = src < n;
    double v[19];
    if (live) {
        if (threadIdx.x == DEL_T - 1) {
            const double *e = edge + (int64_t)blockIdx.x * 19;
#pragma unroll
            for (int a = 0; a < 19; ++a) v[a] = e[a];
        } else {
#pragma unroll
            for (int a = 0; a < 8; ++a) v[a] = A.w1[a][src];
            v[8] = __longlong_as_double(A.i1[src]);
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                double2 q = A.w2[a][src];
                v[9 + 2 * a] = q.x;
                v[10 + 2 * a] = q.y;
            }
        }
    }
    __syncthreads();
    if (live) {
#pragma unroll
        for (int a = 0; a < 8; ++a) A.w1[a][dst] = v[a];
        A.i1[dst] = __double_as_longlong(v[8]);
#pragma unroll
        for (int a = 0; a < 5; ++a) A.w2[a][dst] = make_double2(v[9 + 2 * a], v[10 + 2 * a]);
    }
}

// rows >= n (up to npad) become inert padding: mass 0, far position.
__global__ void pad_tail_kernel(body_arrays A, int64_t n, int64_t npad)
{
    int64_t r = n + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= npad) return;
#pragma unroll
    for (int a = 0; a < 8; ++a) A.w1[a][r] = 0.0;
    A.i1[r] = 0;
#pragma unroll
    for (int a = 0; a < 5; ++a) A.w2[a][r] = make_double2(0.0, 0.0);
    A.w2[0][r] = make_double2(VSB_FAR, VSB_FAR);
}

// swap_with_first-style swap (set_root :293-313): mass den rad pos vel rel_vel parents
__global__ void swap_rows_kernel(body_arrays A, int64_t a, int64_t b)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < 3; ++k) {
        double t = A.w1[k][a];
        A.w1[k][a] = A.w1[k][b];
        A.w1[k][b] = t;
    }
    int64_t ti = A.i1[a];
    A.i1[a] = A.i1[b];
    A.i1[b] = ti;
    for (int k = 0; k < 3; ++k) {
        double2 t = A.w2[k][a];
        A.w2[k][a] = A.w2[k][b];
        A.w2[k][b] = t;
    }
}

__global__ void translate_kernel(int64_t n, double2 *pos, const double2 *diff)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double2 d = *diff;
    double2 p = pos[r];
    p.x -= d.x;
    p.y -= d.y;
    pos[r] = p;
}

// ------------------------------------------------------------------ C7
// One warp per body, lanes over curve points; out[axis][body][k].
__global__ void __launch_bounds__(256)
editor_curve_kernel(int64_t n, int64_t P, const double2 *__restrict__ ecc_v,
                    const double *__restrict__ semi_major, const double *__restrict__ semi_minor,
                    const double *__restrict__ focus, const double *__restrict__ periapsis_arg,
                    const int64_t *__restrict__ parents, const double2 *__restrict__ pos,
                    const double *__restrict__ tabs /* ell_t | par_t, P each */,
                    double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t body = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (body >= n) return;
    const double2 ev = ecc_v[body];
    const double ecc = sqrt(ev.x * ev.x + ev.y * ev.y);
    double sp, cp;
    sincos(periapsis_arg[body], &sp, &cp);
    const double a = semi_major[body], b = semi_minor[body];
    const double fx = focus[body] * cp, fy = focus[body] * sp;
    const double2 pp = pos[parents[body]];
    const double *ell_t = tabs, *par_t = tabs + P;
    double *ox = out + body * P, *oy = out + (n + body) * P;
    for (int64_t k = lane; k < P; k += 32) {
        double x = 0.0, y = 0.0;
        if (ecc < 1) {
            double s, c;
            sincos(ell_t[k], &s, &c);
            x = a * c;
            y = b * s;
        } else if (ecc == 1) {
            double t = par_t[k];
            x = a * (t * t) - a;
            y = 2 * a * t;
        } else if (ecc > 1) {
            double t = ell_t[k];
            x = -a * cosh(t);
            y = b * sinh(t);
        }
        double xr = cp * x + (-sp) * y;
        double yr = sp * x + cp * y;
        ox[k] = xr + fx + pp.x;
        oy[k] = yr + fy + pp.y;
    }
}

// ------------------------------------------------------------------ fused small-N frame
// Up to `max_ticks` editor ticks (gravity(); [body();] check_collision()) for n <= 1024 in ONE
// launch of ONE CTA, all state in shared memory; stops after the tick in which a collision is
// found (the host applies the merge, phys_editor.py:554-565, and calls again for the remaining
// ticks) and runs kepler_basic at the end of an undisturbed frame (editor.py:1503-1523).
// Same device functions and operation order as the per-kernel path above: results are
// bit-identical to it.  hdr = {ticks_done, body_del, body_add, chain_overflow}.
constexpr int FR_MAX = 1024;

__global__ void __launch_bounds__(FR_MAX)
editor_frame_small_kernel(int n, int max_ticks, double gc, double coi_coef, int do_kepler_basic,
                          int do_collision, const int64_t *__restrict__ order, const double *__restrict__ g_mass,
                          double2 *g_pos, double2 *g_vel, double2 *g_rel_vel, double *g_coi,
                          const double *__restrict__ g_rad, int64_t *g_parents, double *g_focus,
                          double2 *g_ecc_v, double *g_semi_major, double *g_semi_minor,
                          double *g_periapsis_arg, int64_t *hdr, double *packed, long long *host_flag, long long seq)
{
    extern __shared__ double2 fr_smem[];
    double2 *s_pos = fr_smem, *s_vel = s_pos + n, *s_rv = s_vel + n, *s_spos = s_rv + n;
    double *s_mass = (double *)(s_spos + n), *s_rad = s_mass + n, *s_scoi2 = s_rad + n;
    int *s_par = (int *)(s_scoi2 + n), *s_rank = s_par + n, *s_ord = s_rank + n,
        *s_changed = s_ord + n;
    __shared__ unsigned long long s_key[32];
    __shared__ int s_overflow;

    const int t = threadIdx.x;
    const bool live = t < n;
    if (t == 0) s_overflow = 0;
    if (live) {
        s_pos[t] = g_pos[t];
        s_vel[t] = g_vel[t];
        s_rv[t] = g_rel_vel[t];
        s_mass[t] = g_mass[t];
        s_rad[t] = g_rad[t];
        s_par[t] = (int)g_parents[t];
        const int body = (int)order[t];   // t plays the role of a rank here
        s_ord[t] = body;
        s_rank[body] = t;
        const double c = g_coi[body];
        s_scoi2[t] = c * c;
    }
    __syncthreads();

    int ticks = 0;
    long long hit_del = -1, hit_add = -1;
    for (; ticks < max_ticks;) {
        // ---- find_parents (:350-354): thread <-> rank
        const int par_old = live ? s_par[t] : 0;
        if (live) s_spos[t] = s_pos[s_ord[t]];
        __syncthreads();
        if (live) {
            const double2 me = s_spos[t];
            int best = -1;
            for (int k = 0; k < t; ++k) {
                double dx = me.x - s_spos[k].x, dy = me.y - s_spos[k].y;
                if (dx * dx + dy * dy < s_scoi2[k]) best = k;
            }
            const int body = s_ord[t];
            s_par[body] = (body == 0 || best < 0) ? 0 : s_ord[best];
        }
        __syncthreads();
        // ---- rel_vel += gravity_one (:361-363): thread <-> body
        int changed = 0;
        if (live && t != 0) {
            const int p = s_par[t];
            double2 a = gravity_one_dev(s_pos[t], s_pos[p], s_mass[t], s_mass[p], gc);
            double2 rv = s_rv[t];
            rv.x += a.x;
            rv.y += a.y;
            s_rv[t] = rv;
            changed = p != par_old;
        }
        if (live) s_changed[t] = changed ? par_old + 1 : 0;   // old parent + 1, 0 = unchanged
        // ---- COI enter/leave fix-up, sequential in index order (:366-373)
        if (__syncthreads_or(changed)) {
            if (t == 0) {
                for (int b = 1; b < n; ++b) {
                    if (!s_changed[b]) continue;
                    const int parent = s_par[b], old = s_changed[b] - 1;
                    if (s_mass[parent] > s_mass[old]) {
                        s_rv[b].x += s_rv[old].x;
                        s_rv[b].y += s_rv[old].y;
                    }
                    if (s_mass[parent] < s_mass[old]) {
                        s_rv[b].x -= s_rv[parent].x;
                        s_rv[b].y -= s_rv[parent].y;
                    }
                }
            }
            __syncthreads();
        }
        // ---- vel compose in descending-mass order + pos += vel (:375-378)
        double2 v = make_double2(0.0, 0.0);
        if (live) {
            int chain[48];
            int depth = 0, node = t;
            for (;;) {
                if (s_rank[node] == 0) {
                    v = s_vel[node];
                    break;
                }
                if (depth == 48) {
                    s_overflow = 1;
                    break;
                }
                chain[depth++] = node;
                const int p = s_par[node];
                if (s_rank[p] < s_rank[node]) {
                    node = p;
                } else {
                    v = s_vel[p];
                    break;
                }
            }
            for (int d = depth - 1; d >= 0; --d) {
                double2 r = s_rv[chain[d]];
                v.x = r.x + v.x;
                v.y = r.y + v.y;
            }
        }
        __syncthreads();
        if (live) {
            s_vel[t] = v;
            double2 p = s_pos[t];
            p.x += v.x;
            p.y += v.y;
            s_pos[t] = p;
        }
        __syncthreads();
        ++ticks;
        if (!do_collision) continue;   // plain Physics.gravity(): no collision phase
        // ---- check_collision (:547-551): first j > i per thread, then the smallest (i,j)
        unsigned long long key = ~0ull;
        if (t < n - 1) {
            const double2 me = s_pos[t];
            const double ri = s_rad[t];
            for (int j = t + 1; j < n; ++j) {
                double dx = me.x - s_pos[j].x, dy = me.y - s_pos[j].y;
                double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                if (d <= __dadd_rn(ri, s_rad[j])) {
                    key = ((unsigned long long)t << 32) | (unsigned)j;
                    break;
                }
            }
        }
        key = warp_min_u64(key);
        if ((t & 31) == 0) s_key[t >> 5] = key;
        __syncthreads();
        if (t < 32) {
            key = t < (int)(blockDim.x >> 5) ? s_key[t] : ~0ull;
            key = warp_min_u64(key);
            if (t == 0) s_key[0] = key;
        }
        __syncthreads();
        key = s_key[0];
        __syncthreads();
        if (key != ~0ull) {
            const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
            if (s_mass[i] <= s_mass[j]) {
                hit_del = i;
                hit_add = j;
            } else {
                hit_del = j;
                hit_add = i;
            }
            break;
        }
    }

    if (live) {
        g_pos[t] = s_pos[t];
        g_vel[t] = s_vel[t];
        g_rel_vel[t] = s_rv[t];
        g_parents[t] = s_par[t];
        kb_out o;
        if (do_kepler_basic && hit_del < 0 && ticks == max_ticks) {
            if (t == 0) {
                o.focus = o.semi_major = o.semi_minor = o.periapsis_arg = o.coi = 0.0;
                o.ecc_v = make_double2(0.0, 0.0);
            } else {
                const int p = s_par[t];
                o = kepler_basic_one_dev(s_pos[t], s_pos[p], s_vel[t], s_vel[p], s_mass[t],
                                         s_mass[p], gc, coi_coef);
            }
            g_focus[t] = o.focus;
            g_ecc_v[t] = o.ecc_v;
            g_semi_major[t] = o.semi_major;
            g_semi_minor[t] = o.semi_minor;
            g_periapsis_arg[t] = o.periapsis_arg;
            g_coi[t] = o.coi;
        } else {
            o.focus = g_focus[t];
            o.ecc_v = g_ecc_v[t];
            o.semi_major = g_semi_major[t];
            o.semi_minor = g_semi_minor[t];
            o.periapsis_arg = g_periapsis_arg[t];
            o.coi = g_coi[t];
        }
        if (packed) {
            // frame record after the 4-word header: pos vel rel_vel parents focus ecc_v
            // semi_major semi_minor periapsis_arg coi  (14 n words)
            double *q = packed + 4;
            ((double2 *)q)[t] = s_pos[t];
            ((double2 *)(q + 2 * n))[t] = s_vel[t];
            ((double2 *)(q + 4 * n))[t] = s_rv[t];
            ((long long *)(q + 6 * n))[t] = s_par[t];
            (q + 7 * n)[t] = o.focus;
            ((double2 *)(q + 8 * n))[t] = o.ecc_v;
            (q + 10 * n)[t] = o.semi_major;
            (q + 11 * n)[t] = o.semi_minor;
            (q + 12 * n)[t] = o.periapsis_arg;
            (q + 13 * n)[t] = o.coi;
        }
    }
    if (host_flag) {
        // `packed` is mapped host memory: every thread's record words are made visible to the host
        // before thread 0 publishes the sequence word the host spins on
        __threadfence_system();
        __syncthreads();
    }
    if (t == 0) {
        hdr[0] = ticks;
        hdr[1] = hit_del;
        hdr[2] = hit_add;
        hdr[3] = s_overflow;
        if (packed) {
            long long *ph = (long long *)packed;
            ph[0] = ticks;
            ph[1] = hit_del;
            ph[2] = hit_add;
            ph[3] = s_overflow;
        }
        if (host_flag) {
            // header copy next to the sequence word (the caller may not have asked for the record)
            host_flag[1] = ticks;
            host_flag[2] = hit_del;
            host_flag[3] = hit_add;
            host_flag[4] = s_overflow;
            __threadfence_system();
            *(volatile long long *)host_flag = seq;
        }
    }
}

// check_collision (phys_editor.py:547-551) of a small system (n <= 1024) in ONE launch of one CTA:
// thread i scans j > i in ascending order for the first touching body (the same un-fused predicate as
// cc_scan_kernel's exact path and the fused frame), the CTA takes the smallest (i, j), the lighter
// body is deleted (ties delete i, :93-97).  The answer goes to out2 (device) and, when host_flag is
// given, into mapped host memory: host_flag[1..2] = (del, add), then host_flag[0] = seq.
__global__ void __launch_bounds__(FR_MAX)
collision_small_kernel(int n, const double2 *__restrict__ g_pos, const double *__restrict__ g_rad,
                       const double *__restrict__ g_mass, int64_t *out2, long long *host_flag,
                       long long seq)
{
    extern __shared__ double2 cs_smem[];
    double2 *s_pos = cs_smem;
    double *s_rad = (double *)(s_pos + n);
    __shared__ unsigned long long s_key[32];
    const int t = threadIdx.x;
    if (t < n) {
        s_pos[t] = g_pos[t];
        s_rad[t] = g_rad[t];
    }
    __syncthreads();
    unsigned long long key = ~0ull;
    if (t < n - 1) {
        const double2 me = s_pos[t];
        const double ri = s_rad[t];
        for (int j = t + 1; j < n; ++j) {
            double dx = me.x - s_pos[j].x, dy = me.y - s_pos[j].y;
            double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            if (d <= __dadd_rn(ri, s_rad[j])) {
                key = ((unsigned long long)t << 32) | (unsigned)j;
                break;
            }
        }
    }
    key = warp_min_u64(key);
    if ((t & 31) == 0) s_key[t >> 5] = key;
    __syncthreads();
    if (t < 32) {
        key = t < (int)(blockDim.x >> 5) ? s_key[t] : ~0ull;
        key = warp_min_u64(key);
        if (t == 0) {
            long long del = -1, add = -1;
            if (key != ~0ull) {
                const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
                const bool first = g_mass[i] <= g_mass[j];
                del = first ? i : j;
                add = first ? j : i;
            }
            out2[0] = del;
            out2[1] = add;
            if (host_flag) {
                host_flag[1] = del;
                host_flag[2] = add;
                __threadfence_system();
                *(volatile long long *)host_flag = seq;
            }
        }
    }
}

// gather several body fields into one contiguous staging buffer (one D2H copy per frame)
struct pack_desc {
    const double *src[16];
    int width[16];
    int count;
};

__global__ void pack_fields_kernel(pack_desc d, int64_t n, double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t off = 0;
    for (int f = 0; f < d.count; ++f) {
        const int64_t words = n * d.width[f];
        if (i < words) out[off + i] = d.src[f][i];
        off += words;
    }
}

// the same for a small system, in one CTA, straight into mapped host memory: `out` and host_flag are
// device aliases of page-locked host words; the sequence word is published after the data
__global__ void __launch_bounds__(1024)
pack_fields_small_kernel(pack_desc d, int64_t n, double *out, long long *host_flag, long long seq)
{
    int64_t off = 0;
    for (int f = 0; f < d.count; ++f) {
        const int64_t words = n * d.width[f];
        for (int64_t i = threadIdx.x; i < words; i += blockDim.x) out[off + i] = d.src[f][i];
        off += words;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *(volatile long long *)host_flag = seq;
}

body_arrays arrays_of(vsb_ctx *c)
{
    body_arrays A;
    A.w1[0] = c->mass; A.w1[1] = c->den; A.w1[2] = c->rad; A.w1[3] = c->coi;
    A.w1[4] = c->focus; A.w1[5] = c->semi_major; A.w1[6] = c->semi_minor;
    A.w1[7] = c->periapsis_arg;
    A.i1 = c->parents;
    A.w2[0] = (double2 *)c->pos; A.w2[1] = (double2 *)c->vel; A.w2[2] = (double2 *)c->rel_vel;
    A.w2[3] = (double2 *)c->acc; A.w2[4] = (double2 *)c->ecc_v;
    return A;
}

}  // namespace

// The MODE_REF tick in two halves so that the find_parents scan can be split over ranks:
//   begin : parents_old snapshot + this part's share of the scan (best[] in c->sorted)
//   finish: best[] -> parents, gravity_one per body, chain composition, integrator
int vsb_launch_gravity_ref_begin(vsb_ctx *c, const int64_t *d_order, int part, int parts)
{
    const int64_t n = c->n;
    if (n <= 0) return VSB_OK;
    VSB_CUDA(cudaMemcpyAsync(c->parents_old, c->parents, sizeof(int64_t) * n,
                             cudaMemcpyDeviceToDevice, c->stream));
    return vsb_launch_find_parents_part(c, d_order, part, parts);
}

int vsb_launch_gravity_ref_finish(vsb_ctx *c, double gc, const int64_t *d_order)
{
    const int64_t n = c->n;
    if (n <= 0) return VSB_OK;
    VSB_TRY(vsb_launch_find_parents_finish(c, d_order));
    const unsigned g = (unsigned)((n + 255) / 256);
    ref_accel_kernel<<<g, 256, 0, c->stream>>>(n, c->mass, (const double2 *)c->pos, c->parents, gc,
                                               (double2 *)c->rel_vel);
    // vel_old snapshot, overflow flag and the slab flags of the fix-up live in scratch
    const int64_t nslabs = (n + 1023) / 1024;
    VSB_TRY(vsb_buf_reserve(c->scratch, 64 + sizeof(double2) * (size_t)c->npad + (size_t)nslabs));
    int *overflow = (int *)c->scratch.p;
    double2 *vel_old = (double2 *)((char *)c->scratch.p + 64);
    unsigned char *slab_flag = (unsigned char *)c->scratch.p + 64 + sizeof(double2) * (size_t)c->npad;
    ref_changed_kernel<<<(unsigned)nslabs, 1024, 0, c->stream>>>(n, c->parents, c->parents_old,
                                                                 slab_flag);
    ref_fixup_kernel<<<1, 1024, 0, c->stream>>>(n, c->mass, c->parents, c->parents_old,
                                                (double2 *)c->rel_vel, slab_flag, nslabs);
    VSB_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), c->stream));
    VSB_CUDA(cudaMemcpyAsync(vel_old, c->vel, sizeof(double2) * n, cudaMemcpyDeviceToDevice,
                             c->stream));
    const int64_t *rank = (const int64_t *)((char *)c->sorted.p + (size_t)c->npad * 24);
    ref_compose_kernel<<<g, 256, 0, c->stream>>>(n, rank, c->parents, (const double2 *)c->rel_vel,
                                                 vel_old, (double2 *)c->vel, (double2 *)c->pos,
                                                 overflow);
    c->launches += 4;
    VSB_CUDA(cudaGetLastError());
    int h_over = 0;
    VSB_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    VSB_CHECK(h_over == 0, VSB_ERR_STATE, "gravity_ref: parent chain deeper than 48 levels");
    return VSB_OK;
}

int vsb_launch_gravity_ref(vsb_ctx *c, double gc, const int64_t *d_order)
{
    VSB_TRY(vsb_launch_gravity_ref_begin(c, d_order, 0, 1));
    return vsb_launch_gravity_ref_finish(c, gc, d_order);
}

int vsb_launch_kepler_basic(vsb_ctx *c, double gc, double coi_coef)
{
    const int64_t n = c->n;
    if (n <= 0) return VSB_OK;
    kepler_basic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        n, c->parents, c->mass, (const double2 *)c->pos, (const double2 *)c->vel, gc, coi_coef,
        c->focus, (double2 *)c->ecc_v, c->semi_major, c->semi_minor, c->periapsis_arg, c->coi);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// kepler_basic over explicit device arrays (the stateless host-array call)
int vsb_launch_kepler_basic_arrays(vsb_ctx *c, int64_t n, const int64_t *parents, const double *mass,
                                   const double *pos, const double *vel, double gc, double coi_coef,
                                   double *focus, double *ecc_v, double *semi_major,
                                   double *semi_minor, double *periapsis_arg, double *coi)
{
    if (n <= 0) return VSB_OK;
    kepler_basic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        n, parents, mass, (const double2 *)pos, (const double2 *)vel, gc, coi_coef, focus,
        (double2 *)ecc_v, semi_major, semi_minor, periapsis_arg, coi);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// largest is written to ((int64_t*)c->scratch.p)[0]
int vsb_launch_simplified_orbit_coi(vsb_ctx *c, double gc, double coi_coef, const int64_t *d_order)
{
    const int64_t n = c->n, npad = c->npad;
    if (n <= 0) return VSB_OK;
    VSB_TRY(vsb_buf_reserve(c->scratch, 64 + (size_t)npad * 16));
    int64_t *d_largest = (int64_t *)c->scratch.p;
    int *best = (int *)((char *)c->scratch.p + 64);
    double *scoi = (double *)((char *)c->scratch.p + 64 + (size_t)npad * 4);
    argmax_kernel<<<1, 1024, 0, c->stream>>>(n, c->mass, d_largest);
    soc_init_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(npad, best, scoi);
    c->launches += 2;
    int64_t h_largest = 0;
    VSB_CUDA(cudaMemcpyAsync(&h_largest, d_largest, sizeof(int64_t), cudaMemcpyDeviceToHost,
                             c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    // From 4096 bodies up: global fixed-point rounds over the O(N) find_parents candidate pass
    // (vsb_parents_grid.cu) -- the same answer as the sweep below, which stays as the path for
    // small systems and for hierarchies nested deeper than the round limit.
    // VSB_PARENTS=scan|grid overrides.
    const char *env = getenv("VSB_PARENTS");
    const bool grid = env ? env[0] == 'g' : n >= 4096;
    if (grid) {
        int converged = 0;
        const char *lim = getenv("VSB_SOC_ROUNDS");      // tests: force the fall-back
        VSB_TRY(vsb_launch_soc_fixed_point(c, d_order, h_largest, gc, coi_coef, lim ? atoi(lim) : 48,
                                           &converged, &c->soc_rounds));
        if (converged) return VSB_OK;
        soc_init_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(npad, best, scoi);
        c->launches += 1;
    }
    for (int64_t r0 = 0; r0 < n; r0 += SOC_T) {
        const int cnt = (int)((n - r0) < SOC_T ? (n - r0) : SOC_T);
        const int64_t rest = n - r0 - cnt;
        soc_slab_kernel<<<1, SOC_T, 0, c->stream>>>(n, r0, d_order, h_largest, c->mass,
                                                    (const double2 *)c->pos,
                                                    (const double2 *)c->vel, gc, coi_coef, best,
                                                    scoi, c->coi);
        c->launches += 1;
        if (rest > 0) {
            soc_update_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, c->stream>>>(
                n, r0, cnt, d_order, (const double2 *)c->pos, scoi, best);
            c->launches += 1;
        }
    }
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_pad_tail(vsb_ctx *c)
{
    const int64_t cnt = c->npad - c->n;
    if (cnt <= 0) return VSB_OK;
    pad_tail_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(arrays_of(c), c->n,
                                                                         c->npad);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_delete_row(vsb_ctx *c, int64_t row)
{
    const int64_t n = c->n;
    const int64_t moves = n - 1 - row;  // rows that shift down
    if (moves > 0) {
        const int64_t ctas = (moves + DEL_T - 1) / DEL_T;
        VSB_TRY(vsb_buf_reserve(c->scratch, sizeof(double) * 19 * (size_t)ctas + 64));
        double *edge = (double *)c->scratch.p;
        delete_edges_kernel<<<(unsigned)((ctas + 255) / 256), 256, 0, c->stream>>>(arrays_of(c),
                                                                                  row, n, edge);
        delete_move_kernel<<<(unsigned)ctas, DEL_T, 0, c->stream>>>(arrays_of(c), row, n, edge);
        c->launches += 2;
    }
    c->n = n - 1;
    if (c->row1 > c->n) c->row1 = c->n;
    if (c->row0 > c->n) c->row0 = c->n;
    // the vacated last row becomes padding again
    pad_tail_kernel<<<1, 32, 0, c->stream>>>(arrays_of(c), c->n, c->n + 1);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_swap_rows(vsb_ctx *c, int64_t a, int64_t b)
{
    swap_rows_kernel<<<1, 32, 0, c->stream>>>(arrays_of(c), a, b);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_translate(vsb_ctx *c, int64_t ref_row)
{
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    VSB_CUDA(cudaMemcpyAsync(c->scratch.p, c->pos + 2 * ref_row, sizeof(double2),
                             cudaMemcpyDeviceToDevice, c->stream));
    translate_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(
        c->n, (double2 *)c->pos, (const double2 *)c->scratch.p);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_editor_curve(vsb_ctx *c, int64_t P, const double *d_tabs, double *d_out)
{
    const int64_t n = c->n;
    if (n <= 0 || P <= 0) return VSB_OK;
    const int64_t threads = n * 32;
    editor_curve_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(
        n, P, (const double2 *)c->ecc_v, c->semi_major, c->semi_minor, c->focus, c->periapsis_arg,
        c->parents, (const double2 *)c->pos, d_tabs, d_out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// Fused small-N frame; hdr (4 x int64) lands in c->scratch.p.  n <= 1024 only.
int vsb_launch_editor_frame_small(vsb_ctx *c, double gc, double coi_coef, const int64_t *d_order,
                                  int max_ticks, int do_kepler_basic, int do_collision,
                                  double *d_packed, long long *d_flag, long long seq)
{
    const int n = (int)c->n;
    VSB_CHECK(n >= 1 && n <= FR_MAX, VSB_ERR_ARG, "editor_frame_small: n=%d", n);
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    const size_t smem = (size_t)n * (4 * sizeof(double2) + 3 * sizeof(double) + 4 * sizeof(int));
    if (smem > 48 * 1024)
        VSB_CUDA(cudaFuncSetAttribute(editor_frame_small_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = (n + 31) / 32 * 32;
    editor_frame_small_kernel<<<1, threads, smem, c->stream>>>(
        n, max_ticks, gc, coi_coef, do_kepler_basic, do_collision, d_order, c->mass,
        (double2 *)c->pos,
        (double2 *)c->vel, (double2 *)c->rel_vel, c->coi, c->rad, c->parents, c->focus,
        (double2 *)c->ecc_v, c->semi_major, c->semi_minor, c->periapsis_arg,
        (int64_t *)c->scratch.p, d_packed, d_flag, seq);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// check_collision of a small system in one launch; (del, add) in c->scratch.p and, with d_flag, in
// mapped host memory (see collision_small_kernel).  n <= 1024 only.
int vsb_launch_check_collision_small(vsb_ctx *c, long long *d_flag, long long seq)
{
    const int n = (int)c->n;
    VSB_CHECK(n >= 1 && n <= FR_MAX, VSB_ERR_ARG, "check_collision_small: n=%d", n);
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    const int threads = (n + 31) / 32 * 32;
    collision_small_kernel<<<1, threads, (size_t)n * (sizeof(double2) + sizeof(double)), c->stream>>>(
        n, (const double2 *)c->pos, c->rad, c->mass, (int64_t *)c->scratch.p, d_flag, seq);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// Packs `nfields` body fields (ids of enum vsb_body_field) into d_out, concatenated.
int vsb_launch_pack_fields(vsb_ctx *c, const int *fields, const void *const *ptrs,
                           const int *widths, int nfields, double *d_out, int64_t *words_out,
                           long long *d_flag, long long seq, int64_t rows)
{
    (void)fields;
    const int64_t n = rows >= 0 ? rows : c->n;     // rows of every field (default: the bodies)
    pack_desc d;
    d.count = nfields;
    int64_t maxw = 0, total = 0;
    for (int f = 0; f < nfields; ++f) {
        d.src[f] = (const double *)ptrs[f];
        d.width[f] = widths[f];
        total += n * widths[f];
        if (n * widths[f] > maxw) maxw = n * widths[f];
    }
    *words_out = total;
    if (maxw == 0) return VSB_OK;
    if (d_flag) {
        pack_fields_small_kernel<<<1, 1024, 0, c->stream>>>(d, n, d_out, d_flag, seq);
        c->launches += 1;
        VSB_CUDA(cudaGetLastError());
        return VSB_OK;
    }
    pack_fields_kernel<<<(unsigned)((maxw + 255) / 256), 256, 0, c->stream>>>(d, n, d_out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
