// vsb_events.cu -- SURVEY 8f row N2: the per-tick vessel event scans of game mode (sm_100a),
// one thread per vessel / event:
//   phys_vessel.Physics.check_warp                 phys_vessel.py:506-560
//   phys_vessel.Physics.predict_enter_coi_service  :744-758 (which vessels get points(v, True))
//   phys_vessel.Physics.cross_coi                  :563-632 (detection + the orbit change)
//   convert.kepler_to_velocity / velocity_to_kepler convert.py:20-131
// The serial loops of the reference become order-preserving reductions: check_warp reports the
// LAST vessel that raises an alarm (its loop overwrites), cross_coi the FIRST that crosses (its
// loop returns).  Compiled with -fmad=false.
#include "vsb_common.cuh"
#include "vsb_orbit_math.cuh"

namespace {

// hold[v]: membership of vessel v in physical_hold.  key_out: max over alarmed vessels of
// (v + 1) * 4 + situation (0 = no alarm); count_out: len(physical_hold) afterwards.
__global__ void __launch_bounds__(256)
check_warp_kernel(int64_t nv, const double *__restrict__ ecc, const double *__restrict__ dr,
                  const double *__restrict__ ea, const double *__restrict__ ma,
                  const double *__restrict__ n, const double *__restrict__ body_impact,
                  const double *__restrict__ body_enter_atm, const double *__restrict__ coi_leave,
                  const double *__restrict__ coi_enter, double warp,
                  unsigned char *__restrict__ hold, unsigned long long *__restrict__ key_out,
                  unsigned long long *__restrict__ count_out)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0;
    unsigned int held = 0;
    if (v < nv) {
        const double all[5] = {body_impact[2 * v], body_enter_atm[2 * v], body_enter_atm[2 * v + 1],
                               coi_enter[6 * v + 1], coi_leave[2 * v]};
        const double e = ecc[v], d = dr[v], ea1 = ea[v];
        unsigned char h = hold[v];
        const int nx = first_intersect(ea1, all, 5, d);
        if (nx >= 0) {
            double ea2;
            if (e < 1) ea2 = pe_solve_kepler_ell(e, ma[v] + n[v] * warp * d * 2, 1e-10);
            else ea2 = pe_newton_hyp(e, ma[v] + n[v] * warp * d * -1 * 2, ea1);
            if (point_between(all[nx], ea1, ea2, d)) {
                const int situation = nx <= 2 ? nx : 3;
                if (nx <= 1) h = 1;
                else if (nx == 2) h = 0;
                key = (unsigned long long)(v + 1) * 4 + situation;
            } else {
                h = 0;
            }
        } else {
            h = 0;
        }
        hold[v] = h;
        held = h;
    }
    // block reduction: max key, sum held
    __shared__ unsigned long long s_key[8];
    __shared__ unsigned int s_cnt[8];
    for (int o = 16; o; o >>= 1) {
        const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
        key = k2 > key ? k2 : key;
        held += __shfl_xor_sync(0xffffffffu, held, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_key[w] = key;
        s_cnt[w] = held;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            key = s_key[i] > key ? s_key[i] : key;
            held += s_cnt[i];
        }
        if (key) atomicMax(key_out, key);
        if (held) atomicAdd(count_out, (unsigned long long)held);
    }
}

__global__ void enter_coi_service_kernel(int64_t nv, const double *__restrict__ dr,
                                         const double *__restrict__ prev_ea,
                                         const double *__restrict__ ea,
                                         const double *__restrict__ coi_enter, int pending,
                                         unsigned char *__restrict__ flags)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const double ce = coi_enter[6 * v];
    unsigned char f = 0;
    if (ce != ce && !pending)
        f = point_between(0.0, prev_ea[v], ea[v], dr[v]) || point_between(PI_D, prev_ea[v], ea[v], dr[v]);
    flags[v] = f;
}

// key_out: min over crossing vessels of v * 4 + kind (1 = enter, 2 = leave); ~0 = none
__global__ void __launch_bounds__(256)
cross_scan_kernel(int64_t nv, const double *__restrict__ ecc, const double *__restrict__ dr,
                  const double *__restrict__ prev_ea, const double *__restrict__ ea,
                  const double *__restrict__ body_impact, const double *__restrict__ coi_leave,
                  const double *__restrict__ coi_enter, unsigned long long *__restrict__ key_out)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = ~0ull;
    if (v < nv) {
        const bool ell = ecc[v] < 1;
        const double all[3] = {body_impact[2 * v + (ell ? 0 : 1)], coi_enter[6 * v + 1],
                               coi_leave[2 * v]};
        const double p = prev_ea[v], d = dr[v];
        const int nx = first_intersect(p, all, 3, d);
        if (nx >= 1 && point_between(all[nx], p, ea[v], d)) key = (unsigned long long)v * 4 + nx;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o);
        key = k2 < key ? k2 : key;
    }
    __shared__ unsigned long long s_key[8];
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) key = s_key[i] < key ? s_key[i] : key;
        if (key != ~0ull) atomicMin(key_out, key);
    }
}

// --- the orbit change of a crossing vessel ---------------------------------------------------
__device__ double ps_compare(double a, double b)
{   // phys_shared.py:35-48
    if (a == b) return 0.0;
    if (a == 0 || b == 0) return 100.0;
    if ((a > 0 && b > 0) || (a < 0 && b < 0)) {
        a = fabs(a);
        b = fabs(b);
        return fabs(a - b) / (a > b ? a : b) * 100;
    }
    return 100;
}

__device__ void kepler_to_velocity(const double *rel_pos, double a, double ecc, double pe_arg,
                                   double u, double dr, double *out2)
{   // convert.py:20-64
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
        if (((xx2 - x1) * (rel_pos[1] - y1)) - ((y2 - y1) * (rel_pos[0] - x1)) < 0) angle += PI_D;
    } else {
        // a is negative; the focus shift uses cosh / sinh of the ANGLE, as the reference does (:45)
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
            if (((xx2 - x1) * (rel_pos[1] - y1)) - ((y2 - y1) * (rel_pos[0] - x1)) < 0) angle += PI_D;
        } else if (PI_D <= pe_arg && pe_arg < 2 * PI_D) {   // the reference's `except ValueError`
            angle += PI_D;
        }
    }
    angle = py_mod(angle, 2 * PI_D);
    const double distance = sqrt(rel_pos[0] * rel_pos[0] + rel_pos[1] * rel_pos[1]);
    const double speed = dr * sqrt(u * ((2 / distance) - (1 / a)));
    out2[0] = speed * cos(angle);
    out2[1] = speed * sin(angle);
}

__device__ void velocity_to_kepler(const double *rel_pos, const double *rel_vel, double u,
                                   int failsafe, double *out5)
{   // convert.py:67-131 -> a, ecc, pe_arg, ma, dr
    const double px = rel_pos[0], py = rel_pos[1];
    double vx = rel_vel[0], vy = rel_vel[1];
    const double mp = sqrt(px * px + py * py);
    const double a = -1 * u / (2 * ((vx * vx + vy * vy) / 2 - u / mp));
    const double mom = px * vy - py * vx;
    {
        const double t = vx;       // rel_vel[0], rel_vel[1] = rel_vel[1], -rel_vel[0]
        vx = vy;
        vy = -t;
    }
    const double ex = (vx * mom / u) - px / mp, ey = (vy * mom / u) - py / mp;
    const double ecc = sqrt(ex * ex + ey * ey);
    double pe_arg = py_mod((3 * PI_D / 2) + atan2(-ex, ey), 2 * PI_D);
    const double dr = copysign(1.0, mom);
    double ta = py_mod(pe_arg - (atan2(py, px) - PI_D), 2 * PI_D);
    if (dr > 0) ta = 2 * PI_D - ta;
    double ea, ma, f = 0, b = 0, np_[2];
    if (ecc < 1) {
        ea = acos((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
        if (PI_D < ta && ta < 2 * PI_D) ea = 2 * PI_D - ea;
        ma = py_mod(ea - ecc * sin(ea), 2 * PI_D);
        if (failsafe) {
            const double new_ea = pe_solve_kepler_ell(ecc, ma, 1e-10);
            f = a * ecc;
            b = sqrt(fabs(f * f - a * a));
            ps_orb2xy(a, b, f, ecc, pe_arg, 0, 0, new_ea, np_);
            if ((ps_compare(px, np_[0]) + ps_compare(py, np_[1])) / 2 > 1) {
                ea = 2 * PI_D - ea;
                ma = py_mod(ea - ecc * sin(ea), 2 * PI_D);
            }
        }
    } else {
        ea = acosh((ecc + cos(ta)) / (1 + (ecc * cos(ta))));
        ma = ecc * sinh(ea) - ea;
        if (failsafe) {
            const double new_ea = pe_newton_hyp(ecc, ma, ma);
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
        {
            const double t = vx;   // undo the swap
            vx = -vy;
            vy = t;
        }
        const double va = py_mod(atan2(vy, vx), 2 * PI_D), pa = py_mod(atan2(py, px), 2 * PI_D);
        if (sin(3 * PI_D / 2 + va - pa) < 0) {
            if (ecc < 1) {
                ea = 2 * PI_D - ea;
                ma = py_mod(ea - ecc * sin(ea), 2 * PI_D);
            } else {
                ea = -ea;
                ma = -ma;
            }
            pe_arg = py_mod(pe_arg + PI_D, 2 * PI_D);
            ps_orb2xy(a, b, f, ecc, pe_arg, 0, 0, ea, np_);
            const double npa = py_mod(atan2(np_[1], np_[0]), 2 * PI_D);
            pe_arg = pe_arg - (npa - pa);
        }
    }
    out5[0] = a;
    out5[1] = ecc;
    out5[2] = pe_arg;
    out5[3] = ma;
    out5[4] = dr;
}

// vessel8 (nev,8) = a b f ecc pea u dr ea_cross; body7 (nb,7) = a ecc pea dr mass pos_x pos_y;
// out6 (nev,6) = a ecc pea ma dr new_ref; status = 1 where the reference returns None
__global__ void cross_apply_kernel(int64_t nev, const int32_t *__restrict__ kind,
                                   const double *__restrict__ vessel8,
                                   const int64_t *__restrict__ ref,
                                   const int64_t *__restrict__ new_ref,
                                   const double *__restrict__ body7, double gc,
                                   double *__restrict__ out6, int32_t *__restrict__ status)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nev) return;
    const double *V = vessel8 + 8 * i;
    const double a = V[0], b = V[1], f = V[2], ecc = V[3], pea = V[4], u = V[5], dr = V[6], eax = V[7];
    const double *R = body7 + 7 * ref[i], *N = body7 + 7 * new_ref[i];
    double *o = out6 + 6 * i;
    if (kind[i] == 1 && ref[i] == new_ref[i]) {
        status[i] = 1;
        for (int k = 0; k < 6; ++k) o[k] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    status[i] = 0;
    const double new_u = gc * N[4];
    double vabs[2], vrel[2], vv[2], rv[2], npos[2], nvel[2], rrel[2];
    ps_orb2xy(a, b, f, ecc, pea, R[5], R[6], eax, vabs);
    vrel[0] = vabs[0] - R[5];
    vrel[1] = vabs[1] - R[6];
    kepler_to_velocity(vrel, a, ecc, pea, u, dr, vv);
    if (kind[i] == 2) {         // leave: up to the parent of the current reference (:577-602)
        rrel[0] = R[5] - N[5];
        rrel[1] = R[6] - N[6];
        kepler_to_velocity(rrel, R[0], R[1], R[2], new_u, R[3], rv);
        npos[0] = vrel[0] + rrel[0];
        npos[1] = vrel[1] + rrel[1];
        nvel[0] = vv[0] + rv[0];
        nvel[1] = vv[1] + rv[1];
        velocity_to_kepler(npos, nvel, new_u, 1, o);
    } else {                    // enter: down to a body orbiting the current reference (:604-629)
        rrel[0] = N[5] - R[5];
        rrel[1] = N[6] - R[6];
        kepler_to_velocity(rrel, N[0], N[1], N[2], new_u, N[3], rv);
        npos[0] = vrel[0] - rrel[0];
        npos[1] = vrel[1] - rrel[1];
        nvel[0] = vv[0] - rv[0];
        nvel[1] = vv[1] - rv[1];
        velocity_to_kepler(npos, nvel, new_u, 2, o);
    }
    o[5] = (double)new_ref[i];
}

// batched convert.kepler_to_velocity / velocity_to_kepler and phys_shared.orb2xy
__global__ void k2v_kernel(int64_t n, const double *__restrict__ rel_pos, const double *__restrict__ a,
                           const double *__restrict__ ecc, const double *__restrict__ pe_arg,
                           const double *__restrict__ u, const double *__restrict__ dr,
                           double *__restrict__ out2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) kepler_to_velocity(rel_pos + 2 * i, a[i], ecc[i], pe_arg[i], u[i], dr[i], out2 + 2 * i);
}

__global__ void v2k_kernel(int64_t n, const double *__restrict__ rel_pos,
                           const double *__restrict__ rel_vel, const double *__restrict__ u,
                           int failsafe, double *__restrict__ out5)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) velocity_to_kepler(rel_pos + 2 * i, rel_vel + 2 * i, u[i], failsafe, out5 + 5 * i);
}

__global__ void orb2xy_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                              const double *__restrict__ f, const double *__restrict__ ecc,
                              const double *__restrict__ pea, const double *__restrict__ ref_pos,
                              const double *__restrict__ ea, double *__restrict__ out2)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        ps_orb2xy(a[i], b[i], f[i], ecc[i], pea[i], ref_pos[2 * i], ref_pos[2 * i + 1], ea[i],
                  out2 + 2 * i);
}

// SoA vessel state -> the (n,12) rows points() works on: a b f ecc ma ea pea n dr pe_d ap_d period
__global__ void pack_vessel12_kernel(int64_t n, const double *__restrict__ a,
                                     const double *__restrict__ b, const double *__restrict__ f,
                                     const double *__restrict__ ecc, const double *__restrict__ ma,
                                     const double *__restrict__ ea, const double *__restrict__ pea,
                                     const double *__restrict__ nmot, const double *__restrict__ dr,
                                     const double *__restrict__ pe_d, const double *__restrict__ ap_d,
                                     const double *__restrict__ period, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *o = out + 12 * i;
    o[0] = a[i]; o[1] = b[i]; o[2] = f[i]; o[3] = ecc[i]; o[4] = ma[i]; o[5] = ea[i];
    o[6] = pea[i]; o[7] = nmot[i]; o[8] = dr[i]; o[9] = pe_d[i]; o[10] = ap_d[i]; o[11] = period[i];
}

}  // namespace

int vsb_launch_k2v(vsb_ctx *c, int64_t n, const double *rel_pos, const double *a, const double *ecc,
                   const double *pe_arg, const double *u, const double *dr, double *out2)
{
    if (n <= 0) return VSB_OK;
    k2v_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(n, rel_pos, a, ecc, pe_arg, u, dr,
                                                                  out2);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_v2k(vsb_ctx *c, int64_t n, const double *rel_pos, const double *rel_vel,
                   const double *u, int failsafe, double *out5)
{
    if (n <= 0) return VSB_OK;
    v2k_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(n, rel_pos, rel_vel, u, failsafe,
                                                                  out5);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_orb2xy(vsb_ctx *c, int64_t n, const double *a, const double *b, const double *f,
                      const double *ecc, const double *pea, const double *ref_pos, const double *ea,
                      double *out2)
{
    if (n <= 0) return VSB_OK;
    orb2xy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, a, b, f, ecc, pea, ref_pos,
                                                                     ea, out2);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_pack_vessel12(vsb_ctx *c, int64_t n, const double *a, const double *b, const double *f,
                             const double *ecc, const double *ma, const double *ea, const double *pea,
                             const double *nmot, const double *dr, const double *pe_d,
                             const double *ap_d, const double *period, double *out12)
{
    if (n <= 0) return VSB_OK;
    pack_vessel12_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        n, a, b, f, ecc, ma, ea, pea, nmot, dr, pe_d, ap_d, period, out12);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_check_warp(vsb_ctx *c, int64_t nv, const double *ecc, const double *dr,
                          const double *ea, const double *ma, const double *n, const double *impact,
                          const double *atm, const double *leave, const double *enter, double warp,
                          unsigned char *hold, unsigned long long *d_key2)
{
    VSB_CUDA(cudaMemsetAsync(d_key2, 0, 16, c->stream));
    if (nv > 0) {
        check_warp_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, c->stream>>>(
            nv, ecc, dr, ea, ma, n, impact, atm, leave, enter, warp, hold, d_key2, d_key2 + 1);
        c->launches += 1;
    }
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_enter_coi_service(vsb_ctx *c, int64_t nv, const double *dr, const double *prev_ea,
                                 const double *ea, const double *enter, int pending,
                                 unsigned char *flags)
{
    if (nv <= 0) return VSB_OK;
    enter_coi_service_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, c->stream>>>(
        nv, dr, prev_ea, ea, enter, pending, flags);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_cross_scan(vsb_ctx *c, int64_t nv, const double *ecc, const double *dr,
                          const double *prev_ea, const double *ea, const double *impact,
                          const double *leave, const double *enter, unsigned long long *d_key)
{
    VSB_CUDA(cudaMemsetAsync(d_key, 0xff, 8, c->stream));
    if (nv > 0) {
        cross_scan_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, c->stream>>>(
            nv, ecc, dr, prev_ea, ea, impact, leave, enter, d_key);
        c->launches += 1;
    }
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_cross_apply(vsb_ctx *c, int64_t nev, const int32_t *kind, const double *vessel8,
                           const int64_t *ref, const int64_t *new_ref, const double *body7,
                           double gc, double *out6, int32_t *status)
{
    if (nev <= 0) return VSB_OK;
    cross_apply_kernel<<<(unsigned)((nev + 63) / 64), 64, 0, c->stream>>>(
        nev, kind, vessel8, ref, new_ref, body7, gc, out6, status);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
