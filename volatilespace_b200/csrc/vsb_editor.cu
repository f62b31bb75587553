// vsb_editor.cu -- O(N) pieces of the editor / body-sim tick (sm_100a):
//   A2+A3  Physics.gravity            phys_editor.py:357-378 (gravity_one :51-60)
//   A6     Physics.kepler_basic       phys_editor.py:381-384 (kepler_basic_one :63-83)
//   A7     simplified_orbit_coi       phys_editor.py:325-346
//   B3     del_body / set_root / re-rooting moves  phys_editor.py:250-313, 844-857
//   C7     editor Physics.curve       phys_editor.py:519-544
// Compiled with -fmad=false: the reference runs un-fused (numba fastmath=False).
#include "vsb_common.cuh"

namespace {

constexpr double PI_D = 3.141592653589793;

__device__ __forceinline__ double py_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

// ------------------------------------------------------------------ A2/A3
// rel_vel += gravity_one(body, parents[body], ...) for every body (:361-363).
__global__ void ref_accel_kernel(int64_t n, const double *__restrict__ mass,
                                 const double2 *__restrict__ pos,
                                 const int64_t *__restrict__ parents, double gc,
                                 double2 *__restrict__ rel_vel)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;  // root: acceleration [0,0] (:60)
    int64_t p = parents[i];
    double2 pi = pos[i], pp = pos[p];
    double rx = pp.x - pi.x, ry = pp.y - pi.y;
    double distance = sqrt(rx * rx + ry * ry);
    double mb = mass[i];
    double force = gc * mb * mass[p] / (distance * distance);
    double angle = atan2(ry, rx);
    double acc = force / mb;
    double sn, cs;
    sincos(angle, &sn, &cs);
    double2 rv = rel_vel[i];
    rv.x += acc * cs;
    rv.y += acc * sn;
    rel_vel[i] = rv;
}

// COI enter/leave fix-up (:366-373): sequential in index order because a body may
// read the rel_vel of a body fixed earlier in the same loop.  One CTA walks the
// bodies in 1024-wide slabs; slabs without a parent change cost one barrier.
__global__ void __launch_bounds__(1024)
ref_fixup_kernel(int64_t n, const double *__restrict__ mass, const int64_t *__restrict__ parents,
                 const int64_t *__restrict__ parents_old, double2 *rel_vel)
{
    __shared__ int changed[1024];
    for (int64_t base = 0; base < n; base += 1024) {
        int64_t b = base + threadIdx.x;
        int ch = (b < n && b != 0 && parents[b] != parents_old[b]) ? 1 : 0;
        changed[threadIdx.x] = ch;
        int any = __syncthreads_or(ch);
        if (any && threadIdx.x == 0) {
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
    int64_t p = parents[b];
    double2 pb = pos[b], pp = pos[p], vb = vel[b], vp = vel[p];
    double rpx = pb.x - pp.x, rpy = pb.y - pp.y;
    double rvx = vb.x - vp.x, rvy = vb.y - vp.y;
    double u = gc * mass[p];
    double magp = sqrt(rpx * rpx + rpy * rpy);
    double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / magp));
    double momentum = rpx * rvy - rpy * rvx;
    double sx = rvy, sy = -rvx;
    double ex = (sx * momentum / u) - rpx / magp;
    double ey = (sy * momentum / u) - rpy / magp;
    double ecc = sqrt(ex * ex + ey * ey);
    double pea = py_mod((3 * PI_D / 2) + atan2(-ex, ey), 2 * PI_D);
    double f = a * ecc;
    double bm = sqrt(fabs(f * f - a * a));
    double c = 0.0;
    if (a > 0) c = a * pow(mass[b] / mass[p], coi_coef);
    focus[b] = f;
    ecc_v[b] = make_double2(ex, ey);
    semi_major[b] = a;
    semi_minor[b] = bm;
    periapsis_arg[b] = pea;
    coi[b] = c;
}

// ------------------------------------------------------------------ A7
// Right-looking blocked sweep over ranks (descending mass).  For a slab of T
// consecutive ranks: soc_slab_kernel resolves the slab sequentially inside one
// CTA (rank s needs coi of every earlier rank), then soc_update_kernel tests all
// later ranks against the slab's T finished COIs (grid-wide, "last match wins" =
// max rank).  best[] holds the rank of the current parent candidate (-1: none).
constexpr int SOC_T = 1024;

__device__ __forceinline__ double soc_coi_of(double2 pb, double2 vb, double mb, double2 pp,
                                             double2 vp, double mp, double gc, double coi_coef)
{
    double rpx = pp.x - pb.x, rpy = pp.y - pb.y;
    double rvx = vp.x - vb.x, rvy = vp.y - vb.y;
    double u = gc * mp;
    double a = -1 * u / (2 * ((rvx * rvx + rvy * rvy) / 2 - u / sqrt(rpx * rpx + rpy * rpy)));
    return a > 0 ? a * pow(mb / mp, coi_coef) : 0.0;
}

__global__ void soc_init_kernel(int64_t npad, int *best, double *scoi)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npad) {
        best[k] = -1;
        scoi[k] = 0.0;
    }
}

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
    s_pos[t] = pb;
    int my_best = live ? best[r] : -1;
    const int cnt = (int)((n - r0) < SOC_T ? (n - r0) : SOC_T);
    __syncthreads();
    for (int s = 0; s < cnt; ++s) {
        if (t == s) {
            double c = 0.0;
            if (r > 0) {  // rank 0 keeps coi 0 (loop starts at bodies_sorted[1:], :332)
                int64_t parent = my_best < 0 ? largest : order[my_best];
                c = soc_coi_of(pb, vel[body], mass[body], pos[parent], vel[parent], mass[parent],
                               gc, coi_coef);
            }
            coi[body] = c;
            scoi[r] = c;
            s_coi2[s] = c * c;
        }
        __syncthreads();
        if (t > s && live) {
            // (pos[body]-pos[pot])**2 summed < coi[pot]**2  (:337)
            double dx = pb.x - s_pos[s].x, dy = pb.y - s_pos[s].y;
            if (dx * dx + dy * dy < s_coi2[s]) my_best = (int)(r0 + s);
        }
    }
    if (live) best[r] = my_best;
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

int vsb_launch_gravity_ref(vsb_ctx *c, double gc, const int64_t *d_order)
{
    const int64_t n = c->n;
    if (n <= 0) return VSB_OK;
    VSB_CUDA(cudaMemcpyAsync(c->parents_old, c->parents, sizeof(int64_t) * n,
                             cudaMemcpyDeviceToDevice, c->stream));
    VSB_TRY(vsb_launch_find_parents(c, d_order));
    const unsigned g = (unsigned)((n + 255) / 256);
    ref_accel_kernel<<<g, 256, 0, c->stream>>>(n, c->mass, (const double2 *)c->pos, c->parents, gc,
                                               (double2 *)c->rel_vel);
    ref_fixup_kernel<<<1, 1024, 0, c->stream>>>(n, c->mass, c->parents, c->parents_old,
                                                (double2 *)c->rel_vel);
    // vel_old snapshot + overflow flag live in scratch
    VSB_TRY(vsb_buf_reserve(c->scratch, 64 + sizeof(double2) * (size_t)c->npad));
    int *overflow = (int *)c->scratch.p;
    double2 *vel_old = (double2 *)((char *)c->scratch.p + 64);
    VSB_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), c->stream));
    VSB_CUDA(cudaMemcpyAsync(vel_old, c->vel, sizeof(double2) * n, cudaMemcpyDeviceToDevice,
                             c->stream));
    const int64_t *rank = (const int64_t *)((char *)c->sorted.p + (size_t)c->npad * 24);
    ref_compose_kernel<<<g, 256, 0, c->stream>>>(n, rank, c->parents, (const double2 *)c->rel_vel,
                                                 vel_old, (double2 *)c->vel, (double2 *)c->pos,
                                                 overflow);
    c->launches += 3;
    VSB_CUDA(cudaGetLastError());
    int h_over = 0;
    VSB_CUDA(cudaMemcpyAsync(&h_over, overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    VSB_CHECK(h_over == 0, VSB_ERR_STATE, "gravity_ref: parent chain deeper than 48 levels");
    return VSB_OK;
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
    for (int64_t r0 = 0; r0 < n; r0 += SOC_T) {
        const int cnt = (int)((n - r0) < SOC_T ? (n - r0) : SOC_T);
        soc_slab_kernel<<<1, SOC_T, 0, c->stream>>>(n, r0, d_order, h_largest, c->mass,
                                                    (const double2 *)c->pos,
                                                    (const double2 *)c->vel, gc, coi_coef, best,
                                                    scoi, c->coi);
        c->launches += 1;
        const int64_t rest = n - r0 - cnt;
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
