// vsb_kepler.cu -- game-mode Keplerian propagation and orbit curves (sm_100a):
//   C1  phys_body.Physics.move            phys_body.py:323-346
//   C2  solve_kepler_ell (ENRKE)          enhanced_kepler_solver.py:18-69
//   C3  newton_root_kepler_hyp            phys_shared.py:114-122
//   C4  phys_vessel.Physics.move          phys_vessel.py:477-494 (+ orb2xy phys_shared.py:187-199)
//   C5  curve_points                      phys_shared.py:202-216
//   C6  curve_move_to                     phys_shared.py:219-222
// Structure-of-arrays elements, one thread per object for the solve (coalesced
// 8-byte loads), one warp per object for the curve (each warp store covers 512
// contiguous bytes, streaming .cs stores).  The moved curve is generated
// straight from the elements (write-only: 16*P bytes per object per tick instead
// of the reference's read-modify-write of two (N,P,2) arrays).  HBM-bound.
// Compiled with -fmad=false so x*cos_p - y*sin_p etc. round exactly as in numpy.
#include "vsb_common.cuh"

namespace {

constexpr double PI_D = 3.141592653589793;

// C2.  The `ecc > 0.99 and mr < 0.0045` branch of the reference updates its
// bisection midpoint OUTSIDE the loop: f stays 0.154, so the first
// pass moves one bracket end to f and every later pass repeats the same
// assignment and the same (false or true) exit test -- two passes reproduce all
// 1000 exactly.
__device__ __forceinline__ double solve_kepler_ell(double ecc, double ma, double tol)
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
    double sn, cs;
    sincos(eapp, &sn, &cs);
    double fpp = ecc * sn;
    double fppp = ecc * cs;
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
        sincos(eapp, &sn, &cs);
        fp = 1 - ecc * cs;
        delta = (mr - eapp + ecc * sn) / fp;
    }
    eapp = eapp + delta;
    return ma + flip * (mr - eapp);
}

// C3 (negated convention M = E - e sinh E).
__device__ __forceinline__ double newton_root_kepler_hyp(double ecc, double ma, double ea_guess)
{
    double ea = ea_guess;
    for (int i = 0; i < 50; ++i) {
        double delta_x = (ea - ecc * sinh(ea) - ma) / (1.0 - ecc * cosh(ea));
        ea -= delta_x;
        if (fabs(delta_x) < 1e-10) return ea;
    }
    return ea;
}

struct kep_view {
    const double *a, *b, *f, *ecc, *pea, *nmot, *dr;
    double *ma, *ea;
    double2 *pos;
    const int64_t *ref;
};

// relative position of the object on its conic, rotated by (pea - pi)
__device__ __forceinline__ double2 conic_rel(double a, double b, double f, double ecc, double pea,
                                             double ea)
{
    double xn, yn;
    if (ecc < 1) {
        double s, c;
        sincos(ea, &s, &c);
        xn = a * c - f;
        yn = b * s;
    } else {
        xn = a * cosh(ea) - f;
        yn = b * sinh(ea);
    }
    double s, c;
    sincos(pea - PI_D, &s, &c);
    return make_double2(xn * c - yn * s, xn * s + yn * c);
}

// C4 per object.  Returns new (ma, ea, pos).
__device__ __forceinline__ void vessel_move_one(double a, double b, double f, double ecc,
                                                double pea, double nmot, double dr, double warp,
                                                double2 refp, double &ma, double &ea, double2 &pos)
{
    double hyp = ecc > 1 ? -1.0 : 1.0;
    double m = ma + dr * nmot * warp * hyp;
    if (ecc < 1 && m > 2 * PI_D) m = m - 2 * PI_D;
    if (ecc < 1 && m < 0) m = m + 2 * PI_D;
    double e = ecc < 1 ? solve_kepler_ell(ecc, m, 1e-10) : newton_root_kepler_hyp(ecc, m, ea);
    double2 p = make_double2(0.0, 0.0);
    if (e != 0.0) {  // orb2xy returns (0,0) for ea == 0
        double2 r = conic_rel(a, b, f, ecc, pea, e);
        p.x = r.x + refp.x;
        p.y = r.y + refp.y;
    }
    ma = m;
    ea = e;
    pos = p;
}

__global__ void __launch_bounds__(256)
vessel_move_kernel(kep_view v, int64_t n, double warp, const double2 *__restrict__ ref_pos)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ma = v.ma[i], ea = v.ea[i];
    double2 pos;
    vessel_move_one(v.a[i], v.b[i], v.f[i], v.ecc[i], v.pea[i], v.nmot[i], v.dr[i], warp,
                    ref_pos[v.ref[i]], ma, ea, pos);
    v.ma[i] = ma;
    v.ea[i] = ea;
    v.pos[i] = pos;
}

// C1 part 1: mean-motion update (wraps every body), solve, relative position.
__global__ void __launch_bounds__(256)
body_solve_kernel(kep_view v, int64_t n, double warp, double2 *__restrict__ pr)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ecc = v.ecc[i];
    double m = v.ma[i] + v.dr[i] * v.nmot[i] * warp;
    if (m > 2 * PI_D) m = m - 2 * PI_D;
    if (m < 0) m = m + 2 * PI_D;
    double e = ecc < 1 ? solve_kepler_ell(ecc, m, 1e-10) : newton_root_kepler_hyp(ecc, m, v.ea[i]);
    v.ma[i] = m;
    v.ea[i] = e;
    pr[i] = conic_rel(v.a[i], v.b[i], v.f[i], ecc, v.pea[i], e);
}

// C1 part 2, one launch per dependency level: pos[body] = pos[ref[body]] + pr.
// Bit 62 of the entry says whether the ref was moved earlier in this tick (read
// the new position) or comes later in the reference's loop (read last tick's).
__global__ void body_compose_kernel(const int64_t *__restrict__ lvl_idx, int64_t count,
                                    const int64_t *__restrict__ ref,
                                    const double2 *__restrict__ pr,
                                    const double2 *__restrict__ pos_old, double2 *pos)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int64_t e = lvl_idx[k];
    int64_t body = e & 0x3fffffffffffffffll;
    bool dep_new = (e >> 62) & 1;
    int64_t r = ref[body];
    double2 base = dep_new ? pos[r] : pos_old[r];
    double2 d = pr[body];
    pos[body] = make_double2(base.x + d.x, base.y + d.y);
}

// tables: ell[k] = (cos t_k, sin t_k), hyp[k] = (cosh t_k, sinh t_k)
__global__ void tables_kernel(const double *__restrict__ t, int64_t P, double2 *tab)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    double s, c;
    sincos(t[k], &s, &c);
    tab[k] = make_double2(c, s);
    tab[P + k] = make_double2(cosh(t[k]), sinh(t[k]));
}

// one object's curve by one warp: C5 then C6, out[k] = (rot(x,y) + focus) + ref_pos
__device__ __forceinline__ void curve_by_warp(int lane, int64_t P, const double2 *__restrict__ tb,
                                              double ax, double by, double cp, double sp, double fx,
                                              double fy, double bx, double byy, bool translate,
                                              double2 *__restrict__ out)
{
    for (int64_t k = lane; k < P; k += 32) {
        double2 t = tb[k];
        double x = ax * t.x;
        double y = by * t.y;
        double xr = x * cp - y * sp;
        double yr = y * cp + x * sp;
        if (translate) {
            xr = xr + fx + bx;
            yr = yr + fy + byy;
        }
        __stcs(out + k, make_double2(xr, yr));
    }
}

// C5(+C6) for objects [first, first+count): warp per object, grid-stride.
// translate=0 gives the reference's relative `curves`, 1 the moved `curves_mov`.
__global__ void __launch_bounds__(256)
curves_kernel(kep_view v, int64_t first, int64_t count, int64_t P, const double2 *__restrict__ tab,
              const double2 *__restrict__ ref_pos, int translate, double2 *__restrict__ out)
{
    extern __shared__ double2 s_tab[];
    for (int64_t k = threadIdx.x; k < 2 * P; k += blockDim.x) s_tab[k] = tab[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < count; w += warps) {
        const int64_t i = first + w;
        const double ecc = v.ecc[i], a = v.a[i], b = v.b[i], pea = v.pea[i];
        double sp, cp;
        sincos(pea, &sp, &cp);
        double fx = 0, fy = 0, bx = 0, by = 0;
        if (translate) {
            double f = v.f[i];
            fx = f * cp;
            fy = f * sp;
            double2 rp = ref_pos[v.ref[i]];
            bx = rp.x;
            by = rp.y;
        }
        const bool ell = ecc < 1;
        curve_by_warp(lane, P, ell ? s_tab : s_tab + P, ell ? a : -a, b, cp, sp, fx, fy, bx, by,
                      translate != 0, out + w * P);
    }
}

// One game tick for VESSEL semantics in a single kernel: every lane moves one
// object (C4 incl. C2/C3), then the warp writes the 32 moved curves (C5+C6).
__global__ void __launch_bounds__(256)
vessel_tick_kernel(kep_view v, int64_t n, double warp, int64_t P, const double2 *__restrict__ tab,
                   const double2 *__restrict__ ref_pos, double2 *__restrict__ out)
{
    extern __shared__ double2 s_tab[];
    for (int64_t k = threadIdx.x; k < 2 * P; k += blockDim.x) s_tab[k] = tab[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t groups = (n + 31) >> 5;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < groups; g += warps) {
        const int64_t i = g * 32 + lane;
        double a = 0, b = 0, cp = 1, sp = 0, fx = 0, fy = 0, bx = 0, by = 0;
        int ell = 1;
        if (i < n) {
            const double ecc = v.ecc[i], f = v.f[i], pea = v.pea[i];
            a = v.a[i];
            b = v.b[i];
            const double2 rp = ref_pos[v.ref[i]];
            double ma = v.ma[i], ea = v.ea[i];
            double2 pos;
            vessel_move_one(a, b, f, ecc, pea, v.nmot[i], v.dr[i], warp, rp, ma, ea, pos);
            v.ma[i] = ma;
            v.ea[i] = ea;
            v.pos[i] = pos;
            sincos(pea, &sp, &cp);
            fx = f * cp;
            fy = f * sp;
            bx = rp.x;
            by = rp.y;
            ell = ecc < 1;
            if (!ell) a = -a;
        }
        const int cnt = (int)((n - g * 32) < 32 ? (n - g * 32) : 32);
        for (int o = 0; o < cnt; ++o) {
            const double oa = __shfl_sync(0xffffffffu, a, o), ob = __shfl_sync(0xffffffffu, b, o);
            const double ocp = __shfl_sync(0xffffffffu, cp, o), osp = __shfl_sync(0xffffffffu, sp, o);
            const double ofx = __shfl_sync(0xffffffffu, fx, o), ofy = __shfl_sync(0xffffffffu, fy, o);
            const double obx = __shfl_sync(0xffffffffu, bx, o), oby = __shfl_sync(0xffffffffu, by, o);
            const int oell = __shfl_sync(0xffffffffu, ell, o);
            curve_by_warp(lane, P, oell ? s_tab : s_tab + P, oa, ob, ocp, osp, ofx, ofy, obx, oby,
                          true, out + (g * 32 + o) * P);
        }
    }
}

__global__ void solve_ell_kernel(int64_t n, const double *__restrict__ ecc,
                                 const double *__restrict__ ma, double tol, double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = solve_kepler_ell(ecc[i], ma[i], tol);
}

__global__ void solve_hyp_kernel(int64_t n, const double *__restrict__ ecc,
                                 const double *__restrict__ ma, const double *__restrict__ guess,
                                 double *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = newton_root_kepler_hyp(ecc[i], ma[i], guess[i]);
}

// C6 as the reference states it: out = (curves + focus) + body_pos[ref]; warp per object.
__global__ void __launch_bounds__(256)
curve_move_to_kernel(int64_t n, int64_t P, const double2 *__restrict__ curves,
                     const double2 *__restrict__ body_pos, const int64_t *__restrict__ ref,
                     const double *__restrict__ f, const double *__restrict__ pea,
                     double2 *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        double s, c;
        sincos(pea[i], &s, &c);
        const double fx = f[i] * c, fy = f[i] * s;
        const double2 bp = body_pos[ref[i]];
        for (int64_t k = lane; k < P; k += 32) {
            double2 p = curves[i * P + k];
            __stcs(out + i * P + k, make_double2(p.x + fx + bp.x, p.y + fy + bp.y));
        }
    }
}

kep_view view_of(vsb_kepler_state &k)
{
    kep_view v;
    v.a = k.a; v.b = k.b; v.f = k.f; v.ecc = k.ecc; v.pea = k.pea; v.nmot = k.nmot; v.dr = k.dr;
    v.ma = k.ma; v.ea = k.ea; v.pos = (double2 *)k.pos; v.ref = k.ref;
    return v;
}

int curve_grid(vsb_ctx *c, int64_t warps_needed)
{
    int64_t ctas = (warps_needed + 7) / 8;
    int64_t cap = (int64_t)c->sm_count * 8;  // 8 CTAs x 8 warps = 64 resident warps per SM
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

}  // namespace

int vsb_kepler_tables(vsb_ctx *c, vsb_kepler_state &k, const double *d_t)
{
    tables_kernel<<<(unsigned)((k.P + 255) / 256), 256, 0, c->stream>>>(d_t, k.P, (double2 *)k.tab);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

static int set_curve_smem(const void *fn, size_t bytes)
{
    if (bytes > 48 * 1024)
        VSB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return VSB_OK;
}

int vsb_launch_kepler_move(vsb_ctx *c, vsb_kepler_state &k, double warp)
{
    const int64_t n = k.n;
    if (n <= 0) return VSB_OK;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (k.kind == VSB_KIND_VESSEL) {
        vessel_move_kernel<<<g, 256, 0, c->stream>>>(view_of(k), n, warp,
                                                     (const double2 *)k.ref_pos);
        c->launches += 1;
    } else {
        VSB_CHECK(k.nlevels > 0, VSB_ERR_STATE,
                  "kepler_move(BODY): vsb_kepler_set_order must be called after load");
        VSB_TRY(vsb_buf_reserve(c->scratch, sizeof(double2) * (size_t)n));
        double2 *pos_old = (double2 *)c->scratch.p;
        VSB_CUDA(cudaMemcpyAsync(pos_old, k.pos, sizeof(double2) * n, cudaMemcpyDeviceToDevice,
                                 c->stream));
        body_solve_kernel<<<g, 256, 0, c->stream>>>(view_of(k), n, warp, (double2 *)k.pr);
        c->launches += 1;
        for (int64_t l = 0; l < k.nlevels; ++l) {
            const int64_t cnt = k.lvl_off[l + 1] - k.lvl_off[l];
            if (cnt <= 0) continue;
            body_compose_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(
                k.lvl_idx + k.lvl_off[l], cnt, k.ref, (const double2 *)k.pr, pos_old,
                (double2 *)k.pos);
            c->launches += 1;
        }
    }
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// fused_move: 0 = curves only (translate per `translate`), 1 = vessel tick (move + curves)
int vsb_launch_kepler_curves(vsb_ctx *c, vsb_kepler_state &k, int64_t first, int64_t count,
                             double *d_out, int translate)
{
    if (count <= 0 || k.P <= 0) return VSB_OK;
    const size_t smem = sizeof(double2) * 2 * (size_t)k.P;
    VSB_CHECK(smem <= 200 * 1024, VSB_ERR_ARG, "curve_points P=%lld too large for shared tables",
              (long long)k.P);
    VSB_TRY(set_curve_smem((const void *)curves_kernel, smem));
    const double2 *ref_pos =
        k.kind == VSB_KIND_VESSEL ? (const double2 *)k.ref_pos : (const double2 *)k.pos;
    curves_kernel<<<curve_grid(c, count), 256, smem, c->stream>>>(
        view_of(k), first, count, k.P, (const double2 *)k.tab, ref_pos, translate,
        (double2 *)d_out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_vessel_tick(vsb_ctx *c, vsb_kepler_state &k, double warp, double *d_out)
{
    if (k.n <= 0) return VSB_OK;
    const size_t smem = sizeof(double2) * 2 * (size_t)k.P;
    VSB_CHECK(smem <= 200 * 1024, VSB_ERR_ARG, "curve_points P=%lld too large for shared tables",
              (long long)k.P);
    VSB_TRY(set_curve_smem((const void *)vessel_tick_kernel, smem));
    vessel_tick_kernel<<<curve_grid(c, (k.n + 31) / 32), 256, smem, c->stream>>>(
        view_of(k), k.n, warp, k.P, (const double2 *)k.tab, (const double2 *)k.ref_pos,
        (double2 *)d_out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_solve_ell(vsb_ctx *c, int64_t n, const double *ecc, const double *ma, double tol,
                         double *out)
{
    if (n <= 0) return VSB_OK;
    solve_ell_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, ecc, ma, tol, out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_solve_hyp(vsb_ctx *c, int64_t n, const double *ecc, const double *ma,
                         const double *guess, double *out)
{
    if (n <= 0) return VSB_OK;
    solve_hyp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, ecc, ma, guess, out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_curve_move_to(vsb_ctx *c, int64_t n, int64_t P, const double *curves,
                             const double *body_pos, const int64_t *ref, const double *f,
                             const double *pea, double *out)
{
    if (n <= 0 || P <= 0) return VSB_OK;
    curve_move_to_kernel<<<curve_grid(c, n), 256, 0, c->stream>>>(
        n, P, (const double2 *)curves, (const double2 *)body_pos, ref, f, pea, (double2 *)out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
