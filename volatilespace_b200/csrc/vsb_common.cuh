// vsb_common.cuh -- shared device/host helpers for libvsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vsb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvsb200 is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------- errors
void vsb_set_error(const char *fmt, ...);

#define VSB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            vsb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                          cudaGetErrorString(e__));                                 \
            return VSB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define VSB_CHECK(cond, code, ...)                                                  \
    do {                                                                            \
        if (!(cond)) {                                                              \
            vsb_set_error(__VA_ARGS__);                                             \
            return (code);                                                          \
        }                                                                           \
    } while (0)

#define VSB_TRY(expr)                                                               \
    do {                                                                            \
        int r__ = (expr);                                                           \
        if (r__ != VSB_OK) return r__;                                              \
    } while (0)

// ---------------------------------------------------------------- context
// All body arrays are padded to a multiple of VSB_PAD entries; padding rows hold
// mass 0 and a far-away position so pair kernels can run whole tiles unmasked.
#define VSB_PAD 2048
#define VSB_FAR 3.0e150

struct vsb_buf {
    void *p = nullptr;
    size_t cap = 0;  // bytes
};

struct vsb_kepler_state {
    int64_t n = 0, npad = 0, P = 0, nref = 0;
    int kind = 0;  // VSB_KIND_BODY / VSB_KIND_VESSEL
    double *a = nullptr, *b = nullptr, *f = nullptr, *ecc = nullptr, *pea = nullptr,
           *nmot = nullptr, *dr = nullptr, *ma = nullptr, *ea = nullptr, *pos = nullptr;
    int64_t *ref = nullptr;
    double *ref_pos = nullptr;      // (nref,2) parent-body positions (vessel kind)
    int ref_pos_borrowed = 0;       // ref_pos aliases the BODY state's pos (not owned)
    int64_t *h_ref = nullptr;       // host copy of ref (level computation)
    double *tab = nullptr;          // 4*P: cos t, sin t, cosh t, sinh t
    double *pr = nullptr;           // (n,2) relative position scratch (body kind)
    int64_t *lvl_idx = nullptr;     // body kind: bodies grouped by dependency level
    int64_t nlevels = 0;
    int64_t lvl_off[66] = {0};
    // vessel kind: orbit points + per-tick event state, device resident (vsb_vessel_events_load)
    vsb_buf ev;
    int64_t ev_n = 0;
    double *ev_impact = nullptr, *ev_atm = nullptr, *ev_leave = nullptr, *ev_enter = nullptr,
           *ev_prev_ea = nullptr, *ev_pe_d = nullptr, *ev_ap_d = nullptr, *ev_period = nullptr;
    unsigned char *ev_hold = nullptr;
    vsb_buf curves;                 // (n,P,2) moved curves, device resident
    vsb_buf quirk_idx;              // compacted ENRKE quirk-branch indices
    size_t bytes_alloc = 0;
};

struct vsb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;

    // body-sim state
    int64_t n = 0, npad = 0, cap = 0;
    int64_t row0 = 0, row1 = 0;  // target rows owned by this rank
    double *mass = nullptr, *den = nullptr, *rad = nullptr, *coi = nullptr;
    double *pos = nullptr, *vel = nullptr, *rel_vel = nullptr, *acc = nullptr;
    int64_t *parents = nullptr, *parents_old = nullptr;
    double *focus = nullptr, *ecc_v = nullptr, *semi_major = nullptr, *semi_minor = nullptr,
           *periapsis_arg = nullptr;

    // scratch
    vsb_buf part;        // all-pairs partial sums [C][rows][2] (or the symmetric kernel's slabs)
    vsb_buf sym_tab;     // symmetric kernel: task table + Q-slab index
    int64_t sym_tasks_n = -1, sym_qf_off = 0, sym_own_off = 0, sym_ownlist_off = 0;
    int sym_world = 0, sym_rank = 0, sym_ntasks = 0, sym_nslots = 0, sym_S = 0, sym_tail = -1;
    // slab-free symmetric kernel (vsb_gravity_sym_lean.cu): ticket list | expectation table | counters
    vsb_buf lean_tab;
    int64_t lean_tasks_n = -1, lean_exp_off = 0, lean_ctl_off = 0, lean_ctl_bytes = 0;
    int lean_world = 0, lean_rank = 0, lean_ntasks = 0, lean_S = 0, lean_chunk = 0;
    int soc_rounds = 0;     // fixed-point rounds of the last simplified_orbit_coi (diagnostics)
    int gravity_mode = -1;  // -1: auto, 0: row kernel (vsb_gravity.cu), 1: symmetric (vsb_gravity_sym.cu)
    vsb_buf order;       // int64 order (descending mass) + rank
    int64_t order_n = -1; // number of valid entries uploaded to `order`
    vsb_buf sorted;      // sorted-source staging for find_parents
    vsb_buf pgrid;       // uniform grid of the find_parents candidate pass
    vsb_buf scratch;     // generic
    vsb_buf host_stage;  // device staging for the stateless host-buffer calls
    vsb_buf march_stats; // work area of the parallel COI-entry march; starts with 4 x u64 counters
    int march_sequential = 2;  // vsb_set_march_mode: bit 0 the one-thread-per-pair serial march, bit 1 exact trigonometry (default on)
    unsigned long long *d_key = nullptr;   // collision key (device)
    unsigned long long *h_key = nullptr;   // pinned mirror
    // fused small-system frame: results written by the kernel straight into mapped page-locked host
    // memory (record + a sequence word the host spins on): no copy call, no stream synchronise
    double *h_frame = nullptr;             // host address; d_frame is the device alias
    double *d_frame = nullptr;
    long long frame_seq = 0;
    bool mapped_refused = false;           // the mapped allocation failed once: copy paths from then on
    // editor curve(): the parameter tables last uploaded (host copy + device copy), re-sent only when
    // they change (reload_settings)
    vsb_buf curve_tabs;
    void *h_small = nullptr;               // plain host scratch of the small-result path
    size_t h_small_cap = 0;
    double *h_curve_tabs = nullptr;
    int64_t curve_tabs_P = 0;
    double *h_pin = nullptr;               // pinned scratch (small)
    size_t h_pin_bytes = 0;

    vsb_kepler_state kep[2];
};

int vsb_buf_reserve(vsb_buf &b, size_t bytes);
void vsb_buf_free(vsb_buf &b);

static inline int64_t vsb_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init()
{
    // make the inits visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// MUFU.RSQ64H: ~2^-22 relative seed for x^-1/2 (only the high word is significant; the gravity
// kernels replace the zero low word, see vsb_gravity.cu).
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

__device__ __forceinline__ long long warp_max_i64(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// ---- shared tile streamer of the O(N^2) predicate scans (vsb_pairs.cu, vsb_convert.cu)
constexpr int PB_BLOCK = 128;
constexpr int PB_TJ = 512;
constexpr int PB_STAGES = 2;

// Streams tiles [t0, t1) of (pos2[], val[]) through shared memory and calls
// body(sp, sv, first_source_index_of_tile) for each; all threads of the CTA must call it.
template <typename F>
__device__ __forceinline__ void stream_tiles(const double2 *__restrict__ gpos,
                                             const double *__restrict__ gval, int64_t t0,
                                             int64_t t1, F body)
{
    __shared__ alignas(128) double2 spos[PB_STAGES][PB_TJ];
    __shared__ alignas(128) double sval[PB_STAGES][PB_TJ];
    __shared__ alignas(8) uint64_t full[PB_STAGES];
    constexpr uint32_t TILE_BYTES = PB_TJ * (sizeof(double2) + sizeof(double));
    const int tid = threadIdx.x;
    const int ntiles = (int)(t1 - t0);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < PB_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < PB_STAGES; ++s)
            if (s < ntiles) {
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_bulk_g2s(spos[s], gpos + (t0 + s) * PB_TJ, PB_TJ * sizeof(double2), &full[s]);
                tma_bulk_g2s(sval[s], gval + (t0 + s) * PB_TJ, PB_TJ * sizeof(double), &full[s]);
            }
    }
    int consumed = 0;
    int issued = ntiles < PB_STAGES ? ntiles : PB_STAGES;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % PB_STAGES;
        mbar_wait(&full[s], (uint32_t)(t / PB_STAGES) & 1u);
        consumed = t + 1;
        const bool keep_going = body(spos[s], sval[s], (t0 + t) * PB_TJ);
        const int go = __syncthreads_or(keep_going ? 1 : 0);  // also: everyone is done with stage s
        if (!go) break;
        if (t + PB_STAGES < ntiles) {
            issued = t + PB_STAGES + 1;
            if (tid == 0) {
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_bulk_g2s(spos[s], gpos + (t0 + t + PB_STAGES) * PB_TJ,
                             PB_TJ * sizeof(double2), &full[s]);
                tma_bulk_g2s(sval[s], gval + (t0 + t + PB_STAGES) * PB_TJ, PB_TJ * sizeof(double),
                             &full[s]);
            }
        }
    }
    // early exit: bulk copies already in flight must land before the CTA may retire
    for (int t = consumed; t < issued; ++t)
        mbar_wait(&full[t % PB_STAGES], (uint32_t)(t / PB_STAGES) & 1u);
}


#endif  // __CUDACC__

// ---------------------------------------------------------------- internal launchers
// gravity (vsb_gravity.cu)
int vsb_launch_allpairs_accel(vsb_ctx *c, double gc);
int vsb_launch_allpairs_integrate(vsb_ctx *c);
int vsb_launch_allpairs_reduce(vsb_ctx *c, double gc, int integrate);
int vsb_launch_dfma_peak(vsb_ctx *c, int iters, int blocks, float *ms_out);
int vsb_allpairs_chunks(int64_t n, int64_t *chunk_out);
// symmetric kernel (vsb_gravity_sym.cu)
int vsb_launch_allpairs_sym(vsb_ctx *c, double gc, int rank, int world, int integrate);
// the same pairs without partial-sum slabs (vsb_gravity_sym_lean.cu): large body counts
int vsb_launch_allpairs_sym_lean(vsb_ctx *c, double gc, int rank, int world, int integrate);
// the same kernels with ptxas's own instruction schedule (VSB_SYM_SCHED=ptxas; see vsb_gravity_sym.cu)
int vsb_launch_allpairs_sym_ptxas(vsb_ctx *c, double gc, int rank, int world, int integrate);
// pair predicates (vsb_pairs.cu)
int vsb_launch_find_parents(vsb_ctx *c, const int64_t *d_order);
int vsb_launch_find_parents_part(vsb_ctx *c, const int64_t *d_order, int part, int parts);
int vsb_launch_find_parents_finish(vsb_ctx *c, const int64_t *d_order);
// phys_vessel.Physics.curve_segments for the listed vessels (vsb_segments.cu)
int vsb_launch_curve_segments(vsb_ctx *c, const vsb_kepler_state &k, const int64_t *d_idx, int64_t count,
                              double *light, double *dark, double *first, int *type, double *time,
                              double *range);
// sort-based candidate pass (vsb_parents_grid.cu): best[] for all sorted ranks
int vsb_launch_find_parents_grid(vsb_ctx *c, const double2 *spos, const double *scoi2, int *best);
// N1: first loop of to_kepler by fixed-point rounds over the candidate pass
int vsb_launch_to_kepler_fixed_point(vsb_ctx *c, const int64_t *d_order, double gc, double coi_coef,
                                     int max_rounds, int64_t *d_ref, int *converged_out);
// A7 by global fixed-point rounds over the candidate pass (vsb_parents_grid.cu)
int vsb_launch_soc_fixed_point(vsb_ctx *c, const int64_t *d_order, int64_t largest, double gc,
                               double coi_coef, int max_rounds, int *converged_out, int *rounds_out);
int vsb_launch_gravity_ref_begin(vsb_ctx *c, const int64_t *d_order, int part, int parts);
int vsb_launch_gravity_ref_finish(vsb_ctx *c, double gc, const int64_t *d_order);
int vsb_launch_check_collision(vsb_ctx *c);
int vsb_launch_check_collision_part(vsb_ctx *c, int part, int parts);
int vsb_launch_check_collision_finish(vsb_ctx *c);
// sort-based candidate pass (vsb_broadphase.cu); result {del, add, fallback} in c->scratch
int vsb_launch_broadphase(vsb_ctx *c);
// editor step pieces (vsb_editor.cu)
int vsb_launch_gravity_ref(vsb_ctx *c, double gc, const int64_t *d_order);
int vsb_launch_kepler_basic(vsb_ctx *c, double gc, double coi_coef);
int vsb_launch_simplified_orbit_coi(vsb_ctx *c, double gc, double coi_coef,
                                    const int64_t *d_order);
int vsb_launch_delete_row(vsb_ctx *c, int64_t row);
int vsb_launch_swap_rows(vsb_ctx *c, int64_t a, int64_t b);
int vsb_launch_translate(vsb_ctx *c, int64_t ref_row);
int vsb_launch_pad_tail(vsb_ctx *c);
int vsb_launch_editor_frame_small(vsb_ctx *c, double gc, double coi_coef, const int64_t *d_order,
                                  int max_ticks, int do_kepler_basic, int do_collision,
                                  double *d_packed, long long *d_flag = nullptr, long long seq = 0);
int vsb_launch_check_collision_small(vsb_ctx *c, long long *d_flag, long long seq);
int vsb_launch_pack_fields(vsb_ctx *c, const int *fields, const void *const *ptrs,
                           const int *widths, int nfields, double *d_out, int64_t *words_out,
                           long long *d_flag = nullptr, long long seq = 0, int64_t rows = -1);
int vsb_launch_editor_curve(vsb_ctx *c, int64_t P, const double *d_tabs, double *d_out);
// culling / visible subset (vsb_visible.cu)
int vsb_launch_culling_flags(vsb_ctx *c, const double *d_pos, const double *d_radius,
                             double r_uniform, int64_t n, const double *sb4, double zoom,
                             unsigned char *d_flags);
int vsb_launch_orbit_flags(vsb_ctx *c, const int64_t *d_ref, const double *d_size, int64_t n,
                           const unsigned char *d_ref_visible, const double *d_ref_coi, double zoom,
                           double screen_x, unsigned char *d_flags);
int vsb_launch_compact(vsb_ctx *c, const unsigned char *d_flags, int64_t n, int *d_counts,
                       int64_t *d_total, int64_t *d_idx);
int vsb_launch_curves_gather(vsb_ctx *c, const double *d_curves, const int64_t *d_idx,
                             int64_t count, int64_t P, double *d_out);
// body thermal / colour / classification (vsb_body.cu)
int vsb_launch_body_thermal(vsb_ctx *c, int64_t n, const double *consts12, const double *d_in6,
                            double *d_atm_h, const long long *d_base_color, double *d_out4,
                            long long *d_color, double *d_types, int *d_stellar);
// Newton <-> Kepler conversion (vsb_convert.cu)
int vsb_launch_to_kepler(vsb_ctx *c, const int64_t *d_order, double gc, double coi_coef,
                         int64_t *d_ref, double *d_el);
int vsb_launch_to_newton(vsb_ctx *c, const double *d_el, const int64_t *d_ref, double gc,
                         int *err_out);
// COI-crossing numeric kernels (vsb_intersect.cu)
int vsb_launch_quartic(vsb_ctx *c, int64_t n, const double *d_coef, double *d_roots);
int vsb_launch_intersect(vsb_ctx *c, int64_t n, const double *d_in6, double *d_ea, int64_t *d_cnt);
int vsb_launch_predict_enter_coi(vsb_ctx *c, int64_t n, const double *d_vd, const double *d_bd,
                                 const unsigned char *d_valid, double tol, int steps_limit,
                                 double *d_out5, void *d_work, int sequential);
size_t vsb_march_work_bytes(int64_t n);
int vsb_launch_check_warp(vsb_ctx *c, int64_t nv, const double *ecc, const double *dr,
                          const double *ea, const double *ma, const double *n, const double *impact,
                          const double *atm, const double *leave, const double *enter, double warp,
                          unsigned char *hold, unsigned long long *d_key2);
int vsb_launch_enter_coi_service(vsb_ctx *c, int64_t nv, const double *dr, const double *prev_ea,
                                 const double *ea, const double *enter, int pending,
                                 unsigned char *flags);
int vsb_launch_cross_scan(vsb_ctx *c, int64_t nv, const double *ecc, const double *dr,
                          const double *prev_ea, const double *ea, const double *impact,
                          const double *leave, const double *enter, unsigned long long *d_key);
int vsb_launch_cross_apply(vsb_ctx *c, int64_t nev, const int32_t *kind, const double *vessel8,
                           const int64_t *ref, const int64_t *new_ref, const double *body7,
                           double gc, double *out6, int32_t *status);
int vsb_launch_k2v(vsb_ctx *c, int64_t n, const double *rel_pos, const double *a, const double *ecc,
                   const double *pe_arg, const double *u, const double *dr, double *out2);
int vsb_launch_v2k(vsb_ctx *c, int64_t n, const double *rel_pos, const double *rel_vel,
                   const double *u, int failsafe, double *out5);
int vsb_launch_orb2xy(vsb_ctx *c, int64_t n, const double *a, const double *b, const double *f,
                      const double *ecc, const double *pea, const double *ref_pos, const double *ea,
                      double *out2);
int vsb_launch_kepler_basic_arrays(vsb_ctx *c, int64_t n, const int64_t *parents, const double *mass,
                                   const double *pos, const double *vel, double gc, double coi_coef,
                                   double *focus, double *ecc_v, double *semi_major,
                                   double *semi_minor, double *periapsis_arg, double *coi);
int vsb_launch_pack_vessel12(vsb_ctx *c, int64_t n, const double *a, const double *b, const double *f,
                             const double *ecc, const double *ma, const double *ea, const double *pea,
                             const double *nmot, const double *dr, const double *pe_d,
                             const double *ap_d, const double *period, double *out12);
int vsb_march_walk_host(double x0, double c, int64_t kmax, int bounded, double bound, int64_t nq,
                        const int64_t *ks, double *xs, int64_t *K_out, int64_t *nseg_out);
int vsb_launch_points_vessel(vsb_ctx *c, int64_t nsel, const int64_t *d_sel, const double *d_vessel12,
                             const int64_t *d_vref, const double *d_body12, int64_t entered_coi,
                             int64_t left_coi, int only_enter_coi, double *d_impact, double *d_atm,
                             double *d_leave, double *d_ea_eff);
int vsb_launch_points_pairs(vsb_ctx *c, int64_t npairs, const int64_t *d_pair_s,
                            const int64_t *d_pair_body, const int64_t *d_sel,
                            const double *d_vessel12, const int64_t *d_vref, const double *d_body12,
                            int64_t left_coi, const double *d_leave, const double *d_ea_eff,
                            double *d_vd10, double *d_bd10, unsigned char *d_valid);
int vsb_launch_points_select(vsb_ctx *c, int64_t nsel, const int64_t *d_sel, const int64_t *d_offsets,
                             const int64_t *d_pair_body, const unsigned char *d_valid,
                             const double *d_out5, const double *d_ea_eff, const double *d_vessel12,
                             double *d_enter);
// kepler (vsb_kepler.cu)
int vsb_launch_kepler_move(vsb_ctx *c, vsb_kepler_state &k, double warp);
int vsb_launch_kepler_curves(vsb_ctx *c, vsb_kepler_state &k, int64_t first, int64_t count,
                             double *d_out, int translate);
int vsb_launch_solve_ell(vsb_ctx *c, int64_t n, const double *ecc, const double *ma, double tol,
                         double *out);
int vsb_launch_solve_hyp(vsb_ctx *c, int64_t n, const double *ecc, const double *ma,
                         const double *guess, double *out);
int vsb_launch_curve_move_to(vsb_ctx *c, int64_t n, int64_t P, const double *curves,
                             const double *body_pos, const int64_t *ref, const double *f,
                             const double *pea, double *out);
