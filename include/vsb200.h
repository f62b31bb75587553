/*
 * vsb200.h -- C ABI of libvsb200.so, the B200 (sm_100a) implementation of
 * VolatileSpace's data-parallel physics hot path.
 *
 * The reference (mzivic7/VolatileSpace v0.5.2, pure Python + numba) has no FFI
 * of its own: its seam is the `Physics` classes in
 * volatilespace/physics/{phys_editor,phys_body,phys_vessel}.py and a few free
 * functions.  Each entry point below names the reference interface
 * (file:line, relative to the reference repo) it stands behind; the ctypes
 * binding a maintainer adds on the reference side is shown in INTEGRATION.md
 * and shipped as volatilespace_b200/_lib.py.
 *
 * Conventions: every function returns VSB_OK (0) or a negative VSB_ERR_* code,
 * with a message available from vsb_last_error(); nothing throws; host pointers
 * are never retained past the call; all floating point is IEEE float64, index
 * arrays are int64, 2-vectors are interleaved (N,2) C-order exactly like the
 * reference's numpy arrays; calls are blocking (the result is visible in the
 * host buffers on return) and expected from one host thread per context.
 * There is no CPU fallback anywhere: without a CUDA device vsb_ctx_create fails.
 */
#ifndef VSB200_H
#define VSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB_ABI_VERSION 1

#define VSB_OK 0
#define VSB_ERR_CUDA (-1)  /* CUDA runtime error, no device, launch failure */
#define VSB_ERR_ARG (-2)   /* bad argument (null pointer, negative size, index out of range) */
#define VSB_ERR_STATE (-3) /* call order violated (e.g. step before upload) */
#define VSB_ERR_NOMEM (-4) /* device or pinned allocation failed */

typedef struct vsb_ctx vsb_ctx;

/* ------------------------------------------------------------------ library */
int vsb_abi_version(void);
const char *vsb_last_error(void);
/* Which instruction schedule of the symmetric pair kernel this library runs: "resched" (the SASS
 * post-pass of volatilespace_b200/sass_resched.py was applied at build time) or "ptxas ..." with
 * the reason in parentheses (post-pass skipped or failed at build time, or VSB_SYM_SCHED=ptxas
 * in the environment).  Same results bit for bit either way; the string is static. */
const char *vsb_sym_schedule(void);
int vsb_device_count(int *count_out);

/* One context per GPU (one per process under torch.distributed). */
int vsb_ctx_create(int device, vsb_ctx **ctx_out);
int vsb_ctx_destroy(vsb_ctx *ctx);
int vsb_ctx_sync(vsb_ctx *ctx);
/* cudaStream_t every kernel of this context is launched on (for event timing /
 * torch.cuda.ExternalStream interop). */
int vsb_ctx_stream(vsb_ctx *ctx, void **cuda_stream_out);
/* Number of kernels this library has launched on the context so far. */
int vsb_ctx_kernel_launches(vsb_ctx *ctx, int64_t *count_out);
/* Measured FP64 FMA peak of this device in FLOP/s (DFMA-chain microbenchmark,
 * 2 flop per DFMA); the roofline denominator for the gravity kernel. */
int vsb_measure_fp64_peak(vsb_ctx *ctx, double *flops_out, double *pipe_ops_out);
/* Pin / unpin a caller-owned host buffer (cudaHostRegister) so the host-buffer entry
 * points below copy at full PCIe rate; optional. */
int vsb_host_register(void *ptr, int64_t bytes);
int vsb_host_unregister(void *ptr);

/* ------------------------------------------------- body-sim state (editor mode)
 * Device mirror of phys_editor.Physics' SoA arrays, phys_editor.py:112-149. */
enum vsb_body_field {
    VSB_F_MASS = 0,      /* (N,)   f64 */
    VSB_F_DEN = 1,       /* (N,)   f64 */
    VSB_F_RAD = 2,       /* (N,)   f64  host-computed np.cbrt radius, phys_editor.py:580-581 */
    VSB_F_COI = 3,       /* (N,)   f64 */
    VSB_F_POS = 4,       /* (N,2)  f64 */
    VSB_F_VEL = 5,       /* (N,2)  f64 */
    VSB_F_REL_VEL = 6,   /* (N,2)  f64 */
    VSB_F_ACC = 7,       /* (N,2)  f64  last all-pairs acceleration (owned rows valid) */
    VSB_F_PARENTS = 8,   /* (N,)   i64 */
    VSB_F_FOCUS = 9,     /* (N,)   f64  kepler_basic outputs, phys_editor.py:381-384 */
    VSB_F_ECC_V = 10,    /* (N,2)  f64 */
    VSB_F_SEMI_MAJOR = 11,
    VSB_F_SEMI_MINOR = 12,
    VSB_F_PERIAPSIS_ARG = 13,
    VSB_F_COUNT_ = 14
};

/* load_system / add_body resize, phys_editor.py:169-201: (re)allocate for n bodies.
 * Contents are undefined until set. */
int vsb_bodies_resize(vsb_ctx *ctx, int64_t n);
int vsb_bodies_count(vsb_ctx *ctx, int64_t *n_out);
/* Copy `count` bodies starting at row `first` host->device / device->host. */
int vsb_bodies_set(vsb_ctx *ctx, int field, const void *src, int64_t first, int64_t count);
int vsb_bodies_get(vsb_ctx *ctx, int field, void *dst, int64_t first, int64_t count);
/* Raw device pointer of a field (zero-copy plumbing for the NCCL all-gather). */
int vsb_bodies_device_ptr(vsb_ctx *ctx, int field, void **dptr_out);
/* Rows [row0,row1) are the target rows this rank computes and integrates
 * (SURVEY 8e row sharding).  Default after resize: [0,n). */
int vsb_bodies_set_rows(vsb_ctx *ctx, int64_t row0, int64_t row1);

/* A1  Physics.find_parents, phys_editor.py:350-354 (find_parent_one :36-48).
 * order_desc_mass = np.argsort(mass)[-1::-1] from the host.
 * Result stays in VSB_F_PARENTS; parents_out may be NULL. */
int vsb_find_parents(vsb_ctx *ctx, const int64_t *order_desc_mass, int64_t *parents_out);

/* A3  Physics.gravity, phys_editor.py:357-378: one tick of the parent/COI model
 * (find_parents, gravity_one :51-60, rel_vel fix-up, vel compose, pos += vel).
 * order_desc_mass may be NULL to reuse the order uploaded by the previous call. */
int vsb_gravity_ref(vsb_ctx *ctx, double gc, const int64_t *order_desc_mass);
/* The same tick with the O(N^2) find_parents scan split over ranks: rank `part` of
 * `parts` scans the row blocks part, part+parts, ... of the mass-sorted bodies (asynchronous);
 * the int32 array vsb_find_parents_best_ptr returns (one entry per body in sorted order: sorted
 * rank of the smallest enclosing COI, -1 = none; rows of other ranks hold -1) is reduced with
 * MAX over the ranks, and every rank calls vsb_gravity_ref_finish (parents, gravity_one,
 * integrator -- all O(N)).  order may be NULL in begin when it is already resident. */
int vsb_gravity_ref_begin_async(vsb_ctx *ctx, const int64_t *order_desc_mass, int part, int parts);
int vsb_find_parents_best_ptr(vsb_ctx *ctx, void **dptr_out);
int vsb_gravity_ref_finish(vsb_ctx *ctx, double gc);

/* A4  documented all-pairs model, documentation/orbital-physics/
 * 01-complex-orbit-model.txt:4-12, pair formula of gravity_one phys_editor.py:51-60.
 * accel: VSB_F_ACC[rows] = gc * sum_{j!=i} m_j (r_j-r_i)/|r_j-r_i|^3 (no softening).
 * integrate: v += a ; P += v on the owned rows.  step = nsteps x (accel, integrate),
 * single GPU (rows must be [0,n)). */
int vsb_allpairs_accel(vsb_ctx *ctx, double gc);
int vsb_allpairs_integrate(vsb_ctx *ctx);
int vsb_allpairs_step(vsb_ctx *ctx, double gc, int64_t nsteps);
/* One tick on the owned rows (accel + reduce + integrate), enqueued on the context
 * stream WITHOUT a host sync: the multi-GPU driver all-gathers VSB_F_POS after it. */
int vsb_allpairs_tick_async(vsb_ctx *ctx, double gc);
/* Acceleration only (accel + reduce into VSB_F_ACC), enqueued without a host sync. */
int vsb_allpairs_accel_async(vsb_ctx *ctx, double gc);
/* Kernel choice for the all-pairs entry points when the context owns every row:
 * -1 = automatic (default: symmetric from 8192 bodies up),
 * 0 = row kernel (13 FP64 instr per directed pair; bit-identical under any row sharding),
 * 1 = symmetric kernel (each unordered pair evaluated once and applied to both bodies,
 *     8 FP64 instr per directed pair; deterministic run to run, last-bit differences from
 *     mode 0).  Sharded rows always use the row kernel. */
int vsb_set_gravity_mode(vsb_ctx *ctx, int mode);
/* Multi-GPU form of the symmetric kernel: rank `rank` of `world` evaluates its share of the
 * unordered block pairs and leaves its PARTIAL accelerations for all bodies in VSB_F_ACC
 * (asynchronous); the caller all-reduces (sum) VSB_F_ACC over the ranks and then calls
 * vsb_allpairs_integrate_async on every rank (all rows owned). */
int vsb_allpairs_sym_partial_async(vsb_ctx *ctx, double gc, int rank, int world);
int vsb_allpairs_integrate_async(vsb_ctx *ctx);

/* Editor frame driver, editor.py:1503-1523: up to max_ticks x (gravity(); check_collision())
 * -- body() only changes `rad` after a merge, which the host applies -- stopping after the
 * tick in which a collision is found: *ticks_done ticks were executed and (*body_del_out,
 * *body_add_out) is that tick's pair ((-1,-1): none).  When every tick ran undisturbed and
 * do_kepler_basic != 0, kepler_basic (phys_editor.py:381-384) runs at the end.  Up to 1024
 * bodies the whole call is ONE kernel launch (one CTA, state in shared memory) and one
 * 32-byte read-back; results are bit-identical to the per-call path.
 * order_desc_mass may be NULL to reuse the order uploaded by the previous call.  frame_out
 * (optional) receives the frame record in one copy: 4 int64 {ticks_done, del, add, 0} followed
 * by pos, vel, rel_vel (2n each), parents (n, int64), focus (n), ecc_v (2n), semi_major,
 * semi_minor, periapsis_arg, coi (n each) -- (4 + 14 n) 8-byte words. */
int vsb_editor_ticks(vsb_ctx *ctx, double gc, double coi_coef, const int64_t *order_desc_mass,
                     int64_t max_ticks, int do_kepler_basic, int64_t *ticks_done,
                     int64_t *body_del_out, int64_t *body_add_out, void *frame_out);
/* Download several fields with one device->host copy: dst receives the fields concatenated
 * in the order given (each n x width 8-byte words; VSB_F_PARENTS as int64). */
int vsb_bodies_get_many(vsb_ctx *ctx, const int *fields, int nfields, void *dst);

/* A6  Physics.kepler_basic, phys_editor.py:381-384 (kepler_basic_one :63-83):
 * fills FOCUS, ECC_V, SEMI_MAJOR, SEMI_MINOR, PERIAPSIS_ARG and COI. */
int vsb_kepler_basic(vsb_ctx *ctx, double gc, double coi_coef);

/* A7  Physics.simplified_orbit_coi, phys_editor.py:325-346: recompute every COI
 * hierarchically in descending-mass order; largest_out = argmax(mass). */
int vsb_simplified_orbit_coi(vsb_ctx *ctx, double gc, double coi_coef,
                             const int64_t *order_desc_mass, int64_t *largest_out);

/* B1  Physics.check_collision, phys_editor.py:547-551 (check_collision_one :86-97):
 * lexicographically smallest colliding pair, returned as (body_del, body_add);
 * (-1,-1) stands for the reference's (None, None).  Bit-exact predicate. */
int vsb_check_collision(vsb_ctx *ctx, int64_t *body_del_out, int64_t *body_add_out);
/* The exhaustive scan split over ranks (SURVEY 8e "collision detect"): rank `part` of `parts`
 * scans the row blocks part, part+parts, ... (asynchronous) and leaves its local minimum key
 * (i << 32 | j; all ones = no collision) in the device word vsb_collision_key_ptr returns; the
 * caller reduces that word with MIN over the ranks (one 64-bit all-reduce) and every rank then
 * calls vsb_check_collision_finish, which applies the reference's del/add rule
 * (phys_editor.py:93-97) to the winning pair. */
int vsb_check_collision_part_async(vsb_ctx *ctx, int part, int parts);
int vsb_collision_key_ptr(vsb_ctx *ctx, void **dptr_out);
int vsb_check_collision_finish(vsb_ctx *ctx, int64_t *del_out, int64_t *add_out);

/* B3  structural pieces of del_body :250-290 / set_root :293-313 /
 * set_body_mass :844-857 on the device arrays.
 * delete_row: np.delete(row) on every body array (order preserved, n -> n-1).
 * swap_rows : swap_with_first-style swap of rows a,b in mass, den, rad, pos, vel,
 *             rel_vel, parents (coi and the orbit arrays are not swapped, as there).
 * translate_to: pos[:] -= pos[row]  (re-rooting, :848-850). */
int vsb_delete_row(vsb_ctx *ctx, int64_t row);
int vsb_swap_rows(vsb_ctx *ctx, int64_t a, int64_t b);
int vsb_translate_to(vsb_ctx *ctx, int64_t row);

/* C7  editor Physics.curve, phys_editor.py:519-544: all conic curves, layout
 * (2,N,P) f64, from the kepler_basic outputs.  ell_t/par_t as :155-157. */
int vsb_editor_curve(vsb_ctx *ctx, int64_t P, const double *ell_t, const double *par_t,
                     double *out_2NP);

/* ------------------------------------------- stateless host-buffer entry points
 * Same operators over caller-owned numpy buffers: upload, run, download inside
 * the call (this is the end-to-end path bench.py times as `e2e`). */
/* A4 over host arrays, in place on pos/vel. */
int vsb_host_allpairs_step(vsb_ctx *ctx, int64_t n, const double *mass, double *pos, double *vel,
                           double gc, int64_t nsteps);
/* A4 acceleration only, acc_out (n,2). */
int vsb_host_allpairs_accel(vsb_ctx *ctx, int64_t n, const double *mass, const double *pos,
                            double gc, double *acc_out);
/* A1 */
int vsb_host_find_parents(vsb_ctx *ctx, int64_t n, const int64_t *order_desc_mass,
                          const double *pos, const double *coi, int64_t *parents_out);
/* B1 */
int vsb_host_check_collision(vsb_ctx *ctx, int64_t n, const double *mass, const double *pos,
                             const double *rad, int64_t *body_del_out, int64_t *body_add_out);
/* C2  enhanced_kepler_solver.solve_kepler_ell, enhanced_kepler_solver.py:18-69 (batched). */
int vsb_host_solve_kepler_ell(vsb_ctx *ctx, int64_t n, const double *ecc, const double *ma,
                              double tol, double *ea_out);
/* C3  phys_shared.newton_root_kepler_hyp, phys_shared.py:114-122 (batched). */
int vsb_host_newton_root_kepler_hyp(vsb_ctx *ctx, int64_t n, const double *ecc, const double *ma,
                                    const double *ea_guess, double *ea_out);
/* C5  phys_shared.curve_points, phys_shared.py:202-216, for n objects: out (n,P,2). */
int vsb_host_curve_points(vsb_ctx *ctx, int64_t n, int64_t P, const double *ecc, const double *a,
                          const double *b, const double *pea, const double *t, double *out);
/* C6  phys_shared.curve_move_to, phys_shared.py:219-222. */
int vsb_host_curve_move_to(vsb_ctx *ctx, int64_t n, int64_t P, int64_t nref, const double *curves,
                           const double *body_pos, const int64_t *ref, const double *f,
                           const double *pea, double *out);

/* ------------------------------------------------ game-mode Keplerian state
 * Device mirror of phys_body.Physics (kind BODY, phys_body.py:122-157) or
 * phys_vessel.Physics (kind VESSEL, phys_vessel.py:139-173) orbit arrays. */
#define VSB_KIND_BODY 0
#define VSB_KIND_VESSEL 1

enum vsb_kepler_field {
    VSB_K_A = 0, VSB_K_B = 1, VSB_K_F = 2, VSB_K_ECC = 3, VSB_K_PEA = 4, VSB_K_N = 5,
    VSB_K_DR = 6, VSB_K_MA = 7, VSB_K_EA = 8, VSB_K_POS = 9 /* (N,2) */, VSB_K_REF = 10 /* i64 */,
    VSB_K_COUNT_ = 11
};

/* Physics.load + the orbit part of initial(): phys_body.py:177-198,221-312 /
 * phys_vessel.py:198-240.  t = np.linspace(-pi, pi, P) (phys_body.py:165).
 * The derived elements b, f, n come from calc_orb_one (C8, host side). */
int vsb_kepler_load(vsb_ctx *ctx, int kind, int64_t n, int64_t P, const double *a,
                    const double *b, const double *f, const double *ecc, const double *pea,
                    const double *n_motion, const double *dr, const double *ma, const double *ea,
                    const int64_t *ref, const double *t);
/* kind BODY only: processing order np.argsort(coi)[-1::-1] of phys_body.py:328; the
 * library turns it into dependency levels (a body whose ref comes earlier in the order
 * sees the ref's NEW position, otherwise last tick's). */
int vsb_kepler_set_order(vsb_ctx *ctx, const int64_t *order_desc_coi);
int vsb_kepler_set(vsb_ctx *ctx, int kind, int field, const void *src, int64_t first, int64_t count);
int vsb_kepler_get(vsb_ctx *ctx, int kind, int field, void *dst, int64_t first, int64_t count);
int vsb_kepler_device_ptr(vsb_ctx *ctx, int kind, int field, void **dptr_out);

/* C1  phys_body.Physics.move(warp), phys_body.py:323-346           (kind BODY;  body_pos NULL)
 * C4  phys_vessel.Physics.move(warp, body_pos, ..), phys_vessel.py:477-494 (kind VESSEL;
 *     body_pos = (nref,2) host array of parent-body positions, or NULL to use the
 *     device-resident BODY state's positions).
 * Includes C2/C3.  Results stay on the device (fetch with vsb_kepler_get). */
int vsb_kepler_move(vsb_ctx *ctx, int kind, double warp, const double *body_pos, int64_t nref);
/* The same move with its results handed back in the same call -- what Physics.move returns
 * (phys_body.py:346 -> (pos, ma, ea); phys_vessel.py:494): pos_out (N,2), ma_out (N,), ea_out (N,),
 * one wait instead of a launch + three copies (small states: through mapped page-locked memory). */
int vsb_kepler_move_state(vsb_ctx *ctx, int kind, double warp, const double *body_pos, int64_t nref,
                          double *pos_out, double *ma_out, double *ea_out);

/* C5+C6 fused: Physics.curve(i) for all i + Physics.curve_move() (phys_body.py:315-320,
 * 405-411; phys_vessel.py:364-369,497-503): writes the moved curves (N,P,2) into the
 * device-resident buffer straight from the elements (write-only traffic). */
int vsb_kepler_curves(vsb_ctx *ctx, int kind);
/* One game tick for kind VESSEL in a single kernel: move + curves (the bench path). */
int vsb_kepler_tick(vsb_ctx *ctx, int kind, double warp, const double *body_pos, int64_t nref);
/* Same tick enqueued without a host sync, parent positions already resident
 * (device-timed benchmarking). */
int vsb_kepler_tick_async(vsb_ctx *ctx, int kind, double warp);
int vsb_kepler_free(vsb_ctx *ctx, int kind);
/* Copy curves of objects [first, first+count) to the host, layout (count,P,2);
 * only visible curves are ever consumed by the caller (game.py:1403-1407). */
int vsb_kepler_curves_get(vsb_ctx *ctx, int kind, int64_t first, int64_t count, double *out);
/* Physics.curve_move() (phys_body.py:405-411, phys_vessel.py:497-503) in one call: vsb_kepler_curves
 * + the copy of all N moved curves (N,P,2) to the host, one wait. */
int vsb_kepler_curve_move(vsb_ctx *ctx, int kind, double *out);
int vsb_kepler_curves_device_ptr(vsb_ctx *ctx, int kind, void **dptr_out);

/* --------------------------------------------- Newton <-> Kepler conversion
 * convert.to_kepler, convert.py:193-261 (opening a Newtonian map in game mode, game.py:318-319)
 * and convert.to_newton, convert.py:132-190 with kepler_to_velocity :20-64 (opening a Keplerian
 * save in the editor, editor.py:295-296), over host arrays; the reference's index quirks are
 * reproduced (see vsb_convert.cu).  to_newton reports the reference's ValueError for hyperbolic
 * bodies as VSB_ERR_ARG "math domain error". */
int vsb_host_to_kepler(vsb_ctx *ctx, int64_t n, const int64_t *order_desc_mass, const double *mass,
                       const double *pos, const double *vel, double gc, double coi_coef,
                       int64_t *ref_out, double *a_out, double *ecc_out, double *pe_arg_out,
                       double *ma_out, double *dir_out);
int vsb_host_to_newton(vsb_ctx *ctx, int64_t n, const double *mass, const double *a,
                       const double *ecc, const double *pe_arg, const double *ma,
                       const int64_t *ref, const double *dir, double gc, double *pos_out,
                       double *vel_out);

/* --------------------------------------------- COI-crossing numeric kernels
 * Batched forms (one item per row) of the numeric routines under phys_vessel.points
 * (phys_vessel.py:372-474); the branchy per-vessel bookkeeping around them stays host side.
 * solve_quartic: quartic_solver.py:57-93, coef5 = (n,5) a,b,c,d,e -> roots8 = (n,4) complex as
 *   interleaved re,im.
 * ell_hyp_intersect_circle: orbit_intersect.py:17-52, in6 = (n,6) a,b,ecc,c_x,c_y,r ->
 *   ea4_out (n,4) eccentric anomalies NaN-padded, count_out (n) (0 = the reference's [nan]).
 * predict_enter_coi: orbit_intersect.py:125-174, vessel_data10 = (n,10) a,b,f,ecc,ma,ea,pea,
 *   period,n,dr; body_data10 = (n,10) a,b,f,ecc,ma,ea,pea,n,dr,coi -> out5 (n,5) ea,b_ea,ma,
 *   b_ma,t (all NaN when the vessel does not enter the COI within one period). */
int vsb_host_solve_quartic(vsb_ctx *ctx, int64_t n, const double *coef5, double *roots8);
int vsb_host_ell_hyp_intersect_circle(vsb_ctx *ctx, int64_t n, const double *in6,
                                      double *ea4_out, int64_t *count_out);
int vsb_host_predict_enter_coi(vsb_ctx *ctx, int64_t n, const double *vessel_data10,
                               const double *body_data10, double tol, int64_t steps_limit,
                               double *out5);

/* phys_vessel.Physics.points(vessel, only_enter_coi), phys_vessel.py:372-474, for the vessels
 * listed in sel (nsel entries; NULL = all nv vessels -- the loop of initial(), :300-302), with
 * next_point / sort_intersect_indices (orbit_intersect.py:55-79) and orbit_time_to
 * (phys_shared.py:58-64).
 *   vessel12 (nv,12) = a b f ecc ma ea pea n dr pe_d ap_d period ; v_ref (nv)
 *   body12   (nb,12) = a b f ecc ma ea pea n dr coi size atm_h   ; body_ref (nb)
 *   entered_coi / left_coi / left_coi_prev_ref: the attributes of the same names (-1 = None)
 *   steps_limit: [game] predict_coi_limit (phys_vessel.py:181)
 * In/out host arrays, rows of the selected vessels rewritten, all other rows untouched:
 *   body_impact (nv,2), body_enter_atm (nv,2), coi_leave (nv,2) (read by only_enter_coi, :441),
 *   coi_enter (nv,6) = body index, ea, b_ea, ma, b_ma, t (all NaN when nothing is entered).
 * Every (vessel, candidate body) pair is marched by one thread block (see vsb_intersect.cu). */
int vsb_host_vessel_points(vsb_ctx *ctx, int64_t nv, int64_t nb, const double *vessel12,
                           const int64_t *v_ref, const double *body12, const int64_t *body_ref,
                           int64_t nsel, const int64_t *sel, int64_t entered_coi, int64_t left_coi,
                           int64_t left_coi_prev_ref, int only_enter_coi, int64_t steps_limit,
                           double *body_impact, double *body_enter_atm, double *coi_leave,
                           double *coi_enter);
/* Modes of the COI-entry march used by the two calls above.  mode bit 0: the reference's own
 * serial loop, one thread per pair (cross-check of the parallel march; the results must be
 * bit-identical).  mode bit 1: "exact libm" -- sin / cos / pow / acos / atan2 of the march and of
 * the orbit-circle quartic are evaluated in double-double arithmetic and rounded to nearest
 * (~20x their normal cost), which reproduces the reference's libm where CUDA's 1-2 ulp routines
 * decide differently: the ill-conditioned marches of high-eccentricity ellipses and the degenerate
 * Ferrari quartic.  Default 2 (parallel march, exact
 * trigonometry); 0 is the fast variant.
 * vsb_march_stats: out4 = chunks marched, chunks re-run after a failed speculation, pairs that
 * fell back to the serial loop, pairs refused (> 2^32 steps) -- of the last call. */
int vsb_set_march_mode(vsb_ctx *ctx, int mode);
/* The closed form the parallel march uses for the reference's running sums (`t += dt`,
 * `ma += dr * n * dt`, orbit_intersect.py:108-121,170): xs[i] = x_{ks[i]} of
 * x_k = fl(x_{k-1} + c), x_0 = x0, with exactly the bits of the serial loop.  K_out = kmax, or
 * (bounded) the trip count of `while x < bound`.  Pure host code, no device needed: it exists so
 * that the closed form can be tested against the plain loop.  VSB_ERR_STATE when no closed form
 * exists (more than 288 segments, or a bounded sum that stops growing below its bound). */
int vsb_host_running_sum(double x0, double c, int64_t kmax, int bounded, double bound, int64_t nq,
                         const int64_t *ks, double *xs, int64_t *K_out, int64_t *nseg_out);
int vsb_march_stats(vsb_ctx *ctx, uint64_t *out3);

/* --------------------------------------------- per-tick vessel events
 * The serial per-vessel loops of the game tick (game.py:1165-1265) as order-preserving parallel
 * scans over host arrays of nv vessels; body_impact / body_enter_atm / coi_leave are (nv,2),
 * coi_enter (nv,6), as filled by vsb_host_vessel_points.
 *
 * check_warp: phys_vessel.Physics.check_warp(warp), phys_vessel.py:506-560.  hold (nv bytes) is
 * the membership mask of `physical_hold`, updated in place.  out3 = situation (0 impact,
 * 1 entering atmosphere, 2 leaving atmosphere, 3 crossing COI; -1 = None), alarm_vessel (the last
 * vessel raising one, -1 = None), len(physical_hold). */
int vsb_host_check_warp(vsb_ctx *ctx, int64_t nv, const double *ecc, const double *dr,
                        const double *ea, const double *ma, const double *n_motion,
                        const double *body_impact, const double *body_enter_atm,
                        const double *coi_leave, const double *coi_enter, double warp,
                        uint8_t *hold, int64_t *out3);
/* predict_enter_coi_service, phys_vessel.py:744-758: flags_out[v] = 1 for the vessels it calls
 * points(v, True) on (feed them to vsb_host_vessel_points as `sel`); crossing_pending != 0 when
 * entered_coi or left_coi is set (the service then does nothing). */
int vsb_host_enter_coi_service(vsb_ctx *ctx, int64_t nv, const double *dr, const double *prev_ea,
                               const double *ea, const double *coi_enter, int crossing_pending,
                               uint8_t *flags_out);
/* cross_coi, phys_vessel.py:563-632, detection: out2 = the FIRST vessel whose next orbit point is
 * a COI crossing passed between prev_ea and ea (-1 = none), kind (1 enter, 2 leave). */
int vsb_host_cross_scan(vsb_ctx *ctx, int64_t nv, const double *ecc, const double *dr,
                        const double *prev_ea, const double *ea, const double *body_impact,
                        const double *coi_leave, const double *coi_enter, int64_t *out2);
/* cross_coi, the orbit change (:577-629) with convert.kepler_to_velocity / velocity_to_kepler
 * (convert.py:20-131), for nev events at once.  vessel8 (nev,8) = a b f ecc pea u dr ea_cross;
 * ref / new_ref: current and new reference body; body7 (nb,7) = a ecc pea dr mass pos_x pos_y.
 * out6 (nev,6) = a ecc pea ma dr new_ref; status[i] = 1 where the reference returns None
 * (entering the body the vessel already orbits). */
int vsb_host_cross_apply(vsb_ctx *ctx, int64_t nev, const int32_t *kind, const double *vessel8,
                         const int64_t *ref, const int64_t *new_ref, int64_t nb,
                         const double *body7, double gc, double *out6, int32_t *status);

/* The same, on device-resident state (no per-vessel array crosses PCIe per tick).  The vessel
 * elements are the KIND_VESSEL Kepler state (vsb_kepler_load / vsb_kepler_move / vsb_kepler_tick,
 * which now also keep the previous tick's ea: self.prev_ea, phys_vessel.py:486);
 * vsb_vessel_events_load adds the orbit points, the physical_hold mask and the derived pe_d /
 * ap_d / period of every vessel (NULL points = NaN, NULL hold = empty, NULL prev_ea = current ea).
 * vsb_vessel_points fills the resident points (body12 / body_ref as in vsb_host_vessel_points);
 * vsb_vessel_check_warp / _enter_coi_service / _cross_scan are the three per-tick scans;
 * vsb_vessel_events_get copies rows [first, first+count) back (NULL = skip an array);
 * vsb_vessel_events_set_orbit refreshes one vessel's derived elements after change_vessel. */
int vsb_vessel_events_load(vsb_ctx *ctx, const double *body_impact, const double *body_enter_atm,
                           const double *coi_leave, const double *coi_enter, const uint8_t *hold,
                           const double *prev_ea, const double *pe_d, const double *ap_d,
                           const double *period);
int vsb_vessel_events_set_orbit(vsb_ctx *ctx, int64_t vessel, double pe_d, double ap_d,
                                double period);
int vsb_vessel_events_get(vsb_ctx *ctx, int64_t first, int64_t count, double *body_impact,
                          double *body_enter_atm, double *coi_leave, double *coi_enter,
                          uint8_t *hold, double *prev_ea);
int vsb_vessel_points(vsb_ctx *ctx, int64_t nb, const double *body12, const int64_t *body_ref,
                      int64_t nsel, const int64_t *sel, int64_t entered_coi, int64_t left_coi,
                      int64_t left_coi_prev_ref, int only_enter_coi, int64_t steps_limit);
int vsb_vessel_check_warp(vsb_ctx *ctx, double warp, int64_t *out3);
int vsb_vessel_enter_coi_service(vsb_ctx *ctx, int crossing_pending, int64_t *idx_out,
                                 int64_t *count_out);
int vsb_vessel_cross_scan(vsb_ctx *ctx, int64_t *out2);

/* --------------------------------------------- per-item element conversions over host arrays
 * convert.kepler_to_velocity (convert.py:20-64; out (n,2)) and convert.velocity_to_kepler
 * (:67-131; failsafe 0/1/2; out5 (n,5) = a, ecc, pe_arg, ma, dr); phys_shared.orb2xy
 * (phys_shared.py:187-199, ref_pos (n,2), (0,0) where ea == 0 exactly); kepler_basic_one for every
 * body (phys_editor.py:63-83 via kepler_basic :381-384) on host arrays instead of the resident
 * body state.  One item per thread. */
int vsb_host_kepler_to_velocity(vsb_ctx *ctx, int64_t n, const double *rel_pos, const double *a,
                                const double *ecc, const double *pe_arg, const double *u,
                                const double *dr, double *vel_out);
int vsb_host_velocity_to_kepler(vsb_ctx *ctx, int64_t n, const double *rel_pos,
                                const double *rel_vel, const double *u, int failsafe, double *out5);
int vsb_host_orb2xy(vsb_ctx *ctx, int64_t n, const double *a, const double *b, const double *f,
                    const double *ecc, const double *pea, const double *ref_pos, const double *ea,
                    double *out);
int vsb_host_kepler_basic(vsb_ctx *ctx, int64_t n, const int64_t *parents, const double *mass,
                          const double *pos, const double *vel, double gc, double coi_coef,
                          double *focus, double *ecc_v, double *semi_major, double *semi_minor,
                          double *periapsis_arg, double *coi);

/* --------------------------------------------- per-frame body passes
 * Physics.body (phys_editor.py:577-599, without the radius, which stays host numpy),
 * Physics.temp_color (:602-651) and Physics.classify (:654-670) in one pass over host arrays.
 * consts12 = {gc_sim, rad_mult, mass_thermal_mult, min_planet_mass, gc_real, c, ls, ms,
 * sqrt(2), pi**(1/4), sigma**(1/4), 4*pi} evaluated by the caller with the reference's own
 * expressions (phys_shared.py:13-17).  atm_h is in/out (bodies without an atmosphere keep their
 * value); base_color / color are (n,3) int64 like the reference's int arrays; stellar_class:
 * 0 Unknown, 1 M, 2 K, 3 G, 4 F, 5 A, 6 B, 7 O. */
int vsb_host_body_thermal(vsb_ctx *ctx, int64_t n, const double *consts12, const double *mass,
                          const double *den, const double *rad, const double *atm_pres0,
                          const double *atm_scale_h, const double *atm_den0, double *atm_h,
                          const int64_t *base_color, double *surf_grav, double *luminosity,
                          double *temp, double *rad_sc, int64_t *color, double *types,
                          int32_t *stellar_class);

/* --------------------------------------------- culling + visible subset */
/* phys_shared.culling, phys_shared.py:224-231, over host arrays: mask_out[i] = 1 if the
 * object is on screen.  screen_bounds = the 2x2 array [[x0,y0],[x1,y1]] flattened. */
int vsb_host_culling(vsb_ctx *ctx, int64_t n, const double *coords, const double *radius,
                     const double *screen_bounds, double zoom, unsigned char *mask_out);
/* The same predicate on the device-resident positions of the Kepler state; returns the visible
 * indices in ascending order (np.arange(n)[truth], phys_body.py:208-210, phys_vessel.py:264-265).
 * radius: host array (n) or NULL for zeros; idx_out needs room for n entries. */
int vsb_kepler_culling(vsb_ctx *ctx, int kind, const double *radius, const double *screen_bounds,
                       double zoom, int64_t *idx_out, int64_t *count_out);
/* Same with one radius for every object -- phys_vessel.Physics.culling's
 * np.array([10] * n) / zoom (phys_vessel.py:264) without an n-element upload per frame. */
int vsb_kepler_culling_uniform(vsb_ctx *ctx, int kind, double radius, const double *screen_bounds,
                               double zoom, int64_t *idx_out, int64_t *count_out);
/* visible_orbits of phys_body.py:212-216 / phys_vessel.py:267-271:
 * ((ref_visible[ref] && ref_coi[ref] > 8/zoom) || ref == 0) && size*zoom < screen_x*15, with
 * size = a (bodies) or pe_d (vessels); ascending indices. */
int vsb_kepler_visible_orbits(vsb_ctx *ctx, int kind, int64_t nref,
                              const unsigned char *ref_visible, const double *ref_coi,
                              const double *size, double zoom, double screen_x, int64_t *idx_out,
                              int64_t *count_out);
/* Gather the moved curves of the listed objects into out (count,P,2) with one copy: only
 * visible curves are drawn (game.py:1403-1407). */
int vsb_kepler_curves_gather(vsb_ctx *ctx, int kind, const int64_t *idx, int64_t count,
                             double *out);

/* phys_vessel.Physics.curve_segments, phys_vessel.py:828-1075 (with concat_wrap :35-79), for the
 * listed vessels (self.visible_orbits), on the device-resident VESSEL state: elements and positions
 * (vsb_kepler_load / _move / _tick), orbit points (vsb_vessel_events_load / vsb_vessel_points) and
 * the moved curves (vsb_kepler_curves / _tick).  Row k of every output belongs to vessel idx[k]:
 * light, dark (count,P,2), NaN padded -- dark is all NaN where the reference never writes it (a
 * vessel without any orbit point); first_intersect (count,2); intersect_type (count): 0 none,
 * 1 impact, 2 COI entry, 3 COI leave; intersect_time (count); select_range (count,2). */
int vsb_vessel_curve_segments(vsb_ctx *ctx, const int64_t *idx, int64_t count, double *light,
                              double *dark, double *first_intersect, int32_t *intersect_type,
                              double *intersect_time, double *select_range);

#ifdef __cplusplus
}
#endif
#endif /* VSB200_H */
