// vsb_gravity_sym_lean.cu -- the symmetric all-pairs kernel (vsb_gravity_sym.cu: every UNORDERED
// pair evaluated once, same pair arithmetic, same macro-block / sub-block / lane rotation) WITHOUT
// partial-sum slabs, for body counts whose slabs (N^2/(2S) * 16 B: 8.6 GB at 1 M bodies, 137 GB at
// 4 M) do not fit or are not wanted.  Selected automatically above VSB_SYM_SLAB_GB (default 16) of
// slab memory, or with VSB_SYM_MODE=lean.
//
// The kernel is PERSISTENT: one CTA per resident slot, tasks (a block Q against a range of blocks
// P <= Q) pulled in ticket order from an atomic counter, the next task's Q block prefetched by a
// 1-D TMA bulk copy into a second shared buffer while the current one computes.  Both sides of a
// block pair end up in acc[] (16 B/body, L2 resident) in a FIXED ORDER per block:
//   * targets' side (block P, registers): read-add-write into acc[P] after every block pair.  The
//     contribution carries the number of contributions to that block that precede it -- a table
//     the host builds by walking the ticket list -- and each warp spins on its own (block,
//     sub-block) counter until that many have landed, adds, and bumps the counter;
//   * sources' side (block Q, shared memory): a task parks its sums in a staging slab (a pool of
//     32 MB, recycled), and the LAST task of row Q to finish adds the row's slabs in ticket order
//     and makes the row's single ordered contribution to acc[Q] (a chain of read-add-writes by the
//     tasks of a row, which all finish together, would serialise them).
// A warp only ever waits for tasks with a LOWER ticket, which are running or done (tickets are
// handed out in order), so the scheme cannot deadlock.  Ticket order: rows Q descending, each row
// cut into P-ranges so that a long row is about one wave of the resident CTAs -- the tasks running
// side by side then belong to one or two rows and to different P-ranges, and the contribution a
// task has to follow (the previous row's, same blocks) comes from a task that started a wave
// earlier: 1.5-3 % of the contributions find their predecessor late (VSB_SYM_DEBUG=s counts them).
// Bit-identical run to run (for a given GPU count).  Measured at 1 M bodies: 555 ms per tick
// (FP64 pipe 85 %; the slab kernel: 526 ms, 90.3 %), DRAM traffic 0.80 GB per launch instead of
// 9.5 GB -- a consumed staging slab is discarded in L2 (discard.global.L2) before it is written
// back; 7.1 GB without that -- and 48 MB of scratch instead of 8.6 GB
// (profiles/r2_dram_gravity1m_lean.csv).
//
// What was tried on the way (profiles/README.md, round 2): ticket order by diagonals of the task
// grid (a contribution made at the END of a long task precedes those of tasks running beside it:
// waits of a whole task); per-task source-side read-add-writes (chains of 100+ within a row);
// parking contributions that are not ready in per-warp rings and applying them later (the warps of
// a CTA meet at barriers; a warp waiting there cannot apply its parked sums: deadlock).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vsb_common.cuh"

namespace {

struct sym_task {
    int q, p0, p1, exp_off;  // block Q, blocks P in [p0, p1); expect[exp_off + (P - p0)] = number of
                             // contributions to block P that precede this task's; entry
                             // [exp_off + (p1 - p0)] is the same for the ROW's contribution to block Q
    int row_t0, row_tasks;   // first ticket and number of tasks of row Q (on this rank)
    int wait_row, pad;       // row whose reduction frees this task's staging slab (-1: none)
};

__device__ __forceinline__ double inv_r3(double dx, double dy, double &r2_out)
{
    const double dy2 = dy * dy;
    double r2 = fma(dx, dx, dy2);
    // The seed only defines the HIGH word of y.  Instead of zeroing the low word (one more
    // instruction on the dispatch port per pair) y takes the low word of the dead temporary
    // dy2: a perturbation below 2^-20, absorbed by the correction below (e stays under 2^-18.6,
    // so the first dropped term 2.19 e^3 is under 3e-17).
    double y = __hiloint2double(__double2hiint(rsqrt_seed(r2)), __double2loint(dy2));
    double t = y * y;
    double e = fma(-r2, t, 1.0);
    double y3 = t * y;
    double p = fma(1.875, e, 1.5);
    double q = e * p;
    r2_out = r2;
    return fma(y3, q, y3);
}

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- ordered accumulation into acc[] -------------------------------------------------------
// A contribution = the T*32 sums of one warp for one (block, sub-block): rows base + k*32 + lane.
// It is added when `expect` earlier contributions to that sub-block have landed (counter `sem`):
// spin (rarely long: see the ticket order in the launcher), read-add-write through L2, publish.
// (Parking a contribution that is not ready and moving on was tried and dropped: the warps of a CTA
// meet at barriers, a warp waiting there cannot apply its parked sums, and the scheme can deadlock.)
template <int T, typename F>
__device__ __forceinline__ void ordered_add(double2 *__restrict__ rows, int *sem, int expect, int lane,
                                            F value, unsigned long long *dbg)
{
    if (lane == 0) {
        unsigned ns = 32;
        bool waited = false;
        while (ld_acquire(sem) != expect) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
            waited = true;
        }
        if (dbg && waited) atomicAdd(dbg, 1ull);
    }
    __syncwarp();
    double2 v[T];
#pragma unroll
    for (int k = 0; k < T; ++k) v[k] = __ldcg(rows + k * 32);
#pragma unroll
    for (int k = 0; k < T; ++k) {
        const double2 a = value(k);
        v[k].x += a.x;
        v[k].y += a.y;
        __stcg(rows + k * 32, v[k]);
    }
    __syncwarp();      // the lanes' stores are ordered before lane 0's release (cumulativity)
    if (lane == 0) st_release(sem, expect + 1);
}

// WARPS warps per CTA, T targets per lane (registers), SU = hand-over of a Q sub-block
// between warps: 1 = CTA barrier, 2 = pairwise named barriers (warp w meets only its ring neighbours).
// Macro-block S = WARPS * T * 32 bodies.  A Q sub-block (T*32 bodies, one per warp and
// sub-round) is cut into NG = T*8 groups of 4 sources (members NG apart, so consecutive lanes
// read consecutive 16-byte words); lane l works in step k on group (l + k) % NG.
// ctl[0] = ticket counter, ctl[1 ...] = the (block, sub-block) counters.
template <int WARPS, int T, int SU, int MB>
__global__ void __launch_bounds__(WARPS * 32, MB)
allpairs_symlean_kernel(const double2 *__restrict__ pos, const double *__restrict__ mass,
                    const sym_task *__restrict__ tasks, const int *__restrict__ expect, int ntasks,
                    int *__restrict__ ctl, int M, double2 *__restrict__ acc,
                    double2 *__restrict__ stage, int pool, unsigned long long *dbg)
{
    constexpr int SB = T * 32;
    constexpr int NG = SB / 4;
    constexpr int S = WARPS * SB;
    extern __shared__ __align__(128) unsigned char sym_smem[];
    double2 *sxy0 = (double2 *)sym_smem;          // 2 x S (double-buffered Q positions)
    double2 *sacc = sxy0 + 2 * S;                 // S
    double *sm0 = (double *)(sacc + S);           // 2 x S (double-buffered Q masses)
    __shared__ alignas(8) uint64_t bar[2];
    __shared__ int s_ticket[2];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int *sems = ctl + 1;
    int *row_done = sems + M * WARPS, *row_reduced = row_done + M;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
        const int t0 = atomicAdd(ctl, 1);
        s_ticket[0] = t0;
        if (t0 < ntasks) {
            const int64_t qb = (int64_t)tasks[t0].q * S;
            mbar_expect_tx(&bar[0], S * (sizeof(double2) + sizeof(double)));
            tma_bulk_g2s(sxy0, pos + qb, S * sizeof(double2), &bar[0]);
            tma_bulk_g2s(sm0, mass + qb, S * sizeof(double), &bar[0]);
        }
    }
    for (int i = tid; i < S; i += WARPS * 32) sacc[i] = make_double2(0.0, 0.0);
    __syncthreads();

    for (int it = 0;; ++it) {
        const int b = it & 1;
        const int ticket = s_ticket[b];
        if (ticket >= ntasks) break;
        const sym_task task = tasks[ticket];
        const double2 *sxy = sxy0 + b * S;
        const double *sm = sm0 + b * S;
        // the next ticket, and its Q block on the way into the other buffer (free since the
        // barrier that ended the previous task)
        if (tid == 0) {
            const int t1 = atomicAdd(ctl, 1);
            s_ticket[b ^ 1] = t1;
            if (t1 < ntasks) {
                const int64_t qb = (int64_t)tasks[t1].q * S;
                mbar_expect_tx(&bar[b ^ 1], S * (sizeof(double2) + sizeof(double)));
                tma_bulk_g2s(sxy0 + (b ^ 1) * S, pos + qb, S * sizeof(double2), &bar[b ^ 1]);
                tma_bulk_g2s(sm0 + (b ^ 1) * S, mass + qb, S * sizeof(double), &bar[b ^ 1]);
            }
        }
        mbar_wait(&bar[b], (uint32_t)(it >> 1) & 1u);

        const bool has_diag = task.p1 == task.q + 1;
        const int p_end = has_diag ? task.q : task.p1;
        const int *exp_row = expect + task.exp_off;
        double xi[T], yi[T], mi[T], ax[T], ay[T];
        for (int P = task.p0; P < p_end; ++P) {
            // targets of this lane: rows pbase + k*32
            const int64_t pbase = (int64_t)P * S + w * SB + lane;
#pragma unroll
            for (int k = 0; k < T; ++k) {
                double2 p = pos[pbase + k * 32];
                xi[k] = p.x;
                yi[k] = p.y;
                mi[k] = -mass[pbase + k * 32];   // negated: the source side gets -m_i w d
                ax[k] = ay[k] = 0.0;
            }
            for (int rho = 0; rho < WARPS; ++rho) {
                const int sb = ((w + rho) & (WARPS - 1)) * SB;
#pragma unroll 1
                for (int step = 0; step < NG; ++step) {
                    const int g = (lane + step) & (NG - 1);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int idx = sb + s * NG + g;
                        const double2 pj = sxy[idx];
                        const double mj = sm[idx];
                        double2 aj = sacc[idx];
#pragma unroll
                        for (int k = 0; k < T; ++k) {
                            double dx = pj.x - xi[k], dy = pj.y - yi[k];
                            double r2;
                            double wv = inv_r3(dx, dy, r2);
                            double wi = wv * mj;
                            ax[k] = fma(wi, dx, ax[k]);
                            ay[k] = fma(wi, dy, ay[k]);
                            double wj = wv * mi[k];
                            aj.x = fma(wj, dx, aj.x);
                            aj.y = fma(wj, dy, aj.y);
                        }
                        sacc[idx] = aj;
                    }
                    __syncwarp();   // the group moves on to the next lane
                }
                // the sub-block moves on to the next warp
                if (SU == 1) {
                    __syncthreads();
                } else {
                    // warp w hands its sub-block to warp w-1 and takes over the one of warp w+1:
                    // rendezvous with each ring neighbour (even warps right first, odd warps left
                    // first, so the ring cannot deadlock); barrier 1+x belongs to the pair (x, x+1)
                    const int right = 1 + w, left = 1 + ((w + WARPS - 1) & (WARPS - 1));
                    const int first = (w & 1) ? left : right, second = (w & 1) ? right : left;
                    asm volatile("bar.sync %0, 64;" ::"r"(first) : "memory");
                    asm volatile("bar.sync %0, 64;" ::"r"(second) : "memory");
                }
            }
            // target-side sums of the pair (P, Q) -> acc[P], in this block's fixed order
            ordered_add<T>(acc + pbase, sems + P * WARPS + w, exp_row[P - task.p0], lane,
                           [&](int k) { return make_double2(ax[k], ay[k]); }, dbg);
        }
        if (has_diag) {
            // diagonal block: directed, masked; its sums join the sources' side below
            const int64_t pbase = (int64_t)task.q * S + w * SB + lane;
#pragma unroll
            for (int k = 0; k < T; ++k) {
                double2 p = pos[pbase + k * 32];
                xi[k] = p.x;
                yi[k] = p.y;
                ax[k] = ay[k] = 0.0;
            }
#pragma unroll 1
            for (int j = 0; j < S; ++j) {
                const double2 pj = sxy[j];
                const double mj = sm[j];
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    double dx = pj.x - xi[k], dy = pj.y - yi[k];
                    const double dy2 = dy * dy;
                    double r2 = fma(dx, dx, dy2);
                    int yh = __double2hiint(rsqrt_seed(r2));
                    yh = __double2hiint(r2) < 0x00100000 ? 0 : yh;   // i == j: exact zero
                    const double y = __hiloint2double(yh, __double2loint(dy2));
                    double t = y * y;
                    double e = fma(-r2, t, 1.0);
                    double ym = y * mj;
                    double y3m = t * ym;
                    double p = fma(1.875, e, 1.5);
                    double q = e * p;
                    double s = fma(y3m, q, y3m);
                    ax[k] = fma(s, dx, ax[k]);
                    ay[k] = fma(s, dy, ay[k]);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < T; ++k) ax[k] = ay[k] = 0.0;
        }
        __syncthreads();   // every warp is done with sacc and with the Q buffer of this task
        {
            // Source-side sums of the task (+ the diagonal pair): parked in the task's staging
            // slab; the LAST task of row Q to finish adds the row's slabs in ticket order and
            // makes the row's one contribution to acc[Q].  (A chain of read-add-writes by the
            // row's tasks -- which all finish at about the same time -- would serialise them.)
            if (task.wait_row >= 0 && tid == 0) {
                unsigned ns = 64;
                while (ld_acquire(row_reduced + task.wait_row) == 0) {     // slab still in use: rare
                    __nanosleep(ns);
                    if (ns < 2048) ns *= 2;
                }
            }
            if (task.wait_row >= 0) __syncthreads();
            double2 *srow = sacc + w * SB + lane;
            double2 *slab = stage + (int64_t)(ticket % pool) * S + w * SB + lane;
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const double2 sq = srow[k * 32];
                __stcg(slab + k * 32, make_double2(sq.x + ax[k], sq.y + ay[k]));
                srow[k * 32] = make_double2(0.0, 0.0);
            }
            __threadfence();
            __syncthreads();   // the slab is complete; sacc is clear; s_ticket[b ^ 1] is visible
            if (tid == 0) s_last = atomicAdd(row_done + task.q, 1) == task.row_tasks - 1;
            __syncthreads();
            if (s_last) {
                __threadfence();
                double2 sum[T];
#pragma unroll
                for (int k = 0; k < T; ++k) sum[k] = make_double2(0.0, 0.0);
                const double2 *base = stage + w * SB + lane;
                int j = 0;
                for (; j + 2 <= task.row_tasks; j += 2) {
                    const double2 *s0 = base + (int64_t)((task.row_t0 + j) % pool) * S;
                    const double2 *s1 = base + (int64_t)((task.row_t0 + j + 1) % pool) * S;
                    double2 a0[T], a1[T];
#pragma unroll
                    for (int k = 0; k < T; ++k) {
                        a0[k] = __ldcg(s0 + k * 32);
                        a1[k] = __ldcg(s1 + k * 32);
                    }
#pragma unroll
                    for (int k = 0; k < T; ++k) {
                        sum[k].x = (sum[k].x + a0[k].x) + a1[k].x;
                        sum[k].y = (sum[k].y + a0[k].y) + a1[k].y;
                    }
                    // the two slabs are dead now: tell L2 not to write their lines back
                    if ((lane & 7) == 0) {
#pragma unroll
                        for (int k = 0; k < T; ++k) {
                            asm volatile("discard.global.L2 [%0], 128;" ::"l"(s0 + k * 32) : "memory");
                            asm volatile("discard.global.L2 [%0], 128;" ::"l"(s1 + k * 32) : "memory");
                        }
                    }
                }
                if (j < task.row_tasks) {
                    const double2 *s0 = base + (int64_t)((task.row_t0 + j) % pool) * S;
#pragma unroll
                    for (int k = 0; k < T; ++k) {
                        const double2 a0 = __ldcg(s0 + k * 32);
                        sum[k].x += a0.x;
                        sum[k].y += a0.y;
                    }
                    if ((lane & 7) == 0) {
#pragma unroll
                        for (int k = 0; k < T; ++k)
                            asm volatile("discard.global.L2 [%0], 128;" ::"l"(s0 + k * 32) : "memory");
                    }
                }
                ordered_add<T>(acc + (int64_t)task.q * S + w * SB + lane, sems + task.q * WARPS + w,
                               exp_row[task.p1 - task.p0], lane, [&](int k) { return sum[k]; },
                               dbg ? dbg + 2 : dbg);
                __syncthreads();
                if (tid == 0) st_release(row_reduced + task.q, 1);
            }
        }
    }
}

// acc[i] *= gc ; optional integrator v += a ; P += v
__global__ void __launch_bounds__(256)
allpairs_symlean_finish_kernel(int64_t n, double gc, double2 *__restrict__ acc, double2 *__restrict__ vel,
                           double2 *__restrict__ pos, int integrate)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 a = acc[i];
    a.x *= gc;
    a.y *= gc;
    acc[i] = a;
    if (integrate) {
        double2 v = vel[i];
        v.x += a.x;
        v.y += a.y;
        vel[i] = v;
        double2 p = pos[i];
        p.x += v.x;
        p.y += v.y;
        pos[i] = p;
    }
}

}  // namespace

// Symmetric all-pairs acceleration for rank `rank` of `world` without slabs (rows of the block
// triangle dealt whole to the ranks).  world == 1: acc is the full result (integrate allowed).
// world > 1: acc holds this rank's partial sum for ALL bodies; the caller all-reduces VSB_F_ACC
// and then integrates.
int vsb_launch_allpairs_sym_lean(vsb_ctx *c, double gc, int rank, int world, int integrate)
{
    const int64_t n = c->n, npad = c->npad;
    if (n <= 0) return VSB_OK;
    // Kernel shape (warps, targets per lane, hand-over, CTAs per SM).  8 targets per lane halve
    // the shared-memory and loop instructions per pair; the macro-block S sets the task grain
    // (small S: better balance at small N).  VSB_SYM_SHAPE=w,t,h,c overrides (tools/sym_ab.py).
    int warps = 4, tpl = 8, su = 2, mb = 3;
    if (npad <= 98304) warps = 2, su = 1, mb = 6;
    const int S = warps * tpl * 32;
    VSB_CHECK(npad % S == 0, VSB_ERR_STATE, "allpairs_sym_lean: padded size %lld not a multiple of %d",
              (long long)npad, S);
    const int M = (int)(npad / S);
    const int slots = c->sm_count * mb;
    // A row Q is cut into P-ranges of c(Q) blocks so that a long row is at least a wave of the
    // resident CTAs (VSB_SYM_WAVE: waves per longest row x 100, default 250; VSB_SYM_CHUNK fixes
    // c): the tasks running side by side then belong to one or two rows and to DIFFERENT P-ranges,
    // and the targets'-side contribution a task has to follow (the previous row's, same blocks)
    // comes from a task that started a whole wave earlier.  Measured at 1 M / 2 M bodies: c = 1
    // 555 / 2202 ms, c = 2 568 / 2219 ms, c = 4 594 / 2277 ms (longer ranges: more waiting).
    int wave_pct = 250, fixed_chunk = 0;
    if (const char *env = getenv("VSB_SYM_WAVE")) wave_pct = atoi(env);
    if (const char *env = getenv("VSB_SYM_CHUNK")) fixed_chunk = atoi(env);
    if (wave_pct < 10) wave_pct = 10;
    const int key_chunk = fixed_chunk * 1000 + wave_pct;
    // staging slabs of the sources' side: 32 MB (L2 resident), at least two long rows and a wave
    int pool = (int)((32u << 20) / (sizeof(double2) * (size_t)S));
    if (pool < 2 * M + 2 * slots) pool = 2 * M + 2 * slots;
    // host-side ticket list + expectation table (cached per n / world / rank / shape)
    if (c->lean_tasks_n != npad || c->lean_world != world || c->lean_rank != rank || c->lean_S != S ||
        c->lean_chunk != key_chunk) {
        // Rows are dealt to the ranks whole (descending, boustrophedon: 0..w-1, w-1..0, ...), so
        // that every rank's rows are waves of its own CTAs and the loads differ by a short row.
        std::vector<sym_task> mine;
        std::vector<int> expect, seen((size_t)M, 0);
        int dealt = 0;
        for (int q = M - 1; q >= 0; --q, ++dealt) {
            const int turn = dealt % (2 * world);
            const int owner = turn < world ? turn : 2 * world - 1 - turn;
            if (owner != rank) continue;
            int chunk = fixed_chunk > 0 ? fixed_chunk
                                        : (int)(((int64_t)(q + 1) * 100 + (int64_t)slots * wave_pct - 1) /
                                                ((int64_t)slots * wave_pct));
            if (chunk < 1) chunk = 1;
            if (chunk > 64) chunk = 64;
            const int row_t0 = (int)mine.size();
            for (int p0 = 0; p0 <= q; p0 += chunk) {
                sym_task t{q, p0, std::min(p0 + chunk, q + 1), (int)expect.size(), row_t0, 0, -1, 0};
                for (int P = t.p0; P < t.p1; ++P) expect.push_back(P == q ? -1 : seen[P]++);
                expect.push_back(seen[q]);          // the row's contribution follows every
                mine.push_back(t);                  // targets'-side one of the earlier rows
            }
            seen[q]++;
            const int row_tasks = (int)mine.size() - row_t0;
            for (int i = row_t0; i < (int)mine.size(); ++i) {
                mine[i].row_tasks = row_tasks;
                if (i >= pool) mine[i].wait_row = mine[i - pool].q;
            }
        }
        const int nm = (int)mine.size();
        const size_t tab_bytes = (sizeof(sym_task) * (size_t)(nm + 1) + 255) / 256 * 256;
        const size_t exp_bytes = (sizeof(int) * (expect.size() + 1) + 255) / 256 * 256;
        const size_t ctl_bytes = (sizeof(int) * ((size_t)M * (warps + 2) + 1) + 255) / 256 * 256;
        VSB_TRY(vsb_buf_reserve(c->lean_tab, tab_bytes + exp_bytes + ctl_bytes));
        VSB_CUDA(cudaMemcpyAsync(c->lean_tab.p, mine.data(), sizeof(sym_task) * (size_t)nm,
                                 cudaMemcpyHostToDevice, c->stream));
        VSB_CUDA(cudaMemcpyAsync((char *)c->lean_tab.p + tab_bytes, expect.data(),
                                 sizeof(int) * expect.size(), cudaMemcpyHostToDevice, c->stream));
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        c->lean_tasks_n = npad;   // the table depends on the padded size only
        c->lean_S = S;
        c->lean_world = world;
        c->lean_rank = rank;
        c->lean_chunk = key_chunk;
        c->lean_ntasks = nm;
        c->lean_exp_off = (int64_t)tab_bytes;
        c->lean_ctl_off = (int64_t)(tab_bytes + exp_bytes);
        c->lean_ctl_bytes = (int64_t)ctl_bytes;
    }
    const sym_task *d_tasks = (const sym_task *)c->lean_tab.p;
    const int *d_expect = (const int *)((char *)c->lean_tab.p + c->lean_exp_off);
    int *d_ctl = (int *)((char *)c->lean_tab.p + c->lean_ctl_off);
    // ticket counter and block counters back to zero, acc to zero (blocks no task of this rank
    // touches must read as zero in the all-reduce)
    VSB_CUDA(cudaMemsetAsync(d_ctl, 0, (size_t)c->lean_ctl_bytes, c->stream));
    VSB_CUDA(cudaMemsetAsync(c->acc, 0, sizeof(double2) * (size_t)npad, c->stream));
    const size_t smem = (size_t)S * (3 * sizeof(double2) + 2 * sizeof(double));
    if (c->lean_ntasks > 0) {
        const int grid = std::min(slots, c->lean_ntasks);
        // the staging slabs of the sources' side (L2 resident)
        VSB_TRY(vsb_buf_reserve(c->part, sizeof(double2) * (size_t)pool * (size_t)S));
        double2 *d_stage = (double2 *)c->part.p;
#define VSB_SYM_CASE(W, T, U, MB)                                                                   \
    if (warps == W && tpl == T && su == U && mb == MB) {                                            \
        VSB_CUDA(cudaFuncSetAttribute(allpairs_symlean_kernel<W, T, U, MB>,                             \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        VSB_CUDA(cudaFuncSetAttribute(allpairs_symlean_kernel<W, T, U, MB>,                             \
                                      cudaFuncAttributePreferredSharedMemoryCarveout,               \
                                      (int)cudaSharedmemCarveoutMaxShared));                        \
        if (dbg_env) {                                                                              \
            int occ = 0;                                                                            \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, allpairs_symlean_kernel<W, T, U, MB>,   \
                                                          W * 32, smem);                            \
            fprintf(stderr, "sym debug: occupancy %d CTAs/SM (want %d), smem %zu\n", occ, MB, smem); \
        }                                                                                           \
        allpairs_symlean_kernel<W, T, U, MB><<<grid, W * 32, smem, c->stream>>>(                        \
            (const double2 *)c->pos, c->mass, d_tasks, d_expect, c->lean_ntasks, d_ctl, M,           \
            (double2 *)c->acc, d_stage, pool, dbg);                                                 \
        launched = true;                                                                            \
    }
        bool launched = false;
        unsigned long long *dbg = nullptr;
        const char *dbg_env = getenv("VSB_SYM_DEBUG");
        if (dbg_env && dbg_env[0] == 's') {
            VSB_TRY(vsb_buf_reserve(c->scratch, 64));
            dbg = (unsigned long long *)c->scratch.p;
            VSB_CUDA(cudaMemsetAsync(dbg, 0, 48, c->stream));
        }
        VSB_SYM_CASE(2, 8, 1, 6)    // S =  512, 12 warps/SM
        VSB_SYM_CASE(4, 8, 2, 3)    // S = 1024, 12 warps/SM, pairwise barriers
#undef VSB_SYM_CASE
        VSB_CHECK(launched, VSB_ERR_ARG, "allpairs_sym_lean: no kernel for shape %d,%d,%d,%d", warps, tpl, su, mb);
        c->launches += 1;
        if (dbg) {
            unsigned long long h[6];
            VSB_CUDA(cudaMemcpyAsync(h, dbg, 48, cudaMemcpyDeviceToHost, c->stream));
            VSB_CUDA(cudaStreamSynchronize(c->stream));
            fprintf(stderr, "sym debug: targets' side waited %llu, rows' side waited %llu; tasks %d, chunk key "
                            "%d, M %d\n", h[0], h[2], c->lean_ntasks, c->lean_chunk, M);
        }
    }
    allpairs_symlean_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        n, gc, (double2 *)c->acc, (double2 *)c->vel, (double2 *)c->pos,
        (world == 1 && integrate) ? 1 : 0);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
