// vsb_gravity_sym.cu -- all-pairs gravity exploiting Newton's third
// law: every UNORDERED pair is evaluated once and applied to both bodies.
//
// Same model and pair arithmetic as vsb_gravity.cu (m_j r^-3 from the MUFU.RSQ64H seed
// with one cubic correction); the shared part of a pair costs 10 FP64-pipe instructions
// and each side 3 more, so a directed interaction costs 8 instead of 13.
//
// Decomposition: bodies are cut into macro-blocks of S = WARPS*128.  A CTA task is one
// block Q (positions, masses AND its accumulators resident in shared memory) against a
// range of blocks P <= Q.  Warp w keeps P's sub-block w in registers (4 targets per
// lane, stride 32) and in sub-round rho works on Q's sub-block (w + rho) % WARPS, so in
// any sub-round every Q sub-block is touched by exactly one warp; inside a sub-block
// lane l works in step k on source group (l + k) % 32, so every shared accumulator has a
// single writer at any time -- no atomics anywhere.  The diagonal task (P == Q) runs
// the directed formula with the i == j mask and writes only target-side sums.
//
// Determinism: the schedule above is fixed, target-side sums of every (P,Q) pair and
// source-side sums of every task go to their own slab, and allpairs_sym_reduce_kernel
// adds the slabs in block order: bit-identical run to run.  (Across different GPU
// counts the result differs in the last bits -- use vsb_gravity.cu when shard
// invariance matters.)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vsb_common.cuh"

namespace {

struct sym_task {
    int q, p0, p1, qslot;  // block Q, blocks P in [p0, p1), slab index of the Q-side sums
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

// WARPS warps per CTA, T targets per lane (registers), SU = hand-over of a Q sub-block
// between warps: 1 = CTA barrier, 2 = pairwise named barriers (warp w meets only its ring neighbours).
// Macro-block S = WARPS * T * 32 bodies.  A Q sub-block (T*32 bodies, one per warp and
// sub-round) is cut into NG = T*8 groups of 4 sources (members NG apart, so consecutive lanes
// read consecutive 16-byte words); lane l works in step k on group (l + k) % NG.
template <int WARPS, int T, int SU, int MB>
__global__ void __launch_bounds__(WARPS * 32, MB)
allpairs_sym_kernel(const double2 *__restrict__ pos, const double *__restrict__ mass,
                    const sym_task *__restrict__ tasks, double2 *__restrict__ part_p,
                    double2 *__restrict__ part_q)
{
    constexpr int SB = T * 32;
    constexpr int NG = SB / 4;
    constexpr int S = WARPS * SB;
    extern __shared__ __align__(128) unsigned char sym_smem[];
    double2 *sxy = (double2 *)sym_smem;          // S
    double2 *sacc = sxy + S;                     // S
    double *sm = (double *)(sacc + S);           // S
    __shared__ alignas(8) uint64_t bar;

    const sym_task task = tasks[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t qbase = (int64_t)task.q * S;

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, S * (sizeof(double2) + sizeof(double)));
        tma_bulk_g2s(sxy, pos + qbase, S * sizeof(double2), &bar);
        tma_bulk_g2s(sm, mass + qbase, S * sizeof(double), &bar);
    }
    for (int i = tid; i < S; i += WARPS * 32) sacc[i] = make_double2(0.0, 0.0);
    mbar_wait(&bar, 0);
    __syncthreads();

    for (int P = task.p0; P < task.p1; ++P) {
        // targets of this lane: rows pbase + w*SB + k*32 + lane
        const int64_t pbase = (int64_t)P * S + w * SB + lane;
        double xi[T], yi[T], mi[T], ax[T], ay[T];
#pragma unroll
        for (int k = 0; k < T; ++k) {
            double2 p = pos[pbase + k * 32];
            xi[k] = p.x;
            yi[k] = p.y;
            mi[k] = -mass[pbase + k * 32];   // negated: the source side gets -m_i w d
            ax[k] = ay[k] = 0.0;
        }
        if (P == task.q) {
            // diagonal block: directed, masked, target side only
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
        }
        // target-side sums of the pair (P, Q): slab q(q+1)/2 + P
        const int64_t slab = (int64_t)task.q * (task.q + 1) / 2 + P;
        double2 *out = part_p + slab * S + w * SB + lane;
#pragma unroll
        for (int k = 0; k < T; ++k) out[k * 32] = make_double2(ax[k], ay[k]);
    }
    __syncthreads();
    double2 *qo = part_q + (int64_t)task.qslot * S;
    for (int i = tid; i < S; i += WARPS * 32) qo[i] = sacc[i];
}

// acc[i] = gc * ( sum of Q-side slabs of block X (tasks in P order) + sum over Q >= X of the
// P-side slab of (X, Q) ), fixed order; optional integrator v += a ; P += v.
// One GPU: the slab addresses are arithmetic.  Several GPUs: a rank sums only the slabs its own
// tasks wrote -- their unified indices (P-side slab s -> s, Q-side slab sl -> pairs + sl; part_q
// follows part_p in memory) are listed per block in `own_list` (same order) -- so a rank's
// reduction costs 1/world of the loads AND of the loop trips.
// The kernel is latency bound (a body's ~2 M 16-byte terms sit 16 S bytes apart), so FOUR lanes
// share a body: lane part p of 4 adds the p-th quarter of the body's slab list in order (eight
// loads in flight at a time), and the quarters are combined as (q0 + q1) + (q2 + q3) -- a fixed
// order, bit-identical run to run.  Lanes part*8 + j of a warp work on body j of the warp's eight,
// so that every load instruction of a part covers 128 contiguous bytes.
__global__ void __launch_bounds__(128)
allpairs_sym_reduce_kernel(const double2 *__restrict__ part_p, const double2 *__restrict__ part_q,
                           const int *__restrict__ qslot_first, const int *__restrict__ own_first,
                           const int *__restrict__ own_list, int S, int M, int64_t n,
                           double gc, double2 *__restrict__ acc, double2 *__restrict__ vel,
                           double2 *__restrict__ pos, int integrate)
{
    const int lane = threadIdx.x & 31, part = lane >> 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t i = warp * 8 + (lane & 7);
    const bool live = i < n;
    const int X = (int)((live ? i : n - 1) / S);
    const int r = (int)((live ? i : n - 1) - (int64_t)X * S);
    double sx = 0.0, sy = 0.0;
    constexpr int RB = 8;
    if (own_first) {
        const int b0 = own_first[X], len = own_first[X + 1] - b0;
        const int e0 = b0 + (int)((int64_t)len * part / 4), e1 = b0 + (int)((int64_t)len * (part + 1) / 4);
        for (int e = e0; e < e1; e += RB) {
            double2 v[RB];
#pragma unroll
            for (int k = 0; k < RB; ++k)
                v[k] = e + k < e1 ? part_p[(int64_t)own_list[e + k] * S + r] : make_double2(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                sx += v[k].x;
                sy += v[k].y;
            }
        }
    } else {
        // the body's list: Q-side slabs qslot_first[X] .. qslot_first[X+1]-1, then P-side slabs of
        // the pairs (X, Q), Q = X .. M-1
        const int q0 = qslot_first[X], nq = qslot_first[X + 1] - q0, len = nq + (M - X);
        const int e0 = (int)((int64_t)len * part / 4), e1 = (int)((int64_t)len * (part + 1) / 4);
        for (int e = e0; e < e1; e += RB) {
            double2 v[RB];
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int ee = e + k;
                if (ee >= e1) {
                    v[k] = make_double2(0.0, 0.0);
                } else if (ee < nq) {
                    v[k] = part_q[(int64_t)(q0 + ee) * S + r];
                } else {
                    const int Q = X + (ee - nq);
                    v[k] = part_p[((int64_t)Q * (Q + 1) / 2 + X) * S + r];
                }
            }
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                sx += v[k].x;
                sy += v[k].y;
            }
        }
    }
    // (q0 + q1) + (q2 + q3)
    sx += __shfl_xor_sync(0xffffffffu, sx, 8);
    sy += __shfl_xor_sync(0xffffffffu, sy, 8);
    sx += __shfl_xor_sync(0xffffffffu, sx, 16);
    sy += __shfl_xor_sync(0xffffffffu, sy, 16);
    if (!live || part != 0) return;
    sx *= gc;
    sy *= gc;
    acc[i] = make_double2(sx, sy);
    if (integrate) {
        double2 v = vel[i];
        v.x += sx;
        v.y += sy;
        vel[i] = v;
        double2 p = pos[i];
        p.x += v.x;
        p.y += v.y;
        pos[i] = p;
    }
}

}  // namespace

// Symmetric all-pairs acceleration for rank `rank` of `world` (tasks dealt round-robin in
// descending cost order).  world == 1: acc is the full result (integrate allowed).
// world > 1: acc holds this rank's partial sum for ALL bodies (zero where it has no task
// touching the block); the caller all-reduces VSB_F_ACC and then integrates.
//
// This file is compiled twice (build.py): as it is, and with -DVSB_SYM_PTXAS_SCHEDULE.  The first
// object goes through sass_resched.py (the accumulates of the pair loop regrouped so that they
// ride the operand-reuse cache: same instructions, same registers, same bits, ~1.5 % faster);
// the second keeps ptxas's schedule and is the reference VSB_SYM_SCHED=ptxas selects -- the GPU
// tests compare the two bit for bit.
#ifdef VSB_SYM_PTXAS_SCHEDULE
#define VSB_SYM_LAUNCHER vsb_launch_allpairs_sym_ptxas
#else
#define VSB_SYM_LAUNCHER vsb_launch_allpairs_sym
#endif
int VSB_SYM_LAUNCHER(vsb_ctx *c, double gc, int rank, int world, int integrate)
{
#ifndef VSB_SYM_PTXAS_SCHEDULE
    if (const char *sched = getenv("VSB_SYM_SCHED"))
        if (strcmp(sched, "ptxas") == 0) return vsb_launch_allpairs_sym_ptxas(c, gc, rank, world, integrate);
#endif
    const int64_t n = c->n, npad = c->npad;
    if (n <= 0) return VSB_OK;
    // The slabs take N^2/(2S) * 16 B (8.6 GB at 1 M bodies, 34 GB at 2 M with S = 1024).  Above
    // VSB_SYM_SLAB_GB (default 16, and never more than 60 % of the free device memory) of slab
    // memory, or with VSB_SYM_MODE=lean, the slab-free kernel
    // of vsb_gravity_sym_lean.cu runs instead: ~4 % slower, 48 MB of scratch, no size limit.
    {
        const char *mode = getenv("VSB_SYM_MODE"), *gb = getenv("VSB_SYM_SLAB_GB");
        double budget = (gb ? atof(gb) : 16.0) * 1073741824.0;
        const double slab_bytes = (double)npad * (double)npad / 2048.0 * 16.0;
        if (!gb && slab_bytes > 1073741824.0 && (double)c->part.cap < slab_bytes) {
            // never more than 60 % of what the device can still give (the Kepler state's curves may
            // hold tens of GB in the same context)
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                budget = std::min(budget, 0.6 * ((double)free_b + (double)c->part.cap));
        }
        if ((mode && mode[0] == 'l') || (!(mode && mode[0] == 's') && slab_bytes > budget))
            return vsb_launch_allpairs_sym_lean(c, gc, rank, world, integrate);
    }
    // Kernel shape (warps, targets per lane, hand-over, CTAs per SM).  8 targets per lane halve
    // the shared-memory and loop instructions per pair (profiles/README.md: 2.6 % at 1 M bodies);
    // the macro-block S sets the task grain (small S: better balance at small N) against the
    // slab memory N^2/(2S) * 16 B (S = 1024: 8.6 GB at 1 M bodies, 34 GB at 2 M).
    // VSB_SYM_SHAPE=w,t,h,c overrides (tools/sym_ab.py).
    // (measured: S = 256 wins at 32 k bodies, 512 at 64 k, 1024 from 128 k: profiles/r1_sym_shapes_ab.txt)
    int warps = 4, tpl = 8, su = 2, mb = 3;
    if (npad <= 98304) warps = 2, su = 1, mb = 6;
    if (npad <= 32768) warps = 1, su = 1, mb = 12;
    if (const char *env = getenv("VSB_SYM_SHAPE")) {
        mb = 0;
        sscanf(env, "%d,%d,%d,%d", &warps, &tpl, &su, &mb);
        if (mb == 0) mb = (warps * tpl <= 32) ? 2 : 1;
    }
    VSB_CHECK(warps > 0 && tpl > 0, VSB_ERR_ARG, "allpairs_sym: bad VSB_SYM_SHAPE");
    const int S = warps * tpl * 32;
    VSB_CHECK(npad % S == 0, VSB_ERR_STATE, "allpairs_sym: padded size %lld not a multiple of %d",
              (long long)npad, S);
    const int M = (int)(npad / S);
    const int64_t pairs = (int64_t)M * (M + 1) / 2;
    // P-range chunk: enough tasks that the heaviest one is a small fraction of an SM's share
    int64_t want_tasks = (int64_t)c->sm_count * 32 * world;
    int chunk = (int)(pairs / want_tasks);
    if (chunk < 1) chunk = 1;
    // host-side task table (cached per padded size / world / rank / shape / tail setting)
    const char *tail_env = getenv("VSB_SYM_TAIL");
    const int tail_split = tail_env ? atoi(tail_env) : 8;
    if (c->sym_tasks_n != npad || c->sym_world != world || c->sym_rank != rank || c->sym_S != S ||
        c->sym_tail != tail_split) {
        // cost unit: one block pair, one side (the diagonal pair runs the directed formula: 1)
        struct item { sym_task t; int cost; };
        std::vector<item> plain;
        plain.reserve((size_t)(pairs / chunk + M + 1));
        for (int q = 0; q < M; ++q)
            for (int p0 = 0; p0 <= q; p0 += chunk) {
                const int p1 = std::min(p0 + chunk, q + 1);
                plain.push_back({sym_task{q, p0, p1, 0}, 2 * (p1 - p0) - (p1 == q + 1 ? 1 : 0)});
            }
        // Tail: the tasks that would run in the last two waves of the resident CTAs -- the last
        // 2 x slots x world entries of the cost-sorted list -- are cut into VSB_SYM_TAIL shorter
        // P-ranges (default 8, 0 = off), so that the ranks and the SMs finish within a small piece
        // of each other.  A cut task keeps its pairs' P-side slabs; each piece gets its own
        // Q-side slab.  (Single block pairs -- the tasks of mid sizes -- are NOT cut: cutting the
        // last two waves of them into their sub-rounds, each piece with its own slabs, was
        // measured slower, 2.113 -> 2.190 ms at 64 k bodies and 8.37 -> 8.50 ms at 128 k,
        // profiles/r2_midsize_subround_pieces_ab.txt: CTAs left alone on an SM run faster, and a
        // piece pays the full Q load and two full slabs for a fraction of the pairs.)
        std::vector<char> cut(plain.size(), 0);
        if (tail_split > 1) {
            int64_t left = std::min<int64_t>(2 * (int64_t)c->sm_count * mb * world, (int64_t)plain.size() / 2);
            const int maxc = 2 * chunk;
            for (int cst = 0; cst <= maxc && left > 0; ++cst)           // cheapest first = run last
                for (int64_t i = (int64_t)plain.size() - 1; i >= 0 && left > 0; --i)
                    if (plain[i].cost == cst) {
                        const sym_task &t = plain[i].t;
                        cut[i] = t.p1 - t.p0 >= 2;
                        --left;
                    }
        }
        std::vector<item> all;
        all.reserve(plain.size() + 16 * (size_t)c->sm_count * mb * world);
        for (size_t i = 0; i < plain.size(); ++i) {
            const sym_task &t = plain[i].t;
            if (!cut[i]) {
                all.push_back(plain[i]);
                continue;
            }
            const int len = (t.p1 - t.p0 + tail_split - 1) / tail_split;
            for (int a0 = t.p0; a0 < t.p1; a0 += len) {
                const int a1 = std::min(a0 + len, t.p1);
                all.push_back({sym_task{t.q, a0, a1, 0}, 2 * (a1 - a0) - (a1 == t.q + 1 ? 1 : 0)});
            }
        }
        // Q-side slab slots: the tasks of block Q in creation (P) order
        std::vector<int> qfirst((size_t)M + 1, 0);
        {
            int slot_run = 0, last_q = -1;
            for (item &it : all) {
                while (last_q < it.t.q) qfirst[++last_q] = slot_run;
                it.t.qslot = slot_run++;
            }
            while (last_q < M) qfirst[++last_q] = slot_run;
        }
        const int slot = qfirst[M];
        // descending cost (stable: creation order within a cost), then deal round-robin to ranks
        std::vector<int> order(all.size());
        for (size_t i = 0; i < all.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return all[x].cost > all[y].cost; });
        std::vector<sym_task> mine;
        for (size_t k = rank; k < order.size(); k += (size_t)world) mine.push_back(all[order[k]].t);
        const int nm = (int)mine.size();
        // several GPUs: per block X the unified indices of the slabs THIS rank's tasks write, in the
        // order the one-GPU reduction adds them (the block's slot list, then P-side by Q)
        std::vector<char> own((size_t)(pairs + slot + 1), 0);
        for (int i = 0; i < nm; ++i) {
            own[pairs + mine[i].qslot] = 1;
            for (int P = mine[i].p0; P < mine[i].p1; ++P)
                own[(int64_t)mine[i].q * (mine[i].q + 1) / 2 + P] = 1;
        }
        std::vector<int> own_first((size_t)M + 1, 0), own_list;
        for (int X = 0; X < M; ++X) {
            own_first[X] = (int)own_list.size();
            for (int sl = qfirst[X]; sl < qfirst[X + 1]; ++sl)
                if (own[pairs + sl]) own_list.push_back((int)(pairs + sl));
            for (int Q = X; Q < M; ++Q) {
                const int64_t sb = (int64_t)Q * (Q + 1) / 2 + X;
                if (own[sb]) own_list.push_back((int)sb);
            }
        }
        own_first[M] = (int)own_list.size();
        const int n_own = (int)own_list.size();
        const size_t tab_bytes = (sizeof(sym_task) * (size_t)(nm + 1) + 255) / 256 * 256;
        const size_t idx_bytes = (sizeof(int) * (size_t)(M + 1) + 255) / 256 * 256;
        const size_t own_bytes = (sizeof(int) * (size_t)(n_own + 1) + 255) / 256 * 256;
        VSB_TRY(vsb_buf_reserve(c->sym_tab, tab_bytes + 2 * idx_bytes + own_bytes + 1024));
        char *qf = (char *)c->sym_tab.p + tab_bytes;
        char *of = qf + idx_bytes, *ol = of + idx_bytes;
        VSB_CUDA(cudaMemcpyAsync(c->sym_tab.p, mine.data(), sizeof(sym_task) * (size_t)nm,
                                 cudaMemcpyHostToDevice, c->stream));
        VSB_CUDA(cudaMemcpyAsync(qf, qfirst.data(), sizeof(int) * (size_t)(M + 1), cudaMemcpyHostToDevice,
                                 c->stream));
        VSB_CUDA(cudaMemcpyAsync(of, own_first.data(), sizeof(int) * (size_t)(M + 1),
                                 cudaMemcpyHostToDevice, c->stream));
        if (n_own)
            VSB_CUDA(cudaMemcpyAsync(ol, own_list.data(), sizeof(int) * (size_t)n_own, cudaMemcpyHostToDevice,
                                     c->stream));
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        c->sym_own_off = (int64_t)(of - (char *)c->sym_tab.p);
        c->sym_ownlist_off = (int64_t)(ol - (char *)c->sym_tab.p);
        c->sym_tasks_n = npad;   // the table depends on the padded size only (a merge rarely changes it)
        c->sym_tail = tail_split;
        c->sym_S = S;
        c->sym_world = world;
        c->sym_rank = rank;
        c->sym_ntasks = nm;
        c->sym_nslots = slot;
        c->sym_qf_off = (int64_t)(qf - (char *)c->sym_tab.p);
    }
    const size_t slab = sizeof(double2) * (size_t)S;
    const size_t bytes_p = slab * (size_t)pairs, bytes_q = slab * (size_t)c->sym_nslots;
    VSB_TRY(vsb_buf_reserve(c->part, bytes_p + bytes_q));
    double2 *part_p = (double2 *)c->part.p;
    double2 *part_q = (double2 *)((char *)c->part.p + bytes_p);
    const sym_task *d_tasks = (const sym_task *)c->sym_tab.p;
    const int *d_qf = (const int *)((char *)c->sym_tab.p + c->sym_qf_off);
    const size_t smem = (size_t)S * (2 * sizeof(double2) + sizeof(double));
    if (c->sym_ntasks > 0) {
#define VSB_SYM_CASE(W, T, U, MB)                                                                   \
    if (warps == W && tpl == T && su == U && mb == MB) {                                                    \
        VSB_CUDA(cudaFuncSetAttribute(allpairs_sym_kernel<W, T, U, MB>,                           \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        allpairs_sym_kernel<W, T, U, MB><<<c->sym_ntasks, W * 32, smem, c->stream>>>(            \
            (const double2 *)c->pos, c->mass, d_tasks, part_p, part_q);                         \
        launched = true;                                                                        \
    }
        bool launched = false;
        VSB_SYM_CASE(1, 8, 1, 12)   // S =  256, 12 warps/SM
        VSB_SYM_CASE(2, 8, 1, 6)    // S =  512, 12 warps/SM
        VSB_SYM_CASE(4, 8, 1, 3)    // S = 1024, 12 warps/SM, CTA barrier
        VSB_SYM_CASE(4, 8, 2, 3)    // S = 1024, 12 warps/SM, pairwise barriers
        VSB_SYM_CASE(8, 8, 1, 1)    // S = 2048,  8 warps/SM
        VSB_SYM_CASE(8, 4, 1, 2)    // S = 1024, 4 targets per lane (first version of the kernel)
        VSB_SYM_CASE(16, 4, 1, 1)   // S = 2048, 4 targets per lane
#undef VSB_SYM_CASE
        VSB_CHECK(launched, VSB_ERR_ARG, "allpairs_sym: no kernel for shape %d,%d,%d,%d", warps, tpl, su, mb);
        c->launches += 1;
    }
    // slabs of other ranks' tasks are never read: no zeroing of the slab memory is needed
    const int *own_first = world > 1 ? (const int *)((char *)c->sym_tab.p + c->sym_own_off) : nullptr;
    const int *own_list = world > 1 ? (const int *)((char *)c->sym_tab.p + c->sym_ownlist_off) : nullptr;
    allpairs_sym_reduce_kernel<<<(unsigned)((n * 4 + 127) / 128), 128, 0, c->stream>>>(
        part_p, part_q, d_qf, own_first, own_list, S, M, n, gc, (double2 *)c->acc, (double2 *)c->vel,
        (double2 *)c->pos, (world == 1 && integrate) ? 1 : 0);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
