// vsb_parents_grid.cu -- sort-based candidate pass for A1 find_parents
// (phys_editor.py:350-354 / find_parent_one :36-48) on sm_100a.
//
// The reference scans, for every body, all more massive bodies in descending-mass order and
// keeps the LAST one whose COI contains it: N^2/2 point-in-disc tests.  In sorted-rank space the
// answer is  best[r] = max { k < r : |pos_r - pos_k|^2 < coi_k^2 },  an order-independent
// maximum, so it can be taken over any candidate set that contains every true match:
//   1. stats: bounding box of the (finite) positions, histogram of the binary exponents of
//      coi^2; the cell size h is the smallest power-of-two COI bound that leaves at most
//      PG_MAX_LARGE larger discs (1.5 * sqrt(box area / N) when that few bodies own a disc);
//   2. discs with coi^2 > h^2 go to the "large" list, every other disc (coi > 0) is
//      counting-sorted into a hashed uniform grid by the cell of its centre;
//   3. a disc of radius <= h can only contain points of its own or an adjacent cell, so each
//      body tests the discs of its 3x3 neighbourhood and then the large list.
// The predicate is the bit-exact one of fp_scan_kernel (vsb_pairs.cu: un-fused dx*dx + dy*dy,
// strict <, coi*coi un-fused), evaluated on every candidate, so best[] -- and with it
// parents[] -- is identical to the exhaustive scan's.  No input needs a fallback: non-finite
// positions never satisfy the predicate on either path and are kept out of the bounding box;
// a cloud collapsed into a few cells only costs time.  O(N) work, ~120 B/body of traffic.
// Compiled with -fmad=false.
#include "vsb_common.cuh"
#include "vsb_grid.cuh"
#include "vsb_orbit_math.cuh"

namespace {

constexpr int PG_MAX_LARGE = 4096;
constexpr int PG_EXP_BINS = 2048;              // biased exponent of coi^2
constexpr double PG_MAX_CELLS = 16777216.0;    // 2^24 cells per axis

struct __align__(32) pg_entry {   // one disc of the grid: a single 32-byte sector per candidate
    double x, y, coi2;
    int rank;
    int vis;      // to_kepler: the step that writes this COI (0 for find_parents: always visible)
};

struct pg_params {
    double xmin, ymin, inv_h, hh;   // hh: discs with coi^2 <= hh are binned
    unsigned mask;
    int n_large;
    int pad[2];
};

__device__ __forceinline__ int exp_bin(double v)   // v > 0
{
    return (int)((unsigned long long)__double_as_longlong(v) >> 52) & 0x7ff;
}

// partial[b] = {min x, min y, max x, max y} over finite coordinates; hist[e] += #discs
__global__ void __launch_bounds__(256)
pg_stats1_kernel(const double2 *__restrict__ spos, const double *__restrict__ scoi2, int64_t n,
                 double *__restrict__ partial, unsigned *__restrict__ hist)
{
    __shared__ double sh[4][256];
    __shared__ unsigned sh_hist[PG_EXP_BINS];
    for (int e = threadIdx.x; e < PG_EXP_BINS; e += 256) sh_hist[e] = 0;
    __syncthreads();
    double x0 = 1e308, y0 = 1e308, x1 = -1e308, y1 = -1e308;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (int64_t)gridDim.x * 256) {
        const double2 p = spos[k];
        if (isfinite(p.x) && isfinite(p.y)) {
            x0 = fmin(x0, p.x); y0 = fmin(y0, p.y);
            x1 = fmax(x1, p.x); y1 = fmax(y1, p.y);
        }
        const double c2 = scoi2[k];
        if (c2 > 0.0) atomicAdd(&sh_hist[exp_bin(c2)], 1u);
    }
    sh[0][threadIdx.x] = x0; sh[1][threadIdx.x] = y0; sh[2][threadIdx.x] = x1; sh[3][threadIdx.x] = y1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] = fmin(sh[0][threadIdx.x], sh[0][threadIdx.x + o]);
            sh[1][threadIdx.x] = fmin(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
            sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) partial[blockIdx.x * 4 + threadIdx.x] = sh[threadIdx.x][0];
    for (int e = threadIdx.x; e < PG_EXP_BINS; e += 256)
        if (sh_hist[e]) atomicAdd(&hist[e], sh_hist[e]);
}

__global__ void __launch_bounds__(256)
pg_stats2_kernel(const double *__restrict__ partial, int nblocks, int64_t n, unsigned mask,
                 const unsigned *__restrict__ hist, pg_params *P)
{
    __shared__ double sh[4][256];
    __shared__ unsigned sh_hist[PG_EXP_BINS];
    for (int e = threadIdx.x; e < PG_EXP_BINS; e += 256) sh_hist[e] = hist[e];
    double x0 = 1e308, y0 = 1e308, x1 = -1e308, y1 = -1e308;
    for (int b = threadIdx.x; b < nblocks; b += 256) {
        x0 = fmin(x0, partial[b * 4]); y0 = fmin(y0, partial[b * 4 + 1]);
        x1 = fmax(x1, partial[b * 4 + 2]); y1 = fmax(y1, partial[b * 4 + 3]);
    }
    sh[0][threadIdx.x] = x0; sh[1][threadIdx.x] = y0; sh[2][threadIdx.x] = x1; sh[3][threadIdx.x] = y1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] = fmin(sh[0][threadIdx.x], sh[0][threadIdx.x + o]);
            sh[1][threadIdx.x] = fmin(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
            sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    x0 = sh[0][0]; y0 = sh[1][0]; x1 = sh[2][0]; y1 = sh[3][0];
    if (!(x1 >= x0) || !(y1 >= y0)) x0 = y0 = x1 = y1 = 0.0;   // no finite position at all
    const double ex = x1 - x0, ey = y1 - y0;
    // density-based cell: 2-3 bodies per cell (a degenerate box falls back to its long side)
    double area = ex * ey;
    double h0 = 1.5 * sqrt(area / (double)n);
    if (!(h0 > 0.0) || !isfinite(h0)) h0 = 2.0 * fmax(ex, ey) / (double)n;
    if (!(h0 > 0.0) || !isfinite(h0)) h0 = 1.0;
    // smallest exponent bound with at most PG_MAX_LARGE discs above it: coi^2 < 2^(e+1-1023)
    unsigned above = 0;
    int e = PG_EXP_BINS - 1;
    while (e >= 0 && above + sh_hist[e] <= (unsigned)PG_MAX_LARGE) above += sh_hist[e--];
    // every disc in bins <= e is binned: hh >= 2^(e + 1 - 1023) (as a double: exponent field e+1)
    // The cells only have to be as large as the binned discs: with hashed buckets smaller cells
    // cost nothing (a query always reads 9 buckets) and keep dense regions cheap.  The density
    // scale h0 is the fallback when at most PG_MAX_LARGE bodies own a disc at all.
    double hh = h0 * h0;
    if (e >= 0)
        hh = e + 1 >= 0x7ff ? 1.7e308 : __longlong_as_double((long long)(e + 1) << 52);
    if (!isfinite(hh)) hh = 1.7e308;
    double h = sqrt(hh) * (1.0 + 1e-6);
    // at most 2^24 cells per axis
    h = fmax(h, fmax(ex, ey) / PG_MAX_CELLS);
    P->xmin = x0;
    P->ymin = y0;
    P->inv_h = 1.0 / h;
    P->hh = hh;
    P->mask = mask;
    P->n_large = 0;
}

__device__ __forceinline__ unsigned cell_of(const pg_params *P, double2 p)
{
    // non-finite coordinates give an arbitrary (but valid) bucket; such bodies never match
    double fx = floor((p.x - P->xmin) * P->inv_h), fy = floor((p.y - P->ymin) * P->inv_h);
    long long cx = isfinite(fx) ? (long long)fx : 0, cy = isfinite(fy) ? (long long)fy : 0;
    return cell_bucket(cx, cy, P->mask);
}

// bucket[k] = grid bucket of disc k, 0xffffffff for a large disc (appended to the list),
// 0xfffffffe for a body that owns no disc
__global__ void __launch_bounds__(256)
pg_count_kernel(const double2 *__restrict__ spos, const double *__restrict__ scoi2, int64_t n,
                pg_params *P, unsigned *__restrict__ bucket, unsigned *__restrict__ counts,
                int *__restrict__ large)
{
    int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const double c2 = scoi2[k];
    if (!(c2 > 0.0)) {
        bucket[k] = 0xfffffffeu;
        return;
    }
    if (c2 > P->hh) {
        int slot = atomicAdd(&P->n_large, 1);
        large[slot] = (int)k;   // capacity n; pg_stats2_kernel keeps the list short (<= PG_MAX_LARGE
                                // unless that many discs have an infinite COI)
        bucket[k] = 0xffffffffu;
        return;
    }
    unsigned b = cell_of(P, spos[k]);
    bucket[k] = b;
    atomicAdd(&counts[b], 1u);
}

__global__ void __launch_bounds__(256)
pg_scatter_kernel(const double2 *__restrict__ spos, const double *__restrict__ scoi2, int64_t n,
                  const unsigned *__restrict__ bucket, unsigned *__restrict__ cursor,
                  const int64_t *__restrict__ vis, pg_entry *__restrict__ ent)
{
    int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    unsigned b = bucket[k];
    if (b >= 0xfffffffeu) return;
    unsigned slot = atomicAdd(&cursor[b], 1u);
    const double2 p = spos[k];
    pg_entry e;
    e.x = p.x;
    e.y = p.y;
    e.coi2 = scoi2[k];
    e.rank = (int)k;
    e.vis = vis ? (int)vis[k] : 0;
    ent[slot] = e;
}

// thread <-> sorted rank r: discs of the 3x3 neighbourhood, then the large list
__global__ void __launch_bounds__(128)
pg_query_kernel(const pg_params *P, const unsigned *__restrict__ starts,
                const pg_entry *__restrict__ ent, const int *__restrict__ large,
                const double2 *__restrict__ spos, const double *__restrict__ scoi2, int64_t n,
                const int64_t *__restrict__ order, const int64_t *__restrict__ vis,
                int *__restrict__ best)
{
    __shared__ double2 lp[128];
    __shared__ double lc[128];
    __shared__ int lr[128];
    const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const bool live = r < n;
    const double2 me = spos[live ? r : 0];
    // to_kepler (vis != NULL): the body at rank r is step order[r] of the reference's loop and only
    // sees COIs written by earlier steps; find_parents: everything is visible
    const int mine = (vis && live) ? (int)order[r] : 0x7fffffff;
    int found = -1;
    if (live) {
        const double fx = floor((me.x - P->xmin) * P->inv_h), fy = floor((me.y - P->ymin) * P->inv_h);
        const long long cx = isfinite(fx) ? (long long)fx : 0, cy = isfinite(fy) ? (long long)fy : 0;
        unsigned seen[9];
        int nseen = 0;
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const unsigned b = cell_bucket(cx + ox, cy + oy, P->mask);
                bool dup = false;   // two neighbour cells hashing to one bucket: scan it once
                for (int k = 0; k < nseen; ++k) dup |= seen[k] == b;
                if (dup) continue;
                seen[nseen++] = b;
                const unsigned e = starts[b + 1];
                for (unsigned s = starts[b]; s < e; ++s) {
                    const pg_entry q = ent[s];
                    const int k = q.rank;
                    if (k >= r || k <= found || q.vis >= mine) continue;
                    const double dx = me.x - q.x, dy = me.y - q.y;
                    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                    if (d2 < q.coi2) found = k;
                }
            }
    }
    const int nl = P->n_large;
    for (int base = 0; base < nl; base += 128) {
        const int cnt = nl - base < 128 ? nl - base : 128;
        __syncthreads();
        if (threadIdx.x < cnt) {
            const int k = large[base + threadIdx.x];
            lr[threadIdx.x] = k;
            lp[threadIdx.x] = spos[k];
            lc[threadIdx.x] = scoi2[k];
        }
        __syncthreads();
        if (!live) continue;
        for (int t = 0; t < cnt; ++t) {
            const int k = lr[t];
            if (k >= r || k <= found || (vis && (int)vis[k] >= mine)) continue;
            const double dx = me.x - lp[t].x, dy = me.y - lp[t].y;
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < lc[t]) found = k;
        }
    }
    if (live) best[r] = found;
}

// ------------------------------------------------------------------ A7 on the same grid
// simplified_orbit_coi (phys_editor.py:325-346) is a sequential recurrence in descending mass:
// a body's COI needs its parent, the parent search needs the COIs of all earlier ranks.  The
// pair (parents, COIs) it ends with is the UNIQUE fixed point of
//     COIs    = f(parents)   (soc_coi_of per body), and
//     parents = g(COIs)      (find_parents: last earlier rank whose COI contains the body)
// (induction over the rank: rank 0 is fixed, and rank r's parent only reads earlier COIs).  So
// iterate the two maps over ALL bodies -- f is one thread per body, g is the O(N) candidate pass
// above -- until a round changes no parent: the sequential answer with the same expressions,
// reached after (depth of the hierarchy + ~2) rounds instead of N^2/2 tests.
__global__ void soc_gather_kernel(const int64_t *__restrict__ order, const double2 *__restrict__ pos,
                                  int64_t n, int64_t npad, double2 *__restrict__ spos,
                                  int *__restrict__ best)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npad) return;
    spos[k] = k < n ? pos[order[k]] : make_double2(VSB_FAR, VSB_FAR);
    best[k] = -1;
}

// COIs from the current parents (best[r]: sorted rank of the parent, -1 = the largest body)
__global__ void soc_coi_all_kernel(int64_t n, const int64_t *__restrict__ order, int64_t largest,
                                   const double *__restrict__ mass, const double2 *__restrict__ pos,
                                   const double2 *__restrict__ vel, double gc, double coi_coef,
                                   const int *__restrict__ best, double *__restrict__ scoi2,
                                   double *__restrict__ coi)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int64_t body = order[r];
    double c = 0.0;
    if (r > 0) {  // rank 0 keeps coi 0 (the loop starts at bodies_sorted[1:], :332)
        const int64_t parent = best[r] < 0 ? largest : order[best[r]];
        c = soc_coi_of(pos[body], vel[body], mass[body], pos[parent], vel[parent], mass[parent], gc,
                       coi_coef);
    }
    coi[body] = c;
    scoi2[r] = __dmul_rn(c, c);
}

__global__ void soc_diff_kernel(int64_t n, const int *__restrict__ a, const int *__restrict__ b,
                                int *__restrict__ changed)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = r < n && a[r] != b[r];
    if (__syncthreads_or(d) && threadIdx.x == 0) atomicOr(changed, 1);
}

// ------------------------------------------------------------------ N1 to_kepler, first loop
// convert.to_kepler (convert.py:205-225) is the same recurrence with its index quirks: step s
// (INDEX order) looks for the last candidate among the first rank[s] bodies of the mass order whose
// COI -- as far as it has been written -- contains pos[s], and then writes the COI of index
// R = rank[s] (the rank used as an index, :219-225).  coi[x] is therefore written by step order[x],
// and the candidate at sorted position l (index order[l]) is visible to step s iff
// order[order[l]] < s.  Same fixed point argument as above, by induction over s.
__global__ void tkf_init_kernel(const int64_t *__restrict__ order, const double2 *__restrict__ pos,
                                int64_t n, int64_t npad, double2 *__restrict__ spos,
                                int *__restrict__ rank, int *__restrict__ best,
                                int64_t *__restrict__ vis)
{
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= npad) return;
    best[l] = -1;
    if (l < n) {
        const int64_t p = order[l];
        spos[l] = pos[p];
        rank[p] = (int)l;
        vis[l] = order[p];
    } else {
        spos[l] = make_double2(VSB_FAR, VSB_FAR);
        vis[l] = 0x7fffffff;
    }
}

// step s with its current parent: ref_l[s] and the COI it writes (squared, by sorted position)
__global__ void tkf_coi_kernel(int64_t n, const int64_t *__restrict__ order,
                               const int *__restrict__ rank, const double *__restrict__ mass,
                               const double2 *__restrict__ pos, const double2 *__restrict__ vel,
                               double gc, double coi_coef, const int *__restrict__ best,
                               double *__restrict__ c2, int64_t *__restrict__ ref_l)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int R = rank[s];
    const int b = best[R];
    const int64_t ref = b < 0 ? 0 : order[b];
    ref_l[s] = ref;
    // quirk: the mass rank R is used as an index from here on (:219-225)
    const double c = soc_coi_of(pos[R], vel[R], mass[R], pos[ref], vel[ref], mass[ref], gc, coi_coef);
    c2[rank[R]] = __dmul_rn(c, c);
}

}  // namespace

namespace {

struct pg_view {
    pg_params *P;
    unsigned *starts;
    pg_entry *ent;
    int *large;
    unsigned M;
};

// counting sort of the discs (coi^2 > 0) of the n sorted ranks into the hashed grid of c->pgrid
int pg_build(vsb_ctx *c, const double2 *spos, const double *scoi2, int64_t n, const int64_t *vis,
             pg_view &v)
{
    // 2 buckets per body, at most 2^22 of them (the three-pass scan handles 4096 chunks of 1024):
    // beyond 2 M bodies the buckets just hold more entries
    unsigned M = 4096;
    while ((int64_t)M < 2 * n && M < (1u << 22)) M <<= 1;
    const int nchunks = (int)(M / 1024);
    const int sblocks = c->sm_count * 4;
    // layout of c->pgrid: params | partial | hist | counts[M] | starts[M+4] | cursor[M] |
    //                     chunk[4096] | bucket[n] | large[n] | ent[n]
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_par = take(sizeof(pg_params)), o_part = take(sizeof(double) * 4 * sblocks),
                 o_hist = take(4 * PG_EXP_BINS), o_cnt = take(4 * (size_t)M),
                 o_st = take(4 * ((size_t)M + 4)), o_cur = take(4 * (size_t)M), o_chk = take(4 * 4096),
                 o_bkt = take(4 * (size_t)n), o_lrg = take(4 * (size_t)n),
                 o_ent = take(sizeof(pg_entry) * (size_t)n);
    VSB_TRY(vsb_buf_reserve(c->pgrid, off));
    char *base = (char *)c->pgrid.p;
    pg_params *P = (pg_params *)(base + o_par);
    double *partial = (double *)(base + o_part);
    unsigned *hist = (unsigned *)(base + o_hist), *counts = (unsigned *)(base + o_cnt),
             *starts = (unsigned *)(base + o_st), *cursor = (unsigned *)(base + o_cur),
             *chunk = (unsigned *)(base + o_chk), *bucket = (unsigned *)(base + o_bkt);
    int *large = (int *)(base + o_lrg);
    pg_entry *ent = (pg_entry *)(base + o_ent);
    const unsigned g = (unsigned)((n + 255) / 256);

    // hist and counts are adjacent: one memset
    VSB_CUDA(cudaMemsetAsync(hist, 0, (o_cnt - o_hist) + 4 * (size_t)M, c->stream));
    pg_stats1_kernel<<<sblocks, 256, 0, c->stream>>>(spos, scoi2, n, partial, hist);
    pg_stats2_kernel<<<1, 256, 0, c->stream>>>(partial, sblocks, n, M - 1, hist, P);
    pg_count_kernel<<<g, 256, 0, c->stream>>>(spos, scoi2, n, P, bucket, counts, large);
    bp_scan_reduce_kernel<<<nchunks, 256, 0, c->stream>>>(counts, chunk);
    bp_scan_chunks_kernel<<<1, 1024, 0, c->stream>>>(chunk, nchunks);
    bp_scan_apply_kernel<<<nchunks, 256, 0, c->stream>>>(counts, chunk, starts, cursor, M);
    pg_scatter_kernel<<<g, 256, 0, c->stream>>>(spos, scoi2, n, bucket, cursor, vis, ent);
    c->launches += 7;
    VSB_CUDA(cudaGetLastError());
    v.P = P;
    v.starts = starts;
    v.ent = ent;
    v.large = large;
    v.M = M;
    return VSB_OK;
}

}  // namespace

// best[r] for ALL sorted ranks from the staging arrays of fp_gather_kernel (spos, scoi2).
int vsb_launch_find_parents_grid(vsb_ctx *c, const double2 *spos, const double *scoi2, int *best)
{
    const int64_t n = c->n;
    pg_view v;
    VSB_TRY(pg_build(c, spos, scoi2, n, nullptr, v));
    pg_query_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(v.P, v.starts, v.ent, v.large,
                                                                         spos, scoi2, n, nullptr,
                                                                         nullptr, best);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// A7 by fixed-point rounds (see soc_coi_all_kernel).  *converged_out = 0 when max_rounds did not
// suffice (a pathological chain of nested bodies): the caller then runs the sequential sweep.
// Uses c->sorted (spos | scoi2 | best_a | best_b) and c->scratch (flag); writes c->coi.
int vsb_launch_soc_fixed_point(vsb_ctx *c, const int64_t *d_order, int64_t largest, double gc,
                               double coi_coef, int max_rounds, int *converged_out, int *rounds_out)
{
    const int64_t n = c->n, npad = c->npad;
    *converged_out = 0;
    VSB_TRY(vsb_buf_reserve(c->sorted, (size_t)npad * (16 + 8 + 4 + 4)));
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    char *base = (char *)c->sorted.p;
    double2 *spos = (double2 *)base;
    double *scoi2 = (double *)(base + (size_t)npad * 16);
    int *best = (int *)(base + (size_t)npad * 24), *best_new = (int *)(base + (size_t)npad * 28);
    int *d_changed = (int *)((char *)c->scratch.p + 32);   // bytes 0-7 hold the caller's argmax result
    const unsigned g = (unsigned)((n + 255) / 256);
    soc_gather_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(
        d_order, (const double2 *)c->pos, n, npad, spos, best);
    c->launches += 1;
    for (int round = 1; round <= max_rounds; ++round) {
        soc_coi_all_kernel<<<g, 256, 0, c->stream>>>(n, d_order, largest, c->mass,
                                                     (const double2 *)c->pos, (const double2 *)c->vel,
                                                     gc, coi_coef, best, scoi2, c->coi);
        VSB_TRY(vsb_launch_find_parents_grid(c, spos, scoi2, best_new));
        VSB_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int), c->stream));
        soc_diff_kernel<<<g, 256, 0, c->stream>>>(n, best, best_new, d_changed);
        c->launches += 2;
        int h_changed = 0;
        VSB_CUDA(cudaMemcpyAsync(&h_changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        if (rounds_out) *rounds_out = round;
        if (!h_changed) {
            *converged_out = 1;     // COIs of this round belong to parents that did not move
            return VSB_OK;
        }
        int *t = best;
        best = best_new;
        best_new = t;
    }
    return VSB_OK;
}

// N1: the first loop of to_kepler by fixed-point rounds: d_ref[s] for every body.  *converged_out
// = 0 when max_rounds did not suffice (the caller runs the blocked sweep).  Uses c->sorted.
int vsb_launch_to_kepler_fixed_point(vsb_ctx *c, const int64_t *d_order, double gc, double coi_coef,
                                     int max_rounds, int64_t *d_ref, int *converged_out)
{
    const int64_t n = c->n, npad = c->npad;
    *converged_out = 0;
    VSB_TRY(vsb_buf_reserve(c->sorted, (size_t)npad * (16 + 8 + 8 + 4 + 4 + 4)));
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    char *base = (char *)c->sorted.p;
    double2 *spos = (double2 *)base;
    double *c2 = (double *)(base + (size_t)npad * 16);
    int64_t *vis = (int64_t *)(base + (size_t)npad * 24);
    int *rank = (int *)(base + (size_t)npad * 32);
    int *best = (int *)(base + (size_t)npad * 36), *best_new = (int *)(base + (size_t)npad * 40);
    int *d_changed = (int *)((char *)c->scratch.p + 32);
    const double2 *pos = (const double2 *)c->pos, *vel = (const double2 *)c->vel;
    const unsigned g = (unsigned)((n + 255) / 256);
    tkf_init_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(d_order, pos, n, npad, spos,
                                                                          rank, best, vis);
    c->launches += 1;
    for (int round = 1; round <= max_rounds; ++round) {
        tkf_coi_kernel<<<g, 256, 0, c->stream>>>(n, d_order, rank, c->mass, pos, vel, gc, coi_coef, best,
                                                 c2, d_ref);
        pg_view v;
        VSB_TRY(pg_build(c, spos, c2, n, vis, v));
        pg_query_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(
            v.P, v.starts, v.ent, v.large, spos, c2, n, d_order, vis, best_new);
        VSB_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int), c->stream));
        soc_diff_kernel<<<g, 256, 0, c->stream>>>(n, best, best_new, d_changed);
        c->launches += 3;
        int h_changed = 0;
        VSB_CUDA(cudaMemcpyAsync(&h_changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        if (!h_changed) {
            *converged_out = 1;     // d_ref of this round belongs to parents that did not move
            return VSB_OK;
        }
        int *t = best;
        best = best_new;
        best_new = t;
    }
    return VSB_OK;
}
