// vsb_broadphase.cu -- sort-based candidate pass for B1 check_collision
// (phys_editor.py:547-551 / check_collision_one :86-97) on sm_100a.
//
// The reference tests all N(N-1)/2 pairs; the answer it returns is the
// lexicographically smallest overlapping index pair.  This pass finds the same
// pair from O(N) candidates:
//   1. stats: mean radius -> r_cut = 2 * mean, cell size h = 2 * r_cut (1 + 1e-6),
//      bounding box;
//   2. bodies with rad > r_cut ("large", a handful: e.g. the central body) go to a
//      list, every other body is counted into a hashed uniform grid (counting sort
//      by cell: histogram -> exclusive scan -> scatter into cell-sorted arrays);
//   3. small-small: two small bodies can only overlap if their cells are equal or
//      adjacent (|dx|,|dy| <= r_i + r_j <= h), so each body tests the 3x3
//      neighbourhood, exact predicate, partner index > own index only;
//   4. large-any: every body is tested against the large list;
//   5. global atomicMin on key (i << 32) | j.
// The predicate is the bit-exact one of vsb_pairs.cu (un-fused squares, IEEE
// sqrt), evaluated for every candidate, and the winner is an order-independent
// minimum, so the result is identical to the exhaustive scan.  If the input does
// not suit a grid (too many large bodies, absurd extent) the pass reports
// "fallback" and the caller runs the exhaustive kernel instead.
// Compiled with -fmad=false.
#include "vsb_common.cuh"
#include "vsb_grid.cuh"

namespace {

constexpr int BP_MAX_LARGE = 4096;
constexpr double BP_CELL_LIMIT = 16777216.0;  // 2^24 cells per axis

struct bp_params {
    double xmin, ymin, inv_h, rcut;
    unsigned mask;        // M - 1
    int n_large;
    int fallback;         // 1: grid unsuitable, run the exhaustive scan
    int pad;
};

__device__ __forceinline__ bool collide_exact(double2 a, double2 b, double ra, double rb)
{
    // a = pos[body1] (smaller index), b = pos[body2]; phys_editor.py:90-92
    double dx = a.x - b.x, dy = a.y - b.y;
    double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    return d <= __dadd_rn(ra, rb);
}

// partial[b] = {sum rad, min x, min y, max x, max y}
__global__ void __launch_bounds__(256)
bp_stats1_kernel(const double2 *__restrict__ pos, const double *__restrict__ rad, int64_t n,
                 double *__restrict__ partial)
{
    __shared__ double sh[5][256];
    double s = 0, x0 = 1e308, y0 = 1e308, x1 = -1e308, y1 = -1e308;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double2 p = pos[i];
        s += rad[i];
        x0 = fmin(x0, p.x); y0 = fmin(y0, p.y);
        x1 = fmax(x1, p.x); y1 = fmax(y1, p.y);
    }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = x0; sh[2][threadIdx.x] = y0;
    sh[3][threadIdx.x] = x1; sh[4][threadIdx.x] = y1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] = fmin(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
            sh[2][threadIdx.x] = fmin(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
            sh[4][threadIdx.x] = fmax(sh[4][threadIdx.x], sh[4][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 5) partial[blockIdx.x * 5 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void __launch_bounds__(256)
bp_stats2_kernel(const double *__restrict__ partial, int nblocks, int64_t n, unsigned mask,
                 bp_params *P, unsigned long long *key)
{
    // fixed-shape tree reduction of the per-CTA partials: deterministic r_cut
    __shared__ double sh[5][256];
    double s = 0, x0 = 1e308, y0 = 1e308, x1 = -1e308, y1 = -1e308;
    for (int b = threadIdx.x; b < nblocks; b += 256) {
        s += partial[b * 5];
        x0 = fmin(x0, partial[b * 5 + 1]); y0 = fmin(y0, partial[b * 5 + 2]);
        x1 = fmax(x1, partial[b * 5 + 3]); y1 = fmax(y1, partial[b * 5 + 4]);
    }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = x0; sh[2][threadIdx.x] = y0;
    sh[3][threadIdx.x] = x1; sh[4][threadIdx.x] = y1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] = fmin(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
            sh[2][threadIdx.x] = fmin(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
            sh[4][threadIdx.x] = fmax(sh[4][threadIdx.x], sh[4][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    s = sh[0][0]; x0 = sh[1][0]; y0 = sh[2][0]; x1 = sh[3][0]; y1 = sh[4][0];
    double rcut = 2.0 * (s / (double)n);
    double h = 2.0 * rcut * (1.0 + 1e-6);
    P->xmin = x0;
    P->ymin = y0;
    P->rcut = rcut;
    P->mask = mask;
    P->n_large = 0;
    bool ok = h > 0.0 && isfinite(h) && isfinite(x0) && isfinite(y0) && isfinite(x1) && isfinite(y1) &&
              (x1 - x0) / h < BP_CELL_LIMIT && (y1 - y0) / h < BP_CELL_LIMIT;
    P->inv_h = ok ? 1.0 / h : 0.0;
    P->fallback = ok ? 0 : 1;
    *key = ~0ull;
}

// bucket[i] = grid bucket of body i, or 0xffffffff for a large body (appended to the list)
__global__ void __launch_bounds__(256)
bp_count_kernel(const double2 *__restrict__ pos, const double *__restrict__ rad, int64_t n,
                bp_params *P, unsigned *__restrict__ bucket, unsigned *__restrict__ counts,
                int *__restrict__ large)
{
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n || P->fallback) return;
    if (rad[i] > P->rcut) {
        int slot = atomicAdd(&P->n_large, 1);
        if (slot < BP_MAX_LARGE) large[slot] = (int)i;
        bucket[i] = 0xffffffffu;
        return;
    }
    double2 p = pos[i];
    long long cx = (long long)floor((p.x - P->xmin) * P->inv_h);
    long long cy = (long long)floor((p.y - P->ymin) * P->inv_h);
    unsigned b = cell_bucket(cx, cy, P->mask);
    bucket[i] = b;
    atomicAdd(&counts[b], 1u);
}

__global__ void __launch_bounds__(256)
bp_scatter_kernel(const double2 *__restrict__ pos, const double *__restrict__ rad, int64_t n,
                  const bp_params *P, const unsigned *__restrict__ bucket,
                  unsigned *__restrict__ cursor, double2 *__restrict__ spos,
                  double *__restrict__ srad, int *__restrict__ sidx)
{
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n || P->fallback) return;
    unsigned b = bucket[i];
    if (b == 0xffffffffu) return;
    unsigned slot = atomicAdd(&cursor[b], 1u);
    spos[slot] = pos[i];
    srad[slot] = rad[i];
    sidx[slot] = (int)i;
}

// small-small candidates: thread <-> slot of the cell-sorted arrays
__global__ void __launch_bounds__(128)
bp_query_kernel(const bp_params *P, const unsigned *__restrict__ starts,
                const double2 *__restrict__ spos, const double *__restrict__ srad,
                const int *__restrict__ sidx, unsigned M, unsigned long long *key)
{
    if (P->fallback) return;
    const unsigned total = starts[M];
    const unsigned s = blockIdx.x * 128 + threadIdx.x;
    if (s >= total) return;
    const int i = sidx[s];
    const double2 pi = spos[s];
    const double ri = srad[s];
    const long long cx = (long long)floor((pi.x - P->xmin) * P->inv_h);
    const long long cy = (long long)floor((pi.y - P->ymin) * P->inv_h);
    int best = 0x7fffffff;
    unsigned seen[9];
    int nseen = 0;
    for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
            const unsigned b = cell_bucket(cx + ox, cy + oy, P->mask);
            bool dup = false;  // two neighbour cells hashing to one bucket: scan it once
            for (int k = 0; k < nseen; ++k) dup |= seen[k] == b;
            if (dup) continue;
            seen[nseen++] = b;
            const unsigned e = starts[b + 1];
            for (unsigned k = starts[b]; k < e; ++k) {
                const int j = sidx[k];
                if (j > i && j < best && collide_exact(pi, spos[k], ri, srad[k])) best = j;
            }
        }
    if (best != 0x7fffffff)
        atomicMin(key, ((unsigned long long)(unsigned)i << 32) | (unsigned long long)(unsigned)best);
}

// large-any candidates: thread <-> body, loop over the large list
__global__ void __launch_bounds__(256)
bp_large_kernel(const bp_params *P, const int *__restrict__ large,
                const double2 *__restrict__ pos, const double *__restrict__ rad, int64_t n,
                unsigned long long *key)
{
    if (P->fallback) return;
    const int nl = P->n_large < BP_MAX_LARGE ? P->n_large : BP_MAX_LARGE;
    if (nl == 0) return;
    __shared__ double2 lp[256];
    __shared__ double lr[256];
    __shared__ int li[256];
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool live = k < n;
    const double2 pk = pos[live ? k : 0];
    const double rk = rad[live ? k : 0];
    unsigned long long best = ~0ull;
    for (int base = 0; base < nl; base += 256) {
        const int cnt = nl - base < 256 ? nl - base : 256;
        __syncthreads();
        if (threadIdx.x < cnt) {
            int l = large[base + threadIdx.x];
            li[threadIdx.x] = l;
            lp[threadIdx.x] = pos[l];
            lr[threadIdx.x] = rad[l];
        }
        __syncthreads();
        if (!live) continue;
        for (int t = 0; t < cnt; ++t) {
            const int l = li[t];
            if (l == (int)k) continue;
            bool hit;
            unsigned long long kk;
            if ((int)k < l) {
                hit = collide_exact(pk, lp[t], rk, lr[t]);
                kk = ((unsigned long long)(unsigned)k << 32) | (unsigned)l;
            } else {
                hit = collide_exact(lp[t], pk, lr[t], rk);
                kk = ((unsigned long long)(unsigned)l << 32) | (unsigned)k;
            }
            if (hit && kk < best) best = kk;
        }
    }
    if (best != ~0ull) atomicMin(key, best);
}

// out3 = {body_del, body_add, fallback}
__global__ void bp_finalize_kernel(const bp_params *P, const unsigned long long *key,
                                   const double *__restrict__ mass, int64_t *out3)
{
    out3[2] = (P->fallback || P->n_large > BP_MAX_LARGE) ? 1 : 0;
    unsigned long long k = *key;
    if (k == ~0ull) {
        out3[0] = -1;
        out3[1] = -1;
        return;
    }
    int64_t i = (int64_t)(k >> 32), j = (int64_t)(k & 0xffffffffull);
    if (mass[i] <= mass[j]) {
        out3[0] = i;
        out3[1] = j;
    } else {
        out3[0] = j;
        out3[1] = i;
    }
}

}  // namespace

// Runs the candidate pass; result {del, add, fallback} lands in c->scratch.p (3 x int64).
int vsb_launch_broadphase(vsb_ctx *c)
{
    const int64_t n = c->n;
    // 2 buckets per body, at most 2^22 of them (the three-pass scan handles 4096 chunks of 1024)
    unsigned M = 4096;
    while ((int64_t)M < 2 * n && M < (1u << 22)) M <<= 1;
    const int nchunks = (int)(M / 1024);
    const int sblocks = c->sm_count * 4;
    // layout of c->sorted: params | partial | counts[M] | starts[M+4] | cursor[M] | chunk[4096] |
    //                      bucket[n] | large[4096] | spos[n] | srad[n] | sidx[n]
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_par = take(sizeof(bp_params)), o_part = take(sizeof(double) * 5 * sblocks);
    const size_t o_cnt = take(4 * (size_t)M), o_st = take(4 * ((size_t)M + 4)),
                 o_cur = take(4 * (size_t)M), o_chk = take(4 * 4096), o_bkt = take(4 * (size_t)n),
                 o_lrg = take(4 * BP_MAX_LARGE), o_sp = take(16 * (size_t)n),
                 o_sr = take(8 * (size_t)n), o_si = take(4 * (size_t)n);
    VSB_TRY(vsb_buf_reserve(c->sorted, off));
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    char *base = (char *)c->sorted.p;
    bp_params *P = (bp_params *)(base + o_par);
    double *partial = (double *)(base + o_part);
    unsigned *counts = (unsigned *)(base + o_cnt), *starts = (unsigned *)(base + o_st),
             *cursor = (unsigned *)(base + o_cur), *chunk = (unsigned *)(base + o_chk),
             *bucket = (unsigned *)(base + o_bkt);
    int *large = (int *)(base + o_lrg), *sidx = (int *)(base + o_si);
    double2 *spos = (double2 *)(base + o_sp);
    double *srad = (double *)(base + o_sr);
    const double2 *pos = (const double2 *)c->pos;
    const unsigned g = (unsigned)((n + 255) / 256);

    VSB_CUDA(cudaMemsetAsync(counts, 0, 4 * (size_t)M, c->stream));
    bp_stats1_kernel<<<sblocks, 256, 0, c->stream>>>(pos, c->rad, n, partial);
    bp_stats2_kernel<<<1, 256, 0, c->stream>>>(partial, sblocks, n, M - 1, P, c->d_key);
    bp_count_kernel<<<g, 256, 0, c->stream>>>(pos, c->rad, n, P, bucket, counts, large);
    bp_scan_reduce_kernel<<<nchunks, 256, 0, c->stream>>>(counts, chunk);
    bp_scan_chunks_kernel<<<1, 1024, 0, c->stream>>>(chunk, nchunks);
    bp_scan_apply_kernel<<<nchunks, 256, 0, c->stream>>>(counts, chunk, starts, cursor, M);
    bp_scatter_kernel<<<g, 256, 0, c->stream>>>(pos, c->rad, n, P, bucket, cursor, spos, srad, sidx);
    bp_query_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(P, starts, spos, srad, sidx,
                                                                         M, c->d_key);
    bp_large_kernel<<<g, 256, 0, c->stream>>>(P, large, pos, c->rad, n, c->d_key);
    bp_finalize_kernel<<<1, 1, 0, c->stream>>>(P, c->d_key, c->mass, (int64_t *)c->scratch.p);
    c->launches += 10;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
