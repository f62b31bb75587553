// vsb_pairs.cu -- the two O(N^2) predicate scans of the reference step (sm_100a):
//   A1 find_parents   (phys_editor.py:36-48, :350-354)  point-in-COI, last match in mass order
//   B1 check_collision (phys_editor.py:86-97, :547-551) first overlapping pair (lexicographic)
// Both are bit-exact restatements: numba emits un-fused dx*dx + dy*dy, an IEEE
// sqrt and plain compares (SURVEY 7 "bit-exact collision predicate"), so this
// file is compiled with -fmad=false and uses the _rn intrinsics where it matters.
//
// Same streaming shape as the gravity kernel: grid (target blocks, source
// chunks), sources staged in shared memory by 1-D TMA bulk copies (UBLKCP) on
// mbarriers, every lane reading the same source (LDS broadcast); results are
// combined with order-independent max / min reductions, hence deterministic.
#include <stdlib.h>

#include "vsb_common.cuh"

namespace {

// ------------------------------------------------------------------ A1 find_parents
// Stage sources in descending-mass order: sorted_pos[k] = pos[order[k]],
// sorted_coi2[k] = coi*coi (un-fused), rank[order[k]] = k, best[k] = -1.
// Padding ranks get coi2 = -1 (never matches) and a far position.
__global__ void fp_gather_kernel(const int64_t *__restrict__ order, const double2 *__restrict__ pos,
                                 const double *__restrict__ coi, int64_t n, int64_t npad,
                                 double2 *__restrict__ spos, double *__restrict__ scoi2,
                                 int64_t *__restrict__ rank, int *__restrict__ best)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npad) return;
    if (k < n) {
        int64_t p = order[k];
        spos[k] = pos[p];
        double c = coi[p];
        scoi2[k] = __dmul_rn(c, c);
        rank[p] = k;
        best[k] = -1;
    } else {
        spos[k] = make_double2(VSB_FAR, VSB_FAR);
        scoi2[k] = -1.0;
        best[k] = -1;
    }
}

// Thread <-> target rank r; sources are ranks k < r.  best[r] = max matching k.
__global__ void __launch_bounds__(PB_BLOCK)
fp_scan_kernel(const double2 *__restrict__ spos, const double *__restrict__ scoi2, int64_t n,
               int64_t chunk_tiles, int *__restrict__ best, int part, int parts)
{
    // row blocks part, part+parts, ... of the mass-sorted bodies (the scan is triangular)
    const int64_t r0 = ((int64_t)blockIdx.x * parts + part) * PB_BLOCK;
    const int64_t r = r0 + threadIdx.x;
    int64_t t0 = (int64_t)blockIdx.y * chunk_tiles;
    int64_t t1 = t0 + chunk_tiles;
    // sources needed by this CTA: k < r0 + PB_BLOCK
    const int64_t tmax = (r0 + PB_BLOCK + PB_TJ - 1) / PB_TJ;
    if (t1 > tmax) t1 = tmax;
    if (t0 >= t1) return;
    const double2 me = spos[r < n ? r : 0];
    int found = -1;
    stream_tiles(spos, scoi2, t0, t1, [&](const double2 *sp, const double *sv, int64_t base) {
        const int lim = (int)((r - base) < PB_TJ ? (r - base) : PB_TJ);  // k < r
#pragma unroll 4
        for (int j = 0; j < PB_TJ; ++j) {
            double dx = me.x - sp[j].x;
            double dy = me.y - sp[j].y;
            double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < sv[j] && j < lim) found = (int)base + j;
        }
        return true;
    });
    if (r < n && found >= 0) atomicMax(&best[r], found);
}

__global__ void fp_finalize_kernel(const int64_t *__restrict__ order, const int *__restrict__ best,
                                   int64_t n, int64_t *__restrict__ parents)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int64_t body = order[k];
    int b = best[k];
    // `if body_s:` -- body index 0 always gets parent 0 (phys_editor.py:38,48)
    parents[body] = (body == 0 || b < 0) ? 0 : order[b];
}

// ------------------------------------------------------------------ B1 check_collision
constexpr double CC_MARGIN = 1.0 + 0x1p-40;  // filter slack, far above every rounding involved

// per-source staged value: rad*(1+margin) so that (ri_m + rj_m)^2 upper-bounds the
// exact threshold; the exact predicate is re-evaluated only for filter hits.
__global__ void cc_prepare_kernel(const double *__restrict__ rad, int64_t n, int64_t npad,
                                  double *__restrict__ radm, unsigned long long *key)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) *key = ~0ull;
    if (k >= npad) return;
    radm[k] = k < n ? rad[k] * CC_MARGIN : __longlong_as_double(0x7ff8000000000000ll);  // NaN: padding never passes
}

__device__ __forceinline__ bool collide_exact(double2 a, double2 b, double ra, double rb)
{
    // check_collision_one, phys_editor.py:90-92: rel = pos[body1]-pos[body2];
    // sqrt(rel0*rel0 + rel1*rel1) <= rad[body1] + rad[body2]
    double dx = a.x - b.x, dy = a.y - b.y;
    double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    return d <= __dadd_rn(ra, rb);
}

__global__ void __launch_bounds__(PB_BLOCK)
cc_scan_kernel(const double2 *__restrict__ pos, const double *__restrict__ rad,
               const double *__restrict__ radm, int64_t n, int64_t chunk_tiles,
               unsigned long long *__restrict__ key, int part, int parts)
{
    // row blocks part, part+parts, ... (multi-GPU: one part per rank; the triangular cost of the
    // scan is balanced by dealing the blocks round-robin)
    const int64_t i0 = ((int64_t)blockIdx.x * parts + part) * PB_BLOCK;
    const int64_t i = i0 + threadIdx.x;
    int64_t t0 = (int64_t)blockIdx.y * chunk_tiles;
    int64_t t1 = t0 + chunk_tiles;
    const int64_t ntiles = (n + PB_TJ - 1) / PB_TJ;
    if (t1 > ntiles) t1 = ntiles;
    // sources j > i >= i0: tiles before the one holding i0+1 are not needed
    const int64_t tmin = (i0 + 1) / PB_TJ;
    if (t0 < tmin) t0 = tmin;
    if (t0 >= t1) return;
    // a smaller row already collided: nothing this CTA finds can win.  The decision must be
    // CTA-uniform (one thread reads the key): the key changes under our feet, and a CTA that
    // lost only some of its threads -- thread 0 among them -- would wait on mbarriers nobody
    // initialised.
    __shared__ int s_skip;
    if (threadIdx.x == 0)
        s_skip = (*(volatile unsigned long long *)key >> 32) < (unsigned long long)i0 ? 1 : 0;
    __syncthreads();
    if (s_skip) return;

    const bool live = i < n - 1;
    const double2 me = pos[live ? i : 0];
    const double my_r = rad[live ? i : 0];
    const double my_rm = live ? my_r * CC_MARGIN : __longlong_as_double(0x7ff8000000000000ll);
    long long hit = -1;
    stream_tiles(pos, radm, t0, t1, [&](const double2 *sp, const double *sv, int64_t base) {
        if (hit < 0) {
#pragma unroll 4
            for (int j = 0; j < PB_TJ; ++j) {
                double dx = me.x - sp[j].x;
                double dy = me.y - sp[j].y;
                double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                double rs = my_rm + sv[j];
                if (d2 <= rs * rs) {  // rare: conservative filter hit -> exact test
                    int64_t jj = base + j;
                    if (jj > i && jj < n && hit < 0 &&
                        collide_exact(me, sp[j], my_r, rad[jj]))
                        hit = jj;
                }
            }
        }
        // stop streaming once every row of the CTA is decided or a smaller row won
        bool more = hit < 0 && live;
        if ((*(volatile unsigned long long *)key >> 32) < (unsigned long long)i0) more = false;
        return more;
    });
    if (hit >= 0) atomicMin(key, ((unsigned long long)i << 32) | (unsigned long long)hit);
}

// (i,j) -> (body_del, body_add): the lighter body is deleted, ties delete i
// (phys_editor.py:93-97).
__global__ void cc_finalize_kernel(const unsigned long long *key, const double *__restrict__ mass,
                                   int64_t *out2)
{
    unsigned long long k = *key;
    if (k == ~0ull) {
        out2[0] = -1;
        out2[1] = -1;
        return;
    }
    int64_t i = (int64_t)(k >> 32), j = (int64_t)(k & 0xffffffffull);
    if (mass[i] <= mass[j]) {
        out2[0] = i;
        out2[1] = j;
    } else {
        out2[0] = j;
        out2[1] = i;
    }
}

int chunk_tiles_for(int64_t tiles, int64_t row_blocks, int sm_count)
{
    // enough CTAs to fill the machine ~8x over, but never more chunks than tiles
    int64_t want = ((int64_t)sm_count * 64 + row_blocks - 1) / row_blocks;
    if (want < 1) want = 1;
    if (want > tiles) want = tiles;
    if (want > 64) want = 64;
    return (int)((tiles + want - 1) / want);
}

}  // namespace

// d_order: device array (n) of body indices in descending-mass order.
// Part `part` of `parts` of the scan: best[r] (sorted rank of the smallest enclosing COI, -1 = none)
// for this part's row blocks; the other rows keep -1, so an elementwise MAX over the parts gives
// the full array.  parts == 1 is the whole scan.
int vsb_launch_find_parents_part(vsb_ctx *c, const int64_t *d_order, int part, int parts)
{
    const int64_t n = c->n, npad = c->npad;
    if (n <= 0) return VSB_OK;
    // sorted staging: spos (npad*16) | scoi2 (npad*8) | rank (npad*8) | best (npad*4)
    VSB_TRY(vsb_buf_reserve(c->sorted, (size_t)npad * (16 + 8 + 8 + 4)));
    char *base = (char *)c->sorted.p;
    double2 *spos = (double2 *)base;
    double *scoi2 = (double *)(base + (size_t)npad * 16);
    int64_t *rank = (int64_t *)(base + (size_t)npad * 24);
    int *best = (int *)(base + (size_t)npad * 32);
    unsigned g = (unsigned)((npad + 255) / 256);
    fp_gather_kernel<<<g, 256, 0, c->stream>>>(d_order, (const double2 *)c->pos, c->coi, n, npad,
                                               spos, scoi2, rank, best);
    c->launches += 1;
    // From 4096 bodies up the sort-based candidate pass gives the same best[] in O(N); it is not
    // split (every part computes all rows; the MAX over parts is then a no-op).
    // VSB_PARENTS=scan|grid overrides.
    const char *env = getenv("VSB_PARENTS");
    const bool grid = env ? env[0] == 'g' : n >= 4096;
    if (grid) return vsb_launch_find_parents_grid(c, spos, scoi2, best);
    const int64_t row_blocks = (n + PB_BLOCK - 1) / PB_BLOCK;
    const int64_t mine = (row_blocks - part + parts - 1) / parts;
    if (mine > 0) {
        const int64_t tiles = npad / PB_TJ;
        const int ct = chunk_tiles_for(tiles, mine, c->sm_count);
        dim3 grid((unsigned)mine, (unsigned)((tiles + ct - 1) / ct));
        fp_scan_kernel<<<grid, PB_BLOCK, 0, c->stream>>>(spos, scoi2, n, ct, best, part, parts);
        c->launches += 1;
    }
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// best[] -> parents[] in body index space
int vsb_launch_find_parents_finish(vsb_ctx *c, const int64_t *d_order)
{
    const int64_t n = c->n, npad = c->npad;
    if (n <= 0) return VSB_OK;
    const int *best = (const int *)((char *)c->sorted.p + (size_t)npad * 32);
    fp_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_order, best, n,
                                                                          c->parents);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_find_parents(vsb_ctx *c, const int64_t *d_order)
{
    VSB_TRY(vsb_launch_find_parents_part(c, d_order, 0, 1));
    return vsb_launch_find_parents_finish(c, d_order);
}

// Exhaustive scan of the row blocks of one part (part of parts); the local minimum key lands in
// c->d_key.  parts == 1 is the whole scan.
int vsb_launch_check_collision_part(vsb_ctx *c, int part, int parts)
{
    const int64_t n = c->n, npad = c->npad;
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    VSB_TRY(vsb_buf_reserve(c->sorted, (size_t)npad * (16 + 8 + 8 + 4)));
    double *radm = (double *)c->sorted.p;
    cc_prepare_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(c->rad, n, npad, radm,
                                                                            c->d_key);
    c->launches += 1;
    if (n >= 2) {
        const int64_t row_blocks = (n - 1 + PB_BLOCK - 1) / PB_BLOCK;
        const int64_t mine = (row_blocks - part + parts - 1) / parts;
        if (mine > 0) {
            const int64_t tiles = (n + PB_TJ - 1) / PB_TJ;
            const int ct = chunk_tiles_for(tiles, mine, c->sm_count);
            dim3 grid((unsigned)mine, (unsigned)((tiles + ct - 1) / ct));
            cc_scan_kernel<<<grid, PB_BLOCK, 0, c->stream>>>((const double2 *)c->pos, c->rad, radm,
                                                             n, ct, c->d_key, part, parts);
            c->launches += 1;
        }
    }
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// (i, j) key in c->d_key -> (body_del, body_add) in c->scratch.p as two int64
int vsb_launch_check_collision_finish(vsb_ctx *c)
{
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    cc_finalize_kernel<<<1, 1, 0, c->stream>>>(c->d_key, c->mass, (int64_t *)c->scratch.p);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// Result (body_del, body_add) is written to c->scratch.p as two int64.
int vsb_launch_check_collision(vsb_ctx *c)
{
    VSB_TRY(vsb_launch_check_collision_part(c, 0, 1));
    return vsb_launch_check_collision_finish(c);
}
