// vsb_gravity.cu -- FP64 all-pairs gravity for sm_100a.
//
// Model: documentation/orbital-physics/01-complex-orbit-model.txt:4-12 with the
// pair formula of gravity_one (phys_editor.py:51-60):
//     a_i = gc * sum_{j != i} m_j (r_j - r_i) / |r_j - r_i|^3 ;  v += a ;  P += v
// 2-D, no softening.
//
// Kernel shape: grid (row blocks, source chunks).  A CTA owns BLOCK target rows
// (TPT targets per thread, in registers) and one chunk of sources, streamed
// through shared memory in TJ-body tiles by 1-D TMA bulk copies
// (cp.async.bulk -> UBLKCP) completing on mbarriers, double buffered.  Every
// lane reads the same source (LDS broadcast), so shared-memory traffic is
// negligible and the kernel is bound by the FP64 pipe: 13 DFMA-class issues per
// pair (2 sub, 2 for r^2, 6 for m_j*r^-3 from the MUFU.RSQ64H seed with one
// cubic correction, 1 for the mass, 2 accumulate FMAs).
//
// Determinism: the chunking depends on N only, every partial sum is formed in a
// fixed order and the partials are reduced in chunk order by the integrate
// kernel, so the result is bit-identical run to run and under any row sharding
// across GPUs (each rank computes only its own rows, same order).
#include <stdlib.h>

#include "vsb_common.cuh"

namespace {

constexpr int GR_BLOCK = 128;  // threads per CTA
constexpr int GR_TJ = 512;     // sources per shared-memory tile
constexpr int GR_STAGES = 2;
constexpr int GR_U = 4;        // independent accumulator chains per target

// m_j * |d|^-3 * d accumulated into (ax, ay).  13 FP64-pipe instructions.
// MASK: tiles that contain the CTA's own target rows can meet i == j (r2 == 0 ->
// seed = +inf).  There the seed is zeroed by an integer select on its high word (no
// FP64-pipe cost) and every later product is an exact 0 contribution; all other
// tiles skip the two integer instructions (the kernel is issue-bound: an FP64
// instruction holds the dispatch port for two cycles, everything else for one).
template <bool MASK>
__device__ __forceinline__ void pair_accum(double xi, double yi, double2 pj, double mj, double &ax,
                                           double &ay)
{
    double dx = pj.x - xi;
    double dy = pj.y - yi;
    const double dy2 = dy * dy;
    double r2 = fma(dx, dx, dy2);
    // MUFU.RSQ64H defines only the HIGH word of the seed.  Its low word is taken from the dead
    // temporary dy2 instead of being zeroed: one instruction less on the dispatch port per pair,
    // and a perturbation below 2^-20 that the correction absorbs (|e| < 2^-18.6, first dropped
    // term 2.19 e^3 < 3e-17).  r2 == 0 implies dy2 == 0, so the masked seed is still an exact 0.
    int yh = __double2hiint(rsqrt_seed(r2));
    if (MASK) yh = __double2hiint(r2) < 0x00100000 ? 0 : yh;
    const double y = __hiloint2double(yh, __double2loint(dy2));
    double t = y * y;
    double e = fma(-r2, t, 1.0);          // e = 1 - r2*y^2  (|e| < 2^-18.6)
    double ym = y * mj;
    double y3m = t * ym;                  // m_j * y^3
    double p = fma(1.875, e, 1.5);        // (1-e)^-3/2 = 1 + e*(3/2 + 15/8 e) + O(e^3)
    double q = e * p;
    double s = fma(y3m, q, y3m);          // m_j * r^-3
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
}

template <int TPT, bool MASK>
__device__ __forceinline__ void tile_accum(const double2 *__restrict__ sp,
                                           const double *__restrict__ sm, const double (&xi)[TPT],
                                           const double (&yi)[TPT], double (&ax)[TPT][GR_U],
                                           double (&ay)[TPT][GR_U])
{
#pragma unroll 2
    for (int j = 0; j < GR_TJ; j += GR_U) {
#pragma unroll
        for (int u = 0; u < GR_U; ++u) {
            double2 pj = sp[j + u];
            double mj = sm[j + u];
#pragma unroll
            for (int k = 0; k < TPT; ++k)
                pair_accum<MASK>(xi[k], yi[k], pj, mj, ax[k][u], ay[k][u]);
        }
    }
}

template <int TPT>
__global__ void __launch_bounds__(GR_BLOCK)
allpairs_accel_kernel(const double2 *__restrict__ pos, const double *__restrict__ mass,
                      double2 *__restrict__ part, int64_t row0, int64_t rows, int64_t chunk,
                      int64_t npad)
{
    __shared__ alignas(128) double2 spos[GR_STAGES][GR_TJ];
    __shared__ alignas(128) double smass[GR_STAGES][GR_TJ];
    __shared__ alignas(8) uint64_t full[GR_STAGES];

    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.y * chunk;
    int64_t j1 = j0 + chunk;
    if (j1 > npad) j1 = npad;
    const int ntiles = (int)((j1 - j0) / GR_TJ);
    constexpr uint32_t TILE_BYTES = GR_TJ * (sizeof(double2) + sizeof(double));

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GR_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GR_STAGES; ++s) {
            if (s < ntiles) {
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_bulk_g2s(spos[s], pos + j0 + (int64_t)s * GR_TJ, GR_TJ * sizeof(double2),
                             &full[s]);
                tma_bulk_g2s(smass[s], mass + j0 + (int64_t)s * GR_TJ, GR_TJ * sizeof(double),
                             &full[s]);
            }
        }
    }

    // global rows this CTA accumulates (for the diagonal-tile test)
    const int64_t cta_row0 = row0 + (int64_t)blockIdx.x * TPT * GR_BLOCK;
    const int64_t cta_row1 = cta_row0 + (int64_t)TPT * GR_BLOCK;
    // targets: row = row0 + (blockIdx.x*TPT + k)*BLOCK + tid
    double xi[TPT], yi[TPT];
    double ax[TPT][GR_U], ay[TPT][GR_U];
    int64_t lrow[TPT];
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        lrow[k] = ((int64_t)blockIdx.x * TPT + k) * GR_BLOCK + tid;
        int64_t g = row0 + (lrow[k] < rows ? lrow[k] : 0);
        double2 p = pos[g];
        xi[k] = p.x;
        yi[k] = p.y;
#pragma unroll
        for (int u = 0; u < GR_U; ++u) ax[k][u] = ay[k][u] = 0.0;
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % GR_STAGES;
        mbar_wait(&full[s], (uint32_t)(t / GR_STAGES) & 1u);
        // does this tile hold any of the CTA's own target rows (possible i == j)?
        const int64_t jt = j0 + (int64_t)t * GR_TJ;
        if (jt < cta_row1 && jt + GR_TJ > cta_row0)
            tile_accum<TPT, true>(spos[s], smass[s], xi, yi, ax, ay);
        else
            tile_accum<TPT, false>(spos[s], smass[s], xi, yi, ax, ay);
        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && t + GR_STAGES < ntiles) {
            const int64_t jn = j0 + (int64_t)(t + GR_STAGES) * GR_TJ;
            mbar_expect_tx(&full[s], TILE_BYTES);
            tma_bulk_g2s(spos[s], pos + jn, GR_TJ * sizeof(double2), &full[s]);
            tma_bulk_g2s(smass[s], mass + jn, GR_TJ * sizeof(double), &full[s]);
        }
    }

#pragma unroll
    for (int k = 0; k < TPT; ++k) {
        if (lrow[k] < rows) {
            double sx = (ax[k][0] + ax[k][1]) + (ax[k][2] + ax[k][3]);
            double sy = (ay[k][0] + ay[k][1]) + (ay[k][2] + ay[k][3]);
            part[(int64_t)blockIdx.y * rows + lrow[k]] = make_double2(sx, sy);
        }
    }
}

// Reduce the per-chunk partials in chunk order, scale by gc, store acc and
// (optionally) apply the reference integrator v += a ; P += v.
__global__ void __launch_bounds__(256)
allpairs_reduce_kernel(const double2 *__restrict__ part, int nchunks, int64_t row0, int64_t rows,
                       double gc, double2 *__restrict__ acc, double2 *__restrict__ vel,
                       double2 *__restrict__ pos, int integrate)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double sx = 0.0, sy = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        double2 p = part[(int64_t)c * rows + r];
        sx += p.x;
        sy += p.y;
    }
    sx *= gc;
    sy *= gc;
    const int64_t g = row0 + r;
    acc[g] = make_double2(sx, sy);
    if (integrate) {
        double2 v = vel[g];
        v.x += sx;
        v.y += sy;
        vel[g] = v;
        double2 p = pos[g];
        p.x += v.x;
        p.y += v.y;
        pos[g] = p;
    }
}

__global__ void integrate_only_kernel(int64_t row0, int64_t rows, const double2 *__restrict__ acc,
                                      double2 *__restrict__ vel, double2 *__restrict__ pos)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int64_t g = row0 + r;
    double2 a = acc[g], v = vel[g], p = pos[g];
    v.x += a.x;
    v.y += a.y;
    p.x += v.x;
    p.y += v.y;
    vel[g] = v;
    pos[g] = p;
}

// DFMA-chain microbenchmark: 16 independent chains per thread, `iters` x 64 DFMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456) out[0] = s;
}

}  // namespace

static inline int64_t pow2ceil(int64_t x)
{
    int64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Source chunk length (a function of N only) and chunk count.
int vsb_allpairs_chunks(int64_t n, int64_t *chunk_out)
{
    int64_t npad = vsb_round_up(n > 0 ? n : 1, VSB_PAD);
    int64_t chunk = pow2ceil(npad) / 32;
    if (chunk < GR_TJ) chunk = GR_TJ;
    chunk = vsb_round_up(chunk, GR_TJ);
    if (chunk_out) *chunk_out = chunk;
    return (int)((npad + chunk - 1) / chunk);
}

static int targets_per_thread(int64_t rows, int nchunks, int sm_count)
{
    // experiments: VSB_GRAVITY_TPT=1|2|4 forces a variant
    const char *env = getenv("VSB_GRAVITY_TPT");
    if (env && (env[0] == '1' || env[0] == '2' || env[0] == '4')) return env[0] - '0';
    // two targets per thread halve the LDS traffic; only worth it when the grid
    // still fills the machine several times over
    int64_t ctas2 = ((rows + 2 * GR_BLOCK - 1) / (2 * GR_BLOCK)) * nchunks;
    return ctas2 >= (int64_t)sm_count * 24 ? 2 : 1;
}

template <int TPT>
static void launch_accel(vsb_ctx *c, int64_t rows, int64_t chunk, int nchunks)
{
    dim3 grid((unsigned)((rows + (int64_t)TPT * GR_BLOCK - 1) / ((int64_t)TPT * GR_BLOCK)),
              (unsigned)nchunks);
    allpairs_accel_kernel<TPT><<<grid, GR_BLOCK, 0, c->stream>>>(
        (const double2 *)c->pos, c->mass, (double2 *)c->part.p, c->row0, rows, chunk, c->npad);
}

int vsb_launch_allpairs_accel(vsb_ctx *c, double gc)
{
    (void)gc;
    const int64_t rows = c->row1 - c->row0;
    if (rows <= 0) return VSB_OK;
    int64_t chunk;
    const int nchunks = vsb_allpairs_chunks(c->n, &chunk);
    VSB_TRY(vsb_buf_reserve(c->part, sizeof(double2) * (size_t)nchunks * (size_t)rows));
    const int tpt = targets_per_thread(rows, nchunks, c->sm_count);
    if (tpt == 4) launch_accel<4>(c, rows, chunk, nchunks);
    else if (tpt == 2) launch_accel<2>(c, rows, chunk, nchunks);
    else launch_accel<1>(c, rows, chunk, nchunks);
    c->launches++;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// reduce partials (+ integrate when `integrate`): second half of an all-pairs tick
int vsb_launch_allpairs_reduce(vsb_ctx *c, double gc, int integrate)
{
    const int64_t rows = c->row1 - c->row0;
    if (rows <= 0) return VSB_OK;
    const int nchunks = vsb_allpairs_chunks(c->n, nullptr);
    unsigned grid = (unsigned)((rows + 255) / 256);
    allpairs_reduce_kernel<<<grid, 256, 0, c->stream>>>(
        (const double2 *)c->part.p, nchunks, c->row0, rows, gc, (double2 *)c->acc,
        (double2 *)c->vel, (double2 *)c->pos, integrate);
    c->launches++;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_allpairs_integrate(vsb_ctx *c)
{
    const int64_t rows = c->row1 - c->row0;
    if (rows <= 0) return VSB_OK;
    unsigned grid = (unsigned)((rows + 255) / 256);
    integrate_only_kernel<<<grid, 256, 0, c->stream>>>(c->row0, rows, (const double2 *)c->acc,
                                                       (double2 *)c->vel, (double2 *)c->pos);
    c->launches++;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_dfma_peak(vsb_ctx *c, int iters, int blocks, float *ms_out)
{
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    dfma_peak_kernel<<<blocks, 256, 0, c->stream>>>((double *)c->scratch.p, 16, 0.999, 0.001);
    VSB_CUDA(cudaEventRecord(c->ev0, c->stream));
    dfma_peak_kernel<<<blocks, 256, 0, c->stream>>>((double *)c->scratch.p, iters, 0.999, 0.001);
    VSB_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    VSB_CUDA(cudaEventSynchronize(c->ev1));
    VSB_CUDA(cudaEventElapsedTime(ms_out, c->ev0, c->ev1));
    return VSB_OK;
}
