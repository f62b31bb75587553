// vsb_visible.cu -- SURVEY 8f row N3: culling + visible-subset curve extraction (sm_100a).
//   phys_shared.culling                      phys_shared.py:224-231
//   phys_body.Physics.culling (orbit part)   phys_body.py:201-218
//   phys_vessel.Physics.culling (orbit part) phys_vessel.py:258-273
// The predicates are evaluated on the device-resident positions, the surviving indices are
// written in ascending order by a stable three-pass compaction (per-CTA counts, scan of the
// counts, ordered scatter), and only the visible curves are gathered and copied to the host
// (game.py:1403-1407 draws nothing else) -- the (N,P,2) array never crosses PCIe.
// Compiled with -fmad=false (the comparisons must round like numpy).
#include "vsb_common.cuh"

namespace {

constexpr int VC_T = 256;

// phys_shared.culling: (x + r) + 2/zoom > sb[0][0] etc.; sb = {x0, y0, x1, y1} as the 2x2
// array [[x0, y0], [x1, y1]] the reference passes (note the swapped y bounds there).
__device__ __forceinline__ bool culling_one(double2 p, double r, const double4 sb, double two_over_zoom)
{
    const bool min_x = (p.x + r) + two_over_zoom > sb.x;   // screen_bounds[0,0]
    const bool min_y = (p.y + r) + two_over_zoom > sb.w;   // screen_bounds[1,1]
    const bool max_x = (p.x - r) + two_over_zoom < sb.z;   // screen_bounds[1,0]
    const bool max_y = (p.y - r) + two_over_zoom < sb.y;   // screen_bounds[0,1]
    return min_x && min_y && max_x && max_y;
}

__global__ void __launch_bounds__(VC_T)
culling_flags_kernel(const double2 *__restrict__ pos, const double *__restrict__ radius,
                     double r_uniform, int64_t n, double4 sb, double two_over_zoom,
                     unsigned char *__restrict__ flags)
{
    int64_t i = (int64_t)blockIdx.x * VC_T + threadIdx.x;
    if (i >= n) return;
    flags[i] = culling_one(pos[i], radius ? radius[i] : r_uniform, sb, two_over_zoom) ? 1 : 0;
}

// phys_body.py:212-215 / phys_vessel.py:267-270:
// ((ref_visible[ref] and ref_coi[ref] > 8/zoom) or ref == 0) and size*zoom < screen_x*15
__global__ void __launch_bounds__(VC_T)
orbit_flags_kernel(const int64_t *__restrict__ ref, const double *__restrict__ size, int64_t n,
                   const unsigned char *__restrict__ ref_visible,
                   const double *__restrict__ ref_coi, double eight_over_zoom, double zoom,
                   double limit, unsigned char *__restrict__ flags)
{
    int64_t i = (int64_t)blockIdx.x * VC_T + threadIdx.x;
    if (i >= n) return;
    const int64_t r = ref[i];
    bool v = ref_visible[r] && ref_coi[r] > eight_over_zoom;
    v = v || r == 0;
    v = v && size[i] * zoom < limit;
    flags[i] = v ? 1 : 0;
}

// stable compaction ---------------------------------------------------------------
__global__ void __launch_bounds__(VC_T)
compact_count_kernel(const unsigned char *__restrict__ flags, int64_t n, int *__restrict__ counts)
{
    int64_t i = (int64_t)blockIdx.x * VC_T + threadIdx.x;
    int f = (i < n && flags[i]) ? 1 : 0;
    int c = __syncthreads_count(f);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

// exclusive scan of up to 2^22 CTA counts by one CTA (each thread owns a contiguous run)
__global__ void __launch_bounds__(1024) compact_scan_kernel(int *counts, int nblocks, int64_t *total)
{
    __shared__ int sh[1024];
    const int per = (nblocks + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    int tot = 0;
    for (int k = 0; k < per; ++k)
        if (b0 + k < nblocks) tot += counts[b0 + k];
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - tot;
    for (int k = 0; k < per; ++k)
        if (b0 + k < nblocks) {
            int c = counts[b0 + k];
            counts[b0 + k] = run;
            run += c;
        }
    if (threadIdx.x == 1023) *total = sh[1023];
}

__global__ void __launch_bounds__(VC_T)
compact_write_kernel(const unsigned char *__restrict__ flags, int64_t n,
                     const int *__restrict__ offsets, int64_t *__restrict__ idx_out)
{
    __shared__ int warp_base[VC_T / 32];
    int64_t i = (int64_t)blockIdx.x * VC_T + threadIdx.x;
    const int f = (i < n && flags[i]) ? 1 : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_base[w] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < VC_T / 32; ++k) {
            int c = warp_base[k];
            warp_base[k] = run;
            run += c;
        }
    }
    __syncthreads();
    if (f) {
        const int rank = warp_base[w] + __popc(ballot & ((1u << lane) - 1u));
        idx_out[offsets[blockIdx.x] + rank] = i;
    }
}

// out[k] = curves[idx[k]] : warp per curve, 16-byte streaming copies
__global__ void __launch_bounds__(VC_T)
curves_gather_kernel(const double2 *__restrict__ curves, const int64_t *__restrict__ idx,
                     int64_t count, int64_t P, double2 *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * VC_T) >> 5;
    for (int64_t k = ((int64_t)blockIdx.x * VC_T + threadIdx.x) >> 5; k < count; k += warps) {
        const double2 *src = curves + idx[k] * P;
        double2 *dst = out + k * P;
        for (int64_t p = lane; p < P; p += 32) dst[p] = src[p];
    }
}

}  // namespace

int vsb_launch_culling_flags(vsb_ctx *c, const double *d_pos, const double *d_radius,
                             double r_uniform, int64_t n, const double *sb4, double zoom,
                             unsigned char *d_flags)
{
    if (n <= 0) return VSB_OK;
    double4 sb = make_double4(sb4[0], sb4[1], sb4[2], sb4[3]);
    culling_flags_kernel<<<(unsigned)((n + VC_T - 1) / VC_T), VC_T, 0, c->stream>>>(
        (const double2 *)d_pos, d_radius, r_uniform, n, sb, 2.0 / zoom, d_flags);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_orbit_flags(vsb_ctx *c, const int64_t *d_ref, const double *d_size, int64_t n,
                           const unsigned char *d_ref_visible, const double *d_ref_coi, double zoom,
                           double screen_x, unsigned char *d_flags)
{
    if (n <= 0) return VSB_OK;
    orbit_flags_kernel<<<(unsigned)((n + VC_T - 1) / VC_T), VC_T, 0, c->stream>>>(
        d_ref, d_size, n, d_ref_visible, d_ref_coi, 8.0 / zoom, zoom, screen_x * 15, d_flags);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

// flags (n bytes) -> ascending indices in d_idx (capacity n); count lands in d_total.
// d_counts must hold ceil(n/256) ints.
int vsb_launch_compact(vsb_ctx *c, const unsigned char *d_flags, int64_t n, int *d_counts,
                       int64_t *d_total, int64_t *d_idx)
{
    if (n <= 0) {
        VSB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), c->stream));
        return VSB_OK;
    }
    const int nblocks = (int)((n + VC_T - 1) / VC_T);
    compact_count_kernel<<<nblocks, VC_T, 0, c->stream>>>(d_flags, n, d_counts);
    compact_scan_kernel<<<1, 1024, 0, c->stream>>>(d_counts, nblocks, d_total);
    compact_write_kernel<<<nblocks, VC_T, 0, c->stream>>>(d_flags, n, d_counts, d_idx);
    c->launches += 3;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}

int vsb_launch_curves_gather(vsb_ctx *c, const double *d_curves, const int64_t *d_idx,
                             int64_t count, int64_t P, double *d_out)
{
    if (count <= 0) return VSB_OK;
    int64_t ctas = (count + 7) / 8;
    if (ctas > (int64_t)c->sm_count * 8) ctas = (int64_t)c->sm_count * 8;
    curves_gather_kernel<<<(unsigned)ctas, VC_T, 0, c->stream>>>(
        (const double2 *)d_curves, d_idx, count, P, (double2 *)d_out);
    c->launches += 1;
    VSB_CUDA(cudaGetLastError());
    return VSB_OK;
}
