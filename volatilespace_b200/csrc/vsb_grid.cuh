// vsb_grid.cuh -- pieces shared by the two sort-based candidate passes (vsb_broadphase.cu for
// check_collision, vsb_parents_grid.cu for find_parents): the cell hash of the uniform grid and
// the three-kernel exclusive scan of the bucket histogram (counting sort by cell).
#pragma once
#include "vsb_common.cuh"

namespace {

__device__ __forceinline__ unsigned cell_bucket(long long cx, long long cy, unsigned mask)
{
    unsigned h = (unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    return h & mask;
}

// exclusive scan of counts[M] in three coalesced passes (1024-entry chunks)
__global__ void __launch_bounds__(256)
bp_scan_reduce_kernel(const unsigned *__restrict__ counts, unsigned *__restrict__ chunk_sum)
{
    __shared__ unsigned sh[256];
    const uint4 v = ((const uint4 *)counts)[(size_t)blockIdx.x * 256 + threadIdx.x];
    sh[threadIdx.x] = v.x + v.y + v.z + v.w;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(1024)
bp_scan_chunks_kernel(unsigned *chunk_sum, int nchunks)
{
    // single CTA, nchunks <= 4096: each thread owns up to 4 consecutive chunks
    __shared__ unsigned sh[1024];
    unsigned v[4], tot = 0;
    for (int k = 0; k < 4; ++k) {
        int c = threadIdx.x * 4 + k;
        v[k] = c < nchunks ? chunk_sum[c] : 0u;
        tot += v[k];
    }
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    unsigned run = sh[threadIdx.x] - tot;  // exclusive prefix of this thread's group
    for (int k = 0; k < 4; ++k) {
        int c = threadIdx.x * 4 + k;
        if (c < nchunks) chunk_sum[c] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(256)
bp_scan_apply_kernel(const unsigned *__restrict__ counts, const unsigned *__restrict__ chunk_off,
                     unsigned *__restrict__ starts, unsigned *__restrict__ cursor, unsigned M)
{
    __shared__ unsigned sh[256];
    const size_t q = (size_t)blockIdx.x * 256 + threadIdx.x;
    const uint4 v = ((const uint4 *)counts)[q];
    const unsigned tot = v.x + v.y + v.z + v.w;
    sh[threadIdx.x] = tot;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        unsigned add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    unsigned base = chunk_off[blockIdx.x] + sh[threadIdx.x] - tot;
    uint4 s;
    s.x = base;
    s.y = s.x + v.x;
    s.z = s.y + v.y;
    s.w = s.z + v.z;
    ((uint4 *)starts)[q] = s;
    ((uint4 *)cursor)[q] = s;
    if (q * 4 + 4 == M) starts[M] = s.w + v.w;  // total
}

}  // namespace
