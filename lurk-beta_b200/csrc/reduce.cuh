// Grid-wide sums of field elements for the reductions of the N4 kernels (sum-check round polynomials, inner products, polynomial
// evaluations): per-thread modular sums -> warp shuffles -> shared memory -> one partial per CTA -> the last CTA to finish adds the
// partials.  One launch, no second kernel, no atomics on field elements.  Device only.
#pragma once
#include "field.cuh"

namespace lurk {

// ------------------------------------------------------------------------------------------------ grid-wide sum of E field elements
template <class F>
__device__ __forceinline__ F shfl_down_fe(const F &x, int off) {
    F r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, x.v[i], off);
    return r;
}
// every thread of a 256-thread CTA calls this with its local sums; result[0..E) is written by the last CTA (Montgomery form)
template <class F, int E>
__device__ void grid_sum(F *acc, F *partial, unsigned *counter, F *result) {
    __shared__ F sh[8][E];
    __shared__ bool last;
#pragma unroll
    for (int e = 0; e < E; e++)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[e] = acc[e] + shfl_down_fe(acc[e], off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int e = 0; e < E; e++) sh[warp][e] = acc[e];
    __syncthreads();
    if (threadIdx.x < E) {
        F s = sh[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) s = s + sh[w][threadIdx.x];
        store_fe(&partial[(size_t)blockIdx.x * E + threadIdx.x], s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket = atomicAdd(counter, 1u);
        last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < E) {
        F s = F::zero();
        for (unsigned b = 0; b < gridDim.x; b++) {
            const uint4 *q = reinterpret_cast<const uint4 *>(&partial[(size_t)b * E + threadIdx.x]);
            uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
            F t;
            t.v[0] = lo.x; t.v[1] = lo.y; t.v[2] = lo.z; t.v[3] = lo.w; t.v[4] = hi.x; t.v[5] = hi.y; t.v[6] = hi.z; t.v[7] = hi.w;
            s = s + t;
        }
        store_fe(&result[threadIdx.x], s);
    }
    if (threadIdx.x == 0) *counter = 0;       // ready for the next launch on this stream
}

}  // namespace lurk
