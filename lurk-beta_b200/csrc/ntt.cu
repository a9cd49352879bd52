// K6: radix-2 number-theoretic transform over the 2-adic subgroup of the proving fields.
//
// north_star asks for an NTT although the reference has no call site for one (lurk-beta compresses
// with Spartan over IPA / HyperKZG, sum-check based); parity is therefore pinned only against the oracle's O(n^2)
// DFT.  HBM-bound kernel family (64 B of traffic per element per pass):
//   * twiddles omega^j, j < n/2, are generated once per (field, log_n, direction) and cached on the device;
//   * the first min(log_n, 10) butterfly stages run inside shared memory on 1024-element tiles after a
//     bit-reversal gather (one global read + one global write for 10 stages);
//   * the remaining stages are strided global passes, two stages (radix-4) per pass where possible; the last pass writes
//     into the caller's buffer (no trailing copy).
// For 256-bit fields the butterflies' multiplier work (12 x 2^24 products = 2.9 ms on the IMAD pipe at 2^24) exceeds the HBM
// time of the passes, so the transform is multiply bound like the rest of the path.  A variant that ran
// stages 11..20 on shared-memory column tiles (4 global passes instead of 9) was measured slower (4.7 ms) in round 1 and dropped.
// Natural order in, natural order out, Montgomery form, in place.
#include "common.cuh"

#include <map>

namespace lurk {

template <class F>
__global__ void __launch_bounds__(256) ntt_twiddle_kernel(F omega, size_t half, F *__restrict__ tw) {
    // tw[j] = omega^j by square-and-multiply per element (one-off, cached)
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += (size_t)gridDim.x * blockDim.x) {
        F acc = F::one(), base = omega;
        for (size_t e = j; e; e >>= 1) { if (e & 1) acc = acc * base; base = base.sqr(); }
        store_fe(tw + j, acc);
    }
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return __brev(x) >> (32 - bits); }

// stages 1..S (S = tile_log) on tiles of 2^S consecutive outputs; input gathered in bit-reversed order.
// out-of-place: src -> dst
template <class F>
__global__ void __launch_bounds__(512) ntt_tile_kernel(const F *__restrict__ src, F *__restrict__ dst, const F *__restrict__ tw, int log_n, int tile_log) {
    extern __shared__ __align__(16) unsigned char ntt_smem[];
    F *sm = reinterpret_cast<F *>(ntt_smem);
    const uint32_t tile = 1u << tile_log;
    const size_t base = (size_t)blockIdx.x * tile;
    for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) {
        size_t g = base + i;
        sm[i] = load_fe<F>(src + bitrev((uint32_t)g, log_n));
    }
    __syncthreads();
    for (int s = 1; s <= tile_log; s++) {
        const uint32_t half = 1u << (s - 1);
        for (uint32_t k = threadIdx.x; k < tile / 2; k += blockDim.x) {
            const uint32_t j = k & (half - 1), i0 = ((k >> (s - 1)) << s) + j, i1 = i0 + half;
            const F w = load_fe<F>(tw + ((size_t)j << (log_n - s)));
            const F u = sm[i0], t = w * sm[i1];
            sm[i0] = u + t;
            sm[i1] = u - t;
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) store_fe(dst + base + i, sm[i]);
}

// one global radix-2 stage s; src == dst (in place) or out of place (every thread reads and writes the same two positions)
template <class F>
__global__ void __launch_bounds__(256) ntt_stage_kernel(const F *src, F *dst, const F *__restrict__ tw, int log_n, int s) {
    const size_t n2 = (size_t)1 << (log_n - 1);
    const size_t half = (size_t)1 << (s - 1);
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += (size_t)gridDim.x * blockDim.x) {
        const size_t j = k & (half - 1), i0 = ((k >> (s - 1)) << s) + j, i1 = i0 + half;
        const F w = load_fe<F>(tw + (j << (log_n - s)));
        const F u = load_fe<F>(src + i0), t = w * load_fe<F>(src + i1);
        store_fe(dst + i0, u + t);
        store_fe(dst + i1, u - t);
    }
}

// two global stages s, s+1 fused (radix-4 butterfly): 4 loads + 4 stores for two stages
template <class F>
__global__ void __launch_bounds__(256) ntt_stage2_kernel(const F *src, F *dst, const F *__restrict__ tw, int log_n, int s) {
    const size_t n4 = (size_t)1 << (log_n - 2);
    const size_t half = (size_t)1 << (s - 1);   // distance in stage s; stage s+1 distance = 2*half
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
        const size_t j = k & (half - 1);
        const size_t i0 = ((k >> (s - 1)) << (s + 1)) + j, i1 = i0 + half, i2 = i0 + 2 * half, i3 = i2 + half;
        const F w1 = load_fe<F>(tw + (j << (log_n - s)));                 // stage s twiddle (same for both pairs)
        const F w2a = load_fe<F>(tw + (j << (log_n - s - 1)));            // stage s+1 twiddle for element j
        const F w2b = load_fe<F>(tw + ((j + half) << (log_n - s - 1)));   // ... and for element j + half
        F x0 = load_fe<F>(src + i0), x1 = load_fe<F>(src + i1), x2 = load_fe<F>(src + i2), x3 = load_fe<F>(src + i3);
        F t = w1 * x1; x1 = x0 - t; x0 = x0 + t;
        t = w1 * x3; x3 = x2 - t; x2 = x2 + t;
        t = w2a * x2; x2 = x0 - t; x0 = x0 + t;
        t = w2b * x3; x3 = x1 - t; x1 = x1 + t;
        store_fe(dst + i0, x0); store_fe(dst + i1, x1); store_fe(dst + i2, x2); store_fe(dst + i3, x3);
    }
}

template <class F>
__global__ void __launch_bounds__(256) ntt_scale_kernel(const F *src, F *dst, size_t n, F k) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        store_fe(dst + i, load_fe<F>(src + i) * k);
}

template <class F>
struct NttCache {
    std::mutex mu;
    std::map<long, F *> tw;        // key: device * 1024 + log_n * 2 + inverse
};

template <class F>
static int ntt_run(void *d_data, int log_n, int inverse, cudaStream_t s) {
    using P = typename F::Params;
    if (log_n < 1 || log_n > P::TWO_ADICITY || log_n > 30) {
        set_error("log_n = %d outside [1, %d] for this field", log_n, P::TWO_ADICITY < 30 ? P::TWO_ADICITY : 30);
        return LURK_ERR_ARG;
    }
    static NttCache<F> cache;
    const size_t n = (size_t)1 << log_n;
    int dev = 0;
    LURK_CUDA_TRY(cudaGetDevice(&dev));
    F *tw = nullptr, *tmp = nullptr;
    {
        std::lock_guard<std::mutex> g(cache.mu);
        long key = (long)dev * 1024 + log_n * 2 + (inverse ? 1 : 0);
        auto it = cache.tw.find(key);
        if (it == cache.tw.end()) {
            // omega = ROOT^(2^(s - log_n)), a primitive 2^log_n-th root of unity; inverse uses omega^-1
            F omega;
            for (int i = 0; i < 8; i++) omega.v[i] = P::ROOT(i);
            for (int i = 0; i < P::TWO_ADICITY - log_n; i++) omega = omega.sqr();
            if (inverse) omega = omega.inv();
            F *d = nullptr;
            LURK_CUDA_TRY(cudaMalloc(&d, (n / 2 ? n / 2 : 1) * sizeof(F)));
            ntt_twiddle_kernel<F><<<sm_count() * 4, 256, 0, s>>>(omega, n / 2, d);
            cudaError_t e = cudaGetLastError();
            // the table is shared by every later call on any stream: publish it only once it is complete
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cudaFree(d); set_error("twiddle generation failed: %s", cudaGetErrorString(e)); return LURK_ERR_CUDA; }
            it = cache.tw.emplace(key, d).first;
        }
        tw = it->second;
    }
    // ping-pong buffer from the stream-ordered allocator: concurrent transforms on different streams never share it
    LURK_CUDA_TRY(cudaMallocAsync((void **)&tmp, n * sizeof(F), s));
    F *a = (F *)d_data;
    const int tile_log = log_n < 10 ? log_n : 10;
    const size_t smem = ((size_t)1 << tile_log) * sizeof(F);
    ntt_tile_kernel<F><<<(unsigned)(n >> tile_log), 512, smem, s>>>(a, tmp, tw, log_n, tile_log);
    const int grid = sm_count() * 8;
    // remaining stages on the ping-pong buffer, two per pass (radix-4; a radix-8 pass -- three stages, 120 registers -- was measured
    // slower: 4.75 ms instead of 4.14 ms at 2^24); the LAST pass of the whole transform writes back into the caller's buffer, so
    // there is no trailing copy
    int st = tile_log + 1;
    const int remaining = log_n - tile_log;
    int passes = (remaining + 1) / 2 + (inverse ? 1 : 0);
    auto dst_of = [&](void) { return --passes == 0 ? a : tmp; };
    while (st <= log_n) {
        F *dst = dst_of();
        if (st + 1 <= log_n) { ntt_stage2_kernel<F><<<grid, 256, 0, s>>>(tmp, dst, tw, log_n, st); st += 2; }
        else { ntt_stage_kernel<F><<<grid, 256, 0, s>>>(tmp, dst, tw, log_n, st); st += 1; }
    }
    cudaError_t e = cudaGetLastError();
    if (inverse) {
        F ninv = F::from_u64((uint64_t)n).inv();
        ntt_scale_kernel<F><<<grid, 256, 0, s>>>(tmp, a, n, ninv);
        if (e == cudaSuccess) e = cudaGetLastError();
    } else if (remaining == 0) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(a, tmp, n * sizeof(F), cudaMemcpyDeviceToDevice, s);   // log_n <= 10: one tile pass only
    }
    cudaFreeAsync(tmp, s);
    LURK_CUDA_TRY(e);
    return LURK_OK;
}

}  // namespace lurk

using namespace lurk;

extern "C" int lurk_ntt_dev(int field_id, void *d_data, int log_n, int inverse, void *stream) {
    LURK_TRY(require_gpu());
    if (!d_data) { set_error("null data"); return LURK_ERR_ARG; }
    return dispatch_field(field_id, [&](auto f) { return ntt_run<decltype(f)>(d_data, log_n, inverse, (cudaStream_t)stream); });
}
