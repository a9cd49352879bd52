// K5: Nova fold helpers on device-resident field vectors (Arecibo NIFS::prove / R1CSShape::commit_T /
// RelaxedR1CSWitness::fold -- SURVEY.md Appendix B steps 3 and 5; call sites src/proof/nova.rs:287-293), plus the
// element-wise format conversion and range check used by the host-buffer entry points.
//
// These are the HBM-bound kernels of the path: AXPY moves 96 B per element for one modular multiply, so they are
// written as grid-stride loops over 32-byte elements with 128-bit loads/stores, one element per thread per
// iteration (a warp touches 1 KiB contiguous per array), grid = a multiple of the SM count.
#include "common.cuh"
#include "reduce.cuh"

#include <initializer_list>

namespace lurk {

static inline int stream_grid(size_t n, int block, int per_sm) {
    size_t want = (n + block - 1) / block;
    size_t cap = (size_t)sm_count() * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

// in == out (in place) is allowed and used (dag.cu), hence no __restrict__: each thread reads element i before writing it
template <class F>
__global__ void __launch_bounds__(256) convert_kernel(const F *in, size_t n, int to_mont, F *out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        F x = load_fe<F>(in + i);
        store_fe(out + i, to_mont ? F::from_canonical(x) : x.to_canonical());
    }
}
template <class F>
int convert_dev(const void *d_in, size_t n, int to_fmt, void *d_out, cudaStream_t s) {
    if (n == 0) return LURK_OK;
    convert_kernel<F><<<stream_grid(n, 256, 8), 256, 0, s>>>((const F *)d_in, n, to_fmt == LURK_FMT_MONTGOMERY, (F *)d_out);
    LURK_CUDA_TRY(cudaGetLastError());
    return LURK_OK;
}

template <class F>
__global__ void __launch_bounds__(256) check_reduced_kernel(const F *__restrict__ in, size_t n, int *bad) {
    int local = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        F x = load_fe<F>(in + i);
        local += x.is_reduced() ? 0 : 1;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(bad, local);
}
template <class F>
int check_reduced_dev(const void *d_in, size_t n, cudaStream_t s, int *bad_host) {
    *bad_host = 0;
    if (n == 0) return LURK_OK;
    // stream-ordered scratch: cudaMalloc / cudaFree would synchronise the whole device inside every host-buffer call
    int *d_bad = nullptr;
    LURK_CUDA_TRY(cudaMallocAsync(&d_bad, sizeof(int), s));
    cudaError_t e = cudaMemsetAsync(d_bad, 0, sizeof(int), s);
    if (e == cudaSuccess) {
        check_reduced_kernel<F><<<stream_grid(n, 256, 8), 256, 0, s>>>((const F *)d_in, n, d_bad);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(bad_host, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaFreeAsync(d_bad, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    LURK_CUDA_TRY(e);
    return LURK_OK;
}

// asynchronous variant: adds the number of unreduced elements to *d_bad (device counter) on stream s
template <class F>
int check_reduced_accumulate_dev(const void *d_in, size_t n, cudaStream_t s, int *d_bad) {
    if (n == 0) return LURK_OK;
    check_reduced_kernel<F><<<stream_grid(n, 256, 8), 256, 0, s>>>((const F *)d_in, n, d_bad);
    LURK_CUDA_TRY(cudaGetLastError());
    return LURK_OK;
}

// out[i] = a[i] + r * b[i]; out may alias a (the fold updates W1 and E1 in place), so only b is __restrict__
template <class F>
__global__ void __launch_bounds__(256) axpy_kernel(const F *a, const F *__restrict__ b, F r, size_t n, F *out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        F x = load_fe<F>(a + i), y = load_fe<F>(b + i);
        store_fe(out + i, x + r * y);
    }
}

// CSR sparse matrix-vector product, one row per thread (R1CS rows carry a handful of non-zeros)
// Rows longer than SPMV_LONG non-zeros are not walked by one thread (a transposed R1CS matrix has a few such rows: the columns of u and
// of the public IO collect one entry per linear constraint -- 10^4..10^5 entries, 11 ms of one thread's time each): the thread that meets
// one appends it to a list and a second kernel gives every listed row a whole CTA.
constexpr uint64_t SPMV_LONG = 1024;
template <class F>
__global__ void __launch_bounds__(256) spmv_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col,
                                                   const F *__restrict__ val, size_t rows, const F *__restrict__ z, F *__restrict__ y,
                                                   unsigned *long_count, uint32_t *long_rows, unsigned long_cap) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t k0 = row_ptr[i], k1 = row_ptr[i + 1];
        if (long_rows && k1 - k0 > SPMV_LONG) {
            const unsigned slot = atomicAdd(long_count, 1u);
            if (slot < long_cap) { long_rows[slot] = (uint32_t)i; continue; }      // list full: fall through and walk the row here
        }
        F acc = F::zero();
        for (uint64_t k = k0; k < k1; k++) acc = acc + load_fe<F>(val + k) * load_fe<F>(z + col[k]);
        store_fe(y + i, acc);
    }
}
template <class F>
__global__ void __launch_bounds__(256) spmv_long_kernel(const uint64_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, const F *__restrict__ val,
                                                        const F *__restrict__ z, F *__restrict__ y, const unsigned *__restrict__ long_count,
                                                        const uint32_t *__restrict__ long_rows, unsigned long_cap) {
    __shared__ F sh[8];
    const unsigned n = *long_count < long_cap ? *long_count : long_cap;
    for (unsigned r = blockIdx.x; r < n; r += gridDim.x) {
        const size_t i = long_rows[r];
        const uint64_t k0 = row_ptr[i], k1 = row_ptr[i + 1];
        F acc = F::zero();
        for (uint64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) acc = acc + load_fe<F>(val + k) * load_fe<F>(z + col[k]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = acc + shfl_down_fe(acc, off);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            F t = sh[0];
            for (int w = 1; w < 8; w++) t = t + sh[w];
            store_fe(y + i, t);
        }
        __syncthreads();
    }
}

// T = az1*bz2 + az2*bz1 - u1*cz2 - u2*cz1 : four products, one lazy reduction for the first two
template <class F>
__global__ void __launch_bounds__(256) cross_term_kernel(const F *__restrict__ az1, const F *__restrict__ bz1, const F *__restrict__ cz1,
                                                         const F *__restrict__ az2, const F *__restrict__ bz2, const F *__restrict__ cz2,
                                                         F u1, F u2, size_t n, F *__restrict__ t) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        WideAcc<typename F::Params> acc;
        acc.clear();
        acc.mul_acc(load_fe<F>(az1 + i), load_fe<F>(bz2 + i));
        acc.mul_acc(load_fe<F>(az2 + i), load_fe<F>(bz1 + i));
        F pos = acc.reduce();
        WideAcc<typename F::Params> neg;
        neg.clear();
        neg.mul_acc(u1, load_fe<F>(cz2 + i));
        neg.mul_acc(u2, load_fe<F>(cz1 + i));
        store_fe(t + i, pos - neg.reduce());
    }
}

template <class F> static F fe_from_bytes(const uint8_t b[32]) { F x; memcpy(x.v, b, 32); return x; }

#define LURK_FOLD_INSTANTIATE(F)                                                         \
    template int convert_dev<F>(const void *, size_t, int, void *, cudaStream_t);        \
    template int check_reduced_dev<F>(const void *, size_t, cudaStream_t, int *);         \
    template int check_reduced_accumulate_dev<F>(const void *, size_t, cudaStream_t, int *);
LURK_FOLD_INSTANTIATE(Fe<Bn254Fr>)
LURK_FOLD_INSTANTIATE(Fe<Bn254Fq>)
LURK_FOLD_INSTANTIATE(Fe<PallasFq>)
LURK_FOLD_INSTANTIATE(Fe<PallasFp>)

}  // namespace lurk

using namespace lurk;

extern "C" {

static int need(size_t n, std::initializer_list<const void *> ptrs) {
    if (n) for (const void *p : ptrs) if (!p) { set_error("null buffer"); return LURK_ERR_ARG; }
    return LURK_OK;
}

int lurk_convert_dev(int field_id, const void *d_in, size_t n, int to_fmt, void *d_out, void *stream) {
    if (to_fmt != LURK_FMT_CANONICAL && to_fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", to_fmt); return LURK_ERR_ARG; }
    LURK_TRY(need(n, {d_in, d_out}));
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return convert_dev<F>(d_in, n, to_fmt, d_out, (cudaStream_t)stream);
    });
}

int lurk_axpy_dev(int field_id, const void *d_a, const void *d_b, const uint8_t r_mont[32], size_t n, void *d_out, void *stream) {
    LURK_TRY(need(n, {d_a, d_b, r_mont, d_out}));
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        F r = fe_from_bytes<F>(r_mont);
        if (!r.is_reduced()) { set_error("scalar r is not reduced"); return LURK_ERR_RANGE; }
        axpy_kernel<F><<<stream_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const F *)d_a, (const F *)d_b, r, n, (F *)d_out);
        LURK_CUDA_TRY(cudaGetLastError());
        return LURK_OK;
    });
}

int lurk_spmv_csr_dev(int field_id, const void *d_row_ptr, const void *d_col, const void *d_val, size_t rows, const void *d_z,
                      void *d_y, void *stream) {
    LURK_TRY(need(rows, {d_row_ptr, d_z, d_y}));   // col / val may be null for a matrix without non-zeros
    LURK_TRY(require_gpu());
    if (rows == 0) return LURK_OK;
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        cudaStream_t s = (cudaStream_t)stream;
        // stream-ordered scratch for the list of long rows (counter + up to LONG_CAP row indices)
        constexpr unsigned LONG_CAP = 4096;
        unsigned *d_long = nullptr;
        LURK_CUDA_TRY(cudaMallocAsync(&d_long, sizeof(unsigned) * (LONG_CAP + 1), s));
        cudaError_t e = cudaMemsetAsync(d_long, 0, sizeof(unsigned), s);
        if (e == cudaSuccess) {
            spmv_kernel<F><<<stream_grid(rows, 256, 8), 256, 0, s>>>((const uint64_t *)d_row_ptr, (const uint32_t *)d_col, (const F *)d_val, rows,
                                                                     (const F *)d_z, (F *)d_y, d_long, d_long + 1, LONG_CAP);
            spmv_long_kernel<F><<<64, 256, 0, s>>>((const uint64_t *)d_row_ptr, (const uint32_t *)d_col, (const F *)d_val, (const F *)d_z, (F *)d_y, d_long,
                                                    d_long + 1, LONG_CAP);
            e = cudaGetLastError();
        }
        cudaFreeAsync(d_long, s);
        LURK_CUDA_TRY(e);
        LURK_CUDA_TRY(cudaGetLastError());
        return LURK_OK;
    });
}

int lurk_cross_term_dev(int field_id, const void *d_az1, const void *d_bz1, const void *d_cz1, const void *d_az2, const void *d_bz2,
                        const void *d_cz2, const uint8_t u1_mont[32], const uint8_t u2_mont[32], size_t n, void *d_t, void *stream) {
    LURK_TRY(need(n, {d_az1, d_bz1, d_cz1, d_az2, d_bz2, d_cz2, u1_mont, u2_mont, d_t}));
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        F u1 = fe_from_bytes<F>(u1_mont), u2 = fe_from_bytes<F>(u2_mont);
        if (!u1.is_reduced() || !u2.is_reduced()) { set_error("scalar u is not reduced"); return LURK_ERR_RANGE; }
        cross_term_kernel<F><<<stream_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
            (const F *)d_az1, (const F *)d_bz1, (const F *)d_cz1, (const F *)d_az2, (const F *)d_bz2, (const F *)d_cz2, u1, u2, n, (F *)d_t);
        LURK_CUDA_TRY(cudaGetLastError());
        return LURK_OK;
    });
}

}  // extern "C"
