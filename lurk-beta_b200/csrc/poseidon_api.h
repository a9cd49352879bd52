// Declarations shared by the per-field Poseidon translation units and the C-ABI front end (poseidon.cu).
#pragma once
#include "common.cuh"
#include "poseidon_params.h"

namespace lurk {

struct PoseidonLayout {
    int rf, rp;
    int off_mds, off_pre, off_sw, off_sv, flat_len;   // in elements
    int block_elems;                                   // witness block size (arity + aux + 1)
};

// defined (explicitly instantiated) in poseidon_f{0..3}.cu
template <class F, bool WITNESS>
int launch_poseidon(int arity, const void *d_pre, size_t n, void *d_out, int in_fmt, int out_fmt, cudaStream_t s,
                    const uint64_t *d_offsets = nullptr);   // optional element offset of every witness block
template <class F>
int poseidon_instance_info(int arity, const PoseidonParams<F> **params, PoseidonLayout *layout);
template <class F>
int launch_bitdecomp(const void *d_values, size_t n, void *d_blocks, int blk, int fmt, cudaStream_t s, const uint64_t *d_offsets = nullptr);
int bitdecomp_block_host(const uint32_t mod[8]);

}  // namespace lurk
