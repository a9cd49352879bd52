// Poseidon kernels instantiated for Fe<Bn254Fr> (one translation unit per field keeps the build parallel).
#include "poseidon_kernel.cuh"
namespace lurk { LURK_POSEIDON_INSTANTIATE(Fe<Bn254Fr>) }
