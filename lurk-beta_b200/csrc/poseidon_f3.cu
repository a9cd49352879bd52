// Poseidon kernels instantiated for Fe<PallasFp> (one translation unit per field keeps the build parallel).
#include "poseidon_kernel.cuh"
namespace lurk { LURK_POSEIDON_INSTANTIATE(Fe<PallasFp>) }
