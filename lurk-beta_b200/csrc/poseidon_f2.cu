// Poseidon kernels instantiated for Fe<PallasFq> (one translation unit per field keeps the build parallel).
#include "poseidon_kernel.cuh"
namespace lurk { LURK_POSEIDON_INSTANTIATE(Fe<PallasFq>) }
