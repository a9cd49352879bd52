// The MSM kernels and pipeline for one curve per object file (-DLURK_C=<curve struct>): parallel build.
#include "msm_impl.cuh"
namespace lurk { LURK_MSM_INSTANTIATE(LURK_C) }
