// Development aid: the other per-warp iteration kernels of the BASELINE configurations in one small translation unit.
#include "../../nbldpc_b200/csrc/tems_warp.cuh"
#include "../../nbldpc_b200/csrc/ems_warp.cuh"
#include "../../nbldpc_b200/csrc/qspa_warp.cuh"
template __global__ void nbl::k_iterate_ems_warp<float, 64, 4, 2>(const nbl::IterParams<float>);
template __global__ void nbl::k_iterate_qspa_warp<float, 256, 4, 2>(const nbl::IterParams<float>);
