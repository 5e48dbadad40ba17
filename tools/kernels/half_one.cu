// Development aid: the GF(16) four-lanes-per-row kernels in one small translation unit.
#include "../../nbldpc_b200/csrc/tems_warp.cuh"
#include "../../nbldpc_b200/csrc/qspa_half.cuh"
template __global__ void nbl::k_iterate_qspa_half<float>(const nbl::IterParams<float>, const int *);
