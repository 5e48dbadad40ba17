// Development aid: a translation unit that instantiates only the headline iteration kernel, so that its SASS can be
// listed with line information in seconds (nvcc -cubin -lineinfo; nvdisasm -g).  Not part of the library.
#include "../../nbldpc_b200/csrc/tems_warp.cuh"
template __global__ void nbl::k_iterate_tems_warp<float, 256, 4, 2, 0>(const nbl::IterParams<float>);
