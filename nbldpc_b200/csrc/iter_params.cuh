// Parameters of one iteration launch, shared by every iteration kernel (decoder.cu and the *_warp / *_half headers).
#pragma once
#include "dev_common.cuh"

namespace nbl {

template <typename T> struct IterParams {
    DevGraph g;
    int B, iter, first, early_stop, method, p0, p1;
    int bitexact;                       // T-EMS tie mode BITEXACT: the sequential replay of the reference's search
    T factor, offset, w_old, w_new;
    const T *lch;
    T *v2c;
    const T *c2v_in;
    T *c2v_out;
    uint8_t *vdec, *dec;
    int *flag;
    const uint8_t *done;
    const uint8_t *pfwd_sw, *pinv_sw;   // warp fast path: the permutation tables with values as split-layout indices
};

}  // namespace nbl
