// Internal: the encoder handle (encoder.cu) and the launch hook the Monte-Carlo path of decoder.cu uses.
#pragma once
#include "internal.h"
#include <cuda_runtime.h>

struct nbl_encoder {
    const nbl_code *code = nullptr;
    int device = 0, q = 0, N = 0, M = 0, K = 0, n_swaps = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::vector<int> G;          // [M][K] parity r as a function of the message symbols
    std::vector<int> perm;       // out[i] = pre[perm[i]]: the recorded column swaps, undone
    uint16_t *d_logG = nullptr, *d_log = nullptr, *d_alog = nullptr;
    int *d_perm = nullptr;
    int *d_msg = nullptr, *d_out = nullptr; size_t msg_cap = 0, out_cap = 0;   // staging for host pointers
};

// device bytes in, device bytes out, on the caller's stream (no synchronisation)
int nbl_encoder_launch_u8(nbl_encoder *e, const uint8_t *msg_dev, int B, uint8_t *out_dev, cudaStream_t stream);
