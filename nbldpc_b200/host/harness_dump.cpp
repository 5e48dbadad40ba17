// Test driver of the harness: writes the same per-frame record as oracle/ref_dump.cpp (received points, transmitted
// symbols, decoder outputs) so that tests can compare this harness with the golden vectors dumped from the reference.
//   harness_dump <profile> <frames> <out.bin> [decode]     decode = 0: host stages only (runs without a GPU)
#include "link_chain.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: harness_dump profile frames out.bin [decode]\n"); return 2; }
    const int frames = atoi(argv[2]);
    const bool decode = argc > 4 ? atoi(argv[4]) != 0 : true;
    CSimulation sim;
    sim.Initial(argv[1]);
    CComm C;
    C.Initial(sim, 0);
    sim.NextSNR();
    C.SetEbN0(sim, 0);
    const int N = C.CODE_SYM_LEN, K = C.MSG_SYM_LEN, q = C.GFq, L = C.MOD_SYM_LEN;
    FILE *f = fopen(argv[3], "wb");
    int32_t hdr[8] = {0x4E424C44, frames, N, q, sim.maxIter, L, K, sim.decodeMethod};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(&C.sigma_n, 8, 1, f);
    for (int fr = 0; fr < frames; ++fr) {
        C.GenerateMessage(); C.Encode(); C.Puncture(); C.Modulate(); C.Channel_AWGN(); C.Demodulate();
        int32_t ri[2] = {0, 0};
        if (decode) { C.Decode(); ri[0] = C.DecodeCorrect; ri[1] = C.iterations; }
        for (int s = 0; s < L; ++s) { fwrite(&C.RX_MOD_SYM[s].Real, 8, 1, f); fwrite(&C.RX_MOD_SYM[s].Image, 8, 1, f); }
        for (int s = 0; s < N; ++s) { int32_t v = C.TX_CODE_SYM[s]; fwrite(&v, 4, 1, f); }
        for (int s = 0; s < K; ++s) { int32_t v = C.TX_MSG_SYM[s]; fwrite(&v, 4, 1, f); }
        fwrite(ri, sizeof ri, 1, f);
        for (int s = 0; s < N; ++s) { int32_t v = decode ? C.RX_DECODE_SYM[s] : 0; fwrite(&v, 4, 1, f); }
        for (int s = 0; s < N; ++s) fwrite(C.RX_LLR_SYM[s], 8, q - 1, f);
        for (int s = 0; s < N; ++s) fwrite(C.NBLDPC.L_post[s], 8, q - 1, f);
    }
    fclose(f);
    return 0;
}
