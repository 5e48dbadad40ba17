// Golden-vector dumper (test infrastructure, ours).  Links against the UNMODIFIED reference
// objects and drives the reference's own public stages -- CComm::GenerateMessage ... Demodulate
// (Comm.h:30-96) and CNBLDPC::Decoding (NBLDPC.h:71) -- then writes, per frame, the received
// points, the channel LLRs, the decoder's return code, iteration count, hard decisions and L_post.
//
// usage (cwd must hold the code/constellation files and ./SRC/ tables):
//     ref_dump <profile.txt> <frames> <out.bin> [skip_frames]
//
// Iteration count: `iter_number` is a local of the reference decoders (NBLDPC.cpp:671,803,971), so it
// is recovered without touching the source as the smallest maxIter for which Decoding() returns 1
// (the decoder is deterministic and re-initialises all state per call, NBLDPC.cpp:934-969).
#include "Simulation.h"
#include "Comm.h"
#include <cstdio>
#include <cstdint>
#include <vector>

static int decode(CComm &C) {
    return C.NBLDPC.Decoding(C.RX_LLR_SYM, C.RX_DECODE_SYM, C.ReliableSeri_Symbol, C.ReliableSeri_BIT);
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_dump profile frames out.bin [skip]\n"); return 2; }
    int frames = atoi(argv[2]);
    int skip = argc > 4 ? atoi(argv[4]) : 0;
    CSimulation sim;
    sim.Initial(argv[1]);
    CComm C;
    C.Initial(sim, 0);
    sim.NextSNR();
    C.SetEbN0(sim, 0);

    const int N = C.CODE_SYM_LEN, K = C.MSG_SYM_LEN, q = C.GFq, L = C.MOD_SYM_LEN;
    const int maxIter = C.NBLDPC.maxIter;
    FILE *f = fopen(argv[3], "wb");
    int32_t hdr[8] = {0x4E424C44, frames, N, q, maxIter, L, K, sim.decodeMethod};
    fwrite(hdr, sizeof(hdr), 1, f);
    double sg = C.sigma_n;
    fwrite(&sg, sizeof sg, 1, f);

    for (int fr = 0; fr < skip + frames; ++fr) {
        C.GenerateMessage(); C.Encode(); C.Puncture(); C.Modulate(); C.Channel_AWGN(); C.Demodulate();
        if (fr < skip) continue;
        int ret = decode(C);
        int iters = maxIter;
        if (ret == 1) {                       // binary search for the first converging maxIter
            int lo = 1, hi = maxIter;
            while (lo < hi) {
                int mid = (lo + hi) / 2;
                C.NBLDPC.maxIter = mid;
                if (decode(C) == 1) hi = mid; else lo = mid + 1;
            }
            iters = lo;
            C.NBLDPC.maxIter = maxIter;
            ret = decode(C);                  // leave outputs / L_post as the full run leaves them
        }
        for (int s = 0; s < L; ++s) { fwrite(&C.RX_MOD_SYM[s].Real, 8, 1, f); fwrite(&C.RX_MOD_SYM[s].Image, 8, 1, f); }
        for (int s = 0; s < N; ++s) { int32_t v = C.TX_CODE_SYM[s]; fwrite(&v, 4, 1, f); }
        for (int s = 0; s < K; ++s) { int32_t v = C.TX_MSG_SYM[s]; fwrite(&v, 4, 1, f); }
        int32_t ri[2] = {ret, iters};
        fwrite(ri, sizeof ri, 1, f);
        for (int s = 0; s < N; ++s) { int32_t v = C.RX_DECODE_SYM[s]; fwrite(&v, 4, 1, f); }
        for (int s = 0; s < N; ++s) fwrite(C.RX_LLR_SYM[s], 8, q - 1, f);
        for (int s = 0; s < N; ++s) fwrite(C.NBLDPC.L_post[s], 8, q - 1, f);
        fprintf(stderr, "frame %d ret %d iters %d\n", fr, ret, iters);
    }
    fclose(f);
    return 0;
}
