#include "link_chain.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

static void reset_source(CComm &c, int seed, int order) {      // Comm.cpp:58-74 / 160-174
    c.Rand.IX = c.Rand.IY = c.Rand.IZ = seed;
    if (order) { c.Rand.IX += order; c.Rand.IY += order; c.Rand.IZ += order; }
    static const int init[11] = {1, 0, 1, 0, 0, 0, 1, 1, 0, 0, 1};
    for (int i = 0; i < 11; ++i) c.regPN[i] = init[i];
    for (int i = 0; i < order; ++i) c.GenPN();
}

bool CComm::Initial(CSimulation &sim, int parallel_order) {
    GFq = sim.GFq; parallel_num = sim.parallel;
    Bit_Len_PerSYM = 0; while ((1 << Bit_Len_PerSYM) < GFq) ++Bit_Len_PerSYM;
    NBLDPC.Initial(sim);
    reset_source(*this, sim.randomseed, parallel_order);
    randomMsg = sim.randomMsg;
    crcLen = sim.crcLen; crcLen_correct = sim.crc_correctLen;
    MSG_SYM_LEN = NBLDPC.CodeLen - NBLDPC.ChkLen; MSG_BIT_LEN = MSG_SYM_LEN * Bit_Len_PerSYM;
    CODE_SYM_LEN = NBLDPC.CodeLen; CODE_BIT_LEN = CODE_SYM_LEN * Bit_Len_PerSYM;
    PUN_SYM_LEN = NBLDPC.PunctureLen; PUN_BIT_LEN = PUN_SYM_LEN * Bit_Len_PerSYM;
    CodeRate = double(MSG_SYM_LEN) / double(CODE_SYM_LEN - PUN_SYM_LEN);       // CRC overhead not included (Comm.cpp:98)
    modOrder = sim.nQAM;
    MOD_BIT_PER_SYM = 0; while ((1 << MOD_BIT_PER_SYM) < modOrder) ++MOD_BIT_PER_SYM;
    MOD_SYM_LEN = (CODE_SYM_LEN - PUN_SYM_LEN) * Bit_Len_PerSYM / MOD_BIT_PER_SYM;
    MOD_BIT_LEN = MOD_SYM_LEN * MOD_BIT_PER_SYM;
    if (modOrder != 2 && modOrder != GFq) {
        std::cout << "This module ( code order ~= modulation order ) haven't been developed!" << std::endl; exit(-1);
    }
    // one int pool, one double pool, one complex pool
    const size_t ni = (size_t)3 * MSG_BIT_LEN + 2 * MSG_SYM_LEN + 3 * CODE_SYM_LEN + 3 * CODE_BIT_LEN + PUN_SYM_LEN + PUN_BIT_LEN + MOD_BIT_LEN + 16;
    ibuf_.assign(ni, 0);
    int *ip = ibuf_.data();
    auto take = [&](int n) { int *r = ip; ip += n; return r; };
    TX_MSG_BIT = take(MSG_BIT_LEN); TX_MSG_BIT_beforeCRC = take(MSG_BIT_LEN); RX_MSG_BIT = take(MSG_BIT_LEN);
    TX_MSG_SYM = take(MSG_SYM_LEN); RX_MSG_SYM = take(MSG_SYM_LEN);
    TX_CODE_SYM = take(CODE_SYM_LEN); RX_DECODE_SYM = take(CODE_SYM_LEN); ReliableSeri_Symbol = take(CODE_SYM_LEN);
    TX_CODE_BIT = take(CODE_BIT_LEN); RX_DECODE_BIT = take(CODE_BIT_LEN); ReliableSeri_BIT = take(CODE_BIT_LEN);
    PUN_SYM = take(PUN_SYM_LEN + 1); PUN_BIT = take(PUN_BIT_LEN + 1); TX_MOD_BIT = take(MOD_BIT_LEN);
    PUN_SYM[PUN_SYM_LEN] = -1; PUN_BIT[PUN_BIT_LEN] = -1;      // sentinels: the reference reads one past the end
    for (int i = 0; i < PUN_SYM_LEN; ++i) {
        PUN_SYM[i] = NBLDPC.PuncturePosition[i];
        for (int b = 0; b < Bit_Len_PerSYM; ++b) PUN_BIT[i * Bit_Len_PerSYM + b] = Bit_Len_PerSYM * PUN_SYM[i] + b;
    }
    cbuf_.assign((size_t)modOrder + 2 * MOD_SYM_LEN, CComplex());
    CONSTELLATION = cbuf_.data(); TX_MOD_SYM = CONSTELLATION + modOrder; RX_MOD_SYM = TX_MOD_SYM + MOD_SYM_LEN;
    std::ifstream fc(sim.ConstellationFileName);
    if (!fc.is_open()) { std::cerr << "Cannot open " << sim.ConstellationFileName << std::endl; exit(-1); }
    std::string label; int index;
    for (int k = 0; k < modOrder; ++k) {          // "Point: idx Real: re Imag: im" (Comm.cpp:117-126)
        fc >> label >> index;
        fc >> label >> CONSTELLATION[index].Real >> label >> CONSTELLATION[index].Image;
    }
    dbuf_.assign((size_t)CODE_BIT_LEN + (size_t)CODE_SYM_LEN * (GFq - 1), 0.0);
    RX_LLR_BIT = dbuf_.data();
    rows_.resize(CODE_SYM_LEN);
    for (int s = 0; s < CODE_SYM_LEN; ++s) rows_[s] = dbuf_.data() + CODE_BIT_LEN + (size_t)s * (GFq - 1);
    RX_LLR_SYM = rows_.data();
    return false;
}

double CComm::SetEbN0(CSimulation &sim, int parallel_order) {
    reset_source(*this, sim.randomseed, parallel_order);        // every SNR point restarts the streams
    const double EbN0 = pow(10.0, sim.EbN0 / 10.0);
    sigma_n = 1.0 / sqrt(2 * MOD_BIT_PER_SYM * CodeRate * EbN0);
    return sigma_n;
}

int CComm::Transmission() {
    GenerateMessage(); Encode(); Puncture(); Modulate(); Channel_AWGN(); DemodulateDevice(); Decode();      // front end + decoder on the GPU
    return 0;
}

int CComm::DemodulateDevice(void) {
    const int per = modOrder == 2 ? 1 : 2;
    std::vector<double> y((size_t)MOD_SYM_LEN * per), cons;
    for (int s = 0; s < MOD_SYM_LEN; ++s) {
        if (per == 1) y[s] = RX_MOD_SYM[s].Real;
        else { y[2 * s] = RX_MOD_SYM[s].Real; y[2 * s + 1] = RX_MOD_SYM[s].Image; }
    }
    if (per == 2) for (int k = 0; k < modOrder; ++k) { cons.push_back(CONSTELLATION[k].Real); cons.push_back(CONSTELLATION[k].Image); }
    return NBLDPC.DemapBatch(modOrder, y.data(), 1, per == 2 ? cons.data() : nullptr, sigma_n, RX_LLR_SYM[0]);   // the rows of RX_LLR_SYM are one dense block
}

int CComm::GenPN(void) {                        // 11-stage LFSR, feedback reg[10]^reg[3] after the shift (Comm.cpp:241-252)
    for (int i = 10; i >= 1; --i) regPN[i] = regPN[i - 1];
    regPN[0] = regPN[10] ^ regPN[3];
    return regPN[10];
}

int CComm::GenerateMessage() {
    for (int b = 0; b < MSG_BIT_LEN - crcLen; ++b) {
        if (randomMsg) { for (int i = 0; i < parallel_num - 1; ++i) GenPN(); TX_MSG_BIT_beforeCRC[b] = GenPN(); }
        else TX_MSG_BIT_beforeCRC[b] = 0;
    }
    CRCEncode(TX_MSG_BIT, TX_MSG_BIT_beforeCRC, MSG_BIT_LEN - crcLen, crcLen, 0);
    for (int s = 0; s < MSG_SYM_LEN; ++s) {       // bit b of symbol s is BIT[s*p + b]
        int v = 0;
        for (int b = Bit_Len_PerSYM - 1; b >= 0; --b) v = 2 * v + TX_MSG_BIT[Bit_Len_PerSYM * s + b];
        TX_MSG_SYM[s] = v;
    }
    return 0;
}

static void crc_poly(int crcLen, int type24, int (&G)[25]) {   // generator taps (Comm.cpp:513-531)
    memset(G, 0, sizeof G);
    static const int t8[] = {0, 1, 3, 4, 7, 8}, t16[] = {0, 5, 12, 16};
    static const int t24a[] = {0, 1, 3, 4, 5, 6, 7, 10, 11, 14, 17, 18, 23, 24}, t24b[] = {0, 1, 5, 6, 23, 24};
    if (crcLen == 8) for (int t : t8) G[t] = 1;
    if (crcLen == 16) for (int t : t16) G[t] = 1;
    if (crcLen == 24) { if (!type24) for (int t : t24a) G[t] = 1; else for (int t : t24b) G[t] = 1; }
}

void CComm::CRCEncode(int *seqOut, int *seqIn, int seqInLen, int crcLen, int crc24Type) {
    int G[25], reg[24] = {0};
    crc_poly(crcLen, crc24Type, G);
    if (seqIn != seqOut) memcpy(seqOut, seqIn, seqInLen * sizeof(int));
    if (crcLen == 0) return;
    for (int i = 0; i < seqInLen; ++i) {
        const int fb = reg[crcLen - 1] ^ seqIn[i];
        for (int j = crcLen - 1; j > 0; --j) reg[j] = reg[j - 1] ^ (G[j] && fb);
        reg[0] = fb;
    }
    for (int i = 0; i < crcLen; ++i) seqOut[i + seqInLen] = reg[crcLen - 1 - i];
}

int CComm::CrcCheck(int *seqIn, int seqInLen, int crcLen, int crc24Type) {
    if (crcLen == 0) return 1;
    int G[25], reg[24] = {0}, ones = 0;
    crc_poly(crcLen, crc24Type, G);
    for (int i = 0; i < seqInLen; ++i) {
        const int fb = reg[crcLen - 1];
        for (int j = crcLen - 1; j > 0; --j) reg[j] = reg[j - 1] ^ (G[j] && fb);
        reg[0] = fb ^ seqIn[i];
        ones += seqIn[i];
    }
    int rem = 0;
    for (int i = 0; i < crcLen; ++i) rem += reg[i];
    return rem == 0 && ones != 0;                // an all-zero sequence never passes (Comm.cpp:618-628)
}

int CComm::Encode(void) {
    if (randomMsg) NBLDPC.Encode(TX_MSG_SYM, TX_CODE_SYM);
    else for (int s = 0; s < CODE_SYM_LEN; ++s) TX_CODE_SYM[s] = 0;
    for (int s = 0; s < CODE_SYM_LEN; ++s)
        for (int b = 0; b < Bit_Len_PerSYM; ++b) TX_CODE_BIT[s * Bit_Len_PerSYM + b] = (TX_CODE_SYM[s] >> b) & 1;
    for (int b = 0; b < MSG_BIT_LEN; ++b) TX_MSG_BIT[b] = TX_CODE_BIT[b];        // errors are counted on these (Comm.cpp:281-285)
    return 0;
}

int CComm::Puncture(void) {
    int pi = 0, mb = 0;
    for (int b = 0; b < CODE_BIT_LEN; ++b) {
        if (PUN_BIT_LEN != 0 && PUN_BIT[pi] == b) ++pi;
        else TX_MOD_BIT[mb++] = TX_CODE_BIT[b];
    }
    return 0;
}

int CComm::Modulate(void) {
    for (int s = 0; s < MOD_SYM_LEN; ++s) {
        int idx = 0;                               // MSB-first index from LSB-first bits, as the reference
        for (int b = 0; b < MOD_BIT_PER_SYM; ++b) idx += TX_MOD_BIT[s * MOD_BIT_PER_SYM + b] << (MOD_BIT_PER_SYM - 1 - b);
        TX_MOD_SYM[s] = CONSTELLATION[idx];
    }
    return 0;
}

int CComm::Channel_AWGN() {
    for (int s = 0; s < MOD_SYM_LEN; ++s) {        // noise on Re then Im, also for BPSK (Comm.cpp:331-335)
        RX_MOD_SYM[s].Real = TX_MOD_SYM[s].Real + Rand.Rand_Norm(0, sigma_n);
        RX_MOD_SYM[s].Image = TX_MOD_SYM[s].Image + Rand.Rand_Norm(0, sigma_n);
    }
    return 0;
}

// Host form of the demapper, kept so that the reference's RX_LLR_SYM / RX_LLR_BIT buffers stay valid for callers
// that read them; the batched path demaps on the device instead (bit-identical: tests/test_decode_gpu.py).
int CComm::Demodulate(void) {
    if (modOrder == 2) {
        int pi = 0;
        for (int b = 0; b < CODE_BIT_LEN; ++b) {
            if (PUN_BIT_LEN != 0 && PUN_BIT[pi] == b) { ++pi; RX_LLR_BIT[b] = 0; }
            else RX_LLR_BIT[b] = -2 * RX_MOD_SYM[b - pi].Real / (sigma_n * sigma_n);
        }
        for (int s = 0; s < CODE_SYM_LEN; ++s)
            for (int q = 1; q < GFq; ++q) {
                double a = 0;
                for (int b = 0; b < Bit_Len_PerSYM; ++b) if (q & (1 << b)) a += RX_LLR_BIT[s * Bit_Len_PerSYM + b];
                RX_LLR_SYM[s][q - 1] = a;
            }
    } else {
        int pi = 0;
        const CComplex c0 = CONSTELLATION[0];
        for (int s = 0; s < CODE_SYM_LEN; ++s) {
            if (PUN_SYM_LEN != 0 && PUN_SYM[pi] == s) { ++pi; for (int q = 1; q < GFq; ++q) RX_LLR_SYM[s][q - 1] = 0; continue; }
            const CComplex y = RX_MOD_SYM[s - pi];
            for (int q = 1; q < GFq; ++q) {
                const CComplex cq = CONSTELLATION[q];
                RX_LLR_SYM[s][q - 1] = ((2 * y.Real - c0.Real - cq.Real) * (cq.Real - c0.Real) +
                                        (2 * y.Image - c0.Image - cq.Image) * (cq.Image - c0.Image)) / (2 * sigma_n * sigma_n);
            }
        }
    }
    return 0;
}

int CComm::Decode(void) {
    DecodeCorrect = NBLDPC.Decoding(RX_LLR_SYM, RX_DECODE_SYM, ReliableSeri_Symbol, ReliableSeri_BIT);   // ⇢ device
    iterations = NBLDPC.lastIterations;
    tems_route_num = 0;
    return UnpackDecision();
}

int CComm::UnpackDecision(void) {
    for (int s = 0; s < CODE_SYM_LEN; ++s)
        for (int b = 0; b < Bit_Len_PerSYM; ++b) RX_DECODE_BIT[s * Bit_Len_PerSYM + b] = (RX_DECODE_SYM[s] >> b) & 1;
    for (int s = 0; s < MSG_SYM_LEN; ++s) RX_MSG_SYM[s] = RX_DECODE_SYM[s];
    for (int b = 0; b < MSG_BIT_LEN; ++b) RX_MSG_BIT[b] = RX_DECODE_BIT[b];
    return 0;
}

int CComm::Err(CSimulation &sim) {              // Comm.cpp:446-503: the first K symbols / K*p bits only
    double errSym = 0, errBit = 0;
    for (int s = 0; s < MSG_SYM_LEN; ++s) errSym += TX_MSG_SYM[s] != RX_MSG_SYM[s];
    for (int b = 0; b < MSG_BIT_LEN; ++b) errBit += TX_MSG_BIT[b] != RX_MSG_BIT[b];
    const int crc_ok = CrcCheck(RX_MSG_BIT, MSG_BIT_LEN, crcLen, 1);           // CRC24B on check, CRC24A on encode
    sim.errSym += errSym; sim.errBit += errBit; sim.errFrame += errSym != 0 ? 1 : 0;
    if (crc_ok && errSym != 0) { sim.U_errSym += errSym; sim.U_errBit += errBit; sim.U_errFrame += 1; }
    sim.sumIterations += iterations;
    const double frames = sim.simCycle * sim.parallel;
    sim.SER = sim.errSym / (frames * MSG_SYM_LEN); sim.BER = sim.errBit / (frames * MSG_BIT_LEN); sim.FER = sim.errFrame / frames;
    sim.U_SER = sim.U_errSym / (frames * MSG_SYM_LEN); sim.U_BER = sim.U_errBit / (frames * MSG_BIT_LEN); sim.U_FER = sim.U_errFrame / frames;
    return 0;
}

void CCommBatch::Run(CComm *w, int count) {
    if (count <= 0) return;
    const int L = w[0].MOD_SYM_LEN, N = w[0].CODE_SYM_LEN, per = w[0].modOrder == 2 ? 1 : 2;
    y_.resize((size_t)count * L * per); out_.resize((size_t)count * N); conv_.resize(count); iters_.resize(count);
    for (int i = 0; i < count; ++i) {
        w[i].GenerateMessage(); w[i].Encode(); w[i].Puncture(); w[i].Modulate(); w[i].Channel_AWGN();
        for (int s = 0; s < L; ++s) {
            if (per == 1) y_[(size_t)i * L + s] = w[i].RX_MOD_SYM[s].Real;
            else { y_[((size_t)i * L + s) * 2] = w[i].RX_MOD_SYM[s].Real; y_[((size_t)i * L + s) * 2 + 1] = w[i].RX_MOD_SYM[s].Image; }
        }
    }
    std::vector<double> cons;
    if (per == 2) for (int k = 0; k < w[0].modOrder; ++k) { cons.push_back(w[0].CONSTELLATION[k].Real); cons.push_back(w[0].CONSTELLATION[k].Image); }
    w[0].NBLDPC.DecodingBatch(w[0].modOrder, y_.data(), count, per == 2 ? cons.data() : nullptr, w[0].sigma_n, out_.data(), conv_.data(), iters_.data());
    for (int i = 0; i < count; ++i) {
        memcpy(w[i].RX_DECODE_SYM, out_.data() + (size_t)i * N, N * sizeof(int));
        w[i].DecodeCorrect = conv_[i] != 0; w[i].iterations = iters_[i];
        w[i].UnpackDecision();
    }
}
