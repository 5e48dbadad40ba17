// One worker of the link simulation with the reference's CComm call surface (Comm.h:13-98): every stage and every
// buffer is public, Transmission() runs one frame.  The stages up to the channel run on the host exactly as the
// reference runs them (PN source, CRC, systematic encoder, LSB-first bit packing, puncturing, constellation
// mapping, the reference's own noise generator) so that the decoder sees identical inputs; Demodulate + Decode are
// the hot path and go to the GPU -- per frame through CNBLDPC::Decoding (the reference's own call), or for a whole
// round of workers at once through CCommBatch (one launch sequence for `parallel` frames).
#pragma once
#include "codec.h"
#include "rand_ref.h"
#include "sim_control.h"
#include <vector>

struct CComplex { double Real = 0, Image = 0; };

class CComm {
public:
    int GFq = 0, parallel_num = 1, Bit_Len_PerSYM = 0;
    int tems_route_num = 0;
    CNBLDPC NBLDPC;
    CRand Rand;
    bool Initial(CSimulation &sim, int parallel_order);
    double SetEbN0(CSimulation &sim, int parallel_order);
    int Transmission(void);

    int randomMsg = 0;
    int *TX_MSG_BIT_beforeCRC = nullptr;
    int regPN[11];
    int GenerateMessage(void);
    int GenPN(void);

    int crcLen = 0, crcLen_correct = 0, MSG_SYM_LEN = 0, MSG_BIT_LEN = 0;
    int *TX_MSG_SYM = nullptr, *TX_MSG_BIT = nullptr;
    void CRCEncode(int *seqOut, int *seqIn, int seqInLen, int crcLen, int crc24Type);
    int CrcCheck(int *seqIn, int seqInLen, int crcLen, int crc24Type);

    int *TX_CODE_SYM = nullptr, *TX_CODE_BIT = nullptr;
    int CODE_SYM_LEN = 0, CODE_BIT_LEN = 0;
    int Encode(void);

    int *PUN_SYM = nullptr; int PUN_SYM_LEN = 0;
    int *PUN_BIT = nullptr; int PUN_BIT_LEN = 0;
    double CodeRate = 0;
    int Puncture(void);

    CComplex *CONSTELLATION = nullptr, *TX_MOD_SYM = nullptr;
    int modOrder = 2, MOD_BIT_PER_SYM = 1, MOD_SYM_LEN = 0;
    int *TX_MOD_BIT = nullptr; int MOD_BIT_LEN = 0;
    int Modulate(void);

    CComplex *RX_MOD_SYM = nullptr;
    double sigma_n = 1;
    int Channel_AWGN(void);

    double *RX_LLR_BIT = nullptr;
    double **RX_LLR_SYM = nullptr;
    int Demodulate(void);                       // host form (runs without a GPU; what the reference computes, same operation order)
    int DemodulateDevice(void);                 // the same LLRs from the device demapper (nbl_demap_bpsk / nbl_demap_qam); Transmission() uses this one

    int *RX_DECODE_SYM = nullptr, *RX_DECODE_BIT = nullptr, *RX_MSG_SYM = nullptr, *RX_MSG_BIT = nullptr;
    int *ReliableSeri_Symbol = nullptr, *ReliableSeri_BIT = nullptr;
    bool DecodeCorrect = false;
    int iterations = 0;                         // iteration count of the last decode (extra, not in the reference)
    int Decode(void);
    int UnpackDecision(void);                   // second half of Decode(): RX_DECODE_SYM -> bits / message
    int Err(CSimulation &sim);

private:
    std::vector<int> ibuf_;
    std::vector<double> dbuf_;
    std::vector<double *> rows_;
    std::vector<CComplex> cbuf_;
};

// One round of `parallel` workers decoded in a single batched call (replaces the parallel_for of main.cpp:46).
class CCommBatch {
public:
    void Run(CComm *workers, int count);        // host stages per worker, ONE device demap+decode, unpack per worker
private:
    std::vector<double> y_;
    std::vector<int> out_, conv_, iters_;
};
