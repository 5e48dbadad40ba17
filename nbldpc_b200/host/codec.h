// The codec object with the reference's CNBLDPC call surface (NBLDPC.h:13-196) -- Initial / Encode / Decoding /
// Decoding_BP / Decoding_EMS / Decoding_TEMS and the public members the link chain reads -- implemented on the
// C ABI of include/nbldpc_b200.h.  Graph, GF tables and the systematic encoder live on the host; every decoder
// call goes to the GPU (there is no CPU decoder here).  DecodingBatch is the batched form the driver uses.
#pragma once
#include "gf_field.h"
#include "sim_control.h"
#include "../../include/nbldpc_b200.h"
#include <vector>

class CNBLDPC {
public:
    CNBLDPC() = default;
    ~CNBLDPC();
    CNBLDPC(const CNBLDPC &) = delete;
    CNBLDPC &operator=(const CNBLDPC &) = delete;

    int GFq = 0;
    CGF GF;
    int maxIter = 0;
    int CodeLen = 0, ChkLen = 0;
    int *PuncturePosition = nullptr;
    int PunctureLen = 0;
    int maxVarDegree = 0, maxChkDegree = 0;
    int *VarDegree = nullptr, *ChkDegree = nullptr;
    int **VarLink = nullptr, **ChkLink = nullptr, **VarLinkGFe = nullptr, **ChkLinkGFe = nullptr;
    int **VarLinkDc = nullptr, **ChkLinkDv = nullptr;
    int DecodeMethod = 0;
    double **L_post = nullptr;                 // posterior of the last Decoding() call, [CodeLen][GFq-1]
    int EMS_Nm = 0, EMS_Nc = 0; double EMS_Correction_Factor = 1, EMS_Correction_Offset = 0;
    int TEMS_Nr = 0, TEMS_Nc = 0; double TEMS_Factor = 1, TEMS_Offset = 0;
    int basic_set_Nm = 0, basic_set_Nc = 0; double basic_set_factor = 1, basic_set_offset = 0;   // NBLDPC.h:169-172
    int OSD_Order = -1, OSD_flag = 0; double OSD_Alpha = 1;   // NBLDPC.h: OSD_Order / OSD_flag / OSD_Alpha
    int CRCLen = 0, CRCrate = 0;                               // crcLen, crc_correctLen
    double tems_route_num = 0;                 // kept for source compatibility; the GPU search does not count leaves
    int lastIterations = 0;                    // iteration count of the last Decoding() call (not exposed by the reference)

    bool Initial(CSimulation &sim);
    int InitialEncode(void);
    int Encode(int *msg_sym, int *code_sym);
    int Decoding(double **L_ch, int *DecodeOutput, int *RelySeri_symbol, int *RelySeri_bit);
    int Decoding_BP(double **L_ch, int *DecodeOutput);
    int Decoding_EMS(double **L_ch, int *DecodeOutput, int *RelySeri_bit);
    int Decoding_BS_TEMS(double **L_ch, int *DecodeOutput, int *RelySeri_symbol, int *RelySeri_bit);   // NBLDPC.h (BS_TEMS_DECODE)
    int Decoding_TEMS(double **L_ch, int *DecodeOutput, int *RelySeri_symbol, int *RelySeri_bit);
    int Decoding_OSD_bit(double **L_ch, int *DecodeOutput, int *ReliableSeri_BIT, int flag);   // OSD.h:5
    int DecideLLRVector(double *LLR);

    // batched drop-in: B frames of received samples (BPSK: [B][mod_len] real parts, QAM: [B][mod_len][2]) ->
    // decisions [B][CodeLen], return codes and iteration counts; demapping happens on the device.
    int DecodingBatch(int modulation, const double *y, int B, const double *constellation, double sigma,
                      int *DecodeOutput, int *converged, int *iterations);
    int DemapBatch(int modulation, const double *y, int B, const double *constellation, double sigma, double *L_ch);
    nbl_code *code() { return code_; }

private:
    nbl_code *code_ = nullptr;
    nbl_decoder *dec_ = nullptr;
    int dec_batch_ = 0, precision_ = 1, device_ = 0;
    std::vector<int> graph_;                   // backing store of the link tables
    std::vector<int *> rows_;
    std::vector<double> post_, lch_;
    std::vector<double *> post_rows_;
    // encoder (InitialEncode): H reduced to [A | L], L unit lower-triangular, plus the recorded column swaps
    std::vector<int> enc_swap_src_, enc_swap_dst_;
    std::vector<std::vector<int>> enc_link_, enc_coef_;
    int ensure_decoder(int batch);
    int run_method(int method, double **L_ch, int *out);
};
