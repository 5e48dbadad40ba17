#include "codec.h"
#include <cstdlib>
#include <iostream>

static void die(const char *what) {
    std::cerr << what << ": " << nbl_last_error() << std::endl;
    exit(-1);                                   // the reference's error convention for configuration faults
}

CNBLDPC::~CNBLDPC() {
    if (dec_) nbl_decoder_destroy(dec_);
    if (code_) nbl_code_destroy(code_);
}

bool CNBLDPC::Initial(CSimulation &sim) {
    GFq = sim.GFq;
    maxIter = sim.maxIter;
    GF.Initial(GFq);
    if (nbl_code_load_file(&code_, sim.NonBinaryFileName.c_str(), 0) != NBL_OK) die("code file");
    nbl_code_info info;
    nbl_code_get_info(code_, &info);
    if (info.q != GFq) { std::cerr << "profile GFq and the code file's field size differ" << std::endl; exit(-1); }
    CodeLen = info.N; ChkLen = info.M; maxVarDegree = info.max_dv; maxChkDegree = info.max_dc;
    const int E = info.E;
    // link tables, exposed with the reference's names (NBLDPC.h:32-39)
    graph_.assign((size_t)(CodeLen + 1) + (ChkLen + 1) + 6 * (size_t)E + CodeLen + ChkLen, 0);
    int *chk_ptr = graph_.data(), *var_ptr = chk_ptr + ChkLen + 1;
    int *chk_var = var_ptr + CodeLen + 1, *chk_coef = chk_var + E, *chk_dv = chk_coef + E;
    int *var_chk = chk_dv + E, *var_coef = var_chk + E, *var_dc = var_coef + E;
    VarDegree = var_dc + E; ChkDegree = VarDegree + CodeLen;
    nbl_code_get_graph(code_, chk_ptr, chk_var, chk_coef, chk_dv, var_ptr, var_chk, var_coef, var_dc);
    rows_.assign(3 * (size_t)(CodeLen + ChkLen), nullptr);
    VarLink = rows_.data(); VarLinkGFe = VarLink + CodeLen; VarLinkDc = VarLinkGFe + CodeLen;
    ChkLink = VarLinkDc + CodeLen; ChkLinkGFe = ChkLink + ChkLen; ChkLinkDv = ChkLinkGFe + ChkLen;
    for (int v = 0; v < CodeLen; ++v) {
        VarDegree[v] = var_ptr[v + 1] - var_ptr[v];
        VarLink[v] = var_chk + var_ptr[v]; VarLinkGFe[v] = var_coef + var_ptr[v]; VarLinkDc[v] = var_dc + var_ptr[v];
    }
    for (int r = 0; r < ChkLen; ++r) {
        ChkDegree[r] = chk_ptr[r + 1] - chk_ptr[r];
        ChkLink[r] = chk_var + chk_ptr[r]; ChkLinkGFe[r] = chk_coef + chk_ptr[r]; ChkLinkDv[r] = chk_dv + chk_ptr[r];
    }
    static thread_local std::vector<int> punct;                 // puncture rule: variables of the given degree
    punct.assign(CodeLen, 0);
    PunctureLen = nbl_code_puncture_set(code_, sim.PuntureVarDegree, punct.data());
    PuncturePosition = new int[PunctureLen > 0 ? PunctureLen : 1];
    for (int i = 0; i < PunctureLen; ++i) PuncturePosition[i] = punct[i];

    DecodeMethod = sim.decodeMethod;
    EMS_Nm = sim.ems_nm; EMS_Nc = sim.ems_nc; EMS_Correction_Factor = sim.ems_factor; EMS_Correction_Offset = sim.ems_offset;
    TEMS_Nr = sim.tems_nr; TEMS_Nc = sim.tems_nc; TEMS_Factor = sim.tems_factor; TEMS_Offset = sim.tems_offset;
    basic_set_Nm = sim.bs_tems_nm; basic_set_Nc = sim.bs_tems_nc; basic_set_factor = sim.bs_tems_factor; basic_set_offset = sim.bs_tems_offset;   // NBLDPC.cpp:334-337
    OSD_Order = sim.OSD_order; OSD_Alpha = sim.OSD_factor; OSD_flag = sim.OSD_flag;     // NBLDPC.cpp:381-383
    CRCLen = sim.crcLen; CRCrate = sim.crc_correctLen;                                  // NBLDPC.cpp:437-438
    precision_ = sim.precision; device_ = sim.device;
    if (DecodeMethod == EMS_DECODE && EMS_Nm > GFq) { std::cerr << "EMS configuration error! EMS_Nm is too large!" << std::endl; exit(-1); }

    post_.assign((size_t)CodeLen * (GFq - 1), 0.0);
    post_rows_.resize(CodeLen);
    for (int v = 0; v < CodeLen; ++v) post_rows_[v] = post_.data() + (size_t)v * (GFq - 1);
    L_post = post_rows_.data();
    if (sim.randomMsg) InitialEncode();
    return true;
}

// Systematic encoder set-up.  Semantics of the reference (NBLDPC.cpp:477-560, 1474-1538): rows are processed from
// the last to the first with pivot column K + row; a zero pivot is fixed by a row swap from above, else by a
// column swap from the left (recorded, undone after encoding); the pivot column is cleared above the pivot and the
// row is scaled to a unit pivot.  What remains left of each pivot is the forward-substitution rule of Encode().
int CNBLDPC::InitialEncode(void) {
    const int N = CodeLen, M = ChkLen, K = N - M;
    std::vector<std::vector<int>> H(M, std::vector<int>(N, 0));
    for (int r = 0; r < M; ++r)
        for (int d = 0; d < ChkDegree[r]; ++d) H[r][ChkLink[r][d]] = ChkLinkGFe[r][d];
    enc_swap_src_.clear(); enc_swap_dst_.clear();
    for (int r = M - 1; r >= 0; --r) {
        const int pc = K + r;
        if (H[r][pc] == 0) {
            bool fixed = false;
            for (int up = r - 1; up >= 0 && !fixed; --up)
                if (H[up][pc] != 0) { std::swap(H[r], H[up]); fixed = true; }
            for (int left = pc - 1; left >= 0 && !fixed; --left)
                if (H[r][left] != 0) {
                    for (int rr = 0; rr < M; ++rr) std::swap(H[rr][pc], H[rr][left]);
                    enc_swap_src_.push_back(pc); enc_swap_dst_.push_back(left);
                    fixed = true;
                }
            if (!fixed) { std::cerr << "NB matrix is not full rank" << std::endl; exit(-1); }
        }
        const int inv = GF.GFInverse(H[r][pc]);
        for (int up = r - 1; up >= 0; --up)
            if (H[up][pc] != 0) {
                const int x = GF.GFMultiply(inv, H[up][pc]);
                for (int c = 0; c < N; ++c) H[up][c] = GF.GFAdd(H[up][c], GF.GFMultiply(x, H[r][c]));
            }
        for (int c = 0; c <= pc; ++c) H[r][c] = GF.GFMultiply(H[r][c], inv);
    }
    enc_link_.assign(M, {}); enc_coef_.assign(M, {});
    for (int r = 0; r < M; ++r)
        for (int c = 0; c < K + r; ++c)
            if (H[r][c] != 0) { enc_link_[r].push_back(c); enc_coef_[r].push_back(H[r][c]); }
    return 0;
}

// NBLDPC.cpp:562-604: forward substitution, undo the column swaps, hand the first K symbols back as the message
int CNBLDPC::Encode(int *msg_sym, int *code_sym) {
    const int N = CodeLen, M = ChkLen, K = N - M;
    for (int c = 0; c < K; ++c) code_sym[c] = msg_sym[c];
    for (int c = K; c < N; ++c) code_sym[c] = 0;
    for (int r = 0; r < M; ++r) {
        int acc = 0;
        for (size_t d = 0; d < enc_link_[r].size(); ++d) acc = GF.GFAdd(acc, GF.GFMultiply(enc_coef_[r][d], code_sym[enc_link_[r][d]]));
        code_sym[K + r] = acc;
    }
    for (int i = (int)enc_swap_src_.size() - 1; i >= 0; --i) std::swap(code_sym[enc_swap_src_[i]], code_sym[enc_swap_dst_[i]]);
    for (int c = 0; c < K; ++c) msg_sym[c] = code_sym[c];
    return 0;
}

int CNBLDPC::DecideLLRVector(double *LLR) {      // NBLDPC.cpp:1542-1562
    double mx = 0; int a = 0;
    for (int k = 0; k < GFq - 1; ++k) if (LLR[k] > mx) { mx = LLR[k]; a = k + 1; }
    return mx <= 0 ? 0 : a;
}

int CNBLDPC::ensure_decoder(int batch) {
    if (dec_ && dec_batch_ >= batch) return 0;
    if (dec_) { nbl_decoder_destroy(dec_); dec_ = nullptr; }
    nbl_decoder_config cfg{};
    cfg.method = DecodeMethod; cfg.max_iter = maxIter; cfg.precision = precision_; cfg.early_stop = 1; cfg.max_batch = batch; cfg.device = device_;
    if (PunctureLen > 0) cfg.tie_mode = NBL_TIES_REFERENCE;    // erased symbols make configuration costs tie exactly
    if (DecodeMethod == EMS_DECODE) { cfg.p0 = EMS_Nm; cfg.p1 = EMS_Nc; cfg.factor = EMS_Correction_Factor; cfg.offset = EMS_Correction_Offset; }
    else if (DecodeMethod == T_EMS_DECODE) { cfg.p0 = TEMS_Nr; cfg.p1 = TEMS_Nc; cfg.factor = TEMS_Factor; cfg.offset = TEMS_Offset; }
    else if (DecodeMethod == BS_TEMS_DECODE) { cfg.p0 = basic_set_Nm; cfg.p1 = basic_set_Nc; cfg.factor = basic_set_factor; cfg.offset = basic_set_offset; }
    else { cfg.factor = 1; }
    if (DecodeMethod == OSD_DECODE) cfg.method = BP_DECODE;     // method 6 runs OSD alone (nbl_osd_batch); the handle still needs a decoder
    if (nbl_decoder_create(&dec_, code_, &cfg, nullptr) != NBL_OK) die("decoder");
    if (OSD_Order >= 0 || DecodeMethod == OSD_DECODE) {         // the reference's tail of every Decoding_* (NBLDPC.cpp:770-775, 921-925, 1135-1139)
        nbl_osd_config oc{};
        oc.order = OSD_Order < 0 ? 0 : OSD_Order; oc.factor = OSD_Alpha; oc.flag = OSD_flag; oc.crc_len = CRCLen; oc.crc_correct_len = CRCrate;
        if (nbl_decoder_enable_osd(dec_, &oc) != NBL_OK) die("OSD");
    }
    dec_batch_ = batch;
    return 0;
}

int CNBLDPC::run_method(int method, double **L_ch, int *out) {
    if (method != DecodeMethod) {                // a direct Decoding_XXX call for another method: rebuild the decoder
        DecodeMethod = method;
        if (dec_) { nbl_decoder_destroy(dec_); dec_ = nullptr; }
    }
    ensure_decoder(1);
    const size_t q1 = GFq - 1;
    lch_.resize((size_t)CodeLen * q1);
    for (int v = 0; v < CodeLen; ++v) std::copy(L_ch[v], L_ch[v] + q1, lch_.begin() + (size_t)v * q1);   // ragged -> dense
    int conv = 0, iters = 0;
    if (method == OSD_DECODE) {                  // Decoding_OSD_bit(L_ch, out, rely, 1) returns 0 (OSD.h:198)
        if (nbl_osd_batch(dec_, lch_.data(), 1, out) != NBL_OK) die("OSD");
        lastIterations = 0;
        return 0;
    }
    if (nbl_decode_batch(dec_, lch_.data(), 1, out, &conv, &iters, post_.data()) != NBL_OK) die("decode");
    lastIterations = iters;
    return conv;
}

int CNBLDPC::Decoding_BP(double **L_ch, int *DecodeOutput) { return run_method(BP_DECODE, L_ch, DecodeOutput); }
int CNBLDPC::Decoding_EMS(double **L_ch, int *DecodeOutput, int *) { return run_method(EMS_DECODE, L_ch, DecodeOutput); }
int CNBLDPC::Decoding_TEMS(double **L_ch, int *DecodeOutput, int *, int *) { return run_method(T_EMS_DECODE, L_ch, DecodeOutput); }
int CNBLDPC::Decoding_BS_TEMS(double **L_ch, int *DecodeOutput, int *, int *) { return run_method(BS_TEMS_DECODE, L_ch, DecodeOutput); }
// OSD.h:5: the stand-alone form (flag 1) only; with flag 0 the reference feeds it the state of the iterative decoder
// that has just failed, which here happens on the device inside nbl_decode_batch
int CNBLDPC::Decoding_OSD_bit(double **L_ch, int *DecodeOutput, int *, int flag) {
    if (!flag) { std::cerr << "Decoding_OSD_bit with flag 0 runs inside the iterative decoders on this path" << std::endl; exit(-1); }
    return run_method(OSD_DECODE, L_ch, DecodeOutput);
}

int CNBLDPC::Decoding(double **L_ch, int *DecodeOutput, int *RelySeri_symbol, int *RelySeri_bit) {
    switch (DecodeMethod) {                      // NBLDPC.cpp:607-641
    case BP_DECODE: return Decoding_BP(L_ch, DecodeOutput);
    case EMS_DECODE: return Decoding_EMS(L_ch, DecodeOutput, RelySeri_bit);
    case T_EMS_DECODE: return Decoding_TEMS(L_ch, DecodeOutput, RelySeri_symbol, RelySeri_bit);
    case MinMax_DECODE: std::cout << "Min Max algorithm has not been developed" << std::endl; exit(-1);
    case T_MinMax_DECODE: std::cout << "Trellis Min Max algorithm has not been developed" << std::endl; exit(-1);
    case BS_TEMS_DECODE: return Decoding_BS_TEMS(L_ch, DecodeOutput, RelySeri_symbol, RelySeri_bit);
    case OSD_DECODE: return Decoding_OSD_bit(L_ch, DecodeOutput, RelySeri_bit, 1);
    default: std::cout << "Decode Method is not specified!" << std::endl; exit(-1);
    }
    return 0;
}

// CComm::Demodulate on the device (bit-identical with the host form, tests/test_decode_gpu.py): y -> L_ch [B][N][q-1] host doubles
int CNBLDPC::DemapBatch(int modulation, const double *y, int B, const double *constellation, double sigma, double *L_ch) {
    ensure_decoder(B);
    const int rc = modulation == 2 ? nbl_demap_bpsk(dec_, y, B, sigma, PuncturePosition, PunctureLen, L_ch)
                                   : nbl_demap_qam(dec_, y, B, constellation, sigma, PuncturePosition, PunctureLen, L_ch);
    if (rc != NBL_OK) die("demap");
    return 0;
}

int CNBLDPC::DecodingBatch(int modulation, const double *y, int B, const double *constellation, double sigma,
                           int *DecodeOutput, int *converged, int *iterations) {
    ensure_decoder(B);
    if (DecodeMethod == OSD_DECODE) {            // method 6: demap, then OSD alone on every frame; "not converged", as the reference returns
        lch_.resize((size_t)B * CodeLen * (GFq - 1));
        const int rc = modulation == 2 ? nbl_demap_bpsk(dec_, y, B, sigma, PuncturePosition, PunctureLen, lch_.data())
                                       : nbl_demap_qam(dec_, y, B, constellation, sigma, PuncturePosition, PunctureLen, lch_.data());
        if (rc != NBL_OK || nbl_osd_batch(dec_, lch_.data(), B, DecodeOutput) != NBL_OK) die("demap+OSD");
        for (int f = 0; f < B; ++f) { converged[f] = 0; iterations[f] = 0; }
        return 0;
    }
    if (nbl_demap_decode_batch(dec_, modulation, y, B, constellation, sigma, PuncturePosition, PunctureLen, DecodeOutput, converged, iterations) != NBL_OK)
        die("demap+decode");
    return 0;
}
