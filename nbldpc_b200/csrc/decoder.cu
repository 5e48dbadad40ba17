// B200 (sm_100a) decode path: kernels + the extern "C" decoder entry points of include/nbldpc_b200.h.
//
// Data layout in HBM (per decoder, frames-in-flight = max_batch; T = float or double):
//   lch  [B][N][Q]   channel LLRs, slot 0 = 0 (the implicit LLR of symbol 0), Q = q
//   v2c  [B][E][Q]   variable->check messages, edge ids check-major so a check row's dc vectors are contiguous
//   c2v  [2][B][E][Q] check->variable messages, double buffered (iteration t reads t-1, writes t)
//   vdec [B][E] u8   decide(v2c) of the previous iteration (the damping test's "before", NBLDPC.cpp:1038)
//   dec  [B][N] u8   tentative hard decisions of the running iteration
// One launch per decoding iteration processes every (frame, check row) pair: a group of Q threads
// recomputes the posteriors of the row's variables, makes the decisions, accumulates the row syndrome,
// forms the damped v2c messages (steps A-C of the reference loop, NBLDPC.cpp:977-1052) and runs the check
// node (step D) on them while they are still in shared memory.  Messages stream HBM -> smem -> HBM once
// per iteration, fully coalesced (thread x <-> element x).
#include "internal.h"
#include "enc_internal.h"
#include "dev_common.cuh"
#include "iter_params.cuh"
#include "cn_tems.cuh"
#include "cn_ems.cuh"
#include "cn_qspa.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <vector>

using namespace nbl;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            nbl_set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_));              \
            return NBL_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

// ======================================================================================================
// kernels
// ======================================================================================================

#include "tems_warp.cuh"
#include "ems_warp.cuh"
#include "qspa_warp.cuh"
#include "osd.cuh"
#include "qspa_half.cuh"
#include "tems_half.cuh"
#include "ems_half.cuh"

// channel LLRs: double [B][N][q-1] -> T [B][N][Q] with a zero in slot 0
template <typename T> __global__ void k_convert_lch(const double *__restrict__ src, T *__restrict__ dst, size_t nvec, int q) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec * q) return;
    size_t v = i / q; int x = (int)(i % q);
    dst[i] = x == 0 ? T(0) : (T)src[v * (q - 1) + x - 1];
}

template <typename T, int Q, int METHOD>
__global__ void __launch_bounds__(Grp<Q>::THREADS) k_iterate(const IterParams<T> P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevGraph &g = P.g;
    const int gid = threadIdx.x / Q, x = threadIdx.x % Q;
    const long long w = (long long)blockIdx.x * Grp<Q>::G + gid;
    if (w >= (long long)P.B * g.M) return;
    const int f = (int)(w / g.M), row = (int)(w % g.M);
    if (P.early_stop && P.done[f]) return;

    size_t per_group;
    TemsSmem<T, Q> St; EmsSmem<T, Q> Se; QspaSmem<T, Q> Sq;
    T *V; int *sdec; RedScratch<typename Real<T>::key_t, Grp<Q>::WARPS, 2> *red;
    if (METHOD == NBL_METHOD_TEMS || METHOD == NBL_METHOD_BSTEMS) {
        // BS-TEMS and the bit-exact T-EMS search use the eta / back-pointer arrays of the n_c >= 3 layout
        const int nca = METHOD == NBL_METHOD_BSTEMS || (P.bitexact && P.p1 < 3) ? 3 : P.p1;
        per_group = TemsSmem<T, Q>::bytes(g.max_dc, nca);
        St.carve(smem_raw + per_group * gid, g.max_dc, nca); V = St.V; sdec = St.sdec; red = St.red;
    } else if (METHOD == NBL_METHOD_EMS) {
        per_group = EmsSmem<T, Q>::bytes(g.max_dc, P.p0, P.p1);
        Se.carve(smem_raw + per_group * gid, g.max_dc, P.p0, P.p1); V = Se.V; sdec = Se.sdec; red = Se.red;
    } else {
        per_group = QspaSmem<T, Q>::bytes(g.max_dc);
        Sq.carve(smem_raw + per_group * gid, g.max_dc); V = Sq.V; sdec = Sq.sdec; red = Sq.red;
    }

    const int e0 = g.chk_ptr[row], dcn = g.chk_ptr[row + 1] - e0;
    const size_t fE = (size_t)f * g.E, fN = (size_t)f * g.N;
    int phase = 0, syn = 0;
    // ---- steps A-C for the row's variables ------------------------------------------------------------
    for (int d = 0; d < dcn; ++d) {
        const int e = e0 + d, col = g.chk_var[e];
        T post = P.lch[(fN + col) * Q + x];
        T vnew = post;
        if (!P.first) {
            T own = T(0);
            for (int s = g.var_ptr[col]; s < g.var_ptr[col + 1]; ++s) {      // file order = summation order
                const int ee = g.var_edge[s];
                const T m = P.c2v_in[(fE + ee) * Q + x];
                post = post + m;
                if (ee == e) own = m;
            }
            vnew = post - own;
        }
        T vals[2] = {post, vnew}; int ds[2];
        group_decide<Q, T, 2>(vals, x, gid, red, phase, ds);
        T vfin = vnew; int dfin = ds[1];
        if (!P.first && P.w_old != T(0) && ds[1] != (int)P.vdec[fE + e]) {      // group-uniform branch
            const T vold = P.v2c[(fE + e) * Q + x];
            vfin = P.w_old * vold + P.w_new * vnew;                             // compiled with -fmad=false
            T v1[2] = {vfin, T(0)}; int d1[2];
            group_decide<Q, T, 2>(v1, x, gid, red, phase, d1);
            dfin = d1[0];
        }
        P.v2c[(fE + e) * Q + x] = vfin;
        V[d * Q + x] = vfin;
        if (x == 0) { P.vdec[fE + e] = (uint8_t)dfin; P.dec[fN + col] = (uint8_t)ds[0]; sdec[d] = dfin; }
        syn ^= g.pfwd[(size_t)e * Q + ds[0]];
    }
    if (x == 0 && syn != 0) P.flag[f] = 1;
    Grp<Q>::sync(gid);
    // ---- step D: check node ---------------------------------------------------------------------------
    T *out = P.c2v_out + fE * Q;
    if (METHOD == NBL_METHOD_TEMS) cn_tems<T, Q>(St, g, gid, x, e0, dcn, P.p0, P.p1, P.bitexact != 0, P.factor, P.offset, out);
    else if (METHOD == NBL_METHOD_BSTEMS) cn_bstems<T, Q>(St, g, gid, x, e0, dcn, P.p0, P.p1, P.factor, P.offset, phase, out);
    else if (METHOD == NBL_METHOD_EMS) cn_ems<T, Q>(Se, g, gid, x, e0, dcn, P.p0, P.p1, P.factor, P.offset, phase, out);
    else cn_qspa<T, Q>(Sq, g, gid, x, e0, dcn, out);
}

// posterior + decision only (the last iteration's steps A-B; also serves the L_post output).
// group per (frame, variable).  want_post: write doubles [B][N][q-1] for frames with sel[f] != 0.
template <typename T, int Q>
__global__ void __launch_bounds__(Grp<Q>::THREADS) k_posterior(DevGraph g, int B, int first, int early_stop, const T *__restrict__ lch,
                                                             const T *__restrict__ c2v_in, const uint8_t *__restrict__ done,
                                                             uint8_t *dec, double *post_out, const uint8_t *__restrict__ sel) {
    __shared__ RedScratch<typename Real<T>::key_t, Grp<Q>::WARPS, 2> red_all[Grp<Q>::G];
    const int gid = threadIdx.x / Q, x = threadIdx.x % Q;
    const long long w = (long long)blockIdx.x * Grp<Q>::G + gid;
    if (w >= (long long)B * g.N) return;
    const int f = (int)(w / g.N), col = (int)(w % g.N);
    if (post_out) { if (!sel[f]) return; }
    else if (early_stop && done[f]) return;
    T post = lch[((size_t)f * g.N + col) * Q + x];
    if (!first)
        for (int s = g.var_ptr[col]; s < g.var_ptr[col + 1]; ++s) post = post + c2v_in[((size_t)f * g.E + g.var_edge[s]) * Q + x];
    if (post_out) {
        if (x != 0) post_out[((size_t)f * g.N + col) * (Q - 1) + x - 1] = (double)post;
        return;
    }
    int phase = 0;
    T vals[2] = {post, T(0)}; int ds[2];
    group_decide<Q, T, 2>(vals, x, gid, &red_all[gid], phase, ds);
    if (x == 0) dec[(size_t)f * g.N + col] = (uint8_t)ds[0];
}

// row syndromes from dec (thread per (frame,row)); sets flag[f] when any row is unsatisfied
__global__ void k_syndrome(DevGraph g, int B, int early_stop, const uint8_t *__restrict__ done, const uint8_t *__restrict__ dec, int *flag) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (long long)B * g.M) return;
    const int f = (int)(w / g.M), row = (int)(w % g.M);
    if (early_stop && done[f]) return;
    int syn = 0;
    for (int e = g.chk_ptr[row]; e < g.chk_ptr[row + 1]; ++e) syn ^= g.pfwd[(size_t)e * g.q + dec[(size_t)f * g.N + g.chk_var[e]]];
    if (syn) flag[f] = 1;
}

// per frame: latch the result at the first zero syndrome (or at max_iter), as the reference returns it.  One warp per frame.
// active[iter - 1] += frames that were still decoding when iteration `iter` ran (one atomic per CTA): the early-stop statistics
__global__ void k_finalize(int B, int N, int iter, int max_iter, int *flag, uint8_t *done, int *iters, int *conv,
                           const uint8_t *__restrict__ dec, int *out_sym, uint8_t *newly, int *active) {
    // one THREAD per frame for the latch test (a warp per frame made 37 888 tiny CTAs per iteration of the GF(16) configuration);
    // the frames of a warp that latch in this iteration -- each frame does so once per decode -- are then copied by the whole
    // warp, one after the other, coalesced
    const int f = (int)(blockIdx.x * (unsigned)blockDim.x + threadIdx.x), lane = threadIdx.x & 31;
    const bool act = f < B && !done[f];
    const int was_active = __syncthreads_count(act);
    if (threadIdx.x == 0 && was_active) atomicAdd(active + iter - 1, was_active);
    int l = 0;
    if (act) {
        const bool ok = flag[f] == 0;
        flag[f] = 0;
        if (ok || iter == max_iter) { done[f] = 1; conv[f] = ok ? 1 : 0; iters[f] = ok ? iter : max_iter; l = 1; }
    }
    if (f < B) newly[f] = (uint8_t)l;
    const int f0 = f - lane;                                        // first frame of my warp
    for (unsigned m = __ballot_sync(0xffffffffu, l != 0); m; m &= m - 1u) {
        const size_t base = (size_t)(f0 + __ffs((int)m) - 1) * N;
        for (int i = lane; i < N; i += 32) out_sym[base + i] = dec[base + i];
    }
}

// ---- soft demapper (CComm::Demodulate, Comm.cpp:340-407); fp64 in the reference's operation order -----
// ysym[s] = offset of symbol s's first sample in the frame's y array, or -1 when punctured
// One CTA handles `syms` consecutive (frame, symbol) pairs.  The p bit LLRs of a symbol, -2 y / sigma^2 (the reference's
// expression, one fp64 division each -- a long dependent sequence, hence many symbols per CTA), are formed once in shared
// memory; then thread <-> (symbol slot, x) adds the ones of x's set bits in ascending bit order, which is the reference's
// summation order (Comm.cpp:361-372).  B * N < 2^31 (host check).
template <typename T>
__global__ void k_demap_bpsk(int B, int N, int q, int p, int ylen, int syms, const double *__restrict__ y, double sigma,
                             const int *__restrict__ ysym, T *lch /*[B][N][q] or null*/, double *out /*[B][N][q-1] or null*/) {
    extern __shared__ double s_bit[];                               // [syms][p]
    const unsigned fs0 = blockIdx.x * (unsigned)syms;               // first (frame, symbol) of this CTA
    const unsigned nfs = (unsigned)B * (unsigned)N;
    for (unsigned i = threadIdx.x; i < (unsigned)(syms * p); i += blockDim.x) {
        const unsigned fs = fs0 + i / p, b = i % p;
        double v = 0;
        if (fs < nfs) {
            const unsigned s = fs % N, f = fs / N;
            const int off = ysym[s];
            if (off >= 0) v = -2 * y[(size_t)f * ylen + off + b] / (sigma * sigma);     // punctured symbols: all-zero LLRs
        }
        s_bit[i] = v;
    }
    __syncthreads();
    const unsigned slots = blockDim.x / q, slot = threadIdx.x / q, x = threadIdx.x % q;
    if (slot >= slots) return;
    for (unsigned ls = slot; ls < (unsigned)syms; ls += slots) {
        const size_t fs = fs0 + ls;
        if (fs >= nfs) break;
        double a = 0;
#pragma unroll 4
        for (int b = 0; b < p; ++b)
            if (x & (1u << b)) a += s_bit[ls * p + b];
        if (lch) lch[fs * q + x] = (T)a;
        if (out && x != 0) out[fs * (q - 1) + x - 1] = a;
    }
}
template <typename T>
__global__ void k_demap_qam(int B, int N, int q, int ylen, const double *__restrict__ y, const double *__restrict__ cs, double sigma,
                            const int *__restrict__ ysym, T *lch, double *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * N * q) return;
    const int x = (int)(i % q); const size_t fs = i / q; const int s = (int)(fs % N); const size_t f = fs / N;
    double a = 0;
    const int off = ysym[s];
    if (off >= 0 && x != 0) {
        const double yr = y[(f * ylen + off) * 2], yi = y[(f * ylen + off) * 2 + 1];
        a = ((2 * yr - cs[0] - cs[2 * x]) * (cs[2 * x] - cs[0]) + (2 * yi - cs[1] - cs[2 * x + 1]) * (cs[2 * x + 1] - cs[1])) /
            (2 * sigma * sigma);
    }
    if (lch) lch[i] = (T)a;
    if (out && x != 0) out[fs * (q - 1) + x - 1] = a;
}

// ======================================================================================================
// host side
// ======================================================================================================
struct nbl_decoder {
    // small batches (the reference's per-frame call shape, B = 1): the launch sequence of one decode is captured once per
    // (batch size, posterior wanted, arithmetic) and replayed as ONE graph launch -- 27 kernel launches cost more than the
    // kernels themselves there.  Dropped when the OSD set-up changes.
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int> graph_launches;
    // large batches with HOST samples (nbl_demap_decode_batch): the samples cross PCIe in chunks on a second stream, and chunk c is
    // demapped and run through iteration 1 while chunk c + 1 is still in flight (iterations 2.. run on the whole batch)
    struct Pipe { bool active = false; int modulation = 0, ylen = 0, chunks = 0, per = 0; double sigma = 0; const double *ysrc = nullptr; } pipe;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    const nbl_code *code = nullptr;
    nbl_decoder_config cfg{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    DevGraph g{};
    size_t tsz = 4;
    // device buffers
    int *d_chk_ptr = nullptr, *d_chk_var = nullptr, *d_var_ptr = nullptr, *d_var_edge = nullptr, *d_ysym = nullptr, *d_rowdesc = nullptr;
    uint8_t *d_pfwd = nullptr, *d_pinv = nullptr, *d_pfwd_sw = nullptr, *d_pinv_sw = nullptr;
    void *d_lch = nullptr, *d_v2c = nullptr, *d_c2v[2] = {nullptr, nullptr};
    uint8_t *d_vdec = nullptr, *d_dec = nullptr, *d_done = nullptr, *d_newly = nullptr;
    int *d_flag = nullptr, *d_iters = nullptr, *d_conv = nullptr, *d_out = nullptr, *d_active = nullptr;
    double *d_stage = nullptr;   size_t stage_bytes = 0;     // staging for host LLR inputs / outputs (doubles)
    double *d_ystage = nullptr;  size_t ystage_bytes = 0;    // staging for host received samples
    std::vector<int> punct_cache; int punct_mod = -1, punct_ylen = 0;
    uint8_t *d_pool = nullptr; size_t pool_cap = 0;          // codeword pool of the Monte-Carlo path
    uint8_t *d_tx = nullptr, *d_msg = nullptr;               // transmitted code words / messages of the running batch [max_batch][N], [max_batch][K]
    nbl_encoder *sim_enc = nullptr;                          // device encoder of the random-message Monte-Carlo mode
    std::vector<int> pool_host;
    unsigned long long *d_counters = nullptr;
    bool timing_pending = false;
    double *d_cons = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> it_ev;                          // [2*max_iter] start/stop around each k_iterate launch
    int it_recorded = 0;
    float last_ms = 0; int last_launches = 0;
    size_t smem_iter = 0;
    bool warp_path = false;      // T-EMS n_c <= 2 or BS-TEMS, dc <= 4, q >= 32: one warp per check row (tems_warp.cuh)
    bool qspa_warp_path = false; // QSPA, dc <= 4, q >= 32: one warp per check row (qspa_warp.cuh)
    bool ems_warp_path = false;  // EMS, n_m a power of two, n_m^(dc-1) <= EW_MAXCOMBOS, dc <= 4, q >= 32: one warp per check row (ems_warp.cuh)
    bool warp_ref = false;       // the T-EMS warp kernel records the configuration the reference's enumeration meets first (tie_mode REFERENCE)
    bool warp_regular = false;   // every check row has TW_DC edges and every variable TW_HOIST: degree guards compiled out
    bool ems_half_path = false;  // EMS over GF(16), dc <= 8, dv <= 4, n_m <= 16, n_c <= 3 (ems_half.cuh)
    bool tems_half_path = false; // T-EMS n_c <= 2 over GF(16), dc <= 8, dv <= 4, not the reference tie order (tems_half.cuh)
    bool qspa_half_path = false; // QSPA over GF(16), dc <= 8, dv <= 4: four lanes per check row, eight rows per warp (qspa_half.cuh)
    int *d_rowdesc8 = nullptr;   // its row descriptors [M][QH_MAXDC][QH_DSC]
    // OSD post-processing of failed frames (osd.cuh); off until nbl_decoder_enable_osd
    bool osd_on = false;
    nbl_osd_config osd_cfg{};
    OsdDev osd{};
    unsigned *d_osd_H0 = nullptr;
    double *d_osd_sum = nullptr;                             // LLR_Sym_sum, the entries OSD reads: [max_batch][N p]
    int *d_osd_status = nullptr;
    size_t osd_smem = 0;
};

extern "C" int nbl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename T> static int upload(T **dst, const std::vector<T> &src) {
    CK(cudaMalloc((void **)dst, std::max<size_t>(1, src.size()) * sizeof(T)));
    CK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return NBL_OK;
}

static size_t smem_for(const nbl_decoder *d, int Q) {
    const int G = Q >= 128 ? 1 : 128 / Q;
    const int mdc = d->code->max_dc;
    size_t per = 0;
    const bool f64 = d->cfg.precision == NBL_PREC_F64;
    const int tnc = d->cfg.tie_mode == NBL_TIES_BITEXACT && d->cfg.p1 < 3 ? 3 : d->cfg.p1;
#define PER(QQ)                                                                                              \
    if (Q == QQ) {                                                                                           \
        if (d->cfg.method == NBL_METHOD_TEMS) per = f64 ? TemsSmem<double, QQ>::bytes(mdc, tnc) : TemsSmem<float, QQ>::bytes(mdc, tnc); \
        else if (d->cfg.method == NBL_METHOD_BSTEMS) per = f64 ? TemsSmem<double, QQ>::bytes(mdc, 3) : TemsSmem<float, QQ>::bytes(mdc, 3); \
        else if (d->cfg.method == NBL_METHOD_EMS) per = f64 ? EmsSmem<double, QQ>::bytes(mdc, d->cfg.p0, d->cfg.p1) : EmsSmem<float, QQ>::bytes(mdc, d->cfg.p0, d->cfg.p1); \
        else per = f64 ? QspaSmem<double, QQ>::bytes(mdc) : QspaSmem<float, QQ>::bytes(mdc);              \
    }
    PER(16) PER(32) PER(64) PER(128) PER(256)
#undef PER
    return per * G;
}

extern "C" void nbl_decoder_destroy(nbl_decoder *d) {
    if (!d) return;
    cudaSetDevice(d->cfg.device);
    void *ptrs[] = {d->d_chk_ptr, d->d_chk_var, d->d_var_ptr, d->d_var_edge, d->d_ysym, d->d_rowdesc, d->d_pfwd, d->d_pinv, d->d_pfwd_sw, d->d_pinv_sw, d->d_lch, d->d_v2c,
                    d->d_c2v[0], d->d_c2v[1], d->d_vdec, d->d_dec, d->d_done, d->d_newly, d->d_flag, d->d_iters, d->d_conv,
                    d->d_out, d->d_active, d->d_stage, d->d_ystage, d->d_cons, d->d_pool, d->d_counters, d->d_tx, d->d_msg, d->d_osd_H0, d->d_osd_sum, d->d_osd_status, d->d_rowdesc8};
    if (d->sim_enc) nbl_encoder_destroy(d->sim_enc);
    for (auto &kv : d->graphs) cudaGraphExecDestroy(kv.second);
    for (cudaEvent_t e : d->chunk_ev) cudaEventDestroy(e);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    for (void *p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t e : d->it_ev) if (e) cudaEventDestroy(e);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->own_stream && d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

extern "C" int nbl_decoder_create(nbl_decoder **out, const nbl_code *code, const nbl_decoder_config *cfg, void *stream) {
    if (!out || !code || !cfg) { nbl_set_error("nbl_decoder_create: null argument"); return NBL_ERR_ARG; }
    if (cfg->method != NBL_METHOD_QSPA && cfg->method != NBL_METHOD_EMS && cfg->method != NBL_METHOD_TEMS && cfg->method != NBL_METHOD_BSTEMS) {
        nbl_set_error("nbl_decoder_create: decode method not supported (1 QSPA, 2 EMS, 4 T-EMS, 7 BS-TEMS)"); return NBL_ERR_ARG;
    }
    if (cfg->method == NBL_METHOD_BSTEMS && (cfg->p0 < 1 || cfg->p0 > 16 || cfg->p0 > code->q - 1 || cfg->p1 < 1 || code->max_dc > 16)) {
        nbl_set_error("BS-TEMS needs 1 <= basic_set_Nm <= min(q - 1, 16), basic_set_Nc >= 1 and check degree <= 16"); return NBL_ERR_ARG;
    }
    if (cfg->method == NBL_METHOD_TEMS && cfg->tie_mode == NBL_TIES_BITEXACT && code->max_dc > 16) {
        nbl_set_error("tie mode BITEXACT needs check degree <= 16"); return NBL_ERR_ARG;
    }
    if (code->q < 16 || code->q > 256) { nbl_set_error("nbl_decoder_create: the decode kernels cover 16 <= q <= 256 (GF(4) and GF(8) tables load, but no decoder is built for them)"); return NBL_ERR_ARG; }
    if (cfg->kernel < NBL_KERNEL_AUTO || cfg->kernel > NBL_KERNEL_WARP_GENERIC) { nbl_set_error("nbl_decoder_create: bad kernel selector"); return NBL_ERR_ARG; }
    if (cfg->max_iter < 1 || cfg->max_batch < 1) { nbl_set_error("nbl_decoder_create: max_iter and max_batch must be >= 1"); return NBL_ERR_ARG; }
    if (cfg->method == NBL_METHOD_EMS && (cfg->p0 < 1 || cfg->p0 > code->q || cfg->p1 < 1)) {
        nbl_set_error("EMS configuration error! EMS_Nm is too large!"); return NBL_ERR_ARG;   // NBLDPC.cpp:282-286
    }
    if (cfg->method == NBL_METHOD_TEMS && (cfg->p0 < 1 || cfg->p1 < 1)) { nbl_set_error("T-EMS needs n_r >= 1 and n_c >= 1"); return NBL_ERR_ARG; }
    if (cfg->method != NBL_METHOD_QSPA && cfg->factor == 0.0) { nbl_set_error("correction factor must be non-zero"); return NBL_ERR_ARG; }
    if (nbl_device_count() <= cfg->device || cfg->device < 0) { nbl_set_error("no such CUDA device (this library has no CPU fallback)"); return NBL_ERR_CUDA; }
    CK(cudaSetDevice(cfg->device));
    nbl_decoder *d = new nbl_decoder;
    d->code = code; d->cfg = *cfg;
    d->tsz = cfg->precision == NBL_PREC_F64 ? 8 : 4;
    int rc = NBL_OK;
    auto fail = [&](int r) { nbl_decoder_destroy(d); return r; };
    if (stream) d->stream = (cudaStream_t)stream;
    else { if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) { nbl_set_error("cudaStreamCreate failed"); return fail(NBL_ERR_CUDA); } d->own_stream = true; }
    const int q = code->q, E = code->E, N = code->N;
    // permutation tables of every edge coefficient
    std::vector<uint8_t> pf((size_t)E * q), pi((size_t)E * q);
    for (int e = 0; e < E; ++e) {
        const int h = code->chk_coef[e], hi = code->inv[h];
        for (int x = 0; x < q; ++x) { pf[(size_t)e * q + x] = (uint8_t)code->mul[h * q + x]; pi[(size_t)e * q + x] = (uint8_t)code->mul[hi * q + x]; }
    }
    if ((rc = upload(&d->d_chk_ptr, code->chk_ptr)) || (rc = upload(&d->d_chk_var, code->chk_var)) ||
        (rc = upload(&d->d_var_ptr, code->var_ptr)) || (rc = upload(&d->d_var_edge, code->var_edge)) ||
        (rc = upload(&d->d_pfwd, pf)) || (rc = upload(&d->d_pinv, pi)))
        return fail(rc);
    // row descriptors of the warp fast path: 32 ints per check row, TW_DSC per edge
    bool desc_ok = code->max_dc <= TW_DC && code->max_dv <= TW_DSC - 3;
    bool all_full = true;
    for (int r = 0; r < code->M; ++r) {
        const int deg = code->chk_ptr[r + 1] - code->chk_ptr[r];
        desc_ok = desc_ok && deg >= 2;
        all_full = all_full && deg == TW_DC;
    }
    for (int v = 0; v < code->N; ++v) all_full = all_full && code->var_ptr[v + 1] - code->var_ptr[v] == TW_HOIST;
    d->warp_regular = all_full && cfg->kernel != NBL_KERNEL_WARP_GENERIC;
    if (desc_ok) {
        std::vector<int> rd((size_t)code->M * 32, 0);
        for (int r = 0; r < code->M; ++r) {
            rd[(size_t)r * 32 + TW_DSC - 1] = code->chk_ptr[r];
            rd[(size_t)r * 32 + 2 * TW_DSC - 1] = code->chk_ptr[r + 1] - code->chk_ptr[r];
            for (int e = code->chk_ptr[r]; e < code->chk_ptr[r + 1]; ++e) {
                int *slot = &rd[(size_t)r * 32 + TW_DSC * (e - code->chk_ptr[r])];
                const int v = code->chk_var[e];
                slot[0] = v; slot[1] = code->var_ptr[v + 1] - code->var_ptr[v];
                for (int s = code->var_ptr[v]; s < code->var_ptr[v + 1]; ++s) slot[2 + s - code->var_ptr[v]] = code->var_edge[s];
            }
        }
        if ((rc = upload(&d->d_rowdesc, rd))) return fail(rc);
        if (q >= 32) {
            // table VALUES re-indexed to the split layout of the warp kernel's shared-memory vectors (PLayout in tems_warp.cuh)
            const int vec = q / 32, e16 = std::min<int>(16 / (int)d->tsz, vec);
            auto swz = [&](int x) { const int k = x / vec, j = x % vec; return (j / e16) * (32 * e16) + k * e16 + j % e16; };
            std::vector<uint8_t> pfs(pf.size()), pis(pi.size());
            for (size_t i = 0; i < pf.size(); ++i) { pfs[i] = (uint8_t)swz(pf[i]); pis[i] = (uint8_t)swz(pi[i]); }
            if ((rc = upload(&d->d_pfwd_sw, pfs)) || (rc = upload(&d->d_pinv_sw, pis))) return fail(rc);
        }
    }
    const bool fast_ok = cfg->kernel != NBL_KERNEL_GROUP;           // NBL_KERNEL_GROUP: every configuration on the generic kernel
    if (q == 16 && code->max_dc <= QH_MAXDC && code->max_dv <= QH_MAXDV && N < (1 << 15) && fast_ok) {      // the GF(16) kernels pack the variable index into 15 bits
        std::vector<int> rd((size_t)code->M * QH_MAXDC * QH_DSC, 0);
        for (int r = 0; r < code->M; ++r)
            for (int e = code->chk_ptr[r]; e < code->chk_ptr[r + 1]; ++e) {
                int *slot = &rd[((size_t)r * QH_MAXDC + (e - code->chk_ptr[r])) * QH_DSC];
                const int v = code->chk_var[e];
                slot[0] = v; slot[1] = code->var_ptr[v + 1] - code->var_ptr[v];
                for (int s = code->var_ptr[v]; s < code->var_ptr[v + 1]; ++s) slot[2 + s - code->var_ptr[v]] = code->var_edge[s];
            }
        if ((rc = upload(&d->d_rowdesc8, rd))) return fail(rc);
        d->qspa_half_path = cfg->method == NBL_METHOD_QSPA;
    }
    d->g = DevGraph{q, code->p, N, code->M, E, code->max_dv, code->max_dc, d->d_chk_ptr, d->d_chk_var, d->d_var_ptr, d->d_var_edge, d->d_pfwd, d->d_pinv, d->d_rowdesc};
    const size_t B = cfg->max_batch;
#define ALLOC(ptr, bytes) if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) { cudaGetLastError(); nbl_set_error("cudaMalloc failed for the decoder workspace"); return fail(NBL_ERR_ALLOC); }
    ALLOC(d->d_lch, B * N * q * d->tsz);
    ALLOC(d->d_v2c, B * E * q * d->tsz);
    ALLOC(d->d_c2v[0], B * E * q * d->tsz);
    ALLOC(d->d_c2v[1], B * E * q * d->tsz);
    ALLOC(d->d_vdec, B * E); ALLOC(d->d_dec, B * N); ALLOC(d->d_done, B); ALLOC(d->d_newly, B);
    ALLOC(d->d_flag, B * sizeof(int)); ALLOC(d->d_iters, B * sizeof(int)); ALLOC(d->d_conv, B * sizeof(int));
    ALLOC(d->d_out, B * N * sizeof(int));
    ALLOC(d->d_active, (size_t)cfg->max_iter * sizeof(int));
    ALLOC(d->d_ysym, N * sizeof(int));
    ALLOC(d->d_cons, 2 * q * sizeof(double));
#undef ALLOC
    if (cudaEventCreate(&d->ev0) != cudaSuccess || cudaEventCreate(&d->ev1) != cudaSuccess) { nbl_set_error("cudaEventCreate failed"); return fail(NBL_ERR_CUDA); }
    d->it_ev.assign(2 * (size_t)cfg->max_iter, nullptr);
    for (auto &e : d->it_ev) if (cudaEventCreate(&e) != cudaSuccess) { nbl_set_error("cudaEventCreate failed"); return fail(NBL_ERR_CUDA); }
    d->smem_iter = smem_for(d, q);
    if (cfg->tie_mode < NBL_TIES_AUTO || cfg->tie_mode > NBL_TIES_BITEXACT) { nbl_set_error("nbl_decoder_create: bad tie_mode"); return fail(NBL_ERR_ARG); }
    const bool ref_ties = cfg->tie_mode == NBL_TIES_REFERENCE || cfg->tie_mode == NBL_TIES_BITEXACT || (cfg->tie_mode == NBL_TIES_AUTO && cfg->precision == NBL_PREC_F64);
    d->warp_path = cfg->method == NBL_METHOD_TEMS && cfg->p1 <= 2 && desc_ok && q >= 32 && cfg->offset >= 0.0 && !ref_ties && fast_ok;
    // the reference's tie order on the same kernel (a second pass over the handled entries records the configuration the reference's
    // enumeration meets first, tems_warp.cuh); n_r >= 3 (a third column can tie) and the bit-exact replay stay on the group kernel
    if (cfg->method == NBL_METHOD_TEMS && cfg->p1 <= 2 && cfg->p0 <= 2 && desc_ok && q >= 32 && cfg->offset >= 0.0 && ref_ties &&
        cfg->tie_mode != NBL_TIES_BITEXACT && fast_ok) { d->warp_path = true; d->warp_ref = true; }
    d->tems_half_path = cfg->method == NBL_METHOD_TEMS && cfg->p1 <= 2 && q == 16 && d->d_rowdesc8 != nullptr && !ref_ties;
    if (cfg->method == NBL_METHOD_BSTEMS && q == 16 && d->d_rowdesc8 != nullptr) d->tems_half_path = true;      // the same kernel with BS = 1
    d->ems_half_path = cfg->method == NBL_METHOD_EMS && q == 16 && d->d_rowdesc8 != nullptr && cfg->p0 <= EH_MAXNM && cfg->p1 + 1 <= EH_MAXNJ;
    // BS-TEMS on the same kernel (its search replaced by the basic-set walk): no tie modes, the walk is the reference's own
    if (cfg->method == NBL_METHOD_BSTEMS && desc_ok && q >= 32 && cfg->offset >= 0.0 && fast_ok) d->warp_path = true;
    {
        const bool pow2 = cfg->p0 >= 1 && (cfg->p0 & (cfg->p0 - 1)) == 0;
        long long combos = 1;
        for (int i = 0; i + 1 < code->max_dc; ++i) combos *= cfg->p0;
        d->qspa_warp_path = cfg->method == NBL_METHOD_QSPA && desc_ok && q >= 32 && fast_ok;
        d->ems_warp_path = cfg->method == NBL_METHOD_EMS && desc_ok && q >= 32 && pow2 && cfg->p0 <= EW_MAXNM && cfg->p0 <= q &&
                           combos <= EW_MAXCOMBOS && fast_ok;
    }
    if (d->smem_iter > 227 * 1024) { nbl_set_error("decoder configuration needs more shared memory than one SM has"); return fail(NBL_ERR_ARG); }
    *out = d;
    return NBL_OK;
}

extern "C" const char *nbl_decoder_kernel(const nbl_decoder *d) {
    if (!d) return "";
    if (d->tems_half_path) return d->cfg.method == NBL_METHOD_BSTEMS ? "bstems_gf16" : "tems_gf16";
    if (d->ems_half_path) return "ems_gf16";
    if (d->qspa_half_path) return "qspa_gf16";
    if (d->warp_path) return d->cfg.method == NBL_METHOD_BSTEMS ? "bstems_warp" : "tems_warp";
    if (d->ems_warp_path) return "ems_warp";
    if (d->qspa_warp_path) return "qspa_warp";
    return "group";
}

#if TW_STATS
extern "C" int nbl_dev_stats(unsigned long long *out, int reset) {      // development builds only
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, nbl::tw_stats, sizeof(unsigned long long) * 128));
    if (reset) { unsigned long long z[128] = {0}; CK(cudaMemcpyToSymbol(nbl::tw_stats, z, sizeof z)); }
    return 0;
}
#endif
extern "C" void *nbl_decoder_stream(nbl_decoder *d) { return d ? (void *)d->stream : nullptr; }
extern "C" int nbl_decoder_sync(nbl_decoder *d) {
    if (!d) { nbl_set_error("null decoder"); return NBL_ERR_ARG; }
    CK(cudaSetDevice(d->cfg.device));
    CK(cudaStreamSynchronize(d->stream));
    return NBL_OK;
}
extern "C" int nbl_decoder_last_timing(nbl_decoder *d, float *ms, int *launches) {
    if (!d) { nbl_set_error("null decoder"); return NBL_ERR_ARG; }
    if (d->timing_pending) {
        CK(cudaSetDevice(d->cfg.device));
        CK(cudaEventSynchronize(d->ev1));
        CK(cudaEventElapsedTime(&d->last_ms, d->ev0, d->ev1));
        d->timing_pending = false;
    }
    if (ms) *ms = d->last_ms;
    if (launches) *launches = d->last_launches;
    return NBL_OK;
}

// duration of each check-node iteration kernel of the last decode (ms[i] for iteration i+1); returns the count
extern "C" int nbl_decoder_iter_timing(nbl_decoder *d, float *ms, int capacity) {
    if (!d) { nbl_set_error("null decoder"); return NBL_ERR_ARG; }
    CK(cudaSetDevice(d->cfg.device));
    CK(cudaStreamSynchronize(d->stream));
    int n = d->it_recorded < capacity ? d->it_recorded : capacity;
    for (int i = 0; i < n; ++i) CK(cudaEventElapsedTime(&ms[i], d->it_ev[2 * i], d->it_ev[2 * i + 1]));
    return n;
}

// frames still decoding when each iteration of the last decode ran: active[i] for iteration i + 1 (max_iter entries)
extern "C" int nbl_decoder_iter_active(nbl_decoder *d, int *active, int capacity) {
    if (!d || !active) { nbl_set_error("nbl_decoder_iter_active: null argument"); return NBL_ERR_ARG; }
    CK(cudaSetDevice(d->cfg.device));
    CK(cudaStreamSynchronize(d->stream));
    const int n = d->cfg.max_iter < capacity ? d->cfg.max_iter : capacity;
    CK(cudaMemcpy(active, d->d_active, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    return n;
}

static int ensure_buf(double **buf, size_t *have, size_t bytes) {
    if (*have >= bytes) return NBL_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *have = 0;
    if (cudaMalloc((void **)buf, bytes) != cudaSuccess) { cudaGetLastError(); nbl_set_error("cudaMalloc failed for a staging buffer"); return NBL_ERR_ALLOC; }
    *have = bytes;
    return NBL_OK;
}
static int ensure_stage(nbl_decoder *d, size_t bytes) { return ensure_buf(&d->d_stage, &d->stage_bytes, bytes); }
static int ensure_ystage(nbl_decoder *d, size_t bytes) { return ensure_buf(&d->d_ystage, &d->ystage_bytes, bytes); }

// ---- launch helpers (dispatch on precision, field size, method) ------------------------------------
template <typename T, int Q, int METHOD> static int launch_iterate_tqm(nbl_decoder *d, const IterParams<T> &P) {
    auto kern = k_iterate<T, Q, METHOD>;
    if (d->smem_iter > 48 * 1024)
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_iter));
    const long long groups = (long long)P.B * P.g.M;
    const unsigned grid = (unsigned)((groups + Grp<Q>::G - 1) / Grp<Q>::G);
    kern<<<grid, Grp<Q>::THREADS, d->smem_iter, d->stream>>>(P);
    CK(cudaGetLastError());
    return NBL_OK;
}
template <typename T, int Q, int DC, int DV, int BS = 0> static int launch_iterate_warp_dc(nbl_decoder *d, const IterParams<T> &P) {
    auto kern = k_iterate_tems_warp<T, Q, DC, DV, BS>;
    constexpr size_t smem = TwSmem<T, Q>::bytes;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // grid: x = group of TW_WARPS check rows (fastest, so that the rows of a frame -- which share messages -- run together),
    // y, z = frame
    const dim3 grid((unsigned)((P.g.M + TW_WARPS - 1) / TW_WARPS), (unsigned)std::min(P.B, 65535), (unsigned)((P.B + 65534) / 65535));
    kern<<<grid, 32 * TW_WARPS, smem, d->stream>>>(P);
    CK(cudaGetLastError());
    return NBL_OK;
}
template <typename T, int Q> static int launch_iterate_warp(nbl_decoder *d, const IterParams<T> &P) {
    if constexpr (Q >= 32) {
        if (d->cfg.method == NBL_METHOD_BSTEMS)
            return d->warp_regular ? launch_iterate_warp_dc<T, Q, TW_DC, TW_HOIST, 1>(d, P) : launch_iterate_warp_dc<T, Q, 0, 0, 1>(d, P);
        if (d->warp_ref)
            return d->warp_regular ? launch_iterate_warp_dc<T, Q, TW_DC, TW_HOIST, 2>(d, P) : launch_iterate_warp_dc<T, Q, 0, 0, 2>(d, P);
        return d->warp_regular ? launch_iterate_warp_dc<T, Q, TW_DC, TW_HOIST>(d, P) : launch_iterate_warp_dc<T, Q, 0, 0>(d, P);
    } else {
        nbl_set_error("internal: warp path needs q >= 32");
        return NBL_ERR_ARG;
    }
}
template <typename T, int Q, int DC, int DV> static int launch_iterate_ems_warp_dc(nbl_decoder *d, const IterParams<T> &P) {
    auto kern = k_iterate_ems_warp<T, Q, DC, DV>;
    constexpr size_t smem = EwSmem<T, Q>::bytes;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)((P.g.M + TW_WARPS - 1) / TW_WARPS), (unsigned)std::min(P.B, 65535), (unsigned)((P.B + 65534) / 65535));
    kern<<<grid, 32 * TW_WARPS, smem, d->stream>>>(P);
    CK(cudaGetLastError());
    return NBL_OK;
}
template <typename T, int Q> static int launch_iterate_ems_warp(nbl_decoder *d, const IterParams<T> &P) {
    if constexpr (Q >= 32) {
        return d->warp_regular ? launch_iterate_ems_warp_dc<T, Q, TW_DC, TW_HOIST>(d, P) : launch_iterate_ems_warp_dc<T, Q, 0, 0>(d, P);
    } else {
        nbl_set_error("internal: warp path needs q >= 32");
        return NBL_ERR_ARG;
    }
}
template <typename T, int Q, int DC, int DV> static int launch_iterate_qspa_warp_dc(nbl_decoder *d, const IterParams<T> &P) {
    auto kern = k_iterate_qspa_warp<T, Q, DC, DV>;
    constexpr size_t smem = QwSmem<T, Q>::bytes;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)((P.g.M + TW_WARPS - 1) / TW_WARPS), (unsigned)std::min(P.B, 65535), (unsigned)((P.B + 65534) / 65535));
    kern<<<grid, 32 * TW_WARPS, smem, d->stream>>>(P);
    CK(cudaGetLastError());
    return NBL_OK;
}
template <typename T, int Q> static int launch_iterate_qspa_warp(nbl_decoder *d, const IterParams<T> &P) {
    if constexpr (Q >= 32) {
        return d->warp_regular ? launch_iterate_qspa_warp_dc<T, Q, TW_DC, TW_HOIST>(d, P) : launch_iterate_qspa_warp_dc<T, Q, 0, 0>(d, P);
    } else {
        nbl_set_error("internal: warp path needs q >= 32");
        return NBL_ERR_ARG;
    }
}
template <typename T, int Q> static int launch_iterate_tq(nbl_decoder *d, const IterParams<T> &P) {
    if constexpr (Q == 16) {
        if (d->tems_half_path) {
            const long long rows = (long long)P.B * P.g.M, per_cta = QH_ROWS * QH_WARPS;
            if (d->cfg.method == NBL_METHOD_BSTEMS) {
                CK(cudaFuncSetAttribute(k_iterate_tems_half<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ThSmem<T>::bytes));
                k_iterate_tems_half<T, 1><<<(unsigned)((rows + per_cta - 1) / per_cta), 32 * QH_WARPS, ThSmem<T>::bytes, d->stream>>>(P, d->d_rowdesc8);
            } else {
                CK(cudaFuncSetAttribute(k_iterate_tems_half<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ThSmem<T>::bytes));
                k_iterate_tems_half<T, 0><<<(unsigned)((rows + per_cta - 1) / per_cta), 32 * QH_WARPS, ThSmem<T>::bytes, d->stream>>>(P, d->d_rowdesc8);
            }
            CK(cudaGetLastError());
            return NBL_OK;
        }
        if (d->ems_half_path) {
            const long long rows = (long long)P.B * P.g.M, per_cta = QH_ROWS * QH_WARPS;
            CK(cudaFuncSetAttribute(k_iterate_ems_half<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EhSmem<T>::bytes));
            k_iterate_ems_half<T><<<(unsigned)((rows + per_cta - 1) / per_cta), 32 * QH_WARPS, EhSmem<T>::bytes, d->stream>>>(P, d->d_rowdesc8);
            CK(cudaGetLastError());
            return NBL_OK;
        }
        if (d->qspa_half_path) {
            const long long rows = (long long)P.B * P.g.M, per_cta = QH_ROWS * QH_WARPS;
            const size_t smem = QhQspaSmem<T>::bytes(P.g.max_dc);
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_iterate_qspa_half<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_iterate_qspa_half<T><<<(unsigned)((rows + per_cta - 1) / per_cta), 32 * QH_WARPS, smem, d->stream>>>(P, d->d_rowdesc8);
            CK(cudaGetLastError());
            return NBL_OK;
        }
    }
    if (d->warp_path) return launch_iterate_warp<T, Q>(d, P);
    if (d->ems_warp_path) return launch_iterate_ems_warp<T, Q>(d, P);
    if (d->qspa_warp_path) return launch_iterate_qspa_warp<T, Q>(d, P);
    switch (d->cfg.method) {
        case NBL_METHOD_TEMS: return launch_iterate_tqm<T, Q, NBL_METHOD_TEMS>(d, P);
        case NBL_METHOD_BSTEMS: return launch_iterate_tqm<T, Q, NBL_METHOD_BSTEMS>(d, P);
        case NBL_METHOD_EMS: return launch_iterate_tqm<T, Q, NBL_METHOD_EMS>(d, P);
        default: return launch_iterate_tqm<T, Q, NBL_METHOD_QSPA>(d, P);
    }
}
// The decision-only form of k_posterior for q >= 32 (the last iteration of every decode): ONE WARP per (frame, variable), VEC = Q / 32
// consecutive elements per lane with 128-bit loads, the same sums in the same order, the decision by warp_decide (tems_warp.cuh) --
// no shared memory, no block barrier; the group form reaches 45 % of the copy bandwidth, this one is bound by it.
template <typename T, int Q>
__global__ void __launch_bounds__(256) k_posterior_warp(DevGraph g, int B, int first, int early_stop, const T *__restrict__ lch,
                                                      const T *__restrict__ c2v_in, const uint8_t *__restrict__ done, uint8_t *dec) {
    constexpr int VEC = Q / 32;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)B * g.N) return;
    const int f = (int)(w / g.N), col = (int)(w % g.N);
    if (early_stop && done[f]) return;
    T post[VEC];
    nbl::ld_vec<T, VEC>(lch + ((size_t)f * g.N + col) * Q + VEC * lane, post);
    if (!first)
        for (int s = g.var_ptr[col]; s < g.var_ptr[col + 1]; ++s) {
            T m[VEC];
            nbl::ld_vec<T, VEC>(c2v_in + ((size_t)f * g.E + g.var_edge[s]) * Q + VEC * lane, m);
#pragma unroll
            for (int i = 0; i < VEC; ++i) post[i] = post[i] + m[i];
        }
    const int dsym = nbl::warp_decide<T, VEC>(post, lane);
    if (lane == 0) dec[(size_t)f * g.N + col] = (uint8_t)dsym;
}

// The same over GF(16): four lanes per (frame, variable) with four elements each (the layout of qspa_half.cuh), eight variables per warp
template <typename T>
__global__ void __launch_bounds__(256) k_posterior_q16(DevGraph g, int B, int first, int early_stop, const T *__restrict__ lch,
                                                     const T *__restrict__ c2v_in, const uint8_t *__restrict__ done, uint8_t *dec) {
    const int l4 = threadIdx.x & 3;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = w < (long long)B * g.N;
    const int f = live ? (int)(w / g.N) : 0, col = live ? (int)(w % g.N) : 0;
    const bool act = live && !(early_stop && done[f]);
    T post[4] = {T(0), T(0), T(0), T(0)};
    if (act) {
        nbl::qh_load4<T>(lch + ((size_t)f * g.N + col) * 16 + 4 * l4, post);
        if (!first)
            for (int s = g.var_ptr[col]; s < g.var_ptr[col + 1]; ++s) {
                T m[4];
                nbl::qh_load4<T>(c2v_in + ((size_t)f * g.E + g.var_edge[s]) * 16 + 4 * l4, m);
#pragma unroll
                for (int i = 0; i < 4; ++i) post[i] = post[i] + m[i];
            }
    }
    const int dsym = nbl::qh_decide<T>(post, 4 * l4);               // every lane takes part in the shuffles
    if (act && l4 == 0) dec[(size_t)f * g.N + col] = (uint8_t)dsym;
}

template <typename T, int Q> static int launch_posterior_tq(nbl_decoder *d, int B, int first, const T *c2v_in, double *post_out, const uint8_t *sel) {
    const long long groups = (long long)B * d->g.N;
    if constexpr (Q == 16) {
        if (!post_out && d->cfg.kernel != NBL_KERNEL_GROUP) {
            k_posterior_q16<T><<<(unsigned)((groups * 4 + 255) / 256), 256, 0, d->stream>>>(d->g, B, first, d->cfg.early_stop, (const T *)d->d_lch, c2v_in, d->d_done, d->d_dec);
            CK(cudaGetLastError());
            return NBL_OK;
        }
    }
    if constexpr (Q >= 32) {
        if (!post_out && d->cfg.kernel != NBL_KERNEL_GROUP) {
            k_posterior_warp<T, Q><<<(unsigned)((groups + 7) / 8), 256, 0, d->stream>>>(d->g, B, first, d->cfg.early_stop, (const T *)d->d_lch, c2v_in, d->d_done, d->d_dec);
            CK(cudaGetLastError());
            return NBL_OK;
        }
    }
    const unsigned grid = (unsigned)((groups + Grp<Q>::G - 1) / Grp<Q>::G);
    k_posterior<T, Q><<<grid, Grp<Q>::THREADS, 0, d->stream>>>(d->g, B, first, d->cfg.early_stop, (const T *)d->d_lch, c2v_in, d->d_done,
                                                               d->d_dec, post_out, sel);
    CK(cudaGetLastError());
    return NBL_OK;
}
#ifdef NBL_DEV_MINIMAL
// development builds (tools/build_variant.sh -DNBL_DEV_MINIMAL): only the GF(256) fp32 instantiations, 10 x faster to compile.
// Never used for the shipped library.
#define DISPATCH_Q(q, CALL)                                                                           \
    switch (q) {                                                                                      \
        case 256: { constexpr int QQ = 256; CALL; } break;                                            \
        default: nbl_set_error("development build: GF(256) only"); return NBL_ERR_ARG;                \
    }
#define run_decode_f64(d, B, post) (nbl_set_error("development build: fp32 only"), NBL_ERR_ARG)
#else
#define DISPATCH_Q(q, CALL)                                                                           \
    switch (q) {                                                                                      \
        case 16: { constexpr int QQ = 16; CALL; } break;                                              \
        case 32: { constexpr int QQ = 32; CALL; } break;                                              \
        case 64: { constexpr int QQ = 64; CALL; } break;                                              \
        case 128: { constexpr int QQ = 128; CALL; } break;                                            \
        case 256: { constexpr int QQ = 256; CALL; } break;                                            \
        default: nbl_set_error("unsupported field size"); return NBL_ERR_ARG;                         \
    }
#define run_decode_f64(d, B, post) run_decode<double>(d, B, post)
#endif

// Runs the iterations on the frames whose channel LLRs are already in d_lch.  Results stay on the device
// (d_out, d_conv, d_iters); post_dev (device doubles) is optional.
static int launch_demap(nbl_decoder *d, int modulation, const double *ysrc, int f0, int nf, int ylen, double sigma, bool to_lch, double *out_dev);
template <typename T> static int enqueue_decode(nbl_decoder *d, int B, double *post_dev, bool timed);
constexpr int NBL_GRAPH_MAX_B = 64;       // batches up to this size go through a captured graph
constexpr int NBL_PIPE_MIN_B = 512;       // host-sample batches from this size on are copied in chunks behind which iteration 1 starts
constexpr int NBL_PIPE_CHUNKS = 4;

static void drop_graphs(nbl_decoder *d) {
    for (auto &kv : d->graphs) cudaGraphExecDestroy(kv.second);
    d->graphs.clear(); d->graph_launches.clear();
}

template <typename T> static int run_decode(nbl_decoder *d, int B, double *post_dev) {
    if (B > NBL_GRAPH_MAX_B) return enqueue_decode<T>(d, B, post_dev, true);
    const int key = (B << 2) | (post_dev ? 1 : 0) | (sizeof(T) == 8 ? 2 : 0);
    auto it = d->graphs.find(key);
    if (it == d->graphs.end()) {
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_decode<T>(d, B, post_dev, false);
        const cudaError_t e = cudaStreamEndCapture(d->stream, &graph);
        if (rc != NBL_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
        if (e != cudaSuccess || !graph) { nbl_set_error(std::string("stream capture of the decode failed: ") + cudaGetErrorString(e)); cudaGetLastError(); return NBL_ERR_CUDA; }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { nbl_set_error(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ei)); return NBL_ERR_CUDA; }
        it = d->graphs.emplace(key, exec).first;
        d->graph_launches[key] = d->last_launches;
    }
    d->last_launches = d->graph_launches[key]; d->it_recorded = 0;     // kernels inside the graph; per-iteration events are not recorded
    CK(cudaEventRecord(d->ev0, d->stream));
    CK(cudaGraphLaunch(it->second, d->stream));
    CK(cudaEventRecord(d->ev1, d->stream));
    d->timing_pending = true;
    return NBL_OK;
}

template <typename T> static int enqueue_decode(nbl_decoder *d, int B, double *post_dev, bool timed) {
    const nbl_decoder_config &c = d->cfg;
    const int q = d->g.q;
    d->last_launches = 0; d->it_recorded = 0;
    CK(cudaMemsetAsync(d->d_flag, 0, (size_t)B * sizeof(int), d->stream));
    CK(cudaMemsetAsync(d->d_done, 0, (size_t)B, d->stream));
    CK(cudaMemsetAsync(d->d_active, 0, (size_t)c.max_iter * sizeof(int), d->stream));
    const bool osd_acc = d->osd_on && d->osd.order >= 0 && !d->osd.flag;
    if (osd_acc) CK(cudaMemsetAsync(d->d_osd_sum, 0, (size_t)B * d->osd.nbit * sizeof(double), d->stream));
    if (timed) CK(cudaEventRecord(d->ev0, d->stream));
    T w_old = 0, w_new = 1;
    if (c.method == NBL_METHOD_QSPA) { w_old = (T)0.5; w_new = (T)0.5; }          // NBLDPC.cpp:739
    if (c.method == NBL_METHOD_TEMS || c.method == NBL_METHOD_BSTEMS) { w_old = (T)0.25; w_new = (T)0.75; }   // NBLDPC.cpp:1046, 1262
    for (int it = 1; it <= c.max_iter; ++it) {
        const T *cin = (const T *)d->d_c2v[(it - 1) & 1];
        T *cout = (T *)d->d_c2v[it & 1];
        int rc;
        if (osd_acc) {                                              // LLR_Sym_sum += this iteration's posterior, frames still decoding
            const long long n = (long long)B * d->osd.nbit;
            k_osd_accumulate<T><<<(unsigned)((n + 255) / 256), 256, 0, d->stream>>>(d->g, B, it == 1, d->osd.alpha, d->d_done, (const T *)d->d_lch, cin, d->d_osd_sum);
            CK(cudaGetLastError());
            d->last_launches += 1;
        }
        if (it < c.max_iter) {
            IterParams<T> P;
            P.g = d->g; P.B = B; P.iter = it; P.first = it == 1; P.early_stop = c.early_stop; P.method = c.method;
            P.p0 = c.p0; P.p1 = c.p1; P.bitexact = c.tie_mode == NBL_TIES_BITEXACT; P.factor = (T)c.factor; P.offset = (T)c.offset; P.w_old = w_old; P.w_new = w_new;
            P.lch = (const T *)d->d_lch; P.v2c = (T *)d->d_v2c; P.c2v_in = cin; P.c2v_out = cout;
            P.vdec = d->d_vdec; P.dec = d->d_dec; P.flag = d->d_flag; P.done = d->d_done;
            P.pfwd_sw = d->d_pfwd_sw; P.pinv_sw = d->d_pinv_sw;
            if (timed) CK(cudaEventRecord(d->it_ev[2 * (it - 1)], d->stream));
            if (it == 1 && d->pipe.active) {
                // host samples arriving in chunks (nbl_demap_decode_batch): chunk c is demapped and iterated as soon as its copy has
                // landed; every per-frame array of the launch is offset to the chunk's first frame
                const int K = d->pipe.chunks, per = d->pipe.per;
                const size_t N = (size_t)d->g.N, E = (size_t)d->g.E;
                for (int ck = 0; ck < K; ++ck) {
                    const int f0 = ck * per, nf = std::min(per, B - f0);
                    if (nf <= 0) break;
                    CK(cudaStreamWaitEvent(d->stream, d->chunk_ev[ck], 0));
                    if ((rc = launch_demap(d, d->pipe.modulation, d->pipe.ysrc, f0, nf, d->pipe.ylen, d->pipe.sigma, true, nullptr))) return rc;
                    IterParams<T> Pc = P;
                    Pc.B = nf; Pc.lch = P.lch + f0 * N * q; Pc.v2c = P.v2c + f0 * E * q; Pc.c2v_in = P.c2v_in + f0 * E * q; Pc.c2v_out = P.c2v_out + f0 * E * q;
                    Pc.vdec = P.vdec + f0 * E; Pc.dec = P.dec + f0 * N; Pc.flag = P.flag + f0; Pc.done = P.done + f0;
                    DISPATCH_Q(q, rc = (launch_iterate_tq<T, QQ>(d, Pc)));
                    if (rc) return rc;
                    d->last_launches += 2;                           // demapper + iteration kernel of the chunk
                }
                rc = NBL_OK; d->last_launches -= 1;                 // (the common count below adds one iteration launch)
            } else
            DISPATCH_Q(q, rc = (launch_iterate_tq<T, QQ>(d, P)));
            if (rc) return rc;
            if (timed) { CK(cudaEventRecord(d->it_ev[2 * (it - 1) + 1], d->stream)); d->it_recorded = it; }
            d->last_launches += 1;
        } else {      // last iteration: the reference's final check-node pass cannot change any output
            DISPATCH_Q(q, rc = (launch_posterior_tq<T, QQ>(d, B, it == 1, cin, nullptr, nullptr)));
            if (rc) return rc;
            const long long rows = (long long)B * d->g.M;
            k_syndrome<<<(unsigned)((rows + 255) / 256), 256, 0, d->stream>>>(d->g, B, c.early_stop, d->d_done, d->d_dec, d->d_flag);
            CK(cudaGetLastError());
            d->last_launches += 2;
        }
        k_finalize<<<(unsigned)((B + 255) / 256), 256, 0, d->stream>>>(B, d->g.N, it, c.max_iter, d->d_flag, d->d_done, d->d_iters, d->d_conv, d->d_dec, d->d_out, d->d_newly, d->d_active);
        CK(cudaGetLastError());
        d->last_launches += 1;
        if (post_dev) {
            DISPATCH_Q(q, rc = (launch_posterior_tq<T, QQ>(d, B, it == 1, cin, post_dev, d->d_newly)));
            if (rc) return rc;
            d->last_launches += 1;
        }
    }
    if (d->osd_on && d->osd.order >= 0) {                           // NBLDPC.cpp:770-775, 921-925, 1135-1139: failed frames only
        k_osd<T><<<B, 256, d->osd_smem, d->stream>>>(d->osd, q, d->d_conv, (const T *)d->d_lch, d->d_osd_sum, d->d_out, d->d_osd_status);
        CK(cudaGetLastError());
        d->last_launches += 1;
    }
    if (timed) { CK(cudaEventRecord(d->ev1, d->stream)); d->timing_pending = true; }
    return NBL_OK;
}

// ---- OSD set-up (NBLDPC.cpp:380-457, OSD.h:440-510) --------------------------------------------------------------------
// CRC parity rows as the reference derives them: the (kbit - crc_len) x kbit band matrix of the CRC generator polynomial
// is brought to systematic form by G_GaussEliminate_bit (rows are added, never swapped; OSD.h:257-313); row i of the
// result says which message bits CRC bit (kbit - crc_rate + i) is the parity of.
static std::vector<std::vector<uint8_t>> osd_crc_rows(int kbit, int crc_len, int crc_rate) {
    const int gr = kbit - crc_len;
    std::vector<std::vector<uint8_t>> G(gr, std::vector<uint8_t>(kbit, 0));
    static const int t8[] = {0, 1, 4, 5, 7, 8}, t16[] = {0, 4, 11, 16}, t24[] = {0, 1, 18, 19, 23, 24};
    const int *taps = crc_len == 8 ? t8 : crc_len == 16 ? t16 : t24;
    const int ntap = crc_len == 16 ? 4 : 6;
    for (int i = 0; i < gr; ++i) for (int t = 0; t < ntap; ++t) G[i][i + taps[t]] = 1;
    // the band matrix has ones on its diagonal, so every pivot is in place: forward and backward elimination only
    for (int r = 0; r < gr; ++r)
        for (int up = r + 1; up < gr; ++up)
            if (G[up][r]) for (int i = 0; i < kbit; ++i) G[up][i] ^= G[r][i];
    for (int r = gr - 1; r > 0; --r)
        for (int up = r - 1; up >= 0; --up)
            if (G[up][r]) for (int i = 0; i < kbit; ++i) G[up][i] ^= G[r][i];
    std::vector<std::vector<uint8_t>> rows(crc_rate, std::vector<uint8_t>(kbit, 0));
    for (int i = 0; i < crc_rate; ++i) {
        rows[i][i + kbit - crc_rate] = 1;
        for (int j = 0; j < gr; ++j) rows[i][j] = G[j][kbit - crc_rate + i];
    }
    return rows;
}

extern "C" int nbl_decoder_enable_osd(nbl_decoder *d, const nbl_osd_config *cfg) {
    if (!d || !cfg) { nbl_set_error("nbl_decoder_enable_osd: null argument"); return NBL_ERR_ARG; }
    drop_graphs(d);                                                 // the captured launch sequences depend on the OSD set-up
    if (cfg->order < 0) { d->osd_on = false; return NBL_OK; }
    if (cfg->order > 3) { nbl_set_error("OSD order must be 0..3"); return NBL_ERR_ARG; }
    if (cfg->crc_len != 8 && cfg->crc_len != 16 && cfg->crc_len != 24) {
        nbl_set_error("OSD needs crcLen 8, 16 or 24 (the reference's CRC matrix set-up does not terminate otherwise, OSD.h:472-499)"); return NBL_ERR_ARG;
    }
    const nbl_code *c = d->code;
    const int P = c->p, nbit = c->N * P, mbit = c->M * P, kbit = nbit - mbit, R = cfg->crc_correct_len + mbit, W = (nbit + 31) / 32;
    if (cfg->crc_correct_len < 0 || cfg->crc_correct_len > cfg->crc_len || kbit <= cfg->crc_len) { nbl_set_error("OSD: bad crc_correct_len"); return NBL_ERR_ARG; }
    if (nbit > 2048) { nbl_set_error("OSD: more than 2048 code bits"); return NBL_ERR_ARG; }
    const size_t smem = OsdSmem::bytes(nbit, kbit, R, W);
    if (smem > 200 * 1024) { nbl_set_error("OSD: the bit matrices of this code do not fit into shared memory"); return NBL_ERR_ARG; }
    CK(cudaSetDevice(d->cfg.device));
    std::vector<unsigned> H0((size_t)R * W, 0u);
    auto setbit = [&](int r, int col) { H0[(size_t)r * W + (col >> 5)] |= 1u << (col & 31); };
    const auto crc = osd_crc_rows(kbit, cfg->crc_len, cfg->crc_correct_len);
    for (int i = 0; i < cfg->crc_correct_len; ++i) for (int j = 0; j < kbit; ++j) if (crc[i][j]) setbit(i, j);
    // binary image of H: block (check, variable) = matrix of the coefficient, entry [k][l] = bit k of g * x^l
    // (NBLDPC.cpp:416-430); the reference's table lacks the matrix of 1/alpha (GF.cpp:137)
    const int dead = cfg->fix_matrepr ? -1 : c->inv[2];
    for (int r = 0; r < c->M; ++r)
        for (int e = c->chk_ptr[r]; e < c->chk_ptr[r + 1]; ++e) {
            const int v = c->chk_var[e], g = c->chk_coef[e];
            if (g == dead) continue;
            for (int l = 0; l < P; ++l) {
                const int prod = c->mul[(size_t)g * c->q + (1 << l)];
                for (int k = 0; k < P; ++k) if ((prod >> k) & 1) setbit(cfg->crc_correct_len + P * r + k, P * v + l);
            }
        }
    if (d->d_osd_H0) { cudaFree(d->d_osd_H0); d->d_osd_H0 = nullptr; }
    if (d->d_osd_sum) { cudaFree(d->d_osd_sum); d->d_osd_sum = nullptr; }
    int rc;
    if ((rc = upload(&d->d_osd_H0, H0))) return rc;
    CK(cudaMalloc((void **)&d->d_osd_sum, (size_t)d->cfg.max_batch * nbit * sizeof(double)));
    if (!d->d_osd_status) { CK(cudaMalloc((void **)&d->d_osd_status, sizeof(int))); CK(cudaMemset(d->d_osd_status, 0, sizeof(int))); }
    CK(cudaFuncSetAttribute(k_osd<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_osd<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    d->osd_cfg = *cfg;
    d->osd.order = cfg->order; d->osd.flag = cfg->flag ? 1 : 0; d->osd.nbit = nbit; d->osd.kbit = kbit; d->osd.R = R; d->osd.W = W;
    d->osd.p = P; d->osd.N = c->N; d->osd.alpha = cfg->factor; d->osd.H0 = d->d_osd_H0;
    // compute_min_distance_bit's own bit length: multiplication first, then truncation (OSD.h:420) -- see osd.cuh
    d->osd.dist_len = (int)(c->N * std::log((double)c->q) / std::log(2.0));
    d->osd_smem = smem;
    d->osd_on = true;
    return NBL_OK;
}

// copies an int result array to the caller (host or device pointer)
static int give_ints(nbl_decoder *d, int *dst, const int *src_dev, size_t n) {
    if (!dst) return NBL_OK;
    CK(cudaMemcpyAsync(dst, src_dev, n * sizeof(int), is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, d->stream));
    return NBL_OK;
}

static int decode_common(nbl_decoder *d, int B, int *out_sym, int *converged, int *iterations, double *L_post) {
    const size_t nvec = (size_t)B * d->g.N, q1 = d->g.q - 1;
    double *post_dev = nullptr;
    int rc;
    if (L_post) {
        if (is_device_ptr(L_post)) post_dev = L_post;
        else { if ((rc = ensure_stage(d, nvec * q1 * sizeof(double)))) return rc; post_dev = d->d_stage; }
    }
    rc = d->cfg.precision == NBL_PREC_F64 ? run_decode_f64(d, B, post_dev) : run_decode<float>(d, B, post_dev);
    if (rc) return rc;
    if ((rc = give_ints(d, out_sym, d->d_out, nvec)) || (rc = give_ints(d, converged, d->d_conv, B)) || (rc = give_ints(d, iterations, d->d_iters, B))) return rc;
    if (L_post && post_dev != L_post) CK(cudaMemcpyAsync(L_post, post_dev, nvec * q1 * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    const bool any_host = (out_sym && !is_device_ptr(out_sym)) || (converged && !is_device_ptr(converged)) ||
                          (iterations && !is_device_ptr(iterations)) || (L_post && !is_device_ptr(L_post));
    if (any_host) CK(cudaStreamSynchronize(d->stream));
    return NBL_OK;
}

extern "C" int nbl_decode_batch(nbl_decoder *d, const double *L_ch, int B, int *out_sym, int *converged, int *iterations, double *L_post) {
    if (!d || !L_ch || !out_sym) { nbl_set_error("nbl_decode_batch: null argument"); return NBL_ERR_ARG; }
    if (B < 0 || B > d->cfg.max_batch) { nbl_set_error("nbl_decode_batch: B exceeds max_batch"); return NBL_ERR_ARG; }
    if (B == 0) return NBL_OK;
    CK(cudaSetDevice(d->cfg.device));
    const size_t nvec = (size_t)B * d->g.N, q = d->g.q;
    const double *src = L_ch;
    int rc;
    if (!is_device_ptr(L_ch)) {
        if ((rc = ensure_stage(d, nvec * (q - 1) * sizeof(double)))) return rc;
        CK(cudaMemcpyAsync(d->d_stage, L_ch, nvec * (q - 1) * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        src = d->d_stage;
    }
    const size_t tot = nvec * q;
    if (d->cfg.precision == NBL_PREC_F64) k_convert_lch<double><<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(src, (double *)d->d_lch, nvec, (int)q);
    else k_convert_lch<float><<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(src, (float *)d->d_lch, nvec, (int)q);
    CK(cudaGetLastError());
    return decode_common(d, B, out_sym, converged, iterations, L_post);
}

// Decoding_OSD_bit with flag 1 (decode method 6, NBLDPC.cpp:625-627): OSD alone on the channel LLRs, every frame
extern "C" int nbl_osd_batch(nbl_decoder *d, const double *L_ch, int B, int *out_sym) {
    if (!d || !L_ch || !out_sym) { nbl_set_error("nbl_osd_batch: null argument"); return NBL_ERR_ARG; }
    if (!d->osd_on) { nbl_set_error("nbl_osd_batch: call nbl_decoder_enable_osd first"); return NBL_ERR_ARG; }
    if (B < 0 || B > d->cfg.max_batch) { nbl_set_error("nbl_osd_batch: B exceeds max_batch"); return NBL_ERR_ARG; }
    if (B == 0) return NBL_OK;
    CK(cudaSetDevice(d->cfg.device));
    const size_t nvec = (size_t)B * d->g.N, q = d->g.q;
    const double *src = L_ch;
    int rc;
    if (!is_device_ptr(L_ch)) {
        if ((rc = ensure_stage(d, nvec * (q - 1) * sizeof(double)))) return rc;
        CK(cudaMemcpyAsync(d->d_stage, L_ch, nvec * (q - 1) * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        src = d->d_stage;
    }
    const size_t tot = nvec * q;
    OsdDev o = d->osd;
    o.flag = 1;
    if (d->cfg.precision == NBL_PREC_F64) {
        k_convert_lch<double><<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(src, (double *)d->d_lch, nvec, (int)q);
        k_osd<double><<<B, 256, d->osd_smem, d->stream>>>(o, (int)q, nullptr, (const double *)d->d_lch, d->d_osd_sum, d->d_out, d->d_osd_status);
    } else {
        k_convert_lch<float><<<(unsigned)((tot + 255) / 256), 256, 0, d->stream>>>(src, (float *)d->d_lch, nvec, (int)q);
        k_osd<float><<<B, 256, d->osd_smem, d->stream>>>(o, (int)q, nullptr, (const float *)d->d_lch, d->d_osd_sum, d->d_out, d->d_osd_status);
    }
    CK(cudaGetLastError());
    if ((rc = give_ints(d, out_sym, d->d_out, nvec))) return rc;
    if (!is_device_ptr(out_sym)) CK(cudaStreamSynchronize(d->stream));
    return NBL_OK;
}

// ---- front end -------------------------------------------------------------------------------------
static int setup_punct(nbl_decoder *d, int modulation, const int *punct_sym, int n_punct, int *ylen) {
    const int N = d->g.N, per = modulation == 2 ? d->g.p : 1;
    std::vector<int> ys(N);
    std::vector<char> isp(N, 0);
    for (int i = 0; i < n_punct; ++i) {
        if (!punct_sym || punct_sym[i] < 0 || punct_sym[i] >= N) { nbl_set_error("bad puncture list"); return NBL_ERR_ARG; }
        isp[punct_sym[i]] = 1;
    }
    int off = 0;
    for (int s = 0; s < N; ++s) { if (isp[s]) ys[s] = -1; else { ys[s] = off; off += per; } }
    *ylen = off;
    if (d->punct_mod == modulation && d->punct_cache == ys) return NBL_OK;      // layout already on the device
    CK(cudaMemcpyAsync(d->d_ysym, ys.data(), N * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    CK(cudaStreamSynchronize(d->stream));     // ys is a local
    d->punct_cache = ys; d->punct_mod = modulation; d->punct_ylen = off;
    return NBL_OK;
}

// BPSK demapper for q >= 32: ONE WARP per (frame, symbol).  The LLR of x is the sum of its set bits' terms in ascending bit order
// (Comm.cpp:340-407); lane l forms the partial sum over bits 0..4 of x = 32 h + l once and extends it by the set bits of h for each
// of its q / 32 outputs -- the same additions in the same order as one thread per output (bit-identical LLRs), 5 + 3 * 2^(p-5) / 2
// of them per lane instead of p * 2^(p-5), and every store a full 128-byte line.
template <typename T>
__global__ void __launch_bounds__(256) k_demap_bpsk_warp(int B, int N, int q, int p, int ylen, const double *__restrict__ y, double sigma,
                                                       const int *__restrict__ ysym, T *lch /*[B][N][q] or null*/, double *out /*[B][N][q-1] or null*/) {
    const int lane = threadIdx.x & 31;
    const long long fs = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (fs >= (long long)B * N) return;
    const int s = (int)(fs % N), f = (int)(fs / N);
    const int off = ysym[s];
    double v = 0;                                                   // lane b < p: the term of bit b (punctured symbols: all-zero LLRs)
    if (off >= 0 && lane < p) v = -2 * y[(size_t)f * ylen + off + lane] / (sigma * sigma);
    double lo = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const double vb = __shfl_sync(0xffffffffu, v, b);
        if (lane & (1 << b)) lo += vb;
    }
    double hb[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) hb[b] = __shfl_sync(0xffffffffu, v, 5 + b);     // p <= 8
    const int nh = q >> 5;
    for (int h = 0; h < nh; ++h) {
        double a = lo;
#pragma unroll
        for (int b = 0; b < 3; ++b) if (h & (1 << b)) a += hb[b];
        const int x = 32 * h + lane;
        if (lch) lch[(size_t)fs * q + x] = (T)a;
        if (out && x != 0) out[(size_t)fs * (q - 1) + x - 1] = a;
    }
}

// the demapper kernels on frames f0 .. f0 + nf - 1 of samples already on the device (ysrc = first frame of the batch)
static int launch_demap(nbl_decoder *d, int modulation, const double *ysrc, int f0, int nf, int ylen, double sigma, bool to_lch, double *out_dev) {
    const int N = d->g.N, q = d->g.q;
    const size_t tot = (size_t)nf * N * q;
    const unsigned grid = (unsigned)((tot + 255) / 256);
    const bool f64 = d->cfg.precision == NBL_PREC_F64;
    const double *ys = ysrc + (size_t)f0 * ylen * (modulation == 2 ? 1 : 2);
    const size_t loff = (size_t)f0 * N * q;
    double *od = out_dev ? out_dev + (size_t)f0 * N * (q - 1) : nullptr;
    if (modulation == 2) {
        const int threads = 256, syms = 8 * (threads / q);          // q <= 256
        const unsigned bgrid = (unsigned)(((size_t)nf * N + syms - 1) / syms);
        const size_t sh = (size_t)syms * d->g.p * sizeof(double);
        if (q >= 32 && d->g.p <= 8) {
            const unsigned wgrid = (unsigned)(((size_t)nf * N + 7) / 8);
            if (f64) k_demap_bpsk_warp<double><<<wgrid, 256, 0, d->stream>>>(nf, N, q, d->g.p, ylen, ys, sigma, d->d_ysym, to_lch ? (double *)d->d_lch + loff : nullptr, od);
            else k_demap_bpsk_warp<float><<<wgrid, 256, 0, d->stream>>>(nf, N, q, d->g.p, ylen, ys, sigma, d->d_ysym, to_lch ? (float *)d->d_lch + loff : nullptr, od);
        } else
        if (f64) k_demap_bpsk<double><<<bgrid, threads, sh, d->stream>>>(nf, N, q, d->g.p, ylen, syms, ys, sigma, d->d_ysym, to_lch ? (double *)d->d_lch + loff : nullptr, od);
        else k_demap_bpsk<float><<<bgrid, threads, sh, d->stream>>>(nf, N, q, d->g.p, ylen, syms, ys, sigma, d->d_ysym, to_lch ? (float *)d->d_lch + loff : nullptr, od);
    } else {
        if (f64) k_demap_qam<double><<<grid, 256, 0, d->stream>>>(nf, N, q, ylen, ys, d->d_cons, sigma, d->d_ysym, to_lch ? (double *)d->d_lch + loff : nullptr, od);
        else k_demap_qam<float><<<grid, 256, 0, d->stream>>>(nf, N, q, ylen, ys, d->d_cons, sigma, d->d_ysym, to_lch ? (float *)d->d_lch + loff : nullptr, od);
    }
    CK(cudaGetLastError());
    return NBL_OK;
}

// demap into d_lch (T) and/or a double output
static int demap_common(nbl_decoder *d, int modulation, const double *y, int B, const double *constellation, double sigma,
                        const int *punct_sym, int n_punct, bool to_lch, double *out_dev) {
    const int N = d->g.N, q = d->g.q;
    if (modulation != 2 && modulation != q) {
        nbl_set_error("This module ( code order ~= modulation order ) haven't been developed!");   // Comm.cpp:400-405
        return NBL_ERR_ARG;
    }
    if (!(sigma > 0)) { nbl_set_error("sigma must be positive"); return NBL_ERR_ARG; }
    if ((long long)B * N >= (1LL << 31)) { nbl_set_error("batch too large: B * N must stay below 2^31"); return NBL_ERR_ARG; }
    int ylen, rc;
    if ((rc = setup_punct(d, modulation, punct_sym, n_punct, &ylen))) return rc;
    const size_t ydbl = (size_t)B * ylen * (modulation == 2 ? 1 : 2);
    const double *ysrc = y;
    if (!is_device_ptr(y)) {
        if ((rc = ensure_ystage(d, ydbl * sizeof(double)))) return rc;
        CK(cudaMemcpyAsync(d->d_ystage, y, ydbl * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        ysrc = d->d_ystage;
    }
    if (modulation != 2) {
        if (!constellation) { nbl_set_error("QAM demapping needs a constellation"); return NBL_ERR_ARG; }
        if (constellation != d->d_cons)                             // nbl_simulate hands in the resident copy
            CK(cudaMemcpyAsync(d->d_cons, constellation, 2 * q * sizeof(double), is_device_ptr(constellation) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d->stream));
    }
    return launch_demap(d, modulation, ysrc, 0, B, ylen, sigma, to_lch, out_dev);
}

static int demap_api(nbl_decoder *d, int modulation, const double *y, int B, const double *cons, double sigma, const int *ps, int np, double *L_ch) {
    if (!d || !y || !L_ch) { nbl_set_error("nbl_demap: null argument"); return NBL_ERR_ARG; }
    if (B < 0 || B > d->cfg.max_batch) { nbl_set_error("nbl_demap: B exceeds max_batch"); return NBL_ERR_ARG; }
    if (B == 0) return NBL_OK;
    CK(cudaSetDevice(d->cfg.device));
    const size_t out_bytes = (size_t)B * d->g.N * (d->g.q - 1) * sizeof(double);
    const bool host_out = !is_device_ptr(L_ch);
    int rc;
    if (host_out && (rc = ensure_stage(d, out_bytes))) return rc;
    rc = demap_common(d, modulation, y, B, cons, sigma, ps, np, false, host_out ? d->d_stage : L_ch);
    if (rc) return rc;
    if (host_out) {
        CK(cudaMemcpyAsync(L_ch, d->d_stage, out_bytes, cudaMemcpyDeviceToHost, d->stream));
        CK(cudaStreamSynchronize(d->stream));
    }
    return NBL_OK;
}
extern "C" int nbl_demap_bpsk(nbl_decoder *d, const double *y_re, int B, double sigma, const int *ps, int np, double *L_ch) {
    return demap_api(d, 2, y_re, B, nullptr, sigma, ps, np, L_ch);
}
extern "C" int nbl_demap_qam(nbl_decoder *d, const double *y, int B, const double *cons, double sigma, const int *ps, int np, double *L_ch) {
    if (!d) { nbl_set_error("null decoder"); return NBL_ERR_ARG; }
    return demap_api(d, d->g.q, y, B, cons, sigma, ps, np, L_ch);
}

extern "C" int nbl_demap_decode_batch(nbl_decoder *d, int modulation, const double *y, int B, const double *cons, double sigma,
                                      const int *ps, int np, int *out_sym, int *converged, int *iterations) {
    if (!d || !y || !out_sym) { nbl_set_error("nbl_demap_decode_batch: null argument"); return NBL_ERR_ARG; }
    if (B < 0 || B > d->cfg.max_batch) { nbl_set_error("nbl_demap_decode_batch: B exceeds max_batch"); return NBL_ERR_ARG; }
    if (B == 0) return NBL_OK;
    CK(cudaSetDevice(d->cfg.device));
    int rc;
    const bool osd_acc = d->osd_on && d->osd.order >= 0 && !d->osd.flag;
    if (B >= NBL_PIPE_MIN_B && (long long)B * d->g.N < (1LL << 31) && !is_device_ptr(y) && d->cfg.max_iter >= 2 && !osd_acc && (modulation == 2 || modulation == d->g.q) && sigma > 0 &&
        (modulation == 2 || cons)) {
        // pipelined: samples in NBL_PIPE_CHUNKS chunks on the copy stream; demapper and iteration 1 of a chunk behind its copy
        int ylen;
        if ((rc = setup_punct(d, modulation, ps, np, &ylen))) return rc;
        const size_t per_frame = (size_t)ylen * (modulation == 2 ? 1 : 2);
        if ((rc = ensure_ystage(d, (size_t)B * per_frame * sizeof(double)))) return rc;
        if (modulation != 2 && cons != d->d_cons)
            CK(cudaMemcpyAsync(d->d_cons, cons, 2 * d->g.q * sizeof(double), is_device_ptr(cons) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d->stream));
        if (!d->copy_stream) CK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
        while ((int)d->chunk_ev.size() < NBL_PIPE_CHUNKS + 1) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); d->chunk_ev.push_back(e); }
        const int per = (B + NBL_PIPE_CHUNKS - 1) / NBL_PIPE_CHUNKS;
        // the staging buffer may still be read by work queued on the decoder's stream (an asynchronous caller): order the copies behind it
        CK(cudaEventRecord(d->chunk_ev[NBL_PIPE_CHUNKS], d->stream));
        CK(cudaStreamWaitEvent(d->copy_stream, d->chunk_ev[NBL_PIPE_CHUNKS], 0));
        for (int ck = 0; ck < NBL_PIPE_CHUNKS; ++ck) {
            const int f0 = ck * per, nf = std::min(per, B - f0);
            if (nf > 0) CK(cudaMemcpyAsync(d->d_ystage + (size_t)f0 * per_frame, y + (size_t)f0 * per_frame, (size_t)nf * per_frame * sizeof(double), cudaMemcpyHostToDevice, d->copy_stream));
            CK(cudaEventRecord(d->chunk_ev[ck], d->copy_stream));
        }
        d->pipe.active = true; d->pipe.modulation = modulation; d->pipe.ylen = ylen; d->pipe.chunks = NBL_PIPE_CHUNKS; d->pipe.per = per;
        d->pipe.sigma = sigma; d->pipe.ysrc = d->d_ystage;
        rc = decode_common(d, B, out_sym, converged, iterations, nullptr);
        d->pipe.active = false;
        return rc;
    }
    rc = demap_common(d, modulation, y, B, cons, sigma, ps, np, true, nullptr);
    if (rc) return rc;
    return decode_common(d, B, out_sym, converged, iterations, nullptr);
}

// ======================================================================================================
// device Monte-Carlo: source -> map -> AWGN -> demap -> decode -> error counters
// (the frame loop of the reference's main.cpp:43-55 with CComm::Transmission / CComm::Err per frame)
// ======================================================================================================
// Philox4x32-10 (Salmon et al., SC'11), counter-based: the noise of a frame depends only on
// (seed, global frame index, sample index), never on batch size, stream or GPU.
__device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two independent N(0,1) from one Philox block (Box-Muller, 53-bit radius uniform => tails to ~8.5 sigma)
__device__ inline void normal_pair(uint64_t seed, uint64_t frame, uint32_t idx, double &z0, double &z1) {
    uint32_t r[4];
    philox4x32_10((uint32_t)frame, (uint32_t)(frame >> 32), idx, 0x4E424Cu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = ((double)((((uint64_t)r[0] << 32) | r[1]) >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
    const double u2 = (double)((((uint64_t)r[2] << 32) | r[3]) >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = rad * c; z1 = rad * s;
}

// transmitted code words of B frames: frame f takes pool entry (first_frame + f) % n_pool, or all zeros
__global__ void k_sim_source_pool(int B, int N, uint64_t first_frame, const uint8_t *__restrict__ pool, int n_pool, uint8_t *__restrict__ tx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    const int f = (int)(i / N), s = (int)(i % N);
    tx[i] = n_pool ? pool[(size_t)((first_frame + (uint64_t)f) % (uint64_t)n_pool) * N + s] : (uint8_t)0;
}
// K uniform message symbols per frame (the reference's randomMsg = 1 source, Comm.cpp:214-249, with a counter-based generator)
__global__ void k_sim_message(int B, int K, int q, uint64_t seed, uint64_t first_frame, uint8_t *__restrict__ msg) {
    const int per = (K + 3) / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * per) return;
    const int f = (int)(i / per), j = (int)(i % per);
    const uint64_t gf = first_frame + (uint64_t)f;
    uint32_t r[4];
    philox4x32_10((uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)j, 0x4D5347u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    for (int h = 0; h < 4; ++h)
        if (4 * j + h < K) msg[(size_t)f * K + 4 * j + h] = (uint8_t)(r[h] & (uint32_t)(q - 1));
}

// CRC generator taps as bit masks of the shift register (Comm.cpp:513-531): bit j set <=> G[j] = 1, j = 0..crc_len-1
__device__ __host__ inline unsigned crc_taps(int crc_len, int type24) {
    if (crc_len == 8) return 0x9Bu;                                 // x^8 + x^7 + x^4 + x^3 + x + 1
    if (crc_len == 16) return 0x1021u;                              // x^16 + x^12 + x^5 + 1
    if (crc_len == 24) return type24 ? 0x800063u : 0x864CFBu;       // CRC24B / CRC24A
    return 0u;
}
// bit b of message symbol s is message bit s * p + b (LSB first, Comm.cpp:234-249)
__device__ inline int msg_bit(const uint8_t *m, int p, int i) { return (m[i / p] >> (i % p)) & 1; }

// CComm::CRCEncode (Comm.cpp:506-560): the last crc_len message bits become the CRC of the first K p - crc_len
__global__ void k_sim_crc(int B, int K, int p, int crc_len, uint8_t *__restrict__ msg) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    uint8_t *m = msg + (size_t)f * K;
    const int n = K * p - crc_len;
    const unsigned taps = crc_taps(crc_len, 0), top = 1u << (crc_len - 1), mask = (top << 1) - 1u;
    unsigned reg = 0;
    for (int i = 0; i < n; ++i) {
        const unsigned fb = ((reg & top) ? 1u : 0u) ^ (unsigned)msg_bit(m, p, i);
        reg = ((reg << 1) & mask & ~1u) ^ (fb ? (taps & ~1u) : 0u) | fb;          // reg[j] = reg[j-1] ^ (G[j] & fb), reg[0] = fb
    }
    for (int i = 0; i < crc_len; ++i) {
        const int bit = (reg >> (crc_len - 1 - i)) & 1, pos = n + i;
        m[pos / p] = (uint8_t)((m[pos / p] & ~(1u << (pos % p))) | ((unsigned)bit << (pos % p)));
    }
}

// CComm::CrcCheck (Comm.cpp:562-630) on the K p decoded message bits
__device__ inline bool crc_check(const int *sym, int K, int p, int crc_len) {
    const unsigned taps = crc_taps(crc_len, 1), top = 1u << (crc_len - 1), mask = (top << 1) - 1u;
    unsigned reg = 0; int ones = 0;
    for (int i = 0; i < K * p; ++i) {
        const unsigned in = (unsigned)((sym[i / p] >> (i % p)) & 1);
        const unsigned fb = (reg & top) ? 1u : 0u;
        reg = ((reg << 1) & mask & ~1u) ^ (fb ? (taps & ~1u) : 0u) | (fb ^ in);
        ones += (int)in;
    }
    return reg == 0 && ones != 0;                                   // an all-zero sequence never passes (Comm.cpp:618-628)
}

// received samples of B frames.  BPSK: y[f][N*p] = (1 - 2 bit) + sigma n  (0 -> +1, 1 -> -1, BPSK.txt; bit b of
// symbol s is sample s*p+b, Comm.cpp:272-279).  QAM: y[f][N][2] = constellation[symbol] + sigma (n_re, n_im).
// ysym[s] = offset of symbol s's first sample inside a frame's y (punctured symbols: -1, not transmitted, Comm.cpp:290-308);
// the noise of a sample depends on its position in the UNPUNCTURED frame, so puncturing does not move the other samples' noise.
__global__ void k_sim_channel(int B, int N, int p, int modulation, uint64_t seed, uint64_t first_frame, double sigma,
                              const uint8_t *__restrict__ tx, const double *__restrict__ cs, const int *__restrict__ ysym, int ylen,
                              double *__restrict__ y) {
    const int per = modulation == 2 ? (N * p + 1) / 2 : N;          // Philox blocks per frame (2 normals each)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * per) return;
    const int f = (int)(i / per), j = (int)(i % per);
    const uint64_t gf = first_frame + (uint64_t)f;
    const uint8_t *cw = tx + (size_t)f * N;
    double z0, z1;
    normal_pair(seed, gf, (uint32_t)j, z0, z1);
    if (modulation == 2) {
        const int L = N * p;
        for (int h = 0; h < 2; ++h) {
            const int b = 2 * j + h;
            if (b >= L) break;
            const int bit = (cw[b / p] >> (b % p)) & 1, off = ysym[b / p];
            if (off >= 0) y[(size_t)f * ylen + off + b % p] = (bit ? -1.0 : 1.0) + sigma * (h ? z1 : z0);
        }
    } else {
        const int sym = cw[j], off = ysym[j];
        if (off >= 0) {
            y[((size_t)f * ylen + off) * 2] = cs[2 * sym] + sigma * z0;
            y[((size_t)f * ylen + off) * 2 + 1] = cs[2 * sym + 1] + sigma * z1;
        }
    }
}

// CComm::Err (Comm.cpp:446-503) per frame on the first K symbols, accumulated into uint64 counters
__global__ void k_sim_count(int B, int N, int K, int p, int crc_len, const uint8_t *__restrict__ tx,
                            const int *__restrict__ out_sym, const int *__restrict__ conv, const int *__restrict__ iters,
                            unsigned long long *counters) {
    __shared__ unsigned long long acc[NBL_NUM_COUNTERS];
    if (threadIdx.x < NBL_NUM_COUNTERS) acc[threadIdx.x] = 0;
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < B) {
        const uint8_t *cw = tx + (size_t)f * N;
        int es = 0, eb = 0;
        for (int s = 0; s < K; ++s) {
            const int diff = out_sym[(size_t)f * N + s] ^ cw[s];
            es += diff != 0; eb += __popc(diff);
        }
        atomicAdd(&acc[NBL_CNT_FRAMES], 1ULL);
        if (es) { atomicAdd(&acc[NBL_CNT_ERR_FRAMES], 1ULL); atomicAdd(&acc[NBL_CNT_ERR_SYMS], (unsigned long long)es); atomicAdd(&acc[NBL_CNT_ERR_BITS], (unsigned long long)eb); }
        if (conv[f]) atomicAdd(&acc[NBL_CNT_CONVERGED], 1ULL);
        if (crc_len > 0) {
            const bool ok = crc_check(out_sym + (size_t)f * N, K, p, crc_len);
            if (ok) atomicAdd(&acc[NBL_CNT_CRC_OK], 1ULL);
            if (ok && es) atomicAdd(&acc[NBL_CNT_UNDETECTED], 1ULL);
        } else if (conv[f] && es) atomicAdd(&acc[NBL_CNT_UNDETECTED], 1ULL);
        atomicAdd(&acc[NBL_CNT_ITERATIONS], (unsigned long long)iters[f]);
    }
    __syncthreads();
    if (threadIdx.x < NBL_NUM_COUNTERS && acc[threadIdx.x]) atomicAdd(&counters[threadIdx.x], acc[threadIdx.x]);
}

static int simulate_impl(nbl_decoder *d, const nbl_sim_config *cfg, unsigned long long *dev_counters) {
    const int N = d->g.N, q = d->g.q, p = d->g.p, K = N - d->g.M;
    if (cfg->modulation != 2 && cfg->modulation != q) {
        nbl_set_error("This module ( code order ~= modulation order ) haven't been developed!"); return NBL_ERR_ARG;
    }
    if (cfg->modulation != 2 && !cfg->constellation) { nbl_set_error("nbl_simulate: QAM needs a constellation"); return NBL_ERR_ARG; }
    if (cfg->n_codewords < 0 || (cfg->n_codewords > 0 && !cfg->codewords)) { nbl_set_error("nbl_simulate: bad codeword pool"); return NBL_ERR_ARG; }
    const double mbits = cfg->modulation == 2 ? 1.0 : (double)p;
    if (cfg->n_punct < 0 || cfg->n_punct >= N || (cfg->n_punct > 0 && !cfg->punct_sym)) { nbl_set_error("nbl_simulate: bad puncture list"); return NBL_ERR_ARG; }
    const double rate = (double)K / (double)(N - cfg->n_punct);                                 // Comm.cpp:98: R = K / (N - punctured)
    const double sigma = 1.0 / sqrt(2.0 * mbits * rate * pow(10.0, cfg->ebn0_db / 10.0));      // Comm.cpp:176-177
    int rc, ylen;
    if ((rc = setup_punct(d, cfg->modulation, cfg->punct_sym, cfg->n_punct, &ylen))) return rc;
    // codeword pool -> device bytes
    if (cfg->n_codewords && !(d->pool_host.size() == (size_t)cfg->n_codewords * N &&
                              memcmp(d->pool_host.data(), cfg->codewords, d->pool_host.size() * sizeof(int)) == 0)) {
        const size_t n = (size_t)cfg->n_codewords * N;
        if (d->pool_cap < n) { if (d->d_pool) cudaFree(d->d_pool); d->d_pool = nullptr; CK(cudaMalloc((void **)&d->d_pool, n)); d->pool_cap = n; }
        std::vector<uint8_t> bytes(n);
        for (size_t i = 0; i < n; ++i) {
            if (cfg->codewords[i] < 0 || cfg->codewords[i] >= q) { nbl_set_error("nbl_simulate: codeword symbol out of range"); return NBL_ERR_ARG; }
            bytes[i] = (uint8_t)cfg->codewords[i];
        }
        CK(cudaMemcpyAsync(d->d_pool, bytes.data(), n, cudaMemcpyHostToDevice, d->stream));
        CK(cudaStreamSynchronize(d->stream));
        d->pool_host.assign(cfg->codewords, cfg->codewords + n);        // same pool next call: nothing to upload
    }
    if (cfg->modulation != 2) CK(cudaMemcpyAsync(d->d_cons, cfg->constellation, 2 * q * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    const size_t ydbl = (size_t)d->cfg.max_batch * ylen * (cfg->modulation == 2 ? 1 : 2);
    if ((rc = ensure_ystage(d, ydbl * sizeof(double)))) return rc;
    const bool rnd = cfg->n_codewords == 0 && cfg->random_msg != 0;
    if (cfg->crc_len != 0 && cfg->crc_len != 8 && cfg->crc_len != 16 && cfg->crc_len != 24) { nbl_set_error("nbl_simulate: crc_len must be 0, 8, 16 or 24"); return NBL_ERR_ARG; }
    if (cfg->crc_len >= K * p) { nbl_set_error("nbl_simulate: crc_len does not fit into the message"); return NBL_ERR_ARG; }
    if (rnd && d->osd_on && d->osd.order >= 0 && d->osd_cfg.crc_correct_len > 0 && cfg->crc_len != d->osd_cfg.crc_len) {
        // the CRC rows of the OSD matrix constrain the last message bits: the messages must carry that CRC
        nbl_set_error("nbl_simulate: OSD with CRC rows needs messages with the same CRC: set nbl_sim_config.crc_len = nbl_osd_config.crc_len "
                      "(or use crc_correct_len = 0)");
        return NBL_ERR_ARG;
    }
    if (!d->d_tx) CK(cudaMalloc((void **)&d->d_tx, (size_t)d->cfg.max_batch * N));
    if (rnd) {
        if (!d->sim_enc && (rc = nbl_encoder_create(&d->sim_enc, d->code, d->cfg.device, d->stream))) return rc;
        if (!d->d_msg) CK(cudaMalloc((void **)&d->d_msg, (size_t)d->cfg.max_batch * K));
    }
    int launches = 0;
    for (uint64_t done = 0; done < cfg->frames;) {
        const int B = (int)std::min<uint64_t>(d->cfg.max_batch, cfg->frames - done);
        const int per = cfg->modulation == 2 ? (N * p + 1) / 2 : N;
        const long long nthr = (long long)B * per, nsym = (long long)B * N;
        if (rnd) {
            const long long nm = (long long)B * ((K + 3) / 4);
            k_sim_message<<<(unsigned)((nm + 255) / 256), 256, 0, d->stream>>>(B, K, q, cfg->seed, cfg->first_frame + done, d->d_msg);
            CK(cudaGetLastError());
            if (cfg->crc_len > 0) {
                k_sim_crc<<<(B + 127) / 128, 128, 0, d->stream>>>(B, K, p, cfg->crc_len, d->d_msg);
                CK(cudaGetLastError());
                launches += 1;
            }
            if ((rc = nbl_encoder_launch_u8(d->sim_enc, d->d_msg, B, d->d_tx, d->stream))) return rc;
            launches += 1;
        } else {
            k_sim_source_pool<<<(unsigned)((nsym + 255) / 256), 256, 0, d->stream>>>(B, N, cfg->first_frame + done, d->d_pool, cfg->n_codewords, d->d_tx);
            CK(cudaGetLastError());
        }
        k_sim_channel<<<(unsigned)((nthr + 255) / 256), 256, 0, d->stream>>>(B, N, p, cfg->modulation, cfg->seed, cfg->first_frame + done, sigma,
                                                                            d->d_tx, d->d_cons, d->d_ysym, ylen, d->d_ystage);
        CK(cudaGetLastError());
        if ((rc = demap_common(d, cfg->modulation, d->d_ystage, B, d->d_cons, sigma, cfg->punct_sym, cfg->n_punct, true, nullptr))) return rc;
        rc = d->cfg.precision == NBL_PREC_F64 ? run_decode_f64(d, B, nullptr) : run_decode<float>(d, B, nullptr);
        if (rc) return rc;
        k_sim_count<<<(B + 127) / 128, 128, 0, d->stream>>>(B, N, K, p, cfg->crc_len, d->d_tx, d->d_out, d->d_conv, d->d_iters, dev_counters);
        CK(cudaGetLastError());
        launches += d->last_launches + 4;
        done += B;
    }
    d->last_launches = launches;
    return NBL_OK;
}

extern "C" int nbl_simulate_async(nbl_decoder *d, const nbl_sim_config *cfg, uint64_t *dev_counters) {
    if (!d || !cfg || !dev_counters) { nbl_set_error("nbl_simulate_async: null argument"); return NBL_ERR_ARG; }
    if (!is_device_ptr(dev_counters)) { nbl_set_error("nbl_simulate_async: counters must be device memory"); return NBL_ERR_ARG; }
    CK(cudaSetDevice(d->cfg.device));
    return simulate_impl(d, cfg, (unsigned long long *)dev_counters);
}

// One round of a sharded Monte-Carlo run: the multi-GPU form of one pass of the reference's frame loop (main.cpp:43-55: `parallel`
// workers simulate one frame each, then their counters are added).  Local rank i (global rank r_i of `world`) simulates the
// frames_per_rank frames starting at  cfg->first_frame + (round_index * world + r_i) * frames_per_rank  on decoders[i]; the
// counter vectors are summed over all ranks on the device (nbl_allreduce_counters) and the total of the round is ADDED to
// totals (host uint64[8]).  Every rank gets the same totals, so every rank takes the same stop decision.
extern "C" int nbl_simulate_sharded(nbl_comm *comm, nbl_decoder *const *decoders, const nbl_sim_config *cfg, uint64_t frames_per_rank,
                                    uint64_t round_index, uint64_t *totals) {
    if (!comm || !decoders || !cfg || !totals) { nbl_set_error("nbl_simulate_sharded: null argument"); return NBL_ERR_ARG; }
    const int n = nbl_comm_local_ranks(comm), world = nbl_comm_size(comm);
    std::vector<uint64_t *> dev(n);
    for (int i = 0; i < n; ++i) {
        nbl_decoder *d = decoders[i];
        if (!d) { nbl_set_error("nbl_simulate_sharded: null decoder"); return NBL_ERR_ARG; }
        CK(cudaSetDevice(d->cfg.device));
        if (!d->d_counters) CK(cudaMalloc((void **)&d->d_counters, NBL_NUM_COUNTERS * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(d->d_counters, 0, NBL_NUM_COUNTERS * sizeof(unsigned long long), d->stream));
        nbl_sim_config c = *cfg;
        c.frames = frames_per_rank;
        c.first_frame = cfg->first_frame + (round_index * (uint64_t)world + (uint64_t)nbl_comm_rank(comm, i)) * frames_per_rank;
        const int rc = simulate_impl(d, &c, d->d_counters);         // asynchronous: the GPUs of this process run concurrently
        if (rc) return rc;
        dev[i] = (uint64_t *)d->d_counters;
    }
    int rc = nbl_allreduce_counters(comm, decoders, dev.data());
    if (rc) return rc;
    unsigned long long host[NBL_NUM_COUNTERS];
    CK(cudaSetDevice(decoders[0]->cfg.device));
    CK(cudaMemcpyAsync(host, decoders[0]->d_counters, sizeof host, cudaMemcpyDeviceToHost, decoders[0]->stream));
    for (int i = 0; i < n; ++i) { CK(cudaSetDevice(decoders[i]->cfg.device)); CK(cudaStreamSynchronize(decoders[i]->stream)); }
    for (int i = 0; i < NBL_NUM_COUNTERS; ++i) totals[i] += host[i];
    return NBL_OK;
}

extern "C" int nbl_simulate(nbl_decoder *d, const nbl_sim_config *cfg, uint64_t *counters) {
    if (!d || !cfg || !counters) { nbl_set_error("nbl_simulate: null argument"); return NBL_ERR_ARG; }
    CK(cudaSetDevice(d->cfg.device));
    if (!d->d_counters) CK(cudaMalloc((void **)&d->d_counters, NBL_NUM_COUNTERS * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(d->d_counters, 0, NBL_NUM_COUNTERS * sizeof(unsigned long long), d->stream));
    int rc = simulate_impl(d, cfg, d->d_counters);
    if (rc) return rc;
    unsigned long long host[NBL_NUM_COUNTERS];
    CK(cudaMemcpyAsync(host, d->d_counters, sizeof host, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    for (int i = 0; i < NBL_NUM_COUNTERS; ++i) counters[i] += host[i];
    return NBL_OK;
}
