// Systematic GF(q) encoder: replaces CNBLDPC::InitialEncode / CNBLDPC::Encode of the reference
// (NBLDPC.cpp:477-604, helpers 1474-1538).
//
// Set-up (host, once per code): the reference reduces H from the last row to the first with pivot column K + row; a zero
// pivot is repaired by a row swap from above, else by a column swap from the left (recorded, undone after encoding); the
// pivot column is cleared above the pivot and the row scaled to a unit pivot.  What is left of each pivot is the rule
// "parity r = sum of coefficients x earlier symbols" that Encode() applies by forward substitution.  We run the same
// elimination and then SUBSTITUTE the rules into each other once, which leaves every parity symbol as a function of the
// K message symbols alone (a dense M x K matrix G over GF(q)); the recorded column swaps become one output permutation.
// GF arithmetic is exact, so the code words are the reference's symbol for symbol.
//
// Encode (device, per batch): thread <-> (frame, output position): a message symbol is copied, a parity symbol is
// sum_c G[r][c] * msg[c] with log/antilog tables in shared memory.  Frames are independent; nothing is sequential.
#include "internal.h"
#include "enc_internal.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>

#define CKE(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            nbl_set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_));              \
            return NBL_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

constexpr uint16_t LOG_ZERO = 0xffff;

// msg: [B][K] symbols (int or u8), out: [B][N] (int and/or u8, either may be null)
template <typename MsgT>
__global__ void k_encode(int B, int N, int K, int q, const MsgT *__restrict__ msg, const uint16_t *__restrict__ logG,
                         const int *__restrict__ perm, const uint16_t *__restrict__ lg, const uint16_t *__restrict__ alog,
                         int *out_i, uint8_t *out_b) {
    extern __shared__ uint16_t s_tab[];                              // log[q], antilog[2q]
    uint16_t *s_log = s_tab, *s_alog = s_tab + q;
    for (int i = threadIdx.x; i < q; i += blockDim.x) s_log[i] = lg[i];
    for (int i = threadIdx.x; i < 2 * q; i += blockDim.x) s_alog[i] = alog[i];
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * N) return;
    const int f = (int)(t / N), i = (int)(t % N);
    const int pre = perm[i];                                         // position before the column swaps are undone
    const MsgT *m = msg + (size_t)f * K;
    int v;
    if (pre < K) v = (int)m[pre];
    else {
        const uint16_t *g = logG + (size_t)(pre - K) * K;
        v = 0;
        for (int c = 0; c < K; ++c) {
            const int x = (int)m[c];
            const uint16_t lgc = g[c];
            if (x != 0 && lgc != LOG_ZERO) v ^= s_alog[lgc + s_log[x]];
        }
    }
    if (out_i) out_i[t] = v;
    if (out_b) out_b[t] = (uint8_t)v;
}

extern "C" void nbl_encoder_destroy(nbl_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    void *ptrs[] = {e->d_logG, e->d_perm, e->d_log, e->d_alog, e->d_msg, e->d_out};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

// host part of the set-up; fills e->G (M x K), e->perm (N), e->n_swaps
static int eliminate(const nbl_code *code, nbl_encoder *e) {
    const int N = code->N, M = code->M, K = N - M, q = code->q;
    if (K <= 0) { nbl_set_error("nbl_encoder_create: the code has no message symbols (N <= M)"); return NBL_ERR_ARG; }
    auto mul = [&](int a, int b) { return code->mul[(size_t)a * q + b]; };
    std::vector<std::vector<int>> H(M, std::vector<int>(N, 0));
    for (int r = 0; r < M; ++r)
        for (int s = code->chk_ptr[r]; s < code->chk_ptr[r + 1]; ++s) H[r][code->chk_var[s]] = code->chk_coef[s];
    std::vector<std::pair<int, int>> swaps;
    for (int r = M - 1; r >= 0; --r) {
        const int pc = K + r;
        if (H[r][pc] == 0) {
            bool fixed = false;
            for (int up = r - 1; up >= 0 && !fixed; --up)
                if (H[up][pc] != 0) { std::swap(H[r], H[up]); fixed = true; }
            for (int left = pc - 1; left >= 0 && !fixed; --left)
                if (H[r][left] != 0) {
                    for (int rr = 0; rr < M; ++rr) std::swap(H[rr][pc], H[rr][left]);
                    swaps.push_back({pc, left});
                    fixed = true;
                }
            if (!fixed) { nbl_set_error("NB matrix is not full rank"); return NBL_ERR_ARG; }   // the reference prints this and exits
        }
        const int inv = code->inv[H[r][pc]];
        for (int up = r - 1; up >= 0; --up)
            if (H[up][pc] != 0) {
                const int x = mul(inv, H[up][pc]);
                for (int c = 0; c < N; ++c) H[up][c] ^= mul(x, H[r][c]);
            }
        for (int c = 0; c <= pc; ++c) H[r][c] = mul(H[r][c], inv);
    }
    // parity r = sum_{c < K + r} H[r][c] * symbol c  (characteristic 2: no sign); substitute earlier parities
    e->G.assign((size_t)M * K, 0);
    for (int r = 0; r < M; ++r) {
        int *g = &e->G[(size_t)r * K];
        for (int c = 0; c < K; ++c) g[c] = H[r][c];
        for (int pr = 0; pr < r; ++pr) {
            const int h = H[r][K + pr];
            if (!h) continue;
            const int *gp = &e->G[(size_t)pr * K];
            for (int c = 0; c < K; ++c) g[c] ^= mul(h, gp[c]);
        }
    }
    // Encode() undoes the swaps last-recorded first; as one gather: out[i] = pre[perm[i]]
    e->perm.resize(N);
    for (int i = 0; i < N; ++i) e->perm[i] = i;
    std::vector<int> pos(N);                                           // simulate the swaps on an index vector
    for (int i = 0; i < N; ++i) pos[i] = i;
    for (int i = (int)swaps.size() - 1; i >= 0; --i) std::swap(pos[swaps[i].first], pos[swaps[i].second]);
    e->perm = pos;
    e->n_swaps = (int)swaps.size();
    return NBL_OK;
}

extern "C" int nbl_encoder_create(nbl_encoder **out, const nbl_code *code, int device, void *stream) {
    if (!out || !code) { nbl_set_error("nbl_encoder_create: null argument"); return NBL_ERR_ARG; }
    nbl_encoder *e = new nbl_encoder;
    e->code = code; e->device = device; e->q = code->q; e->N = code->N; e->M = code->M; e->K = code->N - code->M;
    int rc = eliminate(code, e);
    if (rc) { delete e; return rc; }
    if (nbl_device_count() <= device || device < 0) { delete e; nbl_set_error("no such CUDA device (this library has no CPU fallback)"); return NBL_ERR_CUDA; }
    auto fail = [&](int r) { nbl_encoder_destroy(e); return r; };
    if (cudaSetDevice(device) != cudaSuccess) { nbl_set_error("cudaSetDevice failed"); return fail(NBL_ERR_CUDA); }
    if (stream) e->stream = (cudaStream_t)stream;
    else { if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { nbl_set_error("cudaStreamCreate failed"); return fail(NBL_ERR_CUDA); } e->own_stream = true; }
    const int q = e->q, K = e->K, M = e->M, N = e->N;
    // log / antilog tables from the multiplication table (x = 2 is the primitive element)
    std::vector<uint16_t> lg(q, LOG_ZERO), alog(2 * q, 0);
    int v = 1;
    for (int i = 0; i < q - 1; ++i) { alog[i] = (uint16_t)v; lg[v] = (uint16_t)i; v = code->mul[(size_t)v * q + 2]; }
    for (int i = q - 1; i < 2 * q; ++i) alog[i] = alog[i - (q - 1)];
    std::vector<uint16_t> logG((size_t)M * K);
    for (size_t i = 0; i < logG.size(); ++i) logG[i] = e->G[i] ? lg[e->G[i]] : LOG_ZERO;
    auto up = [&](void **dst, const void *src, size_t bytes) {
        if (cudaMalloc(dst, std::max<size_t>(bytes, 16)) != cudaSuccess) return false;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up((void **)&e->d_logG, logG.data(), logG.size() * 2) || !up((void **)&e->d_perm, e->perm.data(), (size_t)N * 4) ||
        !up((void **)&e->d_log, lg.data(), (size_t)q * 2) || !up((void **)&e->d_alog, alog.data(), (size_t)2 * q * 2)) {
        cudaGetLastError(); nbl_set_error("cudaMalloc/cudaMemcpy failed for the encoder tables"); return fail(NBL_ERR_ALLOC);
    }
    *out = e;
    return NBL_OK;
}

extern "C" int nbl_encoder_get_info(const nbl_encoder *e, int *K, int *n_swaps) {
    if (!e) { nbl_set_error("null encoder"); return NBL_ERR_ARG; }
    if (K) *K = e->K;
    if (n_swaps) *n_swaps = e->n_swaps;
    return NBL_OK;
}

static bool enc_is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int nbl_encoder_launch_u8(nbl_encoder *e, const uint8_t *msg_dev, int B, uint8_t *out_dev, cudaStream_t stream) {
    const long long n = (long long)B * e->N;
    k_encode<uint8_t><<<(unsigned)((n + 255) / 256), 256, (size_t)3 * e->q * 2, stream>>>(B, e->N, e->K, e->q, msg_dev, e->d_logG, e->d_perm,
                                                                                        e->d_log, e->d_alog, nullptr, out_dev);
    CKE(cudaGetLastError());
    return NBL_OK;
}

extern "C" int nbl_encode_batch(nbl_encoder *e, const int *msg, int B, int *code_sym) {
    if (!e || !msg || !code_sym) { nbl_set_error("nbl_encode_batch: null argument"); return NBL_ERR_ARG; }
    if (B < 0) { nbl_set_error("nbl_encode_batch: negative batch"); return NBL_ERR_ARG; }
    if (B == 0) return NBL_OK;
    CKE(cudaSetDevice(e->device));
    const size_t mb = (size_t)B * e->K * sizeof(int), ob = (size_t)B * e->N * sizeof(int);
    const int *msrc = msg; int *odst = code_sym;
    if (!enc_is_device_ptr(msg)) {
        if (e->msg_cap < mb) { if (e->d_msg) cudaFree(e->d_msg); e->d_msg = nullptr; CKE(cudaMalloc((void **)&e->d_msg, mb)); e->msg_cap = mb; }
        CKE(cudaMemcpyAsync(e->d_msg, msg, mb, cudaMemcpyHostToDevice, e->stream));
        msrc = e->d_msg;
    }
    const bool host_out = !enc_is_device_ptr(code_sym);
    if (host_out) {
        if (e->out_cap < ob) { if (e->d_out) cudaFree(e->d_out); e->d_out = nullptr; CKE(cudaMalloc((void **)&e->d_out, ob)); e->out_cap = ob; }
        odst = e->d_out;
    }
    const long long n = (long long)B * e->N;
    k_encode<int><<<(unsigned)((n + 255) / 256), 256, (size_t)3 * e->q * 2, e->stream>>>(B, e->N, e->K, e->q, msrc, e->d_logG, e->d_perm, e->d_log,
                                                                                       e->d_alog, odst, nullptr);
    CKE(cudaGetLastError());
    if (host_out) {
        CKE(cudaMemcpyAsync(code_sym, odst, ob, cudaMemcpyDeviceToHost, e->stream));
        CKE(cudaStreamSynchronize(e->stream));
    }
    return NBL_OK;
}
