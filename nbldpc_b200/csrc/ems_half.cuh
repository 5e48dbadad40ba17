// Fused iteration kernel for EMS over GF(16): four lanes per (frame, check row), four field elements per lane, eight
// check rows per warp -- the layout of qspa_half.cuh, whose row set-up and variable-node phase it shares.
//
// Same check node as the group kernel (cn_ems.cuh; reference NBLDPC.cpp:859-918, 1715-1786): the n_m largest entries of
// every input in the reference's sorted order (ties: larger symbol first), conf(q,1) in gather form, conf(n_m,n_c) as a
// max-plus XOR-convolution with the deviation count as a second state index, sums formed left to right over the inputs.
// What changes is who does it: the 40 arg-max rounds of a d_c = 5 row are two in-lane steps and two shuffles each instead
// of a group reduction through shared-memory scratch, and the rows of a warp run one instruction stream.
// Serves q = 16, d_c <= 8, d_v <= 4, n_m <= 16, n_c <= 3.
#pragma once
#include "qspa_half.cuh"

namespace nbl {

constexpr int EH_MAXNM = 16;
constexpr int EH_MAXNJ = 4;            // n_c + 1

template <typename T> struct EhSmem {
    // per row: V [8][16] v2c of the edges; W [2][4][16] DP buffers; best [16]; sval [8][16] sorted values; ssyn [8][16] bytes
    static constexpr size_t row_elems = 8 * 16 + 2 * EH_MAXNJ * 16 + 16 + 8 * EH_MAXNM;
    static constexpr size_t row_bytes = row_elems * sizeof(T) + 8 * EH_MAXNM;
    static constexpr size_t bytes_per_warp = row_bytes * QH_ROWS;
    static constexpr size_t bytes = bytes_per_warp * QH_WARPS;
};

template <typename T>
__global__ void __launch_bounds__(32 * QH_WARPS, sizeof(T) == 4 ? 4 : 2) k_iterate_ems_half(const IterParams<T> P, const int *__restrict__ desc) {
    constexpr int Q = 16;
    using K = typename Real<T>::key_t;
    extern __shared__ __align__(16) unsigned char eh_raw[];
    const DevGraph &g = P.g;
    QhRow<T> R;
    if (!qh_setup<T>(P, R)) return;
    const int warp = threadIdx.x >> 5, l4 = R.l4, x0 = R.x0, f = R.f, e0 = R.e0, dcn = R.dcn, dmax = R.dmax;
    const T NINF = -Real<T>::inf();
    T *sV = reinterpret_cast<T *>(eh_raw + EhSmem<T>::bytes_per_warp * warp + EhSmem<T>::row_bytes * R.rw);   // [8][16]
    T *sW = sV + 8 * 16;                                            // [2][4][16]
    T *sBest = sW + 2 * EH_MAXNJ * 16;                              // [16]
    T *sval = sBest + 16;                                           // [8][16]
    uint8_t *ssyn = reinterpret_cast<uint8_t *>(sval + 8 * EH_MAXNM);   // [8][16]
    T V[QH_MAXDC][4];
    unsigned tf[QH_MAXDC], ti[QH_MAXDC];
    unsigned sdw;
    qh_vn_phase<T>(P, desc, R, V, tf, ti, sdw);
    const int nm = P.p0, nj = P.p1 + 1;

    // ---- the n_m largest entries of each input, in the reference's sorted order (SortLLRVector, NBLDPC.cpp:1715-1746) --------
#pragma unroll
    for (int d = 0; d < QH_MAXDC; ++d) {
        if (d < dmax) {
            qh_store4<T>(sV + d * 16 + x0, V[d]);
            K key[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) key[i] = order_key(V[d][i]);
#pragma unroll 1
            for (int r = 0; r < nm; ++r) {
                K bk = key[0]; int bi = 0;                          // larger key first, larger symbol among equals; key 0 = taken
#pragma unroll
                for (int i = 1; i < 4; ++i) if (key[i] >= bk) { bk = key[i]; bi = i; }
                int bx = x0 + bi;
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const K ok = __shfl_xor_sync(QH_FULL, bk, o);
                    const int ox = __shfl_xor_sync(QH_FULL, bx, o);
                    if (ok > bk || (ok == bk && ox > bx)) { bk = ok; bx = ox; }
                }
                if ((bx >> 2) == l4) {                              // the owner records it and retires it
                    const int i = bx & 3;
                    if (d < dcn) {
                        sval[d * EH_MAXNM + r] = i == 0 ? V[d][0] : i == 1 ? V[d][1] : i == 2 ? V[d][2] : V[d][3];
                        ssyn[d * EH_MAXNM + r] = (uint8_t)((tf[d] >> (8 * i)) & 255u);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j == i) key[j] = K(0);
                }
            }
        }
    }
    __syncwarp();
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + x0;
#pragma unroll 1
    for (int d = 0; d < dmax; ++d) {
        const bool live = d < dcn;
        // ---- conf(q,1): one input l != d takes any symbol, the others their best one (ConstructConf, :1748-1786) ----------------
        T b1[4] = {NINF, NINF, NINF, NINF};
        if (live) {
            for (int l = 0; l < dcn; ++l) {
                if (l == d) continue;
                int rest0 = 0;                                      // syndrome of the other inputs' best symbols
                for (int m = 0; m < dcn; ++m) if (m != d && m != l) rest0 ^= ssyn[m * EH_MAXNM];
                const uint8_t *pinv_l = g.pinv + (size_t)(e0 + l) * Q;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int y = pinv_l[(x0 + i) ^ rest0];         // the symbol the deviating input must take
                    T sum = T(0);
                    for (int m = 0; m < dcn; ++m) {
                        if (m == d) continue;
                        sum = sum + (m == l ? sV[l * 16 + y] : sval[m * EH_MAXNM]);
                    }
                    b1[i] = sum > b1[i] ? sum : b1[i];
                }
            }
        }
        // ---- conf(n_m, n_c): W[j][s] over the inputs l != d, j = number of non-best picks ------------------------------------------
        T *Wa = sW, *Wb = sW + EH_MAXNJ * 16;
        for (int j = 0; j < nj; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) Wa[j * 16 + x0 + i] = (j == 0 && x0 + i == 0) ? T(0) : NINF;
        for (int l = 0; l < dmax; ++l) {                            // warp-uniform trip count, rows skip their own d and missing inputs
            __syncwarp();
            if (live && l != d && l < dcn) {
                const T s0 = sval[l * EH_MAXNM];
                const int y0 = ssyn[l * EH_MAXNM];
                for (int j = 0; j < nj; ++j) {
                    T w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = Wa[j * 16 + ((x0 + i) ^ y0)] + s0;
                    if (j > 0)
                        for (int k = 1; k < nm; ++k) {
                            const T sk = sval[l * EH_MAXNM + k];
                            const int yk = ssyn[l * EH_MAXNM + k];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const T c = Wa[(j - 1) * 16 + ((x0 + i) ^ yk)] + sk;
                                w[i] = c > w[i] ? c : w[i];
                            }
                        }
#pragma unroll
                    for (int i = 0; i < 4; ++i) Wb[j * 16 + x0 + i] = w[i];
                }
                T *t = Wa; Wa = Wb; Wb = t;
            }
        }
        __syncwarp();
        {
            T bst[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bst[i] = b1[i];
                for (int j = 0; j < nj; ++j) { const T w = Wa[j * 16 + x0 + i]; bst[i] = w > bst[i] ? w : bst[i]; }
            }
            qh_store4<T>(sBest + x0, bst);
        }
        __syncwarp();
        // ---- output (NBLDPC.cpp:899-916) ------------------------------------------------------------------------------------------
        if (live) {
            T out[4];
            const T b0 = sBest[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned tw = 0u;
#pragma unroll
                for (int m = 0; m < QH_MAXDC; ++m) if (m == d) tw = tf[m];
                out[i] = (x0 + i == 0) ? T(0) : soft_threshold<T>(sBest[(tw >> (8 * i)) & 15u] - b0, P.factor, P.offset);
            }
            qh_store4<T>(coutf + (size_t)(e0 + d) * Q, out);
        }
        __syncwarp();                                               // best / W are rewritten for the next output
    }
}

}  // namespace nbl
