// EMS check node, one group of Q threads per check row.
//
// Replaces SortLLRVector / ConstructConf and the output stage of Decoding_EMS (reference
// NBLDPC.cpp:859-918, 1715-1786).  Same result, different algorithm:
//   * the reference insertion-sorts all q entries of every input and then, per output edge, walks
//     conf(q,1) and conf(n_m,n_c) recursively, keeping the best sum per syndrome value;
//   * here the n_m best entries of each input are extracted with n_m group arg-max rounds, conf(q,1) is
//     evaluated in gather form (thread s inverts the one deviating input), and conf(n_m,n_c) is a max-plus
//     XOR-convolution over the other inputs with the deviation count as a second state index
//     (W[j][s], j <= n_c), every thread owning one syndrome value s.  Identical candidate sets, so the
//     maxima agree; sums are formed left to right over the inputs as in the reference.
#pragma once
#include "dev_common.cuh"

namespace nbl {

template <typename T, int Q> struct EmsSmem {
    using K = typename Real<T>::key_t;
    T *V;       // [maxdc][Q]
    T *W;       // [2][nc+1][Q] DP buffers
    T *best;    // [Q]
    T *sval;    // [maxdc][nm] sorted values (descending, ties: larger symbol first)
    int *ssyn;  // [maxdc][nm] h_d * symbol of the sorted entries
    int *ssym;  // [maxdc][nm] symbols
    int *sdec;  // [maxdc] (unused by EMS, keeps the VN stage uniform)
    RedScratch<K, Grp<Q>::WARPS, 1> *red1;
    RedScratch<K, Grp<Q>::WARPS, 2> *red;

    __host__ __device__ static size_t bytes(int maxdc, int nm, int nc) {
        size_t b = sizeof(T) * ((size_t)Q * (maxdc + 2 * (nc + 1) + 1) + (size_t)maxdc * nm);
        b = align_up(b, 8);
        b += sizeof(RedScratch<K, Grp<Q>::WARPS, 1>) + sizeof(RedScratch<K, Grp<Q>::WARPS, 2>);
        b = align_up(b, 4);
        b += sizeof(int) * ((size_t)2 * maxdc * nm + maxdc);
        return align_up(b, 16);
    }
    __device__ void carve(unsigned char *base, int maxdc, int nm, int nc) {
        T *t = reinterpret_cast<T *>(base);
        V = t; t += (size_t)maxdc * Q;
        W = t; t += (size_t)2 * (nc + 1) * Q;
        best = t; t += Q;
        sval = t; t += (size_t)maxdc * nm;
        unsigned char *c = base + align_up((size_t)(reinterpret_cast<unsigned char *>(t) - base), 8);
        red1 = reinterpret_cast<RedScratch<K, Grp<Q>::WARPS, 1> *>(c); c += sizeof(*red1);
        red = reinterpret_cast<RedScratch<K, Grp<Q>::WARPS, 2> *>(c); c += sizeof(*red);
        c = base + align_up((size_t)(c - base), 4);
        ssyn = reinterpret_cast<int *>(c); c += sizeof(int) * (size_t)maxdc * nm;
        ssym = reinterpret_cast<int *>(c); c += sizeof(int) * (size_t)maxdc * nm;
        sdec = reinterpret_cast<int *>(c);
    }
};

template <typename T, int Q>
__device__ void cn_ems(EmsSmem<T, Q> &S, const DevGraph &g, int gid, int x, int e0, int dcn, int nm, int nc,
                       T factor, T offset, int &phase, T *__restrict__ c2v_out) {
    using K = typename Real<T>::key_t;
    const T NINF = -Real<T>::inf();
    // -- the n_m largest entries of each input, in the reference's sorted order ------------------------
    for (int d = 0; d < dcn; ++d) {
        const T v = S.V[d * Q + x];
        bool alive = true;
        for (int r = 0; r < nm; ++r) {
            K key[1] = {alive ? order_key(v) : K(0)}, kmax[1]; int xw[1];
            group_argmax<Q, K, 1, false>(key, x, gid, S.red1, phase, kmax, xw);
            if (x == xw[0]) {
                alive = false;
                S.sval[d * nm + r] = v; S.ssym[d * nm + r] = x;
                S.ssyn[d * nm + r] = g.pfwd[(size_t)(e0 + d) * Q + x];
            }
        }
    }
    Grp<Q>::sync(gid);
    const int nj = nc + 1;
    for (int d = 0; d < dcn; ++d) {
        // -- conf(q,1): one input l != d takes any symbol, the others their best one --------------------
        T b1 = NINF;
        for (int l = 0; l < dcn; ++l) {
            if (l == d) continue;
            int rest = x;                                        // syndrome the deviating input must supply
            for (int m = 0; m < dcn; ++m) if (m != d && m != l) rest ^= S.ssyn[m * nm];
            const int y = g.pinv[(size_t)(e0 + l) * Q + rest];
            T sum = T(0);
            for (int m = 0; m < dcn; ++m) {
                if (m == d) continue;
                sum = sum + (m == l ? S.V[l * Q + y] : S.sval[m * nm]);
            }
            b1 = sum > b1 ? sum : b1;
        }
        // -- conf(n_m, n_c): W[j][s] over the inputs l != d, j = number of non-best picks ---------------
        T *Wa = S.W, *Wb = S.W + (size_t)nj * Q;
        for (int j = 0; j < nj; ++j) Wa[j * Q + x] = (j == 0 && x == 0) ? T(0) : NINF;
        for (int l = 0; l < dcn; ++l) {
            if (l == d) continue;
            Grp<Q>::sync(gid);
            for (int j = 0; j < nj; ++j) {
                T w = Wa[j * Q + (x ^ S.ssyn[l * nm])] + S.sval[l * nm];
                if (j > 0)
                    for (int k = 1; k < nm; ++k) {
                        const T c = Wa[(j - 1) * Q + (x ^ S.ssyn[l * nm + k])] + S.sval[l * nm + k];
                        w = c > w ? c : w;
                    }
                Wb[j * Q + x] = w;
            }
            T *t = Wa; Wa = Wb; Wb = t;
        }
        T bst = b1;
        for (int j = 0; j < nj; ++j) { const T w = Wa[j * Q + x]; bst = w > bst ? w : bst; }
        S.best[x] = bst;
        Grp<Q>::sync(gid);
        // -- output (NBLDPC.cpp:899-916) -----------------------------------------------------------------
        T out = T(0);
        if (x != 0) out = soft_threshold<T>(S.best[g.pfwd[(size_t)(e0 + d) * Q + x]] - S.best[0], factor, offset);
        c2v_out[(size_t)(e0 + d) * Q + x] = out;
        Grp<Q>::sync(gid);                                      // best / W are rewritten for the next output
    }
}

}  // namespace nbl
