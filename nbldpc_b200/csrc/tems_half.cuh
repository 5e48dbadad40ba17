// Fused iteration kernel for T-EMS (n_c <= 2) over GF(16): four lanes per (frame, check row), four field elements per
// lane, eight check rows per warp -- the layout of qspa_half.cuh, whose row set-up and variable-node phase it shares.
//
// The GF(16) codes of the reference have check degree 5, one more than the warp kernel of tems_warp.cuh serves, and a
// 16-element vector is half a warp; they ran on the generic group kernel (cn_tems.cuh).  At q = 16 the configuration
// search is tiny -- deltaW2[s] = min over a != 0 of m0(a) + D_{c0(a)}[s ^ a] (see tems_warp.cuh) is 15 candidates per
// output -- so the kernel is the variable-node phase plus bookkeeping, all of it per lane on four elements:
//   * delta domain: dU[d][eta] = max_d - V_d[h^-1 eta ^ decide(v2c_d)] (the gather index is linear over XOR), gathered
//     through the warp's shared-memory slab;
//   * column order per eta over up to 8 columns (cheapest / second cheapest column, stable: the lower column wins ties);
//   * the D_c vectors, m0 and c0 of the row go to shared memory, the 15 x 4 candidates of a lane are scalar loads;
//   * recorded path, output trellis with a shared-memory atomic min for the two entries of a pair path, fill-in,
//     normal domain and correction as in tems_warp.cuh.
// Ties between equal costs: single deviation before pair, smallest a among pairs (the FAST order of the warp kernel;
// NBL_TIES_REFERENCE / BITEXACT stay on the group kernel).  Serves d_c <= 8, d_v <= 4, any n_r, n_c <= 2.
#pragma once
#include "qspa_half.cuh"

namespace nbl {

template <typename T> struct ThSmem {
    // per row: V [16] (gather buffer), D_c [8][16] (later the output trellis), m0 [16], m1 [16], c0 | c1 << 4 [16] bytes
    //          BS-TEMS: deltaW [16], basic-set costs [16]; recorded subsets [16] x 8 bytes; basic-set eta / column [16] bytes each
    static constexpr size_t row_elems = 16 + 8 * 16 + 16 + 16 + 16 + 16;
    static constexpr size_t row_bytes = row_elems * sizeof(T) + 16 + 16 * 8 + 32;
    static constexpr size_t bytes_per_warp = row_bytes * QH_ROWS;
    static constexpr size_t bytes = bytes_per_warp * QH_WARPS;
};

// BS == 1: BS-TEMS (decode method 7): the pair search is replaced by the basic-set construction and the walk over its
// subsets, as in tems_warp.cuh; here lane 0 of every row walks its own row, eight rows of a warp at a time.
template <typename T, int BS = 0>
__global__ void __launch_bounds__(32 * QH_WARPS, sizeof(T) == 4 ? 4 : 2) k_iterate_tems_half(const IterParams<T> P, const int *__restrict__ desc) {
    constexpr int Q = 16;
    extern __shared__ __align__(16) unsigned char th_raw[];
    const DevGraph &g = P.g;
    QhRow<T> R;
    if (!qh_setup<T>(P, R)) return;
    const int warp = threadIdx.x >> 5, lane = R.lane, l4 = R.l4, x0 = R.x0, f = R.f, e0 = R.e0, dcn = R.dcn, dmax = R.dmax;
    const T INF = Real<T>::inf();
    T *sV = reinterpret_cast<T *>(th_raw + ThSmem<T>::bytes_per_warp * warp + ThSmem<T>::row_bytes * R.rw);
    T *sD = sV + 16;                                                // [8][16]
    T *sM0 = sD + 8 * 16, *sM1 = sM0 + 16;
    T *sDW = sM1 + 16, *sBL = sDW + 16;                             // BS-TEMS: deltaW [16], costs of the basic set [16]
    uint8_t *sC = reinterpret_cast<uint8_t *>(sBL + 16);            // [16]
    unsigned long long *sSet = reinterpret_cast<unsigned long long *>(sC + 16);   // [16] byte d: basic element + 1 deviating in column d
    uint8_t *sBQ = reinterpret_cast<uint8_t *>(sSet + 16), *sBC = sBQ + 16;       // basic set: eta, column
    T V[QH_MAXDC][4];
    unsigned tf[QH_MAXDC], ti[QH_MAXDC];                            // my four bytes of the rows h * x and h^-1 * x
    unsigned sdw;                                                   // nibble d: decide(v2c_d)
    qh_vn_phase<T>(P, desc, R, V, tf, ti, sdw);

    // ---- beta_d = h_d * decide(v2c_d), syndrome of the most reliable symbols (TEMS_Get_Beta, NBLDPC.cpp:1789-1812) ------------
    unsigned betaw = 0u; int syn = 0;                               // nibble d: beta_d
#pragma unroll
    for (int d = 0; d < QH_MAXDC; ++d) {
        if (d < dmax) {
            const int sd = (int)((sdw >> (4 * d)) & 15u);
            const unsigned hw = __shfl_sync(QH_FULL, tf[d], (lane & ~3) + (sd >> 2));
            const int beta = d < dcn ? (int)((hw >> (8 * (sd & 3))) & 255u) : 0;
            betaw |= (unsigned)beta << (4 * d);
            syn ^= beta;
        }
    }
    // ---- delta domain (TEMS_Get_deltaU, :1814-1834): dU[d][eta] = max_d - V_d[h^-1 eta ^ decide] ---------------------------------
    T u[QH_MAXDC][4];
#pragma unroll
    for (int d = 0; d < QH_MAXDC; ++d) {
#pragma unroll
        for (int i = 0; i < 4; ++i) u[d][i] = INF;
        if (d < dmax) {
            qh_store4<T>(sV + x0, V[d]);
            __syncwarp();
            const int sd = (int)((sdw >> (4 * d)) & 15u);
            const T vmax = sV[sd];
            T t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] = vmax - sV[((ti[d] >> (8 * i)) & 15u) ^ (unsigned)sd];
            __syncwarp();
            if (d < dcn) {
#pragma unroll
                for (int i = 0; i < 4; ++i) u[d][i] = t[i];
            }
        }
    }
    // ---- column order per eta (TEMS_Get_Min, :1836-1890): cheapest and second cheapest column, lower column first ----------------
    T m0[4], m1[4]; int k0[4], k1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m0[i] = INF; m1[i] = INF; k0[i] = 0; k1[i] = 0;
#pragma unroll
        for (int d = 0; d < QH_MAXDC; ++d) {
            const T v = u[d][i];
            if (v < m0[i]) { m1[i] = m0[i]; k1[i] = k0[i]; m0[i] = v; k0[i] = d; }
            else if (v < m1[i]) { m1[i] = v; k1[i] = d; }
        }
    }
    {
        const T second_ok = P.p0 >= 2 ? T(0) : INF;                 // n_r = 1: only the cheapest column may deviate
        uint8_t cw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) cw[i] = (uint8_t)(k0[i] | (k1[i] << 4));
        *reinterpret_cast<unsigned *>(sC + x0) = (unsigned)cw[0] | ((unsigned)cw[1] << 8) | ((unsigned)cw[2] << 16) | ((unsigned)cw[3] << 24);
        T rc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rc[i] = (x0 + i == 0) ? INF : m0[i];
        qh_store4<T>(sM0 + x0, rc);
        qh_store4<T>(sM1 + x0, m1);
#pragma unroll
        for (int c = 0; c < QH_MAXDC; ++c) {
            if (c < dmax) {                                         // D_c[x]: cheapest allowed column of x other than c
                T dv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) dv[i] = (x0 + i == 0) ? INF : (k0[i] == c ? m1[i] + second_ok : m0[i]);
                qh_store4<T>(sD + c * 16 + x0, dv);
            }
        }
    }
    __syncwarp();
    T *Tout = sD;
    if constexpr (!BS) {
        // ---- pair search: best[s] = min over a != 0 of m0(a) + D_{c0(a)}[s ^ a]; the first (smallest) a that attains it ---------------
        T best[4]; int arg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { best[i] = INF; arg[i] = 0; }
        if (P.p1 >= 2) {
#pragma unroll 1
            for (int a = 1; a < Q; ++a) {
                const T ma = sM0[a];
                const T *Da = sD + (sC[a] & 15) * 16;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const T cand = ma + Da[(x0 + i) ^ a];
                    if (cand < best[i]) { best[i] = cand; arg[i] = a; }
                }
            }
        }
        // ---- deltaW and the recorded path of each owned s ---------------------------------------------------------------------------
        T dw[4], va[4], vb[4]; int pa[4], pb[4], pca[4], pcb[4]; bool pair[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = x0 + i;
            const T single = (s == 0) ? T(0) : m0[i];
            pair[i] = best[i] < single;                                 // never for s == 0
            dw[i] = pair[i] ? best[i] : single;
            pa[i] = arg[i]; pb[i] = s ^ arg[i];
            pca[i] = pcb[i] = 0; va[i] = vb[i] = T(0);
            if (pair[i]) {
                const int ca = sC[pa[i]] & 15, kb = sC[pb[i]];
                const int cb = ((kb & 15) != ca) ? (kb & 15) : (kb >> 4);
                pca[i] = ca; pcb[i] = cb;
                va[i] = dw[i] - sM0[pa[i]];                             // dU[ca][a]
                vb[i] = dw[i] - sD[ca * 16 + pb[i]];                    // dU[cb][b]
            }
        }
        __syncwarp();                                                  // everyone is done with D
        // ---- output trellis Tout[d][t] (NBLDPC.cpp:1084-1094): own index first, then the entries of pair paths -----------------------
#pragma unroll
        for (int d = 0; d < QH_MAXDC; ++d) {
            if (d < dmax) {
                T tv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int s = x0 + i;
                    const bool uses = s != 0 && (pair[i] ? (pca[i] == d || pcb[i] == d) : (k0[i] == d));
                    tv[i] = uses ? INF : dw[i];
                }
                qh_store4<T>(Tout + d * 16 + x0, tv);
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (pair[i]) {                                              // (values are >= 0: ordered like their bits)
                atomic_min_nonneg(Tout + pca[i] * 16 + pb[i], va[i]);
                atomic_min_nonneg(Tout + pcb[i] * 16 + pa[i], vb[i]);
            }
        }
        __syncwarp();
    } else {
        // ---- BS-TEMS (Decoding_BS_TEMS, NBLDPC.cpp:1145-1396): the all-zero configuration plus the subsets of the basic set ----------
        // (a) Construct_BasicSet (:2023-2101): the n_m cheapest eta != 0 by m0, lowest eta among equals; for n_m <= log2 q
        //     greedily linearly independent over GF(2) (span = 16-bit mask of the row, the same in its four lanes)
        const int nm = P.p0, nc = P.p1;
        T rc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rc[i] = (x0 + i == 0) ? INF : m0[i];
        unsigned avail = 15u, span = 0u;
        int found = 0;
#pragma unroll 1
        for (int k = 0; k < nm; ++k) {                              // warp-uniform trip count; a row that ran out keeps found
            T lm = INF; int lx = 99;
#pragma unroll
            for (int i = 3; i >= 0; --i) if (((avail >> i) & 1u) && rc[i] <= lm && rc[i] < INF) { lm = rc[i]; lx = x0 + i; }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const T om = __shfl_xor_sync(QH_FULL, lm, o);
                const int ox = __shfl_xor_sync(QH_FULL, lx, o);
                if (om < lm || (om == lm && ox < lx)) { lm = om; lx = ox; }
            }
            if (lx < 16 && found == k) {                            // (found == k: the row has not run out before)
                const int w = lx;
                if (l4 == 0) { sBQ[k] = (uint8_t)w; sBC[k] = (uint8_t)(sC[w] & 15); sBL[k] = lm; }
                found = k + 1;
                if (nm > g.p) { if ((w >> 2) == l4) avail &= ~(1u << (w & 3)); }
                else {
                    unsigned moved = 0u;
#pragma unroll
                    for (int x = 0; x < 16; ++x) moved |= ((span >> (x ^ w)) & 1u) << x;
                    span |= moved | (1u << w);
                    avail &= ~((span >> x0) & 15u);
                }
            }
        }
        // (b) BasicSet_TEMS_ConstructConf (:2102-2147): lane 0 of the row walks the subsets (<= n_c elements, no column twice),
        //     include first, with the reference's add / subtract running sum
#pragma unroll
        for (int i = 0; i < 4; ++i) { sDW[x0 + i] = (x0 + i == 0) ? T(0) : INF; sSet[x0 + i] = 0ull; }
        __syncwarp();
        if (l4 == 0 && dcn > 0) {
            unsigned col_sel = 0u; unsigned long long cand = 0ull; int diff = 0, sym = 0, lvl = 0; T sum = T(0);
            uint8_t st[17];
            st[0] = 0;
            while (lvl >= 0) {
                if (lvl == found) {
                    if (sum < sDW[sym]) { sDW[sym] = sum; sSet[sym] = cand; }
                    --lvl;
                    continue;
                }
                const int col = sBC[lvl], bq = sBQ[lvl];
                const T bL = sBL[lvl];
                if (st[lvl] == 0) {
                    if (!((col_sel >> col) & 1u) && diff + 1 <= nc) {
                        diff += 1; col_sel |= 1u << col;
                        sym ^= bq; sum = sum + bL; cand |= (unsigned long long)(lvl + 1) << (8 * col);
                        st[lvl] = 1;
                    } else st[lvl] = 2;
                    ++lvl; st[lvl] = 0;
                } else if (st[lvl] == 1) {
                    sym ^= bq; sum = sum - bL; col_sel &= ~(1u << col); cand &= ~(255ull << (8 * col));
                    diff -= 1;
                    st[lvl] = 2;
                    ++lvl; st[lvl] = 0;
                } else --lvl;
            }
        }
        __syncwarp();
        // (c) output trellis: Tout[d][eta ^ eta_d] = min(deltaW[eta] - dU[d][eta_d]) over the reached eta
        T dw[4]; unsigned long long et[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { dw[i] = sDW[x0 + i]; et[i] = sSet[x0 + i]; }
        {
            T inf_v[4] = {INF, INF, INF, INF};
#pragma unroll
            for (int d = 0; d < QH_MAXDC; ++d) if (d < dmax) qh_store4<T>(Tout + d * 16 + x0, inf_v);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (dw[i] < INF) {
                for (int d = 0; d < dcn; ++d) {
                    const int k1 = (int)((et[i] >> (8 * d)) & 255ull);
                    const int dev = k1 ? sBQ[k1 - 1] : 0;
                    const T du = k1 ? sBL[k1 - 1] : T(0);
                    atomic_min_any(Tout + d * 16 + ((x0 + i) ^ dev), dw[i] - du);      // the running sum's residue can make this negative
                }
            }
        }
        __syncwarp();
    }
    // ---- fill-in (:1095-1102), normal domain, un-permute, correct (:1105-1128) -----------------------------------------------------
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + x0;
#pragma unroll
    for (int d = 0; d < QH_MAXDC; ++d) {
        if (d < dmax) {
            auto getT = [&](int t) -> T {
                T v = Tout[d * 16 + t];
                if (v == INF) { const int kk = sC[t]; v = ((kk & 15) == d) ? sM1[t] : ((t == 0) ? T(0) : sM0[t]); }
                return v;
            };
            const int bs = syn ^ (int)((betaw >> (4 * d)) & 15u);
            const T Tb = getT(bs);
            T out[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = (int)((tf[d] >> (8 * i)) & 15u) ^ bs;
                out[i] = (x0 + i == 0) ? T(0) : soft_threshold<T>(Tb - getT(t), P.factor, P.offset);
            }
            if (d < dcn) qh_store4<T>(coutf + (size_t)(e0 + d) * Q, out);
        }
    }
}

}  // namespace nbl
