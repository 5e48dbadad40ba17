// T-EMS check node (trellis EMS in the delta domain), one group of Q threads per check row.
//
// Replaces CNBLDPC::TEMS_Get_Beta / TEMS_Get_deltaU / TEMS_Get_Min / TEMS_ConstructConf and the output
// stage of Decoding_TEMS (reference NBLDPC.cpp:1055-1130, 1789-1944).  Same result, different algorithm:
//   * the reference enumerates every configuration (<= n_c deviations, each eta deviating only in its n_r
//     cheapest columns) with a recursive depth-first search, ~94 k leaves per GF(256) check;
//   * here deltaW[s] = min(single deviation, best pair) is computed by all q threads at once: for column fd,
//     thread s scans the compact list of deviations (b, dU[fd][b]) allowed in column fd and pairs each with the
//     allowed columns d2 > fd of s ^ b.  Same candidate set, so the same minimum; the winning path (two (column, eta)
//     pairs) is kept for the output stage exactly as TEMS_Eta is.  n_c >= 3 uses the column DP in cn_tems_dp().
// For n_c <= 2 the candidates are scanned in the reference's depth-first visiting order, so exactly tied path costs
// (punctured symbols, quantised LLRs) record the same path as the reference; the n_c >= 3 DP resolves ties in its own
// scan order (see DESIGN.md "exact ties").
#pragma once
#include "dev_common.cuh"
#include <cstring>
#include <type_traits>

namespace nbl {

template <typename T, int Q> struct TemsSmem {
    using K = typename Real<T>::key_t;
    T *V;        // [maxdc][Q] v2c of the row's edges; reused as Tout (per-column output trellis)
    T *dU;       // [maxdc][Q]
    T *P;        // [Q] running minimum over allowed deviations of earlier columns
    T *dW;       // [Q]
    T *lc;       // [Q] list: cost
    T *W2;       // DP buffers (n_c >= 3): [2][n_c+1][Q], else unused
    uint16_t *amask;  // [Q] bit d = eta may deviate in column d
    uint8_t *lb;      // [Q] list: eta
    uint8_t *c0, *c1; // [Q] cheapest / second cheapest column of eta
    uint8_t *pb;      // [Q] path: eta of the later deviation
    uint8_t *pcol;    // [Q] path: c_first | c_second << 4
    uint8_t *pnd;     // [Q] path: number of deviations (0,1,2)
    uint8_t *bp;      // DP back pointers [maxdc][n_c+1][Q]
    uint8_t *eta;     // DP recorded path [maxdc][Q]
    int *wtot;        // [WARPS]
    int *sdec;        // [maxdc] decide() of each input
    RedScratch<K, Grp<Q>::WARPS, 2> *red;

    __host__ __device__ static size_t bytes(int maxdc, int nc) {
        size_t b = 0;
        b += sizeof(T) * (size_t)Q * (2 * maxdc + 3);
        if (nc >= 3) b += sizeof(T) * (size_t)Q * 2 * (nc + 1);
        b = align_up(b, 8);
        b += sizeof(RedScratch<K, Grp<Q>::WARPS, 2>);
        b = align_up(b, 4);
        b += sizeof(int) * (Grp<Q>::WARPS + maxdc);
        b += sizeof(uint16_t) * Q;
        b += 6 * (size_t)Q;
        if (nc >= 3) b += (size_t)maxdc * (nc + 2) * Q;
        return align_up(b, 16);
    }
    __device__ void carve(unsigned char *base, int maxdc, int nc) {
        T *t = reinterpret_cast<T *>(base);
        V = t; t += (size_t)maxdc * Q;
        dU = t; t += (size_t)maxdc * Q;
        P = t; t += Q; dW = t; t += Q; lc = t; t += Q;
        W2 = t; if (nc >= 3) t += (size_t)2 * (nc + 1) * Q;
        unsigned char *c = reinterpret_cast<unsigned char *>(t);
        c = base + align_up((size_t)(c - base), 8);
        red = reinterpret_cast<RedScratch<K, Grp<Q>::WARPS, 2> *>(c); c += sizeof(*red);
        c = base + align_up((size_t)(c - base), 4);
        wtot = reinterpret_cast<int *>(c); c += sizeof(int) * Grp<Q>::WARPS;
        sdec = reinterpret_cast<int *>(c); c += sizeof(int) * maxdc;
        amask = reinterpret_cast<uint16_t *>(c); c += sizeof(uint16_t) * Q;
        lb = c; c += Q; c0 = c; c += Q; c1 = c; c += Q; pb = c; c += Q; pcol = c; c += Q; pnd = c; c += Q;
        bp = c; eta = c + (size_t)maxdc * (nc + 1) * Q;
    }
};

// Configuration search for n_c >= 3: column DP over (syndrome s, number of deviations j) with back pointers.
// W_d[j][s] = min cost of a prefix (eta_0..eta_d) with XOR s and j non-zero deviations; the prefix sums are
// formed left to right over the columns as the reference's recursion forms them.  Paths that repeat a
// non-zero eta (which the reference skips, NBLDPC.cpp:1910-1922) cancel in the XOR and cost >= 0 extra, so
// they never beat the path without the repeated pair.  Ends by building the output trellis Tout.
template <typename T, int Q>
__device__ void cn_tems_dp(TemsSmem<T, Q> &S, int gid, int x, int dcn, int nc, T *Tout) {
    const T INF = Real<T>::inf();
    const int nj = nc + 1;
    T *Wa = S.W2, *Wb = S.W2 + (size_t)nj * Q;
    for (int j = 0; j < nj; ++j) Wa[j * Q + x] = (j == 0 && x == 0) ? T(0) : INF;
    for (int d = 0; d < dcn; ++d) {
        Grp<Q>::sync(gid);
        const bool in = x != 0 && ((S.amask[x] >> d) & 1);
        int len;
        const int pos = group_prefix_count<Q>(in, x, gid, S.wtot, len);
        if (in) { S.lb[pos] = (uint8_t)x; S.lc[pos] = S.dU[d * Q + x]; }
        Grp<Q>::sync(gid);
        for (int j = 0; j < nj; ++j) {
            T w = Wa[j * Q + x]; int b0 = 0;                      // eta_d = 0 first, then ascending eta
            if (j > 0)
                for (int i = 0; i < len; ++i) {
                    const int b = S.lb[i];
                    const T cand = Wa[(j - 1) * Q + (x ^ b)] + S.lc[i];
                    if (cand < w) { w = cand; b0 = b; }
                }
            Wb[j * Q + x] = w;
            S.bp[((size_t)d * nj + j) * Q + x] = (uint8_t)b0;
        }
        T *t = Wa; Wa = Wb; Wb = t;
    }
    // deltaW[s] = best over the deviation counts; walk the back pointers to recover Eta[s][.]
    {
        T w = Wa[x]; int j = 0;
        for (int jj = 1; jj < nj; ++jj) { const T c = Wa[jj * Q + x]; if (c < w) { w = c; j = jj; } }
        S.dW[x] = w;
        int s = x;
        for (int d = dcn - 1; d >= 0; --d) {
            const int b = S.bp[((size_t)d * nj + j) * Q + s];
            S.eta[d * Q + x] = (uint8_t)b;
            if (b) { s ^= b; --j; }
        }
    }
    Grp<Q>::sync(gid);                                          // V is free from here on
    for (int d = 0; d < dcn; ++d) Tout[d * Q + x] = INF;
    Grp<Q>::sync(gid);
    {
        const T w = S.dW[x];
        for (int d = 0; d < dcn; ++d) {
            const int dev = S.eta[d * Q + x];
            atomic_min_nonneg(&Tout[d * Q + (x ^ dev)], w - S.dU[d * Q + dev]);
        }
    }
    Grp<Q>::sync(gid);
}

// Configuration search of tie mode BITEXACT: the reference's depth-first search itself, replayed by one thread.
// TEMS_ConstructConf (NBLDPC.cpp:1892-1944) keeps ONE running cost that it adds to on the way down and subtracts from on
// the way up, so after the first non-trivial pair of operations that sum carries a rounding residue (~1e-14 after a
// GF(256) row) which every later configuration inherits.  T-EMS copies input values into its outputs, so from the
// third iteration on rows meet configurations whose exact costs are equal (a + b against a stored a + b), and the
// reference's choice among them depends on that residue.  Reproducing it needs the same additions in the same order:
// strictly sequential, so this mode is for verification, not for throughput (one thread per check row walks the
// ~1e5 leaves; the other q - 1 threads build the per-column candidate lists before and the output trellis after).
template <typename T, int Q>
__device__ void cn_tems_exact(TemsSmem<T, Q> &S, int gid, int x, int dcn, int nc, T *Tout) {
    const T INF = Real<T>::inf();
    uint8_t *list = S.bp;                                           // [maxdc][Q]: allowed eta of column d, ascending (0 first)
    int *lens = reinterpret_cast<int *>(S.P);                       // [maxdc]
    for (int d = 0; d < dcn; ++d) {
        Grp<Q>::sync(gid);
        const bool in = x == 0 || ((S.amask[x] >> d) & 1);
        int len;
        const int pos = group_prefix_count<Q>(in, x, gid, S.wtot, len);
        if (in) list[d * Q + pos] = (uint8_t)x;
        if (x == 0) lens[d] = len;
    }
    S.dW[x] = INF;
    Grp<Q>::sync(gid);
    if (x == 0) {
        unsigned used[Q / 32 > 0 ? Q / 32 : 1];
        for (int i = 0; i < (Q + 31) / 32; ++i) used[i] = 0;
        int pos[16]; uint8_t cur[16];
        T sum = T(0); int sym = 0, diff = 0, lvl = 0;
        pos[0] = 0;
        while (lvl >= 0) {
            bool down = false;
            if (lvl == dcn) {                                       // a leaf: strict '<' keeps the first of equal costs
                if (sum < S.dW[sym]) {
                    S.dW[sym] = sum;
                    for (int d = 0; d < dcn; ++d) S.eta[d * Q + sym] = cur[d];
                }
            } else {
                const uint8_t *lst = list + lvl * Q;
                const int len = lens[lvl];
                int p = pos[lvl];
                while (p < len && ((used[lst[p] >> 5] >> (lst[p] & 31)) & 1)) ++p;
                if (p < len) {
                    const int e = lst[p];
                    if (diff + (e != 0) <= nc) {                    // else: the reference leaves this column's loop
                        if (e) { ++diff; used[e >> 5] |= 1u << (e & 31); }
                        sym ^= e;
                        sum = sum + S.dU[lvl * Q + e];
                        cur[lvl] = (uint8_t)e; pos[lvl] = p;
                        ++lvl;
                        if (lvl < dcn) pos[lvl] = 0;
                        down = true;
                    }
                }
            }
            if (!down) {                                            // back to the parent: undo its choice, next candidate
                --lvl;
                if (lvl >= 0) {
                    const int e = cur[lvl];
                    sym ^= e;
                    sum = sum - S.dU[lvl * Q + e];
                    if (e) { --diff; used[e >> 5] &= ~(1u << (e & 31)); }
                    ++pos[lvl];
                }
            }
        }
    }
    Grp<Q>::sync(gid);                                              // V is free from here on
    for (int d = 0; d < dcn; ++d) Tout[d * Q + x] = INF;
    Grp<Q>::sync(gid);
    {
        const T w = S.dW[x];
        if (w < INF)
            for (int d = 0; d < dcn; ++d) {
                const int dev = S.eta[d * Q + x];
                atomic_min_any(&Tout[d * Q + (x ^ dev)], w - S.dU[d * Q + dev]);   // the residue can make this negative
            }
    }
    Grp<Q>::sync(gid);
}

// Inputs: S.V[d][x] (final v2c, slot 0 == 0) and S.sdec[d] = decide(V_d) for d < dcn, visible to the group.
// Output: c2v_out[(e0+d)*Q + x] for d < dcn.
template <typename T, int Q>
__device__ void cn_tems(TemsSmem<T, Q> &S, const DevGraph &g, int gid, int x, int e0, int dcn, int nr, int nc, bool bitexact,
                        T factor, T offset, T *__restrict__ c2v_out /* frame base, [E][Q] */) {
    const T INF = Real<T>::inf();
    // -- beta, syndrome of the most reliable symbols (TEMS_Get_Beta) --------------------------------
    int syn = 0;
    for (int d = 0; d < dcn; ++d) syn ^= g.pfwd[(size_t)(e0 + d) * Q + S.sdec[d]];
    // -- delta domain (TEMS_Get_deltaU): dU[d][eta] = max_d - V_d[h^-1 (eta ^ beta_d)] -----------------
    for (int d = 0; d < dcn; ++d) {
        const int sd = S.sdec[d];
        const int beta = g.pfwd[(size_t)(e0 + d) * Q + sd];
        const T vmax = S.V[d * Q + sd];                       // V[.][0] == 0 covers "nothing positive"
        const int k = g.pinv[(size_t)(e0 + d) * Q + (x ^ beta)];
        S.dU[d * Q + x] = vmax - S.V[d * Q + k];
    }
    Grp<Q>::sync(gid);
    // -- column order per eta (TEMS_Get_Min): stable rank of each column by dU[.][x] ------------------
    {
        unsigned am = 0; int k0 = 0, k1 = 0;
        for (int d = 0; d < dcn; ++d) {
            const T ud = S.dU[d * Q + x];
            int rank = 0;
            for (int c = 0; c < dcn; ++c) {
                const T uc = S.dU[c * Q + x];
                rank += (uc < ud) || (uc == ud && c < d);
            }
            if (rank < nr) am |= 1u << d;
            if (rank == 0) k0 = d;
            if (rank == 1) k1 = d;
        }
        S.amask[x] = (uint16_t)am; S.c0[x] = (uint8_t)k0; S.c1[x] = (uint8_t)k1;
        S.P[x] = INF;
    }
    T *Tout = S.V;
    if (bitexact) {
        cn_tems_exact<T, Q>(S, gid, x, dcn, nc, Tout);
    } else if (nc <= 2) {
        // -- configuration search for n_c <= 2, in the reference's own visiting order ------------------------
        // TEMS_ConstructConf walks (eta_0, eta_1, ...) depth first, every eta_d ascending from 0, and keeps the first
        // configuration that attains the minimum (strict '<', NBLDPC.cpp:1897).  For one syndrome s that order is: the
        // LAST deviating column first ("first deviating column" fd descending), then eta_fd ascending -- the single
        // deviation (fd, s) sits at eta_fd = s --, then the column of the second deviation descending.  Scanning the
        // candidates in exactly that order with the same strict comparison reproduces the recorded path also when
        // costs tie exactly (punctured symbols, quantised LLRs).
        T best = INF; int bx = 0, bfd = 0, bd2 = -1;
        for (int fd = dcn - 1; fd >= 0; --fd) {
            Grp<Q>::sync(gid);                                      // previous scan done with the list
            const bool in = x != 0 && ((S.amask[x] >> fd) & 1);
            int len;
            const int pos = group_prefix_count<Q>(in, x, gid, S.wtot, len);
            if (in) { S.lb[pos] = (uint8_t)x; S.lc[pos] = S.dU[fd * Q + x]; }
            Grp<Q>::sync(gid);
            if (x != 0)
                for (int i = 0; i < len; ++i) {
                    const int b = S.lb[i];
                    const T ub = S.lc[i];
                    if (b == x) { if (ub < best) { best = ub; bx = b; bfd = fd; bd2 = -1; } }
                    else if (nc >= 2) {
                        const int y = x ^ b;
                        const unsigned am = S.amask[y];
                        for (int d2 = dcn - 1; d2 > fd; --d2)
                            if ((am >> d2) & 1) {
                                const T cand = ub + S.dU[d2 * Q + y];
                                if (cand < best) { best = cand; bx = b; bfd = fd; bd2 = d2; }
                            }
                    }
                }
        }
        // -- deltaW and the recorded path ------------------------------------------------------------------
        {
            T w; int nd, pc = 0, bb = 0;
            if (x == 0) { w = T(0); nd = 0; }
            else if (bd2 >= 0) { w = best; nd = 2; pc = bfd | (bd2 << 4); bb = x ^ bx; }   // a = bx in column bfd, b in column bd2
            else { w = best; nd = 1; pc = bfd; }
            S.dW[x] = w; S.pb[x] = (uint8_t)bb; S.pcol[x] = (uint8_t)pc; S.pnd[x] = (uint8_t)nd;
        }
        Grp<Q>::sync(gid);                                          // everyone is done with V and P
        // -- output trellis Tout[d][t] (NBLDPC.cpp:1084-1094): identity entries first -----------------------
        {
            const int nd = S.pnd[x], pc = S.pcol[x];
            const int ca = pc & 15, cb = pc >> 4;
            const T w = S.dW[x];
            for (int d = 0; d < dcn; ++d) {
                const bool uses = (nd == 1 && ca == d) || (nd == 2 && (ca == d || cb == d));
                Tout[d * Q + x] = (x == 0) ? T(0) : (uses ? INF : w);
            }
        }
        Grp<Q>::sync(gid);
        if (S.pnd[x] == 2) {                                        // entries that land on another index
            const int pc = S.pcol[x], ca = pc & 15, cb = pc >> 4, b = S.pb[x], a = x ^ b;
            const T w = S.dW[x];
            atomic_min_nonneg(&Tout[ca * Q + b], w - S.dU[ca * Q + a]);
            atomic_min_nonneg(&Tout[cb * Q + a], w - S.dU[cb * Q + b]);
        }
        Grp<Q>::sync(gid);
    } else {
        cn_tems_dp<T, Q>(S, gid, x, dcn, nc, Tout);
    }
    // -- fill-in (NBLDPC.cpp:1095-1102), back to the normal domain, un-permute, correct (1105-1128) -----
    for (int d = 0; d < dcn; ++d) {
        const int e = e0 + d;
        const int bs = syn ^ g.pfwd[(size_t)e * Q + S.sdec[d]];
        auto getT = [&](int t) -> T {
            T v = Tout[d * Q + t];
            if (v == INF) { const int k0 = S.c0[t], k1 = S.c1[t]; v = (d == k0) ? S.dU[k1 * Q + t] : S.dU[k0 * Q + t]; }
            return v;
        };
        T out = T(0);
        if (x != 0) {
            const int t = g.pfwd[(size_t)e * Q + x] ^ bs;
            out = soft_threshold<T>(getT(bs) - getT(t), factor, offset);
        }
        c2v_out[(size_t)e * Q + x] = out;
    }
}


// ---- BS-TEMS check node (CNBLDPC::Decoding_BS_TEMS, NBLDPC.cpp:1145-1396; Construct_BasicSet :2023-2101;
//      BasicSet_TEMS_ConstructConf :2102-2147) ----------------------------------------------------------------------------
// What the reference computes in this mode (TEMS_Nr is never set there, so the T-EMS stage only leaves the all-zero
// configuration): the "basic set" = the n_m cheapest eta by m0(eta) = dU[c0(eta)][eta] -- for n_m <= log2 q chosen
// greedily so that they are linearly independent over GF(2) --, then every subset of it with at most n_c elements and
// no column used twice is a configuration; syndromes no subset reaches keep deltaW = +inf and take the fill-in value.
// The subset enumeration is tiny (2^n_m leaves) and its running sum is formed by add / subtract exactly in the
// reference's order, so one thread walks it with the same recursion; everything around it is data parallel.
template <typename T, int Q> struct BsWalk {
    // BasicSet_TEMS_ConstructConf (NBLDPC.cpp:2102-2147) without the recursion: element `lvl` of the basic set is first
    // included (if its column is free and the deviation budget allows), then excluded; the running sum is added to on the
    // way down and subtracted from on the way back exactly as the reference does, so its rounding residue is the same.
    // st[lvl]: 0 = not visited yet, 1 = inside the "included" subtree, 2 = inside the "excluded" subtree.
    TemsSmem<T, Q> &S;
    int nm, nc, dcn;
    __device__ void walk() {
        unsigned col_sel = 0; int diff = 0, sum_sym = 0; T sum = T(0);
        uint8_t cand[16], st[17];
        for (int d = 0; d < 16; ++d) cand[d] = 0;
        int lvl = 0;
        st[0] = 0;
        while (lvl >= 0) {
            if (lvl == nm) {                                        // a leaf: strict '<' keeps the first of equal costs
                if (sum < S.dW[sum_sym]) {
                    S.dW[sum_sym] = sum;
                    for (int d = 0; d < dcn; ++d) S.eta[d * Q + sum_sym] = cand[d];
                }
                --lvl;
                continue;
            }
            const int col = S.pb[lvl], bq = S.lb[lvl];
            const T bL = S.lc[lvl];
            if (st[lvl] == 0) {
                if (!((col_sel >> col) & 1)) {
                    if (diff + 1 <= nc) {                           // include first
                        diff += 1; col_sel |= 1u << col;
                        sum_sym ^= bq; sum = sum + bL; cand[col] = (uint8_t)bq;
                        st[lvl] = 1;
                    } else { cand[col] = 0; st[lvl] = 2; }          // budget spent: excluded subtree only
                } else st[lvl] = 2;                                 // column already deviates: this element cannot be added
                ++lvl; st[lvl] = 0;
            } else if (st[lvl] == 1) {                              // back from the included subtree: undo, then exclude
                sum_sym ^= bq; sum = sum - bL; col_sel &= ~(1u << col);
                diff -= 1; cand[col] = 0;
                st[lvl] = 2;
                ++lvl; st[lvl] = 0;
            } else --lvl;
        }
    }
};

// Inputs as cn_tems.  nm = basic_set_Nm (1 .. min(q-1, 16)), nc = basic_set_Nc.
template <typename T, int Q>
__device__ void cn_bstems(TemsSmem<T, Q> &S, const DevGraph &g, int gid, int x, int e0, int dcn, int nm, int nc,
                          T factor, T offset, int &phase, T *__restrict__ c2v_out) {
    using K = typename Real<T>::key_t;
    const T INF = Real<T>::inf();
    int syn = 0;
    for (int d = 0; d < dcn; ++d) syn ^= g.pfwd[(size_t)(e0 + d) * Q + S.sdec[d]];
    for (int d = 0; d < dcn; ++d) {                                 // TEMS_Get_deltaU
        const int sd = S.sdec[d];
        const int beta = g.pfwd[(size_t)(e0 + d) * Q + sd];
        const T vmax = S.V[d * Q + sd];
        const int k = g.pinv[(size_t)(e0 + d) * Q + (x ^ beta)];
        S.dU[d * Q + x] = vmax - S.V[d * Q + k];
    }
    Grp<Q>::sync(gid);
    T m0;
    {   // TEMS_Get_Min: stable order of the columns by dU[.][x]
        int k0 = 0, k1 = 0;
        for (int d = 0; d < dcn; ++d) {
            const T ud = S.dU[d * Q + x];
            int rank = 0;
            for (int c = 0; c < dcn; ++c) { const T uc = S.dU[c * Q + x]; rank += (uc < ud) || (uc == ud && c < d); }
            if (rank == 0) k0 = d;
            if (rank == 1) k1 = d;
        }
        S.c0[x] = (uint8_t)k0; S.c1[x] = (uint8_t)k1;
        m0 = S.dU[k0 * Q + x];
        S.dW[x] = x == 0 ? T(0) : INF;                              // the T-EMS stage of this mode: all-zero configuration only
        for (int d = 0; d < dcn; ++d) S.eta[d * Q + x] = 0;
        S.bp[x] = 0;                                                // span flags of the greedy basis
    }
    Grp<Q>::sync(gid);
    // Construct_BasicSet: the nm cheapest eta != 0 (ascending m0), linearly independent when nm <= log2 q
    bool avail = x != 0;
    int found = 0;
    for (int k = 0; k < nm; ++k) {
        K key[2] = {avail ? (K)~order_key(m0) : K(0), K(0)}, kmax[2]; int xw[2];
        group_argmax<Q, K, 2, true>(key, x, gid, S.red, phase, kmax, xw);
        if (kmax[0] == K(0)) break;                                 // nothing left (nm > number of candidates)
        found = k + 1;
        const int w = xw[0];
        if (x == 0) { S.lb[k] = (uint8_t)w; S.pb[k] = S.c0[w]; S.lc[k] = S.dU[(size_t)S.c0[w] * Q + w]; }
        if (nm > g.p) avail = avail && x != w;
        else {
            const int in = S.bp[x] | S.bp[x ^ w] | (x == w);        // span of the chosen elements (without 0)
            Grp<Q>::sync(gid);
            S.bp[x] = (uint8_t)in;
            avail = avail && !in;
            Grp<Q>::sync(gid);
        }
    }
    Grp<Q>::sync(gid);
    if (x == 0) {
        BsWalk<T, Q> wk{S, found, nc, dcn};
        wk.walk();
    }
    Grp<Q>::sync(gid);                                              // deltaW / Eta complete; V is free from here on
    T *Tout = S.V;
    for (int d = 0; d < dcn; ++d) Tout[d * Q + x] = INF;
    Grp<Q>::sync(gid);
    {
        const T w = S.dW[x];
        if (w < INF)
            for (int d = 0; d < dcn; ++d) {
                const int dev = S.eta[d * Q + x];
                atomic_min_any(&Tout[d * Q + (x ^ dev)], w - S.dU[d * Q + dev]);
            }
    }
    Grp<Q>::sync(gid);
    for (int d = 0; d < dcn; ++d) {                                 // fill-in, normal domain, un-permute, correct
        const int e = e0 + d;
        const int bs = syn ^ g.pfwd[(size_t)e * Q + S.sdec[d]];
        auto getT = [&](int t) -> T {
            T v = Tout[d * Q + t];
            if (v == INF) { const int k0 = S.c0[t], k1 = S.c1[t]; v = (d == k0) ? S.dU[k1 * Q + t] : S.dU[k0 * Q + t]; }
            return v;
        };
        T out = T(0);
        if (x != 0) {
            const int t = g.pfwd[(size_t)e * Q + x] ^ bs;
            out = soft_threshold<T>(getT(bs) - getT(t), factor, offset);
        }
        c2v_out[(size_t)e * Q + x] = out;
    }
}

}  // namespace nbl
