// Fast path of the fused iteration kernel for T-EMS with n_c <= 2 (the headline configuration):
// ONE WARP per (frame, check row), VEC = Q/32 consecutive field elements per lane.
//
// Same results as the generic group kernel (decoder.cu / cn_tems.cuh), organised for the issue-bound search:
//   * no block-level barriers: every exchange between lanes is a __syncwarp on the warp's private smem slab;
//   * the pair search uses the structure of the allowed sets.  Let c0(x), c1(x) be the cheapest / second cheapest
//     column of eta = x and m0(x) <= m1(x) their costs (TEMS_Min).  For a pair (a, b) with a ^ b = s the cheapest
//     valid assignment of two DIFFERENT columns is m0(a) + m0(b) when c0(a) != c0(b), else
//     min(m0(a) + m1(b), m1(a) + m0(b)); columns beyond the second never win (c1(b) != c0(b) is always available).
//     So with  D_c[x] = (c0(x) == c ? m1(x) : m0(x))  -- the cheapest column of x other than c --
//         deltaW2[s] = min over a != 0 of  m0(a) + D_{c0(a)}[s ^ a] ,
//     255 x q leaves instead of the ~94 k configurations the reference enumerates (or the 98 k of a per-column
//     scan), every leaf useful.  (n_r = 1: D_c[x] = +inf where c0(x) == c.  n_r >= 2 all give this minimum.)
//   * it is register tiled: lane t owns s = VEC*t .. VEC*t+VEC-1; entry a = VEC*bh + bl loads the VEC values
//     D_c[VEC*(t^bh) ..] with conflict-free 128-bit loads and serves all VEC outputs from registers:
//         best[r] = min(best[r], d[r ^ bl] + m0(a))          -> 1 FADD + 1 FMNMX per leaf, no branches.
//   * arg-min bookkeeping is done per eta-block (2 instructions per output per block: the running minima alternate
//     between two register sets, so "did this block improve it" is one compare); the winning block is re-scanned
//     once at the end to recover the exact pair of the recorded path.
//   * everything outside the search is written as short loops over the row's edges with 128-bit shared-memory
//     accesses, to keep the kernel inside the instruction cache and free of bank conflicts.
#pragma once
#include "dev_common.cuh"
#include <type_traits>

namespace nbl {

constexpr int TW_DC = 4;      // maximum check degree served by the fast path
constexpr int TW_WARPS = 4;   // warps (= check rows) per CTA
constexpr int TW_DSC = 8;     // ints per edge in the row descriptor: variable, degree, up to 5 edge ids, 1 spare
                              // (spare of edge 0 = first edge id of the row, spare of edge 1 = row degree)

template <typename T, int Q> struct TwSmem {
    // per-warp slab:  R1 [TW_DC][Q]   v2c (VN stage) -> D_c, split layout (search) -> output trellis Tout
    //                 M0 [Q]          m0(x), +inf at x = 0 (uniform operand of the search)
    //                 bytes c01[Q] (c0 | c1 << 4);  ints meta[2*TW_DC] (decide, beta per edge)
    static constexpr size_t t_elems = (size_t)Q * (TW_DC + 1);
    static constexpr size_t bytes_per_warp = ((t_elems * sizeof(T) + (size_t)Q + 2 * TW_DC * sizeof(int) + 15) / 16) * 16;
    static constexpr size_t bytes = bytes_per_warp * TW_WARPS;
};

// P is stored so that a lane's VEC values come out of conflict-free 16-byte loads: the vector of block k is cut
// into 16-byte pieces and piece h of all blocks is contiguous:  index(k, j) = (j / E16) * (32 * E16) + k * E16 + j % E16
// with E16 = elements per 16 bytes (4 floats / 2 doubles), or VEC when the whole block is <= 16 bytes.
template <typename T, int VEC> struct PLayout {
    static constexpr int E16 = (int)(16 / sizeof(T)) < VEC ? (int)(16 / sizeof(T)) : VEC;
    __device__ static int index(int k, int j) { return (j / E16) * (32 * E16) + k * E16 + (j % E16); }
};

template <typename T, int N> __device__ __forceinline__ void ld_vec(const T *__restrict__ p, T (&v)[N]) {
    constexpr int E16 = (int)(16 / sizeof(T));
    if constexpr (N >= E16) {
#pragma unroll
        for (int i = 0; i < N / E16; ++i) {
            const int4 w = *reinterpret_cast<const int4 *>(p + i * E16);
            T tmp[E16];
            *reinterpret_cast<int4 *>(tmp) = w;
#pragma unroll
            for (int k = 0; k < E16; ++k) v[i * E16 + k] = tmp[k];
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = p[i];
    }
}
template <typename T, int N> __device__ __forceinline__ void st_vec(T *__restrict__ p, const T (&v)[N]) {
    constexpr int E16 = (int)(16 / sizeof(T));
    if constexpr (N >= E16) {
#pragma unroll
        for (int i = 0; i < N / E16; ++i) {
            T tmp[E16];
#pragma unroll
            for (int k = 0; k < E16; ++k) tmp[k] = v[i * E16 + k];
            *reinterpret_cast<int4 *>(p + i * E16) = *reinterpret_cast<int4 *>(tmp);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = v[i];
    }
}
// VEC consecutive table bytes starting at a VEC-aligned offset, as one little-endian word
template <int VEC> __device__ __forceinline__ unsigned long long ld_bytes(const uint8_t *__restrict__ p) {
    if constexpr (VEC == 8) return *reinterpret_cast<const unsigned long long *>(p);
    else if constexpr (VEC == 4) return *reinterpret_cast<const unsigned *>(p);
    else if constexpr (VEC == 2) return *reinterpret_cast<const unsigned short *>(p);
    else return *p;
}
template <int VEC> __device__ __forceinline__ void st_bytes(uint8_t *p, unsigned long long w) {
    if constexpr (VEC == 8) *reinterpret_cast<unsigned long long *>(p) = w;
    else if constexpr (VEC == 4) *reinterpret_cast<unsigned *>(p) = (unsigned)w;
    else if constexpr (VEC == 2) *reinterpret_cast<unsigned short *>(p) = (unsigned short)w;
    else *p = (uint8_t)w;
}
__device__ __forceinline__ int byte_of(unsigned long long w, int i) { return (int)((w >> (8 * i)) & 255ull); }

// DecideLLRVector over a vector spread VEC-per-lane across the warp (x = VEC*lane + j)
template <typename T, int VEC> __device__ __forceinline__ int warp_decide(const T (&v)[VEC], int lane) {
    using K = typename Real<T>::key_t;
    K bk = 0; int bj = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) { const K k = pos_key<T>(v[j]); if (k > bk) { bk = k; bj = j; } }
    K kmax; int src;
    warp_argmax<true>(bk, 0xffffffffu, kmax, src);
    const int x = __shfl_sync(0xffffffffu, VEC * lane + bj, src);
    return kmax == K(0) ? 0 : x;
}

// Recover the recorded two-deviation path of output s = VEC*lane + R from the winning eta-block `bh` by re-scanning its
// VEC entries in the search order.  Returns a | b << 8 | c_a << 16 | c_b << 20 and the two deviation costs
// dU[c_a][a] (= m0(a)) and dU[c_b][b] (= D_{c_a}[b]).
template <typename T, int Q, int R>
__device__ __forceinline__ unsigned tw_recover_path(const T *__restrict__ sD, const T *__restrict__ sM0,
                                                    const uint8_t *__restrict__ c01, int lane, T best, int bh, T &ua, T &ub) {
    constexpr int VEC = Q / 32;
    using PL = PLayout<T, VEC>;
    const unsigned long long cw = ld_bytes<VEC>(c01 + VEC * bh);
    T cost[VEC];
    ld_vec<T, VEC>(sM0 + VEC * bh, cost);
    const int k = lane ^ bh;
    int bl0 = 0; ua = T(0); ub = T(0);
#pragma unroll
    for (int bl = VEC - 1; bl >= 0; --bl) {                         // descending: the smallest matching eta remains
        const T dv = sD[(size_t)(byte_of(cw, bl) & 15) * Q + PL::index(k, R ^ bl)];
        if (dv + cost[bl] == best) { bl0 = bl; ua = cost[bl]; ub = dv; }
    }
    const int aa = VEC * bh + bl0, bb = (VEC * lane + R) ^ aa, ca = byte_of(cw, bl0) & 15, kb = c01[bb];
    const int cb = ((kb & 15) != ca) ? (kb & 15) : (kb >> 4);
    return (unsigned)aa | ((unsigned)bb << 8) | ((unsigned)ca << 16) | ((unsigned)cb << 20);
}

template <typename T, int Q>
__global__ void __launch_bounds__(32 * TW_WARPS) k_iterate_tems_warp(const IterParams<T> P) {
    constexpr int VEC = Q / 32;
    using PL = PLayout<T, VEC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevGraph &g = P.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * TW_WARPS + warp;
    if (w >= (long long)P.B * g.M) return;
    const int f = (int)(w / g.M), row = (int)(w % g.M);
    if (P.early_stop && P.done[f]) return;

    T *R1 = reinterpret_cast<T *>(smem_raw + TwSmem<T, Q>::bytes_per_warp * warp);
    T *sM0 = R1 + (size_t)TW_DC * Q;
    uint8_t *c01 = reinterpret_cast<uint8_t *>(sM0 + Q);
    int *meta = reinterpret_cast<int *>(c01 + Q);                   // [0..TW_DC) decide(v2c_d), [TW_DC..) beta_d
    const T INF = Real<T>::inf();
    const unsigned FULL = 0xffffffffu;

    // One coalesced 128-byte load brings the whole row descriptor (per edge: variable, its degree, the edge ids of
    // its messages in summation order), so the message loads below do not wait on a chain of index loads.
    const int dsc = g.rowdesc[(size_t)row * 32 + lane];
    const int e0 = __shfl_sync(FULL, dsc, TW_DSC - 1), dcn = __shfl_sync(FULL, dsc, 2 * TW_DSC - 1);
    const size_t fE = (size_t)f * g.E, fN = (size_t)f * g.N;
    const int x0 = VEC * lane;
    int syn_dec = 0, syn = 0;
    // ---- steps A-C for the row's variables (NBLDPC.cpp:977-1052) ---------------------------------------
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        const int e = e0 + d;
        const int col = __shfl_sync(FULL, dsc, TW_DSC * d), nv = __shfl_sync(FULL, dsc, TW_DSC * d + 1);
        T post[VEC], vnew[VEC];
        ld_vec<T, VEC>(P.lch + (fN + col) * Q + x0, post);
        if (!P.first) {
            constexpr int HOIST = 3;                               // messages loaded before the first add
            T m[HOIST][VEC]; int ee[HOIST];
#pragma unroll
            for (int j = 0; j < HOIST; ++j) {
                ee[j] = __shfl_sync(FULL, dsc, TW_DSC * d + 2 + j);
                if (j < nv) ld_vec<T, VEC>(P.c2v_in + (fE + ee[j]) * Q + x0, m[j]);
            }
            const int before = (int)P.vdec[fE + e];
            T own[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) own[j] = T(0);
#pragma unroll
            for (int j = 0; j < HOIST; ++j) {
                if (j < nv) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) { post[i] = post[i] + m[j][i]; if (ee[j] == e) own[i] = m[j][i]; }
                }
            }
#pragma unroll 1
            for (int j = HOIST; j < nv; ++j) {                      // variables of degree > HOIST
                const int ej = __shfl_sync(FULL, dsc, TW_DSC * d + 2 + j);
                T mj[VEC];
                ld_vec<T, VEC>(P.c2v_in + (fE + ej) * Q + x0, mj);
#pragma unroll
                for (int i = 0; i < VEC; ++i) { post[i] = post[i] + mj[i]; if (ej == e) own[i] = mj[i]; }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) vnew[j] = post[j] - own[j];
            const int dpost = warp_decide<T, VEC>(post, lane);
            int dfin = warp_decide<T, VEC>(vnew, lane);
            if (P.w_old != T(0) && dfin != before) {                 // warp-uniform
                T vold[VEC];
                ld_vec<T, VEC>(P.v2c + (fE + e) * Q + x0, vold);
#pragma unroll
                for (int j = 0; j < VEC; ++j) vnew[j] = P.w_old * vold[j] + P.w_new * vnew[j];   // -fmad=false
                dfin = warp_decide<T, VEC>(vnew, lane);
            }
            st_vec<T, VEC>(P.v2c + (fE + e) * Q + x0, vnew);
            st_vec<T, VEC>(R1 + d * Q + x0, vnew);
            const int bt = g.pfwd[(size_t)e * Q + dfin];
            if (lane == 0) { P.vdec[fE + e] = (uint8_t)dfin; P.dec[fN + col] = (uint8_t)dpost; meta[d] = dfin; meta[TW_DC + d] = bt; }
            syn ^= bt;
            syn_dec ^= g.pfwd[(size_t)e * Q + dpost];
        } else {                                                  // iteration 1: c2v == 0, v2c == L_ch, no damping
            const int dpost = warp_decide<T, VEC>(post, lane);
            st_vec<T, VEC>(P.v2c + (fE + e) * Q + x0, post);
            st_vec<T, VEC>(R1 + d * Q + x0, post);
            const int bt = g.pfwd[(size_t)e * Q + dpost];
            if (lane == 0) { P.vdec[fE + e] = (uint8_t)dpost; P.dec[fN + col] = (uint8_t)dpost; meta[d] = dpost; meta[TW_DC + d] = bt; }
            syn ^= bt;
            syn_dec ^= bt;
        }
    }
    if (lane == 0 && syn_dec != 0) P.flag[f] = 1;
    __syncwarp();
    // ---- delta domain (TEMS_Get_deltaU): dU[d][eta] = max_d - V_d[h^-1 (eta ^ beta_d)], own etas in registers -------
    T u[TW_DC][VEC];
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) {
        if (d < dcn) {
            const int sd = meta[d], bt = meta[TW_DC + d];
            const T vmax = R1[d * Q + sd];
            const int bh = bt / VEC, bl = bt % VEC;
            const unsigned long long kk = ld_bytes<VEC>(g.pinv + (size_t)(e0 + d) * Q + VEC * (lane ^ bh));
#pragma unroll
            for (int j = 0; j < VEC; ++j) u[d][j] = vmax - R1[d * Q + byte_of(kk, j ^ bl)];   // eta^beta = VEC*(lane^bh) + (j^bl)
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) u[d][j] = INF;
        }
    }
    __syncwarp();                                                  // all reads of v2c in R1 are done
    // ---- column order per eta (TEMS_Get_Min): c0, c1, m0, m1;  D_c for the search ---------------------------------
    T m0[VEC], m1[VEC]; unsigned long long c01w = 0;
    {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            int k0 = 0, k1 = 0;
#pragma unroll
            for (int d = 0; d < TW_DC; ++d) {
                int rank = 0;
#pragma unroll
                for (int c = 0; c < TW_DC; ++c) {
                    if (c < d) rank += (c < dcn) && (u[c][j] <= u[d][j]);           // stable: the lower column wins ties
                    if (c > d) rank += (c < dcn) && (u[c][j] < u[d][j]);
                }
                if (d < dcn) { if (rank == 0) k0 = d; if (rank == 1) k1 = d; }
            }
            c01w |= (unsigned long long)(k0 | (k1 << 4)) << (8 * j);
            m0[j] = INF; m1[j] = INF;
#pragma unroll
            for (int d = 0; d < TW_DC; ++d) { if (k0 == d) m0[j] = u[d][j]; if (k1 == d) m1[j] = u[d][j]; }
        }
        st_bytes<VEC>(c01 + x0, c01w);
        {
            T mm[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) mm[j] = (x0 + j == 0) ? INF : m0[j];
            st_vec<T, VEC>(sM0 + x0, mm);
        }
        const T second_ok = P.p0 >= 2 ? T(0) : INF;                 // n_r = 1: only the cheapest column may deviate
#pragma unroll
        for (int c = 0; c < TW_DC; ++c) {
            if (c < dcn) {                                          // D_c[x]: cheapest allowed column of x other than c
                constexpr int E16 = PL::E16;
#pragma unroll
                for (int h = 0; h < VEC / E16; ++h) {
                    T tmp[E16];
#pragma unroll
                    for (int i = 0; i < E16; ++i) {
                        const int j = h * E16 + i;
                        const T v = ((byte_of(c01w, j) & 15) == c) ? (m1[j] + second_ok) : m0[j];
                        tmp[i] = (x0 + j == 0) ? INF : v;
                    }
                    st_vec<T, E16>(R1 + (size_t)c * Q + h * (32 * E16) + lane * E16, tmp);
                }
            }
        }
    }
    __syncwarp();
    // ---- pair search: best[s] = min_a m0(a) + D_{c0(a)}[s ^ a] ----------------------------------------------------------
    T best[VEC], alt[VEC]; int tag[VEC];
#pragma unroll
    for (int r = 0; r < VEC; ++r) { best[r] = INF; alt[r] = INF; tag[r] = 0; }
    if (P.p1 >= 2) {
        // one eta-block: `to` = min(`from`, candidates of the block); tag the block where it got smaller
        auto block = [&](int bh, const T (&from)[VEC], T (&to)[VEC]) {
            const unsigned long long cw = ld_bytes<VEC>(c01 + VEC * bh);           // columns of the block's entries (uniform)
            T cost[VEC];
            ld_vec<T, VEC>(sM0 + VEC * bh, cost);                                  // their costs (uniform)
            const int k = lane ^ bh;
#pragma unroll
            for (int bl = 0; bl < VEC; ++bl) {
                const T *Dc = R1 + (size_t)(byte_of(cw, bl) & 15) * Q;
                T p[VEC];
                constexpr int E16 = PL::E16;
#pragma unroll
                for (int h = 0; h < VEC / E16; ++h) {
                    T tmp[E16];
                    ld_vec<T, E16>(Dc + h * (32 * E16) + k * E16, tmp);
#pragma unroll
                    for (int i = 0; i < E16; ++i) p[h * E16 + i] = tmp[i];
                }
#pragma unroll
                for (int r = 0; r < VEC; ++r) to[r] = fmin(bl == 0 ? from[r] : to[r], p[r ^ bl] + cost[bl]);
            }
#pragma unroll
            for (int r = 0; r < VEC; ++r) if (to[r] < from[r]) tag[r] = bh;
        };
#pragma unroll 1
        for (int bh = 0; bh < 32; bh += 2) { block(bh, best, alt); block(bh + 1, alt, best); }
    }
    // ---- deltaW and the recorded path of each owned s: path = a | b << 8 | c_a << 16 | c_b << 20 | ndev << 24 ------
    T dw[VEC], ua[VEC], ub[VEC]; unsigned path[VEC];
    auto finish = [&](auto RC) {
        constexpr int r = decltype(RC)::value;
        const int s = x0 + r;
        dw[r] = m0[r]; ua[r] = T(0); ub[r] = T(0);
        path[r] = (unsigned)s | ((unsigned)byte_of(c01w, r) & 15u) << 16 | (1u << 24);
        if (s == 0) { dw[r] = T(0); path[r] = 0; }
        else if (best[r] < m0[r]) {
            dw[r] = best[r];
            path[r] = tw_recover_path<T, Q, r>(R1, sM0, c01, lane, best[r], tag[r], ua[r], ub[r]) | (2u << 24);
        }
    };
    finish(std::integral_constant<int, 0>{});
    if constexpr (VEC > 1) finish(std::integral_constant<int, 1>{});
    if constexpr (VEC > 2) { finish(std::integral_constant<int, 2>{}); finish(std::integral_constant<int, 3>{}); }
    if constexpr (VEC > 4) {
        finish(std::integral_constant<int, 4>{}); finish(std::integral_constant<int, 5>{});
        finish(std::integral_constant<int, 6>{}); finish(std::integral_constant<int, 7>{});
    }
    __syncwarp();                                                  // everyone is done with D
    // ---- output trellis Tout[d][t] (NBLDPC.cpp:1084-1094): own index first ... ----------------------------------------
    T *Tout = R1;
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        T tv[VEC];
#pragma unroll
        for (int r = 0; r < VEC; ++r) {
            const int nd = path[r] >> 24, ca = (path[r] >> 16) & 15, cb = (path[r] >> 20) & 15;
            const bool uses = (nd == 1 && ca == d) || (nd == 2 && (ca == d || cb == d));
            tv[r] = (nd == 0) ? T(0) : (uses ? INF : dw[r]);
        }
        st_vec<T, VEC>(Tout + d * Q + x0, tv);
    }
    __syncwarp();
    {   // ... then the entries of two-deviation paths, which land on another index: conflict-safe min-store
        unsigned pending = 0;                                       // bits 2r, 2r+1: the two entries of output r
#pragma unroll
        for (int r = 0; r < VEC; ++r) if ((path[r] >> 24) == 2) pending |= 3u << (2 * r);
        while (__any_sync(FULL, pending != 0)) {
#pragma unroll
            for (int r = 0; r < VEC; ++r) {
                if (pending & (3u << (2 * r))) {
                    const int a = path[r] & 255, b = (path[r] >> 8) & 255, ca = (path[r] >> 16) & 15, cb = (path[r] >> 20) & 15;
                    const T va = dw[r] - ua[r], vb = dw[r] - ub[r];
                    if (pending & (1u << (2 * r))) { if (va < Tout[ca * Q + b]) Tout[ca * Q + b] = va; }
                    if (pending & (2u << (2 * r))) { if (vb < Tout[cb * Q + a]) Tout[cb * Q + a] = vb; }
                }
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < VEC; ++r) {
                if (pending & (3u << (2 * r))) {
                    const int a = path[r] & 255, b = (path[r] >> 8) & 255, ca = (path[r] >> 16) & 15, cb = (path[r] >> 20) & 15;
                    const T va = dw[r] - ua[r], vb = dw[r] - ub[r];
                    if (!(va < Tout[ca * Q + b])) pending &= ~(1u << (2 * r));
                    if (!(vb < Tout[cb * Q + a])) pending &= ~(2u << (2 * r));
                }
            }
            __syncwarp();
        }
    }
    // ---- fill-in (NBLDPC.cpp:1095-1102), normal domain, un-permute, correct (1105-1128) ---------------------------------
    T *cout = P.c2v_out + fE * Q;
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        {   // indices no path landed on take the smallest dU of the other columns; the owner of t knows it
            T tv[VEC];
            ld_vec<T, VEC>(Tout + d * Q + x0, tv);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (tv[j] == INF) tv[j] = ((byte_of(c01w, j) & 15) == d) ? m1[j] : m0[j];
            st_vec<T, VEC>(Tout + d * Q + x0, tv);
        }
        __syncwarp();
        const int e = e0 + d;
        const int bs = syn ^ meta[TW_DC + d];
        const T Tb = Tout[d * Q + bs];
        const unsigned long long tt = ld_bytes<VEC>(g.pfwd + (size_t)e * Q + x0);
        T out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            T v = Tb - Tout[d * Q + (byte_of(tt, j) ^ bs)];
            if (P.factor != T(1)) v = v / P.factor;                 // x / 1 == x exactly; skips the IEEE division routine
            out[j] = (x0 + j == 0) ? T(0) : (v < -P.offset ? v + P.offset : (v > P.offset ? v - P.offset : T(0)));
        }
        st_vec<T, VEC>(cout + (size_t)e * Q + x0, out);
    }
}

}  // namespace nbl
