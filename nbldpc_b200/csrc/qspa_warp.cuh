// Fast path of the fused iteration kernel for QSPA: ONE WARP per (frame, check row), VEC = Q/32 consecutive field elements
// per lane, a private shared-memory slab per warp, __syncwarp only.
//
// Same arithmetic as the group form (cn_qspa.cuh):  c2v_d = ln WHT( prod_{l != d} WHT(P_l) ),  P_l[h_l x] = softmax(V_l)[x],
// transform in fp64.  Here a lane holds VEC consecutive points of every vector in registers: the first log2(VEC) butterfly
// stages are in-lane adds, the remaining five are warp shuffles (no shared memory, no barrier), in the same stage order
// and with the same operand order as the group form, so the two produce identical bits.  The products are formed in
// registers; an output vector goes through shared memory once, for the un-permutation gather x -> h_d x.
#pragma once
#include "tems_warp.cuh"
#include "cn_qspa.cuh"

namespace nbl {

template <typename T, int Q> struct QwSmem {
    // per-warp slab:  V [TW_DC][Q] v2c of the row (natural order), later R [Q] doubles: one output vector for the gather;
    //                 F [TW_DC][Q] doubles: the transformed inputs (split layout: a lane's own points, 16-byte pieces)
    static constexpr size_t v_bytes = (size_t)Q * TW_DC * sizeof(T) < (size_t)Q * sizeof(double) ? (size_t)Q * sizeof(double) : (size_t)Q * TW_DC * sizeof(T);
    static constexpr size_t off_f = v_bytes;
    static constexpr size_t raw = off_f + (size_t)Q * TW_DC * sizeof(double);
    static constexpr size_t vec_align = (size_t)Q * sizeof(T) <= 1024 ? (size_t)Q * sizeof(T) : 16;
    static constexpr bool xor_gather = (size_t)Q * sizeof(T) <= 1024;
    static constexpr size_t bytes_per_warp = ((raw + vec_align - 1) / vec_align) * vec_align;
    static constexpr size_t bytes = bytes_per_warp * TW_WARPS;
};

// Q-point WHT of one vector spread VEC-per-lane over the warp (element VEC*lane + j): stages bit = 1, 2, 4, ... as in group_wht
template <int VEC> __device__ __forceinline__ void warp_wht(double (&v)[VEC], int lane) {
#pragma unroll
    for (int bit = 1; bit < VEC; bit <<= 1) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            if (!(j & bit)) {
                const double a = v[j], b = v[j | bit];
                v[j] = a + b;                                       // x & bit == 0:  mine + other
                v[j | bit] = a - b;                                 // x & bit != 0:  other - mine
            }
    }
#pragma unroll
    for (int lb = 1; lb < 32; lb <<= 1) {
        const bool up = lane & lb;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const double other = __shfl_xor_sync(0xffffffffu, v[j], lb);
            v[j] = up ? (other - v[j]) : (v[j] + other);
        }
    }
}

template <typename T, int Q, int DC, int DV>
__global__ void __launch_bounds__(32 * TW_WARPS) k_iterate_qspa_warp(const IterParams<T> P) {
    constexpr int VEC = Q / 32;
    using S = QwSmem<T, Q>;
    using G = Gather<T, S::xor_gather>;
    constexpr int VB = (int)(Q * sizeof(T));
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const DevGraph &g = P.g;
    // warp index out of a shuffle, latch out of a vote: both exits are provably taken by whole warps, so the warp collectives
    // below are not compiled with a second, divergent fall-back path (see tems_warp.cuh)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int f = (int)(blockIdx.z * 65535u + blockIdx.y), row = (int)blockIdx.x * TW_WARPS + warp;
    if (f >= P.B || row >= g.M) return;
    if (P.early_stop && __any_sync(0xffffffffu, P.done[f] != 0)) return;

    unsigned char *slab = smem_raw + S::bytes_per_warp * warp;
    T *V = reinterpret_cast<T *>(slab);
    T *Rv = reinterpret_cast<T *>(slab);                           // one output vector for the un-permutation gather; reuses V once every input is transformed
    double *Fs = reinterpret_cast<double *>(slab + S::off_f);
    const unsigned slab32 = smem_u32(slab);

    const int *rd = g.rowdesc + (size_t)row * 32;
    const int e0 = DC ? DC * row : rd[TW_DSC - 1];
    const int dcn = DC ? DC : rd[2 * TW_DSC - 1];
    const T *lchf = P.lch + (size_t)f * g.N * Q + VEC * lane;
    const T *cinf = P.c2v_in + (size_t)f * g.E * Q + VEC * lane;
    T *v2cf = P.v2c + (size_t)f * g.E * Q + VEC * lane;
    uint8_t *vdecf = P.vdec + (size_t)f * g.E;
    const int x0 = VEC * lane;
    int syn_dec = 0;
    unsigned sdecw = 0;                                             // byte d: decide(v2c_d)
    unsigned long long tfw[TW_DC];                                  // the lane's own bytes of the rows h * x
    // ---- steps A-C for the row's variables (NBLDPC.cpp:676-744): posterior, decision, syndrome, damped v2c --------------
    T nl[2][VEC], nm_[2][TW_HOIST][VEC]; int4 nd[2]; int nbefore[2];
    auto fetch = [&](int d, int slot) {
        nd[slot] = *reinterpret_cast<const int4 *>(rd + TW_DSC * d);
        ld_vec<T, VEC>(lchf + (size_t)nd[slot].x * Q, nl[slot]);
        const int nv = DV ? DV : nd[slot].y;
#pragma unroll
        for (int j = 0; j < TW_HOIST; ++j) {
            if (!P.first && j < nv) ld_vec<T, VEC>(cinf + (size_t)(j == 0 ? nd[slot].z : nd[slot].w) * Q, nm_[slot][j]);
            else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) nm_[slot][j][i] = T(0);
            }
        }
        nbefore[slot] = P.first ? -1 : (int)vdecf[e0 + d];
    };
    auto edge = [&](int d, int slot, unsigned long long tf) {
        const int e = e0 + d, col = nd[slot].x, nv = DV ? DV : nd[slot].y;
        T post[VEC], vnew[VEC], own[VEC];
        const bool mine0 = nd[slot].z == e;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            post[i] = nl[slot][i] + nm_[slot][0][i];
            own[i] = mine0 ? nm_[slot][0][i] : (TW_HOIST > 1 ? nm_[slot][1][i] : T(0));
        }
        if (TW_HOIST > 1 && (DV ? DV > 1 : nv > 1)) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) post[i] = post[i] + nm_[slot][1][i];
        } else if (!mine0) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) own[i] = T(0);
        }
        if (!DV && !P.first) {
#pragma unroll 1
            for (int j = TW_HOIST; j < nv; ++j) {
                const int ej = rd[TW_DSC * d + 2 + j];
                T mj[VEC];
                ld_vec<T, VEC>(cinf + (size_t)ej * Q, mj);
#pragma unroll
                for (int i = 0; i < VEC; ++i) { post[i] = post[i] + mj[i]; if (ej == e) own[i] = mj[i]; }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) vnew[i] = post[i] - own[i];
        const int dpost = warp_decide<T, VEC>(post, lane);
        int dfin = warp_decide<T, VEC>(vnew, lane);
        if (!P.first && P.w_old != T(0) && dfin != nbefore[slot]) { // warp-uniform
            T vold[VEC];
            ld_vec<T, VEC>(v2cf + (size_t)e * Q, vold);
#pragma unroll
            for (int i = 0; i < VEC; ++i) vnew[i] = P.w_old * vold[i] + P.w_new * vnew[i];   // -fmad=false
            dfin = warp_decide<T, VEC>(vnew, lane);
        }
        st_vec<T, VEC>(v2cf + (size_t)e * Q, vnew);
        st_vec<T, VEC>(V + d * Q + x0, vnew);
        if (lane == 0) { vdecf[e] = (uint8_t)dfin; P.dec[(size_t)f * g.N + col] = (uint8_t)dpost; }
        sdecw |= (unsigned)dfin << (8 * d);
        syn_dec ^= row_byte<VEC>(tf, dpost);
    };
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) { tfw[d] = 0; if (DC || d < dcn) tfw[d] = ld_bytes<VEC>(g.pfwd + (size_t)(e0 + d) * Q + x0); }
    fetch(0, 0);
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) {
        if (DC || d < dcn) {
            if (d + 1 < TW_DC && (DC || d + 1 < dcn)) fetch(d + 1, (d + 1) & 1);
            edge(d, d & 1, tfw[d]);
        }
    }
    if (lane == 0 && syn_dec != 0) P.flag[f] = 1;
    unsigned long long tiv[TW_DC];                                  // the lane's own bytes of the rows h^-1 * x
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) { tiv[d] = 0; if (DC || d < dcn) tiv[d] = ld_bytes<VEC>(g.pinv + (size_t)(e0 + d) * Q + x0); }
    __syncwarp();
    // ---- check node: distributions in the syndrome domain, maximum normalised to 1, transformed -------------------------
    // (loops over the edges are kept rolled: straight-line code for four transforms does not fit the instruction cache)
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        const int sd = (sdecw >> (8 * d)) & 255;
        const double vmax = (double)V[d * Q + sd];
        const G gv(slab32 + d * VB, 0);
        unsigned long long tv = 0;
#pragma unroll
        for (int m = 0; m < TW_DC; ++m) if (m == d) tv = tiv[m];
        double Fd[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) Fd[j] = qspa_exp<T>((double)gv.at(tv, j) - vmax);   // P[s] = p[h^-1 s]
        warp_wht<VEC>(Fd, lane);
        st_split<double, VEC>(Fs + d * Q, lane, Fd);
    }
    __syncwarp();                                                  // V is dead from here on: R may overwrite it
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + VEC * lane;
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        double r[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) r[j] = 1.0;
#pragma unroll 1
        for (int l = 0; l < dcn; ++l) {
            if (l == d) continue;
            double Fl[VEC];
            ld_split<double, VEC>(Fs + l * Q, lane, Fl);
#pragma unroll
            for (int j = 0; j < VEC; ++j) r[j] *= Fl[j];
        }
        warp_wht<VEC>(r, lane);
        // floor at the fp64 noise of the transform (see cn_qspa.cuh) and log on the lane's OWN points -- an elementwise map
        // commutes with the un-permutation, so the vector that goes through shared memory for the gather x -> h_d x is already
        // in the message type (fp32 mode: half the stores and half the gather traffic of the fp64 vector)
        const double tot = __shfl_sync(0xffffffffu, r[0], 0);
        const double fl = fabs(tot) * 1.1102230246251565e-16 + 1e-300;
        const double lt = qspa_log<T>(tot > fl ? tot : fl);
        T y[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) y[j] = (T)(qspa_log<T>(r[j] > fl ? r[j] : fl) - lt);
        __syncwarp();                                               // the previous output's gathers are done
        st_vec<T, VEC>(Rv + x0, y);
        __syncwarp();
        unsigned long long tf = 0;
#pragma unroll
        for (int m = 0; m < TW_DC; ++m) if (m == d) tf = tfw[m];
        T out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) out[j] = Rv[byte_of(tf, j)];
        if (lane == 0) out[0] = T(0);
        st_vec<T, VEC>(coutf + (size_t)(e0 + d) * Q, out);
    }
}

}  // namespace nbl
