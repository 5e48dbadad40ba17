// Fused iteration kernel for QSPA over GF(16): FOUR LANES per (frame, check row), four field elements per lane, eight
// check rows per warp.
//
// GF(16) vectors are 16 values, so the warp kernels' "q / 32 elements per lane" does not apply and the group kernel runs
// them as 16-thread groups -- correctly, but its ncu capture on BASELINE configs[0] (profiles/r01_cfg1_v1_ncu_key.txt)
// shows ~1 400 warp instructions per check row and the long-scoreboard stall on top: every message load waits on a chain
// of index loads (chk_var -> var_ptr -> var_edge -> message), decisions go through shared-memory scratch.  A first
// register-only form with one lane per element (16 lanes per row) still needed ~1 050 instructions per row
// (profiles/r01_cfg1_v3_ncu_key.txt): with one element per lane every decision is four shuffle stages plus a vote, every
// transform four shuffle stages, every address is computed sixteen times.  Hence four elements per lane:
//   * a message vector is one 16-byte load per lane; half of the butterfly stages of the 16-point Walsh-Hadamard
//     transform and of every arg-max are in-lane, the other half two shuffles;
//   * a row descriptor (8 edges x {variable, degree, up to 4 edge ids}) makes every address known after ONE load; the
//     descriptor of edge d + 2 and the messages of edge d + 1 are in flight while edge d is computed;
//   * the permutations x -> h x / h^-1 x go through a small shared-memory slab per warp (one 16-byte store, four
//     indexed loads).
// All shuffles and votes are WARP-wide (full mask): offsets 1 and 2 stay inside a row's four lanes, so the eight rows of a
// warp run ONE instruction stream; whatever differs between them (row degree, the damping trigger, a finished frame) is
// predicated, never branched on, unless the branch is warp-uniform.
// Same arithmetic as the other QSPA kernels -- steps A-C of NBLDPC.cpp:676-744, c2v_d = ln WHT(prod_{l != d} WHT(P_l)) --
// with the butterflies of group_wht_reg (cn_qspa.cuh) in the same stage and operand order: bit-identical results.
// Serves check degrees <= 8 and variable degrees <= 4 (all GF(16) codes of the reference: d_c = 5, d_v = 2, 3).
#pragma once
#include "cn_qspa.cuh"

namespace nbl {

constexpr int QH_MAXDC = 8;     // edges per row in the descriptor
constexpr int QH_MAXDV = 4;     // messages of a variable held in registers
constexpr int QH_DSC = 8;       // ints per edge: variable, degree, edge ids (file order = summation order)
constexpr int QH_ROWS = 8;      // check rows per warp
constexpr int QH_WARPS = 4;     // warps per CTA
constexpr unsigned QH_FULL = 0xffffffffu;

template <typename T> struct QhSmem {
    // per warp: the row's current vector in natural order, for the two gathers of an edge
    static constexpr size_t v_bytes = (size_t)QH_ROWS * 16 * sizeof(T);
    static constexpr size_t r_bytes = (size_t)QH_ROWS * 16 * sizeof(double);
    static constexpr size_t bytes_per_warp = v_bytes + r_bytes;
    static constexpr size_t bytes = bytes_per_warp * QH_WARPS;
};

template <typename T> __device__ __forceinline__ void qh_load4(const T *p, T (&v)[4]) {
    if constexpr (sizeof(T) == 4) { const float4 w = *reinterpret_cast<const float4 *>(p); v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w; }
    else {
        const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}
template <typename T> __device__ __forceinline__ void qh_store4(T *p, const T (&v)[4]) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else { *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]); *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]); }
}

// DecideLLRVector (NBLDPC.cpp:1542-1562): arg-max over the non-zero symbols if the maximum is positive, else 0; ties: lowest
template <typename T> __device__ __forceinline__ int qh_decide(const T (&v)[4], int x0) {
    T key[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) key[j] = (x0 + j > 0 && v[j] > T(0)) ? v[j] : T(0);
    T m = key[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) m = key[j] > m ? key[j] : m;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) { const T t = __shfl_xor_sync(QH_FULL, m, o); m = t > m ? t : m; }
    int cand = 99;
#pragma unroll
    for (int j = 3; j >= 0; --j) if (key[j] == m) cand = x0 + j;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) cand = min(cand, __shfl_xor_sync(QH_FULL, cand, o));
    return m > T(0) ? cand : 0;
}

// 16-point WHT, element 4 * l4 + j in v[j] of lane l4 of the row: stages bit = 1, 2 in-lane, bit = 4, 8 by shuffle
__device__ __forceinline__ void qh_wht(double (&v)[4], int l4) {
#pragma unroll
    for (int bit = 1; bit < 4; bit <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (!(j & bit)) { const double a = v[j], b = v[j | bit]; v[j] = a + b; v[j | bit] = a - b; }
    }
#pragma unroll
    for (int lb = 1; lb <= 2; lb <<= 1) {
        const bool up = l4 & lb;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double other = __shfl_xor_sync(QH_FULL, v[j], lb); v[j] = up ? (other - v[j]) : (v[j] + other); }
    }
}

// What the two GF(16) kernels (QSPA here, T-EMS in tems_half.cuh) share: the row a lane works on, and steps A-C.
template <typename T> struct QhRow {
    int lane, l4, rw, x0;              // lane, my lane of the row (0..3), the row's slot in the warp (0..7), first element
    bool active;                       // the row exists and its frame is still decoding
    int f, row, e0, dcn, dmax;         // frame, check row, first edge, row degree (0 if inactive), longest row of the warp
};

// false: nothing to do for the whole warp
template <typename T> __device__ __forceinline__ bool qh_setup(const IterParams<T> &P, QhRow<T> &R) {
    const DevGraph &g = P.g;
    R.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    R.l4 = R.lane & 3; R.rw = R.lane >> 2; R.x0 = 4 * R.l4;
    const long long nrows = (long long)P.B * g.M;
    long long w = ((long long)blockIdx.x * QH_WARPS + warp) * QH_ROWS + R.rw;
    R.active = w < nrows;
    if (!R.active) w = nrows - 1;                                   // the lanes stay for the warp-wide shuffles; they store nothing
    R.f = (int)(w / g.M); R.row = (int)(w % g.M);
    if (P.early_stop && P.done[R.f]) R.active = false;
    if (!__any_sync(QH_FULL, R.active)) return false;               // warp-uniform
    R.e0 = g.chk_ptr[R.row];
    R.dcn = R.active ? g.chk_ptr[R.row + 1] - R.e0 : 0;
    R.dmax = __reduce_max_sync(QH_FULL, R.dcn);                     // the longest of the warp's rows
    return true;
}

// steps A-C (NBLDPC.cpp:676-744 / 977-1052): posterior, decisions, syndrome, damped v2c of the row's edges.
// Out: V[d] = v2c of edge d (my four elements), tf / ti = my four bytes of the rows h * x and h^-1 * x, sdw nibble d =
// decide(v2c_d).  Sets P.flag[f] when the row's syndrome is not zero.
template <typename T>
__device__ __forceinline__ void qh_vn_phase(const IterParams<T> &P, const int *__restrict__ desc, const QhRow<T> &R,
                                            T (&V)[QH_MAXDC][4], unsigned (&tf)[QH_MAXDC], unsigned (&ti)[QH_MAXDC], unsigned &sdw) {
    constexpr int Q = 16;
    const DevGraph &g = P.g;
    const int lane = R.lane, l4 = R.l4, x0 = R.x0, f = R.f, e0 = R.e0, dcn = R.dcn, dmax = R.dmax;
    const int *rd = desc + (size_t)R.row * (QH_MAXDC * QH_DSC);
    const T *lchf = P.lch + (size_t)f * g.N * Q + x0;
    const T *cinf = P.c2v_in + (size_t)f * g.E * Q + x0;
    T *v2cf = P.v2c + (size_t)f * g.E * Q + x0;
    uint8_t *vdecf = P.vdec + (size_t)f * g.E;
    // software pipeline over the edges: descriptor of edge d + 2 and messages of edge d + 1 are in flight during edge d
    int4 da[2]; int2 db[2];                                         // {variable, degree, edge ids 0..1}, {edge ids 2..3}
    T lc[2][4], mm[2][QH_MAXDV][4]; unsigned meta[2];               // meta: own slot | degree << 4 | old decision + 1 << 8 | variable << 17 (N < 2^15, checked by the host)
    auto fetch_desc = [&](int d, int slot) {
        da[slot] = make_int4(0, 0, 0, 0); db[slot] = make_int2(0, 0);
        if (d < dcn) { da[slot] = *reinterpret_cast<const int4 *>(rd + QH_DSC * d); db[slot] = *reinterpret_cast<const int2 *>(rd + QH_DSC * d + 4); }
    };
    auto fetch_msgs = [&](int d, int slot) {                        // slot of the descriptor == slot of the messages (d & 1)
        const int ej[QH_MAXDV] = {da[slot].z, da[slot].w, db[slot].x, db[slot].y};
        const int nv = da[slot].y;
        int own = QH_MAXDV, bef = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) lc[slot][i] = T(0);
#pragma unroll
        for (int j = 0; j < QH_MAXDV; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) mm[slot][j][i] = T(0);
        if (d < dcn) {
            qh_load4<T>(lchf + (size_t)da[slot].x * Q, lc[slot]);
#pragma unroll
            for (int j = 0; j < QH_MAXDV; ++j) {
                if (!P.first && j < nv) qh_load4<T>(cinf + (size_t)ej[j] * Q, mm[slot][j]);
                if (j < nv && ej[j] == e0 + d) own = j;
            }
            bef = P.first ? 0 : (int)vdecf[e0 + d] + 1;
        }
        meta[slot] = (unsigned)own | ((unsigned)nv << 4) | ((unsigned)bef << 8) | ((unsigned)da[slot].x << 17);
    };
    sdw = 0u;
    int syn = 0;
    fetch_desc(0, 0);
    fetch_desc(1, 1);
    fetch_msgs(0, 0);
    // ---- steps A-C (NBLDPC.cpp:676-744): posterior, decisions, syndrome, damped v2c ------------------------------------------
#pragma unroll
    for (int d = 0; d < QH_MAXDC; ++d) {
#pragma unroll
        for (int i = 0; i < 4; ++i) V[d][i] = T(0);
        tf[d] = 0u; ti[d] = 0u;
        if (d < dmax) {                                             // warp-uniform
            const int slot = d & 1;
            if (d + 1 < QH_MAXDC) fetch_msgs(d + 1, slot ^ 1);      // needs descriptor d + 1 (requested one step ago)
            const bool live = d < dcn;
            const int e = e0 + d, own = (int)(meta[slot] & 15u), nv = (int)((meta[slot] >> 4) & 15u), bef = (int)((meta[slot] >> 8) & 511u), col = (int)(meta[slot] >> 17);
            if (live) {
                tf[d] = *reinterpret_cast<const unsigned *>(g.pfwd + (size_t)e * Q + x0);
                ti[d] = *reinterpret_cast<const unsigned *>(g.pinv + (size_t)e * Q + x0);
            }
            T post[4], mine[4], vnew[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { post[i] = lc[slot][i]; mine[i] = T(0); }
#pragma unroll
            for (int j = 0; j < QH_MAXDV; ++j)
                if (j < nv) {                                       // iteration 1: + 0, exact
#pragma unroll
                    for (int i = 0; i < 4; ++i) { post[i] = post[i] + mm[slot][j][i]; if (j == own) mine[i] = mm[slot][j][i]; }
                }
            if (d + 2 < QH_MAXDC) fetch_desc(d + 2, slot);          // this slot's descriptor and messages are consumed
#pragma unroll
            for (int i = 0; i < 4; ++i) vnew[i] = post[i] - mine[i];
            const int dpost = qh_decide<T>(post, x0);
            int dfin = qh_decide<T>(vnew, x0);
            const bool need = live && !P.first && P.w_old != T(0) && dfin != bef - 1;
            if (__any_sync(QH_FULL, need)) {                        // warp-uniform
                T vd[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) vd[i] = vnew[i];
                if (need) {
                    T vold[4];
                    qh_load4<T>(v2cf + (size_t)e * Q, vold);
#pragma unroll
                    for (int i = 0; i < 4; ++i) vd[i] = P.w_old * vold[i] + P.w_new * vnew[i];    // -fmad=false
                }
                const int d2 = qh_decide<T>(vd, x0);
                if (need) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) vnew[i] = vd[i];
                    dfin = d2;
                }
            }
            if (live) {
                qh_store4<T>(v2cf + (size_t)e * Q, vnew);
                if (l4 == 0) { vdecf[e] = (uint8_t)dfin; P.dec[(size_t)f * g.N + col] = (uint8_t)dpost; }
            }
            // h * decision: byte dpost & 3 of the word of lane dpost >> 2 of this row
            const unsigned hw = __shfl_sync(QH_FULL, tf[d], (lane & ~3) + (dpost >> 2));
            if (live) syn ^= (int)((hw >> (8 * (dpost & 3))) & 255u);
#pragma unroll
            for (int i = 0; i < 4; ++i) V[d][i] = vnew[i];
            sdw |= (unsigned)dfin << (4 * d);
        }
    }
    if (R.active && l4 == 0 && syn != 0) P.flag[f] = 1;
}

// The QSPA kernel is written as ROLLED loops over the row's edges (one copy of the variable-node step, of the forward
// transform and of the output stage): unrolled over 8 edges it was ~4 400 instructions, 70 KB, and ran at 92 % of the GPC
// instruction cache's request rate -- bound by instruction fetch, like the T-EMS warp kernel before it was folded.
// Per-edge state that the unrolled form kept in registers lives in shared memory: the transformed distributions F[d].
template <typename T> struct QhQspaSmem {
    // per warp:  F [dc][QH_ROWS][16] doubles (transformed input distributions of the warp's eight rows),
    //            V [QH_ROWS][16] T (the current edge's vector in natural order, for the gathers), R [QH_ROWS][16] doubles
    static constexpr size_t f_bytes_per_edge = (size_t)QH_ROWS * 16 * sizeof(double);
    static constexpr size_t r_bytes = (size_t)QH_ROWS * 16 * sizeof(double);
    static constexpr size_t v_bytes = (size_t)QH_ROWS * 16 * sizeof(T);
    __host__ __device__ static size_t bytes_per_warp(int max_dc) { return f_bytes_per_edge * max_dc + r_bytes + v_bytes; }
    __host__ __device__ static size_t bytes(int max_dc) { return bytes_per_warp(max_dc) * QH_WARPS; }
};

template <typename T>
__global__ void __launch_bounds__(32 * QH_WARPS, sizeof(T) == 4 ? 6 : 2) k_iterate_qspa_half(const IterParams<T> P, const int *__restrict__ desc) {
    constexpr int Q = 16;
    extern __shared__ __align__(16) unsigned char qh_raw[];
    const DevGraph &g = P.g;
    QhRow<T> R;
    if (!qh_setup<T>(P, R)) return;
    const int warp = __shfl_sync(QH_FULL, (int)(threadIdx.x >> 5), 0), lane = R.lane, l4 = R.l4, x0 = R.x0, f = R.f, e0 = R.e0, dcn = R.dcn, dmax = R.dmax;
    unsigned char *slab = qh_raw + QhQspaSmem<T>::bytes_per_warp(g.max_dc) * warp;
    // my four points of edge d live at sF + d * QH_ROWS * 16 as two 16-byte pieces, piece h of all lanes contiguous (only this lane
    // ever reads them back): 128-bit accesses without bank conflicts (a lane's 32 bytes in one run are a 2-way conflict)
    double *sF = reinterpret_cast<double *>(slab) + 2 * lane;
    T *sY = reinterpret_cast<T *>(slab + QhQspaSmem<T>::f_bytes_per_edge * g.max_dc) + R.rw * 16;   // one output vector of my row, message type
    T *sV = reinterpret_cast<T *>(slab + QhQspaSmem<T>::f_bytes_per_edge * g.max_dc + QhQspaSmem<T>::r_bytes) + R.rw * 16;
    const int *rd = desc + (size_t)R.row * (QH_MAXDC * QH_DSC);
    const T *lchf = P.lch + (size_t)f * g.N * Q + x0;
    const T *cinf = P.c2v_in + (size_t)f * g.E * Q + x0;
    T *v2cf = P.v2c + (size_t)f * g.E * Q + x0;
    uint8_t *vdecf = P.vdec + (size_t)f * g.E;
    // ---- steps A-C (NBLDPC.cpp:676-744) and the forward transform, edge by edge.  The descriptor of edge d + 2 and the messages
    //      of edge d + 1 are requested while edge d is worked on -- into the registers edge d has just consumed.
    int4 da, na; int2 db, nb;                                       // {variable, degree, edge ids 0..1}, {edge ids 2..3}: current, next
    T lc[4], mm[QH_MAXDV][4]; unsigned meta;                        // meta: own slot | degree << 4 | old decision + 1 << 8 | variable << 17
    auto fetch_desc = [&](int d, int4 &a, int2 &b) {
        a = make_int4(0, 0, 0, 0); b = make_int2(0, 0);
        if (d < dcn) { a = *reinterpret_cast<const int4 *>(rd + QH_DSC * d); b = *reinterpret_cast<const int2 *>(rd + QH_DSC * d + 4); }
    };
    auto fetch_msgs = [&](int d) {                                  // from the descriptor in (da, db)
        const int ej[QH_MAXDV] = {da.z, da.w, db.x, db.y};
        const int nv = da.y;
        int own = QH_MAXDV, bef = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) lc[i] = T(0);
#pragma unroll
        for (int j = 0; j < QH_MAXDV; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) mm[j][i] = T(0);
        if (d < dcn) {
            qh_load4<T>(lchf + (size_t)da.x * Q, lc);
#pragma unroll
            for (int j = 0; j < QH_MAXDV; ++j) {
                if (!P.first && j < nv) qh_load4<T>(cinf + (size_t)ej[j] * Q, mm[j]);
                if (j < nv && ej[j] == e0 + d) own = j;
            }
            bef = P.first ? 0 : (int)vdecf[e0 + d] + 1;
        }
        meta = (unsigned)own | ((unsigned)nv << 4) | ((unsigned)bef << 8) | ((unsigned)da.x << 17);
    };
    int syn = 0;
    fetch_desc(0, da, db);
    fetch_msgs(0);
    fetch_desc(1, na, nb);
#pragma unroll 1
    for (int d = 0; d < dmax; ++d) {                                // dmax: the longest of the warp's rows (warp-uniform)
        const bool live = d < dcn;
        const int e = e0 + d, own = (int)(meta & 15u), nv = (int)((meta >> 4) & 15u), bef = (int)((meta >> 8) & 511u), col = (int)(meta >> 17);
        unsigned tfw = 0u, tiw = 0u;                                // my four bytes of the rows h * x and h^-1 * x
        if (live) {
            tfw = *reinterpret_cast<const unsigned *>(g.pfwd + (size_t)e * Q + x0);
            tiw = *reinterpret_cast<const unsigned *>(g.pinv + (size_t)e * Q + x0);
        }
        T post[4], mine[4], vnew[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { post[i] = lc[i]; mine[i] = T(0); }
#pragma unroll
        for (int j = 0; j < QH_MAXDV; ++j)
            if (j < nv) {                                           // iteration 1: + 0, exact
#pragma unroll
                for (int i = 0; i < 4; ++i) { post[i] = post[i] + mm[j][i]; if (j == own) mine[i] = mm[j][i]; }
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) vnew[i] = post[i] - mine[i];
        da = na; db = nb;                                           // lc, mm and the descriptor are consumed: the next edge goes out
        if (d + 1 < dmax) fetch_msgs(d + 1);
        if (d + 2 < dmax) fetch_desc(d + 2, na, nb);
        const int dpost = qh_decide<T>(post, x0);
        int dfin = qh_decide<T>(vnew, x0);
        const bool need = live && !P.first && P.w_old != T(0) && dfin != bef - 1;
        if (__any_sync(QH_FULL, need)) {                            // warp-uniform
            T vd[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) vd[i] = vnew[i];
            if (need) {
                T vold[4];
                qh_load4<T>(v2cf + (size_t)e * Q, vold);
#pragma unroll
                for (int i = 0; i < 4; ++i) vd[i] = P.w_old * vold[i] + P.w_new * vnew[i];        // -fmad=false
            }
            const int d2 = qh_decide<T>(vd, x0);
            if (need) {
#pragma unroll
                for (int i = 0; i < 4; ++i) vnew[i] = vd[i];
                dfin = d2;
            }
        }
        if (live) {
            qh_store4<T>(v2cf + (size_t)e * Q, vnew);
            if (l4 == 0) { vdecf[e] = (uint8_t)dfin; P.dec[(size_t)f * g.N + col] = (uint8_t)dpost; }
        }
        // h * decision: byte dpost & 3 of the word of lane dpost >> 2 of this row
        const unsigned hw = __shfl_sync(QH_FULL, tfw, (lane & ~3) + (dpost >> 2));
        if (live) syn ^= (int)((hw >> (8 * (dpost & 3))) & 255u);
        // check node, forward half: the distribution in the syndrome domain, maximum normalised to 1, transformed
        qh_store4<T>(sV + x0, vnew);
        __syncwarp();
        const double vmax = (double)sV[dfin];
        double t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = qspa_exp<T>((double)sV[(tiw >> (8 * i)) & 255u] - vmax);       // P[s] = p[h^-1 s]
        __syncwarp();
        qh_wht(t, l4);
        if (!live) {                                                // edges another row of the warp has and this one lacks: factor 1
#pragma unroll
            for (int i = 0; i < 4; ++i) t[i] = 1.0;
        }
        {
            double *dst = sF + (size_t)d * (QH_ROWS * 16);
            *reinterpret_cast<double2 *>(dst) = make_double2(t[0], t[1]);
            *reinterpret_cast<double2 *>(dst + 64) = make_double2(t[2], t[3]);
        }
    }
    if (R.active && l4 == 0 && syn != 0) P.flag[f] = 1;
    __syncwarp();
    // ---- check node, backward half: product of the other edges' transforms (ascending, * 1.0 is exact), inverse transform, log ----
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + x0;
#pragma unroll 1
    for (int d = 0; d < dmax; ++d) {
        double r[4] = {1.0, 1.0, 1.0, 1.0};
#pragma unroll 1
        for (int l = 0; l < dmax; ++l) {
            if (l != d) {
                const double *srcF = sF + (size_t)l * (QH_ROWS * 16);
                const double2 fa = *reinterpret_cast<const double2 *>(srcF), fb = *reinterpret_cast<const double2 *>(srcF + 64);
                r[0] *= fa.x; r[1] *= fa.y; r[2] *= fb.x; r[3] *= fb.y;
            }
        }
        const unsigned tfw = d < dcn ? *reinterpret_cast<const unsigned *>(g.pfwd + (size_t)(e0 + d) * Q + x0) : 0u;
        qh_wht(r, l4);
        // floor at the fp64 noise of the transform (see cn_qspa.cuh) and log on the lane's OWN points -- an elementwise map
        // commutes with the un-permutation, so what goes through shared memory for the gather is already in the message type
        const double tot = __shfl_sync(0xffffffffu, r[0], lane & ~3);     // point 0 of my row
        const double fl = fabs(tot) * 1.1102230246251565e-16 + 1e-300;
        const double lt = qspa_log<T>(tot > fl ? tot : fl);
        T y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = (T)(qspa_log<T>(r[i] > fl ? r[i] : fl) - lt);
        qh_store4<T>(sY + x0, y);
        __syncwarp();
        T out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = sY[(tfw >> (8 * i)) & 255u];
        __syncwarp();
        if (l4 == 0) out[0] = T(0);
        if (d < dcn) qh_store4<T>(coutf + (size_t)(e0 + d) * Q, out);
    }
}

}  // namespace nbl
