// Device-side building blocks shared by the decode kernels (sm_100a).
//
// Work decomposition used everywhere in this file set: a GROUP of Q threads owns one LLR vector
// position-per-thread (thread x <-> field element x, slot 0 = the implicit zero LLR of symbol 0).
// Q >= 64: a group is Q/32 whole warps and synchronises on its own named barrier; Q <= 32: a group is
// (part of) one warp and synchronises with __syncwarp.  A CTA holds G groups (G*Q >= 128 threads).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>

namespace nbl {

struct DevGraph {
    int q, p, N, M, E, max_dv, max_dc;
    const int *chk_ptr;      // [M+1]
    const int *chk_var;      // [E]   variable of edge e (edge ids are check-major)
    const int *var_ptr;      // [N+1]
    const int *var_edge;     // [E]   edge ids of a variable, in the code file's order (= summation order)
    const uint8_t *pfwd;     // [E][q] h_e * x
    const uint8_t *pinv;     // [E][q] h_e^-1 * x
    const int *rowdesc;      // [M][32] warp fast path: per edge {variable, degree, edge ids...} (or null)
};

template <typename T> struct Real;
template <> struct Real<float> {
    using key_t = unsigned;
    __device__ static float inf() { return __int_as_float(0x7f800000); }
    __device__ static unsigned bits(float v) { return __float_as_uint(v); }
};
template <> struct Real<double> {
    using key_t = unsigned long long;
    __device__ static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    __device__ static unsigned long long bits(double v) { return (unsigned long long)__double_as_longlong(v); }
};

// ---- group geometry ------------------------------------------------------------------------------
template <int Q> struct Grp {
    static constexpr int WARPS = (Q + 31) / 32;
    static constexpr int G = (Q >= 128) ? 1 : 128 / Q;          // groups per CTA
    static constexpr int THREADS = G * Q;
    __device__ static unsigned mask() {
        if constexpr (Q >= 32) return 0xffffffffu;
        else {
            unsigned lane = threadIdx.x & 31u;
            return ((1u << Q) - 1u) << ((lane / Q) * Q);
        }
    }
    // all Q threads of group `gid` (index inside the CTA) rendezvous; shared memory is ordered
    __device__ static void sync(int gid) {
        if (Q > 32) asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(Q) : "memory");
        else __syncwarp(mask());
    }
};

// ---- warp-level arg-max on unsigned keys -----------------------------------------------------------
// winner = largest key; LOW=true: ties -> lowest lane, LOW=false: ties -> highest lane.
template <bool LOW> __device__ inline int pick_lane(unsigned ballot) {
    return LOW ? (__ffs(ballot) - 1) : (31 - __clz(ballot));
}
template <bool LOW> __device__ inline void warp_argmax(unsigned key, unsigned mask, unsigned &kmax, int &lane) {
    kmax = __reduce_max_sync(mask, key);
    lane = pick_lane<LOW>(__ballot_sync(mask, key == kmax));
}
template <bool LOW> __device__ inline void warp_argmax(unsigned long long key, unsigned mask, unsigned long long &kmax, int &lane) {
    unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    unsigned hmax = __reduce_max_sync(mask, hi);
    unsigned lmax = __reduce_max_sync(mask, hi == hmax ? lo : 0u);
    lane = pick_lane<LOW>(__ballot_sync(mask, hi == hmax && lo == lmax));
    kmax = ((unsigned long long)hmax << 32) | lmax;
}

// Scratch for cross-warp reductions: NV values at a time, double-buffered so that one group sync per
// call suffices (a slot is rewritten only after every thread passed the sync of the next call).
template <typename K, int WARPS, int NV> struct RedScratch {
    K key[2][NV][WARPS];
    int idx[2][NV][WARPS];
};

// Group arg-max of NV keys per thread.  Returns, for each value, the winning x (thread index in group)
// and its key.  `phase` is a per-thread counter the caller keeps (toggles the scratch slot).
template <int Q, typename K, int NV, bool LOW>
__device__ inline void group_argmax(const K (&key)[NV], int x, int gid, RedScratch<K, Grp<Q>::WARPS, NV> *sc, int &phase,
                                    K (&kout)[NV], int (&xout)[NV]) {
    constexpr int W = Grp<Q>::WARPS;
    const unsigned m = Grp<Q>::mask();
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        int lane;
        warp_argmax<LOW>(key[v], m, kout[v], lane);
        xout[v] = __shfl_sync(m, x, lane);
    }
    if (W > 1) {
        const int slot = phase & 1;
        phase ^= 1;
        const int w = x >> 5;
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int v = 0; v < NV; ++v) { sc->key[slot][v][w] = kout[v]; sc->idx[slot][v][w] = xout[v]; }
        }
        Grp<Q>::sync(gid);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            K bk = sc->key[slot][v][0];
            int bx = sc->idx[slot][v][0];
#pragma unroll
            for (int ww = 1; ww < W; ++ww) {
                K k2 = sc->key[slot][v][ww];
                bool take = LOW ? (k2 > bk) : (k2 >= bk);
                if (take) { bk = k2; bx = sc->idx[slot][v][ww]; }
            }
            kout[v] = bk; xout[v] = bx;
        }
    }
}

// DecideLLRVector (NBLDPC.cpp:1542-1562): arg-max over the strictly positive entries, ties to the lowest
// symbol, 0 when nothing is positive.  Positive IEEE values order like their bit patterns.
template <typename T> __device__ inline typename Real<T>::key_t pos_key(T v) {
    return v > T(0) ? Real<T>::bits(v) : typename Real<T>::key_t(0);
}
// total-order key for any finite value (EMS sorting): larger value <-> larger key, key 0 is "dead"
__device__ inline unsigned order_key(float v) {
    unsigned u = __float_as_uint(v == 0.f ? 0.f : v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline unsigned long long order_key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v == 0.0 ? 0.0 : v);
    return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
}

template <int Q, typename T, int NV>
__device__ inline void group_decide(const T (&val)[NV], int x, int gid,
                                    RedScratch<typename Real<T>::key_t, Grp<Q>::WARPS, NV> *sc, int &phase, int (&out)[NV]) {
    using K = typename Real<T>::key_t;
    K key[NV], kmax[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) key[v] = pos_key<T>(val[v]);
    group_argmax<Q, K, NV, true>(key, x, gid, sc, phase, kmax, out);
#pragma unroll
    for (int v = 0; v < NV; ++v) if (kmax[v] == K(0)) out[v] = 0;
}

// exclusive prefix count of `flag` over the group's threads (ascending x) and the group total
template <int Q> __device__ inline int group_prefix_count(bool flag, int x, int gid, int *warp_tot /*[WARPS]*/, int &total) {
    constexpr int W = Grp<Q>::WARPS;
    const unsigned m = Grp<Q>::mask();
    unsigned b = __ballot_sync(m, flag) & m;
    unsigned lane = threadIdx.x & 31u;
    int pre = __popc(b & ((1u << lane) - 1u));
    if (W == 1) { total = __popc(b); return pre; }
    if (lane == 0) warp_tot[x >> 5] = __popc(b);
    Grp<Q>::sync(gid);
    int off = 0; total = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) { int t = warp_tot[w]; if (w < (x >> 5)) off += t; total += t; }
    return pre + off;
}

template <typename T> __device__ inline T soft_threshold(T v, T factor, T offset) {
    v = v / factor;                       // EMS NBLDPC.cpp:903-915, T-EMS :1114-1126
    if (v < -offset) return v + offset;
    if (v > offset) return v - offset;
    return T(0);
}

// order-preserving atomic min for non-negative values in shared memory
__device__ inline void atomic_min_nonneg(float *addr, float v) { atomicMin((unsigned *)addr, __float_as_uint(v)); }
__device__ inline void atomic_min_nonneg(double *addr, double v) {
    atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

template <typename T> __device__ inline void atomic_min_any(T *addr, T v) {      // values of either sign (CAS loop)
    using U = typename std::conditional<sizeof(T) == 4, unsigned, unsigned long long>::type;
    U *a = reinterpret_cast<U *>(addr);
    U old = *a;
    while (true) {
        T cur;
        memcpy(&cur, &old, sizeof(T));
        if (!(v < cur)) return;
        U nv; memcpy(&nv, &v, sizeof(T));
        const U prev = atomicCAS(a, old, nv);
        if (prev == old) return;
        old = prev;
    }
}

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace nbl
