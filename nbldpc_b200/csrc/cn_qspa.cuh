// QSPA (sum-product) check node as a fast Walsh-Hadamard transform, one group of Q threads per check row.
//
// Replaces L_Back / L_Forward / LLR_BoxPlus of the reference (NBLDPC.cpp:747-767, 1565-1712), which evaluates
// every pairwise box-plus in the log domain with q^2 Jacobian terms (O(dc^2 q^2) exp/log per check).
// GF(2^p) addition is XOR, so the check constraint sum_l h_l x_l = 0 makes the extrinsic distribution of
// h_d x_d the XOR-convolution of the other inputs' (permuted) distributions; the Walsh-Hadamard transform
// diagonalises XOR-convolution:   c2v_d = ln( WHT( prod_{l != d} WHT(P_l) ) ),  P_l[h_l x] = softmax(V_l)[x].
// O(dc q log q) per check.  The transform runs in fp64 in both precision modes (probabilities below
// ~1e-7 of the maximum would vanish in an fp32 transform).
#pragma once
#include "dev_common.cuh"

namespace nbl {

template <typename T, int Q> struct QspaSmem {
    using K = typename Real<T>::key_t;
    T *V;         // [maxdc][Q]
    double *F;    // [maxdc][Q] transformed inputs
    double *R;    // [Q] product / output work vector
    int *sdec;    // [maxdc]
    RedScratch<K, Grp<Q>::WARPS, 1> *red1;
    RedScratch<K, Grp<Q>::WARPS, 2> *red;
    __host__ __device__ static size_t bytes(int maxdc) {
        size_t b = sizeof(double) * (size_t)Q * (maxdc + 1) + sizeof(T) * (size_t)Q * maxdc;
        b = align_up(b, 8);
        b += sizeof(RedScratch<K, Grp<Q>::WARPS, 1>) + sizeof(RedScratch<K, Grp<Q>::WARPS, 2>);
        b = align_up(b, 4);
        b += sizeof(int) * maxdc;
        return align_up(b, 16);
    }
    __device__ void carve(unsigned char *base, int maxdc) {
        double *f = reinterpret_cast<double *>(base);
        F = f; f += (size_t)maxdc * Q; R = f; f += Q;
        T *t = reinterpret_cast<T *>(f);
        V = t; t += (size_t)maxdc * Q;
        unsigned char *c = base + align_up((size_t)(reinterpret_cast<unsigned char *>(t) - base), 8);
        red1 = reinterpret_cast<RedScratch<K, Grp<Q>::WARPS, 1> *>(c); c += sizeof(*red1);
        red = reinterpret_cast<RedScratch<K, Grp<Q>::WARPS, 2> *>(c); c += sizeof(*red);
        c = base + align_up((size_t)(c - base), 4);
        sdec = reinterpret_cast<int *>(c);
    }
};

// in-place Q-point WHT of `n` vectors stored back to back in shared memory (buf[v*Q + x]); unnormalised
template <int Q> __device__ inline void group_wht(double *buf, int n, int x, int gid) {
    for (int bit = 1; bit < Q; bit <<= 1) {
        Grp<Q>::sync(gid);
        double mine[16], other[16];
        for (int v = 0; v < n; ++v) { mine[v] = buf[v * Q + x]; other[v] = buf[v * Q + (x ^ bit)]; }
        Grp<Q>::sync(gid);
        for (int v = 0; v < n; ++v) buf[v * Q + x] = (x & bit) ? (other[v] - mine[v]) : (mine[v] + other[v]);
    }
    Grp<Q>::sync(gid);
}

// Q-point WHT of up to NV vectors held one element per thread in registers (v[i] = vector i at index x); unnormalised.
// The stages inside a warp are shuffles; only the stages that cross warps (bit >= 32, q >= 64) go through shared memory
// (scratch [NV][Q]).  Same butterflies in the same order as group_wht.
template <int Q, int NV> __device__ inline void group_wht_reg(double (&v)[NV], int n, double *scratch, int x, int gid) {
    const unsigned m = Grp<Q>::mask();
#pragma unroll
    for (int bit = 1; bit < (Q < 32 ? Q : 32); bit <<= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < n) {
                const double other = __shfl_xor_sync(m, v[i], bit);
                v[i] = (x & bit) ? (other - v[i]) : (v[i] + other);
            }
    }
#pragma unroll
    for (int bit = 32; bit < Q; bit <<= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) if (i < n) scratch[i * Q + x] = v[i];
        Grp<Q>::sync(gid);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < n) {
                const double other = scratch[i * Q + (x ^ bit)];
                v[i] = (x & bit) ? (other - v[i]) : (v[i] + other);
            }
        Grp<Q>::sync(gid);
    }
}

// exp / log of the probability-domain transform.  Parity mode (T = double): fp64 library routines.  Fast mode (T = float): the
// messages are fp32 anyway, so fp32 ACCURACY (relative error ~1e-7) is enough -- but not fp32 RANGE: with 256-QAM the
// inputs span hundreds of nats and the tails between e^-87 and e^-745 still steer the decoder (measured: an fp32-range
// exp raises the FER of cfg4 from 0 to 3e-2 at 16 dB).  So: 2^n from the fp64 exponent field, 2^frac / log(mantissa)
// from the fp32 special-function unit.
template <typename T> __device__ inline double qspa_exp(double x);              // x <= 0
template <> __device__ inline double qspa_exp<double>(double x) { return exp(x); }
template <> __device__ inline double qspa_exp<float>(double x) {
    const double t = x * 1.4426950408889634;                         // log2(e)
    if (t < -1021.0) return 0.0;                                     // below the normal range: flush (fp64 exp would give a denormal)
    const double fn = floor(t);
    const float p = exp2f((float)(t - fn));                          // [1, 2)
    const double d = (double)p;
    return __hiloint2double(__double2hiint(d) + ((int)fn << 20), __double2loint(d));
}
template <typename T> __device__ inline double qspa_log(double x);              // x > 0, normal
template <> __device__ inline double qspa_log<double>(double x) { return log(x); }
template <> __device__ inline double qspa_log<float>(double x) {
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
    return (double)e * 0.6931471805599453 + (double)(__log2f((float)m) * 0.693147181f);   // lg2.approx: abs error < 3e-7
}

constexpr int QSPA_REG_DC = 8;      // check degrees served by the register / shuffle form (larger ones: shared-memory form)

template <typename T, int Q>
__device__ void cn_qspa(QspaSmem<T, Q> &S, const DevGraph &g, int gid, int x, int e0, int dcn,
                        T *__restrict__ c2v_out) {
    if (dcn <= QSPA_REG_DC) {
        // distributions in the syndrome domain, maximum normalised to 1 (the scale cancels in the LLR); all inputs are
        // transformed together, all dcn products are formed in registers and transformed back together.
        double F[QSPA_REG_DC], R[QSPA_REG_DC];
#pragma unroll
        for (int d = 0; d < QSPA_REG_DC; ++d) {
            F[d] = 1.0;
            if (d < dcn) {
                const double vmax = (double)S.V[d * Q + S.sdec[d]];
                const int k = g.pinv[(size_t)(e0 + d) * Q + x];     // P[s] = p[h^-1 s]
                F[d] = qspa_exp<T>((double)S.V[d * Q + k] - vmax);
            }
        }
        group_wht_reg<Q, QSPA_REG_DC>(F, dcn, S.F, x, gid);
#pragma unroll
        for (int d = 0; d < QSPA_REG_DC; ++d) {
            double r = 1.0;
#pragma unroll
            for (int l = 0; l < QSPA_REG_DC; ++l) if (l != d && l < dcn) r *= F[l];
            R[d] = r;
        }
        group_wht_reg<Q, QSPA_REG_DC>(R, dcn, S.F, x, gid);
#pragma unroll
        for (int d = 0; d < QSPA_REG_DC; ++d) if (d < dcn) S.F[d * Q + x] = R[d];
        Grp<Q>::sync(gid);
#pragma unroll
        for (int d = 0; d < QSPA_REG_DC; ++d) {
            if (d < dcn) {
                // The true inverse transform is a sum of non-negative products; entries whose probability is below the
                // fp64 rounding noise of the transform (~2^-53 of the total) come out as noise of either sign, so they
                // are floored there before the log (their LLR saturates near -37 relative to symbol 0).
                const double tot = S.F[d * Q];
                const double fl = fabs(tot) * 1.1102230246251565e-16 + 1e-300;
                T out = T(0);
                if (x != 0) {
                    const double a = S.F[d * Q + g.pfwd[(size_t)(e0 + d) * Q + x]];
                    out = (T)(qspa_log<T>(a > fl ? a : fl) - qspa_log<T>(tot > fl ? tot : fl));
                }
                c2v_out[(size_t)(e0 + d) * Q + x] = out;
            }
        }
        return;
    }
    for (int d = 0; d < dcn; ++d) {
        const double vmax = (double)S.V[d * Q + S.sdec[d]];
        const int k = g.pinv[(size_t)(e0 + d) * Q + x];         // P[s] = p[h^-1 s]
        S.F[d * Q + x] = exp((double)S.V[d * Q + k] - vmax);
    }
    group_wht<Q>(S.F, dcn, x, gid);
    for (int d = 0; d < dcn; ++d) {
        double r = 1.0;
        for (int l = 0; l < dcn; ++l) if (l != d) r *= S.F[l * Q + x];
        S.R[x] = r;
        group_wht<Q>(S.R, 1, x, gid);
        const double tot = S.R[0];
        const double fl = fabs(tot) * 1.1102230246251565e-16 + 1e-300;
        T out = T(0);
        if (x != 0) {
            const double a = S.R[g.pfwd[(size_t)(e0 + d) * Q + x]];
            out = (T)(log(a > fl ? a : fl) - log(tot > fl ? tot : fl));
        }
        c2v_out[(size_t)(e0 + d) * Q + x] = out;
        Grp<Q>::sync(gid);
    }
}

}  // namespace nbl
