// Fast path of the fused iteration kernel for EMS (n_m a power of two, n_m^(dc-1) small): ONE WARP per (frame, check row),
// VEC = Q/32 consecutive field elements per lane, a private shared-memory slab per warp, __syncwarp only.
//
// Replaces SortLLRVector / ConstructConf and the output stage of Decoding_EMS (NBLDPC.cpp:859-918, 1715-1786) with the
// same candidate sets:
//   * the n_m largest entries of every input (ties: larger symbol first, the order the reference's insertion sort leaves)
//     by n_m warp arg-max rounds on the values while they are still in registers;
//   * conf(q,1) in gather form: the lane that owns syndrome s inverts the one deviating input
//     (h^-1 (s ^ rest) = h^-1 s ^ h^-1 rest: own table bytes and one uniform byte);
//   * conf(n_m,n_c) by direct enumeration, exactly the reference's n_m^(dc-1) combinations spread over the lanes, each
//     landing with a shared-memory atomic max on the order-preserving key of its sum (2 808 leaves per GF(64) check
//     where a per-syndrome dynamic programme would evaluate 24 576);
//   * sums are formed left to right over the inputs as the reference's recursion forms them; the maximum per syndrome is
//     order independent (strict '>' keeps a value, not a path).
#pragma once
#include "tems_warp.cuh"

namespace nbl {

constexpr int EW_MAXNM = 16;          // n_m served by this path (power of two)
constexpr int EW_MAXCOMBOS = 4096;    // n_m^(dc-1) served by this path

template <typename T, int Q> struct EwSmem {
    using K = typename Real<T>::key_t;
    // per-warp slab:  V [TW_DC][Q] v2c of the row (natural order);  B [Q] keys -> values of the running output;
    //                 sval [TW_DC][EW_MAXNM] sorted values;  ssyn [TW_DC][EW_MAXNM] h_d * symbol of the sorted entries
    static constexpr size_t off_b = (size_t)Q * TW_DC * sizeof(T);
    static constexpr size_t off_sval = off_b + (size_t)Q * sizeof(K);
    static constexpr size_t off_ssyn = off_sval + (size_t)TW_DC * EW_MAXNM * sizeof(T);
    static constexpr size_t raw = off_ssyn + (size_t)TW_DC * EW_MAXNM * sizeof(int);
    static constexpr size_t vec_align = (size_t)Q * sizeof(T) <= 1024 ? (size_t)Q * sizeof(T) : 16;
    static constexpr bool xor_gather = (size_t)Q * sizeof(T) <= 1024;
    static constexpr size_t bytes_per_warp = ((raw + vec_align - 1) / vec_align) * vec_align;
    static constexpr size_t bytes = bytes_per_warp * TW_WARPS;
};

__device__ __forceinline__ float key_value(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }
__device__ __forceinline__ double key_value(unsigned long long k) {
    return __longlong_as_double((long long)((k & 0x8000000000000000ULL) ? (k ^ 0x8000000000000000ULL) : ~k));
}

template <typename T, int Q, int DC, int DV>
__global__ void __launch_bounds__(32 * TW_WARPS) k_iterate_ems_warp(const IterParams<T> P) {
    constexpr int VEC = Q / 32;
    using S = EwSmem<T, Q>;
    using K = typename Real<T>::key_t;
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
    K *Bk = reinterpret_cast<K *>(slab + S::off_b);
    T *Bv = reinterpret_cast<T *>(slab + S::off_b);                 // same storage once the keys are turned back into values
    T *sval = reinterpret_cast<T *>(slab + S::off_sval);
    int *ssyn = reinterpret_cast<int *>(slab + S::off_ssyn);
    const unsigned slab32 = smem_u32(slab);
    const unsigned FULL = 0xffffffffu;
    const int nm = P.p0, nc = P.p1;
    const int lg = 31 - __clz(nm);                                  // n_m = 1 << lg (host check)

    const int *rd = g.rowdesc + (size_t)row * 32;
    const int e0 = DC ? DC * row : rd[TW_DSC - 1];
    const int dcn = DC ? DC : rd[2 * TW_DSC - 1];
    const T *lchf = P.lch + (size_t)f * g.N * Q + VEC * lane;
    const T *cinf = P.c2v_in + (size_t)f * g.E * Q + VEC * lane;
    const int x0 = VEC * lane;
    int syn_dec = 0;
    unsigned long long tfw[TW_DC];                                  // the lane's own bytes of the rows h * x
    // ---- steps A-C for the row's variables (NBLDPC.cpp:808-857): no damping in EMS; the n_m best of each message --------
    T nl[2][VEC], nm_[2][TW_HOIST][VEC]; int4 nd[2];
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
        if (lane == 0) P.dec[(size_t)f * g.N + col] = (uint8_t)dpost;
        syn_dec ^= row_byte<VEC>(tf, dpost);
        st_vec<T, VEC>(V + d * Q + x0, vnew);
        // SortLLRVector, first n_m places: descending, ties -> larger symbol first
        K key[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) key[j] = order_key(vnew[j]);
#pragma unroll 1
        for (int r = 0; r < nm; ++r) {
            K lk = key[0]; int lj = 0;
#pragma unroll
            for (int j = 1; j < VEC; ++j) if (key[j] >= lk) { lk = key[j]; lj = j; }       // larger symbol wins ties
            K kmax; int src;
            warp_argmax<false>(lk, FULL, kmax, src);                                       // highest lane among ties
            const int xw = __shfl_sync(FULL, x0 + lj, src);
            if (lane == src) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) if (j == lj) key[j] = K(0);                  // taken
            }
            if (lane == 0) { sval[d * EW_MAXNM + r] = key_value(kmax); ssyn[d * EW_MAXNM + r] = xw; }   // symbol; h * symbol below
        }
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
    __syncwarp();
    for (int i = lane; i < dcn * EW_MAXNM; i += 32) {               // symbols -> syndromes h_d * symbol, all lists at once
        const int d = i / EW_MAXNM, r = i % EW_MAXNM;
        if (r < nm) ssyn[i] = g.pfwd[(size_t)(e0 + d) * Q + ssyn[i]];
    }
    unsigned long long tiv[TW_DC];                                  // the lane's own bytes of the rows h^-1 * x
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) { tiv[d] = 0; if (DC || d < dcn) tiv[d] = ld_bytes<VEC>(g.pinv + (size_t)(e0 + d) * Q + x0); }
    __syncwarp();
    // ---- check node, one output edge at a time (NBLDPC.cpp:883-916) ------------------------------------------------------
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + VEC * lane;
    const int no = dcn - 1;                                         // inputs taking part
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) {
        if (DC || d < dcn) {
            // best syndromes / values of the other inputs
            int o[TW_DC - 1] = {0, 0, 0}; int n = 0;
#pragma unroll
            for (int m = 0; m < TW_DC; ++m) if (m != d && (DC || m < dcn)) { if (n < TW_DC - 1) o[n] = m; ++n; }
#pragma unroll
            for (int j = 0; j < VEC; ++j) Bk[x0 + j] = K(0);        // -inf
            // conf(q,1): input l deviates to any symbol, the others take their best
            T b1[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) b1[j] = -Real<T>::inf();
#pragma unroll
            for (int li = 0; li < TW_DC - 1; ++li) {
                if (li < no) {
                    const int l = o[li];
                    int rest = 0;
#pragma unroll
                    for (int mi = 0; mi < TW_DC - 1; ++mi) if (mi < no && mi != li) rest ^= ssyn[o[mi] * EW_MAXNM];
                    // y = h_l^-1 (s ^ rest) = h_l^-1 s ^ h_l^-1 rest
                    unsigned long long tl = 0;
#pragma unroll
                    for (int m = 0; m < TW_DC; ++m) if (m == l) tl = tiv[m];
                    const int yr = row_byte<VEC>(tl, rest);
                    const G gv(slab32 + l * VB, yr);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        const T vy = gv.at(tl, j);
                        T sum = T(0);
#pragma unroll
                        for (int mi = 0; mi < TW_DC - 1; ++mi) if (mi < no) sum = sum + (mi == li ? vy : sval[o[mi] * EW_MAXNM]);
                        b1[j] = fmax(b1[j], sum);
                    }
                }
            }
            __syncwarp();
            // conf(n_m, n_c): every combination of the n_m best of the other inputs with at most n_c non-best picks.
            // A lane takes prefixes (picks of all but the last input) and walks the last input's list; sums are formed
            // left to right over the inputs.  A sum is never -0 (it starts from +0), so its key needs no normalisation.
            {
                const int last = no - 1;
                const int prefixes = 1 << (lg * last);
                const T *lv = sval + o[last < 0 ? 0 : last] * EW_MAXNM;
                const int *ls = ssyn + o[last < 0 ? 0 : last] * EW_MAXNM;
#pragma unroll 1
                for (int c = lane; c < prefixes; c += 32) {
                    int dev = 0, sy = 0; T sum = T(0);
#pragma unroll
                    for (int mi = 0; mi < TW_DC - 2; ++mi) {
                        if (mi < last) {
                            const int k = (c >> (lg * mi)) & (nm - 1);
                            dev += k != 0;
                            sum = sum + sval[o[mi] * EW_MAXNM + k];
                            sy ^= ssyn[o[mi] * EW_MAXNM + k];
                        }
                    }
                    const int kend = dev < nc ? nm : (dev == nc ? 1 : 0);   // the last input may deviate only if dev < n_c
#pragma unroll 1
                    for (int k = 0; k < kend; ++k) {
                        const T t = sum + lv[k];
                        K key;
                        if constexpr (sizeof(T) == 4) { const unsigned u = __float_as_uint(t); key = u ^ ((unsigned)((int)u >> 31) | 0x80000000u); }
                        else key = order_key(t);
                        atomicMax(&Bk[sy ^ ls[k]], key);
                    }
                }
            }
            __syncwarp();
            // merge, back to values (in place), then un-permute and correct
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const K k = Bk[x0 + j];
                const T v = k == K(0) ? -Real<T>::inf() : key_value(k);
                Bv[x0 + j] = fmax(v, b1[j]);
            }
            __syncwarp();
            const T b0 = Bv[0];
            const G gb(slab32 + (unsigned)S::off_b, 0);
            unsigned long long td = 0;
#pragma unroll
            for (int m = 0; m < TW_DC; ++m) if (m == d) td = tfw[m];
            T out[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) out[j] = gb.at(td, j) - b0;
            if (P.factor != T(1)) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) out[j] = div_call(out[j], P.factor);
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const T v = out[j];
                out[j] = v < -P.offset ? v + P.offset : (v > P.offset ? v - P.offset : T(0));
            }
            if (lane == 0) out[0] = T(0);                           // slot 0 is the implicit LLR of symbol 0
            st_vec<T, VEC>(coutf + (size_t)(e0 + d) * Q, out);
            __syncwarp();                                           // B is re-initialised for the next output
        }
    }
}

}  // namespace nbl
