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
    // per-warp slab:  V [TW_DC][Q] v2c of the row (natural order);  B [2][Q] keys -> values of the running output (second copy:
    //                 the same vector with the two elements of every lane swapped, fp32 / q = 64 only);
    //                 sval [TW_DC][EW_MAXNM] sorted values;  ssyn [TW_DC][EW_MAXNM] h_d * symbol of the sorted entries;
    //                 skof [TW_DC][EW_MAXNM] the byte offset that syndrome XORs into a lane's address in B (two_copies only)
    static constexpr bool two_copies = sizeof(T) == 4 && Q == 64;
    static constexpr size_t off_b = (size_t)Q * TW_DC * sizeof(T);
    static constexpr size_t off_sval = off_b + (size_t)Q * sizeof(K) * (two_copies ? 2 : 1);
    static constexpr size_t off_ssyn = off_sval + (size_t)TW_DC * EW_MAXNM * sizeof(T);
    static constexpr size_t off_skof = off_ssyn + (size_t)TW_DC * EW_MAXNM * sizeof(int);
    static constexpr size_t raw = off_skof + (two_copies ? (size_t)TW_DC * EW_MAXNM * sizeof(int) : 0);
    static constexpr size_t vec_align = two_copies ? 2 * (size_t)Q * sizeof(T) : ((size_t)Q * sizeof(T) <= 1024 ? (size_t)Q * sizeof(T) : 16);
    static constexpr bool xor_gather = (size_t)Q * sizeof(T) <= 1024;
    static constexpr size_t bytes_per_warp = ((raw + vec_align - 1) / vec_align) * vec_align;
    static constexpr size_t bytes = bytes_per_warp * TW_WARPS;
};

__device__ __forceinline__ float key_value(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }
__device__ __forceinline__ double key_value(unsigned long long k) {
    return __longlong_as_double((long long)((k & 0x8000000000000000ULL) ? (k ^ 0x8000000000000000ULL) : ~k));
}

// VEC consecutive elements (a lane-block of a natural-order vector): one load when the block is 8 bytes or a multiple of 16
template <typename T, int VEC> __device__ __forceinline__ void ld_block(const T *p, T (&v)[VEC]) {
    if constexpr (VEC * sizeof(T) == 8 && sizeof(T) == 4) {
        const float2 w = *reinterpret_cast<const float2 *>(p);
        v[0] = w.x; v[1] = w.y;
    } else ld_vec<T, VEC>(p, v);
}

template <typename T, int VEC> __device__ __forceinline__ void st_block(T *p, const T (&v)[VEC]) {
    if constexpr (VEC * sizeof(T) == 8 && sizeof(T) == 4) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    else st_vec<T, VEC>(p, v);
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
    int *skof = reinterpret_cast<int *>(slab + S::off_skof);        // two_copies only
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
        // SortLLRVector, first n_m places: descending, ties -> larger symbol first.  Only place 0 and the SET of the other
        // n_m - 1 entries matter downstream (a maximum over all combinations), so the r-th winner simply stays in lane r
        // until the list is written once, with its syndrome h * symbol and its address offset in B.
        K rk = K(0); int rs = 0;
        if constexpr (VEC == 2) {
            // two elements per lane (q = 64): the lane keeps them ORDERED (hi first; equal keys: the larger symbol first), so a
            // round is one REDUX over the hi keys, a vote for the highest lane holding the maximum, one shuffle for its symbol,
            // and in that lane "lo moves up"
            const K k0 = order_key(vnew[0]), k1 = order_key(vnew[1]);
            const bool up = k1 >= k0;
            K hi = up ? k1 : k0, lo = up ? k0 : k1;
            int xhi = x0 + (up ? 1 : 0), xlo = x0 + (up ? 0 : 1);
#pragma unroll 1
            for (int r = 0; r < nm; ++r) {
                K kmax; int src;
                warp_argmax<false>(hi, FULL, kmax, src);                                   // highest lane among ties
                const int xw = __shfl_sync(FULL, xhi, src);
                if (lane == src) { hi = lo; xhi = xlo; lo = K(0); }                         // taken
                if (lane == r) { rk = kmax; rs = xw; }
            }
        } else {
        K key[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) key[j] = order_key(vnew[j]);
        K lk = key[0];
#pragma unroll
        for (int j = 1; j < VEC; ++j) lk = key[j] >= lk ? key[j] : lk;
#pragma unroll 1
        for (int r = 0; r < nm; ++r) {
            K kmax; int src;
            warp_argmax<false>(lk, FULL, kmax, src);                                       // highest lane among ties
            int lj = 0;
#pragma unroll
            for (int j = 1; j < VEC; ++j) if (key[j] == kmax) lj = j;                      // larger symbol wins ties
            const int xw = __shfl_sync(FULL, x0 + lj, src);
            if (lane == src) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) if (j == lj) key[j] = K(0);                  // taken
                lk = key[0];
#pragma unroll
                for (int j = 1; j < VEC; ++j) lk = key[j] >= lk ? key[j] : lk;
            }
            if (lane == r) { rk = kmax; rs = xw; }
        }
        }
        if (lane < nm) {
            const int sy = g.pfwd[(size_t)e * Q + rs];
            sval[d * EW_MAXNM + lane] = key_value(rk);
            ssyn[d * EW_MAXNM + lane] = sy;
            if constexpr (S::two_copies) skof[d * EW_MAXNM + lane] = ((sy >> 1) << 3) | ((sy & 1) ? Q * (int)sizeof(T) : 0);
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
    unsigned long long tiv[TW_DC];                                  // the lane's own bytes of the rows h^-1 * x
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) { tiv[d] = 0; if (DC || d < dcn) tiv[d] = ld_bytes<VEC>(g.pinv + (size_t)(e0 + d) * Q + x0); }
    __syncwarp();
    // ---- check node, one output edge at a time (NBLDPC.cpp:883-916): ONE copy of the code, rolled over the outputs ------
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + VEC * lane;
    const int no = dcn - 1;                                         // inputs taking part
    const int last = no - 1;
    const int prefixes = 1 << (lg * (last < 0 ? 0 : last));
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        int o[TW_DC - 1];                                           // the other inputs, ascending
#pragma unroll
        for (int li = 0; li < TW_DC - 1; ++li) o[li] = li < d ? li : li + 1;
#pragma unroll
        for (int j = 0; j < VEC; ++j) Bk[x0 + j] = K(0);            // -inf
        // conf(q,1): input l deviates to any symbol, the others take their best
        T top[TW_DC - 1]; int tsy[TW_DC - 1];
#pragma unroll
        for (int mi = 0; mi < TW_DC - 1; ++mi) { top[mi] = T(0); tsy[mi] = 0; if (mi < no) { top[mi] = sval[o[mi] * EW_MAXNM]; tsy[mi] = ssyn[o[mi] * EW_MAXNM]; } }
        T b1[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) b1[j] = -Real<T>::inf();
#pragma unroll
        for (int li = 0; li < TW_DC - 1; ++li) {
            if (li < no) {
                const int l = o[li];
                int rest = 0;
#pragma unroll
                for (int mi = 0; mi < TW_DC - 1; ++mi) if (mi < no && mi != li) rest ^= tsy[mi];
                // y = h_l^-1 (s ^ rest) = h_l^-1 s ^ h_l^-1 rest
                const unsigned long long tl = pick<unsigned long long, TW_DC>(tiv, l);
                const int yr = row_byte<VEC>(tl, rest);
                const G gv(slab32 + l * VB, yr);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const T vy = gv.at(tl, j);
                    T sum = T(0);
#pragma unroll
                    for (int mi = 0; mi < TW_DC - 1; ++mi) if (mi < no) sum = sum + (mi == li ? vy : top[mi]);
                    b1[j] = fmax(b1[j], sum);
                }
            }
        }
        __syncwarp();
        // conf(n_m, n_c): every combination of the n_m best of the other inputs with at most n_c non-best picks; sums are
        // formed left to right over the inputs.  A sum is never -0 (it starts from +0), so its key needs no normalisation.
        const int ol = o[last < 0 ? 0 : last];
        const T *lv = sval + ol * EW_MAXNM;
        const int *ls = ssyn + ol * EW_MAXNM;
        T bout[VEC];
        if (nc >= no) {
            // n_c does not bind (cfg2: three inputs, n_c = 3): TWO STAGES.  (1) the prefixes (picks of all but the last input)
            // land on P[t] = max over the prefixes of syndrome t -- n_m^(dc-2) atomics instead of n_m^(dc-1); (2) every lane
            // walks the last input's list for its OWN syndromes, B[s] = max_k fl(P[s ^ syn_k] + val_k).  Rounding is monotone,
            // so the maximum over the prefixes commutes with the last addition: fl(max_p p + v) = max_p fl(p + v) --
            // bit-identical results.
#pragma unroll 1
            for (int c = lane; c < prefixes; c += 32) {
                int sy = 0; T sum = T(0);
#pragma unroll
                for (int mi = 0; mi < TW_DC - 2; ++mi) {
                    if (mi < last) {
                        const int k = (c >> (lg * mi)) & (nm - 1);
                        sum = sum + sval[o[mi] * EW_MAXNM + k];
                        sy ^= ssyn[o[mi] * EW_MAXNM + k];
                    }
                }
                K key;
                if constexpr (sizeof(T) == 4) { const unsigned u = __float_as_uint(sum); key = u ^ ((unsigned)((int)u >> 31) | 0x80000000u); }
                else key = order_key(sum);
                atomicMax(&Bk[sy], key);
            }
            __syncwarp();
            T pown[VEC];                                            // my own entries of P back to values, in place
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const K k = Bk[x0 + j];
                pown[j] = k == K(0) ? -Real<T>::inf() : key_value(k);
            }
            st_block<T, VEC>(Bv + x0, pown);
            T b2[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) b2[j] = -Real<T>::inf();
            if constexpr (S::two_copies) {
                // fp32, q = 64 (two elements per lane).  A second copy of P holds every lane's two elements swapped, so the
                // element order syn_k & 1 asks for is a matter of WHICH copy: the address is (B + 8 * lane) ^ skof_k, the pair
                // takes one FADD2, and two entries of the list share one FMNMX3 per output.
                { const T sw[2] = {pown[1], pown[0]}; st_block<T, VEC>(Bv + Q + x0, sw); }
                __syncwarp();
                const unsigned X = slab32 + (unsigned)S::off_b + 8u * (unsigned)lane;
                const int *lo = skof + ol * EW_MAXNM;
                auto pair = [&](int k, bool two) {
                    const T v0 = lv[k], v1 = two ? lv[k + 1] : -Real<T>::inf();
                    Pair2<float> pa, pb;
                    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(pa.v) : "r"(X ^ (unsigned)lo[k]) : "memory");
                    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(pb.v) : "r"(X ^ (unsigned)lo[two ? k + 1 : k]) : "memory");
                    pa = pa.plus(v0); pb = pb.plus(v1);
                    b2[0] = max3(b2[0], pa.lo(), pb.lo());
                    b2[1] = max3(b2[1], pa.hi(), pb.hi());
                };
                int k = 0;
#pragma unroll 1
                for (; k + 8 <= nm; k += 8) {                       // n_m = 8, 16: four pairs per trip, list offsets as immediates
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) pair(k + 2 * kk, true);
                }
#pragma unroll 1
                for (; k < nm; k += 2) pair(k, k + 1 < nm);
            } else {
                __syncwarp();
#pragma unroll 1
                for (int k = 0; k < nm; ++k) {
                    const int sk = ls[k];
                    const T vk = lv[k];
                    T pv[VEC];
                    ld_block<T, VEC>(Bv + VEC * (lane ^ (sk / VEC)), pv);
#pragma unroll
                    for (int bit = 1; bit < VEC; bit <<= 1) {       // element j of my block pairs with element j ^ (sk % VEC) of that one
                        const bool sw = (sk & bit) != 0;
#pragma unroll
                        for (int j = 0; j < VEC; ++j)
                            if (!(j & bit)) { const T lo = pv[j], hi = pv[j | bit]; pv[j] = sw ? hi : lo; pv[j | bit] = sw ? lo : hi; }
                    }
#pragma unroll
                    for (int j = 0; j < VEC; ++j) b2[j] = fmax(b2[j], pv[j] + vk);
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) bout[j] = fmax(b2[j], b1[j]);
            __syncwarp();                                           // everyone has read P
        } else {
            // A lane takes prefixes and walks the last input's list; every leaf lands with an atomic max on its syndrome
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
            __syncwarp();
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const K k = Bk[x0 + j];
                bout[j] = fmax(k == K(0) ? -Real<T>::inf() : key_value(k), b1[j]);
            }
        }
        // merged running output as values (in place), then un-permute and correct
        st_block<T, VEC>(Bv + x0, bout);
        __syncwarp();
        const T b0 = Bv[0];
        const G gb(slab32 + (unsigned)S::off_b, 0);
        const unsigned long long td = pick<unsigned long long, TW_DC>(tfw, d);
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
        if (lane == 0) out[0] = T(0);                               // slot 0 is the implicit LLR of symbol 0
        st_vec<T, VEC>(coutf + (size_t)(e0 + d) * Q, out);
        __syncwarp();                                               // B is re-initialised for the next output
    }
}

}  // namespace nbl
