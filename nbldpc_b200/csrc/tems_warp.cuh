// Fast path of the fused iteration kernel for T-EMS with n_c <= 2 (the headline configuration):
// ONE WARP per (frame, check row), VEC = Q/32 consecutive field elements per lane.
//
// Same results as the generic group kernel (decoder.cu / cn_tems.cuh).  What shaped it, in the order the measurements forced it
// (DESIGN.md 4.1 has the numbers):
//   * INSTRUCTION FETCH.  Unrolled over the lane's elements, the row's edges and the outputs, the kernel was 3 800 SASS
//     instructions (60 KB) and ran at 85-95 % of the request rate of the GPC-level instruction cache
//     (gcc__cache_requests_type_instruction): every round of optimisation that only removed instructions left its duration
//     where it was.  So there is ONE copy of everything now -- the variable-node step is a rolled loop over the edges, the delta
//     domain and the column order a rolled loop over the lane's elements, the pair recovery an out-of-line function -- and both
//     exits of the kernel are provably taken by whole warps (warp index out of a shuffle, latch out of a vote), which stops the
//     compiler from emitting a second, divergent version of every warp collective.  1 960 instructions, 31 KB: the cache
//     requests fell from 5 M to 15 k per launch.
//   * the pair search uses the structure of the allowed sets.  Let c0(x), c1(x) be the cheapest / second cheapest column of
//     eta = x and m0(x) <= m1(x) their costs (TEMS_Min).  For a pair (a, b) with a ^ b = s the cheapest valid assignment of two
//     DIFFERENT columns is m0(a) + m0(b) when c0(a) != c0(b), else min(m0(a) + m1(b), m1(a) + m0(b)); columns beyond the second
//     never win.  So with  D_c[x] = (c0(x) == c ? m1(x) : m0(x))  -- the cheapest column of x other than c --
//         deltaW2[s] = min over a != 0 of  m0(a) + D_{c0(a)}[s ^ a]
//     (255 x q candidates where the reference enumerates ~94 k configurations; n_r = 1: D_c[x] = +inf where c0(x) == c);
//   * and it is PRUNED BY VALUE, exactly: an entry a whose m0(a) exceeds half of the largest deltaW any output can still have
//     cannot be the lighter element of a winning pair and is skipped -- its pairs are reached from their other element (see the
//     search for the one case that needs a second kind of candidate).  The search runs in TWO ROUNDS -- the light entries
//     (m0 <= U0/5) first, then what is still live against the refreshed bound -- and 35-43 of the 255 entries are handled in the
//     first iterations, ~12 in the last; the minima and the recorded pairs are those of the full enumeration;
//   * it is register tiled: lane t owns s = VEC*t .. VEC*t+VEC-1; entry a = VEC*bh + bl reads block k = t ^ bh of D_{c0(a)}
//     (its column is part of the ADDRESS: no dispatch), permuted by bl.  Entries are taken TWO AT A TIME: 16 candidates by 8
//     FADD2 (two sums per instruction, sm_100a), then per output one FMNMX3, one compare and one select of a tag that names both
//     entries (which of the two won is settled once per output, in the path recovery).  For fp32 / q = 256 only bits 0-1 of bl
//     are compile-time constants (four bodies); bit 2 selects the 16-byte piece, i.e. it is part of the address, and the entries
//     of a class reach their bodies through a compacted list of 8-byte records built lane-parallel from two ballots;
//   * tie_mode REFERENCE (template parameter BS == 2): a second pass finds the configurations whose cost EQUALS deltaW and records
//     the one the reference's depth-first enumeration meets first;
//   * no block-level barriers: every exchange between lanes is a __syncwarp on the warp's private smem slab;
//   * all message-sized vectors in shared memory (v2c, D_c, the output trellis) use the same split layout, so that the 128-bit
//     accesses of a lane to its own elements are conflict free; gathers go through table bytes that are stored pre-swizzled to
//     that layout, for the lane's own elements only:  h^-1 (eta ^ beta) = h^-1 eta ^ decide(v2c)  and  h x ^ (syn ^ beta)  need
//     no data-dependent table row, an address is  base ^ (byte ^ key) << 2, and the scalar look-ups beta = h * decide(.) are
//     two shuffles out of registers.
// BS-TEMS (decode method 7) runs on the same kernel with BS = 1: everything around the search is shared, the search itself
// is replaced by the basic-set construction and the walk over its subsets (see the BS block in the kernel).
#pragma once
#include "dev_common.cuh"
#include "iter_params.cuh"
#include <type_traits>

namespace nbl {

constexpr int TW_DC = 4;      // maximum check degree served by the fast path
#ifndef TW_WARPS_VALUE
#define TW_WARPS_VALUE 2
#endif
constexpr int TW_WARPS = TW_WARPS_VALUE;   // warps (= check rows) per CTA
constexpr int TW_DSC = 8;     // ints per edge in the row descriptor: variable, degree, up to 5 edge ids, 1 spare
                              // (spare of edge 0 = first edge id of the row, spare of edge 1 = row degree)
constexpr int TW_HOIST = 2;   // messages of a variable fetched ahead (its degree may be larger: slow loop)
#ifndef TW_MIN_CTAS_VALUE
#define TW_MIN_CTAS_VALUE 5
#endif
constexpr int TW_MIN_CTAS = TW_MIN_CTAS_VALUE;   // resident groups of 4 warps per SM the register allocation aims for (round 2, two-at-a-time search: 5 at 96 registers
                                                 // = 6 at 80 + 3 %; CTAs of 2 warps instead of 4: + 1.5 %)
#ifndef TW_MINB_VALUE
#define TW_MINB_VALUE (TW_MIN_CTAS * 4 / TW_WARPS)
#endif

// A message-sized vector in shared memory is cut into 16-byte pieces; piece h of all 32 lane-blocks is contiguous:
//   index(k, j) = (j / E16) * (32 * E16) + k * E16 + j % E16     (element VEC*k + j; E16 = elements per 16 bytes, or VEC
// when a whole block is <= 16 bytes).  A bit permutation of the element index, hence linear over XOR.
template <typename T, int VEC> struct PLayout {
    static constexpr int E16 = (int)(16 / sizeof(T)) < VEC ? (int)(16 / sizeof(T)) : VEC;
    static constexpr int NP = VEC / E16;                          // pieces per block
    static constexpr int PSTRIDE = 32 * E16 * (int)sizeof(T);     // bytes between pieces
    __host__ __device__ static constexpr int index(int k, int j) { return (j / E16) * (32 * E16) + k * E16 + (j % E16); }
    __host__ __device__ static constexpr int swz(int x) { return index(x / VEC, x % VEC); }
    __device__ static __forceinline__ int index_rt(int k, int j) { return (j / E16) * (32 * E16) + k * E16 + (j % E16); }   // j known at run time only
};

template <typename T, int Q> struct TwSmem {
    // per-warp slab:  R1 [TW_DC][Q]   v2c (VN stage) -> D_c (search) -> output trellis Tout, split layout
    //                 BLK [32]        per eta-block: m0 of its VEC entries (+inf at x = 0), their columns c0 (2 bits each);
    //                                 warp-uniform operands of the search.  The stride is an odd number of 16-byte units so
    //                                 that the per-lane reads of the path recovery (every lane another block) spread out
    //                 M1 [Q]          m1(x), split layout
    //                 bytes c01[Q]    c0 | c1 << 4
    static constexpr int VEC = Q / 32;
    static constexpr size_t blk_raw = ((VEC * sizeof(T) + 4 + 15) / 16) * 16;
    static constexpr size_t blk_stride = (blk_raw / 16) % 2 ? blk_raw : blk_raw + 16;
    static constexpr size_t blk_cls = VEC * sizeof(T);                           // offset of the column word inside a record
    static constexpr size_t off_blk = (size_t)Q * TW_DC * sizeof(T);
    static constexpr size_t off_m1 = off_blk + 32 * blk_stride;
    static constexpr size_t off_c01 = off_m1 + (size_t)Q * sizeof(T);
    // slabs are aligned to one message vector (at most 1 KB) so that a gather address is  base ^ (index * sizeof(T))
    static constexpr size_t vec_align = (size_t)Q * sizeof(T) <= 1024 ? (size_t)Q * sizeof(T) : 16;
    static constexpr bool xor_gather = (size_t)Q * sizeof(T) <= 1024;
    static constexpr size_t bytes_per_warp = ((off_c01 + (size_t)Q + vec_align - 1) / vec_align) * vec_align;
    static constexpr size_t bytes = bytes_per_warp * TW_WARPS;
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
// N consecutive elements as one store when they make 16 or 8 bytes
template <typename T, int N> __device__ __forceinline__ void st_run(T *p, const T (&v)[N]) {
    if constexpr (N * sizeof(T) == 8) {
        unsigned long long w;
        if constexpr (sizeof(T) == 4) w = ((unsigned long long)__float_as_uint((float)v[1]) << 32) | __float_as_uint((float)v[0]);
        else w = (unsigned long long)__double_as_longlong((double)v[0]);
        *reinterpret_cast<unsigned long long *>(p) = w;
    } else st_vec<T, N>(p, v);
}
#ifndef TW_CO_PE
#define TW_CO_PE 4        // elements per trip of the column-order loop (one store per trip and array)
#endif
// a lane's VEC elements of a split-layout vector (block k): NP pieces of 16 bytes, PSTRIDE apart
template <typename T, int VEC> __device__ __forceinline__ void ld_split(const T *__restrict__ vec, int k, T (&v)[VEC]) {
    using PL = PLayout<T, VEC>;
#pragma unroll
    for (int h = 0; h < PL::NP; ++h) {
        T tmp[PL::E16];
        ld_vec<T, PL::E16>(vec + h * (32 * PL::E16) + k * PL::E16, tmp);
#pragma unroll
        for (int i = 0; i < PL::E16; ++i) v[h * PL::E16 + i] = tmp[i];
    }
}
template <typename T, int VEC> __device__ __forceinline__ void st_split(T *__restrict__ vec, int k, const T (&v)[VEC]) {
    using PL = PLayout<T, VEC>;
#pragma unroll
    for (int h = 0; h < PL::NP; ++h) {
        T tmp[PL::E16];
#pragma unroll
        for (int i = 0; i < PL::E16; ++i) tmp[i] = v[h * PL::E16 + i];
        st_vec<T, PL::E16>(vec + h * (32 * PL::E16) + k * PL::E16, tmp);
    }
}
// the same from the byte address of the block's first piece
template <typename T, int VEC> __device__ __forceinline__ void ld_split_bytes(const unsigned char *src, T (&v)[VEC]) {
    using PL = PLayout<T, VEC>;
#pragma unroll
    for (int h = 0; h < PL::NP; ++h) {
        T tmp[PL::E16];
        ld_vec<T, PL::E16>(reinterpret_cast<const T *>(src + h * PL::PSTRIDE), tmp);
#pragma unroll
        for (int i = 0; i < PL::E16; ++i) v[h * PL::E16 + i] = tmp[i];
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
// byte i of a little-endian table word; 32-bit halves so that a constant i costs one instruction
__device__ __forceinline__ int byte_of(unsigned long long w, int i) {
    const unsigned h = i < 4 ? (unsigned)w : (unsigned)(w >> 32);
    return (int)((h >> (8 * (i & 3))) & 255u);
}

#ifndef TW_TH0
#define TW_TH0 0.2       // round 0 of the pair search takes the entries with m0 <= TW_TH0 * U0
#endif
#ifndef TW_ROUNDS
#define TW_ROUNDS 2
#endif
#ifndef TW_TH1
#define TW_TH1 0.3
#endif
#ifndef TW_STATS
#define TW_STATS 0
#endif
#if TW_STATS
__device__ unsigned long long tw_stats[128];     // per iteration: rows, live entries, candidates of the second kind, blocks with a live entry
#endif
// tag of an output of the pruned search: a1 | a2 << 8 = the two entries of the body whose candidate won (a1 == a2: a lone
// entry), TW_SECOND | a | a << 8 = a candidate of the second kind of entry a, TW_NOTAG = the single deviation stands
constexpr int TW_SECOND = 0x10000;
constexpr int TW_NOTAG = 0x20000;
// IEEE division kept out of line (one copy of the routine; the correction factor is 1 in every shipped profile)
template <typename T> __device__ __noinline__ T div_call(T a, T b) { return a / b; }

// shared-memory loads by 32-bit shared-window address (lets the gathers fold base and index into one LOP3)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_addr(unsigned a, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double lds_addr(unsigned a, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
// V[(byte j of w) ^ x] for a vector V at shared address `base` (aligned to its size when XG): one shift, one LOP3, one LDS.
// Bytes and x are split-layout indices (the layout is linear over XOR).
template <typename T, bool XG> struct Gather {
    static constexpr int SH = sizeof(T) == 4 ? 2 : 3;
    unsigned key;                                                    // XG: base ^ (x << SH);  else x << SH
    unsigned base;
    __device__ __forceinline__ Gather(unsigned base_, int x) : key(XG ? (base_ ^ ((unsigned)x << SH)) : ((unsigned)x << SH)), base(base_) {}
    __device__ __forceinline__ T at(unsigned long long w, int j) const {
        const unsigned h = j < 4 ? (unsigned)w : (unsigned)(w >> 32);
        const int jj = j & 3;
        const unsigned b = (jj == 0 ? (h << SH) : (h >> (8 * jj - SH))) & (255u << SH);
        return lds_addr(XG ? (b ^ key) : (base + (b ^ key)), T());
    }
};

// 3-input minimum / maximum: one FMNMX3 on sm_100a for fp32
__device__ __forceinline__ float min3(float a, float b, float c) {
    float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ double min3(double a, double b, double c) { return fmin(fmin(a, b), c); }
__device__ __forceinline__ double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

// two adjacent elements of a D_c vector + one scalar: a single FADD2 on sm_100a for fp32
template <typename T> struct Pair2;
template <> struct Pair2<float> {
    unsigned long long v;
    __device__ __forceinline__ float lo() const { return __uint_as_float((unsigned)v); }
    __device__ __forceinline__ float hi() const { return __uint_as_float((unsigned)(v >> 32)); }
    __device__ __forceinline__ Pair2 plus(float c) const {
        Pair2 r; unsigned long long cc;
        asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(v), "l"(cc));
        return r;
    }
};
template <> struct Pair2<double> {
    double a, b;
    __device__ __forceinline__ double lo() const { return a; }
    __device__ __forceinline__ double hi() const { return b; }
    __device__ __forceinline__ Pair2 plus(double c) const { Pair2 r; r.a = a + c; r.b = b + c; return r; }
};
// best = min(best, c1, c2); the tag is rewritten only when the minimum FALLS (strict: the first of equal costs stays).  fp32: one
// FMNMX3, one FSETP, one SEL per output for two candidates.
__device__ __forceinline__ void min_tag2(float &best, int &tag, float c1, float c2, int tg) {
    asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tmin.f32 t, %1, %2, %3;\n\tsetp.lt.f32 p, t, %1;\n\tselp.b32 %0, %4, %0, p;\n\tmov.f32 %1, t;\n\t}"
        : "+r"(tag), "+f"(best) : "f"(c1), "f"(c2), "r"(tg));
}
__device__ __forceinline__ void min_tag2(double &best, int &tag, double c1, double c2, int tg) {
    const double t = fmin(fmin(best, c1), c2);
    if (t < best) tag = tg;
    best = t;
}
// 16 bytes of a split-layout vector as two FADD2 operands, by 32-bit shared address
__device__ __forceinline__ void lds_pairs(unsigned addr, Pair2<float> &p0, Pair2<float> &p1) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(p0.v), "=l"(p1.v) : "r"(addr) : "memory");
}
// the VEC/2 pairs of block k of a split-layout vector from shared memory (src = first piece of the block)
template <int VEC> __device__ __forceinline__ void ld_pairs(const unsigned char *src, Pair2<float> (&p)[VEC / 2]) {
    if constexpr (VEC >= 4) {
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {
            const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(src + h * 512);
            p[2 * h].v = w.x; p[2 * h + 1].v = w.y;
        }
    } else p[0].v = *reinterpret_cast<const unsigned long long *>(src);
}
template <int VEC> __device__ __forceinline__ void ld_pairs(const unsigned char *src, Pair2<double> (&p)[VEC / 2]) {
#pragma unroll
    for (int h = 0; h < VEC / 2; ++h) {
        const double2 w = *reinterpret_cast<const double2 *>(src + h * 512);
        p[h].a = w.x; p[h].b = w.y;
    }
}

template <typename T, int N> __device__ __forceinline__ T vec_max(const T (&v)[N]) {
    if constexpr (N == 8) return max3(max3(v[0], v[1], v[2]), max3(v[3], v[4], v[5]), fmax(v[6], v[7]));
    else if constexpr (N == 4) return fmax(max3(v[0], v[1], v[2]), v[3]);
    else if constexpr (N == 2) return fmax(v[0], v[1]);
    else return v[0];
}

// v[r] for an index known at run time only (N a power of two): a tree of N - 1 selects, no local memory
template <typename V, int N> __device__ __forceinline__ V pick(const V (&v)[N], int r) {
    if constexpr (N == 1) return v[0];
    else {
        V t[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) t[i] = (r & 1) ? v[2 * i + 1] : v[2 * i];
        return pick<V, N / 2>(t, r >> 1);
    }
}
// warp maximum of non-negative values (they order like their bit patterns): one REDUX for fp32
__device__ __forceinline__ float warp_max_nonneg(float v, float) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v))); }
__device__ __forceinline__ double warp_max_nonneg(double v, double) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max_nonneg(T v) { return warp_max_nonneg(v, T()); }

// DecideLLRVector over a vector spread VEC-per-lane across the warp (x = VEC*lane + j): the largest strictly positive
// entry, ties to the lowest symbol, 0 when nothing is positive.
template <typename T, int VEC> __device__ __forceinline__ int warp_decide(const T (&v)[VEC], int lane) {
    using K = typename Real<T>::key_t;
    const T lm = vec_max<T, VEC>(v);
    K kmax; int src;
    warp_argmax<true>(pos_key<T>(lm), 0xffffffffu, kmax, src);      // lowest lane holding the maximum
    int idx = VEC - 1;
#pragma unroll
    for (int j = VEC - 2; j >= 0; --j) if (Real<T>::bits(v[j]) == kmax) idx = j;   // lowest element inside that lane
    const int x = __shfl_sync(0xffffffffu, VEC * lane + idx, src);
    return kmax == K(0) ? 0 : x;
}

// byte `i` (runtime) of the table row whose bytes VEC*l .. VEC*l+VEC-1 lane l holds in `w`
template <int VEC> __device__ __forceinline__ int row_byte(unsigned long long w, int i) {
    const int src = i / VEC, sh = 8 * (i % VEC);
    unsigned lo = __shfl_sync(0xffffffffu, (unsigned)w, src);
    if constexpr (VEC == 8) {
        const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(w >> 32), src);
        lo = (sh & 32) ? hi : lo;
    }
    return (int)((lo >> (sh & 31)) & 255u);
}

// per eta-block record: VEC costs + column word
template <typename T, int VEC> __device__ __forceinline__ void ld_blk(const unsigned char *rec, T (&cost)[VEC], unsigned &cls) {
    ld_vec<T, VEC>(reinterpret_cast<const T *>(rec), cost);
    cls = *reinterpret_cast<const unsigned *>(rec + VEC * sizeof(T));
}

// The pair behind a tag of the pruned search, for the output s (split-layout index sidx) whose deltaW is dw: element a in column
// ca at cost ua and b = s ^ a in column cb at cost ub.  Returns the packed path  sb | sa << 8 | ca << 16 | cb << 18 | (deviating
// columns) << 20 | 1 << 24  and the two landing values dw - ua, dw - ub.  Out of line on purpose: ONE copy for a lane's VEC outputs.
template <typename T> struct TwPair { unsigned path; T va, vb; };
template <typename T, int Q>
__device__ __noinline__ TwPair<T> tw_pair_of_tag(const unsigned char *slab, int tg, int s, unsigned sidx, T dw) {
    constexpr int VEC = Q / 32;
    using PL = PLayout<T, VEC>;
    using S = TwSmem<T, Q>;
    const unsigned char *blk = slab + S::off_blk;
    const T *sM1 = reinterpret_cast<const T *>(slab + S::off_m1);
    const uint8_t *c01 = slab + S::off_c01;
    int aa = tg & 255;
    const int a2 = (tg >> 8) & 255;
    if (a2 != aa) {                                                 // which entry of the body attained dw: the first one unless its candidate differs
        const T c = *reinterpret_cast<const T *>(blk + (aa / VEC) * (int)S::blk_stride + (aa % VEC) * (int)sizeof(T)) +
                    reinterpret_cast<const T *>(slab)[(c01[aa] & 15) * Q + (int)(sidx ^ (unsigned)PL::swz(aa))];
        if (c != dw) aa = a2;
    }
    const int bb = s ^ aa, ka = c01[aa], kb = c01[bb];
    const unsigned sa = (unsigned)PL::swz(aa), sb = sidx ^ sa;          // the layout is linear over XOR
    const T m0b = *reinterpret_cast<const T *>(blk + (bb / VEC) * (int)S::blk_stride + (bb % VEC) * (int)sizeof(T));
    int ca, cb; T ua, ub;
    if (!(tg & TW_SECOND)) {                                        // a pays its cheapest column, b the cheapest one that is not ca
        ca = ka & 15;
        const bool same = (kb & 15) == ca;
        cb = same ? (kb >> 4) : (kb & 15);
        ua = *reinterpret_cast<const T *>(blk + (aa / VEC) * (int)S::blk_stride + (aa % VEC) * (int)sizeof(T));
        ub = same ? sM1[sb] : m0b;                                  // D_ca[b]
    } else {                                                        // a pays its SECOND-cheapest column, b its cheapest one (= c0(a))
        ca = ka >> 4; cb = ka & 15;
        ua = sM1[sa];
        ub = m0b;
    }
    TwPair<T> out;
    out.path = sb | (sa << 8) | ((unsigned)ca << 16) | ((unsigned)cb << 18) | (0x100000u << ca) | (0x100000u << cb) | (1u << 24);
    out.va = dw - ua; out.vb = dw - ub;
    return out;
}

// Rank of a configuration in the reference's depth-first enumeration (smaller = met earlier): first deviating column fc (later
// is earlier), its eta fv (ascending), second deviating column sc (4 = none, which comes first; else later is earlier), its eta.
__device__ __forceinline__ unsigned tw_key(int fc, int fv, int sc, int sv) {
    return ((unsigned)(3 - fc) << 19) | ((unsigned)fv << 11) | ((unsigned)(4 - sc) << 8) | (unsigned)sv;
}
// All column assignments of the pair {a, b = s ^ a} whose cost equals dw exactly; the one the reference records.  key = ~0: none.
template <typename T> struct TwRef { unsigned key, path; T va, vb; };
template <typename T, int Q>
__device__ __noinline__ TwRef<T> tw_ref_pair(const unsigned char *slab, int a, int s, unsigned sidx, T dw, bool second) {
    constexpr int VEC = Q / 32;
    using PL = PLayout<T, VEC>;
    using S = TwSmem<T, Q>;
    const unsigned char *blk = slab + S::off_blk;
    const T *sM1 = reinterpret_cast<const T *>(slab + S::off_m1);
    const uint8_t *c01 = slab + S::off_c01;
    const int b = s ^ a, ka = c01[a], kb = c01[b];
    const unsigned sa = (unsigned)PL::swz(a), sb = sidx ^ sa;           // the layout is linear over XOR
    const T ua[2] = {*reinterpret_cast<const T *>(blk + (a / VEC) * (int)S::blk_stride + (a % VEC) * (int)sizeof(T)), sM1[sa]};
    const T ub[2] = {*reinterpret_cast<const T *>(blk + (b / VEC) * (int)S::blk_stride + (b % VEC) * (int)sizeof(T)), sM1[sb]};
    TwRef<T> out; out.key = 0xffffffffu; out.path = 0u; out.va = T(0); out.vb = T(0);
    if (b == 0) return out;
    const int lim = second ? 2 : 1;
    for (int i = 0; i < lim; ++i)
        for (int j = 0; j < lim; ++j) {
            const int ca = i ? (ka >> 4) : (ka & 15), cb = j ? (kb >> 4) : (kb & 15);
            if (ca == cb || !(ua[i] + ub[j] == dw)) continue;
            const unsigned key = ca < cb ? tw_key(ca, a, cb, b) : tw_key(cb, b, ca, a);
            if (key < out.key) {
                out.key = key;
                out.path = sb | (sa << 8) | ((unsigned)ca << 16) | ((unsigned)cb << 18) | (0x100000u << ca) | (0x100000u << cb) | (1u << 24);
                out.va = dw - ua[i]; out.vb = dw - ub[j];
            }
        }
    return out;
}

// DC > 0: every check row has exactly DC edges (no degree guards in the unrolled stages); DC == 0: row degree from the
// descriptor.  DV > 0: every variable has exactly DV <= TW_HOIST edges; DV == 0: variable degree from the descriptor.
// BS == 2: T-EMS with the reference's tie order among configurations of equal cost (second pass after the search, see there).
// BS == 1: BS-TEMS (decode method 7) -- same variable-node step, delta domain, column order, fill-in and output stage; the
// pair search is replaced by the basic-set construction and the walk over its subsets (see the BS block below).
template <typename T, int Q, int DC, int DV, int BS = 0>
__global__ void __launch_bounds__(32 * TW_WARPS, sizeof(T) == 4 ? (BS == 1 ? 5 * 4 / TW_WARPS : TW_MINB_VALUE) : 1) k_iterate_tems_warp(const IterParams<T> P) {
    constexpr int VEC = Q / 32;
    using PL = PLayout<T, VEC>;
    using S = TwSmem<T, Q>;
    using G = Gather<T, S::xor_gather>;
    constexpr int VB = (int)(Q * sizeof(T));                        // bytes of one message vector
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const DevGraph &g = P.g;
    // The warp index comes out of a shuffle and the latch out of a vote, so that the compiler KNOWS both exits are taken by whole
    // warps: otherwise every warp collective below is compiled twice (a convergence test plus a divergent fall-back path).
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int f = (int)(blockIdx.z * 65535u + blockIdx.y), row = (int)blockIdx.x * TW_WARPS + warp;   // see launch_iterate_warp_dc
    if (f >= P.B || row >= g.M) return;
    if (P.early_stop && __any_sync(0xffffffffu, P.done[f] != 0)) return;

    unsigned char *slab = smem_raw + S::bytes_per_warp * warp;
    T *R1 = reinterpret_cast<T *>(slab);
    unsigned char *blk = slab + S::off_blk;
    T *sM1 = reinterpret_cast<T *>(slab + S::off_m1);
    uint8_t *c01 = slab + S::off_c01;
    const unsigned slab32 = smem_u32(slab);
    const T INF = Real<T>::inf();

    // Row descriptor: per edge {variable, its degree, the edge ids of its messages in summation order}; every lane reads
    // the same 16 bytes (one broadcast transaction), so the message loads do not wait on a chain of index loads.
    const int *rd = g.rowdesc + (size_t)row * 32;
    const int e0 = DC ? DC * row : rd[TW_DSC - 1];
    const int dcn = DC ? DC : rd[2 * TW_DSC - 1];
    const T *lchf = P.lch + (size_t)f * g.N * Q + VEC * lane;        // this frame, this lane's elements
    const T *cinf = P.c2v_in + (size_t)f * g.E * Q + VEC * lane;
    T *v2cf = P.v2c + (size_t)f * g.E * Q + VEC * lane;
    uint8_t *vdecf = P.vdec + (size_t)f * g.E;
    const int x0 = VEC * lane;
    int syn_dec = 0, syn = 0;
    unsigned sdecw = 0, sbetaw = 0;                                 // byte d: decide(v2c_d), beta_d = h_d * decide(v2c_d)
    // ---- steps A-C for the row's variables (NBLDPC.cpp:977-1052) -----------------------------------------------------------
    // ONE copy of the code, rolled over the row's edges (the kernel is bound by instruction fetch: see DESIGN.md): the inputs of
    // edge d + 1 are requested into the SAME registers as soon as edge d has formed its sums, so they are in flight during the
    // decisions, the stores and the table look-ups of edge d.
    T a0[VEC], a1[VEC], a2[VEC];                                    // channel LLRs, first and second message of the variable
    int4 nd; int nbefore;
    unsigned long long ntf;                                         // the lane's own bytes of the edge's h * x row
    auto fetch = [&](int d) {
        nd = *reinterpret_cast<const int4 *>(rd + TW_DSC * d);      // {variable, degree, edge id 0, edge id 1}
        ld_vec<T, VEC>(lchf + (size_t)nd.x * Q, a0);
        ntf = ld_bytes<VEC>(g.pfwd + (size_t)(e0 + d) * Q + x0);
        const int nv = DV ? DV : nd.y;
        if (!P.first && 0 < nv) ld_vec<T, VEC>(cinf + (size_t)nd.z * Q, a1);
        else {                                                      // iteration 1: c2v == 0 (x + 0 and x - 0 are exact)
#pragma unroll
            for (int i = 0; i < VEC; ++i) a1[i] = T(0);
        }
        if (!P.first && 1 < nv) ld_vec<T, VEC>(cinf + (size_t)nd.w * Q, a2);
        else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) a2[i] = T(0);
        }
        nbefore = P.first ? -1 : (int)vdecf[e0 + d];
    };
    fetch(0);
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        const int e = e0 + d, col = nd.x, nv = DV ? DV : nd.y, before = nbefore;
        const unsigned long long tf = ntf;
        const bool mine0 = nd.z == e;                               // warp-uniform
        T post[VEC], vnew[VEC], own[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            post[i] = a0[i] + a1[i];
            own[i] = mine0 ? a1[i] : a2[i];
        }
        if (DV ? DV > 1 : nv > 1) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) post[i] = post[i] + a2[i];
        } else if (!mine0) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) own[i] = T(0);
        }
        if (!DV && !P.first) {
#pragma unroll 1
            for (int j = TW_HOIST; j < nv; ++j) {                   // variables of degree > TW_HOIST
                const int ej = rd[TW_DSC * d + 2 + j];
                T mj[VEC];
                ld_vec<T, VEC>(cinf + (size_t)ej * Q, mj);
#pragma unroll
                for (int i = 0; i < VEC; ++i) { post[i] = post[i] + mj[i]; if (ej == e) own[i] = mj[i]; }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) vnew[i] = post[i] - own[i];
        if (d + 1 < dcn) fetch(d + 1);                              // a0 .. a2 are consumed: the next edge's loads go out now
        const int dpost = warp_decide<T, VEC>(post, lane);
        int dfin = warp_decide<T, VEC>(vnew, lane);
        if (!P.first && P.w_old != T(0) && dfin != before) {        // warp-uniform
            T vold[VEC];
            ld_vec<T, VEC>(v2cf + (size_t)e * Q, vold);
#pragma unroll
            for (int i = 0; i < VEC; ++i) vnew[i] = P.w_old * vold[i] + P.w_new * vnew[i];   // -fmad=false
            dfin = warp_decide<T, VEC>(vnew, lane);
        }
        st_vec<T, VEC>(v2cf + (size_t)e * Q, vnew);
        st_split<T, VEC>(R1 + d * Q, lane, vnew);
        const int bt = row_byte<VEC>(tf, dfin);
        if (lane == 0) { vdecf[e] = (uint8_t)dfin; P.dec[(size_t)f * g.N + col] = (uint8_t)dpost; }
        sdecw |= (unsigned)dfin << (8 * d); sbetaw |= (unsigned)bt << (8 * d);
        syn ^= bt;
        int bp = bt;
        if (dpost != dfin) bp = row_byte<VEC>(tf, dpost);           // warp-uniform
        syn_dec ^= bp;
    }
    if (lane == 0 && syn_dec != 0) P.flag[f] = 1;
    // the lane's own bytes of the rows h^-1 * x, as split-layout indices
    unsigned long long tiv[TW_DC];
#pragma unroll
    for (int d = 0; d < TW_DC; ++d) { tiv[d] = 0; if (DC || d < dcn) tiv[d] = ld_bytes<VEC>(P.pinv_sw + (size_t)(e0 + d) * Q + x0); }
    __syncwarp();
    // ---- delta domain (TEMS_Get_deltaU): dU[d][eta] = max_d - V_d[h^-1 (eta ^ beta_d)],  h^-1 (eta ^ beta_d) = h^-1 eta ^ decide,
    //      and the column order per eta (TEMS_Get_Min): c0, c1, m0, m1 -- rolled over the lane's VEC elements (one copy of the
    //      four gathers and of the ordering network); the results go to the block records, M1 and c01, none of which the gathers read
    {
        unsigned key[TW_DC]; T vmx[TW_DC];
#pragma unroll
        for (int d = 0; d < TW_DC; ++d) {
            key[d] = 0u; vmx[d] = INF;
            if (DC || d < dcn) {
                const int sd = PL::swz((sdecw >> (8 * d)) & 255);
                vmx[d] = R1[d * Q + sd];
                key[d] = S::xor_gather ? ((slab32 + d * VB) ^ ((unsigned)sd << G::SH)) : ((unsigned)sd << G::SH);
            }
        }
        unsigned char *mine = blk + lane * (int)S::blk_stride;
        unsigned cls = 0;
        // rolled over the 16-byte pieces of the lane's block, unrolled inside a piece: the piece's m0, m1 and column bytes leave
        // as ONE store each (a scalar store per element meets 4-way bank conflicts on the 48-byte / 16-byte lane strides)
        constexpr int PE = TW_CO_PE < PL::E16 ? TW_CO_PE : PL::E16;
#pragma unroll 1
        for (int jp = 0; jp < VEC; jp += PE) {
            T m0p[PE], m1p[PE]; unsigned ckp = 0u;
#pragma unroll
            for (int jj = 0; jj < PE; ++jj) {
                const int j = jp + jj;
                T uu[TW_DC];
#pragma unroll
                for (int d = 0; d < TW_DC; ++d) {
                    uu[d] = INF;
                    if (DC || d < dcn) {
                        const unsigned bb = ((unsigned)(tiv[d] >> (8 * j)) & 255u) << G::SH;
                        uu[d] = vmx[d] - lds_addr(S::xor_gather ? (bb ^ key[d]) : (slab32 + d * VB + (bb ^ key[d])), T());
                    }
                }
                const T a = uu[0], b = uu[1], c = uu[2], dd = uu[3];
                const T m0 = fmin(min3(a, b, c), dd);
                const int k0 = (a == m0) ? 0 : (b == m0) ? 1 : (c == m0) ? 2 : 3;               // stable: the lower column wins ties
                const T a1 = k0 == 0 ? INF : a, b1 = k0 == 1 ? INF : b, c1 = k0 == 2 ? INF : c, d1 = k0 == 3 ? INF : dd;
                const T m1 = fmin(min3(a1, b1, c1), d1);
                int k1 = (k0 != 3) ? 3 : 2;                         // lowest column other than k0 that attains m1 (row degree >= 2)
                if (c1 == m1) k1 = 2;
                if (b1 == m1) k1 = 1;
                if (a1 == m1) k1 = 0;
                m0p[jj] = (x0 + j == 0) ? INF : m0;                 // the search never takes eta = 0
                m1p[jj] = m1;
                ckp |= (unsigned)(k0 | (k1 << 4)) << (8 * jj);
                cls |= (unsigned)k0 << (2 * j);
            }
            st_run<T, PE>(reinterpret_cast<T *>(mine + jp * (int)sizeof(T)), m0p);
            st_run<T, PE>(sM1 + PL::index_rt(lane, jp), m1p);
            st_bytes<PE>(c01 + x0 + jp, ckp);
        }
        *reinterpret_cast<unsigned *>(mine + S::blk_cls) = cls;
    }
    __syncwarp();                                                  // all reads of v2c in R1 are done; records, M1 and c01 are visible
    if constexpr (BS != 1) {
        // D_c[x]: cheapest allowed column of x other than c (the candidate vectors of the search), over the dead v2c
        T m0[VEC], m1[VEC];
        ld_vec<T, VEC>(reinterpret_cast<const T *>(blk + lane * (int)S::blk_stride), m0);
        ld_split<T, VEC>(sM1, lane, m1);
        const unsigned long long cw = ld_bytes<VEC>(c01 + x0);
        const T second_ok = P.p0 >= 2 ? T(0) : INF;                 // n_r = 1: only the cheapest column may deviate
#pragma unroll
        for (int c = 0; c < TW_DC; ++c) {
            if (DC || c < dcn) {
                T dcv[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const T v = ((byte_of(cw, j) & 15) == c) ? (m1[j] + second_ok) : m0[j];
                    dcv[j] = (x0 + j == 0) ? INF : v;
                }
                st_split<T, VEC>(R1 + c * Q, lane, dcv);
            }
        }
    }
    __syncwarp();
    const unsigned long long c01w = ld_bytes<VEC>(c01 + x0);
    T *Tout = R1;
    if constexpr (BS != 1) {
        // ---- pair search: deltaW[s] = min(m0(s), min_a m0(a) + D_{c0(a)}[s ^ a]), with EXACT pruning by value ---------------------
        // best starts at the single-deviation cost m0(s), so it is at all times an upper bound of deltaW[s]; a tag is written only
        // when a pair is STRICTLY cheaper.  tag = a1 | a2 << 8: the two entries of the body whose candidate (first kind) won;
        // TW_SECOND | a | a << 8: a candidate of the second kind of entry a; TW_NOTAG: the single deviation stands.
        //
        // Let U >= max_s deltaW[s] (any upper bound: the running max of best).  A pair {x, y} can only be the cost of some s if
        // cost <= U, and cost >= m0(x) + m0(y), so its lighter element has m0 <= U/2: an entry a with m0(a) > U/2 ("dead") is
        // skipped altogether, because every pair it belongs to is also reached from its partner --
        //   * different cheapest columns: m0(x) + m0(y) is the same sum from either side;
        //   * same cheapest column c: the pair costs min(m0(x) + m1(y), m1(x) + m0(y)); the first kind of candidate, m0(a) +
        //     D_c[.], covers the term in which the LIVE element pays m0; the term in which the live element a pays m1 while its
        //     dead partner pays m0 is covered by a second kind of candidate, m1(a) + (m0(y) if c0(y) == c else inf), needed only
        //     when m1(a) <= U/2 as well (else the partner is live and reaches the pair itself).
        // Same minima, same recorded pairs as the full enumeration (no two configurations tie exactly); the full enumeration with
        // its four-way column dispatch per entry, and the one-entry-at-a-time form with a bound refresh per block range, which this
        // replaces, are in the history of this file.  A live entry loads the block of ITS column only (the address carries the
        // column).
        T best[VEC]; int tag[VEC];
        const unsigned char *myrec = blk + lane * (int)S::blk_stride;
        ld_vec<T, VEC>(reinterpret_cast<const T *>(myrec), best);
        if (lane == 0) best[0] = T(0);
#pragma unroll
        for (int r = 0; r < VEC; ++r) tag[r] = TW_NOTAG;
        unsigned done = 0u;                                             // bit j: first kind handled, bit 8 + j: second kind handled (my block)
        if (P.p1 >= 2) {
            const bool second = P.p0 >= 2;                              // n_r = 1: the second-cheapest column may not deviate
            // TWO ROUNDS.  Round 0 takes the light entries, m0 <= U0/5 (U0 = max_s m0(s), the bound before any pair is known): they
            // are the ones that pull the bound down.  Round 1 takes what is still live against the refreshed bound, m0 <= U1/2, and
            // has not been handled.  Any schedule is exact as long as every entry with m0 <= U/2 for the FINAL bound U has been
            // handled at the end (U only falls); this one keeps fewer entries live than a refresh per block range (offline on dumped
            // rows: 51 against 68 in iteration 2, 72 against 79 later), and -- the point -- it leaves enough entries per round that
            // they can be taken TWO AT A TIME: within a round the entries with the same position bl inside their eta-block (bl fixes
            // the register permutation, so it must be a compile-time constant of the body) come out of one ballot, and a body
            // handles two of them: 16 candidates by 8 FADD2, one FMNMX3 + one compare + one predicated tag write per output, where
            // one entry at a time costs a compare and two selects per candidate.  The tag names both entries; which of the two
            // attained the minimum is settled once per output in tw_pair_of_tag.
#pragma unroll 1
            T U0 = T(0);
            for (int round = 0; round < TW_ROUNDS; ++round) {
                const T U = warp_max_nonneg<T>(vec_max<T, VEC>(best));
                if (round == 0) U0 = U;
                const T th = round == TW_ROUNDS - 1 ? T(0.5) * U : round == 0 ? T(TW_TH0) * U : fmin(T(0.5) * U, T(TW_TH1) * U0);
                unsigned mymask = 0u;
                T c0v[VEC];                                             // m0 of my block's entries (kept for the entry list)
                ld_vec<T, VEC>(reinterpret_cast<const T *>(myrec), c0v);                           // +inf at eta = 0: never live
                {
                    T c1v[VEC];
                    ld_split<T, VEC>(sM1, lane, c1v);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        if (c0v[j] <= th) mymask |= 1u << j;
                        if (second && c1v[j] <= th && c0v[j] <= th) mymask |= 256u << j;
                    }
                }
                mymask &= ~done; done |= mymask;
                if constexpr (sizeof(T) == 4 && VEC == 8) {
                    // fp32, q = 256.  Of the three bits of bl only the lower two permute registers (which pair of a 16-byte piece, which
                    // half of the pair); bit 2 selects the PIECE, i.e. it is part of the address.  So there are four bodies (classes
                    // bc = bl & 3), each serving both pieces, and a class has enough entries per round to fill its bodies in pairs.
                    // The entries of a class are compacted into a list of 8-byte records {m0(a), w} first -- lane-parallel: every lane
                    // (= eta-block) places its up to two entries by a prefix count over two ballots -- with
                    //   w = bh << 4 | (bl >> 2) << 9 | c0(a) << 10 | a << 16:  block address of D_c0(a) = ((slab + lane * 16) ^ (w & 0x3f0)) + (w & 0xc00),
                    // so that a body costs one broadcast LDS.128 for both records and three instructions per address.  The list lives
                    // in the c01 area (+ the slab's spare bytes); c01 is put back from registers after the search.
                    static_assert(S::bytes_per_warp - S::off_c01 >= 512, "the entry list needs 64 records");
                    uint2 *list = reinterpret_cast<uint2 *>(slab + S::off_c01);
                    const unsigned list32 = slab32 + (unsigned)S::off_c01, X = slab32 + ((unsigned)lane << 4);
                    const unsigned lt = (1u << lane) - 1u;
                    const unsigned mycls = *reinterpret_cast<const unsigned *>(myrec + S::blk_cls);
#pragma unroll
                    for (int bc = 0; bc < 4; ++bc) {
                        const bool lA = (mymask >> bc) & 1u, lB = (mymask >> (bc + 4)) & 1u;
                        const unsigned mA = __ballot_sync(0xffffffffu, lA), mB = __ballot_sync(0xffffffffu, lB);
                        const int nA = __popc(mA), n = nA + __popc(mB);
                        if (n == 0) continue;                           // warp-uniform
#if TW_STATS
                        if (lane == 0) { atomicAdd(tw_stats + 8 * (P.iter & 15) + 1, (unsigned long long)n); atomicAdd(tw_stats + 8 * (P.iter & 15) + 3, (unsigned long long)((n + 1) / 2)); }
#endif
                        if (lA) list[__popc(mA & lt)] = make_uint2(__float_as_uint(c0v[bc]),
                                                                   ((unsigned)lane << 4) | (((mycls >> (2 * bc)) & 3u) << 10) | ((unsigned)(VEC * lane + bc) << 16));
                        if (lB) list[nA + __popc(mB & lt)] = make_uint2(__float_as_uint(c0v[bc + 4]),
                                                                        ((unsigned)lane << 4) | 512u | (((mycls >> (2 * bc + 8)) & 3u) << 10) | ((unsigned)(VEC * lane + bc + 4) << 16));
                        if (lane == 0 && (n & 1)) list[n] = make_uint2(__float_as_uint(INF), 0u);   // a lone last entry: its partner's candidates are +inf
                        __syncwarp();
#pragma unroll 1
                        for (int i = 0; i < n; i += 2) {
                            unsigned c1b, w1, c2b, w2;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c1b), "=r"(w1), "=r"(c2b), "=r"(w2) : "r"(list32 + 8u * (unsigned)i) : "memory");
                            const unsigned ad1 = (X ^ (w1 & 0x3f0u)) + (w1 & 0xc00u), ad2 = (X ^ (w2 & 0x3f0u)) + (w2 & 0xc00u);
                            const int tg = (int)__byte_perm(w1, w2, 0x3362);    // a1 | a2 << 8
                            const float cost1 = __uint_as_float(c1b), cost2 = __uint_as_float(c2b);
                            Pair2<float> p1[4], p2[4];                  // [0..1]: the piece that feeds outputs 0..3, [2..3]: outputs 4..7
                            lds_pairs(ad1, p1[0], p1[1]); lds_pairs(ad1 ^ 512u, p1[2], p1[3]);
                            lds_pairs(ad2, p2[0], p2[1]); lds_pairs(ad2 ^ 512u, p2[2], p2[3]);
#pragma unroll
                            for (int k = 0; k < 4; ++k) { p1[k] = p1[k].plus(cost1); p2[k] = p2[k].plus(cost2); }
#pragma unroll
                            for (int r = 0; r < VEC; ++r) {
                                const int j = (r & 4) | ((r ^ bc) & 3);
                                const float c1 = (j & 1) ? p1[j >> 1].hi() : p1[j >> 1].lo();
                                const float c2 = (j & 1) ? p2[j >> 1].hi() : p2[j >> 1].lo();
                                min_tag2(best[r], tag[r], c1, c2, tg);
                            }
                        }
                        __syncwarp();                                   // the list is rewritten by the next class
                    }
                } else {
#pragma unroll
                for (int bl = 0; bl < VEC; ++bl) {
                    unsigned mb = __ballot_sync(0xffffffffu, (mymask >> bl) & 1u);                 // bit bh: entry VEC * bh + bl is to be handled
#pragma unroll 1
                    while (mb) {                                        // warp-uniform
                        const int bh1 = __ffs((int)mb) - 1; mb &= mb - 1u;
                        const bool two = mb != 0u;
                        const int bh2 = two ? __ffs((int)mb) - 1 : bh1; mb &= mb - 1u;
#if TW_STATS
                        if (lane == 0) { atomicAdd(tw_stats + 8 * (P.iter & 15) + 1, two ? 2ULL : 1ULL); atomicAdd(tw_stats + 8 * (P.iter & 15) + 3, 1ULL); }
#endif
                        const unsigned char *rec1 = blk + bh1 * (int)S::blk_stride, *rec2 = blk + bh2 * (int)S::blk_stride;
                        const T cost1 = *reinterpret_cast<const T *>(rec1 + bl * (int)sizeof(T));
                        const T cost2r = *reinterpret_cast<const T *>(rec2 + bl * (int)sizeof(T));
                        const T cost2 = two ? cost2r : INF;             // a lone entry: its partner's candidates are +inf
                        const unsigned col1 = (*reinterpret_cast<const unsigned *>(rec1 + S::blk_cls) >> (2 * bl)) & 3u;
                        const unsigned col2 = (*reinterpret_cast<const unsigned *>(rec2 + S::blk_cls) >> (2 * bl)) & 3u;
                        const unsigned char *src1 = slab + (lane ^ bh1) * (PL::E16 * (int)sizeof(T)) + col1 * VB;
                        const unsigned char *src2 = slab + (lane ^ bh2) * (PL::E16 * (int)sizeof(T)) + col2 * VB;
                        const int tg = (VEC * bh1 + bl) | ((VEC * bh2 + bl) << 8);
                        if constexpr (VEC >= 2) {
                            Pair2<T> p1[VEC / 2], p2[VEC / 2];
                            ld_pairs<VEC>(src1, p1);
                            ld_pairs<VEC>(src2, p2);
#pragma unroll
                            for (int i = 0; i < VEC / 2; ++i) { p1[i] = p1[i].plus(cost1); p2[i] = p2[i].plus(cost2); }
#pragma unroll
                            for (int r = 0; r < VEC; ++r) {
                                const int j = r ^ bl;
                                const T c1 = (j & 1) ? p1[j >> 1].hi() : p1[j >> 1].lo();
                                const T c2 = (j & 1) ? p2[j >> 1].hi() : p2[j >> 1].lo();
                                min_tag2(best[r], tag[r], c1, c2, tg);
                            }
                        } else {
                            const T c1 = *reinterpret_cast<const T *>(src1) + cost1, c2 = *reinterpret_cast<const T *>(src2) + cost2;
                            min_tag2(best[0], tag[0], c1, c2, tg);
                        }
                    }
                }
                }
                // second kind (rare): ONE copy of the code for any entry; the tile is brought into the order of the lane's outputs
                // by three conditional exchange steps (index j = r ^ bl)
                unsigned b2 = __ballot_sync(0xffffffffu, (mymask >> 8) != 0u);
#pragma unroll 1
                while (b2) {
                    const int bh = __ffs((int)b2) - 1; b2 &= b2 - 1u;
                    const unsigned char *rec = blk + bh * (int)S::blk_stride;
                    const unsigned cls = *reinterpret_cast<const unsigned *>(rec + S::blk_cls);
                    const unsigned char *Dk = slab + (lane ^ bh) * (PL::E16 * (int)sizeof(T));
                    for (unsigned m2 = __shfl_sync(0xffffffffu, mymask, bh) >> 8; m2; m2 &= m2 - 1u) {
                        const int bl = __ffs((int)m2) - 1, a = VEC * bh + bl;
#if TW_STATS
                        if (lane == 0) atomicAdd(tw_stats + 8 * (P.iter & 15) + 2, 1ULL);
#endif
                        const T m1a = sM1[PL::index_rt(bh, bl)];
                        const unsigned c = (cls >> (2 * bl)) & 3u;
                        T dc[VEC], dx[VEC], cand[VEC];                  // D_c and D of another column of the row (degree >= 2: column
                        ld_split_bytes<T, VEC>(Dk + c * VB, dc);        // 0 or 1 exists) differ exactly where c0(y) == c
                        ld_split_bytes<T, VEC>(Dk + (c ? 0u : 1u) * VB, dx);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) cand[j] = m1a + (dc[j] > dx[j] ? dx[j] : INF);
#pragma unroll
                        for (int bit = 1; bit < VEC; bit <<= 1) {
                            const bool sw = (bl & bit) != 0;
#pragma unroll
                            for (int j = 0; j < VEC; ++j)
                                if (!(j & bit)) { const T lo = cand[j], hi = cand[j | bit]; cand[j] = sw ? hi : lo; cand[j | bit] = sw ? lo : hi; }
                        }
#pragma unroll
                        for (int r = 0; r < VEC; ++r) if (cand[r] < best[r]) { best[r] = cand[r]; tag[r] = TW_SECOND | a | (a << 8); }
                    }
                }
            }
#if TW_STATS
            if (lane == 0) atomicAdd(tw_stats + 8 * (P.iter & 15), 1ULL);
#endif
        }
        if constexpr (sizeof(T) == 4 && VEC == 8) {                    // the entry list of the search lay over c01
            st_bytes<VEC>(c01 + x0, c01w);
            __syncwarp();
        }
        // ---- the recorded path of each owned s, straight from the tag (tw_pair_of_tag: one out-of-line copy of the code for the
        //      lane's VEC outputs -- the kernel must stay inside the instruction cache), and the output trellis (NBLDPC.cpp:1084-1094):
        //      Tout[d][t] = min over s of (deltaW[s] - dU[d][eta_d(s)]) at t = s ^ eta_d(s); column d of the path of s lands on s
        //      itself when the path does not deviate there (value deltaW[s]), on the partner's index when it does
        unsigned path[VEC];   // sb | sa << 8 | ca << 16 | cb << 18 | (columns the path deviates in) << 20 | pair << 24
        T va[VEC], vb[VEC];                                             // dw - ua, dw - ub of pair paths
        if constexpr (BS == 2) {
            // ---- REFERENCE TIE ORDER (tie_mode NBL_TIES_REFERENCE; n_r <= 2).  deltaW is what the search left in best; which of
            // the configurations of EXACTLY that cost is recorded decides the messages once inputs tie (punctured symbols,
            // quantised LLRs).  The reference keeps the first one its depth-first enumeration meets (NBLDPC.cpp:1897, strict '<'),
            // i.e. the lexicographically smallest tuple (q_0, .., q_dc-1), 0 = "column does not deviate" first: the first
            // deviating column as late as possible, its eta as small as possible, then the same for the second deviation (none
            // beats any).  That order is a key (tw_key); a second pass over the entries the search handled recomputes their
            // candidates, and where one EQUALS deltaW[s] (bit-equal sums) tw_ref_pair looks at all column assignments of that
            // pair.  Every tied pair is met: its lighter element pays <= deltaW/2 <= U/2, so it was handled.
            const bool second = P.p0 >= 2;
            unsigned kbest[VEC];
            {
                T m0s[VEC], m1s[VEC];
                ld_vec<T, VEC>(reinterpret_cast<const T *>(blk + lane * (int)S::blk_stride), m0s);
                ld_split<T, VEC>(sM1, lane, m1s);
#pragma unroll
                for (int r = 0; r < VEC; ++r) {
                    const int s = x0 + r, c0 = byte_of(c01w, r) & 15, c1 = byte_of(c01w, r) >> 4;
                    va[r] = T(0); vb[r] = T(0); path[r] = 0u; kbest[r] = s == 0 ? 0u : 0xffffffffu;
                    if (s != 0 && m0s[r] == best[r]) { kbest[r] = tw_key(c0, s, 4, 0); path[r] = 0x100000u << c0; }
                    if (s != 0 && second && m1s[r] == best[r] && tw_key(c1, s, 4, 0) < kbest[r]) { kbest[r] = tw_key(c1, s, 4, 0); path[r] = 0x100000u << c1; }
                }
            }
            // Almost every (entry, output slot) step has SOME lane with a hit, and a hit examined on the spot occupies the whole warp.
            // So the pass first only COLLECTS: up to four hit entries per output (one byte each) and their number; the assignments
            // are examined afterwards, all lanes together.  Continuous inputs have one or two hits per output; when any output of
            // the warp has more than four (integer LLRs, erasures) the pass runs once more and examines every hit where it is met.
            unsigned hl[VEC]; int hc[VEC];
#pragma unroll
            for (int r = 0; r < VEC; ++r) { hl[r] = 0u; hc[r] = 0; }
            bool direct = false;
#pragma unroll 1
            for (int attempt = 0; attempt < 2; ++attempt) {
#pragma unroll 1
                for (int bh = 0; bh < 32; ++bh) {
                    const unsigned m = __shfl_sync(0xffffffffu, done, bh);
                    if (!m) continue;                                   // warp-uniform
                    const unsigned char *rec = blk + bh * (int)S::blk_stride;
                    const unsigned cls = *reinterpret_cast<const unsigned *>(rec + S::blk_cls);
                    const unsigned char *Dk = slab + (lane ^ bh) * (PL::E16 * (int)sizeof(T));
#pragma unroll 1
                    for (unsigned mm = m & ((1u << VEC) - 1u); mm; mm &= mm - 1u) {   // (an entry of the second kind is one of the first kind too)
                        const int bl = __ffs((int)mm) - 1, a = VEC * bh + bl;
                        const T m0a = *reinterpret_cast<const T *>(rec + bl * (int)sizeof(T));
                        const T m1a = sM1[PL::index_rt(bh, bl)];
                        const unsigned c = (cls >> (2 * bl)) & 3u;
                        T dc[VEC], dx[VEC], cand[VEC];
                        ld_split_bytes<T, VEC>(Dk + c * VB, dc);
                        ld_split_bytes<T, VEC>(Dk + (c ? 0u : 1u) * VB, dx);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) cand[j] = second ? fmin(m0a + dc[j], m1a + (dc[j] > dx[j] ? dx[j] : INF)) : m0a + dc[j];   // cheapest assignment of the pair {a, b}
#pragma unroll
                        for (int bit = 1; bit < VEC; bit <<= 1) {
                            const bool sw = (bl & bit) != 0;
#pragma unroll
                            for (int j = 0; j < VEC; ++j)
                                if (!(j & bit)) { const T lo = cand[j], hi = cand[j | bit]; cand[j] = sw ? hi : lo; cand[j | bit] = sw ? lo : hi; }
                        }
                        if (!direct) {
#pragma unroll
                            for (int r = 0; r < VEC; ++r)
                                if (cand[r] == best[r] && x0 + r != 0) { hl[r] = (hl[r] << 8) | (unsigned)a; hc[r] += 1; }
                        } else {
#pragma unroll
                            for (int r = 0; r < VEC; ++r) {
                                if (cand[r] == best[r] && x0 + r != 0) {
                                    const TwRef<T> pr = tw_ref_pair<T, Q>(slab, a, x0 + r, (unsigned)PL::index(lane, r), best[r], second);
                                    if (pr.key < kbest[r]) { kbest[r] = pr.key; path[r] = pr.path; va[r] = pr.va; vb[r] = pr.vb; }
                                }
                            }
                        }
                    }
                }
                if (direct) break;
                int most = 0;
#pragma unroll
                for (int r = 0; r < VEC; ++r) most = hc[r] > most ? hc[r] : most;
                if (__any_sync(0xffffffffu, most > 4)) { direct = true; continue; }   // warp-uniform: examine every hit where it is met
#pragma unroll 1
                for (int i = 0; i < 4; ++i) {
                    if (!__any_sync(0xffffffffu, i < most)) break;
#pragma unroll
                    for (int r = 0; r < VEC; ++r) {
                        if (i < hc[r]) {
                            const TwRef<T> pr = tw_ref_pair<T, Q>(slab, (int)((hl[r] >> (8 * i)) & 255u), x0 + r, (unsigned)PL::index(lane, r), best[r], second);
                            if (pr.key < kbest[r]) { kbest[r] = pr.key; path[r] = pr.path; va[r] = pr.va; vb[r] = pr.vb; }
                        }
                    }
                }
                break;
            }
        } else {
#pragma unroll
        for (int r = 0; r < VEC; ++r) {
            const int s = x0 + r;
            va[r] = T(0); vb[r] = T(0);
            path[r] = s == 0 ? 0u : (0x100000u << (byte_of(c01w, r) & 15));      // single deviation of s in c0(s)
            if (tag[r] != TW_NOTAG) {
                const TwPair<T> pr = tw_pair_of_tag<T, Q>(slab, tag[r], s, (unsigned)PL::index(lane, r), best[r]);
                path[r] = pr.path; va[r] = pr.va; vb[r] = pr.vb;
            }
        }
        }
        __syncwarp();                                                  // everyone is done with D: its area becomes the output trellis
#pragma unroll
        for (int d = 0; d < TW_DC; ++d) {
            if (DC || d < dcn) {
                T tv[VEC];
#pragma unroll
                for (int r = 0; r < VEC; ++r) tv[r] = (path[r] & (0x100000u << d)) ? INF : best[r];
                st_split<T, VEC>(Tout + d * Q, lane, tv);
            }
        }
        __syncwarp();
        // ... then the landings of the two-deviation paths (values are >= 0: ordered like their bit patterns)
#pragma unroll
        for (int r = 0; r < VEC; ++r) {
            if (path[r] >> 24) {
                atomic_min_nonneg(Tout + ((path[r] >> 16) & 3u) * Q + (path[r] & 255u), va[r]);
                atomic_min_nonneg(Tout + ((path[r] >> 18) & 3u) * Q + ((path[r] >> 8) & 255u), vb[r]);
            }
        }
        __syncwarp();
    } else {
        // ---- BS-TEMS (Decoding_BS_TEMS, NBLDPC.cpp:1145-1396).  As the reference runs it the T-EMS stage only leaves the
        //      all-zero configuration (TEMS_Nr is never set in this mode); every other syndrome is reached by a subset of the
        //      BASIC SET or takes the fill-in value.
        // (a) Construct_BasicSet (:2023-2101): the n_m cheapest eta != 0 by m0(eta), lowest eta among equals; when
        //     n_m <= log2 q chosen greedily so that they stay linearly independent over GF(2) (a span flag per eta).
        //     Element k of the set lives on lane k: eta, its cheapest column, its cost.
        const int nm = P.p0, nc = P.p1;
        T rc[VEC];                                                  // m0 of my etas, +inf at eta = 0
        ld_vec<T, VEC>(reinterpret_cast<const T *>(blk + lane * (int)S::blk_stride), rc);
        unsigned avail = (1u << VEC) - 1u, span = 0u;
        int found = 0, my_bq = 0, my_col = 0; T my_bL = T(0);
        for (int k = 0; k < nm; ++k) {
            T lm = INF; int lj = 0;
#pragma unroll
            for (int j = VEC - 1; j >= 0; --j) if (((avail >> j) & 1u) && rc[j] <= lm) { lm = rc[j]; lj = j; }
            if (!(lm < INF)) lj = 0;
            T wm = lm;
#pragma unroll
            for (int o = 16; o; o >>= 1) wm = fmin(wm, __shfl_xor_sync(0xffffffffu, wm, o));
            if (!(wm < INF)) break;                                 // nothing left (n_m > number of candidates); warp-uniform
            const int src = __ffs(__ballot_sync(0xffffffffu, lm == wm)) - 1;
            const int w = VEC * src + __shfl_sync(0xffffffffu, lj, src);
            const int wcol = c01[w] & 15;
            found = k + 1;
            if (lane == k) { my_bq = w; my_col = wcol; my_bL = wm; }
            if (nm > g.p) { if (lane == src) avail &= ~(1u << lj); }
            else {                                                  // span of the chosen elements: in[x] = in[x] | in[x ^ w] | (x == w)
                const unsigned other = __shfl_sync(0xffffffffu, span, lane ^ (w / VEC));
                unsigned moved = 0u;
#pragma unroll
                for (int j = 0; j < VEC; ++j) moved |= ((other >> (j ^ (w % VEC))) & 1u) << j;
                span |= moved | (lane == src ? (1u << lj) : 0u);
                avail &= ~span;
            }
        }
        // (b) BasicSet_TEMS_ConstructConf (:2102-2147): one lane walks the subsets (<= n_c elements, no column twice) with the
        //     reference's include-first order and its add / subtract running sum; deltaW and the recorded subset per syndrome
        //     go through the (now unused) D area: dWs[eta], sets[eta] = index + 1 of the basic element per column.
        T *dWs = R1;
        unsigned *sets = reinterpret_cast<unsigned *>(R1 + Q);
        int *bsq = reinterpret_cast<int *>(sets + Q); int *bsc = bsq + 16; T *bsl = reinterpret_cast<T *>(bsc + 16);
        {
            T dinit[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) { dinit[j] = (x0 + j == 0) ? T(0) : INF; sets[x0 + j] = 0u; }
            st_vec<T, VEC>(dWs + x0, dinit);
            if (lane < found) { bsq[lane] = my_bq; bsc[lane] = my_col; bsl[lane] = my_bL; }
        }
        __syncwarp();
        if (lane == 0) {
            unsigned col_sel = 0u, cand = 0u; int diff = 0, sym = 0, lvl = 0; T sum = T(0);
            uint8_t st[17];                                         // 0 = not visited, 1 = inside "included", 2 = inside "excluded"
            st[0] = 0;
            while (lvl >= 0) {
                if (lvl == found) {                                 // a leaf: strict '<' keeps the first of equal costs
                    if (sum < dWs[sym]) { dWs[sym] = sum; sets[sym] = cand; }
                    --lvl;
                    continue;
                }
                const int col = bsc[lvl], bq = bsq[lvl];
                const T bL = bsl[lvl];
                if (st[lvl] == 0) {
                    if (!((col_sel >> col) & 1u)) {
                        if (diff + 1 <= nc) {
                            diff += 1; col_sel |= 1u << col;
                            sym ^= bq; sum = sum + bL; cand |= (unsigned)(lvl + 1) << (8 * col);
                            st[lvl] = 1;
                        } else st[lvl] = 2;
                    } else st[lvl] = 2;
                    ++lvl; st[lvl] = 0;
                } else if (st[lvl] == 1) {
                    sym ^= bq; sum = sum - bL; col_sel &= ~(1u << col); cand &= ~(255u << (8 * col));
                    diff -= 1;
                    st[lvl] = 2;
                    ++lvl; st[lvl] = 0;
                } else --lvl;
            }
        }
        __syncwarp();
        const int my_bq_sw = PL::swz(my_bq);
        T dw[VEC]; unsigned et[VEC];
        ld_vec<T, VEC>(dWs + x0, dw);
#pragma unroll
        for (int j = 0; j < VEC; ++j) et[j] = sets[x0 + j];
        __syncwarp();                                               // the D area becomes the output trellis
        // (c) output trellis (NBLDPC.cpp:1364-1374): Tout[d][eta ^ eta_d] = min(deltaW[eta] - dU[d][eta_d]) over the reached eta
        {
            T inf_v[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) inf_v[j] = INF;
#pragma unroll
            for (int d = 0; d < TW_DC; ++d) if (DC || d < dcn) st_split<T, VEC>(Tout + d * Q, lane, inf_v);
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < VEC; ++r) {
#pragma unroll
            for (int d = 0; d < TW_DC; ++d) {
                const int k1 = (int)((et[r] >> (8 * d)) & 255u);    // basic element deviating in column d, + 1
                const int ksrc = k1 ? k1 - 1 : 0;
                const int dev_sw = __shfl_sync(0xffffffffu, my_bq_sw, ksrc);
                const T du = __shfl_sync(0xffffffffu, my_bL, ksrc);
                if ((DC || d < dcn) && dw[r] < INF) {               // the layout is linear over XOR: swz(s ^ dev) = swz(s) ^ swz(dev)
                    const int t_sw = PL::index(lane, r) ^ (k1 ? dev_sw : 0);
                    atomic_min_any(Tout + d * Q + t_sw, k1 ? dw[r] - du : dw[r] - T(0));   // the running sum's residue can make this negative
                }
            }
        }
        __syncwarp();
    }
    // ---- fill-in (NBLDPC.cpp:1095-1102): indices no path landed on take the smallest dU of the other columns -------------
    {
        T m0[VEC], m1[VEC];
        ld_vec<T, VEC>(reinterpret_cast<const T *>(blk + lane * (int)S::blk_stride), m0);
        ld_split<T, VEC>(sM1, lane, m1);
#pragma unroll 1
        for (int d = 0; d < dcn; ++d) {
            T tv[VEC];
            ld_split<T, VEC>(Tout + d * Q, lane, tv);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (tv[j] == INF) tv[j] = ((byte_of(c01w, j) & 15) == d) ? m1[j] : m0[j];   // (t = 0 is never +inf)
            st_split<T, VEC>(Tout + d * Q, lane, tv);
        }
    }
    __syncwarp();
    // ---- normal domain, un-permute, correct (NBLDPC.cpp:1105-1128) ------------------------------------------------------
    T *coutf = P.c2v_out + (size_t)f * g.E * Q + VEC * lane;
    unsigned long long tnext = ld_bytes<VEC>(P.pfwd_sw + (size_t)e0 * Q + x0);
#pragma unroll 1
    for (int d = 0; d < dcn; ++d) {
        const unsigned long long tt = tnext;                        // the lane's bytes of h * x, as split-layout indices
        if (d + 1 < dcn) tnext = ld_bytes<VEC>(P.pfwd_sw + (size_t)(e0 + d + 1) * Q + x0);
        const int bs = PL::swz(syn ^ (int)((sbetaw >> (8 * d)) & 255u));
        const T Tb = Tout[d * Q + bs];
        const G gt(slab32 + d * VB, bs);
        T out[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) out[j] = Tb - gt.at(tt, j);
        if (P.factor != T(1)) {                                     // x / 1 == x exactly; skips the IEEE division routine
#pragma unroll
            for (int j = 0; j < VEC; ++j) out[j] = div_call(out[j], P.factor);
        }
        // soft threshold toward 0 by the offset: sign(v) * max(|v| - offset, 0), the same roundings as v -+ offset; "+ 0"
        // turns a -0 into the reference's +0.  (x = 0 needs no special case: h * 0 = 0 reads Tout[bs] itself.)
#pragma unroll
        for (int j = 0; j < VEC; ++j) out[j] = copysign(fmax(fabs(out[j]) - P.offset, T(0)), out[j]) + T(0);
        st_vec<T, VEC>(coutf + (size_t)(e0 + d) * Q, out);
    }
}

}  // namespace nbl
