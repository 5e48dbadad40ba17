// Bit-level ordered-statistics post-processing of the frames the iterative decoder gave up on.
//
// Replaces CNBLDPC::Decoding_OSD_bit and its helpers (reference OSD.h:5-199, 202-231, 326-420) and the matrix set-up of
// NBLDPC.cpp:380-457 / OSD.h:440-510.  What the reference computes, per failed frame:
//   1. rank the N p code bits by |LLR_Sym_sum[sym][2^k - 1]| (the posteriors accumulated over the iterations with the
//      forgetting factor OSD_factor; the channel LLRs when OSD_flag is set), most reliable first;
//   2. bring [CRC rows ; binary image of H] to systematic form with the pivots on the LEAST reliable positions
//      (H_GaussEliminate_bit: bottom row first, rows are added not swapped, a column with no pivot left moves to the
//      front of the order);
//   3. re-encode the hard decisions of the most reliable (independent) bits, then with every 1, 2 (3) of the first
//      MsgLen_bit of them flipped, and keep the candidate closest to the channel.
// Same results, different algorithm: one CTA per frame keeps the bit-packed matrix in shared memory (row operations
// are word-wide XORs spread over the threads); the systematic form gives one generator column per independent bit, so a
// candidate is base ^ column_i ^ column_j instead of a full re-encoding; all candidates are evaluated in parallel.
//
// Three quirks of the reference decide results and are reproduced on purpose:
//   * min_distance is an int (NBLDPC.h:127): the best distance is stored truncated, so a candidate only replaces the
//     current one if its distance is below floor(current) -- an order-dependent rule, resolved here by repeated
//     "first candidate after the current one below the threshold" searches (a handful of rounds);
//   * compute_min_distance_bit recomputes the bit length as int(CodeLen * log(GFq) / log(2)) (OSD.h:420), which is
//     N p - 1 for e.g. N = 96, q = 64: the last bit is then neither counted nor copied (dist_len, computed on the host
//     with the same expression);
//   * the binary image of the field element 1/alpha is all-zero because the Mat.Repr reader stops one record early
//     (GF.cpp:137) -- fix_matrepr = 0 keeps that.
// Ties between equal reliabilities: the reference uses std::sort, which leaves their order unspecified; here the lower
// bit index goes first.
#pragma once
#include "dev_common.cuh"

namespace nbl {

struct OsdDev {
    int order, flag, nbit, kbit, R, W, dist_len, p, N;
    double alpha;
    const unsigned *H0;      // [R][W] bit-packed: CRC rows, then the binary image of H
};

struct OsdSmem {
    // carve-up of the dynamic shared memory of k_osd
    unsigned *Hm, *D, *vec;  // [R][W], [kbit][W], [4][W]: base candidate, expected bit, valid mask, best candidate
    int *order;              // [nbit]
    double *key, *lbit;      // [nbit]
    __host__ __device__ static size_t bytes(int nbit, int kbit, int R, int W) {
        return sizeof(double) * 2 * (size_t)nbit + sizeof(unsigned) * ((size_t)R * W + (size_t)kbit * W + 4 * (size_t)W) + sizeof(int) * (size_t)nbit;
    }
    __device__ void carve(unsigned char *base, int nbit, int kbit, int R, int W) {
        key = reinterpret_cast<double *>(base); lbit = key + nbit;
        Hm = reinterpret_cast<unsigned *>(lbit + nbit); D = Hm + (size_t)R * W; vec = D + (size_t)kbit * W;
        order = reinterpret_cast<int *>(vec + 4 * (size_t)W);
    }
};

// LLR_Sym_sum restricted to the entries OSD reads: sum[f][j p + k] = alpha * sum + L_post[j][2^k - 1], once per iteration
// for every frame still decoding (NBLDPC.cpp:687, 820, 989, 1205).  The posterior is formed as the iteration kernels form
// it: channel value, then the variable's messages in file order.
template <typename T>
__global__ void k_osd_accumulate(DevGraph g, int B, int first, double alpha, const uint8_t *__restrict__ done,
                                 const T *__restrict__ lch, const T *__restrict__ c2v_in, double *sum) {
    const int nbit = g.N * g.p;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * nbit) return;
    const int f = (int)(i / nbit), n = (int)(i % nbit), j = n / g.p, x = 1 << (n % g.p);
    if (done[f]) return;
    T post = lch[((size_t)f * g.N + j) * g.q + x];
    if (!first)
        for (int s = g.var_ptr[j]; s < g.var_ptr[j + 1]; ++s) post = post + c2v_in[((size_t)f * g.E + g.var_edge[s]) * g.q + x];
    sum[i] = __dadd_rn(__dmul_rn(alpha, sum[i]), (double)post);
}

__device__ __forceinline__ int osd_bit(const unsigned *row, int pos) { return (row[pos >> 5] >> (pos & 31)) & 1; }

// candidate c -> the (up to three) flipped entries of the reliability order, in the reference's loop order:
// c = 0 no flip; then i = 0..kb-1; then i < j lexicographic; then i < j < k lexicographic
__device__ inline int osd_unrank(long long c, int kb, int &i, int &j, int &k) {
    i = j = k = -1;
    if (c == 0) return 0;
    c -= 1;
    if (c < kb) { i = (int)c; return 1; }
    c -= kb;
    const long long n2 = (long long)kb * (kb - 1) / 2;
    auto pair = [](long long r, int lo, int kb2, int &a, int &b) {          // r-th pair (a < b) with a >= lo
        // pairs starting at a: kb2 - 1 - a of them; before a (from lo): sum_{t=lo}^{a-1} (kb2 - 1 - t)
        int l = lo, h = kb2 - 2;
        auto before = [&](int a2) { const long long m = a2 - lo; return m * (kb2 - 1 - lo) - m * (m - 1) / 2; };
        while (l < h) { const int mid = (l + h + 1) / 2; if (before(mid) <= r) l = mid; else h = mid - 1; }
        a = l; b = a + 1 + (int)(r - before(l));
    };
    if (c < n2) { pair(c, 0, kb, i, j); return 2; }
    c -= n2;
    // triples: before first index a: C(kb,3) - C(kb-a,3)
    auto c3 = [](long long n) { return n < 3 ? 0LL : n * (n - 1) * (n - 2) / 6; };
    int l = 0, h = kb - 3;
    while (l < h) { const int mid = (l + h + 1) / 2; if (c3(kb) - c3(kb - mid) <= c) l = mid; else h = mid - 1; }
    i = l;
    pair(c - (c3(kb) - c3(kb - l)), l + 1, kb, j, k);
    return 3;
}

// One CTA per frame.  sel: frames with conv[f] != 0 are left alone (null: every frame, the stand-alone method 6).
template <typename T>
__global__ void __launch_bounds__(256) k_osd(OsdDev o, int Q, const int *__restrict__ conv, const T *__restrict__ lch,
                                             const double *__restrict__ sym_sum, int *out_sym, int *status) {
    extern __shared__ __align__(16) unsigned char osd_raw[];
    const int f = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    if (conv && conv[f]) return;
    OsdSmem S;
    S.carve(osd_raw, o.nbit, o.kbit, o.R, o.W);
    const int nbit = o.nbit, W = o.W, R = o.R, dep = nbit - R, kb = o.kbit, p = o.p;
    unsigned *b0 = S.vec, *expect = S.vec + W, *valid = S.vec + 2 * W, *bestv = S.vec + 3 * W;
    __shared__ int s_found, s_rot;
    __shared__ long long s_first;
    __shared__ int s_thr;

    // ---- step 1: reference decisions, bit LLRs of the channel, reliabilities ---------------------------------------------
    for (int w = tid; w < 4 * W; w += NT) S.vec[w] = 0;
    __syncthreads();
    for (int j = tid; j < o.N; j += NT) {
        const T *L = lch + ((size_t)f * o.N + j) * Q;
        int rec;
        if (o.flag) {                                               // DecideLLRVector of the channel vector (NBLDPC.cpp:1542-1562)
            T mx = T(0); rec = 0;
            for (int x = 1; x < Q; ++x) if (L[x] > mx) { mx = L[x]; rec = x; }
            out_sym[(size_t)f * o.N + j] = rec;
        } else rec = out_sym[(size_t)f * o.N + j];
        for (int k = 0; k < p; ++k) {
            const int n = j * p + k;
            const double l = (double)L[1 << k];
            S.lbit[n] = l;
            S.key[n] = fabs(o.flag ? l : sym_sum[(size_t)f * nbit + n]);
            if ((rec >> k) & 1) atomicOr(&bestv[n >> 5], 1u << (n & 31));           // infor_record_bit == near_optimal_bit
            if (n < o.dist_len) {
                if (l > 0) atomicOr(&expect[n >> 5], 1u << (n & 31));
                if (l != 0) atomicOr(&valid[n >> 5], 1u << (n & 31));
            }
        }
    }
    __syncthreads();
    // ---- step 2: OSD_permute -- descending reliability (rank sort) -------------------------------------------------------
    for (int n = tid; n < nbit; n += NT) {
        const double kn = S.key[n];
        int rank = 0;
        for (int m = 0; m < nbit; ++m) { const double km = S.key[m]; rank += (km > kn) || (km == kn && m < n); }
        S.order[rank] = n;
    }
    for (int w = tid; w < R * W; w += NT) S.Hm[w] = o.H0[w];
    __syncthreads();
    // ---- step 3: H_GaussEliminate_bit (OSD.h:326-384) ---------------------------------------------------------------------
    if (tid == 0) s_rot = 0;
    for (int row = R - 1; row >= 0;) {
        const int nt = row + dep, col = S.order[nt];
        if (tid == 0) s_found = -1;
        __syncthreads();
        const bool has = osd_bit(S.Hm + (size_t)row * W, col);
        if (!has) {
            for (int up = tid; up < row; up += NT) if (osd_bit(S.Hm + (size_t)up * W, col)) atomicMax(&s_found, up);
            __syncthreads();
            const int found = s_found;
            if (found >= 0) {
                for (int w = tid; w < W; w += NT) S.Hm[(size_t)row * W + w] ^= S.Hm[(size_t)found * W + w];
                __syncthreads();
            } else {                                                // no pivot in this column: it moves to the front, same row again
                int keep[8];                                        // nbit <= 8 * NT (host check)
                for (int t = 0, i = tid; i < nt; i += NT, ++t) keep[t] = S.order[i];
                __syncthreads();
                for (int t = 0, i = tid; i < nt; i += NT, ++t) S.order[i + 1] = keep[t];
                if (tid == 0) { S.order[0] = col; s_rot += 1; }
                __syncthreads();
                if (s_rot > nbit) break;                            // rank-deficient beyond repair (the reference would not terminate)
                continue;
            }
        }
        for (int up = tid; up < row; up += NT)
            if (osd_bit(S.Hm + (size_t)up * W, col))
                for (int w = 0; w < W; ++w) S.Hm[(size_t)up * W + w] ^= S.Hm[(size_t)row * W + w];
        __syncthreads();
        --row;
    }
    __syncthreads();
    if (s_rot > nbit) { if (tid == 0 && status) atomicOr(status, 1); return; }
    for (int i = 0; i < R; ++i) {
        const int pos = S.order[dep + i];
        for (int j = i + 1 + tid; j < R; j += NT)
            if (osd_bit(S.Hm + (size_t)j * W, pos))
                for (int w = 0; w < W; ++w) S.Hm[(size_t)j * W + w] ^= S.Hm[(size_t)i * W + w];
        __syncthreads();
    }
    // ---- step 4: generator columns.  Flipping the independent bit order[jj] changes itself and the parity bit of every
    //      row with a one in that column (OSD_Encode_bit, OSD.h:386-401); entries jj >= dep of the flip loops are
    //      dependent positions that the encoding overwrites: flipping them changes nothing (zero column). -------------------
    for (int w = tid; w < kb * W; w += NT) S.D[w] = 0;
    __syncthreads();
    for (int jj = tid; jj < dep; jj += NT) {
        const int pos = S.order[jj];
        unsigned *d = S.D + (size_t)jj * W;
        d[pos >> 5] |= 1u << (pos & 31);
        for (int i = 0; i < R; ++i)
            if (osd_bit(S.Hm + (size_t)i * W, pos)) { const int pp = S.order[dep + i]; d[pp >> 5] |= 1u << (pp & 31); }
    }
    __syncthreads();
    for (int w = tid; w < W; w += NT) {                             // order-0 candidate = sum of the columns of the set bits
        unsigned acc = 0;
        for (int jj = 0; jj < dep; ++jj) { const int pos = S.order[jj]; if ((bestv[pos >> 5] >> (pos & 31)) & 1) acc ^= S.D[(size_t)jj * W + w]; }
        b0[w] = acc;
    }
    __syncthreads();
    // ---- step 5: candidates in the reference's order; accept = distance < (int) current best ---------------------------------
    long long ncand = 1;
    if (o.order >= 1) ncand += kb;
    if (o.order >= 2) ncand += (long long)kb * (kb - 1) / 2;
    if (o.order >= 3) ncand += (long long)kb * (kb - 1) * (kb - 2) / 6;
    auto distance = [&](long long c) {
        int i, j, k;
        osd_unrank(c, kb, i, j, k);
        double d = 0;
        for (int w = 0; w < W; ++w) {
            unsigned x = b0[w];
            if (i >= 0) x ^= S.D[(size_t)i * W + w];
            if (j >= 0) x ^= S.D[(size_t)j * W + w];
            if (k >= 0) x ^= S.D[(size_t)k * W + w];
            unsigned m = (x ^ expect[w]) & valid[w];
            while (m) { const int b = __ffs(m) - 1; m &= m - 1; d = d + fabs(S.lbit[32 * w + b]); }   // ascending bit index
        }
        return d;
    };
    long long cur = -1;
    if (tid == 0) { s_thr = 1000000; s_first = -1; }
    __syncthreads();
    for (int round = 0; round < 4096; ++round) {
        if (tid == 0) s_first = ncand;
        __syncthreads();
        const int thr = s_thr;
        // thread t scans its candidates cur + 1 + t, cur + 1 + t + NT, ... and stops at its first hit
        for (long long c = cur + 1 + tid; c < ncand; c += NT) {
            if (c > *(volatile long long *)&s_first) break;
            if (distance(c) < (double)thr) { atomicMin((unsigned long long *)&s_first, (unsigned long long)c); break; }
        }
        __syncthreads();
        const long long first = s_first;
        if (first >= ncand) break;
        cur = first;
        __syncthreads();
        if (tid == 0) s_thr = (int)distance(cur);
        __syncthreads();
    }
    // ---- output: the accepted candidate, bits dist_len.. keep the reference decision (see header) -----------------------------
    if (cur >= 0) {
        int i, j, k;
        osd_unrank(cur, kb, i, j, k);
        for (int w = tid; w < W; w += NT) {
            unsigned x = b0[w];
            if (i >= 0) x ^= S.D[(size_t)i * W + w];
            if (j >= 0) x ^= S.D[(size_t)j * W + w];
            if (k >= 0) x ^= S.D[(size_t)k * W + w];
            unsigned keep = 0;                                      // bits >= dist_len
            for (int b = 0; b < 32; ++b) if (32 * w + b >= o.dist_len) keep |= 1u << b;
            bestv[w] = (x & ~keep) | (bestv[w] & keep);
        }
    }
    __syncthreads();
    for (int j = tid; j < o.N; j += NT) {
        int v = 0;
        for (int k = 0; k < p; ++k) { const int n = j * p + k; v |= ((bestv[n >> 5] >> (n & 31)) & 1) << k; }
        out_sym[(size_t)f * o.N + j] = v;
    }
}

}  // namespace nbl
