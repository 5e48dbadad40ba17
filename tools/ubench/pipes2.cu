// Micro-benchmark: sustained issue rate of the instruction mixes the T-EMS pair search is built from, on sm_100a.
// Every mix is written in PTX inside asm volatile, 8 independent dependency chains per thread, and the SASS of this
// file is checked to contain exactly the instructions named (see tools/ubench/README).  Rates are computed per SM from
// clock64() (the SM's own clock): instructions executed on an SM / (last CTA end - first CTA start on that SM), so they
// do not depend on the wall clock or on the SM frequency.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu
//   output: one line per mix, and a final JSON line {"leaf_per_clk_sm": ..} = the FADD2+FMNMX3 leaf rate
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define ITER 4096
enum { M_FADD, M_FADD2, M_FMNMX, M_FMNMX3, M_LEAF, M_LEAF_LOP, M_LEAF_TAG, M_LOP3, M_SETPSEL, M_IMADMOV, M_LEAF_LDS, M_LDS128, M_LDS32, M_FADD_FMNMX, M_COUNT };

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, const float *in, long long *rec, int lds_stride) {
    __shared__ float4 sm[1024];
    float b[8], p[8];
    unsigned long long pp[4];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(in[i & 63], 1.f, 2.f, 3.f);
    for (int i = 0; i < 8; ++i) { b[i] = 1e30f + in[i]; p[i] = in[(threadIdx.x + i) & 63]; }
    for (int i = 0; i < 4; ++i) pp[i] = ((unsigned long long)__float_as_uint(p[2 * i + 1]) << 32) | __float_as_uint(p[2 * i]);
    float c = in[threadIdx.x & 7];
    unsigned u = threadIdx.x, tag0 = 0, tag1 = 0;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * lds_stride;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == M_FADD) {
#pragma unroll
            for (int r = 0; r < 8; ++r) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[r]) : "f"(c));
        } else if (MODE == M_FADD2) {
#pragma unroll
            for (int r = 0; r < 8; ++r) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pp[r & 3]) : "l"(pp[(r + 1) & 3]));
        } else if (MODE == M_FMNMX) {
#pragma unroll
            for (int r = 0; r < 8; ++r) asm volatile("min.f32 %0, %0, %1;" : "+f"(b[r]) : "f"(p[r]));
        } else if (MODE == M_FMNMX3) {
#pragma unroll
            for (int r = 0; r < 8; ++r) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(b[r]) : "f"(p[r]), "f"(p[r ^ 1]));
        } else if (MODE == M_FADD_FMNMX) {      // 8 leaves: 8 FADD + 8 FMNMX
#pragma unroll
            for (int r = 0; r < 8; ++r) { float t; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(t) : "f"(p[r]), "f"(c)); asm volatile("min.f32 %0, %0, %1;" : "+f"(b[r]) : "f"(t)); }
        } else if (MODE == M_LEAF || MODE == M_LEAF_LOP || MODE == M_LEAF_TAG || MODE == M_LEAF_LDS) {
            // 16 leaves: 8 FADD2 + 8 FMNMX3 (two entries against 8 outputs), the search loop's arithmetic
            unsigned long long q0[4], q1[4];
            if (MODE == M_LEAF_LDS) {                  // + the block's share of D loads: 8 LDS.128 per 64 leaf instructions -> 2 per 16
                float4 v0, v1;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(sbase + ((it & 7) << 9)));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(sbase + ((it & 7) << 9) + 4096));
                pp[0] = ((unsigned long long)__float_as_uint(v0.y) << 32) | __float_as_uint(v0.x);
                pp[1] = ((unsigned long long)__float_as_uint(v0.w) << 32) | __float_as_uint(v0.z);
                pp[2] = ((unsigned long long)__float_as_uint(v1.y) << 32) | __float_as_uint(v1.x);
                pp[3] = ((unsigned long long)__float_as_uint(v1.w) << 32) | __float_as_uint(v1.z);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm volatile("{.reg .b64 cc; mov.b64 cc, {%2, %2}; add.rn.f32x2 %0, %1, cc;}" : "=l"(q0[i]) : "l"(pp[i]), "f"(c));
                asm volatile("{.reg .b64 cc; mov.b64 cc, {%2, %2}; add.rn.f32x2 %0, %1, cc;}" : "=l"(q1[i]) : "l"(pp[i ^ 1]), "f"(p[0]));
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float a0 = __uint_as_float((unsigned)(r & 1 ? q0[r >> 1] >> 32 : q0[r >> 1]));
                const float a1 = __uint_as_float((unsigned)(r & 1 ? q1[r >> 1] >> 32 : q1[r >> 1]));
                if (MODE == M_LEAF_TAG && (r & 3) == 0) {      // + 2 of the 8 tag updates a block does per 4 entry pairs
                    float nb; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(nb) : "f"(b[r]), "f"(a0), "f"(a1));
                    asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; selp.b32 %0, %3, %0, q;}" : "+r"(tag0) : "f"(nb), "f"(b[r]), "r"(it));
                    b[r] = nb;
                } else asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(b[r]) : "f"(a0), "f"(a1));
            }
            if (MODE == M_LEAF_LOP) {                  // + the dispatch tests of two entries (4 LOP3 with predicate output ~ 4 ALU ops)
#pragma unroll
                for (int r = 0; r < 4; ++r) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(tag0 + r), "r"(tag1));
            }
        } else if (MODE == M_LOP3) {
            unsigned w[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) w[r] = __float_as_uint(b[r]);
#pragma unroll
            for (int r = 0; r < 8; ++r) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[r]) : "r"(u), "r"(tag1));
#pragma unroll
            for (int r = 0; r < 8; ++r) b[r] = __uint_as_float(w[r]);
        } else if (MODE == M_SETPSEL) {            // 4 x (FSETP + SEL)
#pragma unroll
            for (int r = 0; r < 4; ++r) asm volatile("{.reg .pred q; setp.lt.f32 q, %1, %2; selp.b32 %0, %3, %0, q;}" : "+r"(tag0) : "f"(b[r]), "f"(p[r]), "r"(it + r));
#pragma unroll
            for (int r = 0; r < 4; ++r) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[r]) : "f"(c));   // keeps the compares live
        } else if (MODE == M_IMADMOV) {
            unsigned w[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) w[r] = __float_as_uint(b[r]);
#pragma unroll
            for (int r = 0; r < 8; ++r) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[r]) : "r"(u), "r"(tag1));
#pragma unroll
            for (int r = 0; r < 8; ++r) b[r] = __uint_as_float(w[r]);
        } else if (MODE == M_LDS128 || MODE == M_LDS32) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (MODE == M_LDS128) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sbase + (((it + r) & 15) << 9)));
                    b[r] += v.x + v.w;
                } else {
                    float v;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + (((it + r) & 15) << 9)));
                    b[r] += v;
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += b[i];
    for (int i = 0; i < 4; ++i) s += (float)(pp[i] >> 40);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u + tag0 + tag1);
    if (threadIdx.x == 0) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rec[3 * blockIdx.x] = smid; rec[3 * blockIdx.x + 1] = t0; rec[3 * blockIdx.x + 2] = t1;
    }
}

static double g_leaf = 0, g_fadd = 0;
template <int MODE> void run(const char *name, double counted_per_iter, const char *what, int ctas_per_sm = 4, int lds_stride = 16) {
    const int blocks = 148 * ctas_per_sm;
    float *out, *in; long long *rec;
    cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&rec, blocks * 24);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, in, rec, lds_stride); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, in, rec, lds_stride); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(3 * blocks); cudaMemcpy(h.data(), rec, blocks * 24, cudaMemcpyDeviceToHost);
    std::vector<long long> lo(256, -1), hi(256, 0), n(256, 0);
    for (int i = 0; i < blocks; ++i) { int s = (int)h[3 * i]; if (lo[s] < 0 || h[3 * i + 1] < lo[s]) lo[s] = h[3 * i + 1]; hi[s] = std::max(hi[s], h[3 * i + 2]); n[s]++; }
    std::vector<double> rate;
    for (int s = 0; s < 256; ++s) if (n[s]) rate.push_back((double)n[s] * 8 /*warps*/ * ITER * counted_per_iter / (double)(hi[s] - lo[s]));
    std::sort(rate.begin(), rate.end());
    const double med = rate[rate.size() / 2];
    const double wall = (double)blocks * 8 * ITER * counted_per_iter / 148.0 / (ms * 1e-3);     // per SM per second
    printf("%-28s %8.3f ms  %7.3f %s / clk / SM (clock64, median of %zu SMs; min %.3f max %.3f)   implied SM clock %.0f MHz\n", name, ms, med, what, rate.size(), rate.front(), rate.back(), wall / med / 1e6);
    if (MODE == M_LEAF) g_leaf = med * 32;
    if (MODE == M_FADD) g_fadd = med;
    cudaFree(out); cudaFree(in); cudaFree(rec);
}
int main() {
    run<M_FADD>("FADD", 8, "warp-instr");
    run<M_FADD2>("FADD2", 8, "warp-instr");
    run<M_FMNMX>("FMNMX", 8, "warp-instr");
    run<M_FMNMX3>("FMNMX3", 8, "warp-instr");
    run<M_LOP3>("LOP3", 8, "warp-instr");
    run<M_IMADMOV>("IMAD", 8, "warp-instr");
    run<M_SETPSEL>("FSETP+SEL (+FADD) x4", 12, "warp-instr");
    run<M_FADD_FMNMX>("FADD+FMNMX leaf", 8, "warp-leaves");
    run<M_LEAF>("FADD2+FMNMX3 leaf", 16, "warp-leaves");
    run<M_LEAF_LOP>("leaf + dispatch tests", 16, "warp-leaves");
    run<M_LEAF_TAG>("leaf + tags", 16, "warp-leaves");
    run<M_LEAF_LDS>("leaf + D loads", 16, "warp-leaves");
    run<M_LDS128>("LDS.128 conflict-free", 8, "warp-instr");
    run<M_LDS32>("LDS.32 conflict-free", 8, "warp-instr", 4, 4);
    printf("{\"leaf_per_clk_sm\": %.2f, \"fadd_warp_instr_per_clk_sm\": %.3f}\n", g_leaf, g_fadd);
    return 0;
}
