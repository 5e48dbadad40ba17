// Micro-benchmark (development aid): issue rate of the candidate min-plus leaf sequences on sm_100a.
//   A: FADD + FMNMX            (fp32 leaf as used by the T-EMS search)
//   B: FADD,FADD + 3-input min (fminf(fminf(a,b),c) -> FMNMX3 if the compiler fuses it)
//   C: __viaddmin_s32          (DPX fused add+min on int32 fixed point)
//   D: __viaddmin_s16x2        (DPX, two 16-bit lanes per register)
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE> __global__ void k(float *out, const float *in, int n) {
    float b[8]; int ib[8];
    float p[8], c = in[threadIdx.x & 7];
    for (int i = 0; i < 8; ++i) { b[i] = 1e30f; p[i] = in[(threadIdx.x + i) & 63]; ib[i] = 0x7fffffff; }
    int ic = __float_as_int(c) >> 8;
    int ip[8]; for (int i = 0; i < 8; ++i) ip[i] = __float_as_int(p[i]) >> 8;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (MODE == 0) b[r] = fminf(b[r], p[r ^ u] + c);
                if (MODE == 1) b[r] = fminf(fminf(b[r], p[r ^ u] + c), p[r ^ (u + 4)] + c);
                if (MODE == 2) ib[r] = __viaddmin_s32(ip[r ^ u], ic, ib[r]);
                if (MODE == 3) ib[r] = (int)__viaddmin_s16x2((unsigned)ip[r ^ u], (unsigned)ic, (unsigned)ib[r]);
            }
            c += 1.0f; ic += 3;
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += b[i] + (float)ib[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double leaves_per_inner) {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&in, 256); cudaMemset(in, 0, 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, in, 0); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, in, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double leaves = 148.0 * 8 * 256 * ITER * 4 * 8 * leaves_per_inner;
    printf("%-28s %.3f ms  %.2f T leaf/s  (%.1f leaves/clk/SM @1.965GHz)\n", name, ms, leaves / ms / 1e9, leaves / (ms * 1e-3) / 148 / 1.965e9);
}
int main() { run<0>("FADD+FMNMX", 1); run<1>("2xFADD+min3", 2); run<2>("viaddmin_s32", 1); run<3>("viaddmin_s16x2", 2); return 0; }
