// Micro-benchmark (development aid): issue rates of the instruction mixes a min-plus leaf can be built from, sm_100a.
// Every variant keeps 8 independent accumulators per thread and feeds operands that change every step, so that
// nothing can be shared between leaves (checked in SASS: one add + one min per leaf, or two adds + one 3-input min
// per two leaves).  Output: warp-instructions per clock per SM and leaves per clock per SM at the measured SM clock.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define ITER 2048
template <int MODE> __global__ void __launch_bounds__(256) k(float *out, const float *in, long long *clk) {
    float b[8], p[8], q[8];
    int ib[8], ip[8];
    for (int i = 0; i < 8; ++i) {
        b[i] = 1e30f; p[i] = in[(threadIdx.x + i) & 63]; q[i] = in[(threadIdx.x + 2 * i + 1) & 63];
        ib[i] = 0x7fffffff; ip[i] = __float_as_int(p[i]) >> 8;
    }
    float c = in[threadIdx.x & 7], d = in[threadIdx.x & 3];
    int ic = __float_as_int(c) >> 8;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (MODE == 0) b[r] = b[r] + p[r ^ u];                                         // FADD
                if (MODE == 1) b[r] = (u & 1) ? fminf(b[r], p[r ^ u]) : fmaxf(b[r], q[r ^ u]);   // FMNMX (min/max alternate: no 3-input fusion)
                if (MODE == 2) b[r] = fminf(fminf(b[r], p[r ^ u]), q[r ^ u]);                    // FMNMX3
                if (MODE == 3) b[r] = fminf(b[r], p[r ^ u] + c);                                // FADD + FMNMX
                if (MODE == 4) b[r] = fminf(fminf(b[r], p[r ^ u] + c), q[r ^ u] + d);            // 2 FADD + FMNMX3 (2 leaves)
                if (MODE == 5) ib[r] = __viaddmin_s32(ip[r ^ u], ic, ib[r]);                    // VIADDMNMX
                if (MODE == 6) ib[r] = min(ib[r], ip[r ^ u] + ic);                              // IADD + VIMNMX
                if (MODE == 7) ib[r] = min(min(ib[r], ip[r ^ u] + ic), ip[r ^ u ^ 1] - ic);      // 2 IADD + VIMNMX3 ?
                if (MODE == 8) b[r] = (u & 1) ? fminf(b[r], p[r ^ u] + c) : fmaxf(b[r], q[r ^ u] + d);   // FADD + FMNMX, unfusable
            }
            c += 1.0f; d += 0.5f; ic += 3;
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += b[i] + (float)ib[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int MODE> void run(const char *name, double leaves_per_inner, double instr_per_inner) {
    const int blocks = 148 * 4;
    float *out, *in; long long *clk;
    cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&in, 256); cudaMemset(in, 0, 256); cudaMalloc(&clk, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, in, clk); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, in, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, clk, 8, cudaMemcpyDeviceToHost);
    // 4 CTAs of 8 warps per SM all resident: per-SM inner steps = 4 CTAs * 256 thr * ITER * 64; cycles from clock64
    const double inner_per_sm = 4.0 * 256 * ITER * 64;
    // SMSP cycles per warp per loop iteration (8 warps per SMSP, 64 inner steps per iteration): compare with the SASS body
    printf("%-22s %.3f ms  %.1f SMSP-cycles per warp-iteration (64 inner steps)  %.1f leaves/clk/SM (clock64)  %.1f leaves/clk/SM (@1.965GHz wall)\n",
           name, ms, (double)cyc / ITER / 8, inner_per_sm * leaves_per_inner / cyc,
           148.0 * inner_per_sm * leaves_per_inner / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    run<0>("FADD", 1, 1); run<1>("FMNMX", 1, 1); run<2>("FMNMX3", 2, 1); run<3>("FADD+FMNMX", 1, 2); run<4>("2FADD+FMNMX3", 2, 3);
    run<5>("VIADDMNMX.s32", 1, 1); run<6>("IADD+VIMNMX", 1, 2); run<7>("2IADD+VIMNMX3", 2, 3); run<8>("FADD+FMNMX nofuse", 1, 2);
    return 0;
}
