// Microbenchmark: FP32 scalar vs packed (f32x2) throughput on sm_100a. Developer tool, not product.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out, float a, float b, unsigned long long* cyc)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // scalar FFMA, 16 independent
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {  // packed fma.rn.f32x2, 8 independent pairs
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, aa, bb;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(aa) : "f"(a));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(bb) : "f"(b));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 2) {  // scalar FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + a;
        } else if (MODE == 3) {  // packed add
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long v, aa;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[i]), "f"(x[i + 1]));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(aa) : "f"(a));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(aa));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(v));
            }
        } else if (MODE == 4) {  // FFMA with immediate-ish constants
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], 1.0001f, 0.5f);
        } else if (MODE == 5) {  // FMUL scalar
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] * a;
        }
    }
    unsigned long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_sm, int flops_per_inst)
{
    float* out; unsigned long long* cyc;
    int blocks = 148, threads = warps_per_sm * 32;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double lane_ops = (double)ITERS * 16 * threads;  // per SM: scalar-equivalent element ops
    printf("%-28s warps/SM=%2d  cycles=%9.0f  elem-ops/clk/SM=%7.1f  (%.3f ms, %.1f Gelem-op/s chip)\n", name, warps_per_sm, c,
           lane_ops / c, ms, lane_ops * 148 / ms / 1e6);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA scalar (reg)", w, 2);
        run<4>("FFMA scalar (imm)", w, 2);
        run<1>("fma.rn.f32x2", w, 4);
        run<2>("FADD scalar", w, 1);
        run<3>("add.rn.f32x2", w, 2);
        run<5>("FMUL scalar", w, 1);
    }
    return 0;
}
