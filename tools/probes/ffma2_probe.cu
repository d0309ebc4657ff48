// Microbenchmark (GPU box): issue / pipe throughput of scalar FFMA vs packed FFMA2 / FMUL2 / FADD2 on sm_100a, as used
// by scgaussian_b200/csrc/render.cu.  Each thread runs CHAINS independent dependency chains of ITERS instructions;
// reports warp-instructions per cycle per SM and lane-FMAs per cycle per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>

#define ITERS 4096
template <int MODE, int CHAINS>
__global__ void __launch_bounds__(256) probe(float* out, float a, float b) {
    float2 acc[CHAINS];
    for (int c = 0; c < CHAINS; c++) acc[c] = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f - c);
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) {            // scalar FFMA x2 (same work as one FFMA2)
                acc[c].x = fmaf(acc[c].x, A.x, B.x);
                acc[c].y = fmaf(acc[c].y, A.y, B.y);
            } else if (MODE == 1) {     // FFMA2
                unsigned long long ra, rb, rc;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(acc[c].x), "f"(acc[c].y));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(A.x), "f"(A.y));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(B.x), "f"(B.y));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(ra) : "l"(rb), "l"(rc));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[c].x), "=f"(acc[c].y) : "l"(ra));
            } else if (MODE == 2) {     // FMUL2
                unsigned long long ra, rb;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(acc[c].x), "f"(acc[c].y));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(A.x), "f"(A.y));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(ra) : "l"(rb));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[c].x), "=f"(acc[c].y) : "l"(ra));
            } else {                    // scalar FMUL x2
                acc[c].x *= A.x;
                acc[c].y *= A.y;
            }
        }
    }
    float s = 0.f;
    for (int c = 0; c < CHAINS; c++) s += acc[c].x + acc[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CHAINS>
void run(const char* name, int warps_per_sm) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
    float* out;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, CHAINS><<<blocks, threads>>>(out, 1.0001f, 1e-6f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) probe<MODE, CHAINS><<<blocks, threads>>>(out, 1.0001f, 1e-6f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double per_thread_ops = (double)ITERS * CHAINS;            // "pair" operations (2 lane-FMAs each)
    const double warp_pairs = per_thread_ops * blocks * threads / 32.0;
    const double cycles = ms * 1e-3 * khz * 1e3;                     // at the nominal max clock
    printf("%-22s warps/SM %2d chains %d: %.3f ms  pair-ops/clk/SM %.2f  (lane-FMA/clk/SM %.1f)\n", name, warps_per_sm, CHAINS, ms,
           warp_pairs / cycles / sms, warp_pairs * 64.0 / cycles / sms);
    cudaFree(out);
}

int main() {
    for (int w : {16, 32}) {
        run<0, 4>("2x scalar FFMA", w); run<1, 4>("FFMA2", w); run<3, 4>("2x scalar FMUL", w); run<2, 4>("FMUL2", w);
        run<0, 8>("2x scalar FFMA", w); run<1, 8>("FFMA2", w);
    }
    return 0;
}
