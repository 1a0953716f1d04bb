// ffma2_probe.cu -- issue cost, pipe throughput and dependent-issue latency of the packed fp32 FMA of sm_100
// (FFMA2, PTX fma.rn.f32x2 / __ffma2_rn) against the scalar FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int PACKED, int CHAINS>
__global__ void probe(float* out, int iters, float a, float b, long long* cyc) {
    float2 x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = make_float2(threadIdx.x + c, threadIdx.x - c);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (PACKED) x[c] = __ffma2_rn(x[c], a2, b2);
            else { x[c].x = fmaf(x[c].x, a, b); x[c].y = fmaf(x[c].y, a, b); }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c].x + x[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int PACKED, int CHAINS>
void run(const char* name, int blocks, int threads) {
    float* out; long long* cyc; long long h = 0;
    cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    probe<PACKED, CHAINS><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<PACKED, CHAINS><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fmas = 2.0 * CHAINS * iters * (double)blocks * threads;
    printf("%-34s blocks %4d x %4d: %8.1f cycles per iteration of warp 0 (%d chains x 2 FMAs), %7.2f TFLOP/s\n", name, blocks,
           threads, (double)h / iters, CHAINS, 2.0 * fmas / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    // latency: one warp, one chain (two scalar FMAs are independent, so the scalar loop shows the latency of one FFMA)
    run<0, 1>("scalar, 1 warp, 1 chain pair", 1, 32);
    run<1, 1>("packed, 1 warp, 1 chain", 1, 32);
    run<0, 8>("scalar, 1 warp, 8 chain pairs", 1, 32);
    run<1, 8>("packed, 1 warp, 8 chains", 1, 32);
    // one warp per scheduler: issue-bound
    run<0, 8>("scalar, 4 warps/SM, 8 chain pairs", sms, 128);
    run<1, 8>("packed, 4 warps/SM, 8 chains", sms, 128);
    // full machine
    run<0, 8>("scalar, 32 warps/SM, 8 chain pairs", sms, 1024);
    run<1, 8>("packed, 32 warps/SM, 8 chains", sms, 1024);
    return 0;
}
