// Dependent-issue latency of FFMA vs FFMA2 (one warp), and throughput vs number of independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS, bool PACKED>
__global__ void k(float* sink, long long* cyc, int iters, float a, float b) {
    float2 v[CHAINS];
    for (int i = 0; i < CHAINS; ++i) v[i] = make_float2(threadIdx.x + i, i);
    const float2 bb = make_float2(b, b * 0.5f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (PACKED) v[i] = __ffma2_rn(make_float2(a, a), v[i], bb);
                else v[i].x = fmaf(a, v[i].x, b);
            }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < CHAINS; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CHAINS, bool PACKED>
void run(float* sink, long long* cyc, const char* name) {
    const int iters = 4096;
    k<CHAINS, PACKED><<<1, 32>>>(sink, cyc, iters, 0.999f, 0.001f);
    k<CHAINS, PACKED><<<1, 32>>>(sink, cyc, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s chains=%2d: %.2f cycles per instruction (one warp)\n", name, CHAINS, (double)c / (iters * 8.0 * CHAINS));
}
int main() {
    float* sink; long long* cyc; cudaMalloc(&sink, 4); cudaMalloc(&cyc, 8);
    run<1, false>(sink, cyc, "FFMA "); run<2, false>(sink, cyc, "FFMA "); run<4, false>(sink, cyc, "FFMA "); run<8, false>(sink, cyc, "FFMA ");
    run<1, true>(sink, cyc, "FFMA2"); run<2, true>(sink, cyc, "FFMA2"); run<3, true>(sink, cyc, "FFMA2"); run<4, true>(sink, cyc, "FFMA2");
    run<5, true>(sink, cyc, "FFMA2"); run<6, true>(sink, cyc, "FFMA2"); run<8, true>(sink, cyc, "FFMA2"); run<12, true>(sink, cyc, "FFMA2"); run<20, true>(sink, cyc, "FFMA2");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
