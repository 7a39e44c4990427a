// FP32 throughput probe (sm_100a): scalar FFMA vs packed FFMA2, chain form and register-tile form.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_chain1(float* sink, int iters, float a, float b) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) k_chain2(float* sink, int iters, float a, float b) {
    float2 v[8];
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.5f);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ffma2_rn(v[i], aa, bb);
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) sink[0] = s;
}
// 8x8 register tile, scalar FFMA, operands rotate every iteration (no memory traffic)
__global__ void __launch_bounds__(128) k_tile1(float* sink, int iters, float seed) {
    float acc[8][8], a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed - i; for (int j = 0; j < 8; ++j) acc[i][j] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            float t = a[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) a[i] = a[i + 1];
            a[7] = t;
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    if (s == 123.456f) sink[0] = s;
}
// 8x8 register tile with FFMA2: pairs along N, a broadcast into both halves
__global__ void __launch_bounds__(128) k_tile2(float* sink, int iters, float seed) {
    float2 acc[8][4], b[4]; float a[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0, 0); }
    for (int j = 0; j < 4; ++j) b[j] = make_float2(seed - j, seed + j);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(aa, b[j], acc[i][j]);
            }
            float t = a[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) a[i] = a[i + 1];
            a[7] = t;
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
    if (s == 123.456f) sink[0] = s;
}
// FFMA2 tile, pairs along K: both operands natural pairs, two partial sums per output
__global__ void __launch_bounds__(128) k_tile2k(float* sink, int iters, float seed) {
    float2 acc[8][8], a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed); b[i] = make_float2(seed - i, seed + i); for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0, 0); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __ffma2_rn(a[i], b[j], acc[i][j]);
            float2 t = a[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) a[i] = a[i + 1];
            a[7] = t;
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
    if (s == 123.456f) sink[0] = s;
}

template <typename F>
double timeit(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5 * 1e-3;
}

int main() {
    float* sink; cudaMalloc(&sink, 4);
    const int iters = 2048;
    int nsm = 148;
    for (int warps_per_sm : {64, 32, 16, 12, 8, 4}) {
        int blocks256 = nsm * warps_per_sm / 8, blocks128 = nsm * warps_per_sm / 4;
        double t;
        t = timeit([&] { k_chain1<<<blocks256, 256>>>(sink, iters, 0.999f, 0.001f); });
        double c1 = 2.0 * 8 * 16 * iters * 256.0 * blocks256 / t / 1e12;
        t = timeit([&] { k_chain2<<<blocks256, 256>>>(sink, iters, 0.999f, 0.001f); });
        double c2 = 2.0 * 2 * 8 * 16 * iters * 256.0 * blocks256 / t / 1e12;
        t = timeit([&] { k_tile1<<<blocks128, 128>>>(sink, iters, 0.5f); });
        double t1 = 2.0 * 64 * 4 * iters * 128.0 * blocks128 / t / 1e12;
        t = timeit([&] { k_tile2<<<blocks128, 128>>>(sink, iters, 0.5f); });
        double t2 = 2.0 * 64 * 4 * iters * 128.0 * blocks128 / t / 1e12;
        t = timeit([&] { k_tile2k<<<blocks128, 128>>>(sink, iters, 0.5f); });
        double t2k = 2.0 * 128 * 2 * iters * 128.0 * blocks128 / t / 1e12;
        printf("warps/SM %2d: chain FFMA %.1f | chain FFMA2 %.1f | 8x8 tile FFMA %.1f | 8x8 tile FFMA2(N pairs,a dup) %.1f | 8x8 tile FFMA2(K pairs) %.1f TFLOP/s\n",
               warps_per_sm, c1, c2, t1, t2, t2k);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
