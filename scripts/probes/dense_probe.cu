// Dense phase of the CSR GCN kernel in isolation: h = relu(agg.W1 + b1) (Fh wide), z = h.W2 on station PAIRS (FFMA2),
// weights either broadcast from shared memory (LDS.128) or from constant memory (uniform registers, LDCU).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dense_probe dense_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int FH = 128;
__constant__ float4 c_w1[FH * 4];
__constant__ float4 c_w2[FH * 4];
__constant__ float c_b1[FH];

template <int PAIRS, bool CONST, int UNR>
__global__ void __launch_bounds__(512, 1) probe(const float* __restrict__ gw, float* out, int iters) {
    __shared__ __align__(16) float w1t[FH * 16], w2p[FH * 16], b1s[FH];
    for (int e = threadIdx.x; e < FH * 16; e += blockDim.x) { w1t[e] = gw[e]; w2p[e] = gw[FH * 16 + e]; }
    for (int e = threadIdx.x; e < FH; e += blockDim.x) b1s[e] = gw[2 * FH * 16 + e];
    __syncthreads();
    float2 ag[13][PAIRS], z[13][PAIRS];
#pragma unroll
    for (int f = 0; f < 13; ++f)
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) {
            ag[f][q] = make_float2(threadIdx.x * 0.001f + f, threadIdx.x * 0.002f + q);
            z[f][q] = make_float2(0.f, 0.f);
        }
    for (int it = 0; it < iters; ++it) {
#pragma unroll UNR
        for (int fh = 0; fh < FH; ++fh) {
            float2 h[PAIRS];
            const float bb = CONST ? c_b1[fh] : b1s[fh];
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) h[q] = make_float2(bb, bb);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 w = CONST ? c_w1[fh * 4 + v] : *reinterpret_cast<const float4*>(w1t + fh * 16 + 4 * v);
                const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * v + j < 13)
#pragma unroll
                        for (int q = 0; q < PAIRS; ++q) h[q] = __ffma2_rn(make_float2(wv[j], wv[j]), ag[4 * v + j][q], h[q]);
            }
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) { h[q].x = fmaxf(h[q].x, 0.f); h[q].y = fmaxf(h[q].y, 0.f); }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 w = CONST ? c_w2[fh * 4 + v] : *reinterpret_cast<const float4*>(w2p + fh * 16 + 4 * v);
                const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * v + j < 13)
#pragma unroll
                        for (int q = 0; q < PAIRS; ++q) z[4 * v + j][q] = __ffma2_rn(make_float2(wv[j], wv[j]), h[q], z[4 * v + j][q]);
            }
        }
#pragma unroll
        for (int f = 0; f < 13; ++f)
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) { ag[f][q].x += z[f][q].y * 1e-9f; ag[f][q].y += z[f][q].x * 1e-9f; }
    }
    float s = 0;
#pragma unroll
    for (int f = 0; f < 13; ++f)
#pragma unroll
        for (int q = 0; q < PAIRS; ++q) s += z[f][q].x + z[f][q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int PAIRS, bool CONST, int UNR>
void run(const char* name, const float* gw, float* out, int threads) {
    const int iters = 200, blocks = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<PAIRS, CONST, UNR><<<blocks, threads>>>(gw, out, iters);
    cudaEventRecord(e0);
    probe<PAIRS, CONST, UNR><<<blocks, threads>>>(gw, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)blocks * threads * iters * FH * 26.0 * PAIRS * 4.0;
    printf("%-44s threads %4d  %.3f ms  %.1f TFLOP/s  (%s)\n", name, threads, ms, flop / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* gw; float* out;
    cudaMalloc(&gw, (2 * FH * 16 + FH) * 4); cudaMalloc(&out, 148 * 512 * 4);
    float* hw = new float[2 * FH * 16 + FH];
    for (int i = 0; i < 2 * FH * 16 + FH; ++i) hw[i] = 0.01f * ((i * 37) % 19 - 9);
    cudaMemcpy(gw, hw, (2 * FH * 16 + FH) * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_w1, hw, FH * 16 * 4); cudaMemcpyToSymbol(c_w2, hw + FH * 16, FH * 16 * 4);
    cudaMemcpyToSymbol(c_b1, hw + 2 * FH * 16, FH * 4);
    for (int threads : {512, 384, 256}) {
        run<1, false, 2>("smem weights, 1 pair, unroll 2", gw, out, threads);
        run<1, true, 2>("const weights, 1 pair, unroll 2", gw, out, threads);
        run<1, true, 4>("const weights, 1 pair, unroll 4", gw, out, threads);
        run<2, false, 2>("smem weights, 2 pairs, unroll 2", gw, out, threads);
        run<2, true, 1>("const weights, 2 pairs, unroll 1", gw, out, threads);
        run<2, true, 2>("const weights, 2 pairs, unroll 2", gw, out, threads);
    }
    return 0;
}
