// GCN layers on a large sparse graph (CSR adjacency, wide hidden layer) — the scaled-shape path
// (BASELINE.json configs[3]: 4096 stations, kNN(k=8), GCN hidden width 128).
//
// Reference arithmetic: GraphConvLayer.forward twice (src/step5_gcn_layer_model.py:13-23 via
// src/step6_gcn_gru_combined_model.py:17,20) with A_hat stored as CSR: (A_hat . X)[s] is a sum
// over the ~10 neighbours of s instead of a dense S-long dot product.
//   layer 1:  G1 = relu((A.X).W1 + b1)                       [S, F_hid]  (never materialised)
//   layer 2:  G2 = relu((A.G1).W2 + b2) = relu(A.(G1.W2) + b2)
// The second form is what runs: Z = G1.W2 is only F_out wide, so the F_hid-wide hidden layer
// lives in registers of the thread that produced it and only Z [rows, S, F_out] goes through a
// scratch buffer between the two kernels.  (A.G1).W2 and A.(G1.W2) are the same sum in a
// different association; in fp32 they differ by rounding only (covered by the 1e-5 parity bar).
//
// Mapping: a CTA = 128 consecutive rows (one K-major output tile of the projection GEMM) x a
// chunk of stations; lane = row, so the CSR entries and the weights are warp-uniform (broadcast)
// and the tiled output store is coalesced; each lane gathers its own row's F-wide station slabs.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kSpThreads = 128;
constexpr int kSpF = 16;  // max narrow feature width (F_in, F_out)

// kernel 1: Z[r][s][:] = relu((A.X[r])[s] . W1 + b1) . W2          (x: [R][S][Fi], Z: [R][S][Fo])
__global__ void __launch_bounds__(kSpThreads)
    gcn_sparse_l1_kernel(const float* __restrict__ X, const int* __restrict__ rowptr,
                         const int* __restrict__ colidx, const float* __restrict__ vals,
                         const float* __restrict__ W1, const float* __restrict__ b1,
                         const float* __restrict__ W2, float* __restrict__ Z, long long R, int S, int Fi,
                         int Fh, int Fo, int s_chunk) {
    extern __shared__ __align__(16) float smem[];
    float* w1t = smem;                    // [Fh][16]: w1t[fh][f] = W1[f][fh]
    float* w2p = w1t + (size_t)Fh * kSpF; // [Fh][16]: w2p[fh][fo] = W2[fh][fo]
    float* b1s = w2p + (size_t)Fh * kSpF; // [Fh]
    for (int e = threadIdx.x; e < Fh * kSpF; e += kSpThreads) {
        const int fh = e / kSpF, f = e % kSpF;
        w1t[e] = f < Fi ? W1[(size_t)f * Fh + fh] : 0.0f;
        w2p[e] = f < Fo ? W2[(size_t)fh * Fo + f] : 0.0f;
    }
    for (int e = threadIdx.x; e < Fh; e += kSpThreads) b1s[e] = b1[e];
    __syncthreads();

    const long long r = (long long)blockIdx.x * kSpThreads + threadIdx.x;
    if (r >= R) return;
    const float* xr = X + (size_t)r * S * Fi;
    float* zr = Z + (size_t)r * S * Fo;
    const int s0 = blockIdx.y * s_chunk;
    const int s1 = s0 + s_chunk < S ? s0 + s_chunk : S;
    for (int s = s0; s < s1; ++s) {
        float agg[kSpF];
#pragma unroll
        for (int f = 0; f < kSpF; ++f) agg[f] = 0.0f;
        const int e0 = rowptr[s], e1 = rowptr[s + 1];
        for (int e = e0; e < e1; ++e) {
            const float a = __ldg(vals + e);
            const float* xs = xr + (size_t)__ldg(colidx + e) * Fi;
#pragma unroll
            for (int f = 0; f < kSpF; ++f)
                if (f < Fi) agg[f] = fmaf(a, __ldg(xs + f), agg[f]);
        }
        float z[kSpF];
#pragma unroll
        for (int f = 0; f < kSpF; ++f) z[f] = 0.0f;
#pragma unroll 2
        for (int fh = 0; fh < Fh; ++fh) {
            float h = 0.0f;
#pragma unroll
            for (int v = 0; v < kSpF / 4; ++v) {
                const float4 w = *reinterpret_cast<const float4*>(w1t + fh * kSpF + 4 * v);
                h = fmaf(agg[4 * v + 0], w.x, h);
                h = fmaf(agg[4 * v + 1], w.y, h);
                h = fmaf(agg[4 * v + 2], w.z, h);
                h = fmaf(agg[4 * v + 3], w.w, h);
            }
            h += b1s[fh];
            h = h < 0.0f ? 0.0f : h;  // ReLU of layer 1
#pragma unroll
            for (int v = 0; v < kSpF / 4; ++v) {
                const float4 w = *reinterpret_cast<const float4*>(w2p + fh * kSpF + 4 * v);
                z[4 * v + 0] = fmaf(h, w.x, z[4 * v + 0]);
                z[4 * v + 1] = fmaf(h, w.y, z[4 * v + 1]);
                z[4 * v + 2] = fmaf(h, w.z, z[4 * v + 2]);
                z[4 * v + 3] = fmaf(h, w.w, z[4 * v + 3]);
            }
        }
#pragma unroll
        for (int f = 0; f < kSpF; ++f)
            if (f < Fo) zr[(size_t)s * Fo + f] = z[f];
    }
}

// kernel 2: U[r][s*Fo + fo] = relu((A.Z[r])[s][fo] + b2[fo]), written as K-major 128-row tiles
// [R/128][ldo][128] with the columns beyond S*Fo zeroed
__global__ void __launch_bounds__(kSpThreads)
    gcn_sparse_l2_kernel(const float* __restrict__ Z, const int* __restrict__ rowptr,
                         const int* __restrict__ colidx, const float* __restrict__ vals,
                         const float* __restrict__ b2, float* __restrict__ U, long long R, int S, int Fo,
                         int ldo, int s_chunk) {
    const long long tile = blockIdx.x;
    const long long r = tile * kSpThreads + threadIdx.x;
    const bool live = r < R;
    const float* zr = Z + (size_t)(live ? r : 0) * S * Fo;
    float* ut = U + (size_t)tile * ldo * kSpThreads + threadIdx.x;
    const int s0 = blockIdx.y * s_chunk;
    const int s1 = s0 + s_chunk < S ? s0 + s_chunk : S;
    for (int s = s0; s < s1; ++s) {
        float acc[kSpF];
#pragma unroll
        for (int f = 0; f < kSpF; ++f) acc[f] = 0.0f;
        const int e0 = rowptr[s], e1 = rowptr[s + 1];
        for (int e = e0; e < e1; ++e) {
            const float a = __ldg(vals + e);
            const float* zs = zr + (size_t)__ldg(colidx + e) * Fo;
#pragma unroll
            for (int f = 0; f < kSpF; ++f)
                if (f < Fo) acc[f] = fmaf(a, __ldg(zs + f), acc[f]);
        }
#pragma unroll
        for (int f = 0; f < kSpF; ++f) {
            if (f < Fo) {
                float v = acc[f] + __ldg(b2 + f);
                v = v < 0.0f ? 0.0f : v;
                ut[(size_t)(s * Fo + f) * kSpThreads] = live ? v : 0.0f;
            }
        }
    }
    // zero the K padding once (the CTA of the last station chunk does it)
    if (s1 == S)
        for (int c = S * Fo; c < ldo; ++c) ut[(size_t)c * kSpThreads] = 0.0f;
}

}  // namespace wg
