// GCN layers — relu((A_hat . X) . W + b), aggregate-then-transform, one or two layers fused.
//
// Reference: GraphConvLayer.forward, src/step5_gcn_layer_model.py:13-23 (matmul order :15,:18,
// ReLU :21); chained twice by GCN_GRU.forward, src/step6_gcn_gru_combined_model.py:17,20, whose
// `.view(1, T, S*13)` (:20) is the flat index s*F_out + f used for the output rows here.
//
// Work decomposition (dense A_hat, S small enough for A_hat^T to sit in shared memory):
//   * a CTA owns RB consecutive rows (a row = one (sequence, timestep) pair = an [S, F] slab);
//   * a thread owns SG consecutive stations of one row: it keeps the SG x F aggregate
//     in registers, runs the S-long aggregation as register FFMAs (A_hat^T column group by
//     LDS.128, the S x F input slab by warp-broadcast LDS), then applies the F x F transform,
//     bias and ReLU from registers and writes the result slab to shared memory;
//   * layer 2 reads layer 1's slab from shared memory; only the final slab goes to HBM,
//     zero-padded to `ldo` columns so the input-projection GEMM needs no K-edge handling.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kGcnThreads = 128;

// Floats of dynamic shared memory the kernel needs.
template <int FP, int SG>
__host__ __device__ inline size_t gcn_smem_floats(int S, int Fi, int Fh, int Fo, int RB, int layers) {
    constexpr int SGS = (SG + 3) & ~3;
    constexpr int FPAD = (FP + 3) & ~3;
    const int NSG = ceil_div(S, SG);
    const int fmaxA = layers == 2 ? (Fi > Fo ? Fi : Fo) : Fi;
    size_t n = 0;
    n += (size_t)S * NSG * SGS;             // adjT
    n += 2 * (size_t)FPAD * FPAD;           // w1t, w2t ([fo][FPAD], fo < FPAD)
    n += 2 * (size_t)FPAD;                  // b1, b2
    n += (size_t)round_up(RB * S * fmaxA, 4);  // bufA
    n += (size_t)round_up(RB * S * (layers == 2 ? Fh : Fo), 4);  // bufB
    return n;
}

// One GCN layer on the CTA's row block, shared memory -> shared memory.
//   in  [rows][S*Fi], outp [rows][S*Fo]; adjT [S][NSG*SGS] with adjT[sp][q*SGS+i] = A_hat[q*SG+i][sp]
//   (zero where the station index is out of range); Wt [Fo][FPAD] (f contiguous, zero padded).
template <int FP, int SG, bool EXACT>
__device__ __forceinline__ void gcn_layer_smem(const float* __restrict__ in, float* __restrict__ outp,
                                               const float* __restrict__ adjT,
                                               const float* __restrict__ Wt,
                                               const float* __restrict__ bias, int S, int Fi, int Fo,
                                               int NSG, int row_local, int q) {
    constexpr int SGS = (SG + 3) & ~3;
    constexpr int FPAD = (FP + 3) & ~3;
    float acc[SG][FP];
#pragma unroll
    for (int i = 0; i < SG; ++i)
#pragma unroll
        for (int f = 0; f < FP; ++f) acc[i][f] = 0.0f;

    const float* xrow = in + (size_t)row_local * S * Fi;
    const float* arow = adjT + q * SGS;
    const int astride = NSG * SGS;
#pragma unroll 2
    for (int sp = 0; sp < S; ++sp) {
        float a[SGS];
#pragma unroll
        for (int v = 0; v < SGS / 4; ++v) {
            const float4 t = *reinterpret_cast<const float4*>(arow + (size_t)sp * astride + 4 * v);
            a[4 * v + 0] = t.x;
            a[4 * v + 1] = t.y;
            a[4 * v + 2] = t.z;
            a[4 * v + 3] = t.w;
        }
        float x[FP];
#pragma unroll
        for (int f = 0; f < FP; ++f) x[f] = (EXACT || f < Fi) ? xrow[sp * Fi + f] : 0.0f;
#pragma unroll
        for (int i = 0; i < SG; ++i)
#pragma unroll
            for (int f = 0; f < FP; ++f) acc[i][f] = fmaf(a[i], x[f], acc[i][f]);
    }

    float* orow = outp + (size_t)row_local * S * Fo;
    for (int fo = 0; fo < Fo; ++fo) {
        float w[FPAD];
#pragma unroll
        for (int v = 0; v < FPAD / 4; ++v) {
            const float4 t = *reinterpret_cast<const float4*>(Wt + fo * FPAD + 4 * v);
            w[4 * v + 0] = t.x;
            w[4 * v + 1] = t.y;
            w[4 * v + 2] = t.z;
            w[4 * v + 3] = t.w;
        }
        const float b = bias[fo];
#pragma unroll
        for (int i = 0; i < SG; ++i) {
            const int s = q * SG + i;
            float v = 0.0f;
#pragma unroll
            for (int f = 0; f < FP; ++f) v = fmaf(acc[i][f], w[f], v);
            v += b;
            v = v < 0.0f ? 0.0f : v;  // ReLU; NaN propagates like torch.relu
            if (s < S) orow[s * Fo + fo] = v;
        }
    }
}

// LAYERS == 2: out[r][c] (ld = ldo, c < ldo) = flatten(relu-GCN2(relu-GCN1(x[r]))) zero padded.
// LAYERS == 1: out[r][c] = flatten(relu-GCN1(x[r])), W2/b2 unused, Fh := Fo.
template <int FP, int SG, bool EXACT, int LAYERS>
__global__ void __launch_bounds__(kGcnThreads)
    gcn_kernel(const float* __restrict__ X, const float* __restrict__ adj, const float* __restrict__ W1,
               const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
               float* __restrict__ out, long long R, int S, int Fi, int Fh, int Fo, int ldo, int RB) {
    constexpr int SGS = (SG + 3) & ~3;
    constexpr int FPAD = (FP + 3) & ~3;
    extern __shared__ __align__(16) float smem[];
    const int NSG = ceil_div(S, SG);
    const int tid = threadIdx.x;
    const int Flast = LAYERS == 2 ? Fo : Fh;  // features of the final slab
    const int fmaxA = LAYERS == 2 ? (Fi > Fo ? Fi : Fo) : Fi;

    float* adjT = smem;
    float* w1t = adjT + (size_t)S * NSG * SGS;
    float* w2t = w1t + FPAD * FPAD;
    float* b1s = w2t + FPAD * FPAD;
    float* b2s = b1s + FPAD;
    float* bufA = b2s + FPAD;
    float* bufB = bufA + round_up(RB * S * fmaxA, 4);

    // ---- stage the (tiny) graph and layer parameters once per CTA ----
    for (int e = tid; e < S * NSG * SGS; e += kGcnThreads) {
        const int sp = e / (NSG * SGS);
        const int c = e % (NSG * SGS);
        const int qq = c / SGS, i = c % SGS;
        const int s = qq * SG + i;
        adjT[e] = (i < SG && s < S) ? adj[(size_t)s * S + sp] : 0.0f;
    }
    for (int e = tid; e < FPAD * FPAD; e += kGcnThreads) {
        const int fo = e / FPAD, f = e % FPAD;
        w1t[e] = (fo < Fh && f < Fi) ? W1[f * Fh + fo] : 0.0f;
        if (LAYERS == 2) w2t[e] = (fo < Fo && f < Fh) ? W2[f * Fo + fo] : 0.0f;
    }
    for (int e = tid; e < FPAD; e += kGcnThreads) {
        b1s[e] = e < Fh ? b1[e] : 0.0f;
        if (LAYERS == 2) b2s[e] = e < Fo ? b2[e] : 0.0f;
    }
    __syncthreads();

    const int row_local = tid / NSG;
    const int q = tid % NSG;
    const long long nblocks = (R + RB - 1) / RB;
    const int in_cols = S * Fi;
    const int out_cols = S * Flast;

    for (long long rb = blockIdx.x; rb < nblocks; rb += gridDim.x) {
        const long long r0 = rb * RB;
        const int nrows = (int)((R - r0) < RB ? (R - r0) : RB);
        // ---- coalesced copy of the row block (contiguous in HBM) ----
        const float* src = X + (size_t)r0 * in_cols;
        const int n_in = nrows * in_cols;
        for (int e = tid; e < n_in; e += kGcnThreads) bufA[e] = __ldg(src + e);
        __syncthreads();

        const bool active = row_local < nrows;
        if (active) gcn_layer_smem<FP, SG, EXACT>(bufA, bufB, adjT, w1t, b1s, S, Fi, Fh, NSG, row_local, q);
        __syncthreads();
        const float* fin = bufB;
        if (LAYERS == 2) {
            if (active) gcn_layer_smem<FP, SG, EXACT>(bufB, bufA, adjT, w2t, b2s, S, Fh, Fo, NSG, row_local, q);
            __syncthreads();
            fin = bufA;
        }
        // ---- coalesced store of the final slab, zero padded to ldo columns ----
        float* dst = out + (size_t)r0 * ldo;
        if ((ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            const int ld4 = ldo >> 2;
            const int n4 = nrows * ld4;
            for (int e = tid; e < n4; e += kGcnThreads) {
                const int row = e / ld4;
                const int c = (e % ld4) * 4;
                const float* p = fin + row * out_cols;
                float4 v;
                v.x = (c + 0 < out_cols) ? p[c + 0] : 0.0f;
                v.y = (c + 1 < out_cols) ? p[c + 1] : 0.0f;
                v.z = (c + 2 < out_cols) ? p[c + 2] : 0.0f;
                v.w = (c + 3 < out_cols) ? p[c + 3] : 0.0f;
                *reinterpret_cast<float4*>(dst + (size_t)row * ldo + c) = v;
            }
        } else {
            const int n_out = nrows * ldo;
            for (int e = tid; e < n_out; e += kGcnThreads) {
                const int row = e / ldo, c = e % ldo;
                dst[e] = c < out_cols ? fin[row * out_cols + c] : 0.0f;
            }
        }
        __syncthreads();
    }
}

}  // namespace wg
