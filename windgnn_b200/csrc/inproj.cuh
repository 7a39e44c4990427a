// GRU input projection — GI[r][n] = sum_k U[r][k] * W_ih[n][k] + bias[n] for every (sequence,
// timestep) row r at once.  It has no dependence on the hidden state, so it is hoisted out of
// the recurrence and run as one throughput GEMM (65 % of the path's FLOPs at S = 34).
//
// Reference: the `gi = W_ih u + b_ih` half of nn.GRU, called at
// src/step6_gcn_gru_combined_model.py:23 (module built at :11).  `bias` additionally carries
// b_hh for the r and z gates (they are added to the same pre-activation).
//
// FP32 FFMA GEMM, "NT" form (both operands K-contiguous):
//   CTA tile 128 (rows) x 64 (gate columns), K tile 16, 3-stage cp.async ring;
//   128 threads, 8 x 8 accumulators each; operand fragments are read with LDS.128 along K
//   from rows padded to 20 floats, with rows interleaved over the thread grid (row = ty + 16 i,
//   col = tx + 8 j) so that every LDS.128 is bank-conflict free (offsets 20*ty mod 32 distinct).
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kIpBM = 128;
constexpr int kIpBN = 64;
constexpr int kIpBK = 16;
constexpr int kIpLds = kIpBK + 4;  // padded row length in floats
constexpr int kIpStages = 3;
constexpr int kIpThreads = 128;
constexpr int kIpSmemBytes = kIpStages * (kIpBM + kIpBN) * kIpLds * 4;

// A  [M][K]   (U, K a multiple of 16, 16-B aligned rows)
// Bw [NP][K]  (packed w_ih, NP a multiple of 64, rows >= N are zero)
// C  [M][ldc] (GI); only columns < N are written.
__global__ void __launch_bounds__(kIpThreads, 3)
    inproj_kernel(const float* __restrict__ A, const float* __restrict__ Bw,
                  const float* __restrict__ bias, float* __restrict__ C, long long M, int N, int K,
                  int ldc, int n_tiles) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                   // [stage][BM][LDS]
    float* Bs = smem + kIpStages * kIpBM * kIpLds;      // [stage][BN][LDS]

    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const int nt = (int)(tile % n_tiles);
    const long long mt = tile / n_tiles;
    const long long m0 = mt * kIpBM;
    const int n0 = nt * kIpBN;

    const int ty = tid >> 3;  // 0..15
    const int tx = tid & 7;   // 0..7

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int KT = K / kIpBK;

    auto load_tile = [&](int kt, int slot) {
        const int k0 = kt * kIpBK;
        float* as = As + slot * kIpBM * kIpLds;
        float* bs = Bs + slot * kIpBN * kIpLds;
#pragma unroll
        for (int c = 0; c < (kIpBM * 4) / kIpThreads; ++c) {
            const int idx = tid + c * kIpThreads;
            const int row = idx >> 2, kc = idx & 3;
            const long long gm = m0 + row;
            const bool ok = gm < M;
            const float* src = A + (size_t)(ok ? gm : 0) * K + k0 + kc * 4;
            cp_async16(as + row * kIpLds + kc * 4, src, ok);
        }
#pragma unroll
        for (int c = 0; c < (kIpBN * 4) / kIpThreads; ++c) {
            const int idx = tid + c * kIpThreads;
            const int row = idx >> 2, kc = idx & 3;
            const float* src = Bw + (size_t)(n0 + row) * K + k0 + kc * 4;
            cp_async16(bs + row * kIpLds + kc * 4, src, true);
        }
    };

#pragma unroll
    for (int s = 0; s < kIpStages - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<kIpStages - 2>();
        __syncthreads();
        {
            const int nk = kt + kIpStages - 1;
            if (nk < KT) load_tile(nk, nk % kIpStages);
            cp_async_commit();
        }
        const float* as = As + (kt % kIpStages) * kIpBM * kIpLds;
        const float* bs = Bs + (kt % kIpStages) * kIpBN * kIpLds;
#pragma unroll
        for (int k4 = 0; k4 < kIpBK / 4; ++k4) {
            float4 a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = *reinterpret_cast<const float4*>(as + (ty + 16 * i) * kIpLds + k4 * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                b[j] = *reinterpret_cast<const float4*>(bs + (tx + 8 * j) * kIpLds + k4 * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: + bias, predicated store ----
    float bj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + tx + 8 * j;
        bj[j] = n < N ? __ldg(bias + n) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gm = m0 + ty + 16 * i;
        if (gm < M) {
            float* crow = C + (size_t)gm * ldc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + tx + 8 * j;
                if (n < N) crow[n] = acc[i][j] + bj[j];
            }
        }
    }
}

}  // namespace wg
