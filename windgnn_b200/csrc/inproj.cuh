// GRU input projection — GI[r][n] = sum_k U[r][k] * W_ih[n][k] + bias[n] for every (sequence,
// timestep) row r at once.  It has no dependence on the hidden state, so it is hoisted out of
// the recurrence and run as one throughput GEMM (65 % of the path's FLOPs at S = 34).
//
// Reference: the `gi = W_ih u + b_ih` half of nn.GRU, called at
// src/step6_gcn_gru_combined_model.py:23 (module built at :11).  `bias` additionally carries
// b_hh for the r and z gates (they are added to the same pre-activation).
//
// FP32 GEMM on the FMA pipe with packed FFMA2 (sm_100a `fma.rn.f32x2`: two fp32 FMAs per
// instruction, so the 8 x 8 register tile needs 32 issue slots per k instead of 64 and its
// operands are 64-bit register pairs — measured at the full FP32 rate where scalar FFMA tops
// out at ~90 % on register-bank conflicts, see scripts/probes/ffma_probe.cu):
//   acc2[i][jp] += (a[i][k], a[i][k]) * (b[2jp][k], b[2jp+1][k])
//
// Both operands are stored K-major in HBM, in tiles, by their producers (the GCN kernel writes
// U as [M/128][K][128], the pack kernel writes w_ih as [N/64][K][64]), so one pipeline stage
// (16 k's: an 8 KB A slab + a 4 KB B slab) is TWO bulk async copies (cp.async.bulk, the 1-D TMA
// path) issued by one thread and completed on an mbarrier — no per-thread address arithmetic in
// the main loop.  CTA tile 128 x 64, 128 threads, 4-stage ring, <= 4 CTAs per SM.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kIpBM = 128;
constexpr int kIpBN = 64;
constexpr int kIpBK = 16;
constexpr int kIpStages = 4;
constexpr int kIpThreads = 128;
constexpr int kIpStageFloats = kIpBK * (kIpBM + kIpBN);
constexpr int kIpStageBytes = kIpStageFloats * 4;
constexpr int kIpSmemBytes = kIpStages * kIpStageBytes + 64;  // + mbarriers

// At [M/128][K][128] (U, K a multiple of 16), Bt [NP/64][K][64] (packed w_ih, zero padded),
// C [M][ldc] row-major (GI); columns < ldc (a multiple of 4) are written.
__global__ void __launch_bounds__(kIpThreads, 4)
    inproj_kernel(const float* __restrict__ At, const float* __restrict__ Bt,
                  const float* __restrict__ bias, float* __restrict__ C, long long M, int K, int ldc,
                  int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kIpStages * kIpStageBytes);

    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const int nt = (int)(tile % n_tiles);
    const long long mt = tile / n_tiles;
    const int KT = K / kIpBK;
    const float* a_src = At + (size_t)mt * K * kIpBM;
    const float* b_src = Bt + (size_t)nt * K * kIpBN;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kIpStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int kt) {  // one thread: arm the barrier, launch the two bulk copies
        const int slot = kt % kIpStages;
        float* dst = stage0 + slot * kIpStageFloats;
        mbar_expect_tx(&bars[slot], kIpStageBytes);
        bulk_g2s(dst, a_src + (size_t)kt * kIpBK * kIpBM, kIpBK * kIpBM * 4, &bars[slot]);
        bulk_g2s(dst + kIpBK * kIpBM, b_src + (size_t)kt * kIpBK * kIpBN, kIpBK * kIpBN * 4, &bars[slot]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kIpStages - 1; ++s)
            if (s < KT) issue(s);
    }

    const int ty = tid >> 3;  // 0..15 : rows ty*4 + {0..3} and 64 + ty*4 + {0..3}
    const int tx = tid & 7;   // 0..7  : cols tx*4 + {0..3} and 32 + tx*4 + {0..3}

    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.0f, 0.0f);

    for (int kt = 0; kt < KT; ++kt) {
        __syncthreads();  // every thread is done with tile kt-1, whose slot is refilled next
        if (tid == 0 && kt + kIpStages - 1 < KT) issue(kt + kIpStages - 1);
        const int slot = kt % kIpStages;
        mbar_wait(&bars[slot], (kt / kIpStages) & 1);
        const float* as = stage0 + slot * kIpStageFloats + ty * 4;
        const float* bs = stage0 + slot * kIpStageFloats + kIpBK * kIpBM + tx * 4;
#pragma unroll
        for (int k = 0; k < kIpBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(as + k * kIpBM);
            const float4 a1 = *reinterpret_cast<const float4*>(as + k * kIpBM + 64);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + k * kIpBN);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + k * kIpBN + 32);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w),
                                  make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(aa, bp[j], acc[i][j]);
            }
        }
    }

    // ---- epilogue: + bias, 16-byte stores (a warp row covers 128 contiguous bytes) ----
    const int n0 = nt * kIpBN;
    const float4 bia0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + tx * 4));
    const float4 bia1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + 32 + tx * 4));
    const long long m0 = mt * kIpBM;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm < M) {
            float* crow = C + (size_t)gm * ldc;
            const int c0 = n0 + tx * 4, c1 = n0 + 32 + tx * 4;
            if (c0 < ldc)
                *reinterpret_cast<float4*>(crow + c0) = make_float4(
                    acc[i][0].x + bia0.x, acc[i][0].y + bia0.y, acc[i][1].x + bia0.z, acc[i][1].y + bia0.w);
            if (c1 < ldc)
                *reinterpret_cast<float4*>(crow + c1) = make_float4(
                    acc[i][2].x + bia1.x, acc[i][2].y + bia1.y, acc[i][3].x + bia1.z, acc[i][3].y + bia1.w);
        }
    }
}

}  // namespace wg
