// FP32 GEMM for the backward pass — C[M][N] = sum_k A(m,k) * B(k,n), FFMA2, optional split-K.
//
// The training step (reference: `loss.backward()` at src/main.py:76 through the nn.GRU built at
// src/step6_gcn_gru_combined_model.py:11 and the two GraphConvLayers) needs three contractions
// that are plain GEMMs over the B*T (sequence, timestep) rows:
//     dU    [BT x I ]  = dGI [BT x 3H] . W_ih [3H x I]
//     dW_ih [3H x I ]  = dGI^T . U                       (K = B*T: split-K, deterministic reduce)
//     dW_hh [3H x H ]  = dGH^T . H_prev                  (K = B*T, H_prev = out shifted one step)
// Their operands sit in HBM in whatever layout their producers wrote (row-major with K or with
// the output index contiguous, the forward's K-major 128-row tiles of U, `out` read one step
// back), so the loader is described per operand by `SgOperand` instead of demanding a re-pack.
//
// CTA tile 128 x 128 x 16, 256 threads, 8 x 8 register tile as float2[8][4]; operands are staged
// through registers (coalesced 16-byte loads along whichever index is contiguous, transposed on
// the way into shared memory when that index is k) and double-buffered in shared memory, so the
// global loads of tile kt+1 are in flight while tile kt is multiplied:
//     acc2[i][jp] += (a[i][k], a[i][k]) * (b[k][2jp], b[k][2jp+1])          (FFMA2)
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kSgBM = 128;
constexpr int kSgBN = 128;
constexpr int kSgBK = 16;
constexpr int kSgThreads = 256;
constexpr int kSgLd = kSgBM + 4;  // shared-memory row stride (floats): 16-byte aligned rows
constexpr int kSgSmemBytes = 2 * 2 * kSgBK * kSgLd * 4;

// One GEMM operand as a function (r, k) -> element, r = the output index (m for A, n for B).
struct SgOperand {
    const float* p;
    long long ld;   // leading dimension in floats
    int kcontig;    // 1: (r, k) at p[r * ld + k]         0: (r, k) at p[k * ld + r]
    int tiled;      // 1 (kcontig): the forward's U tiles, (r, k) at p[(k >> 7) * ld * 128 + r * 128 + (k & 127)]
    int vec;        // 16-byte loads are legal (p and ld 16-byte aligned)
    int period;     // > 0 (r-contiguous only): row k reads row k - 1, rows with k % period == 0 read zeros
                    //   (H_prev[b, t] = out[b, t - 1], zero at t = 0)
};

// 128 (r) x 16 (k) tile of `op` starting at (r0, k0) into registers: two 4-element groups per thread.
__device__ __forceinline__ void sg_load(const SgOperand& op, long long r0, long long R, long long k0,
                                        long long Kend, int tid, float4 (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * kSgThreads;
        float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (op.kcontig) {
            const long long r = r0 + (idx >> 2);
            const long long k = k0 + (idx & 3) * 4;
            if (r < R && k < Kend) {
                const float* src = op.tiled ? op.p + (k >> 7) * op.ld * 128 + r * 128 + (k & 127)
                                            : op.p + r * op.ld + k;
                if (op.vec && k + 3 < Kend) {
                    t = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    t.x = __ldg(src);
                    if (k + 1 < Kend) t.y = __ldg(src + 1);
                    if (k + 2 < Kend) t.z = __ldg(src + 2);
                    if (k + 3 < Kend) t.w = __ldg(src + 3);
                }
            }
        } else {
            const long long k = k0 + (idx >> 5);
            const long long r = r0 + (idx & 31) * 4;
            bool ok = k < Kend && r < R;
            long long krow = k;
            if (op.period > 0) {
                ok = ok && (k % op.period) != 0;
                krow = k - 1;
            }
            if (ok) {
                const float* src = op.p + krow * op.ld + r;
                if (op.vec && r + 3 < R) {
                    t = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    t.x = __ldg(src);
                    if (r + 1 < R) t.y = __ldg(src + 1);
                    if (r + 2 < R) t.z = __ldg(src + 2);
                    if (r + 3 < R) t.w = __ldg(src + 3);
                }
            }
        }
        v[i] = t;
    }
}

// registers -> shared tile S[k][r] (row stride kSgLd)
__device__ __forceinline__ void sg_store(const SgOperand& op, float* S, int tid, const float4 (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * kSgThreads;
        if (op.kcontig) {
            const int r = idx >> 2, k = (idx & 3) * 4;
            S[(k + 0) * kSgLd + r] = v[i].x;
            S[(k + 1) * kSgLd + r] = v[i].y;
            S[(k + 2) * kSgLd + r] = v[i].z;
            S[(k + 3) * kSgLd + r] = v[i].w;
        } else {
            const int k = idx >> 5, r = (idx & 31) * 4;
            *reinterpret_cast<float4*>(S + k * kSgLd + r) = v[i];
        }
    }
}

// C (splits == 1): [M][ldc] row-major.  splits > 1: partial sums go to Cpart[z][M][N] (dense) and
// sg_reduce_kernel adds them in a fixed order.  grid = (n tiles, m tiles, splits).
__global__ void __launch_bounds__(kSgThreads, 2)
    sgemm_kernel(SgOperand A, SgOperand Bo, float* __restrict__ C, long long ldc, long long M, int N,
                 long long K, long long k_per_split) {
    extern __shared__ __align__(16) float sg_smem[];
    float* As = sg_smem;                       // [2][16][kSgLd]
    float* Bs = sg_smem + 2 * kSgBK * kSgLd;   // [2][16][kSgLd]
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.y * kSgBM;
    const long long n0 = (long long)blockIdx.x * kSgBN;
    const long long kb = (long long)blockIdx.z * k_per_split;
    long long Kend = kb + k_per_split;
    if (Kend > K) Kend = K;
    const int KT = Kend > kb ? (int)((Kend - kb + kSgBK - 1) / kSgBK) : 0;

    const int tx = tid & 15;  // cols tx*4 + {0..3} and 64 + tx*4 + {0..3}
    const int ty = tid >> 4;  // rows ty*4 + {0..3} and 64 + ty*4 + {0..3}
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.0f, 0.0f);

    float4 ra[2], rb[2];
    if (KT > 0) {
        sg_load(A, m0, M, kb, Kend, tid, ra);
        sg_load(Bo, n0, N, kb, Kend, tid, rb);
        sg_store(A, As, tid, ra);
        sg_store(Bo, Bs, tid, rb);
    }
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < KT) {  // global loads of the next tile fly during this tile's FMAs
            sg_load(A, m0, M, kb + (long long)(kt + 1) * kSgBK, Kend, tid, ra);
            sg_load(Bo, n0, N, kb + (long long)(kt + 1) * kSgBK, Kend, tid, rb);
        }
        const float* as = As + cur * kSgBK * kSgLd + ty * 4;
        const float* bs = Bs + cur * kSgBK * kSgLd + tx * 4;
#pragma unroll
        for (int k = 0; k < kSgBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(as + k * kSgLd);
            const float4 a1 = *reinterpret_cast<const float4*>(as + k * kSgLd + 64);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + k * kSgLd);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + k * kSgLd + 64);
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
        if (kt + 1 < KT) {
            sg_store(A, As + (cur ^ 1) * kSgBK * kSgLd, tid, ra);
            sg_store(Bo, Bs + (cur ^ 1) * kSgBK * kSgLd, tid, rb);
        }
        __syncthreads();
    }

    // ---- epilogue ----
    float* dst = C;
    long long ld = ldc;
    if (gridDim.z > 1) {
        dst = C + (size_t)blockIdx.z * M * N;
        ld = N;
    }
    const bool pair_ok = (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
        float* crow = dst + (size_t)gm * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gn = n0 + (j < 2 ? tx * 4 + 2 * j : 64 + tx * 4 + 2 * (j - 2));
            if (gn + 1 < N && pair_ok) {
                *reinterpret_cast<float2*>(crow + gn) = acc[i][j];
            } else {
                if (gn < N) crow[gn] = acc[i][j].x;
                if (gn + 1 < N) crow[gn + 1] = acc[i][j].y;
            }
        }
    }
}

// out[m * s_m + n * s_n] = sum_z part[z][m][n], added in a fixed order (deterministic); also used to finish
// column sums (M == 1).
// ldp: row stride of the partial matrices (>= N).
// 8 lanes per output element: lane q adds the partials z = q, q + 8, ... (two independent chains), then the eight
// lane sums meet in a fixed shuffle tree — deterministic, and 8 x the threads of a one-thread-per-element loop (that
// loop was latency-bound: 54 us for 296 partials of a 102 x 102 result).  grid = ceil(8 M N / 256) blocks of 256.
__global__ void sg_reduce_kernel(const float* __restrict__ part, int splits, long long M, int N,
                                 float* __restrict__ out, long long s_m, long long s_n, int ldp) {
    const long long total = M * N;
    const size_t zstride = (size_t)M * ldp;
    const int q = threadIdx.x & 7;
    const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool ok = e < total;
    const long long ee = ok ? e : 0;
    const long long m = ee / N, n = ee - m * N;
    const float* pz = part + (size_t)m * ldp + n;
    float s0 = 0.0f, s1 = 0.0f;
    int z = q;
    for (; z + 8 < splits; z += 16) {
        s0 += pz[(size_t)z * zstride];
        s1 += pz[(size_t)(z + 8) * zstride];
    }
    if (z < splits) s0 += pz[(size_t)z * zstride];
    float sum = s0 + s1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (ok && q == 0) out[m * s_m + n * s_n] = sum;
}

}  // namespace wg
