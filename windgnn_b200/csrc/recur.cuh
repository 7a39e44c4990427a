// GRU recurrence — the serial part of nn.GRU (h0 = 0), all T steps inside ONE launch.
//
// Reference: `gru_out, _ = self.gru(hidden2)` at src/step6_gcn_gru_combined_model.py:23 (module
// built at :11), PyTorch gate convention (rows r, z, n):
//     gh = W_hh h + b_hh
//     r = sigma(gi_r + gh_r)   z = sigma(gi_z + gh_z)   n = tanh(gi_n + r * gh_n)
//     h' = (h - n) * z + n
// gi (with b_ih, and b_hh for r and z, already added) comes from the input-projection kernel.
//
// A CTA owns 32 sequences for all T steps.  W_hh^T ([k][n], zero padded to KP x NP) lives in
// shared memory for the whole kernel, the hidden state of the 32 sequences too; nothing but gi
// (read) and h (written) touches HBM inside the time loop.  The warps form TWO independent
// groups of 16 sequences that only ever synchronise among themselves (named barriers), so one
// group's gate phase (MUFU-bound) overlaps the other group's GEMM (FMA-bound).  Per step:
//   GEMM  : acc[b][n] = sum_k h[b][k] * W[k][n] on the FMA pipe with packed FFMA2 (pairs along n;
//           the hidden state is kept DUPLICATED in shared memory, (h, h), so both FFMA2 operands
//           are plain LDS.128 register pairs).  Warp tile 16 sequences x 32 gate columns, thread
//           tile 4 x 4, operand fragments double-buffered in registers.
//   merge : each thread adds its accumulators onto the gi tile that IT prefetched with cp.async
//           during the GEMM (16-byte chunks, no barrier needed: a thread only waits for its own
//           copies); the n-gate part of gh is kept apart because r multiplies it.
//   gates : one (sequence, hidden unit) item per thread-slot, lanes along the hidden index so
//           the h stores to HBM are coalesced.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kRcBT = 32;    // sequences per CTA
constexpr int kRcMaxQ = 6;   // gate items per thread (BT*H <= kRcMaxQ * threads)

// hidden state rows hold (h, h) pairs: 2*KP floats (+ pad so 4 consecutive rows hit distinct banks)
__host__ __device__ inline int recur_hs_stride(int KP) { return ((2 * KP / 4) & 1) ? 2 * KP : 2 * KP + 4; }
__host__ __device__ inline size_t recur_smem_floats(int KP, int NP, int GP, bool w_smem) {
    size_t n = 0;
    if (w_smem) n += (size_t)KP * NP;
    n += (size_t)kRcBT * recur_hs_stride(KP);  // hs2
    n += (size_t)kRcBT * GP;                   // gi tile (r,z columns become gi + gh)
    n += (size_t)kRcBT * KP;                   // gh of the n gate
    n += (size_t)KP;                           // b_hn
    return n;
}

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NWARPS, bool W_SMEM>
__global__ void __launch_bounds__(NWARPS * 32, 1)
    gru_recur_kernel(const float* __restrict__ GI, const float* __restrict__ WhT,
                     const float* __restrict__ bhn, float* __restrict__ out, long long B, int T, int H,
                     int ldg, int KP, int NP) {
    constexpr int NT = NWARPS * 32;
    constexpr int WG = NWARPS / 2;   // warps per group
    constexpr int NG = WG * 32;      // threads per group
    constexpr int GB = kRcBT / 2;    // sequences per group
    extern __shared__ __align__(16) float smem[];
    const int RS = recur_hs_stride(KP);
    float* Ws = smem;
    float* hs = smem + (W_SMEM ? (size_t)KP * NP : 0);
    float* gis = hs + kRcBT * RS;          // [32][ldg]
    float* ghn = gis + kRcBT * ldg;        // [32][KP]
    float* bns = ghn + kRcBT * KP;         // [KP]
    const float* Wsrc = W_SMEM ? Ws : WhT;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int grp = warp / WG;             // 0 or 1: which 16 sequences
    const int gwarp = warp - grp * WG;     // warp index inside the group
    const int gtid = tid - grp * NG;
    const long long b0 = (long long)blockIdx.x * kRcBT;
    const int H2 = 2 * H;

    if (W_SMEM) {
        const int n4 = KP * NP / 4;
        const float4* src = reinterpret_cast<const float4*>(WhT);
        float4* dst = reinterpret_cast<float4*>(Ws);
        for (int e = tid; e < n4; e += NT) dst[e] = __ldg(src + e);
    }
    for (int e = tid; e < kRcBT * RS; e += NT) hs[e] = 0.0f;
    for (int e = tid; e < kRcBT * ldg; e += NT) gis[e] = 0.0f;
    for (int e = tid; e < kRcBT * KP; e += NT) ghn[e] = 0.0f;
    for (int e = tid; e < KP; e += NT) bns[e] = e < H ? __ldg(bhn + e) : 0.0f;

    // ---- gate-phase items of this thread inside its group: (b, j) packed as b<<16 | j ----
    const int n_items = GB * H;
    int item_bj[kRcMaxQ];
#pragma unroll
    for (int q = 0; q < kRcMaxQ; ++q) {
        const int item = gtid + q * NG;
        int b = -1, j = 0;
        if (item < n_items) {
            b = item / H;
            j = item - b * H;
            b += grp * GB;
            if (b0 + b >= B) b = -1;  // ragged last CTA
        }
        item_bj[q] = b < 0 ? -1 : ((b << 16) | j);
    }

    // ---- GEMM-phase coordinates ----
    const int ng = lane & 7;   // column group within the warp tile
    const int bg = lane >> 3;  // row group within the warp tile
    const int n_nb = NP / 32;  // 32-column blocks; this warp takes gwarp, gwarp + WG, ...
    const int rbase = grp * GB + bg;       // rows rbase + 4 i

    // this thread's part of the gi tile of step t: 4 rows x 16 bytes per column block it owns
    auto prefetch_gi = [&](int t) {
        for (int nb = gwarp; nb < n_nb; nb += WG) {
            const int nbase = nb * 32 + ng * 4;
            if (nbase < ldg) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = rbase + 4 * i;
                    if (b0 + b < B)
                        cp_async16(gis + b * ldg + nbase, GI + ((size_t)(b0 + b) * T + t) * ldg + nbase, true);
                }
            }
        }
        cp_async_commit();
    };

    __syncthreads();   // zero fills are done before any async copy may land
    prefetch_gi(0);

    const float* hrow = hs + rbase * RS;
    for (int t = 0; t < T; ++t) {
        // ================= GEMM + merge =================
        for (int nb = gwarp; nb < n_nb; nb += WG) {
            const int nbase = nb * 32 + ng * 4;
            const float* wcol = Wsrc + nbase;
            float2 acc[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.0f, 0.0f);

            auto load_frag = [&](int k2, float4 (&hv)[4], float4 (&wv)[2]) {
#pragma unroll
                for (int i = 0; i < 4; ++i)   // (h[k2], h[k2], h[k2+1], h[k2+1]) of row rbase + 4 i
                    hv[i] = *reinterpret_cast<const float4*>(hrow + (4 * i) * RS + 2 * k2);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const float* wp = wcol + (size_t)(k2 + kk) * NP;
                    wv[kk] = W_SMEM ? *reinterpret_cast<const float4*>(wp)
                                    : __ldg(reinterpret_cast<const float4*>(wp));
                }
            };
            auto mma_frag = [&](const float4 (&hv)[4], const float4 (&wv)[2]) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 h0 = make_float2(hv[i].x, hv[i].y), h1 = make_float2(hv[i].z, hv[i].w);
                    acc[i][0] = __ffma2_rn(h0, make_float2(wv[0].x, wv[0].y), acc[i][0]);
                    acc[i][1] = __ffma2_rn(h0, make_float2(wv[0].z, wv[0].w), acc[i][1]);
                    acc[i][0] = __ffma2_rn(h1, make_float2(wv[1].x, wv[1].y), acc[i][0]);
                    acc[i][1] = __ffma2_rn(h1, make_float2(wv[1].z, wv[1].w), acc[i][1]);
                }
            };
            if (t > 0) {  // h_{-1} = 0: the product is zero at t == 0
                float4 hA[4], wA[2], hB[4], wB[2];
                load_frag(0, hA, wA);
#pragma unroll 1
                for (int k2 = 0; k2 < KP; k2 += 4) {  // KP is a multiple of 4
                    load_frag(k2 + 2, hB, wB);
                    mma_frag(hA, wA);
                    if (k2 + 4 < KP) load_frag(k2 + 4, hA, wA);
                    mma_frag(hB, wB);
                }
            }
            cp_async_wait<0>();  // this thread's chunks of gi(t) have landed
            if (nbase < ldg) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = rbase + 4 * i;
                    float* g = gis + b * ldg + nbase;
                    const float a4[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
                    if (nbase + 3 < H2) {          // whole chunk in the r / z gates: gi + gh
                        float4 v = *reinterpret_cast<float4*>(g);
                        v.x += a4[0]; v.y += a4[1]; v.z += a4[2]; v.w += a4[3];
                        *reinterpret_cast<float4*>(g) = v;
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int n = nbase + c;
                            if (n < H2) g[c] += a4[c];
                            else if (n < 3 * H) ghn[b * KP + (n - H2)] = a4[c];  // n gate: keep gh apart
                        }
                    }
                }
            }
        }
        group_barrier(1 + grp, NG);

        // ================= gate phase =================
#pragma unroll
        for (int q = 0; q < kRcMaxQ; ++q) {
            if (item_bj[q] >= 0) {
                const int b = item_bj[q] >> 16, j = item_bj[q] & 0xffff;
                const float* g = gis + b * ldg + j;
                const float r = sigmoid_f(g[0]);
                const float z = sigmoid_f(g[H]);
                const float n = tanh_f(g[H2] + r * (ghn[b * KP + j] + bns[j]));
                const float hold = hs[b * RS + 2 * j];
                const float hnew = (hold - n) * z + n;
                *reinterpret_cast<float2*>(hs + b * RS + 2 * j) = make_float2(hnew, hnew);
                out[((size_t)(b0 + b) * T + t) * H + j] = hnew;
            }
        }
        group_barrier(1 + grp, NG);
        if (t + 1 < T) prefetch_gi(t + 1);  // lands during the next GEMM
    }
}

}  // namespace wg
