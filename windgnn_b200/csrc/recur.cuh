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
//   GEMM  : acc[b][n] = sum_k h[b][k] * W[k][n] on the FMA pipe with packed FFMA2 (sm_100a):
//           acc2[b][(n, n+1)] += h[b][k] (scalar operand, broadcast) * (W[k][n], W[k][n+1]).
//           A warp covers 16 sequences x 80 gate columns, a thread 4 x 10 (five column pairs);
//           operand fragments of four k's are double-buffered in registers so the LDS latency
//           hides behind the previous fragment's 80 FFMA2.
//   merge : each thread adds its accumulators onto the gi chunks that IT prefetched with
//           cp.async during the GEMM (no barrier needed: a thread only waits for its own
//           copies); the n-gate part of gh is kept apart because r multiplies it.
//   gates : one (sequence, hidden unit) item per thread-slot, lanes along the hidden index so
//           the h stores to HBM are coalesced.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kRcBT = 32;    // sequences per CTA (two groups of 16)
constexpr int kRcCB = 80;    // gate columns per warp
constexpr int kRcMaxQ = 14;  // gate items per thread: 16 * H <= kRcMaxQ * 32 * warps_per_group

__host__ __device__ inline int recur_np(int G) { return round_up(G, kRcCB); }
// hidden-state rows: KP floats padded so that four consecutive rows start in distinct bank groups
__host__ __device__ inline int recur_hs_stride(int KP) { return ((KP / 4) & 1) ? KP : KP + 4; }
__host__ __device__ inline size_t recur_smem_floats(int KP, int NP, int GP, bool w_smem) {
    size_t n = 0;
    if (w_smem) n += (size_t)KP * NP;
    n += (size_t)kRcBT * recur_hs_stride(KP);  // hs
    n += (size_t)kRcBT * GP;                   // gi tile (r,z columns become gi + gh)
    n += (size_t)kRcBT * KP;                   // gh of the n gate
    n += (size_t)KP;                           // b_hn
    return n;
}

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void group_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

template <int N>
struct IntC { static constexpr int value = N; };

#ifndef WG_GATE_EXP
#define WG_GATE_EXP 0
#endif
#ifdef WG_RC_TRACE
// debug build only: per-phase clock stamps of CTA 0 (8 stamps per step per group)
__device__ long long g_rc_trace[2 * 8 * 256];
#define WG_TRACE(slot)                                                                          \
    do {                                                                                        \
        if (blockIdx.x == 0 && gtid == 0 && t < 256) g_rc_trace[(grp * 256 + t) * 8 + (slot)] = clock64(); \
    } while (0)
#else
#define WG_TRACE(slot) do { } while (0)
#endif

// one operand fragment: four k's of this thread's 4 rows and 10 columns
struct RcFrag {
    float4 h[4];      // h[i] = hs[row i][k .. k+3]
    float4 wa[4];     // W[k + kk][cA .. cA+3]
    float4 wb[4];     // W[k + kk][cB .. cB+3]
    float2 wc[4];     // W[k + kk][cC .. cC+1]
};

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
    const int ln = lane & 7;   // column slot within the warp's 80 columns
    const int bg = lane >> 3;  // row group
    const int n_cb = NP / kRcCB;           // 80-column blocks; this warp takes gwarp, gwarp + WG, ...
    const int rbase = grp * GB + bg;       // rows rbase + 4 i

    // this thread's chunks of the gi tile of step t: per row 16 + 16 + 8 bytes per column block
    auto prefetch_gi = [&](int t) {
        for (int cb = gwarp; cb < n_cb; cb += WG) {
            const int cA = cb * kRcCB + ln * 4, cB = cA + 32, cC = cb * kRcCB + 64 + ln * 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = rbase + 4 * i;
                if (b0 + b < B) {
                    const float* src = GI + ((size_t)(b0 + b) * T + t) * ldg;
                    float* dst = gis + b * ldg;
                    if (cA < ldg) cp_async16(dst + cA, src + cA, true);
                    if (cB < ldg) cp_async16(dst + cB, src + cB, true);
                    if (cC < ldg) cp_async8(dst + cC, src + cC);
                }
            }
        }
        cp_async_commit();
    };

    __syncthreads();   // zero fills are done before any async copy may land
    prefetch_gi(0);

    const float* hrow = hs + rbase * RS;
    const int h_step = 4 * RS;
#ifndef WG_RC_ALTERNATE
#define WG_RC_ALTERNATE 1
#endif
    for (int t = 0; t < T; ++t) {
        WG_TRACE(0);
#if WG_RC_ALTERNATE
        // The two groups take turns on the FMA pipe: group 0 runs GEMM(t) while group 1 does the
        // merge / gates of step t-1, then they swap.  Barrier 3: "group 0 finished its GEMM",
        // barrier 4: "group 1 finished its GEMM" (arrive = signal, sync = wait; NT participants).
        if (grp == 0) {
            if (t > 0) group_barrier(4, NT);
        } else {
            group_barrier(3, NT);
        }
#endif
        // ================= GEMM + merge =================
        WG_TRACE(1);
        bool handed_over = false;
        for (int cb = gwarp; cb < n_cb; cb += WG) {
            const int cA = cb * kRcCB + ln * 4, cB = cA + 32, cC = cb * kRcCB + 64 + ln * 2;
            float2 acc[4][5];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int p = 0; p < 5; ++p) acc[i][p] = make_float2(0.0f, 0.0f);

            const float* hp = hrow;            // + 4 floats per fragment
            const float* wp = Wsrc + cA;       // + 4 * NP floats per fragment
            auto load_frag = [&](RcFrag& f) {
#pragma unroll
                for (int i = 0; i < 4; ++i) f.h[i] = *reinterpret_cast<const float4*>(hp + i * h_step);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float* w = wp + kk * NP;
                    if (W_SMEM) {
                        f.wa[kk] = *reinterpret_cast<const float4*>(w);
                        f.wb[kk] = *reinterpret_cast<const float4*>(w + 32);
                        f.wc[kk] = *reinterpret_cast<const float2*>(w + 64 - 2 * ln);
                    } else {
                        f.wa[kk] = __ldg(reinterpret_cast<const float4*>(w));
                        f.wb[kk] = __ldg(reinterpret_cast<const float4*>(w + 32));
                        f.wc[kk] = __ldg(reinterpret_cast<const float2*>(w + 64 - 2 * ln));
                    }
                }
                hp += 4;
                wp += 4 * NP;
            };
            auto mma_frag = [&](const RcFrag& f) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float hv = kk == 0 ? f.h[i].x : kk == 1 ? f.h[i].y : kk == 2 ? f.h[i].z : f.h[i].w;
                        const float2 hh = make_float2(hv, hv);  // FFMA2 broadcasts a scalar operand
                        acc[i][0] = __ffma2_rn(hh, make_float2(f.wa[kk].x, f.wa[kk].y), acc[i][0]);
                        acc[i][1] = __ffma2_rn(hh, make_float2(f.wa[kk].z, f.wa[kk].w), acc[i][1]);
                        acc[i][2] = __ffma2_rn(hh, make_float2(f.wb[kk].x, f.wb[kk].y), acc[i][2]);
                        acc[i][3] = __ffma2_rn(hh, make_float2(f.wb[kk].z, f.wb[kk].w), acc[i][3]);
                        acc[i][4] = __ffma2_rn(hh, f.wc[kk], acc[i][4]);
                    }
                }
            };
            if (t > 0) {  // h_{-1} = 0: the product is zero at t == 0
                // software pipeline over KP / 4 fragments; branch-free body so the fragment
                // registers are written by the loads themselves
                RcFrag fa, fb;
                load_frag(fa);
                int k4 = 4;
#pragma unroll 1
                for (; k4 + 4 < KP; k4 += 8) {
                    load_frag(fb);
                    mma_frag(fa);
                    load_frag(fa);
                    mma_frag(fb);
                }
                if (k4 < KP) {   // even number of fragments: one more pair
                    load_frag(fb);
                    mma_frag(fa);
                    mma_frag(fb);
                } else {
                    mma_frag(fa);
                }
            }
#if WG_RC_ALTERNATE
            if (cb + WG >= n_cb) {  // last column block of this warp: hand the FMA pipe over
                if (grp == 0) group_arrive(3, NT);
                else if (t + 1 < T) group_arrive(4, NT);
                handed_over = true;
            }
#endif
            WG_TRACE(2);
            cp_async_wait<0>();  // this thread's chunks of gi(t) have landed
            // merge one aligned chunk of W columns [n0, n0 + W): r/z columns accumulate onto gi, n-gate
            // columns go to ghn; chunks that straddle a gate boundary take the element path
            auto merge_chunk = [&](float* g, float* gn, int n0, const float* a, auto width) {
                constexpr int W = decltype(width)::value;
                if (n0 + W <= H2) {
                    if (W == 4) {
                        float4 v = *reinterpret_cast<float4*>(g + n0);
                        v.x += a[0]; v.y += a[1]; v.z += a[2]; v.w += a[3];
                        *reinterpret_cast<float4*>(g + n0) = v;
                    } else {
                        float2 v = *reinterpret_cast<float2*>(g + n0);
                        v.x += a[0]; v.y += a[1];
                        *reinterpret_cast<float2*>(g + n0) = v;
                    }
                } else if (n0 >= H2 && n0 + W <= 3 * H) {
#pragma unroll
                    for (int c = 0; c < W; ++c) gn[n0 - H2 + c] = a[c];
                } else {
#pragma unroll
                    for (int c = 0; c < W; ++c) {
                        const int n = n0 + c;
                        if (n < H2) g[n] += a[c];
                        else if (n < 3 * H) gn[n - H2] = a[c];
                    }
                }
            };
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = rbase + 4 * i;
                float* g = gis + b * ldg;
                float* gn = ghn + b * KP;
                const float aA[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
                const float aB[4] = {acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y};
                const float aC[2] = {acc[i][4].x, acc[i][4].y};
                merge_chunk(g, gn, cA, aA, IntC<4>{});
                merge_chunk(g, gn, cB, aB, IntC<4>{});
                merge_chunk(g, gn, cC, aC, IntC<2>{});
            }
        }
#if WG_RC_ALTERNATE
        if (!handed_over) {  // a warp without a column block still takes part in the hand-over
            if (grp == 0) group_arrive(3, NT);
            else if (t + 1 < T) group_arrive(4, NT);
        }
#else
        (void)handed_over;
#endif
        WG_TRACE(3);
        group_barrier(1 + grp, NG);
        WG_TRACE(4);

        // ================= gate phase =================
        // Two passes: first every item's new state is computed into registers (loads and MUFU
        // chains of different items are independent, so they overlap), only then the stores —
        // a store to hs in between would order every later shared-memory load behind it.
        {
            float hnew[kRcMaxQ];
#pragma unroll
            for (int q = 0; q < kRcMaxQ; ++q) {
                const int pk = item_bj[q] >= 0 ? item_bj[q] : (grp * GB) << 16;  // invalid slot: harmless item
                const int b = pk >> 16, j = pk & 0xffff;
                const float* g = gis + b * ldg + j;
#if WG_GATE_EXP == 2
                const float r = g[0] * 0.25f;
                const float z = g[H] * 0.25f;
                const float n = (g[H2] + r * (ghn[b * KP + j] + bns[j])) * 0.25f;
#else
                const float r = sigmoid_f(g[0]);
                const float z = sigmoid_f(g[H]);
                const float n = tanh_f(g[H2] + r * (ghn[b * KP + j] + bns[j]));
#endif
                hnew[q] = (hs[b * RS + j] - n) * z + n;
            }
            float* out_t = out + (size_t)b0 * T * H + (size_t)t * H;
#pragma unroll
            for (int q = 0; q < kRcMaxQ; ++q) {
                if (item_bj[q] >= 0) {
                    const int b = item_bj[q] >> 16, j = item_bj[q] & 0xffff;
#if WG_GATE_EXP != 3
                    hs[b * RS + j] = hnew[q];
#endif
#if WG_GATE_EXP != 1
                    out_t[(size_t)b * T * H + j] = hnew[q];
#endif
                }
            }
        }
        WG_TRACE(5);
        group_barrier(1 + grp, NG);
        WG_TRACE(6);
        if (t + 1 < T) prefetch_gi(t + 1);  // lands during the next GEMM
        WG_TRACE(7);
    }
}

}  // namespace wg
