// dW_ih = dGI^T . U of the training step, reading both operands WHERE THE PRODUCERS LEFT THEM.
//
// Reference: `loss.backward()` at src/main.py:76 through the nn.GRU input half built at
// src/step6_gcn_gru_combined_model.py:11 — dW_ih [3H x I] = dGI^T . U, a contraction over the B*T
// (sequence, timestep) rows (split-K, partial sums added in a fixed order by sg_reduce_kernel).
// dGI is the row-major [B*T][4H] buffer the BPTT kernel writes (gate columns contiguous); U is in
// the forward's K-major 128-row tiles [r / 128][IP][128] (rows contiguous).  Round 1 re-laid both
// (rows_to_colblocks / tiles_to_colblocks: 1.3 ms and 2.3 GB of scratch per 4096-sequence step)
// so that the forward's projection kernel could bulk-copy them, and padded the 306 gate columns
// to 384.  Here U is contiguous along the CONTRACTION index, which is exactly what the broadcast
// side of the FFMA2 register tile wants:
//     acc2[i][jp] += (u[m_i][k], u[m_i][k]) * (dgi[k][2jp], dgi[k][2jp+1])
// u[m][k .. k+3] is ONE 16-byte shared-memory load when the A tile is stored [m][k], and the
// scalar operand of FFMA2 is an ordinary 32-bit register; dGI is stored [k][n] as in inproj.cuh.
// m = input feature i, n = gate column g, k = row.  Operand tiles arrive by 16-byte cp.async
// (LDGSTS, zero fill outside the matrices) on a 4-stage ring.  CTA tile 64 (i) x 160 (g): 7 x 2
// tiles cover 442 x 306 with 5.7 % padding; 8 x 10 register tile; 128 threads = one warp per
// scheduler (a 160-thread variant with 8 x 8 tiles ran at 51 % of the FMA pipe: the fifth warp
// doubles one scheduler's work and the per-tile barrier makes everyone wait for it; this one 70 %).
// (Tried: writing dGI transposed into the dU GEMM's K-major tiles from here as the tiles pass through shared
// memory — +0.28 ms on this kernel against the 0.32 ms of the separate rows_to_tiles pass: dropped.)
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kKbBK = 16;
constexpr int kKbLdA = kKbBK;       // A tile row stride (floats): 4 chunks of 16 bytes, chunk c of row m stored at
                                    // c ^ ((m >> 1) & 3) — LDGSTS quarter-warps (2 rows x 4 chunks) and the LDS of 2 / 4
                                    // consecutive rows at one k both touch every bank group once
constexpr int kKbStages = 4;
constexpr int kKbThreads = 128;     // 4 warps: one per scheduler (a 5-warp CTA leaves one scheduler with twice the work)

// BM x BN CTA tile, 8 x TN register tile per thread, (BM/8) x (BN/TN) = 128 threads
template <int BM, int BN, int TN>
struct KbCfg {
    static constexpr int kBM = BM, kBN = BN, kTN = TN;
    static constexpr int kThreads = kKbThreads;
    static constexpr int kStageFloats = BM * kKbLdA + kKbBK * BN;
    static constexpr int kSmemBytes = kKbStages * kStageFloats * 4;
    static_assert((BM / 8) * (BN / TN) == kKbThreads, "thread grid");
};

// 64 (i) x 160 (g), 8 x 10 per thread, <= 3 CTAs per SM
using KbWih = KbCfg<64, 160, 10>;

struct KbArgs {
    const float* a;      // U tiles
    const float* b;      // dGI rows
    float* c;            // partials [z][I][ldc]
    long long R;         // B * T rows
    int I;               // number of input features (rows of the partial matrices)
    int IP;              // U tile column count (K padding of the forward's projection)
    int ld_dg;           // row stride of dGI (4H rounded up to 4)
    int ldc;
    int m_tiles, n_tiles;
    long long kt_per_split;  // k-tiles per blockIdx.y
};

template <int BM, int BN, int TN>
__global__ void __launch_bounds__(kKbThreads, 3) gemm_kb_kernel(KbArgs g) {
    using Cfg = KbCfg<BM, BN, TN>;
    constexpr int NT = kKbThreads, TX = BN / TN, JP = TN / 2;
    static_assert(TN == 8 || (TN == 10 && BN == 160), "register tile widths served");
    static_assert((BM / 8) % 8 == 0, "thread rows share one swizzle");
    extern __shared__ __align__(128) float kb_smem[];
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int nt = (int)(blockIdx.x % g.n_tiles), mt = (int)(blockIdx.x / g.n_tiles);

    const long long KTall = (g.R + kKbBK - 1) / kKbBK;
    const long long kt0 = (long long)blockIdx.y * g.kt_per_split;
    const long long KT = KTall - kt0 < g.kt_per_split ? (KTall - kt0 > 0 ? KTall - kt0 : 0) : g.kt_per_split;

    // ---- loader: every thread owns a fixed set of 16-byte chunks of a stage — NA of the A tile (rows 32 apart,
    //      same k offset) and NB of the B tile (one k row, columns 32 floats apart) — so a stage costs NA + NB
    //      LDGSTS and two pointer increments.  Bounds that do not move with k are folded into the sizes
    //      (0 = zero fill) once. ----
    constexpr int NA = BM * 4 / NT, NB = (BN / 4) / 8;
    static_assert(BM * 4 % NT == 0 && (BN / 4) % 8 == 0 && NT / 8 == kKbBK, "chunk ownership");
    const int a_ml = tid >> 2, a_kc = (tid & 3) * 4;
    const int b_kl = tid >> 3, b_nc = (tid & 7) * 4;
    const long long k_first = kt0 * kKbBK;
    const int i0 = mt * BM + a_ml;
    const float* a_src;      // chunk 0 of the tile loaded next
    {
        const long long r = k_first + a_kc;
        a_src = g.a + ((size_t)(r >> 7) * g.IP + i0) * 128 + (r & 127);
    }
    const int col = nt * BN + b_nc;
    const float* b_src = g.b + (size_t)(k_first + b_kl) * g.ld_dg + col;
    int a_sz[NA], b_sz[NB];
#pragma unroll
    for (int j = 0; j < NA; ++j) a_sz[j] = i0 + 32 * j < g.IP ? 16 : 0;
#pragma unroll
    for (int j = 0; j < NB; ++j) b_sz[j] = col + 32 * j + 3 < g.ld_dg ? 16 : 0;   // columns 3H .. 4H-1 are harmless
    auto cp16 = [](float* dst, const float* src, int sz) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
    };
    long long kt_issued = 0;   // tiles handed to load_next so far (they are issued in order)
    auto load_next = [&]() {
        float* As = kb_smem + (size_t)(kt_issued % kKbStages) * Cfg::kStageFloats;
        float* a_dst = As + a_ml * kKbLdA + (a_kc ^ (((a_ml >> 1) & 3) << 2));   // rows 32 apart share the swizzle
        float* b_dst = As + BM * kKbLdA + b_kl * BN + b_nc;
        const long long k0 = k_first + kt_issued * kKbBK;
        if (k0 + kKbBK <= g.R) {   // all 16 rows exist
#pragma unroll
            for (int j = 0; j < NA; ++j) cp16(a_dst + j * 32 * kKbLdA, a_sz[j] ? a_src + (size_t)j * 32 * 128 : g.a, a_sz[j]);
#pragma unroll
            for (int j = 0; j < NB; ++j) cp16(b_dst + 32 * j, b_sz[j] ? b_src + 32 * j : g.b, b_sz[j]);
        } else {   // the tile that straddles the end: rows >= R hold no data (zero both operands there)
            const long long r = k0 + a_kc;
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                float* d = a_dst + j * 32 * kKbLdA;
                const float* sp = a_src + (size_t)j * 32 * 128;
#pragma unroll
                for (int e = 0; e < 4; ++e) d[e] = (a_sz[j] && r + e < g.R) ? __ldg(sp + e) : 0.0f;
            }
            const bool row_ok = k0 + b_kl < g.R;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int sz = row_ok ? b_sz[j] : 0;
                cp16(b_dst + 32 * j, sz ? b_src + 32 * j : g.b, sz);
            }
        }
        // advance one k-tile: 16 rows further inside the 128-row U tile, or on to the next tile
        a_src += (((k0 >> 4) & 7) == 7) ? (long long)g.IP * 128 - 112 : 16;
        b_src += (long long)kKbBK * g.ld_dg;
        ++kt_issued;
    };

#pragma unroll
    for (int s = 0; s < kKbStages - 1; ++s) {
        if (s < KT) load_next();
        cp_async_commit();
    }

    float2 acc[8][JP];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < JP; ++j) acc[i][j] = make_float2(0.0f, 0.0f);

    for (long long kt = 0; kt < KT; ++kt) {
        cp_async_wait<kKbStages - 2>();   // this thread's copies of tile kt have landed
        __syncthreads();                  // everyone's have; everyone is done with tile kt-1 (slot refilled below)
        if (kt + kKbStages - 1 < KT) load_next();
        cp_async_commit();
        const float* As = kb_smem + (size_t)(kt % kKbStages) * Cfg::kStageFloats;
        const float* ap = As + ty * kKbLdA;          // thread rows: ty + i * (BM / 8); BM / 8 is a multiple of 8, so
        const int swz = ((ty >> 1) & 3) << 2;        // all of them share ty's swizzle
        const float* bp0 = As + BM * kKbLdA + tx * 4;
#pragma unroll
        for (int kv = 0; kv < kKbBK; kv += 4) {
            float av[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 t = *reinterpret_cast<const float4*>(ap + i * (BM / 8) * kKbLdA + (kv ^ swz));
                av[i][0] = t.x; av[i][1] = t.y; av[i][2] = t.z; av[i][3] = t.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float* br = bp0 + (kv + kk) * BN;
                float2 bp[JP];
                {
                    const float4 b0 = *reinterpret_cast<const float4*>(br);
                    const float4 b1 = *reinterpret_cast<const float4*>(br + (TN == 8 ? BN / 2 : 64));
                    bp[0] = make_float2(b0.x, b0.y); bp[1] = make_float2(b0.z, b0.w);
                    bp[2] = make_float2(b1.x, b1.y); bp[3] = make_float2(b1.z, b1.w);
                    if (TN == 10) bp[JP - 1] = *reinterpret_cast<const float2*>(br + 128 - tx * 2);   // column 128 + tx*2
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 aa = make_float2(av[i][kk], av[i][kk]);
#pragma unroll
                    for (int j = 0; j < JP; ++j) acc[i][j] = __ffma2_rn(aa, bp[j], acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: rows ty + i * (BM / 8); columns tx*4 .. +3, (BN/2 | 64) + tx*4 .. +3 and, for the 10-wide
    //      tile, 128 + tx*2 .. +1 ----
    float* C = g.c + (size_t)blockIdx.y * (size_t)g.I * g.ldc;
    const int c0 = nt * BN + tx * 4, c1 = nt * BN + (TN == 8 ? BN / 2 : 64) + tx * 4, c2 = nt * BN + 128 + tx * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = mt * BM + ty + i * (BM / 8);
        if (gm < g.I) {
            float* crow = C + (size_t)gm * g.ldc;
            if (c0 < g.ldc)
                *reinterpret_cast<float4*>(crow + c0) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
            if (c1 < g.ldc)
                *reinterpret_cast<float4*>(crow + c1) = make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y);
            if (TN == 10 && c2 < g.ldc) *reinterpret_cast<float2*>(crow + c2) = acc[i][JP - 1];
        }
    }
}

}  // namespace wg
