// Backward of the two fused GCN layers of the 13-feature model, second generation:
// lane = (row, half), half-warp = station group.
//
// Reference: autograd through GraphConvLayer.forward (src/step5_gcn_layer_model.py:13-23) applied
// twice (src/step6_gcn_gru_combined_model.py:17,20); the math is stated at the top of gcn_bwd.cuh:
//   AX = A.X   G1 = relu(AX.W1 + b1)   AG = A.G1   Z2 = AG.W2 + b2
//   dZ2 = dU * [Z2 > 0]     dW2 += AG^T dZ2     db2 += colsum(dZ2)
//   dAG = dZ2 W2^T          dG1 = A^T dAG       dZ1 = dG1 * [G1 > 0]
//   dW1 += AX^T dZ1         db1 += colsum(dZ1)
// AX, G1 and AG are recomputed from X with exactly the forward's instruction sequence
// (gcn_rows.cuh: s' ascending, f ascending, bias last), so the ReLU masks are the forward's.
//
// gcn_bwd_kernel (thread = (row, 4 stations), 24-row blocks, three 16-float-padded slabs = 222 KB,
// one 8-warp CTA per SM) ran at 36 % of the FMA pipe and was the slowest kernel of the training
// step.  Here
//   * a CTA owns 16 consecutive rows; a warp is two half-warps, lane & 15 = row, and each of the
//     CTA's 8 half-warps ("units") owns up to 3 station PAIRS.  The adjacency pair is uniform
//     over a half-warp (one broadcast LDS.128 + LDS.64 per s'), the feature value is per lane:
//         acc2[(s, s+1)][f] += x[s'][f] * (A[s][s'], A[s+1][s'])                      (FFMA2)
//     The 13 x 13 transforms use the accumulator pair as the operand pair (gcn_rows.cuh);
//   * 16-row slabs with 14 floats per station (8-byte loads, row stride odd in 8-byte units:
//     conflict free) make three slabs + tables 110 KB: TWO 4-warp CTAs per SM, one warp of each
//     per scheduler; the unit with one more pair rotates with the CTA's wave;
//   * the two 13 x 13 outer-product sums over all (row, station) pairs are a small register-tiled
//     GEMM out of the slabs: thread = 8 (f) x 4 (fo) tile x one of 16 slices of the (row, station)
//     index, FFMA2 with the A value as the broadcast scalar.  Column 13 of the A-side slabs holds
//     1.0, so row 13 of each product IS the bias gradient.  The accumulators persist in registers
//     across the CTA's row blocks; partials go to HBM once (gcn_bwd_finish_kernel adds them in a
//     fixed order);
//   * dZ2 is formed in place in the dU block (which arrives by one bulk copy); the next block's X
//     is prefetched as soon as the slab it lands in is dead.
// Serves F_in = F_hid = F_out = 13 (the widths the reference's forward hard-codes, step6:16) and
// S <= 48; other shapes run gcn_bwd_kernel.
#pragma once

#include "gcn_rows.cuh"
#include "wg_common.cuh"

namespace wg {

constexpr int kGb2Rows = 16;       // rows per block (= lanes of a half-warp)
constexpr int kGb2Warps = 4;
constexpr int kGb2Threads = kGb2Warps * 32;
constexpr int kGb2Units = 2 * kGb2Warps;
constexpr int kGb2MaxPairs = 3;    // station pairs per unit (accumulators: 3 x 13 float2)
constexpr int kGb2FS = 14;         // slab floats per station: 13 features + one slot (1.0 on the A side)
constexpr int kGb2AP = 8;          // adjacency floats per (s', unit): up to 3 pairs, 16-byte aligned
constexpr int kGb2Slices = kGb2Threads / 8;   // reduction: 8 output tiles x 16 slices
constexpr int kGb2PartFloats = 2 * 256 + 32;  // per-slice partials: dW1 16x16, dW2 16x16, db1 16, db2 16

__host__ __device__ inline bool gcn_bwd_rows_applies(int S, int Fi, int Fh, int Fo) {
    return Fi == kGrF && Fh == kGrF && Fo == kGrF && ceil_div(ceil_div(S, 2), kGb2Units) <= kGb2MaxPairs;
}
// slab row stride (floats): even, and odd in 8-byte units, so that the 16 rows of a half-warp hit 16 distinct
// 8-byte bank groups
__host__ __device__ inline int gcn_bwd_rows_rs(int S) { return S * kGb2FS + ((S & 1) ? 4 : 2); }
__host__ __device__ inline int gcn_bwd_rows_region(int S, int ldu) {   // floats per region: a slab, X block or dU block
    int n = kGb2Rows * gcn_bwd_rows_rs(S);
    if (kGb2Rows * S * kGrF > n) n = kGb2Rows * S * kGrF;
    if (kGb2Rows * ldu > n) n = kGb2Rows * ldu;
    return round_up(n, 4);
}
__host__ __device__ inline size_t gcn_bwd_rows_smem_floats(int S, int ldu) {
    size_t n = 2 * (size_t)S * kGb2Units * kGb2AP;          // adjT (A[own][s']) and adjN (A[s'][own])
    n += 3 * (size_t)kGcnFS * kGcnFS + 2 * kGcnFS;         // W1, W2, W2^T (zero padded 16 x 16), b1, b2
    n += 3 * (size_t)gcn_bwd_rows_region(S, ldu);
    n += 4;                                                 // two mbarriers
    return n;
}

// acc[p][f] = sum_{s'} x[s'][f] * (a[2p], a[2p+1])  with x a 14-float-per-station slab row
template <int NPW>
__device__ __forceinline__ void gb2_aggregate(float2 (&acc)[NPW][kGrF], const float* __restrict__ xrow,
                                              const float* __restrict__ arow, int astride, int S) {
#pragma unroll
    for (int p = 0; p < NPW; ++p)
#pragma unroll
        for (int f = 0; f < kGrF; ++f) acc[p][f] = make_float2(0.0f, 0.0f);
#pragma unroll 2
    for (int sp = 0; sp < S; ++sp) {
        float a[8];
        const float* ap = arow + (size_t)sp * astride;
        {
            const float4 t = *reinterpret_cast<const float4*>(ap);
            a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
            if (NPW > 2) {
                const float2 u = *reinterpret_cast<const float2*>(ap + 4);
                a[4] = u.x; a[5] = u.y;
            }
        }
        float x[kGb2FS];
        const float* xp = xrow + sp * kGb2FS;
#pragma unroll
        for (int q = 0; q < kGb2FS / 2; ++q) {
            const float2 t = *reinterpret_cast<const float2*>(xp + 2 * q);
            x[2 * q] = t.x; x[2 * q + 1] = t.y;
        }
#pragma unroll
        for (int p = 0; p < NPW; ++p) {
            const float2 aa = make_float2(a[2 * p], a[2 * p + 1]);
#pragma unroll
            for (int f = 0; f < kGrF; ++f) acc[p][f] = __ffma2_rn(make_float2(x[f], x[f]), aa, acc[p][f]);
        }
    }
}

// store the accumulators of this unit's stations as slab rows; column 13 = c13
template <int NPW>
__device__ __forceinline__ void gb2_store_acc(float* __restrict__ srow, const float2 (&acc)[NPW][kGrF], int s0, int s_end,
                                              float c13) {
#pragma unroll
    for (int p = 0; p < NPW; ++p) {
        const int s = s0 + 2 * p;
        if (s < s_end) {
            float* d = srow + s * kGb2FS;
#pragma unroll
            for (int q = 0; q < 6; ++q) *reinterpret_cast<float2*>(d + 2 * q) = make_float2(acc[p][2 * q].x, acc[p][2 * q + 1].x);
            *reinterpret_cast<float2*>(d + 12) = make_float2(acc[p][12].x, c13);
        }
        if (s + 1 < s_end) {
            float* d = srow + (s + 1) * kGb2FS;
#pragma unroll
            for (int q = 0; q < 6; ++q) *reinterpret_cast<float2*>(d + 2 * q) = make_float2(acc[p][2 * q].y, acc[p][2 * q + 1].y);
            *reinterpret_cast<float2*>(d + 12) = make_float2(acc[p][12].y, c13);
        }
    }
}

// Persistent accumulators of one thread of the reduction: an 8 (f) x 4 (fo) tile of dW1 and of dW2.
struct Gb2Tiles {
    float2 w1[8][2], w2[8][2];
};

// C[f0 + i][fo0 .. fo0+3] += sum_k A[k][f0 + i] * B[k][fo0 .. fo0+3], k = (row, station) pairs slice, slice + nslices, ...
// A: slab (14 floats per station, row stride rsa).  B: station stride bs floats, row stride rsb; BVEC: 8-byte loads.
// Loads run past the 13 (14) valid columns of a station into the next station / the row padding (always
// inside the region): those products land in rows / columns 13 .. 15 of the tiles that nobody reads.
template <bool BVEC>
__device__ __forceinline__ void gb2_reduce(float2 (&c)[8][2], const float* __restrict__ A, int rsa,
                                           const float* __restrict__ B, int rsb, int bs, int nrows, int S, int f0, int fo0,
                                           int slice, int nslices) {
    const int K = nrows * S;
    const int cnt = slice < K ? (K - slice + nslices - 1) / nslices : 0;
    int row = 0, s = slice;
    while (s >= S) { s -= S; ++row; }
    const float* pa = A + (size_t)row * rsa + s * kGb2FS + f0;
    const float* pb = B + (size_t)row * rsb + s * bs + fo0;
    // one step = nslices stations further; q_wrap row wraps of it are certain, one more when s runs past S
    const int q_wrap = nslices / S, s_step = nslices - q_wrap * S;
    const int da = nslices * kGb2FS + q_wrap * (rsa - S * kGb2FS), db = nslices * bs + q_wrap * (rsb - S * bs);
    const int wa = rsa - S * kGb2FS, wb = rsb - S * bs;
#pragma unroll 2
    for (int n = 0; n < cnt; ++n) {
        float a[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 t = *reinterpret_cast<const float2*>(pa + 2 * q);
            a[2 * q] = t.x; a[2 * q + 1] = t.y;
        }
        float2 b0, b1;
        if (BVEC) {
            b0 = *reinterpret_cast<const float2*>(pb);
            b1 = *reinterpret_cast<const float2*>(pb + 2);
        } else {
            b0 = make_float2(pb[0], pb[1]);
            b1 = make_float2(pb[2], pb[3]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 aa = make_float2(a[i], a[i]);
            c[i][0] = __ffma2_rn(aa, b0, c[i][0]);
            c[i][1] = __ffma2_rn(aa, b1, c[i][1]);
        }
        s += s_step;
        pa += da;
        pb += db;
        if (s >= S) { s -= S; pa += wa; pb += wb; }
    }
}

// One 16-row block for a warp whose two units own up to NPW station pairs each.
//   ra: AX slab          rb: X block -> G1 slab -> AG slab -> dAG slab (-> next X block)
//   rc: dU block -> dZ2 in place -> dZ1 slab
template <int NPW>
__device__ __forceinline__ void gb2_block(float* __restrict__ ra, float* __restrict__ rb, float* __restrict__ rc,
                                          const float* __restrict__ adjT, const float* __restrict__ adjN, int astride,
                                          const float* __restrict__ w1d, const float* __restrict__ b1s,
                                          const float* __restrict__ w2d, const float* __restrict__ b2s,
                                          const float* __restrict__ w2t, int S, int ldu, int s0, int s_end, int r,
                                          int nrows, Gb2Tiles& tl, int f0, int fo0, int slice, int nslices, uint64_t* bar_d,
                                          unsigned phase_d, bool du_bulk, const float* __restrict__ next_x,
                                          unsigned next_x_bytes, uint64_t* bar_x) {
    const int in_cols = S * kGrF, RS = gcn_bwd_rows_rs(S);
    float2 acc[NPW][kGrF];
    unsigned m1[NPW];   // bit f: G1[s][f] > 0, bit 16 + f: G1[s+1][f] > 0

    // ---- P1: AX = A.X (own stations) -> ra ; G1 = relu(AX.W1 + b1) -> rb ----
    if ((in_cols & 1) == 0) gr_aggregate<NPW, false, true>(acc, rb + (size_t)r * in_cols, adjT, astride, S);
    else gr_aggregate<NPW, false, false>(acc, rb + (size_t)r * in_cols, adjT, astride, S);
    __syncthreads();   // every read of the X block is done: rb becomes the G1 slab
    gb2_store_acc<NPW>(ra + (size_t)r * RS, acc, s0, s_end, 1.0f);
    {
        float* g1row = rb + (size_t)r * RS;
#pragma unroll
        for (int p = 0; p < NPW; ++p) m1[p] = 0u;
        float2 o[NPW][4];
        auto put = [&](int c0, auto width) {
            constexpr int W = decltype(width)::value;
            const float4 bb = *reinterpret_cast<const float4*>(b1s + c0);
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int p = 0; p < NPW; ++p) {
                const int s = s0 + 2 * p;
                float v0[4], v1[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    v0[c] = c < W ? gr_relu(o[p][c].x + bv[c]) : 0.0f;
                    v1[c] = c < W ? gr_relu(o[p][c].y + bv[c]) : 0.0f;
                    if (c < W) m1[p] |= (v0[c] > 0.0f ? 1u : 0u) << (c0 + c) | (v1[c] > 0.0f ? 1u : 0u) << (16 + c0 + c);
                }
                if (s < s_end) {
                    *reinterpret_cast<float2*>(g1row + s * kGb2FS + c0) = make_float2(v0[0], v0[1]);
                    if (W == 4) *reinterpret_cast<float2*>(g1row + s * kGb2FS + c0 + 2) = make_float2(v0[2], v0[3]);
                }
                if (s + 1 < s_end) {
                    *reinterpret_cast<float2*>(g1row + (s + 1) * kGb2FS + c0) = make_float2(v1[0], v1[1]);
                    if (W == 4) *reinterpret_cast<float2*>(g1row + (s + 1) * kGb2FS + c0 + 2) = make_float2(v1[2], v1[3]);
                }
            }
        };
#pragma unroll 1
        for (int c0 = 0; c0 < 12; c0 += 4) {
            gr_transform_chunk<NPW, 4>(o, acc, w1d, c0);
            put(c0, IntC<4>{});
        }
        gr_transform_chunk<NPW, 1>(o, acc, w1d, 12);
        put(12, IntC<1>{});
    }
    __syncthreads();   // G1 complete

    // ---- P2: AG = A.G1 -> rb (after everyone has read G1) ; dZ2 = dU * [AG.W2 + b2 > 0] in place in rc ----
    gb2_aggregate<NPW>(acc, rb + (size_t)r * RS, adjT, astride, S);
    __syncthreads();
    gb2_store_acc<NPW>(rb + (size_t)r * RS, acc, s0, s_end, 1.0f);
    if (du_bulk) mbar_wait(bar_d, phase_d);   // the dU block (its bulk copy was issued before P1)
    {
        float* durow = rc + (size_t)r * ldu;
        float2 o[NPW][4];
        auto mask = [&](int c0, auto width) {
            constexpr int W = decltype(width)::value;
            const float4 bb = *reinterpret_cast<const float4*>(b2s + c0);
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int p = 0; p < NPW; ++p) {
                const int s = s0 + 2 * p;
#pragma unroll
                for (int c = 0; c < W; ++c) {
                    if (s < s_end && !(o[p][c].x + bv[c] > 0.0f)) durow[s * kGrF + c0 + c] = 0.0f;
                    if (s + 1 < s_end && !(o[p][c].y + bv[c] > 0.0f)) durow[(s + 1) * kGrF + c0 + c] = 0.0f;
                }
            }
        };
#pragma unroll 1
        for (int c0 = 0; c0 < 12; c0 += 4) {
            gr_transform_chunk<NPW, 4>(o, acc, w2d, c0);
            mask(c0, IntC<4>{});
        }
        gr_transform_chunk<NPW, 1>(o, acc, w2d, 12);
        mask(12, IntC<1>{});
    }
    __syncthreads();   // AG and dZ2 complete

    // ---- P3: dW2 += AG^T dZ2 (row 13: db2) ; then dAG = dZ2 . W2^T -> rb ----
    if (slice >= 0) gb2_reduce<false>(tl.w2, rb, RS, rc, ldu, kGrF, nrows, S, f0, fo0, slice, nslices);
    {   // own dZ2 as operand pairs (reads only)
        const float* dz = rc + (size_t)r * ldu;
#pragma unroll
        for (int p = 0; p < NPW; ++p) {
            const int s = s0 + 2 * p;
#pragma unroll
            for (int f = 0; f < kGrF; ++f)
                acc[p][f] = make_float2(s < s_end ? dz[s * kGrF + f] : 0.0f, s + 1 < s_end ? dz[(s + 1) * kGrF + f] : 0.0f);
        }
    }
    __syncthreads();   // the reduction has read every AG: rb becomes the dAG slab
    {
        float* dagrow = rb + (size_t)r * RS;
        float2 o[NPW][4];
        auto put = [&](int c0, auto width) {
            constexpr int W = decltype(width)::value;
#pragma unroll
            for (int p = 0; p < NPW; ++p) {
                const int s = s0 + 2 * p;
                if (s < s_end) {
                    *reinterpret_cast<float2*>(dagrow + s * kGb2FS + c0) = make_float2(o[p][0].x, W == 4 ? o[p][1].x : 0.0f);
                    if (W == 4) *reinterpret_cast<float2*>(dagrow + s * kGb2FS + c0 + 2) = make_float2(o[p][2].x, o[p][3].x);
                }
                if (s + 1 < s_end) {
                    *reinterpret_cast<float2*>(dagrow + (s + 1) * kGb2FS + c0) = make_float2(o[p][0].y, W == 4 ? o[p][1].y : 0.0f);
                    if (W == 4) *reinterpret_cast<float2*>(dagrow + (s + 1) * kGb2FS + c0 + 2) = make_float2(o[p][2].y, o[p][3].y);
                }
            }
        };
#pragma unroll 1
        for (int c0 = 0; c0 < 12; c0 += 4) {
            gr_transform_chunk<NPW, 4>(o, acc, w2t, c0);
            put(c0, IntC<4>{});
        }
        gr_transform_chunk<NPW, 1>(o, acc, w2t, 12);
        put(12, IntC<1>{});
    }
    __syncthreads();   // dAG complete

    // ---- P4: dG1 = A^T dAG ; dZ1 = dG1 * [G1 > 0] -> rc (dZ2 is dead) ----
    gb2_aggregate<NPW>(acc, rb + (size_t)r * RS, adjN, astride, S);
#pragma unroll
    for (int p = 0; p < NPW; ++p)
#pragma unroll
        for (int f = 0; f < kGrF; ++f) {
            if (!((m1[p] >> f) & 1u)) acc[p][f].x = 0.0f;
            if (!((m1[p] >> (16 + f)) & 1u)) acc[p][f].y = 0.0f;
        }
    gb2_store_acc<NPW>(rc + (size_t)r * RS, acc, s0, s_end, 0.0f);
    __syncthreads();   // dZ1 complete; every read of dAG is done: rb is free
    if (next_x_bytes != 0 && threadIdx.x == 0) {   // the next block's X lands during the last reduction
        mbar_expect_tx(bar_x, next_x_bytes);
        bulk_g2s(rb, next_x, next_x_bytes, bar_x);
    }

    // ---- P5: dW1 += AX^T dZ1 (row 13: db1) ----
    if (slice >= 0) gb2_reduce<true>(tl.w1, ra, RS, rc, RS, kGb2FS, nrows, S, f0, fo0, slice, nslices);
    __syncthreads();   // rc and ra are free for the next block
}

// part: [gridDim.x * 16][2 * 256 + 32] per-slice partials (the layout gcn_bwd_finish_kernel adds up).
__global__ void __launch_bounds__(kGb2Threads, 2)
    gcn_bwd_rows_kernel(const float* __restrict__ X, const float* __restrict__ dU, const float* __restrict__ adj,
                        const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                        const float* __restrict__ b2, float* __restrict__ part, long long R, int S, int ldu) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int in_cols = S * kGrF;
    // station pairs are dealt out to the 8 units as evenly as possible; the warps rotate with the CTA's wave so
    // that the two co-resident CTAs do not put their heavier units on one scheduler
    const int NPT = ceil_div(S, 2), base = NPT / kGb2Units, extra = NPT % kGb2Units;
    const int w = ((tid >> 5) + (blockIdx.x / kNumSMs)) % kGb2Warps;
    const int u = 2 * w + (lane >> 4);
    const int npu = base + (u < extra ? 1 : 0);
    const int s0 = 2 * (u * base + (u < extra ? u : extra));
    const int s_end = (s0 + 2 * npu) < S ? (s0 + 2 * npu) : S;
    const int npw = base + (2 * w < extra ? 1 : 0);   // the warp's wider unit (the even one)
    const int r = lane & 15;

    const int REG = gcn_bwd_rows_region(S, ldu);
    float* adjT = smem;                                          // [S][8 units][8]: A[s0(u) + i][s']
    float* adjN = adjT + (size_t)S * kGb2Units * kGb2AP;         //                  A[s'][s0(u) + i]
    float* w1d = adjN + (size_t)S * kGb2Units * kGb2AP;
    float* w2d = w1d + kGcnFS * kGcnFS;
    float* w2t = w2d + kGcnFS * kGcnFS;
    float* b1s = w2t + kGcnFS * kGcnFS;
    float* b2s = b1s + kGcnFS;
    float* ra = b2s + kGcnFS;
    float* rb = ra + REG;
    float* rc = rb + REG;
    uint64_t* bar_x = reinterpret_cast<uint64_t*>(rc + REG);
    uint64_t* bar_d = bar_x + 1;

    for (int e = tid; e < S * kGb2Units * kGb2AP; e += kGb2Threads) {
        const int sp = e / (kGb2Units * kGb2AP), c = e % (kGb2Units * kGb2AP);
        const int uu = c / kGb2AP, i = c % kGb2AP;
        const int un = base + (uu < extra ? 1 : 0), us0 = 2 * (uu * base + (uu < extra ? uu : extra));
        const int s = us0 + i;
        const bool ok = i < 2 * un && s < S;
        adjT[e] = ok ? adj[(size_t)s * S + sp] : 0.0f;
        adjN[e] = ok ? adj[(size_t)sp * S + s] : 0.0f;
    }
    for (int e = tid; e < kGcnFS * kGcnFS; e += kGb2Threads) {
        const int f = e / kGcnFS, fo = e % kGcnFS;
        const bool in = f < kGrF && fo < kGrF;
        w1d[e] = in ? W1[f * kGrF + fo] : 0.0f;
        w2d[e] = in ? W2[f * kGrF + fo] : 0.0f;
        w2t[e] = in ? W2[fo * kGrF + f] : 0.0f;   // w2t[fo'][f'] = W2[f'][fo']
    }
    for (int e = tid; e < kGcnFS; e += kGb2Threads) {
        b1s[e] = e < kGrF ? b1[e] : 0.0f;
        b2s[e] = e < kGrF ? b2[e] : 0.0f;
    }
    if (tid == 0) {
        mbar_init(bar_x, 1);
        mbar_init(bar_d, 1);
        fence_mbar_init();
    }
    __syncthreads();

    // reduction coordinates: 8 (f) x 4 (fo) output tile, slice of the (row, station) index.  The warps whose units
    // own one more station pair than the others (S = 34: one warp with 3 + 2 pairs against 2 + 2) leave the
    // reductions to the lighter warps — that evens out the FMA work between the CTA's barriers.
    const int heavy_warps = (extra + 1) / 2;
    const bool all_reduce = heavy_warps == 0 || heavy_warps == kGb2Warps;
    const int nslices = (all_reduce ? kGb2Warps : kGb2Warps - heavy_warps) * 4;
    const int f0 = (lane & 1) * 8, fo0 = ((lane >> 1) & 3) * 4;
    const int slice = all_reduce ? w * 4 + (lane >> 3) : (w >= heavy_warps ? (w - heavy_warps) * 4 + (lane >> 3) : -1);
    const int part_slot = tid >> 3;   // where this thread's partial sums go (0 .. 15)
    Gb2Tiles tl;
#pragma unroll
    for (int i = 0; i < 8; ++i) tl.w1[i][0] = tl.w1[i][1] = tl.w2[i][0] = tl.w2[i][1] = make_float2(0.0f, 0.0f);

    const float* arT = adjT + u * kGb2AP;
    const float* arN = adjN + u * kGb2AP;
    const int astride = kGb2Units * kGb2AP;
    const long long nblocks = (R + kGb2Rows - 1) / kGb2Rows;
    auto x_span = [&](long long blk, const float*& src, unsigned& bytes, int& nr) {
        const long long r0 = blk * kGb2Rows;
        nr = (int)((R - r0) < kGb2Rows ? (R - r0) : kGb2Rows);
        src = X + (size_t)r0 * in_cols;
        bytes = (unsigned)((size_t)nr * in_cols * 4);
        return ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0);
    };
    unsigned phase_x = 0, phase_d = 0;
    bool prefetched = false;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const float* xsrc;
        unsigned xbytes;
        int nrows;
        const bool x_bulk = x_span(blk, xsrc, xbytes, nrows);
        const float* dsrc = dU + (size_t)blk * kGb2Rows * ldu;
        const unsigned dbytes = (unsigned)((size_t)nrows * ldu * 4);
        const bool d_bulk = ((reinterpret_cast<uintptr_t>(dsrc) & 15) == 0) && ((dbytes & 15) == 0);
        if (tid == 0) {
            if (x_bulk && !prefetched) {
                mbar_expect_tx(bar_x, xbytes);
                bulk_g2s(rb, xsrc, xbytes, bar_x);
            }
            if (d_bulk) {   // rc was released by the previous block's last barrier
                mbar_expect_tx(bar_d, dbytes);
                bulk_g2s(rc, dsrc, dbytes, bar_d);
            }
        }
        if (!x_bulk)
            for (int e = tid; e < nrows * in_cols; e += kGb2Threads) rb[e] = __ldg(xsrc + e);
        if (!d_bulk)
            for (int e = tid; e < nrows * ldu; e += kGb2Threads) rc[e] = __ldg(dsrc + e);
        if (!x_bulk || !d_bulk) __syncthreads();
        if (x_bulk) {
            mbar_wait(bar_x, phase_x);
            phase_x ^= 1;
        }
        const float* nsrc = nullptr;
        unsigned nbytes = 0;
        int nn;
        prefetched = false;
        if (blk + gridDim.x < nblocks && x_span(blk + gridDim.x, nsrc, nbytes, nn)) prefetched = true;
        else nbytes = 0;
#define WG_GB2(N)                                                                                                   \
    case N:                                                                                                         \
        gb2_block<N>(ra, rb, rc, arT, arN, astride, w1d, b1s, w2d, b2s, w2t, S, ldu, s0, s_end, r, nrows, tl, f0,   \
                     fo0, slice, nslices, bar_d, phase_d, d_bulk, nsrc, nbytes, bar_x);                             \
        break;
        switch (npw) {   // warp-uniform
            WG_GB2(1) WG_GB2(2) WG_GB2(3)
            default:   // a warp without stations (S < 16) still takes part in the barriers and the reductions
                gb2_block<1>(ra, rb, rc, arT, arN, astride, w1d, b1s, w2d, b2s, w2t, S, ldu, s0, s_end, r, nrows, tl, f0,
                             fo0, slice, nslices, bar_d, phase_d, d_bulk, nsrc, nbytes, bar_x);
                break;
        }
#undef WG_GB2
        if (d_bulk) phase_d ^= 1;
    }

    // ---- per-thread partials: part[cta * 16 + slice][...]; row 13 of each product is the bias gradient ----
    float* pp = part + ((size_t)blockIdx.x * kGb2Slices + part_slot) * kGb2PartFloats;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int f = f0 + i;
        const float4 v1 = make_float4(tl.w1[i][0].x, tl.w1[i][0].y, tl.w1[i][1].x, tl.w1[i][1].y);
        const float4 v2 = make_float4(tl.w2[i][0].x, tl.w2[i][0].y, tl.w2[i][1].x, tl.w2[i][1].y);
        if (f < kGrF) {
            *reinterpret_cast<float4*>(pp + f * 16 + fo0) = v1;
            *reinterpret_cast<float4*>(pp + 256 + f * 16 + fo0) = v2;
        } else if (f == kGrF) {
            *reinterpret_cast<float4*>(pp + 512 + fo0) = v1;
            *reinterpret_cast<float4*>(pp + 528 + fo0) = v2;
        }
    }
}

}  // namespace wg
