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
// (read) and h (written) touches HBM inside the time loop.  Each step has two phases:
//   GEMM  : gs[b][n] = sum_k hs[b][k] * Ws[k][n].  Warp tile 16 sequences x 32 gate columns,
//           thread tile 4 x 4, all operands by LDS.128 (hs rows padded so the four row
//           offsets of a warp fall in distinct bank groups; Ws rows are read 128 B contiguous).
//   gates : one (sequence, hidden unit) item per thread-slot, lanes along the hidden index so
//           the gi loads / h stores are coalesced; gi for step t+1 is prefetched into registers
//           while step t's gates and step t+1's GEMM run.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kRcBT = 32;    // sequences per CTA
constexpr int kRcMaxQ = 6;   // gate items per thread (BT*H <= kRcMaxQ * threads)

__host__ __device__ inline int recur_hs_stride(int KP) { return ((KP / 4) & 1) ? KP : KP + 4; }
__host__ __device__ inline int recur_gs_stride(int NP) { return NP + 4; }
__host__ __device__ inline size_t recur_smem_floats(int KP, int NP, bool w_smem) {
    size_t n = 0;
    if (w_smem) n += (size_t)KP * NP;
    n += (size_t)kRcBT * recur_hs_stride(KP);
    n += (size_t)kRcBT * recur_gs_stride(NP);
    return n;
}

template <int NWARPS, bool W_SMEM>
__global__ void __launch_bounds__(NWARPS * 32, 1)
    gru_recur_kernel(const float* __restrict__ GI, const float* __restrict__ WhT,
                     const float* __restrict__ bhn, float* __restrict__ out, long long B, int T, int H,
                     int ldg, int KP, int NP) {
    constexpr int NT = NWARPS * 32;
    extern __shared__ __align__(16) float smem[];
    const int RS = recur_hs_stride(KP);
    const int GS = recur_gs_stride(NP);
    float* Ws = smem;
    float* hs = smem + (W_SMEM ? (size_t)KP * NP : 0);
    float* gs = hs + kRcBT * RS;
    const float* Wsrc = W_SMEM ? Ws : WhT;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const long long b0 = (long long)blockIdx.x * kRcBT;

    if (W_SMEM) {
        const int n4 = KP * NP / 4;
        const float4* src = reinterpret_cast<const float4*>(WhT);
        float4* dst = reinterpret_cast<float4*>(Ws);
        for (int e = tid; e < n4; e += NT) dst[e] = __ldg(src + e);
    }
    for (int e = tid; e < kRcBT * RS; e += NT) hs[e] = 0.0f;

    // ---- gate-phase items of this thread: (b, j) packed as b<<16 | j, j fastest over lanes ----
    const int n_items = kRcBT * H;
    int item_bj[kRcMaxQ];
    float bn[kRcMaxQ];
    float gr[kRcMaxQ], gz[kRcMaxQ], gn[kRcMaxQ];
#pragma unroll
    for (int q = 0; q < kRcMaxQ; ++q) {
        const int item = tid + q * NT;
        int b = -1, j = 0;
        if (item < n_items) {
            b = item / H;
            j = item - b * H;
            if (b0 + b >= B) b = -1;  // ragged last CTA
        }
        item_bj[q] = b < 0 ? -1 : ((b << 16) | j);
        bn[q] = b < 0 ? 0.0f : __ldg(bhn + j);
        gr[q] = gz[q] = gn[q] = 0.0f;
    }
    auto load_gi = [&](int t) {
#pragma unroll
        for (int q = 0; q < kRcMaxQ; ++q) {
            if (item_bj[q] >= 0) {
                const int b = item_bj[q] >> 16, j = item_bj[q] & 0xffff;
                const float* p = GI + ((size_t)(b0 + b) * T + t) * ldg + j;
                gr[q] = __ldg(p);
                gz[q] = __ldg(p + H);
                gn[q] = __ldg(p + 2 * H);
            }
        }
    };
    load_gi(0);

    // ---- GEMM-phase coordinates ----
    const int ng = lane & 7;   // column group within the warp tile
    const int bg = lane >> 3;  // row group within the warp tile
    const int n_ntiles = NP / 32;
    const int n_tiles = 2 * n_ntiles;

    __syncthreads();

    for (int t = 0; t < T; ++t) {
        // ================= GEMM phase =================
        for (int wt = warp; wt < n_tiles; wt += NWARPS) {
            const int wr = wt & 1;        // which 16-sequence half
            const int nb = wt >> 1;       // which 32-column block
            const int nbase = nb * 32 + ng * 4;
            const float* hrow = hs + (wr * 16 + bg) * RS;  // rows bg + 4 i
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
            if (t > 0) {  // h_{-1} = 0: the product is zero at t == 0
#pragma unroll 2
                for (int k4 = 0; k4 < KP; k4 += 4) {
                    float4 hv[4], wv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        hv[i] = *reinterpret_cast<const float4*>(hrow + (4 * i) * RS + k4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float* wp = Wsrc + (size_t)(k4 + kk) * NP + nbase;
                        wv[kk] = W_SMEM ? *reinterpret_cast<const float4*>(wp)
                                        : __ldg(reinterpret_cast<const float4*>(wp));
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][0] = fmaf(hv[i].x, wv[0].x, acc[i][0]);
                        acc[i][1] = fmaf(hv[i].x, wv[0].y, acc[i][1]);
                        acc[i][2] = fmaf(hv[i].x, wv[0].z, acc[i][2]);
                        acc[i][3] = fmaf(hv[i].x, wv[0].w, acc[i][3]);
                        acc[i][0] = fmaf(hv[i].y, wv[1].x, acc[i][0]);
                        acc[i][1] = fmaf(hv[i].y, wv[1].y, acc[i][1]);
                        acc[i][2] = fmaf(hv[i].y, wv[1].z, acc[i][2]);
                        acc[i][3] = fmaf(hv[i].y, wv[1].w, acc[i][3]);
                        acc[i][0] = fmaf(hv[i].z, wv[2].x, acc[i][0]);
                        acc[i][1] = fmaf(hv[i].z, wv[2].y, acc[i][1]);
                        acc[i][2] = fmaf(hv[i].z, wv[2].z, acc[i][2]);
                        acc[i][3] = fmaf(hv[i].z, wv[2].w, acc[i][3]);
                        acc[i][0] = fmaf(hv[i].w, wv[3].x, acc[i][0]);
                        acc[i][1] = fmaf(hv[i].w, wv[3].y, acc[i][1]);
                        acc[i][2] = fmaf(hv[i].w, wv[3].z, acc[i][2]);
                        acc[i][3] = fmaf(hv[i].w, wv[3].w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                *reinterpret_cast<float4*>(gs + (wr * 16 + bg + 4 * i) * GS + nbase) = v;
            }
        }
        __syncthreads();

        // ================= gate phase =================
        float cr[kRcMaxQ], cz[kRcMaxQ], cn[kRcMaxQ];
#pragma unroll
        for (int q = 0; q < kRcMaxQ; ++q) {
            cr[q] = gr[q];
            cz[q] = gz[q];
            cn[q] = gn[q];
        }
        if (t + 1 < T) load_gi(t + 1);  // in flight during the gates and the next GEMM phase
#pragma unroll
        for (int q = 0; q < kRcMaxQ; ++q) {
            if (item_bj[q] >= 0) {
                const int b = item_bj[q] >> 16, j = item_bj[q] & 0xffff;
                const float* g = gs + b * GS + j;
                const float r = sigmoid_f(cr[q] + g[0]);
                const float z = sigmoid_f(cz[q] + g[H]);
                const float n = tanh_f(cn[q] + r * (g[2 * H] + bn[q]));
                const float hold = hs[b * RS + j];
                const float hnew = (hold - n) * z + n;
                hs[b * RS + j] = hnew;
                out[((size_t)(b0 + b) * T + t) * H + j] = hnew;
            }
        }
        __syncthreads();
    }
}

}  // namespace wg
