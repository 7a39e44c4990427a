// Backward of the GRU recurrence (BPTT) — all T steps, last to first, inside ONE launch.
//
// Reference: `loss.backward()` (src/main.py:76) through `self.gru(hidden2)`
// (src/step6_gcn_gru_combined_model.py:23).  With the forward's saved gate values
// (r, z, n, hn = W_hn h + b_hn) and h_prev = out[t-1] (0 at t = 0):
//     g     = dL/dout[t] + dL/dh_t carried from step t+1
//     dn    = g (1 - z)            dz = g (h_prev - n)
//     da_n  = dn (1 - n^2)         da_z = dz z (1 - z)        da_r = da_n hn r (1 - r)
//     dgi   = [da_r, da_z, da_n]   dgh  = [da_r, da_z, da_n r]
//     dL/dh_{t-1} = g z + dgh . W_hh
// The kernel writes DG[b, t, :] = [da_r | da_z | da_n | da_n r] (4H columns) for the three weight
// / input GEMMs that follow (sgemm.cuh), and per-CTA column sums of DG for the bias gradients.
//
// A CTA owns 16 sequences; W_hh ([3H][H] as PyTorch stores it — already "contraction-major" for
// dgh . W_hh) stays in shared memory for the whole kernel.  Per step: the gate phase (thread = hidden
// unit x half of the sequences, lanes along the hidden index so every global access is coalesced;
// its six input rows per sequence were prefetched with cp.async during the previous step's GEMM),
// then the [16 x 3H] . [3H x H] product on the FMA pipe with FFMA2, the contraction split in two
// halves over 2 x 104 threads (thread tile 4 sequences x 4 hidden units).
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kGbBT = 16;        // sequences per CTA (large batches); 4 when the batch would not fill the SMs
constexpr int kGbThreads = 256;
// contraction splits: (BT / 4) row groups x 26 column groups x KS splits = 208 threads for H = 102
__host__ __device__ constexpr int gru_bwd_ksplit(int BT) { return 32 / BT; }
__host__ __device__ inline int gru_bwd_gr(int H, int BT) { return round_up(3 * H, 4 * gru_bwd_ksplit(BT)); }

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, bool pred) {
    const int src_bytes = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async8z(void* smem_dst, const void* gmem_src, bool pred) {
    const int src_bytes = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(src_bytes)
                 : "memory");
}

// HP = H rounded up to 4 (row stride of the per-sequence vectors), GR = 3H rounded up so that each of
// the KS contraction splits is a multiple of 4 long.
__host__ __device__ inline size_t gru_bwd_smem_floats(int HP, int GR, int BT = kGbBT) {
    size_t n = 0;
    n += (size_t)GR * HP;                              // W_hh rows (zero padded)
    n += (size_t)BT * (GR + 4);                        // dgh of the current step
    n += (size_t)(1 + gru_bwd_ksplit(BT)) * BT * HP;   // D (g z) and the KS partial products of dgh . W_hh
    n += 6 * (size_t)BT * HP;                          // staged inputs of one step: dout, r, z, n, hn, h_prev
    return n;
}

template <int BT>
__global__ void __launch_bounds__(kGbThreads, 1)
    gru_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ out, const float* __restrict__ dout,
                   const float* __restrict__ w_hh, float* __restrict__ DG, float* __restrict__ bias_part,
                   long long B, int T, int H, int LD4, int HP, int GR) {
    extern __shared__ __align__(16) float smem[];
    const int DS = GR + 4;  // dgh row stride: rows 4 apart land in different banks
    float* Ws = smem;                         // [GR][HP]
    constexpr int KS = gru_bwd_ksplit(BT);
    float* dgh = Ws + (size_t)GR * HP;        // [BT][DS]
    float* Dd = dgh + BT * DS;                // [BT][HP]
    float* Pp = Dd + BT * HP;                 // [KS][BT][HP]
    float* stg = Pp + KS * BT * HP;           // [6][BT][HP]

    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * BT;
    const int G = 3 * H;

    for (int e = tid; e < GR * HP; e += kGbThreads) {
        const int n = e / HP, k = e - n * HP;
        Ws[e] = (n < G && k < H) ? __ldg(w_hh + (size_t)n * H + k) : 0.0f;
    }
    for (int e = tid; e < BT * DS; e += kGbThreads) dgh[e] = 0.0f;
    for (int e = tid; e < (1 + KS) * BT * HP; e += kGbThreads) Dd[e] = 0.0f;      // D and the partials
    for (int e = tid; e < 6 * BT * HP; e += kGbThreads) stg[e] = 0.0f;

    // ---- staging of one step's inputs: six H-long rows per sequence ----
    const bool even = (H & 1) == 0 && (LD4 & 1) == 0 &&
                      ((reinterpret_cast<uintptr_t>(gates) | reinterpret_cast<uintptr_t>(out) |
                        reinterpret_cast<uintptr_t>(dout)) & 7) == 0;
    // a warp takes (input, sequence) rows w, w + 8, ...; its lanes walk the row (no divisions)
    auto prefetch = [&](int t) {
        const int per_row = even ? (H >> 1) : H;       // copies per row
        for (int rb = tid >> 5; rb < 6 * BT; rb += kGbThreads / 32) {
            const int b = rb % BT, q = rb / BT;
            const bool ok = (b0 + b < B) && !(q == 5 && t == 0);
            const size_t row = (size_t)(b0 + b) * T + t;
            const float* src = dout;                         // any valid address when !ok (zero fill)
            if (ok) {
                if (q == 0) src = dout + row * H;
                else if (q == 5) src = out + (row - 1) * H;  // h_prev = out[b, t - 1]
                else src = gates + row * LD4 + (size_t)(q - 1) * H;
            }
            float* dst = stg + ((size_t)q * BT + b) * HP;
            if (even) {
                for (int c = tid & 31; c < per_row; c += 32) cp_async8z(dst + 2 * c, ok ? src + 2 * c : src, ok);
            } else {
                for (int c = tid & 31; c < per_row; c += 32) cp_async4(dst + c, ok ? src + c : src, ok);
            }
        }
        cp_async_commit();
    };

    __syncthreads();
    prefetch(T - 1);

    // gate-phase coordinates: hidden unit j (+128 ...) x half of the sequences
    const int gj = tid & 127;
    const int gh_ = tid >> 7;              // 0 / 1: first / second half of the CTA's sequences
    // GEMM coordinates: contraction split kh, row group rg (4 sequences), column group cg (4 units)
    const int n_cg = HP >> 2;
    const int per_half = (BT / 4) * n_cg;
    const int kh = tid / per_half;         // >= KS: idle in the GEMM
    const int rem = tid - kh * per_half;
    const int rg = rem / n_cg, cg = rem - rg * n_cg;
    const int KH = GR / KS;

    // per-thread column sums of DG over this CTA's sequences and all steps (hidden units gj, gj+128, ...)
    constexpr int kMaxJ = 1;   // H <= 128 (the launcher enforces HP <= 128)
    float bsum[kMaxJ][4];
#pragma unroll
    for (int u = 0; u < kMaxJ; ++u) bsum[u][0] = bsum[u][1] = bsum[u][2] = bsum[u][3] = 0.0f;

    for (int t = T - 1; t >= 0; --t) {
        cp_async_wait<0>();
        __syncthreads();   // staged inputs of step t and the previous GEMM's P0 / P1 are visible

        // ================= gate phase =================
#pragma unroll
        for (int u = 0; u < kMaxJ; ++u) {
            const int j = gj + u * 128;
            if (j < H) {
#pragma unroll
                for (int bl = 0; bl < BT / 2; ++bl) {
                    const int b = gh_ * (BT / 2) + bl;
                    const int o = b * HP + j;
                    float g = stg[o] + Dd[o];
#pragma unroll
                    for (int p = 0; p < KS; ++p) g += Pp[p * BT * HP + o];
                    const float r = stg[1 * BT * HP + o], z = stg[2 * BT * HP + o];
                    const float n = stg[3 * BT * HP + o], hn = stg[4 * BT * HP + o];
                    const float hp = stg[5 * BT * HP + o];
                    const float dn = g * (1.0f - z);
                    const float dz = g * (hp - n);
                    const float da_n = dn * (1.0f - n * n);
                    const float da_z = dz * z * (1.0f - z);
                    const float da_r = da_n * hn * r * (1.0f - r);
                    const float da_nr = da_n * r;
                    Dd[o] = g * z;
                    float* d = dgh + b * DS;
                    d[j] = da_r;
                    d[H + j] = da_z;
                    d[2 * H + j] = da_nr;
                    if (b0 + b < B) {
                        float* dg = DG + ((size_t)(b0 + b) * T + t) * LD4;
                        dg[j] = da_r;
                        dg[H + j] = da_z;
                        dg[2 * H + j] = da_n;
                        dg[3 * H + j] = da_nr;
                    }
                    bsum[u][0] += da_r; bsum[u][1] += da_z; bsum[u][2] += da_n; bsum[u][3] += da_nr;
                }
            }
        }
        __syncthreads();   // dgh complete; the staged inputs are free
        if (t == 0) break;
        prefetch(t - 1);   // lands during the GEMM

        // ================= P = dgh . W_hh (KS contraction splits) =================
        if (kh < KS) {
            float2 acc[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.0f, 0.0f);
            const float* ap = dgh + (rg * 4) * DS + kh * KH;
            const float* wp = Ws + (size_t)(kh * KH) * HP + cg * 4;
#pragma unroll 2
            for (int n = 0; n < KH; n += 4) {
                float4 a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(ap + i * DS + n);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(wp + (size_t)(n + kk) * HP);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                        const float2 aa = make_float2(av, av);
                        acc[i][0] = __ffma2_rn(aa, make_float2(w[kk].x, w[kk].y), acc[i][0]);
                        acc[i][1] = __ffma2_rn(aa, make_float2(w[kk].z, w[kk].w), acc[i][1]);
                    }
                }
            }
            float* P = Pp + kh * BT * HP;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(P + (rg * 4 + i) * HP + cg * 4) =
                    make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
        }
    }

    // ---- per-CTA column sums of DG: bias_part[(cta * 2 + half)][LD4] ----
    float* bp = bias_part + ((size_t)blockIdx.x * 2 + gh_) * LD4;
#pragma unroll
    for (int u = 0; u < kMaxJ; ++u) {
        const int j = gj + u * 128;
        if (j < H) {
            bp[j] = bsum[u][0];
            bp[H + j] = bsum[u][1];
            bp[2 * H + j] = bsum[u][2];
            bp[3 * H + j] = bsum[u][3];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BPTT with W_hh IN REGISTERS (compile-time padded hidden size HPT).
//
// gru_bwd_kernel keeps W_hh (125 KB at H = 102) in shared memory: one CTA per SM, 16 sequences per
// CTA (256 CTAs = 1.7 waves at 4096 sequences), 8.2 us per timestep — 36 % of the FMA pipe (ncu).
// Here thread (gate g, k-half kh, hidden-unit PAIR p) keeps W_hh[g*H + k][2p, 2p+1] for its HPT/2
// contraction indices k in registers for the whole kernel (104 registers at H = 102) and computes, for
// every sequence of the CTA, the partial product over ITS gate and k-half:
//     P[g][kh][b][2p, 2p+1] = sum_{k in half kh} dgh[b][g*H + k] * W_hh[g*H + k][2p, 2p+1]
// as FFMA2 with the unit pair as the vector operand and dgh[b][k] as the broadcast scalar.  dgh lives in
// shared memory as [gate][k][sequence], so one broadcast 16-byte load (four sequences at one k) feeds FOUR
// FFMA2.  (First version: thread = one column, FFMA2 over sequence pairs — two FFMA2 per 16-byte load, and a
// broadcast LDS.128 still costs two shared-memory wavefronts: ncu showed 19 K wavefronts and 30 K cycles per
// timestep, 41 % of the warp samples waiting on shared-memory data.)  The six partials of a unit meet in the
// next step's gate phase, added in a fixed order (deterministic).  Without the weights shared memory holds
// 28 sequences: 4096 sequences are ONE wave of 147 CTAs.  Sequences are processed in passes of 16 so that
// the accumulators stay at 32 registers.  Gate phase, staging (cp.async during the product), DG / bias-
// partial outputs as in gru_bwd_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kGrwThreads = 320;   // 3 gates x 2 k-halves x 52 unit pairs = 312 product threads
// dgh row stride (floats): a multiple of 4 with an odd number of 16-byte groups, so that the gate phase's
// 16-byte stores (lanes = consecutive hidden units) are conflict free
__host__ __device__ constexpr int gru_bwd_regw_btp(int BT) { return ((BT / 4) % 2 == 1) ? BT : BT + 4; }
// staging slots (floats): a row is fetched by ONE bulk copy from the 16-byte-aligned address at or below its start,
// so element j of the row sits at slot[m + j], m = the row's misalignment in floats (0 .. 3)
__host__ __device__ constexpr int gru_bwd_regw_sg(int HPT) { return 4 * HPT + 4; }   // 4H floats + m, rounded up to 4
__host__ __device__ constexpr int gru_bwd_regw_sr(int HPT) { return HPT + 4; }       // H floats + m, rounded up to 4
__host__ __device__ constexpr int gru_bwd_regw_nbuf(int BT) { return BT <= 8 ? 2 : 1; }   // staging buffers
__host__ __device__ inline size_t gru_bwd_regw_smem_floats(int HPT, int BT) {
    return 3 * (size_t)HPT * gru_bwd_regw_btp(BT)      // dgh [3][HPT][BTP]
           + (size_t)BT * HPT                          // D = g z
           + 6 * (size_t)BT * HPT                      // P: the six (gate, k-half) partials of dgh . W_hh
           + (size_t)gru_bwd_regw_nbuf(BT) * BT * (gru_bwd_regw_sg(HPT) + 2 * gru_bwd_regw_sr(HPT))   // staged rows:
                                                       // [r|z|n|hn], dout and h_prev of one step, per buffer
           + 4;                                        // mbarriers
}

template <int BT, int HPT>
__global__ void __launch_bounds__(kGrwThreads, 1)
    gru_bwd_regw_kernel(const float* __restrict__ gates, const float* __restrict__ out, const float* __restrict__ dout,
                        const float* __restrict__ w_hh, float* __restrict__ DG, float* __restrict__ bias_part,
                        long long B, int T, int H, int LD4) {
    static_assert(BT % 4 == 0 && HPT % 4 == 0 && 3 * HPT <= kGrwThreads && HPT <= 128, "geometry");
    extern __shared__ __align__(16) float smem[];
    constexpr int HP = HPT, BTP = gru_bwd_regw_btp(BT), NG4 = BT / 4;
    constexpr int NP = HPT / 2;               // unit pairs
    constexpr int KH = HPT / 2;               // contraction indices per k-half
    float* dgh = smem;                        // [3][HPT][BTP]
    float* Dd = dgh + 3 * HPT * BTP;          // [BT][HP]
    float* Pp = Dd + BT * HP;                 // [6][BT][HP]   slot = gate * 2 + k-half
    constexpr int SG = gru_bwd_regw_sg(HPT), SR = gru_bwd_regw_sr(HPT);
    // staging buffers: step t uses buffer t % NBUF and is fetched NBUF steps ahead.  Small CTAs (short products)
    // need the second buffer to cover the global-memory latency; at 16 / 28 sequences the product covers it.
    constexpr int NBUF = gru_bwd_regw_nbuf(BT);
    constexpr int SBUF = BT * (SG + 2 * SR);  // floats per staging buffer
    float* stage0 = Pp + 6 * BT * HP;         // [NBUF] x { [BT][SG] gate rows, [BT][SR] dout rows, [BT][SR] h_prev rows }
    uint64_t* sbar0 = reinterpret_cast<uint64_t*>(stage0 + NBUF * SBUF);   // [NBUF]

    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * BT;

    // product coordinates and this thread's weights
    const int pslot = tid / NP, pp = tid - pslot * NP;
    const int pg = pslot >> 1, pkh = pslot & 1;
    const bool prod_thread = tid < 6 * NP;
    float2 w2[KH];
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        const int k = pkh * KH + kk, j = 2 * pp;
        const float* wr = w_hh + ((size_t)pg * H + k) * H + j;
        w2[kk].x = (prod_thread && k < H && j < H) ? __ldg(wr) : 0.0f;
        w2[kk].y = (prod_thread && k < H && j + 1 < H) ? __ldg(wr + 1) : 0.0f;
    }

    for (int e = tid; e < 3 * HPT * BTP; e += kGrwThreads) dgh[e] = 0.0f;
    for (int e = tid; e < 7 * BT * HP; e += kGrwThreads) Dd[e] = 0.0f;      // D and the partials
    for (int e = tid; e < NBUF * SBUF; e += kGrwThreads) stage0[e] = 0.0f;
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&sbar0[i], 3 * BT);
        fence_mbar_init();
    }

    // ---- staging of one step's inputs: per sequence one gate row (4H floats), one dout row, one h_prev row ----
    // Thread i < 3 BT fetches row (kind = i / BT, sequence i % BT) with ONE bulk copy and arrives on the mbarrier
    // (count 3 BT) with the copy's byte count.  (First version: 8-byte cp.async per thread — as many instructions
    // per step as the whole product.)  A row the aligned copy would take outside its tensor (the tensor's first /
    // last row at most), and h_prev at t = 0, are written by the thread itself.
    const int skind = tid / BT, sb = tid - skind * BT;
    const size_t n_rows = (size_t)B * T;
    auto stage_row = [&](uint64_t* sbar, float* slot, const float* base, size_t total, size_t g0, int n) {
        const int m = (int)((reinterpret_cast<uintptr_t>(base + g0) >> 2) & 3);
        const int L = (m + n + 3) & ~3;
        if (g0 >= (size_t)m && g0 - m + L <= total) {
            mbar_expect_tx(sbar, (unsigned)(L * 4));
            bulk_g2s(slot, base + g0 - m, (unsigned)(L * 4), sbar);
        } else {
            for (int j = 0; j < n; ++j) slot[m + j] = base[g0 + j];
            fence_proxy_async_smem();
            mbar_arrive(sbar);
        }
    };
    auto prefetch = [&](int t) {
        if (tid < 3 * BT && t >= 0) {
            const int buf = t % NBUF;
            uint64_t* sbar = &sbar0[buf];
            float* stG = stage0 + buf * SBUF;
            float* stD = stG + BT * SG;
            float* stH = stD + BT * SR;
            if (b0 + sb >= B) {
                mbar_arrive(sbar);   // no such sequence: its slots stay zero
            } else {
                const size_t row = (size_t)(b0 + sb) * T + t;
                if (skind == 0) {
                    stage_row(sbar, stG + sb * SG, gates, n_rows * LD4, row * LD4, 4 * H);
                } else if (skind == 1) {
                    stage_row(sbar, stD + sb * SR, dout, n_rows * H, row * H, H);
                } else if (t > 0) {
                    stage_row(sbar, stH + sb * SR, out, n_rows * H, (row - 1) * H, H);   // h_prev = out[b, t - 1]
                } else {
                    for (int j = 0; j < SR; ++j) stH[sb * SR + j] = 0.0f;          // h_{-1} = 0
                    fence_proxy_async_smem();
                    mbar_arrive(sbar);
                }
            }
        }
    };
    // base misalignment (floats) of the three staged tensors; a row's own is derived from it in the gate phase
    const int aG = (int)((reinterpret_cast<uintptr_t>(gates) >> 2) & 3);
    const int aD = (int)((reinterpret_cast<uintptr_t>(dout) >> 2) & 3);
    const int aH = (int)((reinterpret_cast<uintptr_t>(out) >> 2) & 3);

    __syncthreads();
    fence_proxy_async_smem();   // the zero fills above are ordered before the first bulk copies
#pragma unroll
    for (int i = 1; i <= NBUF; ++i) prefetch(T - i);

    // gate-phase coordinates: hidden unit gj x every third group of 4 sequences (3 x HPT threads)
    const int gq = tid / HPT;              // 0 .. 2: sequence groups gq, gq + 3, ...; 3: none
    const int gj = tid - gq * HPT;
    const bool gate_thread = gq < 3 && gj < H;
    const long long left = B - b0;
    const int nvalid = left >= BT ? BT : (int)left;       // sequences of this CTA that exist
    const unsigned dg_seq = (unsigned)T * (unsigned)LD4;  // floats between consecutive sequences of DG
    const int tm = T & 3, hm = H & 3;
    const int r00 = (int)((b0 * T) & 3);
    float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};

    for (int t = T - 1; t >= 0; --t) {
        const int buf = t % NBUF;
        const float* stG = stage0 + buf * SBUF;
        const float* stD = stG + BT * SG;
        const float* stH = stD + BT * SR;
        // staged inputs of step t have landed: the buffer's uses are steps t, t + NBUF, ... counted from T - 1 down
        mbar_wait(&sbar0[buf], (unsigned)(((T - 1 - t) / NBUF) & 1));
        __syncthreads();   // the previous product's partials (and thread-written slots) are visible

        // ================= gate phase =================
        if (gate_thread) {
            float* dgt = DG + ((size_t)b0 * T + t) * LD4 + gj;   // this step's DG row of the CTA's first sequence
            const int r0 = (r00 + t) & 3;                        // (row of sequence 0) mod 4
            const float* sgp = stG + aG + gj;                    // gate rows: stride LD4 is a multiple of 4
#pragma unroll
            for (int grp = 0; grp < (NG4 + 2) / 3; ++grp) {
                const int g4 = gq + 3 * grp;
                if (g4 < NG4) {
                    float dr[4], dz_[4], dnr[4];
#pragma unroll
                    for (int bl = 0; bl < 4; ++bl) {
                        const int b = g4 * 4 + bl;
                        const int o = b * HP + gj;
                        const int rl = (r0 + b * tm) & 3;                          // row mod 4
                        const int mD = (aD + rl * hm) & 3;
                        const int mH = t > 0 ? (aH + ((rl + 3) & 3) * hm) & 3 : 0;  // row - 1
                        const float* sg = sgp + b * SG;
                        const float g = ((stD[b * SR + mD + gj] + Dd[o]) + (Pp[o] + Pp[BT * HP + o])) +
                                        ((Pp[2 * BT * HP + o] + Pp[3 * BT * HP + o]) +
                                         (Pp[4 * BT * HP + o] + Pp[5 * BT * HP + o]));
                        const float r = sg[0], z = sg[H];
                        const float n = sg[2 * H], hn = sg[3 * H];
                        const float hp = stH[b * SR + mH + gj];
                        const float dn = g * (1.0f - z);
                        const float dz = g * (hp - n);
                        const float da_n = dn * (1.0f - n * n);
                        const float da_z = dz * z * (1.0f - z);
                        const float da_r = da_n * hn * r * (1.0f - r);
                        const float da_nr = da_n * r;
                        Dd[o] = g * z;
                        dr[bl] = da_r; dz_[bl] = da_z; dnr[bl] = da_nr;
                        if (b < nvalid) {
                            float* dg = dgt + (size_t)b * dg_seq;
                            dg[0] = da_r;
                            dg[H] = da_z;
                            dg[2 * H] = da_n;
                            dg[3 * H] = da_nr;
                        }
                        bsum[0] += da_r; bsum[1] += da_z; bsum[2] += da_n; bsum[3] += da_nr;
                    }
                    float* d = dgh + gj * BTP + g4 * 4;
                    *reinterpret_cast<float4*>(d) = make_float4(dr[0], dr[1], dr[2], dr[3]);
                    *reinterpret_cast<float4*>(d + HPT * BTP) = make_float4(dz_[0], dz_[1], dz_[2], dz_[3]);
                    *reinterpret_cast<float4*>(d + 2 * HPT * BTP) = make_float4(dnr[0], dnr[1], dnr[2], dnr[3]);
                }
            }
        }
        __syncthreads();   // dgh complete; the staged inputs are free
        if (t == 0) break;
        prefetch(t - NBUF);   // into the buffer this step has just finished with

        // ===== P[g][kh] = dgh[g][k-half] . W_hh[g][k-half] (this thread: gate pg, half pkh, unit pair pp, all sequences) =====
        if (prod_thread) {
            const float* dp = dgh + (pg * HPT + pkh * KH) * BTP;
            float* P = Pp + (size_t)pslot * BT * HP + 2 * pp;
#pragma unroll
            for (int q0 = 0; q0 < NG4; q0 += 4) {   // passes of up to 16 sequences
                float2 acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int kk = 0; kk < KH; ++kk) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        if (q0 + qq < NG4) {
                            const float4 d4 = *reinterpret_cast<const float4*>(dp + kk * BTP + 4 * (q0 + qq));
                            acc[4 * qq + 0] = __ffma2_rn(make_float2(d4.x, d4.x), w2[kk], acc[4 * qq + 0]);
                            acc[4 * qq + 1] = __ffma2_rn(make_float2(d4.y, d4.y), w2[kk], acc[4 * qq + 1]);
                            acc[4 * qq + 2] = __ffma2_rn(make_float2(d4.z, d4.z), w2[kk], acc[4 * qq + 2]);
                            acc[4 * qq + 3] = __ffma2_rn(make_float2(d4.w, d4.w), w2[kk], acc[4 * qq + 3]);
                        }
                    }
                }
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    if (q0 + qq < NG4) {
#pragma unroll
                        for (int bl = 0; bl < 4; ++bl)
                            *reinterpret_cast<float2*>(P + (size_t)(4 * (q0 + qq) + bl) * HP) = acc[4 * qq + bl];
                    }
                }
            }
        }
    }

    // ---- per-CTA column sums of DG: bias_part[(cta * 3 + gq)][LD4] ----
    if (gate_thread) {
        float* bp = bias_part + ((size_t)blockIdx.x * 3 + gq) * LD4;
        bp[gj] = bsum[0];
        bp[H + gj] = bsum[1];
        bp[2 * H + gj] = bsum[2];
        bp[3 * H + gj] = bsum[3];
    }
}

// db_ih = [sum da_r | sum da_z | sum da_n], db_hh = [sum da_r | sum da_z | sum da_n r], summed over
// the per-CTA partials in a fixed order.
__global__ void gru_bias_grad_kernel(const float* __restrict__ part, int nparts, int H, int LD4,
                                     float* __restrict__ db_ih, float* __restrict__ db_hh) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 4 * H) return;
    float s = 0.0f;
    for (int z = 0; z < nparts; ++z) s += part[(size_t)z * LD4 + c];
    if (c < 2 * H) {
        db_ih[c] = s;
        db_hh[c] = s;
    } else if (c < 3 * H) {
        db_ih[c] = s;
    } else {
        db_hh[c - H] = s;
    }
}

}  // namespace wg
