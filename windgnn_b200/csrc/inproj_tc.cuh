// GRU input projection on the 5th-generation tensor cores (tcgen05), error-compensated TF32:
//     GI = U . W_ih^T + b   with   U = U_hi + U_lo,  W = W_hi + W_lo  (hi = fp32 rounded to TF32)
//     U.W ~= U_hi.W_hi + U_hi.W_lo + U_lo.W_hi                        ("3xTF32", fp32 accumulate)
// The dropped U_lo.W_lo term and the truncation of the lo parts are O(2^-21) relative, so the
// result is as close to the fp64 truth as the plain fp32 FMA path (2.7e-6 vs
// 3.0e-6 normalised at S = 34) and stays inside the 1e-5 parity bar.
//
// Reference: the `gi = W_ih u + b_ih` half of nn.GRU, src/step6_gcn_gru_combined_model.py:11,23.
//
// Accumulation: TMEM accumulators add with truncation (measured, scripts/probes/tc_gemm_probe.cu:
// the error is a bias towards zero that grows linearly with the number of accumulator updates),
// so the big hi.hi products alternate between TWO accumulators (halving the updates each sees)
// and the small hi.lo / lo.hi products go to a third one whose truncation is 2^-11 smaller; the
// epilogue adds the three in fp32 with round-to-nearest.  Three accumulators of N columns fit the
// 512 TMEM columns for N <= 160, so a CTA tile is 128 rows x one column slice of <= 160.
//
// Structure (one persistent CTA per SM, 6 warps):
//   warp 4 — producer: per pipeline stage (16 k's) four bulk async copies (cp.async.bulk / UBLKCP)
//            land A_hi, A_lo (128 rows) and B_hi, B_lo (the tile's gate columns) in shared memory; both
//            operands are stored by their producers in the UMMA canonical K-major no-swizzle
//            layout ([k/4][row][4 floats]: 8-row x 16-byte core matrices, SBO = 128 B,
//            LBO = rows * 16 B), so a stage is a plain contiguous image;
//   warp 5 — MMA issuer: one thread issues tcgen05.mma.kind::tf32 (M = 128, N <= 256, K = 8) from
//            shared-memory descriptors into a TMEM accumulator (128 lanes x N columns, fp32),
//            tcgen05.commit releases the stage / publishes the finished tile through mbarriers;
//   warps 0-3 — epilogue: tcgen05.ld 32 lanes x 32 columns per warp, + bias, 16-byte stores.
#pragma once

#include "wg_common.cuh"

namespace wg {

constexpr int kTcBM = 128;
constexpr int kTcBK = 16;
constexpr int kTcThreads = 192;

struct TcShape {
    int NP;      // padded gate columns = n_nt * N_each
    int n_nt;    // column slices (CTA tiles along N)
    int N_each;  // columns per slice = MMA N, multiple of 16, <= 160 (three accumulators in 512 TMEM columns)
    int stages;  // pipeline depth that fits shared memory
    size_t stage_bytes;
    size_t smem_bytes;
    bool ok;
};

__host__ inline TcShape tc_shape(int G) {
    TcShape s{};
    const int np16 = round_up(G, 16);
    s.n_nt = ceil_div(np16, 160);
    s.N_each = round_up(ceil_div(np16, s.n_nt), 16);
    s.NP = s.n_nt * s.N_each;
    s.ok = s.N_each <= 160;
    s.stage_bytes = (size_t)2 * (kTcBM + s.N_each) * kTcBK * 4;  // hi + lo of A and B
    s.stages = (int)((kMaxSmemOptin - 1024) / s.stage_bytes);
    if (s.stages > 8) s.stages = 8;
    if (s.stages < 2) s.ok = false;
    s.smem_bytes = s.stage_bytes * s.stages + 1024;
    return s;
}

// ---- tcgen05 / TMEM helpers -------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE: start >> 4 in bits [0,14),
    // leading byte offset >> 4 in [16,30) (next 16-byte chunk along K), stride byte offset >> 4 in
    // [32,46) (next group of 8 rows), descriptor version 1 in [46,48), layout type 0 in [61,64)
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    // instruction descriptor: D = F32 (1 at bit 4), A = B = TF32 (2 at bits 7 and 10), both
    // K-major (bits 15, 16 = 0), N >> 3 at bits [17,23), M >> 4 at bits [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// A_hi / A_lo: [M/128][K/4][128][4]   B_hi / B_lo: [n_nt][K/4][N_each][4]   C: [M][ldc] row-major (GI)
__global__ void __launch_bounds__(kTcThreads, 1)
    inproj_tc_kernel(const float* __restrict__ A_hi, const float* __restrict__ A_lo,
                     const float* __restrict__ B_hi, const float* __restrict__ B_lo,
                     const float* __restrict__ bias, float* __restrict__ C, long long M, int K, int ldc, int n_nt,
                     int N_each, int stages) {
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    unsigned char* const smem_raw = smem_tc;
    const uint32_t a_bytes = kTcBM * kTcBK * 4;              // one of A_hi / A_lo per stage
    const uint32_t b_bytes = (uint32_t)N_each * kTcBK * 4;   // one of B_hi / B_lo per stage
    const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);
    unsigned char* tail = smem_raw + (size_t)stage_bytes * stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);        // [stages]
    uint64_t* empty_bar = full_bar + 8;                             // [stages]
    uint64_t* tmem_full = empty_bar + 8;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m_tiles = (M + kTcBM - 1) / kTcBM;
    const long long n_tiles = m_tiles * n_nt;   // tile = mt * n_nt + nt: the slices of one row tile run together
    const int KB = K / kTcBK;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);  // one arrival per epilogue warp
        fence_mbar_init();
    }
    if (warp == 0) {  // TMEM: the whole 512-column space of this SM (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // accumulators: [0, N) and [N, 2N) take the hi.hi products of even / odd k-steps, [2N, 3N) the
    // hi.lo and lo.hi corrections
    const uint32_t acc_cols = (uint32_t)N_each;

    if (warp == 4) {
        // ===================== producer =====================
        int s = 0;
        uint32_t phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long mt = tile / n_nt;
            const int nt = (int)(tile - mt * n_nt);
            const float* a_hi = A_hi + (size_t)mt * K * kTcBM;
            const float* a_lo = A_lo + (size_t)mt * K * kTcBM;
            const float* b_hi = B_hi + (size_t)nt * K * N_each;
            const float* b_lo = B_lo + (size_t)nt * K * N_each;
            for (int kb = 0; kb < KB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&empty_bar[s], phase ^ 1);  // slot free (first pass: passes immediately)
                    unsigned char* st = smem_raw + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full_bar[s], stage_bytes);
                    bulk_g2s(st, a_hi + (size_t)kb * kTcBK * kTcBM, a_bytes, &full_bar[s]);
                    bulk_g2s(st + a_bytes, a_lo + (size_t)kb * kTcBK * kTcBM, a_bytes, &full_bar[s]);
                    bulk_g2s(st + 2 * a_bytes, b_hi + (size_t)kb * kTcBK * N_each, b_bytes, &full_bar[s]);
                    bulk_g2s(st + 2 * a_bytes + b_bytes, b_lo + (size_t)kb * kTcBK * N_each, b_bytes, &full_bar[s]);
                }
                __syncwarp();
                if (++s == stages) { s = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = umma_idesc_tf32(kTcBM, N_each);
        const uint32_t lbo_a = kTcBM * 16, lbo_b = (uint32_t)N_each * 16, sbo = 128;
        int s = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (lane == 0) {
                mbar_wait(tmem_empty, acc_phase ^ 1);  // epilogue has drained the accumulators
                tc_fence_after();
            }
            __syncwarp();
            for (int kb = 0; kb < KB; ++kb) {
                if (lane == 0) {
                    mbar_wait(&full_bar[s], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem_raw + (size_t)s * stage_bytes);
                    const uint32_t sa_hi = st, sa_lo = st + a_bytes;
                    const uint32_t sb_hi = st + 2 * a_bytes, sb_lo = sb_hi + b_bytes;
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 8; ++ks) {      // K = 8 per MMA: two 16-byte chunks
                        const uint64_t da_hi = umma_desc_kmajor(sa_hi + ks * 2 * lbo_a, lbo_a, sbo);
                        const uint64_t da_lo = umma_desc_kmajor(sa_lo + ks * 2 * lbo_a, lbo_a, sbo);
                        const uint64_t db_hi = umma_desc_kmajor(sb_hi + ks * 2 * lbo_b, lbo_b, sbo);
                        const uint64_t db_lo = umma_desc_kmajor(sb_lo + ks * 2 * lbo_b, lbo_b, sbo);
                        // kTcBK / 8 == 2: k-step parity == ks, so each main accumulator is first
                        // written (not accumulated) in k-block 0
                        umma_tf32(tmem_base + ks * acc_cols, da_hi, db_hi, idesc, kb != 0);        // hi . hi
                        umma_tf32(tmem_base + 2 * acc_cols, da_hi, db_lo, idesc, (kb | ks) != 0);  // hi . lo
                        umma_tf32(tmem_base + 2 * acc_cols, da_lo, db_hi, idesc, 1);               // lo . hi
                    }
                    umma_commit(&empty_bar[s]);                  // stage free once these MMAs have read it
                    if (kb == KB - 1) umma_commit(tmem_full);    // accumulators complete
                }
                __syncwarp();
                if (++s == stages) { s = 0; phase ^= 1; }
            }
            acc_phase ^= 1;
        }
    } else {
        // ===================== epilogue (warps 0-3 <-> TMEM lanes 32w .. 32w+31) =====================
        uint32_t acc_phase = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const long long mt = tile / n_nt;
            const int nt = (int)(tile - mt * n_nt);
            mbar_wait(tmem_full, acc_phase);
            tc_fence_after();
            const long long row = mt * kTcBM + warp * 32 + lane;
            float* crow = C + (size_t)(row < M ? row : 0) * ldc;
            const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < N_each; c0 += 32) {
                float v0[32], v1[32], v2[32];
                tmem_ld32(lane_base + c0, v0);
                tmem_ld32(lane_base + acc_cols + c0, v1);
                tmem_ld32(lane_base + 2 * acc_cols + c0, v2);
                if (row < M) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int cl = c0 + 4 * q;            // column inside the slice
                        const int c = nt * N_each + cl;       // gate column
                        if (cl < N_each && c < ldc) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
                            float4 o;
                            o.x = ((v0[4 * q] + v1[4 * q]) + v2[4 * q]) + b4.x;
                            o.y = ((v0[4 * q + 1] + v1[4 * q + 1]) + v2[4 * q + 1]) + b4.y;
                            o.z = ((v0[4 * q + 2] + v1[4 * q + 2]) + v2[4 * q + 2]) + b4.z;
                            o.w = ((v0[4 * q + 3] + v1[4 * q + 3]) + v2[4 * q + 3]) + b4.w;
                            *reinterpret_cast<float4*>(crow + c) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace wg
