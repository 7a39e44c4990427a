// Accuracy probe of the tcgen05 TF32 GEMM kernel: how does the TMEM fp32 accumulation round?
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../windgnn_b200/csrc/inproj_tc.cuh"

static float tf32r(float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

int main() {
    const int M = 256, NP = 320, N_each = 160, n_nt = 2;
    for (int K : {64, 448, 1792}) {
        for (int mode = 0; mode < 3; ++mode) {  // 0: tf32-exact inputs, lo = 0 ; 1: full fp32 inputs, 3xTF32 ; 2: positive inputs 3xTF32
            std::vector<float> A((size_t)M * K), B((size_t)NP * K);
            srand(1);
            for (auto& v : A) { v = (float)rand() / RAND_MAX * (mode == 2 ? 1.f : 2.f) - (mode == 2 ? 0.f : 1.f); if (mode == 0) v = tf32r(v); }
            for (auto& v : B) { v = (float)rand() / RAND_MAX * (mode == 2 ? 1.f : 2.f) - (mode == 2 ? 0.f : 1.f); if (mode == 0) v = tf32r(v); }
            // layouts: A [M/128][K/4][128][4], B [K/4][NP][4]
            std::vector<float> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
            for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
                size_t o = (((size_t)(m / 128) * (K / 4) + k / 4) * 128 + m % 128) * 4 + k % 4;
                float v = A[(size_t)m * K + k], h = tf32r(v); Ah[o] = h; Al[o] = v - h;
            }
            for (int n = 0; n < NP; ++n) for (int k = 0; k < K; ++k) {
                size_t o = (((size_t)(n / N_each) * (K / 4) + k / 4) * N_each + n % N_each) * 4 + k % 4;
                float v = B[(size_t)n * K + k], h = tf32r(v); Bh[o] = h; Bl[o] = v - h;
            }
            float *dAh, *dAl, *dBh, *dBl, *dC, *dbias;
            cudaMalloc(&dAh, A.size() * 4); cudaMalloc(&dAl, A.size() * 4); cudaMalloc(&dBh, B.size() * 4); cudaMalloc(&dBl, B.size() * 4);
            cudaMalloc(&dC, (size_t)M * NP * 4); cudaMalloc(&dbias, 512 * 4); cudaMemset(dbias, 0, 512 * 4);
            cudaMemcpy(dAh, Ah.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dAl, Al.data(), A.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(dBh, Bh.data(), B.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBl, Bl.data(), B.size() * 4, cudaMemcpyHostToDevice);
            wg::TcShape t = wg::tc_shape(306);
            cudaFuncSetAttribute(wg::inproj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem_bytes);
            wg::inproj_tc_kernel<<<4, wg::kTcThreads, t.smem_bytes>>>(dAh, dAl, dBh, dBl, dbias, dC, M, K, NP, n_nt, N_each, t.stages);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> C((size_t)M * NP);
            cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
            double maxref = 0, maxerr = 0, sumerr = 0, maxerr32 = 0, bias = 0;
            for (int m = 0; m < M; ++m) for (int n = 0; n < NP; ++n) {
                double r = 0; float f = 0;
                for (int k = 0; k < K; ++k) { r += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k]; f = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f); }
                maxref = fmax(maxref, fabs(r));
                double d = C[(size_t)m * NP + n] - r;
                maxerr = fmax(maxerr, fabs(d)); sumerr += fabs(d); bias += d * (r >= 0 ? 1 : -1);
                maxerr32 = fmax(maxerr32, fabs((double)f - r));
            }
            printf("K=%4d mode=%d (%s): max|err|/max|ref| = %.2e  mean|err|/max = %.2e  signed bias/max = %+.2e | fp32 fma chain: %.2e   [%s]\n", K, mode,
                   mode == 0 ? "tf32-exact, lo=0" : mode == 1 ? "fp32 inputs, 3xTF32" : "positive fp32, 3xTF32", maxerr / maxref, sumerr / C.size() / maxref,
                   bias / C.size() / maxref, maxerr32 / maxref, cudaGetErrorString(e));
            cudaFree(dAh); cudaFree(dAl); cudaFree(dBh); cudaFree(dBl); cudaFree(dC); cudaFree(dbias);
        }
    }
    return 0;
}
