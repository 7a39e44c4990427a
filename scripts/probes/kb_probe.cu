// Standalone timing / correctness probe for gemm_kb.cuh (dW_ih = dGI^T . U with both operands read in place).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o kb_probe kb_probe.cu && ./kb_probe [B]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../windgnn_b200/csrc/gemm_kb.cuh"

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                  \
        }                                                                             \
    } while (0)

__global__ void fill(float* p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = ((h & 0xffff) / 65536.0f) - 0.5f;
    }
}
// reference dW_ih[g][i] over a row sample: rows r with r % stride == 0 -> not comparable; use full sum in double on few (g,i)
__global__ void ref_wih(const float* U, const float* DG, double* out, long long R, int IP, int ld, const int* gi, int n) {
    const int e = blockIdx.x;
    if (e >= n) return;
    const int g = gi[2 * e], i = gi[2 * e + 1];
    double s = 0;
    for (long long r = threadIdx.x; r < R; r += blockDim.x)
        s += (double)DG[r * ld + g] * (double)U[((r >> 7) * IP + i) * 128 + (r & 127)];
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[e] = sh[0];
}
__global__ void reduce_parts(const float* part, int splits, int I, int ldc, int G, float* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= I * G) return;
    const int i = e / G, g = e % G;
    float s = 0;
    for (int z = 0; z < splits; ++z) s += part[((size_t)z * I + i) * ldc + g];
    out[(size_t)g * I + i] = s;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 4096;
    const int T = 168, S = 34, H = 102, I = S * 13, IP = 448, G = 3 * H, LD4 = 4 * H, KPd = 320, NPd = 448, ldu = 444;
    const long long R = (long long)B * T;
    const long long rt = (R + 127) / 128 * 128;
    float *U, *DG, *part, *dW, *wpb, *dU;
    CK(cudaMalloc(&U, rt * IP * 4));
    CK(cudaMalloc(&DG, R * LD4 * 4));
    CK(cudaMalloc(&dW, (size_t)G * I * 4));
    CK(cudaMalloc(&wpb, (size_t)NPd * KPd * 4));
    CK(cudaMalloc(&dU, R * ldu * 4));
    fill<<<1024, 256>>>(U, rt * IP, 1);
    fill<<<1024, 256>>>(DG, R * LD4, 2);
    fill<<<64, 256>>>(wpb, (size_t)NPd * KPd, 3);
    using C0 = wg::KbWih;
    const int mt = (IP + C0::kBM - 1) / C0::kBM, nt = (G + C0::kBN - 1) / C0::kBN;
    const long long ktall = (R + 15) / 16;
    long long sp = (3LL * 148) / (mt * nt);
    if (sp > ktall) sp = ktall;
    const long long ktper = (ktall + sp - 1) / sp;
    const int splits = (int)((ktall + ktper - 1) / ktper);
    const int ldc = nt * C0::kBN;
    CK(cudaMalloc(&part, (size_t)splits * I * ldc * 4));
    auto k0 = wg::gemm_kb_kernel<C0::kBM, C0::kBN, C0::kTN>;
    CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, C0::kSmemBytes));
    wg::KbArgs a0{};
    a0.a = U; a0.b = DG; a0.c = part; a0.R = R; a0.I = I; a0.IP = IP; a0.ld_dg = LD4; a0.ldc = ldc;
    a0.m_tiles = mt; a0.n_tiles = nt; a0.kt_per_split = ktper;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) k0<<<dim3(mt * nt, splits), C0::kThreads, C0::kSmemBytes>>>(a0);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("dW_ih  gemm_kb<%d,%d,0>: %.3f ms  (%.1f TFLOP/s useful)  grid %d x %d\n", C0::kBM, C0::kBN, ms / 5,
                        2.0 * R * G * I / (ms / 5 * 1e-3) / 1e12, mt * nt, splits);
    }
    // correctness: dW_ih on 64 sampled entries against a double-precision sum
    reduce_parts<<<(I * G + 255) / 256, 256>>>(part, splits, I, ldc, G, dW);
    std::vector<int> gi(128);
    for (int e = 0; e < 64; ++e) { gi[2 * e] = (e * 37) % G; gi[2 * e + 1] = (e * 101 + 7) % I; }
    gi[0] = G - 1; gi[1] = I - 1;
    int* dgi; double* dref;
    CK(cudaMalloc(&dgi, 128 * 4)); CK(cudaMalloc(&dref, 64 * 8));
    CK(cudaMemcpy(dgi, gi.data(), 128 * 4, cudaMemcpyHostToDevice));
    ref_wih<<<64, 256>>>(U, DG, dref, R, IP, LD4, dgi, 64);
    std::vector<double> ref(64);
    std::vector<float> got((size_t)G * I);
    CK(cudaMemcpy(ref.data(), dref, 64 * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got.data(), dW, got.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (int e = 0; e < 64; ++e) {
        const double d = fabs(got[(size_t)gi[2 * e] * I + gi[2 * e + 1]] - ref[e]);
        if (d > worst) worst = d;
        if (fabs(ref[e]) > scale) scale = fabs(ref[e]);
    }
    printf("dW_ih check: max abs err %.3e (scale %.3e)\n", worst, scale);
    return 0;
}
