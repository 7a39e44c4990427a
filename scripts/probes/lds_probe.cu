// Shared-memory wavefront probe: cycles per LDS for several lane->address patterns (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_probe lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>  // floats per load: 1, 2, 4
__global__ void probe(const int* __restrict__ lane_off, float* out, int iters, long long* cycles) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 0.001f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int off = lane_off[lane] + warp * 4;  // in floats; warps shifted a little
    float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + off * 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            unsigned a = base + u * 1024;  // stays inside the 32 KB buffer
            if (VEC == 4) {
                float x, y, z, w;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
                acc0 = __int_as_float(__float_as_int(acc0) ^ __float_as_int(x));
            } else if (VEC == 2) {
                float x, y;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
                acc0 = __int_as_float(__float_as_int(acc0) ^ __float_as_int(x));
            } else {
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
                acc0 = __int_as_float(__float_as_int(acc0) ^ __float_as_int(x));
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc0 + acc1 + acc2 + acc3 == 1234.5f) out[0] = acc0;
}

struct Pattern { const char* name; int vec; int off[32]; };

int main() {
    Pattern pats[24];
    int np = 0;
    auto add = [&](const char* n, int vec, auto f) {
        pats[np].name = n; pats[np].vec = vec;
        for (int l = 0; l < 32; ++l) pats[np].off[l] = f(l);
        ++np;
    };
    add("v4 all-same", 4, [](int l) { return 0; });
    add("v4 one addr per quarter (lane>>3)*4", 4, [](int l) { return (l >> 3) * 4; });
    add("v4 8 distinct same in each quarter (lane&7)*4", 4, [](int l) { return (l & 7) * 4; });
    add("v4 all distinct lane*4", 4, [](int l) { return l * 4; });
    add("v4 4 distinct lm-fastest (lane&3)*4", 4, [](int l) { return (l & 3) * 4; });
    add("v4 8 distinct, 2 per quarter (lane>>2)*4", 4, [](int l) { return (l >> 2) * 4; });
    add("v4 2 distinct per quarter, 4 rows stride 20 (lane>>3)*20", 4, [](int l) { return (l >> 3) * 20; });
    add("v4 8 rows stride 20 (lane&7)*20", 4, [](int l) { return (l & 7) * 20; });
    add("v4 rows stride 212, 4 rows (lane>>3)*212 [recur hs]", 4, [](int l) { return (l >> 3) * 212; });
    add("v4 conflict: 8 rows stride 32 (lane&7)*32", 4, [](int l) { return (l & 7) * 32; });
    add("v4 16 distinct contiguous (lane>>1)*4", 4, [](int l) { return (l >> 1) * 4; });
    add("v2 4 distinct (lane>>3)*2", 2, [](int l) { return (l >> 3) * 2; });
    add("v2 8 distinct (lane&7)*2", 2, [](int l) { return (l & 7) * 2; });
    add("v2 all distinct lane*2", 2, [](int l) { return l * 2; });
    add("v1 4 distinct (lane>>3)", 1, [](int l) { return (l >> 3); });
    add("v1 8 distinct (lane&7)", 1, [](int l) { return (l & 7); });
    add("v1 all distinct", 1, [](int l) { return l; });
    add("v1 all same", 1, [](int l) { return 0; });

    int* d_off; float* d_out; long long* d_cyc;
    cudaMalloc(&d_off, 32 * 4); cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, 8 * 148);
    const int iters = 2000, threads = 512, blocks = 148;
    for (int p = 0; p < np; ++p) {
        cudaMemcpy(d_off, pats[p].off, 32 * 4, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            if (rep == 1) cudaEventRecord(e0);
            if (pats[p].vec == 4) probe<4><<<blocks, threads, 32768>>>(d_off, d_out, iters, d_cyc);
            else if (pats[p].vec == 2) probe<2><<<blocks, threads, 32768>>>(d_off, d_out, iters, d_cyc);
            else probe<1><<<blocks, threads, 32768>>>(d_off, d_out, iters, d_cyc);
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long c[148];
        cudaMemcpy(c, d_cyc, 8 * 148, cudaMemcpyDeviceToHost);
        double warps = threads / 32.0;
        double per = (double)c[0] / ((double)iters * 16 * warps);
        double ns_per = ms * 1e6 / ((double)iters * 16 * warps);
        printf("%-58s clock64/LDS %.2f | %.3f ns/LDS per SM = %.2f cyc @1.92GHz (%s)\n", pats[p].name, per, ns_per,
               ns_per * 1.92, cudaGetErrorString(e));
    }
    return 0;
}
