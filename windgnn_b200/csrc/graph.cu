// Graph build on the device, bit-identical to the reference.  THIS FILE IS COMPILED WITH
// -fmad=false: the reference evaluates X*X + Y*Y with two roundings (NumPy/Python floats never
// fuse), and every other step is a single IEEE fp64 operation (+ - * / sqrt are correctly
// rounded on sm_100a).
//
// Reference: build_graph, src/step2_graph_builder.py:16-40
//   A_hat[i][i] = 1;  A_hat[i][j] = 1 / sqrt((X*X + Y*Y) / 100000000)            (:24-31)
//   D = diag(np.sum(A_hat, axis=0))   -- row-by-row accumulation, column j sums i = 0..S-1 (:34)
//   A* = D^-1/2 . A_hat . D^-1/2 via scipy.linalg.fractional_matrix_power           (:37-38)
// SciPy evaluates the -1/2 power of the diagonal matrix as inv(D) . diag(sqrt(d)), hence
//   dh_i = fl(fl(1/d_i) * fl(sqrt(d_i))),   A*_ij = fl(fl(dh_i * a_ij) * dh_j)
// (pinned by tests/golden/adj_ref_*.npy).  The Mercator projection
// (:8-13, libm log/tan) stays on the host.
//
// k > 0 selects the synthetic large-graph generator: same weights and normalisation on the
// symmetrised k-nearest-neighbour pattern (no counterpart in the reference; defined by
// oracle/graph_oracle.py:knn_pattern).

#include "../../include/windgnn_b200.h"

#include <cuda_runtime.h>

#include <cstdio>

// thread-local error text lives in abi.cu; this TU reports through a tiny shim
extern "C" int wg_internal_fail(int code, const char* msg);

namespace {

__device__ __forceinline__ double edge_key(const double* __restrict__ xy, int i, int j) {
    const double X = xy[2 * i] - xy[2 * j];
    const double Y = xy[2 * i + 1] - xy[2 * j + 1];
    return (X * X + Y * Y) / 100000000.0;  // step2:27-30 (no FMA: -fmad=false)
}

__global__ void a_hat_kernel(const double* __restrict__ xy, double* __restrict__ A, int S) {
    const long long n = (long long)S * S;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / S), j = (int)(e % S);
        A[e] = (i == j) ? 1.0 : 1.0 / sqrt(edge_key(xy, i, j));
    }
}

// one warp per station: the k nearest others by (key, index) ascending
__global__ void knn_kernel(const double* __restrict__ xy, int* __restrict__ nbr, int S, int k) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= S) return;
    const int i = warp;
    double last_key = -1.0;
    int last_j = -1;
    for (int r = 0; r < k; ++r) {
        double best = __longlong_as_double(0x7ff0000000000000LL);  // +inf
        int best_j = 0x7fffffff;
        for (int j = lane; j < S; j += 32) {
            if (j == i) continue;
            const double key = edge_key(xy, i, j);
            const bool after = (key > last_key) || (key == last_key && j > last_j);
            if (after && (key < best || (key == best && j < best_j))) {
                best = key;
                best_j = j;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
            if (ob < best || (ob == best && oj < best_j)) {
                best = ob;
                best_j = oj;
            }
        }
        if (lane == 0) nbr[(size_t)i * k + r] = best_j;
        last_key = best;
        last_j = best_j;
    }
}

__global__ void mask_scatter_kernel(const int* __restrict__ nbr, unsigned char* __restrict__ mask, int S,
                                    int k) {
    const long long n = (long long)S * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / k);
        const int j = nbr[e];
        if (j >= 0 && j < S) {
            mask[(size_t)i * S + j] = 1;
            mask[(size_t)j * S + i] = 1;
        }
    }
}

__global__ void mask_apply_kernel(double* __restrict__ A, const unsigned char* __restrict__ mask, int S) {
    const long long n = (long long)S * S;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / S), j = (int)(e % S);
        if (i != j && !mask[e]) A[e] = 0.0;
    }
}

// d_j = ((0 + A[0][j]) + A[1][j]) + ...   then dh_j = (1/d_j) * sqrt(d_j)
__global__ void degree_kernel(const double* __restrict__ A, double* __restrict__ dh, int S) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    double d = 0.0;
    for (int i = 0; i < S; ++i) d = d + A[(size_t)i * S + j];
    dh[j] = (1.0 / d) * sqrt(d);
}

__global__ void normalise_kernel(const double* __restrict__ A, const double* __restrict__ dh,
                                 double* __restrict__ out64, float* __restrict__ out32, int S) {
    const long long n = (long long)S * S;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / S), j = (int)(e % S);
        const double v = (dh[i] * A[e]) * dh[j];
        if (out64) out64[e] = v;
        if (out32) out32[e] = (float)v;  // round-to-nearest-even, = torch .float() (main.py:26)
    }
}


// ---- CSR output of the kNN graph: no S x S matrix anywhere (2 MB of pattern bits at S = 4096) ----
// The values are bit-identical to the dense path: a column sum that skips the exact zeros of the dense
// matrix is the same sequence of additions (x + 0.0 == x), the pattern is symmetric and a_ij == a_ji
// bit for bit (X and -X square to the same double), so d_j is row j's entries added in ascending column order.
__global__ void bits_scatter_kernel(const int* __restrict__ nbr, unsigned* __restrict__ bits, int S, int k, int W) {
    const long long n = (long long)S * (k + 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / (k + 1)), r = (int)(e % (k + 1));
        const int j = r == k ? i : nbr[(size_t)i * k + r];      // r == k: the self loop
        if (j >= 0 && j < S) {
            atomicOr(bits + (size_t)i * W + (j >> 5), 1u << (j & 31));
            atomicOr(bits + (size_t)j * W + (i >> 5), 1u << (i & 31));
        }
    }
}
__global__ void row_count_kernel(const unsigned* __restrict__ bits, int* __restrict__ cnt, int S, int W) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= S) return;
    int c = 0;
    for (int w = lane; w < W; w += 32) c += __popc(bits[(size_t)warp * W + w]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[warp] = c;
}
// exclusive scan of cnt[0..S) into rowptr[0..S]; one block (S is a station count)
__global__ void row_scan_kernel(const int* __restrict__ cnt, int* __restrict__ rowptr, int S) {
    __shared__ long long part[1024];
    const int t = threadIdx.x, n = blockDim.x;
    const int per = (S + n - 1) / n, lo = t * per, hi = lo + per < S ? lo + per : S;
    long long s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < n; ++i) { const long long v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    long long run = part[t];
    for (int i = lo; i < hi; ++i) { rowptr[i] = (int)run; run += cnt[i]; }
    if (t == n - 1) rowptr[S] = (int)run;   // threads past S have lo >= S: run == total for the last thread
}
// one warp per row: columns ascending, a_ij; lane 0 then adds the row in order -> dh_i
__global__ void csr_fill_kernel(const double* __restrict__ xy, const unsigned* __restrict__ bits,
                                const int* __restrict__ rowptr, int* __restrict__ colidx, double* __restrict__ aval,
                                double* __restrict__ dh, int S, int W, long long capacity) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= S) return;
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    if ((long long)e1 > capacity) return;   // the caller's buffers are too small: reported by the host
    int e = e0;
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        const unsigned b = w < W ? bits[(size_t)i * W + w] : 0u;
        const int c = __popc(b);
        int pre = c;   // inclusive scan of the counts over the lanes
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += v;
        }
        int pos = e + pre - c;
        unsigned bb = b;
        while (bb) {
            const int bit = __ffs(bb) - 1;
            bb &= bb - 1;
            const int j = (w << 5) + bit;
            colidx[pos] = j;
            aval[pos] = (j == i) ? 1.0 : 1.0 / sqrt(edge_key(xy, i, j));
            ++pos;
        }
        e += __shfl_sync(0xffffffffu, pre, 31);
    }
    __syncwarp();
    if (lane == 0) {
        double d = 0.0;
        for (int q = e0; q < e1; ++q) d = d + aval[q];
        dh[i] = (1.0 / d) * sqrt(d);
    }
}
__global__ void csr_normalise_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                     const double* __restrict__ aval, const double* __restrict__ dh,
                                     double* __restrict__ v64, float* __restrict__ v32, int S, long long capacity) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= S) return;
    const int e0 = rowptr[i], e1 = rowptr[i + 1];
    if ((long long)e1 > capacity) return;
    for (int e = e0 + lane; e < e1; e += 32) {
        const double v = (dh[i] * aval[e]) * dh[colidx[e]];
        if (v64) v64[e] = v;
        if (v32) v32[e] = (float)v;
    }
}

__device__ __forceinline__ unsigned long long splitmix64_nth(unsigned long long seed, unsigned long long n) {
    unsigned long long z = seed + n * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void synth_coords_kernel(double* __restrict__ latlon, int S, unsigned long long seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    // bounding box of the shipped stations (data/ACISStationCoordinates.csv:2-36)
    const double lat_lo = 50.10, lat_hi = 51.59, lon_lo = -113.36, lon_hi = -110.09;
    const double u_lat = (double)(splitmix64_nth(seed, 2ULL * i + 1) >> 11) * 0x1.0p-53;
    const double u_lon = (double)(splitmix64_nth(seed, 2ULL * i + 2) >> 11) * 0x1.0p-53;
    latlon[2 * i] = lat_lo + u_lat * (lat_hi - lat_lo);
    latlon[2 * i + 1] = lon_lo + u_lon * (lon_hi - lon_lo);
}

size_t al(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" {

size_t wg_build_graph_workspace_bytes(int S, int k) {
    if (S <= 0 || k < 0) return 0;
    const int kk = k > S - 1 ? S - 1 : k;
    size_t n = al((size_t)S * S * 8) + al((size_t)S * 8);
    if (kk > 0) n += al((size_t)S * kk * 4) + al((size_t)S * S);
    return n;
}

int wg_build_graph_f64(const double* xy, double* adj_f64, float* adj_f32, int S, int k, void* workspace,
                       size_t workspace_bytes, int device, void* stream) {
    if (S <= 0 || k < 0) return wg_internal_fail(WG_ERR_BAD_ARG, "build_graph: S must be > 0 and k >= 0");
    if (!xy || (!adj_f64 && !adj_f32)) return wg_internal_fail(WG_ERR_BAD_ARG, "build_graph: null pointer");
    const size_t need = wg_build_graph_workspace_bytes(S, k);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return wg_internal_fail(WG_ERR_WORKSPACE, "build_graph: workspace NULL, misaligned or too small");
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || (prev != device && cudaSetDevice(device) != cudaSuccess))
        return wg_internal_fail(WG_ERR_CUDA, "build_graph: cannot select CUDA device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int kk = k > S - 1 ? S - 1 : k;
    char* base = static_cast<char*>(workspace);
    double* A = reinterpret_cast<double*>(base);
    double* dh = reinterpret_cast<double*>(base + al((size_t)S * S * 8));
    const long long n = (long long)S * S;
    const int threads = 256;
    int blocks = (int)((n + threads - 1) / threads);
    if (blocks > 148 * 16) blocks = 148 * 16;

    a_hat_kernel<<<blocks, threads, 0, st>>>(xy, A, S);
    if (kk > 0) {
        int* nbr = reinterpret_cast<int*>(base + al((size_t)S * S * 8) + al((size_t)S * 8));
        unsigned char* mask =
            reinterpret_cast<unsigned char*>(base + al((size_t)S * S * 8) + al((size_t)S * 8) + al((size_t)S * kk * 4));
        cudaMemsetAsync(mask, 0, (size_t)S * S, st);
        knn_kernel<<<(S * 32 + threads - 1) / threads, threads, 0, st>>>(xy, nbr, S, kk);
        const long long ne = (long long)S * kk;
        int b2 = (int)((ne + threads - 1) / threads);
        if (b2 > 148 * 16) b2 = 148 * 16;
        mask_scatter_kernel<<<b2, threads, 0, st>>>(nbr, mask, S, kk);
        mask_apply_kernel<<<blocks, threads, 0, st>>>(A, mask, S);
    }
    degree_kernel<<<(S + 127) / 128, 128, 0, st>>>(A, dh, S);
    normalise_kernel<<<blocks, threads, 0, st>>>(A, dh, adj_f64, adj_f32, S);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "build_graph: %s", cudaGetErrorString(e));
        return wg_internal_fail(WG_ERR_CUDA, buf);
    }
    return WG_OK;
}


size_t wg_build_graph_csr_workspace_bytes(int S, int k) {
    if (S <= 0 || k <= 0) return 0;
    const int kk = k > S - 1 ? S - 1 : k;
    const size_t W = (size_t)(S + 31) / 32;
    // neighbour lists, pattern bits, row counts, a_ij per stored entry (capacity S (2k + 1)), dh
    return al((size_t)S * kk * 4) + al((size_t)S * W * 4) + al((size_t)S * 4) + al((size_t)S * (2 * kk + 1) * 8) +
           al((size_t)S * 8);
}

int wg_build_graph_csr_f64(const double* xy, int S, int k, int32_t* rowptr, int32_t* colidx, double* vals_f64,
                           float* vals_f32, int64_t capacity, void* workspace, size_t workspace_bytes, int device,
                           void* stream) {
    if (S <= 0 || k <= 0) return wg_internal_fail(WG_ERR_BAD_ARG, "build_graph_csr: S and k must be > 0");
    if (!xy || !rowptr || !colidx || (!vals_f64 && !vals_f32))
        return wg_internal_fail(WG_ERR_BAD_ARG, "build_graph_csr: null pointer");
    const int kk = k > S - 1 ? S - 1 : k;
    if (capacity < (int64_t)S * (2 * kk + 1))
        return wg_internal_fail(WG_ERR_BAD_ARG, "build_graph_csr: capacity must be at least S * (2k + 1) entries");
    const size_t need = wg_build_graph_csr_workspace_bytes(S, k);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return wg_internal_fail(WG_ERR_WORKSPACE, "build_graph_csr: workspace NULL, misaligned or too small");
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || (prev != device && cudaSetDevice(device) != cudaSuccess))
        return wg_internal_fail(WG_ERR_CUDA, "build_graph_csr: cannot select CUDA device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int W = (S + 31) / 32;
    char* base = static_cast<char*>(workspace);
    int* nbr = reinterpret_cast<int*>(base);                     base += al((size_t)S * kk * 4);
    unsigned* bits = reinterpret_cast<unsigned*>(base);          base += al((size_t)S * W * 4);
    int* cnt = reinterpret_cast<int*>(base);                     base += al((size_t)S * 4);
    double* aval = reinterpret_cast<double*>(base);              base += al((size_t)S * (2 * kk + 1) * 8);
    double* dh = reinterpret_cast<double*>(base);
    const int threads = 256;
    const int warp_blocks = (int)(((long long)S * 32 + threads - 1) / threads);
    cudaMemsetAsync(bits, 0, (size_t)S * W * 4, st);
    knn_kernel<<<warp_blocks, threads, 0, st>>>(xy, nbr, S, kk);
    const long long ne = (long long)S * (kk + 1);
    int b2 = (int)((ne + threads - 1) / threads);
    if (b2 > 148 * 16) b2 = 148 * 16;
    bits_scatter_kernel<<<b2, threads, 0, st>>>(nbr, bits, S, kk, W);
    row_count_kernel<<<warp_blocks, threads, 0, st>>>(bits, cnt, S, W);
    row_scan_kernel<<<1, 1024, 0, st>>>(cnt, rowptr, S);
    csr_fill_kernel<<<warp_blocks, threads, 0, st>>>(xy, bits, rowptr, colidx, aval, dh, S, W, (long long)capacity);
    csr_normalise_kernel<<<warp_blocks, threads, 0, st>>>(rowptr, colidx, aval, dh, vals_f64, vals_f32, S,
                                                          (long long)capacity);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof(buf), "build_graph_csr: %s", cudaGetErrorString(e));
        return wg_internal_fail(WG_ERR_CUDA, buf);
    }
    return WG_OK;
}

int wg_synthetic_coordinates_f64(double* latlon, int S, uint64_t seed, int device, void* stream) {
    if (S <= 0 || !latlon) return wg_internal_fail(WG_ERR_BAD_ARG, "synthetic_coordinates: bad argument");
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || (prev != device && cudaSetDevice(device) != cudaSuccess))
        return wg_internal_fail(WG_ERR_CUDA, "synthetic_coordinates: cannot select CUDA device");
    synth_coords_kernel<<<(S + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        latlon, S, (unsigned long long)seed);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return wg_internal_fail(WG_ERR_CUDA, cudaGetErrorString(e));
    return WG_OK;
}

}  // extern "C"
