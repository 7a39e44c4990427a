// step4 window extraction and the caller's de-normalise / last-step slice, on the device.
// Pure data movement (HBM-bound): coalesced, vectorised where alignment allows.
//
// Reference:
//   __create_sequences, src/step4_sequence_preparer.py:7-27 — window i of the pivoted table
//   data[time, station, column]:  x = data[i*L:(i+1)*L, :, 2:15]             (:13)
//                                 y = concat_k data[i*L+k : (i+1)*L+k, :, 13], k = 1, 2, 3   (:14-18)
//   followed by an (unseeded) shuffle of the windows (:23-26) — here an optional permutation input.
//   De-normalisation of the model output, src/main.py:103:  y * (wind_max - wind_min) + wind_min,
//   of which the evaluation keeps the last timestep only (src/main.py:116,131,146).
// The device table holds the numeric block (columns 2:15) only: feature f = column f + 2, so the
// label column 13 ("Wind Speed 10 m Avg.", src/step3_feature_extractor.py:17) is feature 11.

#include "../../include/windgnn_b200.h"

#include <cuda_runtime.h>

extern "C" int wg_internal_fail(int code, const char* msg);

namespace {

__global__ void windows_x_kernel(const float* __restrict__ table, const long long* __restrict__ perm,
                                 float* __restrict__ x, long long win_elems, long long N) {
    // one window = L*S*F contiguous floats in both source and destination
    const long long total = win_elems * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / win_elems, r = e - n * win_elems;
        const long long w = perm ? perm[n] : n;
        x[e] = __ldg(table + w * win_elems + r);
    }
}

__global__ void windows_y_kernel(const float* __restrict__ table, const long long* __restrict__ perm,
                                 float* __restrict__ y, int S, int F, int L, int label_f, int K, long long N) {
    const long long per_win = (long long)L * K * S;
    const long long total = per_win * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / per_win;
        long long r = e - n * per_win;
        const int l = (int)(r / (K * S));
        r -= (long long)l * K * S;
        const int k = (int)(r / S), s = (int)(r - (long long)k * S);
        const long long w = perm ? perm[n] : n;
        const long long t = w * L + l + 1 + k;  // wind speed k+1 hours after the window row
        y[e] = __ldg(table + (t * S + s) * F + label_f);
    }
}

__global__ void denorm_last_kernel(const float* __restrict__ out, float* __restrict__ pred, long long B, int T,
                                   int H, float scale, float vmin) {
    const long long total = B * H;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / H;
        const int j = (int)(e - b * H);
        const float v = __ldg(out + (b * T + (T - 1)) * H + j);
        pred[e] = __fadd_rn(__fmul_rn(v, scale), vmin);  // two roundings, as NumPy (no FMA)
    }
}

int grid_for(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

struct DevGuard {
    int prev = -1;
    bool ok = true;
    explicit DevGuard(int dev) {
        ok = cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess);
    }
    ~DevGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};


// Pivot of the long table (src/step4_sequence_preparer.py:36-47): the reference walks the stations in
// np.unique order and stacks each station's rows, in file order, along a new station axis:
//   table[t][s][:] = values of the t-th row (file order) whose station is s.
// One CTA per station scans the station-id column once; a row's time index is the number of earlier
// rows of the same station (block-wide exclusive prefix count of the match flags).
__global__ void __launch_bounds__(256) pivot_kernel(const int* __restrict__ station, const float* __restrict__ values,
                                                    float* __restrict__ table, int* __restrict__ counts,
                                                    long long n_rows, int S, int F, long long Ttot) {
    __shared__ int warp_sum[8];
    __shared__ int running;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (long long base = 0; base < n_rows; base += 256) {
        const long long i = base + tid;
        const bool hit = i < n_rows && station[i] == s;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_sum[warp] = __popc(m);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_sum[w];
        const long long t = before + __popc(m & ((1u << lane) - 1u));
        if (hit && t < Ttot) {
            const float* src = values + i * F;
            float* dst = table + (t * S + s) * F;
            for (int f = 0; f < F; ++f) dst[f] = __ldg(src + f);
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_sum[w];
            running += tot;
        }
        __syncthreads();
    }
    if (tid == 0) counts[s] = running;
}

}  // namespace

extern "C" {

int64_t wg_num_windows(int64_t Ttot, int L, int horizons) {
    if (Ttot < 0 || L <= 0 || horizons < 0) return -1;
    // the reference takes floor(Ttot / L) windows (step4:10); the labels of the last one need
    // `horizons` more rows, so a table that ends exactly on a window boundary loses that window here
    int64_t n = Ttot / L;
    while (n > 0 && n * L + horizons > Ttot) --n;
    return n;
}

int wg_make_windows_f32(const float* table, const int64_t* perm, float* x, float* y, int64_t Ttot, int S,
                        int F, int L, int label_f, int horizons, int64_t N, int device, void* stream) {
    if (Ttot <= 0 || S <= 0 || F <= 0 || L <= 0 || horizons < 0 || N < 0 || label_f < 0 || label_f >= F)
        return wg_internal_fail(WG_ERR_BAD_ARG, "make_windows: bad dimension");
    if (N * L + horizons > Ttot)
        return wg_internal_fail(WG_ERR_BAD_ARG, "make_windows: N windows plus the label horizon exceed the table");
    if (N == 0) return WG_OK;
    if (!table || (!x && !y)) return wg_internal_fail(WG_ERR_BAD_ARG, "make_windows: null pointer");
    DevGuard g(device);
    if (!g.ok) return wg_internal_fail(WG_ERR_CUDA, "make_windows: cannot select CUDA device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long* p = reinterpret_cast<const long long*>(perm);
    const long long win = (long long)L * S * F;
    if (x) windows_x_kernel<<<grid_for(win * N), 256, 0, st>>>(table, p, x, win, N);
    if (y && horizons > 0)
        windows_y_kernel<<<grid_for((long long)L * horizons * S * N), 256, 0, st>>>(table, p, y, S, F, L, label_f,
                                                                                    horizons, N);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return wg_internal_fail(WG_ERR_CUDA, cudaGetErrorString(e));
    return WG_OK;
}

int wg_denorm_last_step_f32(const float* out, float* pred, int64_t B, int T, int H, double vmin, double vmax,
                            int device, void* stream) {
    if (B < 0 || T <= 0 || H <= 0) return wg_internal_fail(WG_ERR_BAD_ARG, "denorm_last_step: bad dimension");
    if (B == 0) return WG_OK;
    if (!out || !pred) return wg_internal_fail(WG_ERR_BAD_ARG, "denorm_last_step: null pointer");
    DevGuard g(device);
    if (!g.ok) return wg_internal_fail(WG_ERR_CUDA, "denorm_last_step: cannot select CUDA device");
    // NumPy: float32 array * python float -> the scalar is rounded to float32 first
    const float scale = (float)(vmax - vmin), lo = (float)vmin;
    denorm_last_kernel<<<grid_for(B * H), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, pred, B, T, H, scale, lo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return wg_internal_fail(WG_ERR_CUDA, cudaGetErrorString(e));
    return WG_OK;
}


int wg_pivot_table_f32(const int32_t* station, const float* values, float* table, int32_t* counts, int64_t n_rows,
                       int S, int F, int64_t Ttot, int device, void* stream) {
    if (n_rows < 0 || S <= 0 || F <= 0 || Ttot < 0) return wg_internal_fail(WG_ERR_BAD_ARG, "pivot: bad dimension");
    if (!counts || (n_rows > 0 && (!station || !values)) || (Ttot > 0 && !table))
        return wg_internal_fail(WG_ERR_BAD_ARG, "pivot: null pointer");
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || (prev != device && cudaSetDevice(device) != cudaSuccess))
        return wg_internal_fail(WG_ERR_CUDA, "pivot: cannot select CUDA device");
    pivot_kernel<<<S, 256, 0, static_cast<cudaStream_t>(stream)>>>(station, values, table, counts, n_rows, S, F, Ttot);
    cudaError_t e = cudaGetLastError();
    if (prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return wg_internal_fail(WG_ERR_CUDA, cudaGetErrorString(e));
    return WG_OK;
}

}  // extern "C"
