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

}  // extern "C"
