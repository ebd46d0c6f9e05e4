// Per-(case, channel) input normalisation: HBM-bound, one CTA per image plane.
//
//  * DWI: z-score with unbiased std (clamped at 1e-6), clip to [z_lo, z_hi], affine to
//    [0,1]; the last channel is written as zeros when `skip_last` (the reference's
//    adc=True default).  Follows code/dataset.py:14-41 (DWINormalize.__call__).
//  * DCE: Nyul histogram standardisation: 11 order-statistic landmarks of the plane
//    (numpy "linear" percentile rule) -> piece-wise linear map onto the fitted average
//    landmarks -> piece-wise linear map onto the standard scale, all in float64 with
//    numpy.interp's exact branch structure.  Follows code/preprocess_helpers.py:85-120
//    (NyulStandardizer.transform).
//
// Both kernels read every input element from HBM once and write every output once.
#include <cfloat>

#include "b200_fusion.h"
#include "common.cuh"

namespace b200 {

constexpr int kNormThreads = 256;

// ------------------------------------------------------------------ DWI ----
template <int VEC>  // float4 values per thread kept in registers
__global__ void __launch_bounds__(kNormThreads)
dwi_normalize_reg_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int n, int skip_last,
                         float z_lo, float z_hi, float* __restrict__ plane_mean) {
    __shared__ double scratch[33];
    const int plane = blockIdx.x;
    const int c = plane % C;
    const int n4 = n >> 2;
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(plane) * n);
    float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(plane) * n);
    if (skip_last && c == C - 1) {
        for (int i = threadIdx.x; i < n4; i += kNormThreads) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (plane_mean != nullptr && threadIdx.x == 0) plane_mean[plane] = 0.f;
        return;
    }
    float4 v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int i = threadIdx.x + j * kNormThreads;
        v[j] = i < n4 ? __ldcs(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const double total = block_sum<double>(static_cast<double>(s), scratch);
    const float mean = static_cast<float>(total / n);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int i = threadIdx.x + j * kNormThreads;
        if (i < n4) {
            const float a = v[j].x - mean, b = v[j].y - mean, cc = v[j].z - mean, d = v[j].w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const double ss = block_sum<double>(static_cast<double>(q), scratch);
    // torch.std(): unbiased (n-1); n == 1 gives NaN there as well.
    const float sd = fmaxf(static_cast<float>(sqrt(ss / static_cast<double>(n - 1))), 1e-6f);
    const float range = z_hi - z_lo;
    float osum = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int i = threadIdx.x + j * kNormThreads;
        if (i < n4) {
            float4 o;
            o.x = (fminf(fmaxf((v[j].x - mean) / sd, z_lo), z_hi) - z_lo) / range;
            o.y = (fminf(fmaxf((v[j].y - mean) / sd, z_lo), z_hi) - z_lo) / range;
            o.z = (fminf(fmaxf((v[j].z - mean) / sd, z_lo), z_hi) - z_lo) / range;
            o.w = (fminf(fmaxf((v[j].w - mean) / sd, z_lo), z_hi) - z_lo) / range;
            osum += (o.x + o.y) + (o.z + o.w);
            __stcs(dst + i, o);
        }
    }
    if (plane_mean != nullptr) {
        const double om = block_sum<double>(static_cast<double>(osum), scratch);
        if (threadIdx.x == 0) plane_mean[plane] = static_cast<float>(om / n);
    }
}

// Any plane size / alignment: three sweeps, the 2nd and 3rd hit L2 (a plane is <= 200 KB).
__global__ void __launch_bounds__(kNormThreads)
dwi_normalize_stream_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int n, int skip_last,
                            float z_lo, float z_hi, float* __restrict__ plane_mean) {
    __shared__ double scratch[33];
    const int plane = blockIdx.x;
    const int c = plane % C;
    const float* src = x + static_cast<size_t>(plane) * n;
    float* dst = out + static_cast<size_t>(plane) * n;
    if (skip_last && c == C - 1) {
        for (int i = threadIdx.x; i < n; i += kNormThreads) dst[i] = 0.f;
        if (plane_mean != nullptr && threadIdx.x == 0) plane_mean[plane] = 0.f;
        return;
    }
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kNormThreads) s += static_cast<double>(src[i]);
    const float mean = static_cast<float>(block_sum<double>(s, scratch) / n);
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += kNormThreads) {
        const float d = src[i] - mean;
        q += static_cast<double>(d * d);
    }
    const double ss = block_sum<double>(q, scratch);
    const float sd = fmaxf(static_cast<float>(sqrt(ss / static_cast<double>(n - 1))), 1e-6f);
    const float range = z_hi - z_lo;
    double osum = 0.0;
    for (int i = threadIdx.x; i < n; i += kNormThreads) {
        const float o = (fminf(fmaxf((src[i] - mean) / sd, z_lo), z_hi) - z_lo) / range;
        osum += static_cast<double>(o);
        dst[i] = o;
    }
    if (plane_mean != nullptr) {
        const double om = block_sum<double>(osum, scratch);
        if (threadIdx.x == 0) plane_mean[plane] = static_cast<float>(om / n);
    }
}

// ----------------------------------------------------------------- Nyul ----
constexpr int kMaxLandmarks = 16;
constexpr int kNyulThreads = 512;

// numpy.interp for one sample: xp ascending (ties allowed), float64 throughout, no FMA
// contraction (matches the C loop in numpy's compiled_base.c).
__device__ __forceinline__ double np_interp(double xv, const double* xp, const double* fp, const double* slope, int L) {
    if (xv != xv) return xv;
    if (xv < xp[0]) return fp[0];
    if (xv > xp[L - 1]) return fp[L - 1];
    int j = 0;  // largest j with xp[j] <= xv
#pragma unroll 1
    for (int i = 1; i < L; ++i)
        if (xp[i] <= xv) j = i;
    if (j == L - 1) return fp[j];
    if (xp[j] == xv) return fp[j];
    double r = __dadd_rn(__dmul_rn(slope[j], __dadd_rn(xv, -xp[j])), fp[j]);
    if (r != r) {
        r = __dadd_rn(__dmul_rn(slope[j], __dadd_rn(xv, -xp[j + 1])), fp[j + 1]);
        if (r != r && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

__global__ void __launch_bounds__(kNyulThreads)
nyul_transform_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int n, int npad, int L,
                      const double* __restrict__ avg_landmarks,  // [C, L]
                      const double* __restrict__ standard_scale,  // [L]
                      const int* __restrict__ prev_index,         // [L] floor(q*(n-1))
                      const double* __restrict__ gamma,           // [L] fractional part
                      float* __restrict__ plane_mean) {
    extern __shared__ float sorted[];  // npad floats
    __shared__ double s_orig[kMaxLandmarks], s_avg[kMaxLandmarks], s_std[kMaxLandmarks];
    __shared__ double s_slope1[kMaxLandmarks], s_slope2[kMaxLandmarks];
    __shared__ double scratch[33];
    const int plane = blockIdx.x;
    const int c = plane % C;
    const float* src = x + static_cast<size_t>(plane) * n;
    float* dst = out + static_cast<size_t>(plane) * n;

    for (int i = threadIdx.x; i < npad; i += kNyulThreads) sorted[i] = i < n ? __ldcs(src + i) : FLT_MAX;
    __syncthreads();
    // Bitonic sort, ascending.
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npad >> 1); t += kNyulThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
                const int l = i | j;
                const bool up = (i & k) == 0;
                const float a = sorted[i], b = sorted[l];
                if ((a > b) == up) {
                    sorted[i] = b;
                    sorted[l] = a;
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x < L) {
        const int t = threadIdx.x;
        const int lo = prev_index[t];
        const int hi = min(lo + 1, n - 1);
        const float a = sorted[lo], b = sorted[hi];
        const float diff = b - a;  // numpy subtracts in the array dtype (float32) first
        const double g = gamma[t];
        double pv;
        if (g >= 0.5) pv = __dadd_rn(static_cast<double>(b), -__dmul_rn(static_cast<double>(diff), __dadd_rn(1.0, -g)));
        else pv = __dadd_rn(static_cast<double>(a), __dmul_rn(static_cast<double>(diff), g));
        s_orig[t] = pv;
        s_avg[t] = avg_landmarks[c * L + t];
        s_std[t] = standard_scale[t];
    }
    __syncthreads();
    if (threadIdx.x < L - 1) {
        const int t = threadIdx.x;
        s_slope1[t] = __ddiv_rn(__dadd_rn(s_avg[t + 1], -s_avg[t]), __dadd_rn(s_orig[t + 1], -s_orig[t]));
        s_slope2[t] = __ddiv_rn(__dadd_rn(s_std[t + 1], -s_std[t]), __dadd_rn(s_avg[t + 1], -s_avg[t]));
    }
    __syncthreads();
    double osum = 0.0;
    for (int i = threadIdx.x; i < n; i += kNyulThreads) {
        const double xv = static_cast<double>(src[i]);  // second read of the plane: L2/L1 hit
        const double mid = np_interp(xv, s_orig, s_avg, s_slope1, L);
        const float o = static_cast<float>(np_interp(mid, s_avg, s_std, s_slope2, L));
        osum += static_cast<double>(o);
        __stcs(dst + i, o);
    }
    if (plane_mean != nullptr) {
        const double om = block_sum<double>(osum, scratch);
        if (threadIdx.x == 0) plane_mean[plane] = static_cast<float>(om / n);
    }
}

__global__ void __launch_bounds__(kNormThreads)
plane_mean_kernel(const float* __restrict__ x, int n, float* __restrict__ plane_mean) {
    __shared__ double scratch[33];
    const float* src = x + static_cast<size_t>(blockIdx.x) * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kNormThreads) s += static_cast<double>(src[i]);
    const double t = block_sum<double>(s, scratch);
    if (threadIdx.x == 0) plane_mean[blockIdx.x] = static_cast<float>(t / n);
}

}  // namespace b200

extern "C" int b200_dwi_normalize(const float* x, float* out, int planes, int C, int n, int skip_last, float z_lo,
                                  float z_hi, float* plane_mean, void* stream) {
    using namespace b200;
    if (planes < 0 || C <= 0 || n <= 0 || planes % C != 0) return -1;
    if (planes == 0) return 0;
    if (x == nullptr || out == nullptr) return -2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool aligned = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (aligned && n <= kNormThreads * 4 * 4)
        dwi_normalize_reg_kernel<4><<<planes, kNormThreads, 0, s>>>(x, out, C, n, skip_last, z_lo, z_hi, plane_mean);
    else if (aligned && n <= kNormThreads * 4 * 8)
        dwi_normalize_reg_kernel<8><<<planes, kNormThreads, 0, s>>>(x, out, C, n, skip_last, z_lo, z_hi, plane_mean);
    else
        dwi_normalize_stream_kernel<<<planes, kNormThreads, 0, s>>>(x, out, C, n, skip_last, z_lo, z_hi, plane_mean);
    return launch_status();
}

extern "C" int b200_nyul_transform(const float* x, float* out, int planes, int C, int n, int L,
                                   const double* avg_landmarks, const double* standard_scale, const int* prev_index,
                                   const double* gamma, float* plane_mean, void* stream) {
    using namespace b200;
    if (planes < 0 || C <= 0 || n <= 0 || planes % C != 0 || L < 2 || L > kMaxLandmarks) return -1;
    if (planes == 0) return 0;
    if (x == nullptr || out == nullptr || avg_landmarks == nullptr || standard_scale == nullptr ||
        prev_index == nullptr || gamma == nullptr)
        return -2;
    int npad = 2;
    while (npad < n) npad <<= 1;
    const size_t smem = static_cast<size_t>(npad) * sizeof(float);
    if (smem > 200 * 1024) return -3;  // planes above 51200 samples need the multi-pass select (not built yet)
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(nyul_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = smem;
    }
    nyul_transform_kernel<<<planes, kNyulThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        x, out, C, n, npad, L, avg_landmarks, standard_scale, prev_index, gamma, plane_mean);
    return launch_status();
}

extern "C" int b200_plane_mean(const float* x, int planes, int n, float* plane_mean, void* stream) {
    using namespace b200;
    if (planes < 0 || n <= 0) return -1;
    if (planes == 0) return 0;
    if (x == nullptr || plane_mean == nullptr) return -2;
    plane_mean_kernel<<<planes, kNormThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, n, plane_mean);
    return launch_status();
}
