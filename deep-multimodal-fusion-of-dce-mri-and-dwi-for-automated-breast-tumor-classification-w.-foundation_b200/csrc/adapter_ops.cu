// HBM-bound helpers of the backbone-adapter path (ViT-B/16 features at 14 x 14 feeding the CNN blocks):
//   * per-case channel sums of an NHWC map (global average pool for SE / classifier when the GEMM epilogue
//     cannot stage the tile, i.e. for maps whose width does not divide 128),
//   * the backbone mix  GroupNorm(C, C)(alpha * f_b + (1 - alpha) * f)   (reference model_module.py:596-597,
//     :673-675, :688-690; groups == channels, i.e. an instance norm per (case, channel) with affine),
//   * AdaptiveAvgPool2d to the projector grid (reference :531-534, :707-710) for bf16 NHWC maps and for the
//     1-channel fp32 reconstruction maps,
//   * the element-wise sum of two maps (f2 + f1_aligned when FeatureDownAlign is the identity, :682-683).
// All maps are NHWC bf16 with the channel count a multiple of 8; one thread moves 8 channels (16 bytes).
#include "b200_fusion.h"
#include "common.cuh"

namespace b200 {

constexpr int kPixLanes = 32;  // pixel lanes per CTA; x 8 channel octets = 256 threads = 64 channels

// grid (C/64, B); thread = (octet o = tid & 7, pixel lane pl = tid >> 3)
__global__ void __launch_bounds__(256)
channel_sums_kernel(const __nv_bfloat16* __restrict__ x, int ld, int npix, int C, float* __restrict__ out) {
    __shared__ float part[kPixLanes][65];
    const int o = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c0 = blockIdx.x * 64 + o * 8;
    const size_t b = blockIdx.y;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
        const __nv_bfloat16* base = x + b * static_cast<size_t>(npix) * ld + c0;
        for (int p = pl; p < npix; p += kPixLanes) {
            float f[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(p) * ld)), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += f[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[pl][o * 8 + k] = s[k];
    __syncthreads();
    if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < C) {
        float t = 0.f;
#pragma unroll 8
        for (int r = 0; r < kPixLanes; ++r) t += part[r][threadIdx.x];
        out[b * C + blockIdx.x * 64 + threadIdx.x] = t;
    }
}

// grid (C/64, B).  Three passes over the case's [npix, 64]-channel slab (mean, centred variance, write); the
// slab is at most a few hundred KB so passes two and three hit L2 / L1.
__global__ void __launch_bounds__(256)
mix_instnorm_kernel(const __nv_bfloat16* __restrict__ fb, const __nv_bfloat16* __restrict__ f, int npix, int C,
                    const float* __restrict__ weight_logit, const float* __restrict__ gn_w,
                    const float* __restrict__ gn_b, float eps, __nv_bfloat16* __restrict__ out) {
    __shared__ float part[kPixLanes][65];
    __shared__ float stat[2][64];
    const int o = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c0 = blockIdx.x * 64 + o * 8;
    const bool live = c0 < C;
    const size_t off = blockIdx.y * static_cast<size_t>(npix) * C + c0;
    const float alpha = sigmoidf_(*weight_logit), beta = 1.f - alpha;
    auto mixed = [&](int p, float* m) {
        float a[8], c[8];
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(fb + off + static_cast<size_t>(p) * C)), a);
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(f + off + static_cast<size_t>(p) * C)), c);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = alpha * a[k] + beta * c[k];
    };
    auto reduce_to = [&](const float* s, float* dst, float scale) {
#pragma unroll
        for (int k = 0; k < 8; ++k) part[pl][o * 8 + k] = s[k];
        __syncthreads();
        if (threadIdx.x < 64) {
            float t = 0.f;
#pragma unroll 8
            for (int r = 0; r < kPixLanes; ++r) t += part[r][threadIdx.x];
            dst[threadIdx.x] = t * scale;
        }
        __syncthreads();
    };
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) {
        for (int p = pl; p < npix; p += kPixLanes) {
            float m[8];
            mixed(p, m);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += m[k];
        }
    }
    reduce_to(s, stat[0], 1.f / npix);  // mean
    float mu[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mu[k] = stat[0][o * 8 + k];
        s[k] = 0.f;
    }
    if (live) {
        for (int p = pl; p < npix; p += kPixLanes) {
            float m[8];
            mixed(p, m);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += (m[k] - mu[k]) * (m[k] - mu[k]);
        }
    }
    reduce_to(s, stat[1], 1.f / npix);  // biased variance (GroupNorm)
    if (!live) return;
    float g[8], h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float r = rsqrtf(stat[1][o * 8 + k] + eps) * gn_w[c0 + k];
        g[k] = r;
        h[k] = gn_b[c0 + k] - mu[k] * r;
    }
    for (int p = pl; p < npix; p += kPixLanes) {
        float m[8];
        mixed(p, m);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = m[k] * g[k] + h[k];
        *reinterpret_cast<uint4*>(out + off + static_cast<size_t>(p) * C) = pack_bf16x8(m);
    }
}

// AdaptiveAvgPool2d: output cell (i, j) averages rows [floor(i*H/Ho), ceil((i+1)*H/Ho)) x the same along w.
__device__ __forceinline__ void pool_window(int i, int n_in, int n_out, int& lo, int& hi) {
    lo = (i * n_in) / n_out;
    hi = ((i + 1) * n_in + n_out - 1) / n_out;
}

__global__ void adaptive_pool_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int Ho, int Wo, int act,
                                     __nv_bfloat16* __restrict__ out, size_t total_vec) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total_vec) return;
    const int cv = C >> 3;
    const int c0 = static_cast<int>(i % cv) << 3;
    size_t r = i / cv;
    const int ow = static_cast<int>(r % Wo); r /= Wo;
    const int oh = static_cast<int>(r % Ho);
    const size_t b = r / Ho;
    int h0, h1, w0, w1;
    pool_window(oh, H, Ho, h0, h1);
    pool_window(ow, W, Wo, w0, w1);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) {
            float f[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(x + ((b * H + h) * W + w) * C + c0)), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += f[k];
        }
    const float inv = 1.f / static_cast<float>((h1 - h0) * (w1 - w0));
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        float2 v = make_float2(s[k] * inv, s[k + 1] * inv);
        if (act == 1) v = gelu_fast2(v);
        s[k] = v.x;
        s[k + 1] = v.y;
    }
    reinterpret_cast<uint4*>(out)[i] = pack_bf16x8(s);
}

__global__ void adaptive_pool_c1_kernel(const float* __restrict__ x, int H, int W, int Ho, int Wo,
                                        float* __restrict__ out, size_t total) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total) return;
    size_t r = i;
    const int ow = static_cast<int>(r % Wo); r /= Wo;
    const int oh = static_cast<int>(r % Ho);
    const size_t b = r / Ho;
    int h0, h1, w0, w1;
    pool_window(oh, H, Ho, h0, h1);
    pool_window(ow, W, Wo, w0, w1);
    float s = 0.f;
    for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) s += __ldg(x + (b * H + h) * W + w);
    out[i] = s / static_cast<float>((h1 - h0) * (w1 - w0));
}

__global__ void add_maps_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                size_t total_vec) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_vec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float f[8], g[8];
        unpack_bf16x8(__ldg(a + i), f);
        unpack_bf16x8(__ldg(b + i), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += g[k];
        out[i] = pack_bf16x8(f);
    }
}

// Test-time-augmentation flips (reference train.py:916-923: torch.flip over the last and / or second-last
// dimension): every [H, W] plane of a [planes, H, W] fp32 tensor, one output row per (blockIdx.y, plane).
__global__ void flip_planes_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, int flip_w,
                                   int flip_h) {
    const int h = blockIdx.y;
    const size_t plane = blockIdx.z;
    const float* src = x + (plane * H + (flip_h ? H - 1 - h : h)) * W;
    float* dst = out + (plane * H + h) * W;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x)
        dst[w] = __ldg(src + (flip_w ? W - 1 - w : w));
}

}  // namespace b200

using namespace b200;

extern "C" int b200_channel_sums(const void* x, int x_ld, int B, int npix, int C, float* out, void* stream) {
    if (B < 0 || npix <= 0 || C <= 0 || C % 8 != 0 || x_ld % 8 != 0 || x_ld < C) return -1;
    if (B == 0) return 0;
    if (x == nullptr || out == nullptr || (reinterpret_cast<uintptr_t>(x) & 15)) return -2;
    if (B > 65535) return -3;
    const dim3 grid((C + 63) / 64, B);
    channel_sums_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), x_ld,
                                                                             npix, C, out);
    return launch_status();
}

extern "C" int b200_mix_instnorm(const void* fb, const void* f, int B, int npix, int C, const float* weight_logit,
                                 const float* gn_w, const float* gn_b, float eps, void* out, void* stream) {
    if (B < 0 || npix <= 0 || C <= 0 || C % 8 != 0) return -1;
    if (B == 0) return 0;
    if (fb == nullptr || f == nullptr || weight_logit == nullptr || gn_w == nullptr || gn_b == nullptr || out == nullptr)
        return -2;
    if (B > 65535) return -3;
    const dim3 grid((C + 63) / 64, B);
    mix_instnorm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(fb), static_cast<const __nv_bfloat16*>(f), npix, C, weight_logit, gn_w, gn_b,
        eps, static_cast<__nv_bfloat16*>(out));
    return launch_status();
}

extern "C" int b200_adaptive_pool(const void* x, int x_f32, int B, int H, int W, int C, int Ho, int Wo, int act,
                                  void* out, void* stream) {
    if (B < 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || C <= 0) return -1;
    if (B == 0) return 0;
    if (x == nullptr || out == nullptr) return -2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (x_f32) {  // 1-channel fp32 map (reconstruction heads), fp32 out
        if (C != 1 || act != 0) return -1;
        const size_t total = static_cast<size_t>(B) * Ho * Wo;
        adaptive_pool_c1_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
            static_cast<const float*>(x), H, W, Ho, Wo, static_cast<float*>(out), total);
        return launch_status();
    }
    if (C % 8 != 0 || (act != 0 && act != 1)) return -1;
    const size_t total_vec = static_cast<size_t>(B) * Ho * Wo * (C / 8);
    adaptive_pool_kernel<<<static_cast<unsigned>((total_vec + 255) / 256), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(x), H, W, C, Ho, Wo, act, static_cast<__nv_bfloat16*>(out), total_vec);
    return launch_status();
}

extern "C" int b200_add_maps(const void* a, const void* b, long long n_elems, void* out, void* stream) {
    if (n_elems < 0 || n_elems % 8 != 0) return -1;
    if (n_elems == 0) return 0;
    if (a == nullptr || b == nullptr || out == nullptr) return -2;
    const size_t total_vec = static_cast<size_t>(n_elems / 8);
    size_t blocks = (total_vec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    add_maps_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), total_vec);
    return launch_status();
}

extern "C" int b200_flip_planes(const float* x, float* out, long long planes, int H, int W, int flip_w, int flip_h,
                                void* stream) {
    if (planes < 0 || H <= 0 || W <= 0 || H > 65535) return -1;
    if (planes == 0) return 0;
    if (x == nullptr || out == nullptr || x == out) return -2;
    const int threads = W >= 256 ? 256 : (W >= 128 ? 128 : 64);
    for (long long p0 = 0; p0 < planes; p0 += 65535) {  // gridDim.z limit
        const unsigned nz = static_cast<unsigned>(planes - p0 < 65535 ? planes - p0 : 65535);
        flip_planes_kernel<<<dim3((W + threads - 1) / threads, H, nz), threads, 0, static_cast<cudaStream_t>(stream)>>>(
            x + p0 * H * W, out + p0 * H * W, H, W, flip_w, flip_h);
    }
    return launch_status();
}
