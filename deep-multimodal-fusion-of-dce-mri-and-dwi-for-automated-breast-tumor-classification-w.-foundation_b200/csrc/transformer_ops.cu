// Row-wise pieces of the transformer blocks that are not GEMMs.
//   layernorm : nn.LayerNorm over the channel dimension of a token matrix (code/transformer_model.py:16,
//               :71, :73; timm ViT blocks), bf16 in / bf16 out, statistics in fp32.  One warp per row.
#include "b200_fusion.h"
#include "common.cuh"

namespace b200 {

template <int VEC, bool IN32, bool OUT32>  // 8-element vectors per lane: C = 256 * VEC
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ xv, long long rows, int C, const float* __restrict__ w,
                 const float* __restrict__ b, float eps, void* __restrict__ yv) {
    const int lane = threadIdx.x & 31;
    const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    if (row >= rows) return;
    float f[VEC][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (IN32) {
            const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(xv) + row * C) + (v * 32 + lane) * 2;
            const float4 a = __ldg(src), c = __ldg(src + 1);
            f[v][0] = a.x; f[v][1] = a.y; f[v][2] = a.z; f[v][3] = a.w;
            f[v][4] = c.x; f[v][5] = c.y; f[v][6] = c.z; f[v][7] = c.w;
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(xv) + row * C);
            unpack_bf16x8(__ldg(src + v * 32 + lane), f[v]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += f[v][k];
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float d = f[v][k] - mean;
            q += d * d;
        }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int c0 = (v * 32 + lane) * 8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
        float o[8];
        o[0] = (f[v][0] - mean) * rstd * w0.x + b0.x;
        o[1] = (f[v][1] - mean) * rstd * w0.y + b0.y;
        o[2] = (f[v][2] - mean) * rstd * w0.z + b0.z;
        o[3] = (f[v][3] - mean) * rstd * w0.w + b0.w;
        o[4] = (f[v][4] - mean) * rstd * w1.x + b1.x;
        o[5] = (f[v][5] - mean) * rstd * w1.y + b1.y;
        o[6] = (f[v][6] - mean) * rstd * w1.z + b1.z;
        o[7] = (f[v][7] - mean) * rstd * w1.w + b1.w;
        if (OUT32) {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(yv) + row * C) + (v * 32 + lane) * 2;
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
            reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(yv) + row * C)[v * 32 + lane] = pack_bf16x8(o);
        }
    }
}

// ---------------------------------------------------------------- ViT glue -
// patchify: fp32 NCHW image -> bf16 patch matrix [B * gh * gw, C * P * P], k = (c*P + ky)*P + kx, i.e. the
// flattening of Conv2d(C, E, P, P).weight, so the patch embedding (timm PatchEmbed / foundation_model.py
// :388-412) becomes one tcgen05 GEMM.  One thread moves one 8-pixel run of a patch row.
__global__ void patchify_kernel(const float* __restrict__ x, const float* __restrict__ gate, int C, int H, int W, int P,
                                __nv_bfloat16* __restrict__ out, size_t total_runs) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total_runs) return;
    const int runs_per_row = P / 8;
    const int gw = W / P, gh = H / P;
    size_t r = i;
    const int run = static_cast<int>(r % runs_per_row); r /= runs_per_row;
    const int ky = static_cast<int>(r % P); r /= P;
    const int c = static_cast<int>(r % C); r /= C;
    const int px = static_cast<int>(r % gw); r /= gw;
    const int py = static_cast<int>(r % gh);
    const size_t b = r / gh;
    const float* src = x + ((b * C + c) * H + py * P + ky) * W + px * P + run * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), d = __ldg(reinterpret_cast<const float4*>(src) + 1);
    const float g = gate != nullptr ? __ldg(gate + b * C + c) : 1.f;  // modality-attention gate of this plane
    const float f[8] = {a.x * g, a.y * g, a.z * g, a.w * g, d.x * g, d.y * g, d.z * g, d.w * g};
    const size_t token = (b * gh + py) * gw + px;
    __nv_bfloat16* dst = out + token * (static_cast<size_t>(C) * P * P) + (static_cast<size_t>(c) * P + ky) * P + run * 8;
    *reinterpret_cast<uint4*>(dst) = pack_bf16x8(f);
}

// tokens: t[b,0,:] = cls + pos[0]; t[b,1+i,:] = patch[b,i,:] + pos[1+i]  (fp32 residual stream)
__global__ void vit_tokens_kernel(const __nv_bfloat16* __restrict__ patches, const float* __restrict__ cls,
                                  const float* __restrict__ pos, int n_patch, int E, float* __restrict__ t,
                                  size_t total_vec) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total_vec) return;
    const int ev = E / 8;
    const int e0 = static_cast<int>(i % ev) * 8;
    const size_t row = i / ev;
    const int tok = static_cast<int>(row % (n_patch + 1));
    const size_t b = row / (n_patch + 1);
    float f[8];
    if (tok == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = cls[e0 + k];
    } else {
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(patches + (b * n_patch + tok - 1) * E + e0)), f);
    }
    float* dst = t + row * E + e0;
    const float* pp = pos + static_cast<size_t>(tok) * E + e0;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = f[k] + pp[k];
}

// feature map of one block: fp32 stream [B, 1+n, E] -> bf16 [B, n, E] (cls token stripped; NHWC map)
__global__ void vit_feature_kernel(const float* __restrict__ t, int n_patch, int E, __nv_bfloat16* __restrict__ out,
                                   int out_ld, size_t total_vec) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= total_vec) return;
    const int ev = E / 8;
    const int e0 = static_cast<int>(i % ev) * 8;
    const size_t row = i / ev;
    const int tok = static_cast<int>(row % n_patch);
    const size_t b = row / n_patch;
    const float4* src = reinterpret_cast<const float4*>(t + (b * (n_patch + 1) + tok + 1) * E + e0);
    const float4 a = __ldg(src), d = __ldg(src + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
    *reinterpret_cast<uint4*>(out + row * out_ld + e0) = pack_bf16x8(f);
}

}  // namespace b200

extern "C" int b200_patchify(const float* x, const float* gate, int B, int C, int H, int W, int P, void* out,
                             void* stream) {
    using namespace b200;
    if (B < 0 || C <= 0 || P <= 0 || P % 8 != 0 || H % P != 0 || W % P != 0) return -1;
    if (B == 0) return 0;
    if (x == nullptr || out == nullptr || (reinterpret_cast<uintptr_t>(x) & 15)) return -2;
    const size_t total = static_cast<size_t>(B) * (H / P) * (W / P) * C * P * (P / 8);
    patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, gate, C, H, W, P, static_cast<__nv_bfloat16*>(out), total);
    return launch_status();
}

extern "C" int b200_vit_tokens(const void* patches, const float* cls, const float* pos, int B, int n_patch, int E,
                               float* t, void* stream) {
    using namespace b200;
    if (B < 0 || n_patch <= 0 || E % 8 != 0) return -1;
    if (B == 0) return 0;
    if (patches == nullptr || cls == nullptr || pos == nullptr || t == nullptr) return -2;
    const size_t total = static_cast<size_t>(B) * (n_patch + 1) * (E / 8);
    vit_tokens_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(patches), cls, pos, n_patch, E, t, total);
    return launch_status();
}

extern "C" int b200_vit_feature(const float* t, int B, int n_patch, int E, void* out, int out_ld, void* stream) {
    using namespace b200;
    if (B < 0 || n_patch <= 0 || E % 8 != 0 || out_ld % 8 != 0 || out_ld < E) return -1;
    if (reinterpret_cast<uintptr_t>(out) & 15) return -2;
    if (B == 0) return 0;
    if (t == nullptr || out == nullptr) return -2;
    const size_t total = static_cast<size_t>(B) * n_patch * (E / 8);
    vit_feature_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        t, n_patch, E, static_cast<__nv_bfloat16*>(out), out_ld, total);
    return launch_status();
}

template <bool IN32, bool OUT32>
static void ln_launch(int vec, unsigned blocks, cudaStream_t s, const void* x, long long rows, int C, const float* w,
                      const float* b, float eps, void* y) {
    switch (vec) {
        case 1: b200::layernorm_kernel<1, IN32, OUT32><<<blocks, 256, 0, s>>>(x, rows, C, w, b, eps, y); break;
        case 2: b200::layernorm_kernel<2, IN32, OUT32><<<blocks, 256, 0, s>>>(x, rows, C, w, b, eps, y); break;
        case 3: b200::layernorm_kernel<3, IN32, OUT32><<<blocks, 256, 0, s>>>(x, rows, C, w, b, eps, y); break;
        default: b200::layernorm_kernel<4, IN32, OUT32><<<blocks, 256, 0, s>>>(x, rows, C, w, b, eps, y); break;
    }
}

extern "C" int b200_layernorm(const void* x, int x_f32, long long rows, int C, const float* w, const float* b, float eps,
                              void* y, int y_f32, void* stream) {
    using namespace b200;
    if (rows < 0 || C <= 0 || C % 256 != 0 || C > 1024) return -1;
    if (rows == 0) return 0;
    if (x == nullptr || w == nullptr || b == nullptr || y == nullptr) return -2;
    const unsigned blocks = static_cast<unsigned>((rows * 32 + 255) / 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (x_f32 && y_f32) ln_launch<true, true>(C / 256, blocks, s, x, rows, C, w, b, eps, y);
    else if (x_f32) ln_launch<true, false>(C / 256, blocks, s, x, rows, C, w, b, eps, y);
    else if (y_f32) ln_launch<false, true>(C / 256, blocks, s, x, rows, C, w, b, eps, y);
    else ln_launch<false, false>(C / 256, blocks, s, x, rows, C, w, b, eps, y);
    return launch_status();
}
