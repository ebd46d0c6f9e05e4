// Row-wise pieces of the transformer blocks that are not GEMMs.
//   layernorm : nn.LayerNorm over the channel dimension of a token matrix (code/transformer_model.py:16,
//               :71, :73; timm ViT blocks), bf16 in / bf16 out, statistics in fp32.  One warp per row.
#include "b200_fusion.h"
#include "common.cuh"

namespace b200 {

template <int VEC>  // uint4 (8 x bf16) vectors per lane: C = 256 * VEC
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C, const float* __restrict__ w,
                 const float* __restrict__ b, float eps, __nv_bfloat16* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    if (row >= rows) return;
    const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
    float f[VEC][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        unpack_bf16x8(__ldg(src + v * 32 + lane), f[v]);
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
    uint4* dst = reinterpret_cast<uint4*>(y + row * C);
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
        dst[v * 32 + lane] = pack_bf16x8(o);
    }
}

}  // namespace b200

extern "C" int b200_layernorm(const void* x, long long rows, int C, const float* w, const float* b, float eps, void* y,
                              void* stream) {
    using namespace b200;
    if (rows < 0 || C <= 0 || C % 256 != 0 || C > 1024) return -1;
    if (rows == 0) return 0;
    if (x == nullptr || w == nullptr || b == nullptr || y == nullptr) return -2;
    const long long blocks = (rows * 32 + 255) / 256;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* xi = static_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* yo = static_cast<__nv_bfloat16*>(y);
    switch (C / 256) {
        case 1: layernorm_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xi, rows, C, w, b, eps, yo); break;
        case 2: layernorm_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xi, rows, C, w, b, eps, yo); break;
        case 3: layernorm_kernel<3><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xi, rows, C, w, b, eps, yo); break;
        default: layernorm_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xi, rows, C, w, b, eps, yo); break;
    }
    return launch_status();
}
