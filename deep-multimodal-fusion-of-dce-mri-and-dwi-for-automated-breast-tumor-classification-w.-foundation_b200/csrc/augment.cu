// Training-time augmentation of the reference's data transforms, batched on the device
// (code/prepare_single_model.py:107-113: transforms.RandomAffine(degrees=90, translate=(0.1, 0.1), shear=(0.1, 0.1)),
// RandomHorizontalFlip, RandomVerticalFlip - applied per sample on the CPU there, ahead of Resize and the normaliser).
// One pass: out[b,c,i,j] = x[b,c, src(i', j')] with (i', j') the un-flipped position and src the nearest-neighbour
// inverse affine map torchvision's tensor backend builds (torchvision/transforms/_functional_tensor.py: _gen_affine_grid
// on half-integer pixel centres, grid_sample(mode="nearest", padding_mode="zeros", align_corners=False)); pixels that
// map outside the image take `fill`.  The per-case 2x3 inverse matrices and flip flags are sampled on the host with
// torchvision's own random-number call sequence (dataset.BatchAugment).
#include "b200_fusion.h"
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256)
augment_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W,
               const float* __restrict__ theta, const int* __restrict__ flips, float fill) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    int i = p / W, j = p % W;
    const int fl = flips != nullptr ? flips[b] : 0;
    if (fl & 2) i = H - 1 - i;  // vertical flip is applied last in the reference's Compose: undo it first
    if (fl & 1) j = W - 1 - j;
    const float* t = theta + static_cast<size_t>(b) * 6;
    // _gen_affine_grid: base grid on pixel centres relative to the image centre, theta rescaled by half the size
    const float xb = static_cast<float>(j) - 0.5f * W + 0.5f, yb = static_cast<float>(i) - 0.5f * H + 0.5f;
    const float hw = 0.5f * W, hh = 0.5f * H;
    const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xb, __fdiv_rn(t[0], hw)), __fmul_rn(yb, __fdiv_rn(t[1], hw))),
                               __fdiv_rn(t[2], hw));
    const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xb, __fdiv_rn(t[3], hh)), __fmul_rn(yb, __fdiv_rn(t[4], hh))),
                               __fdiv_rn(t[5], hh));
    // grid_sample, align_corners=False: pixel = ((g + 1) * size - 1) / 2, nearest = round half to even
    const float sx = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(W)), -1.0f), 2.0f);
    const float sy = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(H)), -1.0f), 2.0f);
    const float rx = nearbyintf(sx), ry = nearbyintf(sy);
    const bool inside = rx >= 0.f && rx <= static_cast<float>(W - 1) && ry >= 0.f && ry <= static_cast<float>(H - 1);
    const size_t plane = static_cast<size_t>(H) * W;
    const size_t src = inside ? static_cast<size_t>(ry) * W + static_cast<size_t>(rx) : 0;
    const float* xin = x + static_cast<size_t>(b) * C * plane;
    float* o = out + static_cast<size_t>(b) * C * plane + p;
    for (int c = 0; c < C; ++c) __stcs(o + c * plane, inside ? __ldg(xin + c * plane + src) : fill);
}

}  // namespace b200

extern "C" int b200_augment(const float* x, float* out, int B, int C, int H, int W, const float* theta,
                            const int* flips, float fill, void* stream) {
    if (B < 0 || C <= 0 || H <= 0 || W <= 0) return -1;
    if (B == 0) return 0;
    if (x == nullptr || out == nullptr || theta == nullptr || x == out) return -2;
    dim3 grid((H * W + 255) / 256, B);
    b200::augment_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, C, H, W, theta, flips, fill);
    return b200::launch_status();
}
