// Weight gradient of a 1x1 / 3x3 (stride 1, padding taps/2) convolution on the tensor cores, sm_100a.
//
//   dW[co, ci, tap] += sum_{pixels p} dY[p, co] * X[p + offset(tap), ci]
//
// i.e. a GEMM whose REDUCTION dimension is the pixel index.  Both operands live in HBM as NHWC bf16 maps, so
// both are "MN-major" for this product (the channel index is contiguous, the reduction index strides): the same
// TMA boxes the forward kernel loads - {64 channels, BW, BH, 1} pixels, 128-byte swizzle, the tap shift as a
// coordinate offset and the zero padding as TMA out-of-bounds fill - are handed to tcgen05.mma through MN-major
// shared-memory descriptors (instruction descriptor a_major = b_major = 1).  No transposed copy of either map
// is ever made.
//
//   tile         M = 128 output channels x N = BN input channels for ONE tap, fp32 accumulator in TMEM
//   k-block      64 pixels: A = two 8 KB boxes (64 co each), B = BN/64 boxes of 8 KB (64 ci each)
//   split-K      the (co tile, ci tile, tap) units are few (9 .. 36), so each unit's pixel range is split over
//                grid / units CTAs; every CTA reduces its slice in TMEM and adds it into dW with fp32 RED atomics
//                (the caller zeroes dW, or keeps accumulating micro-batches into it)
//   warps        0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM -> atomics; warp 2 also owns TMEM alloc)
//
// Serves the backward pass of every tensor-core convolution of the encoders and the fusion head
// (reference: torch autograd of nn.Conv2d, code/model_module.py:259-269, :113-118, :337-345, :386-390, :857-858).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "b200_fusion.h"
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

struct WgradParams {
    int H, W, Cin, Cout, taps;
    int BW, BH;          // pixel box of one k-block: BW * BH == 64
    int blocks_w;        // W / BW
    int blocks_per_case; // blocks_w * (H / BH)
    int total_blocks;    // B * blocks_per_case
    int n_units;         // co tiles * ci tiles * taps
    int ci_tiles;
    int n_splits;        // CTAs per unit
    int stages;
    float* dw;           // [Cout, Cin, taps] fp32, accumulated
};

constexpr int kWgKP = 64;                  // pixels per k-block
constexpr int kWgBox = kWgKP * 128;        // one {64 ch x 64 px} box: 8 KB
constexpr int kWgThreads = 6 * 32;
constexpr int kWgMaxStages = 8;

// MN-major operand under the 128-byte swizzle: 64 channels (128 B) are contiguous, 8 consecutive pixels form a
// 1 KB swizzle atom (SBO = 1024), the next 64 channels live in the next box (LBO = box bytes).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(uint32_t m, uint32_t n) {
    return umma_idesc_bf16(m, n) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
    constexpr int kABytes = 2 * kWgBox;
    constexpr int kBBytes = (BN / 64) * kWgBox;
    constexpr int kStageBytes = kABytes + kBBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw)) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * kStageBytes);
    uint64_t* empty = full + kWgMaxStages;
    uint64_t* tfull = empty + kWgMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = static_cast<int>(blockIdx.x) % p.n_units;
    const int split = static_cast<int>(blockIdx.x) / p.n_units;
    const int tap = unit % p.taps;
    const int ci_tile = (unit / p.taps) % p.ci_tiles;
    const int co_tile = unit / (p.taps * p.ci_tiles);
    const int per = (p.total_blocks + p.n_splits - 1) / p.n_splits;
    const int kb0 = split * per;
    const int kb1 = min(kb0 + per, p.total_blocks);
    const int n_kb = kb1 > kb0 ? kb1 - kb0 : 0;
    int dy = 0, dx = 0;
    if (p.taps == 9) {
        dy = tap / 3 - 1;
        dx = tap % 3 - 1;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int b = kb / p.blocks_per_case;
                const int r = kb - b * p.blocks_per_case;
                const int w0 = (r % p.blocks_w) * p.BW;
                const int h0 = (r / p.blocks_w) * p.BH;
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], kStageBytes);
                uint8_t* dst = smem + stage * kStageBytes;
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    tma_load_4d(dst + i * kWgBox, &tmDY, &full[stage], co_tile * 128 + i * 64, w0, h0, b);
#pragma unroll
                for (int i = 0; i < BN / 64; ++i)
                    tma_load_4d(dst + kABytes + i * kWgBox, &tmX, &full[stage], ci_tile * BN + i * 64, w0 + dx, h0 + dy, b);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_kb > 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_mn(128, BN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
                const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                for (int k = 0; k < kWgKP / 16; ++k) {
                    // 16 pixels = two 8-pixel swizzle atoms = 2 KB further into every box
                    const uint64_t da = umma_desc_mn_sw128(a_addr + k * 2048, kWgBox);
                    const uint64_t db = umma_desc_mn_sw128(b_addr + k * 2048, kWgBox);
                    umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty[stage]);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(tfull);
        }
    } else if (n_kb > 0) {
        // epilogue: warp w may read TMEM lanes [32 * (w & 3), +32) = output channels co_tile * 128 + that range
        const int q = warp & 3;
        const int co = co_tile * 128 + q * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t tm_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* const row = p.dw + (static_cast<long long>(co) * p.Cin + ci_tile * BN) * p.taps + tap;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
            uint32_t r[32];
            tmem_ld_32x32(tm_row + ch * 32, r);
            tmem_ld_wait32(r);
            if (co < p.Cout) {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(row + static_cast<long long>(ch * 32 + j) * p.taps, __uint_as_float(r[j]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn wg_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static int wg_encode_map(EncodeTiledFn enc, CUtensorMap* tm, const void* base, int C, int ld, int W, int H, int B,
                         int BW, int BH) {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * W * 2,
                                   static_cast<cuuint64_t>(ld) * W * H * 2};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(BW), static_cast<cuuint32_t>(BH), 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -300 - static_cast<int>(r);
}

template <int BN>
static int wg_launch(const CUtensorMap& tmDY, const CUtensorMap& tmX, WgradParams& p, int num_sms, cudaStream_t stream) {
    constexpr int kStageBytes = 2 * kWgBox + (BN / 64) * kWgBox;
    int stages = (220 * 1024) / kStageBytes;
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    p.stages = stages;
    const int smem = stages * kStageBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
    static int configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = smem;
    }
    p.ci_tiles = p.Cin / BN;
    p.n_units = ((p.Cout + 127) / 128) * p.ci_tiles * p.taps;
    int splits = num_sms / p.n_units;
    if (splits < 1) splits = 1;
    if (splits > p.total_blocks) splits = p.total_blocks;
    p.n_splits = splits;
    conv_wgrad_kernel<BN><<<p.n_units * splits, kWgThreads, smem, stream>>>(tmDY, tmX, p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace b200

extern "C" int b200_conv_wgrad(const void* dy, int dy_ld, const void* x, int x_ld, float* dw, int B, int H, int W,
                               int Cin, int Cout, int taps, void* stream) {
    using namespace b200;
    if (dy == nullptr || x == nullptr || dw == nullptr || B <= 0 || H <= 0 || W <= 0) return -1;
    if (Cin % 64 != 0 || Cout % 64 != 0 || (taps != 1 && taps != 9)) return -2;
    if (dy_ld % 8 != 0 || x_ld % 8 != 0 || dy_ld < Cout || x_ld < Cin) return -3;
    if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(x) & 15)) return -5;
    WgradParams p{};
    p.BW = W < kWgKP ? W : kWgKP;
    if (kWgKP % p.BW != 0 || W % p.BW != 0) return -7;
    p.BH = kWgKP / p.BW;
    if (H % p.BH != 0) return -7;
    p.H = H;
    p.W = W;
    p.Cin = Cin;
    p.Cout = Cout;
    p.taps = taps;
    p.blocks_w = W / p.BW;
    p.blocks_per_case = p.blocks_w * (H / p.BH);
    p.total_blocks = B * p.blocks_per_case;
    p.dw = dw;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) return -9;
    }
    EncodeTiledFn enc = wg_encode_fn();
    if (enc == nullptr) return -8;
    CUtensorMap tmDY, tmX;
    int rc;
    if ((rc = wg_encode_map(enc, &tmDY, dy, Cout, dy_ld, W, H, B, p.BW, p.BH)) != 0) return rc - 1000;
    if ((rc = wg_encode_map(enc, &tmX, x, Cin, x_ld, W, H, B, p.BW, p.BH)) != 0) return rc - 2000;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (Cin % 256 == 0) return wg_launch<256>(tmDY, tmX, p, num_sms, s);
    if (Cin % 128 == 0) return wg_launch<128>(tmDY, tmX, p, num_sms, s);
    return wg_launch<64>(tmDY, tmX, p, num_sms, s);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight packing for the training step: fp32 master weights [Cout, Cin, kh, kw] (the nn.Conv2d layout) ->
//   w_fwd   bf16 [Cout, taps * Cin]   k = tap * Cin + ci        (forward implicit GEMM)
//   w_dgrad bf16 [Cin, taps * Cout]   k = tap' * Cout + co, tap' = taps - 1 - tap (the data gradient of a stride-1
//           'same' convolution is the convolution of dY with the flipped, transposed filter)
// ---------------------------------------------------------------------------------------------------------------
namespace b200 {
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                         __nv_bfloat16* __restrict__ w_fwd, __nv_bfloat16* __restrict__ w_dgrad) {
    const long long n = static_cast<long long>(Cout) * Cin * taps;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % taps);
        const int ci = static_cast<int>((i / taps) % Cin);
        const int co = static_cast<int>(i / (static_cast<long long>(taps) * Cin));
        const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
        if (w_fwd != nullptr) w_fwd[(static_cast<long long>(co) * taps + tap) * Cin + ci] = v;
        if (w_dgrad != nullptr) w_dgrad[(static_cast<long long>(ci) * taps + (taps - 1 - tap)) * Cout + co] = v;
    }
}
}  // namespace b200

extern "C" int b200_pack_conv_weights(const float* w, int Cout, int Cin, int taps, void* w_fwd, void* w_dgrad,
                                      void* stream) {
    using namespace b200;
    if (w == nullptr || Cout <= 0 || Cin <= 0 || taps <= 0) return -1;
    const long long n = static_cast<long long>(Cout) * Cin * taps;
    const int grid = static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    pack_conv_weights_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, Cout, Cin, taps, static_cast<__nv_bfloat16*>(w_fwd), static_cast<__nv_bfloat16*>(w_dgrad));
    return launch_status();
}
