// Fusion-head fine-tuning step (BASELINE config C5, frozen-encoder phase): forward + backward of the logits path of
// FusionModel.forward (code/model_module.py:919-1000) under the classification objective of
// LightningFusionModel._shared_step (code/train_fusion.py:238-242: LabelSmoothing -> Soft(Weighted)FocalLoss,
// code/loss.py:133-213), and the AdamW update (code/selector_helpers.py:222-229).
//
// The logits depend on the encoder maps only through 4x4-pooled tokens: proj_in_* are bias-free 1x1 convolutions
// (they commute with average pooling), GAP(p) is the mean of the tokens, and GAP(bilinear_up(lowres)) is a fixed
// linear map of the low-resolution tokens (`up_coef`).  So the step pools f3 once (fusion_tokens, the only pass
// over the maps) and everything after it is fp32 work on [B*T, C] token matrices:
//
//   sgemm            C = op(A) op(B) (+bias, +broadcast residual, GELU, split-K accumulate) - every Linear forward,
//                    data gradient and weight gradient of the head (fp32 SIMT: these are 0.2 % of the step's FLOPs
//                    and the gradient parity is held to 1e-4 against the fp32 oracle)
//   colsum           bias gradients
//   mha_fwd / mha_bwd   per (case, head) softmax attention on T <= 32 tokens, probabilities saved
//   ln_fwd / ln_bwd  token LayerNorm of attn_ffn
//   gelu_bwd         dH = dG * gelu'(H)
//   head_loss        per case: gating softmax, pooled fused vector, SE gate, classifier, smoothed focal loss and the
//                    whole backward of that chain down to the token gradients
//   adamw            torch.optim.AdamW on one flat parameter buffer
#include "b200_fusion.h"
#include "common.cuh"

namespace b200 {

// --------------------------------------------------------------------------------------------------------------
// sgemm
// --------------------------------------------------------------------------------------------------------------
struct SgemmArgs {
    const float* A;
    const float* B;
    float* C;
    float* pre;
    const float* bias;
    const float* res;
    long long lda, ldb, ldc, ldres;
    int M, N, K, res_div, act, beta, kchunk;
};

constexpr int SG_BK = 16;

// Tile (16 TM) x (16 TN), 256 threads, each a TM x TN register block (TN = 8: two 4-wide column groups 64 apart, so a
// warp's B reads stay one contiguous 256-byte segment); (TM, TN) = (4, 4) for small / split-K problems, (8, 8) for the
// token-matrix products.
template <int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmArgs a) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    __shared__ __align__(16) float As[SG_BK][BM + 4];
    __shared__ __align__(16) float Bs[SG_BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * a.kchunk, kend = min(a.K, kbeg + a.kchunk);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // global -> register prefetch of the next k-slab while the current one is consumed from shared memory
    constexpr int NA = BM * SG_BK / 256, NB = BN * SG_BK / 256;
    float ra[NA], rb[NB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int idx = tid + i * 256;
            const int m = TA ? idx % BM : idx / SG_BK, k = TA ? idx / BM : idx % SG_BK;
            const int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < a.M && gk < kend) ? (TA ? __ldg(a.A + gk * a.lda + gm) : __ldg(a.A + gm * a.lda + gk)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int idx = tid + i * 256;
            const int n = TB ? idx / SG_BK : idx % BN, k = TB ? idx % SG_BK : idx / BN;
            const int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < a.N && gk < kend) ? (TB ? __ldg(a.B + gn * a.ldb + gk) : __ldg(a.B + gk * a.ldb + gn)) : 0.f;
        }
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += SG_BK) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int idx = tid + i * 256;
            const int m = TA ? idx % BM : idx / SG_BK, k = TA ? idx / BM : idx % SG_BK;
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int idx = tid + i * 256;
            const int n = TB ? idx / SG_BK : idx % BN, k = TB ? idx % SG_BK : idx / BN;
            Bs[k][n] = rb[i];
        }
        __syncthreads();
        if (k0 + SG_BK < kend) fetch(k0 + SG_BK);
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                av[i] = t.x, av[i + 1] = t.y, av[i + 2] = t.z, av[i + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&Bs[k][(j >> 2) * 64 + tx * 4]);
                bv[j] = t.x, bv[j + 1] = t.y, bv[j + 2] = t.z, bv[j + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + ty * TM + i;
        if (row >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
            if (col >= a.N) continue;
            float v = acc[i][j];
            float* dst = a.C + row * a.ldc + col;
            if (split) {
                atomicAdd(dst, v);
                continue;
            }
            if (a.bias != nullptr) v += a.bias[col];
            if (a.res != nullptr) v += a.res[(row / a.res_div) * a.ldres + col];
            if (a.pre != nullptr) a.pre[row * a.ldc + col] = v;
            if (a.act == 1) v = gelu_exact(v);
            if (a.beta) v += *dst;
            *dst = v;
        }
    }
}

// out[n] += sum_r X[r][n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ld, int R, int N,
                                                     float* __restrict__ out) {
    __shared__ float part[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (col < N)
        for (int r = blockIdx.y * 8 + ry; r < R; r += gridDim.y * 8) s += X[r * ld + col];
    part[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][cx];
        atomicAdd(out + col, t);
    }
}

// --------------------------------------------------------------------------------------------------------------
// multi-head attention on a few tokens (nn.MultiheadAttention, batch_first, no dropout, no mask)
// --------------------------------------------------------------------------------------------------------------
// q [B*Tq, ldq] (head h at columns h*DH), k / v [B*Tk, ldkv]; probs [B, NH, Tq, Tk]; ctx [B*Tq, ldc]
__global__ void __launch_bounds__(128)
mha_fwd_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, const float* __restrict__ v,
               long long ldkv, int Tq, int Tk, int DH, float* __restrict__ probs, float* __restrict__ ctx,
               long long ldc) {
    extern __shared__ float sm[];
    float* Q = sm;                 // [Tq][DH+1]
    float* K = Q + Tq * (DH + 1);  // [Tk][DH+1]
    float* V = K + Tk * (DH + 1);  // [Tk][DH+1]
    float* S = V + Tk * (DH + 1);  // [Tq][Tk]
    const int b = blockIdx.x, h = blockIdx.y, NH = gridDim.y, tid = threadIdx.x;
    for (int i = tid; i < Tq * DH; i += blockDim.x) {
        const int t = i / DH, d = i % DH;
        Q[t * (DH + 1) + d] = q[(static_cast<long long>(b) * Tq + t) * ldq + h * DH + d];
    }
    for (int i = tid; i < Tk * DH; i += blockDim.x) {
        const int t = i / DH, d = i % DH;
        const long long off = (static_cast<long long>(b) * Tk + t) * ldkv + h * DH + d;
        K[t * (DH + 1) + d] = k[off];
        V[t * (DH + 1) + d] = v[off];
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(DH));
    for (int i = tid; i < Tq * Tk; i += blockDim.x) {
        const int tq = i / Tk, tk = i % Tk;
        float acc = 0.f;
        for (int d = 0; d < DH; ++d) acc = fmaf(Q[tq * (DH + 1) + d], K[tk * (DH + 1) + d], acc);
        S[i] = acc * scale;
    }
    __syncthreads();
    for (int r = tid; r < Tq; r += blockDim.x) {
        float* row = S + r * Tk;
        float mx = row[0];
        for (int j = 1; j < Tk; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
        for (int j = 0; j < Tk; ++j) {
            row[j] = expf(row[j] - mx);
            sum += row[j];
        }
        const float inv = 1.0f / sum;
        for (int j = 0; j < Tk; ++j) row[j] *= inv;
    }
    __syncthreads();
    float* pdst = probs + (static_cast<long long>(b) * NH + h) * Tq * Tk;
    for (int i = tid; i < Tq * Tk; i += blockDim.x) pdst[i] = S[i];
    for (int i = tid; i < Tq * DH; i += blockDim.x) {
        const int tq = i / DH, d = i % DH;
        float acc = 0.f;
        for (int tk = 0; tk < Tk; ++tk) acc = fmaf(S[tq * Tk + tk], V[tk * (DH + 1) + d], acc);
        ctx[(static_cast<long long>(b) * Tq + tq) * ldc + h * DH + d] = acc;
    }
}

__global__ void __launch_bounds__(128)
mha_bwd_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, const float* __restrict__ v,
               long long ldkv, const float* __restrict__ probs, const float* __restrict__ dctx, long long ldc, int Tq,
               int Tk, int DH, float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
    extern __shared__ float sm[];
    float* Q = sm;
    float* K = Q + Tq * (DH + 1);
    float* V = K + Tk * (DH + 1);
    float* dO = V + Tk * (DH + 1);   // [Tq][DH+1]
    float* P = dO + Tq * (DH + 1);   // [Tq][Tk]
    float* dS = P + Tq * Tk;         // [Tq][Tk]
    const int b = blockIdx.x, h = blockIdx.y, NH = gridDim.y, tid = threadIdx.x;
    for (int i = tid; i < Tq * DH; i += blockDim.x) {
        const int t = i / DH, d = i % DH;
        Q[t * (DH + 1) + d] = q[(static_cast<long long>(b) * Tq + t) * ldq + h * DH + d];
        dO[t * (DH + 1) + d] = dctx[(static_cast<long long>(b) * Tq + t) * ldc + h * DH + d];
    }
    for (int i = tid; i < Tk * DH; i += blockDim.x) {
        const int t = i / DH, d = i % DH;
        const long long off = (static_cast<long long>(b) * Tk + t) * ldkv + h * DH + d;
        K[t * (DH + 1) + d] = k[off];
        V[t * (DH + 1) + d] = v[off];
    }
    const float* psrc = probs + (static_cast<long long>(b) * NH + h) * Tq * Tk;
    for (int i = tid; i < Tq * Tk; i += blockDim.x) P[i] = psrc[i];
    __syncthreads();
    for (int i = tid; i < Tq * Tk; i += blockDim.x) {  // dP
        const int tq = i / Tk, tk = i % Tk;
        float acc = 0.f;
        for (int d = 0; d < DH; ++d) acc = fmaf(dO[tq * (DH + 1) + d], V[tk * (DH + 1) + d], acc);
        dS[i] = acc;
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(DH));
    for (int r = tid; r < Tq; r += blockDim.x) {  // softmax backward, 1/sqrt(DH) folded in
        float dot = 0.f;
        for (int j = 0; j < Tk; ++j) dot = fmaf(dS[r * Tk + j], P[r * Tk + j], dot);
        for (int j = 0; j < Tk; ++j) dS[r * Tk + j] = P[r * Tk + j] * (dS[r * Tk + j] - dot) * scale;
    }
    __syncthreads();
    for (int i = tid; i < Tq * DH; i += blockDim.x) {
        const int tq = i / DH, d = i % DH;
        float acc = 0.f;
        for (int tk = 0; tk < Tk; ++tk) acc = fmaf(dS[tq * Tk + tk], K[tk * (DH + 1) + d], acc);
        dq[(static_cast<long long>(b) * Tq + tq) * ldq + h * DH + d] = acc;
    }
    for (int i = tid; i < Tk * DH; i += blockDim.x) {
        const int tk = i / DH, d = i % DH;
        float ak = 0.f, av = 0.f;
        for (int tq = 0; tq < Tq; ++tq) {
            ak = fmaf(dS[tq * Tk + tk], Q[tq * (DH + 1) + d], ak);
            av = fmaf(P[tq * Tk + tk], dO[tq * (DH + 1) + d], av);
        }
        const long long off = (static_cast<long long>(b) * Tk + tk) * ldkv + h * DH + d;
        dk[off] = ak;
        dv[off] = av;
    }
}

// --------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension, one warp per row; statistics saved for the backward
// --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, int R, int C, const float* __restrict__ w, const float* __restrict__ bvec,
              float eps, float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const float* xr = x + static_cast<long long>(r) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / C;
    float s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float d = xr[c] - mean;
        s2 = fmaf(d, d, s2);
    }
    const float rstd = rsqrtf(warp_sum(s2) / C + eps);
    for (int c = lane; c < C; c += 32) y[static_cast<long long>(r) * C + c] = (xr[c] - mean) * rstd * w[c] + bvec[c];
    if (lane == 0) {
        mean_out[r] = mean;
        rstd_out[r] = rstd;
    }
}

// dx = dres + LN'(dy);  dyxhat = dy * xhat (its column sums are the weight gradient, dy's the bias gradient)
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ dres, int R, int C,
              const float* __restrict__ w, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              float* __restrict__ dx, float* __restrict__ dyxhat) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const long long base = static_cast<long long>(r) * C;
    const float mean = mean_in[r], rstd = rstd_in[r];
    float sg = 0.f, sgx = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float xh = (x[base + c] - mean) * rstd;
        const float g = dy[base + c] * w[c];
        sg += g;
        sgx = fmaf(g, xh, sgx);
    }
    sg = warp_sum(sg) / C;
    sgx = warp_sum(sgx) / C;
    for (int c = lane; c < C; c += 32) {
        const float xh = (x[base + c] - mean) * rstd;
        const float d = dy[base + c];
        float v = rstd * (d * w[c] - sg - xh * sgx);
        if (dres != nullptr) v += dres[base + c];
        dx[base + c] = v;
        dyxhat[base + c] = d * xh;
    }
}

__device__ __forceinline__ float gelu_grad(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

__global__ void gelu_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dg, long long n,
                                float* __restrict__ out) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = dg[i] * gelu_grad(pre[i]);
}

// --------------------------------------------------------------------------------------------------------------
// per-case head: gating, pooled fused vector, SE gate, classifier, loss and their backward
// --------------------------------------------------------------------------------------------------------------
constexpr int HEAD_MAX_K = 16;

__global__ void __launch_bounds__(128) head_loss_kernel(const b200_head_train a) {
    extern __shared__ float sm[];
    const int C = a.C, T = a.T, Cm = a.se_mid, K = a.num_classes;
    float* s_pd = sm;          // [C]
    float* s_pc = s_pd + C;    // [C]
    float* s_gf = s_pc + C;    // [C] pooled fused vector
    float* s_g = s_gf + C;     // [C] SE gate
    float* s_dgf = s_g + C;    // [C]
    float* s_da2 = s_dgf + C;  // [C]
    float* s_a1 = s_da2 + C;   // [Cm]
    float* s_h = s_a1 + Cm;    // [Cm]
    float* s_da1 = s_h + Cm;   // [Cm]
    float* s_dl = s_da1 + Cm;  // [HEAD_MAX_K] logits, then dlogits
    float* s_misc = s_dl + HEAD_MAX_K;  // [8]
    __shared__ double scratch[33];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const long long tokbase = static_cast<long long>(b) * T * C;
    const float invT = 1.0f / T;

    const bool given_pvec = a.pvec_dwi != nullptr && a.pvec_dce != nullptr;
    const float pscale = given_pvec ? a.pvec_scale : invT;  // factor on the pooled-vector gradients
    for (int c = tid; c < C; c += blockDim.x) {
        if (given_pvec) {
            s_pd[c] = a.pvec_dwi[static_cast<long long>(b) * C + c] * a.pvec_scale;
            s_pc[c] = a.pvec_dce[static_cast<long long>(b) * C + c] * a.pvec_scale;
            continue;
        }
        float sd = 0.f, sc = 0.f;
        for (int t = 0; t < T; ++t) {
            sd += a.tok_dwi[tokbase + t * C + c];
            sc += a.tok_dce[tokbase + t * C + c];
        }
        s_pd[c] = sd * invT;
        s_pc[c] = sc * invT;
    }
    const bool use_mask = a.use_mask_attention != 0;
    const int in_dim = 2 * C + (use_mask ? 2 : 0);
    float cd = 0.f, cc = 0.f;
    if (use_mask) {
        double dd = 0.0, dc = 0.0;
        for (int p = tid; p < a.npix_mask; p += blockDim.x) {
            dd += a.mask_dwi[static_cast<long long>(b) * a.npix_mask + p];
            dc += a.mask_dce[static_cast<long long>(b) * a.npix_mask + p];
        }
        cd = static_cast<float>(block_sum<double>(dd, scratch) / a.npix_mask);
        cc = static_cast<float>(block_sum<double>(dc, scratch) / a.npix_mask);
    }
    __syncthreads();
    // ---- gating (model_module.py:745-780) ----
    {
        float p0 = 0.f, p1 = 0.f;
        for (int i = tid; i < 2 * C; i += blockDim.x) {
            const float xv = i < C ? s_pd[i] : s_pc[i - C];
            p0 = fmaf(a.gate_w[i], xv, p0);
            p1 = fmaf(a.gate_w[in_dim + i], xv, p1);
            a.gx_out[static_cast<long long>(b) * in_dim + i] = xv;
        }
        const double g0 = block_sum<double>(p0, scratch);
        const double g1 = block_sum<double>(p1, scratch);
        if (tid == 0) {
            float l0 = static_cast<float>(g0) + a.gate_b[0], l1 = static_cast<float>(g1) + a.gate_b[1];
            if (use_mask) {
                l0 += a.gate_w[2 * C] * cd + a.gate_w[2 * C + 1] * cc;
                l1 += a.gate_w[in_dim + 2 * C] * cd + a.gate_w[in_dim + 2 * C + 1] * cc;
                a.gx_out[static_cast<long long>(b) * in_dim + 2 * C] = cd;
                a.gx_out[static_cast<long long>(b) * in_dim + 2 * C + 1] = cc;
            }
            const float mx = fmaxf(l0, l1);
            const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
            s_misc[0] = e0 / (e0 + e1);
            s_misc[1] = e1 / (e0 + e1);
            if (a.gating_out != nullptr) {
                a.gating_out[b * 2] = s_misc[0];
                a.gating_out[b * 2 + 1] = s_misc[1];
            }
        }
    }
    __syncthreads();
    const float al0 = s_misc[0], al1 = s_misc[1];
    // ---- pooled fused vector (:958, :962-974 under GAP) ----
    for (int c = tid; c < C; c += blockDim.x) {
        float g = al0 * s_pd[c] + al1 * s_pc[c];
        if (a.lowres != nullptr)
            for (int t = 0; t < T; ++t) g = fmaf(a.up_coef[t], a.lowres[tokbase + t * C + c], g);
        s_gf[c] = g;
        a.gf_out[static_cast<long long>(b) * C + c] = g;
    }
    __syncthreads();
    // ---- SE gate (SEBlock, model_module.py:25-43) ----
    if (a.use_se) {
        for (int m = warp; m < Cm; m += nwarps) {
            float acc = 0.f;
            for (int c = lane; c < C; c += 32) acc = fmaf(a.se_w1[m * C + c], s_gf[c], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const float pre = acc + a.se_b1[m];
                s_a1[m] = pre;
                s_h[m] = gelu_exact(pre);
                a.h_out[static_cast<long long>(b) * Cm + m] = s_h[m];
            }
        }
        __syncthreads();
        for (int c = tid; c < C; c += blockDim.x) {
            float acc = a.se_b2[c];
            for (int m = 0; m < Cm; ++m) acc = fmaf(a.se_w2[c * Cm + m], s_h[m], acc);
            s_g[c] = sigmoidf_(acc);
        }
    } else {
        for (int c = tid; c < C; c += blockDim.x) s_g[c] = 1.f;
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        a.z_out[static_cast<long long>(b) * C + c] = s_gf[c] * s_g[c];
        if (a.gate_out != nullptr) a.gate_out[static_cast<long long>(b) * C + c] = s_g[c];
        if (a.u_out != nullptr && a.mask_v != nullptr) a.u_out[static_cast<long long>(b) * C + c] = a.mask_v[c] * s_g[c];
    }
    // ---- classifier (:895-899) ----
    for (int k = warp; k < K; k += nwarps) {
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) acc = fmaf(a.cls_w[k * C + c], s_gf[c] * s_g[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_dl[k] = acc + a.cls_b[k];
    }
    __syncthreads();
    if (a.forward_only) {
        if (a.logits_out != nullptr)
            for (int k = tid; k < K; k += blockDim.x) a.logits_out[static_cast<long long>(b) * K + k] = s_dl[k];
        return;
    }
    // ---- LabelSmoothing (loss.py:190-213) + Soft(Weighted)FocalLoss (loss.py:133-188), mean over the batch ----
    if (tid == 0) {
        const int label = static_cast<int>(a.labels[b]);
        float mx = s_dl[0];
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, s_dl[k]);
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(s_dl[k] - mx);
        const float lse = mx + logf(se);
        float G[HEAD_MAX_K], p[HEAD_MAX_K];
        float loss = 0.f, gsum = 0.f;
        for (int k = 0; k < K; ++k) {
            const float logit = s_dl[k];
            if (a.logits_out != nullptr) a.logits_out[static_cast<long long>(b) * K + k] = logit;
            const float lp = logit - lse;
            p[k] = expf(lp);
            const float y = (k == label) ? 1.0f - a.smoothing : a.smoothing / (K - 1);
            const float cw = a.class_weights != nullptr ? a.class_weights[k] : 1.0f;
            const float om = fmaxf(1.0f - p[k], 0.f);
            const float fw = powf(om, a.gamma);
            loss -= y * cw * fw * lp;
            // d/dlp of -y cw (1-p)^gamma lp, p = exp(lp)
            const float fwm1 = om > 0.f ? powf(om, a.gamma - 1.0f) : 0.f;
            G[k] = -y * cw * (fw - a.gamma * fwm1 * p[k] * lp);
            gsum += G[k];
        }
        for (int k = 0; k < K; ++k) {
            const float dl = (G[k] - p[k] * gsum) * a.loss_scale;
            s_dl[k] = dl;
            a.dlogits_out[static_cast<long long>(b) * K + k] = dl;
        }
        atomicAdd(a.loss_out, loss * a.loss_scale);
    }
    __syncthreads();
    // ---- backward: classifier, SE (+ the fused-mask dice term arriving through u = v * gate) ----
    const bool mask_term = a.mask_v != nullptr && a.mk_tmpd != nullptr;
    float e0 = 0.f, e1 = 0.f;  // gradient of the gating weights coming from the mask term
    for (int c = tid; c < C; c += blockDim.x) {
        float dz = 0.f;
        for (int k = 0; k < K; ++k) dz = fmaf(a.cls_w[k * C + c], s_dl[k], dz);
        const float g = s_g[c];
        s_dgf[c] = dz * g;
        float dg = dz * s_gf[c];
        if (mask_term) {
            const long long bc = static_cast<long long>(b) * C + c;
            const float v = a.mask_v[c], u = v * g, td = a.mk_tmpd[bc], tc = a.mk_tmpc[bc];
            float du = al0 * td + al1 * tc;
            if (a.lowres != nullptr)
                for (int t = 0; t < T; ++t) du = fmaf(a.mk_q[b * T + t], a.lowres[tokbase + t * C + c], du);
            dg = fmaf(du, v, dg);
            a.dug_out[bc] = du * g;
            a.aud_out[bc] = al0 * u;
            a.auc_out[bc] = al1 * u;
            e0 = fmaf(u, td, e0);
            e1 = fmaf(u, tc, e1);
        }
        const float da2 = a.use_se ? dg * g * (1.0f - g) : 0.f;
        s_da2[c] = da2;
        if (a.use_se) a.da2_out[static_cast<long long>(b) * C + c] = da2;
    }
    if (mask_term) {
        e0 = static_cast<float>(block_sum<double>(e0, scratch));
        e1 = static_cast<float>(block_sum<double>(e1, scratch));
    }
    __syncthreads();
    if (a.use_se) {
        for (int m = tid; m < Cm; m += blockDim.x) {
            float dh = 0.f;
            for (int c = 0; c < C; ++c) dh = fmaf(a.se_w2[c * Cm + m], s_da2[c], dh);
            const float da1 = dh * gelu_grad(s_a1[m]);
            s_da1[m] = da1;
            a.da1_out[static_cast<long long>(b) * Cm + m] = da1;
        }
        __syncthreads();
        for (int c = tid; c < C; c += blockDim.x) {
            float acc = s_dgf[c];
            for (int m = 0; m < Cm; ++m) acc = fmaf(a.se_w1[m * C + c], s_da1[m], acc);
            s_dgf[c] = acc;
        }
        __syncthreads();
    }
    // ---- backward: gating softmax and the pooled vectors ----
    float q0 = 0.f, q1 = 0.f;
    for (int c = tid; c < C; c += blockDim.x) {
        q0 = fmaf(s_dgf[c], s_pd[c], q0);
        q1 = fmaf(s_dgf[c], s_pc[c], q1);
    }
    const float da0 = static_cast<float>(block_sum<double>(q0, scratch)) + e0;
    const float da1g = static_cast<float>(block_sum<double>(q1, scratch)) + e1;
    const float dots = al0 * da0 + al1 * da1g;
    const float dgl0 = al0 * (da0 - dots), dgl1 = al1 * (da1g - dots);
    if (tid == 0) {
        a.dgl_out[b * 2] = dgl0;
        a.dgl_out[b * 2 + 1] = dgl1;
    }
    for (int c = tid; c < C; c += blockDim.x) {
        const float dgf = s_dgf[c];
        a.dpd_out[static_cast<long long>(b) * C + c] =
            (al0 * dgf + a.gate_w[c] * dgl0 + a.gate_w[in_dim + c] * dgl1) * pscale;
        a.dpc_out[static_cast<long long>(b) * C + c] =
            (al1 * dgf + a.gate_w[C + c] * dgl0 + a.gate_w[in_dim + C + c] * dgl1) * pscale;
        if (a.dlowres_out != nullptr) {
            const float u = mask_term ? a.mask_v[c] * s_g[c] : 0.f;
            for (int t = 0; t < T; ++t)
                a.dlowres_out[tokbase + t * C + c] =
                    a.up_coef[t] * dgf + (mask_term ? a.mk_q[b * T + t] * u : 0.f);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// fused-mask dice term (see b200_fusion.h)
// --------------------------------------------------------------------------------------------------------------
// D[b,p] = omega[b] . f3[b,p,:]; one warp per pixel, omega[b] in registers (8 * 32 * NJ channels)
template <int NJ>
__global__ void __launch_bounds__(256)
mask_dot_kernel(const __nv_bfloat16* __restrict__ f3, const float* __restrict__ omega, int npix, int Cin,
                float* __restrict__ D) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nvec = Cin >> 3;  // uint4 = 8 channels
    float w[NJ][8];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = (lane + 32 * j) * 8 + k;
            w[j][k] = (lane + 32 * j) < nvec ? omega[static_cast<long long>(b) * Cin + c] : 0.f;
        }
    // two pixels per trip (four cost occupancy: 77 registers, measured slower): loads of both rows precede the reductions
    constexpr int NP = 2;
    const int stride = gridDim.x * nwarps;
    for (int p = blockIdx.x * nwarps + warp; p < npix; p += NP * stride) {
        uint4 r[NP][NJ];
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const int q = min(p + u * stride, npix - 1);
            const uint4* row = reinterpret_cast<const uint4*>(f3 + (static_cast<long long>(b) * npix + q) * Cin);
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (lane + 32 * j < nvec) r[u][j] = __ldg(row + lane + 32 * j);
        }
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (lane + 32 * j < nvec) {
                    float f[8];
                    unpack_bf16x8(r[u][j], f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc = fmaf(w[j][k], f[k], acc);
                }
            acc = warp_sum(acc);
            const int q = p + u * stride;
            if (lane == 0 && q < npix) D[static_cast<long long>(b) * npix + q] = acc;
        }
    }
}

// s[b,c] = sum_p dm[b,p] f3[b,p,c]; grid (chunks, B): each CTA sums a pixel chunk, 8 channels per thread
__global__ void __launch_bounds__(128)
mask_wsum_kernel(const __nv_bfloat16* __restrict__ f3, const float* __restrict__ dm, int npix, int Cin,
                 float* __restrict__ s) {
    const int b = blockIdx.y, nvec = Cin >> 3;
    const int per = (npix + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
    for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
        for (int p = p0; p < p1; ++p) {
            const float d = __ldg(dm + static_cast<long long>(b) * npix + p);
            float f[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(f3 + (static_cast<long long>(b) * npix + p) * Cin) + vi), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(d, f[k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(s + static_cast<long long>(b) * Cin + vi * 8 + k, acc[k]);
    }
}

// Same sum for Cin / 8 <= 256: a thread owns 8 channels of one pixel slice (blockDim = nvec * slices, no idle lanes),
// 8 independent 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256)
mask_wsum_vec_kernel(const __nv_bfloat16* __restrict__ f3, const float* __restrict__ dm, int npix, int Cin,
                     float* __restrict__ s) {
    const int b = blockIdx.y, nvec = Cin >> 3;
    const int vi = threadIdx.x % nvec, pg = threadIdx.x / nvec, npg = blockDim.x / nvec;
    const int per = (npix + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
    const uint4* base = reinterpret_cast<const uint4*>(f3 + static_cast<long long>(b) * npix * Cin) + vi;
    const float* d = dm + static_cast<long long>(b) * npix;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int p = p0 + pg; p < p1; p += 8 * npg) {
        uint4 r[8];
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int q = p + u * npg;
            if (q < p1) {
                r[u] = __ldg(base + static_cast<long long>(q) * nvec);
                w[u] = __ldg(d + q);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (p + u * npg < p1) {
                float f[8];
                unpack_bf16x8(r[u], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[u], f[k], acc[k]);
            }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(s + static_cast<long long>(b) * Cin + vi * 8 + k, acc[k]);
}

__device__ __forceinline__ float bce_logits(float m, float t) {  // F.binary_cross_entropy_with_logits, one element
    return fmaxf(m, 0.f) - m * t + log1pf(expf(-fabsf(m)));
}

// One case's mask loss: loss_type 0 = SoftDiceLoss (loss.py:45-62: eps in numerator and denominator),
// 1 = DiceBCELoss(1, 1) (loss.py:11-43: mean BCE-with-logits + dice with eps in the denominator only).
__device__ __forceinline__ float mask_loss_of(const float* __restrict__ logits, const float* __restrict__ target,
                                              int npix, float eps, int loss_type, double* scratch) {
    float si = 0.f, sp = 0.f, st = 0.f, sb = 0.f;
    for (int p = threadIdx.x; p < npix; p += blockDim.x) {
        const float m = logits[p], pr = sigmoidf_(m), t = target[p];
        si = fmaf(pr, t, si);
        sp += pr;
        st += t;
        if (loss_type == 1) sb += bce_logits(m, t);
    }
    const double I = block_sum<double>(si, scratch), P = block_sum<double>(sp, scratch),
                 Tt = block_sum<double>(st, scratch), Bc = block_sum<double>(sb, scratch);
    const double num = 2.0 * I + (loss_type == 0 ? eps : 0.0);
    return static_cast<float>(1.0 - num / (P + Tt + eps) + Bc / npix);
}

// bilinear source taps of F.interpolate(align_corners=False): n_in -> n_out along one axis
__device__ __forceinline__ void bilinear_taps(int o, int n_in, float ratio, int& i0, int& i1, float& lam) {
    const float s = fmaxf((o + 0.5f) * ratio - 0.5f, 0.f);
    i0 = min(static_cast<int>(s), n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    lam = s - i0;
}

// One CTA per case.  Map resolution H x W (D_*, dm_out), mask resolution Ho x Wo (target, encoder masks, m_out).
// Ho x Wo != H x W is MaskHeadResize's interpolation dispatch (code/model_module.py:197-211): pre -> bilinear resize ->
// out, still linear, = bilinear resize of the map-resolution logit; its gradient returns through the transposed resize.
__global__ void __launch_bounds__(256)
mask_dice_kernel(const float* __restrict__ D_dwi, const float* __restrict__ D_dce, const float* __restrict__ gating,
                 const float* __restrict__ u, const float* __restrict__ lowres, const float* __restrict__ pre_b,
                 const float* __restrict__ out_w, const float* __restrict__ out_b, int mid,
                 const float* __restrict__ target, const float* __restrict__ enc_dwi,
                 const float* __restrict__ enc_dce, int H, int W, int Ho, int Wo, int Hp, int Wp, int C, float scale,
                 float eps, int loss_type, float* __restrict__ m_out, float* __restrict__ dm_out,
                 float* __restrict__ q_out, float* __restrict__ dc0_out, float* __restrict__ loss_out) {
    extern __shared__ float sm[];
    const int npix = H * W, npo = Ho * Wo, T = Hp * Wp;
    const bool same = (Ho == H && Wo == W);
    float* s_mh = sm;            // [npix] map-resolution logits
    float* s_dmh = s_mh + npix;  // [npix] gradient at the map-resolution logits
    float* s_p = s_dmh + npix;   // [npo] probabilities at the mask resolution
    float* s_r = s_p + npo;      // [T] u . lowres[t]
    float* s_q = s_r + T;        // [T]
    __shared__ double scratch[33];
    __shared__ float s_c0;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const long long pb = static_cast<long long>(b) * npix, po = static_cast<long long>(b) * npo;
    for (int t = warp; t < T; t += nwarps) {
        float acc = 0.f;
        if (lowres != nullptr)
            for (int c = lane; c < C; c += 32)
                acc = fmaf(u[static_cast<long long>(b) * C + c], lowres[(static_cast<long long>(b) * T + t) * C + c], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            s_r[t] = acc;
            s_q[t] = 0.f;
        }
    }
    if (warp == 0) {
        float acc = 0.f;
        for (int i = lane; i < mid; i += 32) acc = fmaf(out_w[i], pre_b[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_c0 = acc + out_b[0];
    }
    __syncthreads();
    const float a0 = gating[b * 2], a1 = gating[b * 2 + 1], c0 = s_c0;
    const float sh = static_cast<float>(Hp) / H, sw = static_cast<float>(Wp) / W;
    // ---- logits at the map resolution: the cross-attention tokens are up-sampled bilinearly (model_module.py:972) ----
    for (int p = tid; p < npix; p += blockDim.x) {
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_taps(p / W, Hp, sh, y0, y1, ly);
        bilinear_taps(p % W, Wp, sw, x0, x1, lx);
        const float up = (1.f - ly) * ((1.f - lx) * s_r[y0 * Wp + x0] + lx * s_r[y0 * Wp + x1]) +
                         ly * ((1.f - lx) * s_r[y1 * Wp + x0] + lx * s_r[y1 * Wp + x1]);
        s_mh[p] = c0 + a0 * D_dwi[pb + p] + a1 * D_dce[pb + p] + up;
        s_dmh[p] = 0.f;
    }
    __syncthreads();
    // ---- logits and loss sums at the mask resolution ----
    const float rh = static_cast<float>(H) / Ho, rw = static_cast<float>(W) / Wo;
    float si = 0.f, sp = 0.f, st = 0.f, sb = 0.f;
    for (int o = tid; o < npo; o += blockDim.x) {
        float m;
        if (same) {
            m = s_mh[o];
        } else {
            int y0, y1, x0, x1;
            float ly, lx;
            bilinear_taps(o / Wo, H, rh, y0, y1, ly);
            bilinear_taps(o % Wo, W, rw, x0, x1, lx);
            m = (1.f - ly) * ((1.f - lx) * s_mh[y0 * W + x0] + lx * s_mh[y0 * W + x1]) +
                ly * ((1.f - lx) * s_mh[y1 * W + x0] + lx * s_mh[y1 * W + x1]);
        }
        m_out[po + o] = m;
        const float pr = sigmoidf_(m), t = target[po + o];
        s_p[o] = pr;
        si = fmaf(pr, t, si);
        sp += pr;
        st += t;
        if (loss_type == 1) sb += bce_logits(m, t);
    }
    const double I = block_sum<double>(si, scratch), P = block_sum<double>(sp, scratch),
                 Tt = block_sum<double>(st, scratch), Bc = block_sum<double>(sb, scratch);
    const double S = P + Tt + eps, num = 2.0 * I + (loss_type == 0 ? eps : 0.0);
    const float own = static_cast<float>(1.0 - num / S + Bc / npo);  // this case's fused-mask loss
    const float bce_g = loss_type == 1 ? scale / npo : 0.f;
    // ---- gradient at the mask resolution, carried back to the map resolution (transposed resize) ----
    for (int o = tid; o < npo; o += blockDim.x) {
        const float pr = s_p[o], t = target[po + o];
        // loss = scale * (1 - dice): d/dp = -scale * (2 t S - num) / S^2, then through the sigmoid
        const float dm = -scale * static_cast<float>((2.0 * t * S - num) / (S * S)) * pr * (1.f - pr) + bce_g * (pr - t);
        if (same) {
            s_dmh[o] = dm;
        } else {
            int y0, y1, x0, x1;
            float ly, lx;
            bilinear_taps(o / Wo, H, rh, y0, y1, ly);
            bilinear_taps(o % Wo, W, rw, x0, x1, lx);
            atomicAdd(&s_dmh[y0 * W + x0], dm * (1.f - ly) * (1.f - lx));
            atomicAdd(&s_dmh[y0 * W + x1], dm * (1.f - ly) * lx);
            atomicAdd(&s_dmh[y1 * W + x0], dm * ly * (1.f - lx));
            atomicAdd(&s_dmh[y1 * W + x1], dm * ly * lx);
        }
    }
    __syncthreads();
    float sdm = 0.f;
    for (int p = tid; p < npix; p += blockDim.x) {
        const float dm = s_dmh[p];
        dm_out[pb + p] = dm;
        sdm += dm;
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_taps(p / W, Hp, sh, y0, y1, ly);
        bilinear_taps(p % W, Wp, sw, x0, x1, lx);
        atomicAdd(&s_q[y0 * Wp + x0], dm * (1.f - ly) * (1.f - lx));
        atomicAdd(&s_q[y0 * Wp + x1], dm * (1.f - ly) * lx);
        atomicAdd(&s_q[y1 * Wp + x0], dm * ly * (1.f - lx));
        atomicAdd(&s_q[y1 * Wp + x1], dm * ly * lx);
    }
    const float tot = static_cast<float>(block_sum<double>(sdm, scratch));
    float extra = 0.f;  // the encoder masks' loss terms: constants for the head, part of the reported loss
    if (enc_dwi != nullptr) extra += mask_loss_of(enc_dwi + po, target + po, npo, eps, loss_type, scratch);
    if (enc_dce != nullptr) extra += mask_loss_of(enc_dce + po, target + po, npo, eps, loss_type, scratch);
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) q_out[b * T + t] = s_q[t];
    if (tid == 0) {
        atomicAdd(dc0_out, tot);
        atomicAdd(loss_out, scale * (own + extra));
    }
}

// v = pre_w^T out_w, c0 = out_w . pre_b + out_b:  dpre_w = out_w (x) dv, dout_w = pre_w dv + pre_b dc0,
// dpre_b = out_w dc0, dout_b = dc0.  One CTA.
__global__ void __launch_bounds__(256)
mask_head_grads_kernel(const float* __restrict__ dv, const float* __restrict__ dc0, const float* __restrict__ pre_w,
                       const float* __restrict__ pre_b, const float* __restrict__ out_w, int mid, int C,
                       float* __restrict__ g_pre_w, float* __restrict__ g_pre_b, float* __restrict__ g_out_w,
                       float* __restrict__ g_out_b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float d0 = dc0[0];
    for (int i = tid; i < mid * C; i += blockDim.x) g_pre_w[i] += out_w[i / C] * dv[i % C];
    for (int i = warp; i < mid; i += nwarps) {
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) acc = fmaf(pre_w[i * C + c], dv[c], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            g_out_w[i] += acc + pre_b[i] * d0;
            g_pre_b[i] += out_w[i] * d0;
        }
    }
    if (tid == 0) g_out_b[0] += d0;
}

// torch.optim.AdamW (decoupled weight decay, no amsgrad) on a flat buffer; g is scaled by grad_scale first
// (1 / world_size after a summing all-reduce).
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2_sqrt, float grad_scale) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * grad_scale;
        float pi = p[i] * (1.0f - lr * wd);
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

// The same update with PER-ELEMENT learning rate, weight decay and first step (parameter groups of
// LightningFusionOptimizerFactory, code/selector_helpers.py:456-518: discriminative learning rates / regularisation by
// depth; groups that join the optimiser later - gradual unfreezing - start their own bias-correction count).
__global__ void adamw_groups_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, long long n, const float* __restrict__ lr_vec,
                                    const float* __restrict__ wd_vec, const int* __restrict__ step0, float lr_mult,
                                    float beta1, float beta2, float eps, int step, float grad_scale) {
    const float l1 = logf(beta1), l2 = logf(beta2);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float lr = lr_vec[i] * lr_mult;
        if (lr == 0.f) continue;  // alignment padding / parameters parked outside every group
        const float k = static_cast<float>(step - (step0 != nullptr ? step0[i] : 0));
        const float bc1 = 1.0f - expf(k * l1), bc2_sqrt = sqrtf(1.0f - expf(k * l2));
        const float gi = g[i] * grad_scale;
        float pi = p[i] * (1.0f - lr * wd_vec[i]);
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        pi -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        p[i] = pi;
    }
}

template <int TM, int TN>
static int launch_sgemm(const SgemmArgs& a, int ta, int tb, dim3 grid, cudaStream_t s) {
    if (ta && tb)
        sgemm_kernel<TM, TN, true, true><<<grid, 256, 0, s>>>(a);
    else if (ta)
        sgemm_kernel<TM, TN, true, false><<<grid, 256, 0, s>>>(a);
    else if (tb)
        sgemm_kernel<TM, TN, false, true><<<grid, 256, 0, s>>>(a);
    else
        sgemm_kernel<TM, TN, false, false><<<grid, 256, 0, s>>>(a);
    return launch_status();
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sgemm(const float* A, long long lda, int trans_a, const float* B, long long ldb, int trans_b,
                          float* C, long long ldc, int M, int N, int K, const float* bias, const float* res,
                          long long ldres, int res_div, float* pre, int act, int beta, int split_k, void* stream) {
    if (M < 0 || N < 0 || K < 0 || split_k < 1) return -1;
    if (M == 0 || N == 0) return 0;
    if (A == nullptr || B == nullptr || C == nullptr) return -2;
    if (split_k > 1 && (bias != nullptr || res != nullptr || pre != nullptr || act != 0 || beta != 1)) return -3;
    if (act != 0 && act != 1) return -4;
    if (res != nullptr && res_div < 1) return -5;
    SgemmArgs a;
    a.A = A, a.B = B, a.C = C, a.pre = pre, a.bias = bias, a.res = res;
    a.lda = lda, a.ldb = ldb, a.ldc = ldc, a.ldres = ldres;
    a.M = M, a.N = N, a.K = K, a.res_div = res_div < 1 ? 1 : res_div, a.act = act, a.beta = beta;
    int kchunk = (K + split_k - 1) / split_k;
    kchunk = (kchunk + SG_BK - 1) / SG_BK * SG_BK;
    if (kchunk == 0) kchunk = SG_BK;
    a.kchunk = kchunk;
    const int splits = K == 0 ? 1 : (K + kchunk - 1) / kchunk;
    if (splits > 1 && split_k == 1) return -6;
    // 64 x 64 tiles (3-4 CTAs per SM hide the per-slab barrier) unless 128 x 128 tiles (8 x 8 register blocks, one
    // 160-register CTA per SM) still give every SM several tiles.  Measured on the head's [16384, 128..512] products:
    // 128 one-per-SM tiles run at 15 TFLOP/s, 512 small tiles at 18.6 - so the large tile is for larger problems only.
    const long long big_tiles = static_cast<long long>((M + 127) / 128) * ((N + 127) / 128);
    const bool big = splits == 1 && N >= 128 && big_tiles >= 4 * 148;
    const int bm = big ? 128 : 64, bn = big ? 128 : 64;
    dim3 grid((N + bn - 1) / bn, (M + bm - 1) / bm, splits);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return big ? launch_sgemm<8, 8>(a, trans_a, trans_b, grid, s) : launch_sgemm<4, 4>(a, trans_a, trans_b, grid, s);
}

extern "C" int b200_colsum(const float* X, long long ld, int R, int N, float* out, void* stream) {
    if (R < 0 || N < 0) return -1;
    if (R == 0 || N == 0) return 0;
    if (X == nullptr || out == nullptr) return -2;
    int gy = (R + 63) / 64;
    if (gy > 64) gy = 64;
    dim3 grid((N + 31) / 32, gy);
    colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, ld, R, N, out);
    return launch_status();
}

static int mha_smem(int Tq, int Tk, int DH, bool bwd) {
    const int qkv = (Tq + 2 * Tk) * (DH + 1);
    return static_cast<int>(sizeof(float)) * (bwd ? qkv + Tq * (DH + 1) + 2 * Tq * Tk : qkv + Tq * Tk);
}

extern "C" int b200_mha_fwd(const float* q, long long ldq, const float* k, const float* v, long long ldkv, int B,
                            int heads, int Tq, int Tk, int DH, float* probs, float* ctx, long long ldc, void* stream) {
    if (B < 0 || heads <= 0 || Tq <= 0 || Tk <= 0 || DH <= 0 || Tq > 32 || Tk > 32 || DH > 128) return -1;
    if (B == 0) return 0;
    if (q == nullptr || k == nullptr || v == nullptr || probs == nullptr || ctx == nullptr) return -2;
    const int smem = mha_smem(Tq, Tk, DH, false);
    if (smem > 48 * 1024) return -3;
    mha_fwd_kernel<<<dim3(B, heads), 128, smem, static_cast<cudaStream_t>(stream)>>>(q, ldq, k, v, ldkv, Tq, Tk, DH,
                                                                                     probs, ctx, ldc);
    return launch_status();
}

extern "C" int b200_mha_bwd(const float* q, long long ldq, const float* k, const float* v, long long ldkv,
                            const float* probs, const float* dctx, long long ldc, int B, int heads, int Tq, int Tk,
                            int DH, float* dq, float* dk, float* dv, void* stream) {
    if (B < 0 || heads <= 0 || Tq <= 0 || Tk <= 0 || DH <= 0 || Tq > 32 || Tk > 32 || DH > 128) return -1;
    if (B == 0) return 0;
    if (q == nullptr || k == nullptr || v == nullptr || probs == nullptr || dctx == nullptr || dq == nullptr ||
        dk == nullptr || dv == nullptr)
        return -2;
    const int smem = mha_smem(Tq, Tk, DH, true);
    static int configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = smem;
    }
    mha_bwd_kernel<<<dim3(B, heads), 128, smem, static_cast<cudaStream_t>(stream)>>>(q, ldq, k, v, ldkv, probs, dctx,
                                                                                     ldc, Tq, Tk, DH, dq, dk, dv);
    return launch_status();
}

extern "C" int b200_ln_fwd(const float* x, int R, int C, const float* w, const float* b, float eps, float* y,
                           float* mean, float* rstd, void* stream) {
    if (R < 0 || C <= 0) return -1;
    if (R == 0) return 0;
    if (x == nullptr || w == nullptr || b == nullptr || y == nullptr || mean == nullptr || rstd == nullptr) return -2;
    ln_fwd_kernel<<<(R + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, R, C, w, b, eps, y, mean, rstd);
    return launch_status();
}

extern "C" int b200_ln_bwd(const float* x, const float* dy, const float* dres, int R, int C, const float* w,
                           const float* mean, const float* rstd, float* dx, float* dyxhat, void* stream) {
    if (R < 0 || C <= 0) return -1;
    if (R == 0) return 0;
    if (x == nullptr || dy == nullptr || w == nullptr || mean == nullptr || rstd == nullptr || dx == nullptr ||
        dyxhat == nullptr)
        return -2;
    ln_bwd_kernel<<<(R + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, dres, R, C, w, mean, rstd, dx,
                                                                              dyxhat);
    return launch_status();
}

extern "C" int b200_gelu_bwd(const float* pre, const float* dg, long long n, float* out, void* stream) {
    if (n < 0) return -1;
    if (n == 0) return 0;
    if (pre == nullptr || dg == nullptr || out == nullptr) return -2;
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    gelu_bwd_kernel<<<static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(pre, dg, n, out);
    return launch_status();
}

extern "C" int b200_head_loss(const b200_head_train* args, int B, void* stream) {
    if (args == nullptr || B < 0) return -1;
    if (B == 0) return 0;
    const b200_head_train& a = *args;
    if (a.C <= 0 || a.T <= 0 || a.num_classes < 2 || a.num_classes > HEAD_MAX_K) return -2;
    if (a.use_se && a.se_mid <= 0) return -2;
    if (a.tok_dwi == nullptr || a.tok_dce == nullptr || a.gate_w == nullptr || a.gate_b == nullptr ||
        a.cls_w == nullptr || a.cls_b == nullptr || a.z_out == nullptr || a.gf_out == nullptr || a.gx_out == nullptr)
        return -3;
    if (!a.forward_only && (a.labels == nullptr || a.loss_out == nullptr || a.dlogits_out == nullptr ||
                            a.dgl_out == nullptr || a.dpd_out == nullptr || a.dpc_out == nullptr))
        return -3;
    if (a.mk_tmpd != nullptr && (a.mask_v == nullptr || a.mk_tmpc == nullptr || a.mk_q == nullptr ||
                                 a.dug_out == nullptr || a.aud_out == nullptr || a.auc_out == nullptr))
        return -8;
    if (a.use_mask_attention && (a.mask_dwi == nullptr || a.mask_dce == nullptr || a.npix_mask <= 0)) return -4;
    if (a.use_se && (a.se_w1 == nullptr || a.se_b1 == nullptr || a.se_w2 == nullptr || a.se_b2 == nullptr ||
                     a.h_out == nullptr || a.da1_out == nullptr || a.da2_out == nullptr))
        return -5;
    if ((a.lowres != nullptr) != (a.dlowres_out != nullptr)) return -6;
    if (a.lowres != nullptr && a.up_coef == nullptr) return -6;
    const size_t smem = (static_cast<size_t>(6) * a.C + 3 * a.se_mid + HEAD_MAX_K + 8) * sizeof(float);
    if (smem > 48 * 1024) return -7;
    head_loss_kernel<<<B, 128, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return launch_status();
}

extern "C" int b200_mask_dot(const void* f3, const float* omega, int B, int npix, int Cin, float* D, void* stream) {
    if (B < 0 || npix <= 0 || Cin <= 0 || Cin % 8 != 0 || Cin > 1024) return -1;
    if (B == 0) return 0;
    if (f3 == nullptr || omega == nullptr || D == nullptr) return -2;
    int gx = (npix + 63) / 64;  // 8 warps x 8 pixels per CTA
    if (gx > 16) gx = 16;
    if (Cin <= 512)
        mask_dot_kernel<2><<<dim3(gx, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(f3), omega, npix, Cin, D);
    else
        mask_dot_kernel<4><<<dim3(gx, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(f3), omega, npix, Cin, D);
    return launch_status();
}

extern "C" int b200_mask_wsum(const void* f3, const float* dm, int B, int npix, int Cin, float* s, void* stream) {
    if (B < 0 || npix <= 0 || Cin <= 0 || Cin % 8 != 0) return -1;
    if (B == 0) return 0;
    if (f3 == nullptr || dm == nullptr || s == nullptr) return -2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(s, 0, static_cast<size_t>(B) * Cin * sizeof(float), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    int chunks = B >= 148 ? 4 : (592 + B - 1) / B;  // >= 4 CTAs per SM in flight; 64 threads of a CTA carry loads
    if (chunks > npix) chunks = npix;
    const int nvec = Cin / 8;
    if (nvec <= 256) {
        const int threads = nvec * (256 / nvec);
        mask_wsum_vec_kernel<<<dim3(chunks, B), threads, 0, st>>>(static_cast<const __nv_bfloat16*>(f3), dm, npix, Cin, s);
    } else {
        mask_wsum_kernel<<<dim3(chunks, B), 128, 0, st>>>(static_cast<const __nv_bfloat16*>(f3), dm, npix, Cin, s);
    }
    return launch_status();
}

extern "C" int b200_mask_dice(const float* D_dwi, const float* D_dce, const float* gating, const float* u,
                              const float* lowres, const float* pre_b, const float* out_w, const float* out_b,
                              int mid, const float* target, const float* enc_mask_dwi, const float* enc_mask_dce,
                              int B, int H, int W, int Ho, int Wo, int Hp, int Wp, int C, float scale, float eps,
                              int loss_type, float* m_out, float* dm_out, float* q_out, float* dc0_out,
                              float* loss_out, void* stream) {
    if (loss_type != 0 && loss_type != 1) return -1;
    if (B < 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0 || Hp <= 0 || Wp <= 0 || C <= 0 || mid <= 0 ||
        H * W > 4096 || Ho * Wo > 4096 || Hp * Wp > 64)
        return -1;
    if (B == 0) return 0;
    if (D_dwi == nullptr || D_dce == nullptr || gating == nullptr || u == nullptr || pre_b == nullptr ||
        out_w == nullptr || out_b == nullptr || target == nullptr || m_out == nullptr || dm_out == nullptr ||
        q_out == nullptr || dc0_out == nullptr || loss_out == nullptr)
        return -2;
    const size_t smem = (static_cast<size_t>(2) * H * W + static_cast<size_t>(Ho) * Wo + 2 * Hp * Wp) * sizeof(float);
    if (smem > 48 * 1024) return -3;
    mask_dice_kernel<<<B, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        D_dwi, D_dce, gating, u, lowres, pre_b, out_w, out_b, mid, target, enc_mask_dwi, enc_mask_dce, H, W, Ho, Wo, Hp,
        Wp, C, scale, eps, loss_type, m_out, dm_out, q_out, dc0_out, loss_out);
    return launch_status();
}

extern "C" int b200_mask_head_grads(const float* dv, const float* dc0, const float* pre_w, const float* pre_b,
                                    const float* out_w, int mid, int C, float* g_pre_w, float* g_pre_b,
                                    float* g_out_w, float* g_out_b, void* stream) {
    if (mid <= 0 || C <= 0) return -1;
    if (dv == nullptr || dc0 == nullptr || pre_w == nullptr || pre_b == nullptr || out_w == nullptr ||
        g_pre_w == nullptr || g_pre_b == nullptr || g_out_w == nullptr || g_out_b == nullptr)
        return -2;
    mask_head_grads_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(dv, dc0, pre_w, pre_b, out_w, mid, C,
                                                                             g_pre_w, g_pre_b, g_out_w, g_out_b);
    return launch_status();
}

extern "C" int b200_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    if (n < 0 || step < 1) return -1;
    if (n == 0) return 0;
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr) return -2;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    long long grid = (n + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    adamw_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
        grad_scale);
    return launch_status();
}

extern "C" int b200_adamw_groups(float* p, const float* g, float* m, float* v, long long n, const float* lr_vec,
                                 const float* wd_vec, const int* step0, float lr_mult, float beta1, float beta2, float eps,
                                 int step, float grad_scale, void* stream) {
    if (n < 0 || step < 1) return -1;
    if (n == 0) return 0;
    if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || lr_vec == nullptr || wd_vec == nullptr) return -2;
    long long grid = (n + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    adamw_groups_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, n, lr_vec, wd_vec, step0, lr_mult, beta1, beta2, eps, step, grad_scale);
    return launch_status();
}
