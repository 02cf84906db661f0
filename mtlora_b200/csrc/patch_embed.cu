// PatchEmbed forward for sm_100a: 4x4 stride-4 patch projection (Conv2d k4 s4) + bias + LayerNorm in one pass.
//
// Reference: models/swin_transformer_mtlora.py:568-611 (PatchEmbed.forward: self.proj(x).flatten(2).transpose(1, 2),
// then self.norm). The reference runs a cuDNN convolution on an NCHW image, two layout transposes and a LayerNorm;
// here every token (one 4x4x3 patch = 48 inputs) is read once from the fp32 image and three bf16 rows are written:
// the projection (LayerNorm input, kept for backward), the normalised token and the im2col'd patch (operand of
// dW = d_proj^T . patches in backward, mtl_xty). HBM traffic: 77 MB in, 77 + 77 + 38 MB out at batch 32 / 448 px.
//
// One warp handles two horizontally adjacent tokens per iteration: lanes 0-11 / 16-27 fetch the 12 float4 row segments
// of token A / B (adjacent 16-byte pieces -> full 32-byte sectors), the patches go through a warp-private smem
// slot, and lane l accumulates channels l, l + 32, ... against W^T held in shared memory (conflict-free reads, each
// weight read once for both tokens). LayerNorm statistics are warp reductions; all stores are 64-byte coalesced.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {

namespace {

constexpr int PE_K = 48;        // in_chans (3) * patch (4) * patch (4)
constexpr int PE_WARPS = 8;

struct PatchEmbedParams {
  const float* x;       // [B, 3, H, W]
  const float* w;       // [E, 48]
  const float* bias;    // [E]
  const float* gamma;   // [E] or null (no norm: y = projection)
  const float* beta;    // [E]
  __nv_bfloat16* proj;  // [B * L, E] projection + bias (null: not stored)
  __nv_bfloat16* y;     // [B * L, E]
  __nv_bfloat16* patches;  // [B * L, 48] (null: not stored)
  float* mean;          // [B * L]
  float* rstd;          // [B * L]
  int B, H, W, E;
  float eps;
};

template <int EPL>   // channels per lane: E = 32 * EPL
__global__ void __launch_bounds__(32 * PE_WARPS) patch_embed_fwd_kernel(const PatchEmbedParams p) {
  extern __shared__ float smem[];
  float* wt = smem;                                   // [48][E]  W^T
  float* slot = smem + PE_K * p.E;                    // [PE_WARPS][2][48] patches of the two tokens of each warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < PE_K * p.E; i += blockDim.x) {
    const int k = i / p.E, c = i - k * p.E;
    wt[i] = p.w[c * PE_K + k];
  }
  __syncthreads();
  float bias[EPL], gam[EPL], bet[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    bias[j] = p.bias != nullptr ? p.bias[lane + 32 * j] : 0.f;
    gam[j] = p.gamma != nullptr ? p.gamma[lane + 32 * j] : 1.f;
    bet[j] = p.gamma != nullptr ? p.beta[lane + 32 * j] : 0.f;
  }
  const int Pw = p.W >> 2, Ph = p.H >> 2;            // patch grid
  const long n_tok = static_cast<long>(p.B) * Ph * Pw;
  const long n_pair = (n_tok + 1) >> 1;
  float* my = slot + warp * 2 * PE_K;
  const int half = lane >> 4, l16 = lane & 15;       // lanes 0-11 -> token A, 16-27 -> token B
  const int ch = l16 >> 2, r = l16 & 3;              // input channel, row inside the patch
  const float inv_e = 1.f / p.E;

  for (long pr = static_cast<long>(blockIdx.x) * PE_WARPS + warp; pr < n_pair; pr += static_cast<long>(gridDim.x) * PE_WARPS) {
    const long t = 2 * pr + half;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l16 < 12 && t < n_tok) {
      const long b = t / (static_cast<long>(Ph) * Pw);
      const int rem = static_cast<int>(t - b * Ph * Pw);
      const int py = rem / Pw, px = rem - py * Pw;
      v = __ldg(reinterpret_cast<const float4*>(p.x + ((b * 3 + ch) * p.H + 4 * py + r) * p.W + 4 * px));
    }
    __syncwarp();   // previous iteration finished reading the slot
    if (l16 < 12) {
      *reinterpret_cast<float4*>(my + half * PE_K + 4 * l16) = v;
      if (p.patches != nullptr && t < n_tok)
        *reinterpret_cast<uint2*>(p.patches + t * PE_K + 4 * l16) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    __syncwarp();
    float a0[EPL], a1[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) a0[j] = a1[j] = bias[j];
#pragma unroll 4
    for (int k4 = 0; k4 < PE_K / 4; ++k4) {
      const float4 pa = *reinterpret_cast<const float4*>(my + 4 * k4);            // broadcast reads
      const float4 pb = *reinterpret_cast<const float4*>(my + PE_K + 4 * k4);
      const float xa[4] = {pa.x, pa.y, pa.z, pa.w}, xb[4] = {pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          const float wv = wt[(4 * k4 + q) * p.E + lane + 32 * j];
          a0[j] = fmaf(xa[q], wv, a0[j]);
          a1[j] = fmaf(xb[q], wv, a1[j]);
        }
      }
    }
    // LayerNorm of both tokens (two-pass statistics in registers); the statistics are taken on the bf16-rounded
    // projection, which is what backward (mtl_layernorm_bwd on `proj`) will see
#pragma unroll
    for (int tk = 0; tk < 2; ++tk) {
      const long tt = 2 * pr + tk;
      if (tt >= n_tok) break;
      float* a = tk ? a1 : a0;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        a[j] = __bfloat162float(__float2bfloat16_rn(a[j]));
        s += a[j];
      }
      const float mean = warp_sum(s) * inv_e;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < EPL; ++j) q += (a[j] - mean) * (a[j] - mean);
      const float rstd = rsqrtf(warp_sum(q) * inv_e + p.eps);
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const size_t o = static_cast<size_t>(tt) * p.E + lane + 32 * j;
        if (p.proj != nullptr) p.proj[o] = __float2bfloat16_rn(a[j]);
        const float yv = p.gamma != nullptr ? (a[j] - mean) * rstd * gam[j] + bet[j] : a[j];
        p.y[o] = __float2bfloat16_rn(yv);
      }
      if (lane == 0 && p.mean != nullptr) {
        p.mean[tt] = mean;
        p.rstd[tt] = rstd;
      }
    }
  }
}

}  // namespace

int launch_patch_embed_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                           void* proj, void* y, void* patches, float* mean, float* rstd, int B, int H, int W, int E,
                           float eps, cudaStream_t stream) {
  MTL_REQUIRE(x != nullptr && w != nullptr && y != nullptr, "patch_embed: NULL argument");
  MTL_REQUIRE(B > 0 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, "patch_embed: image %dx%d must be a multiple of the 4x4 patch", H, W);
  MTL_REQUIRE(E == 96 || E == 128, "patch_embed: embed_dim %d unsupported (96 or 128)", E);
  MTL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "patch_embed: image must be 16-byte aligned");
  MTL_REQUIRE(gamma == nullptr || beta != nullptr, "patch_embed: beta missing");
  PatchEmbedParams p{x, w, bias, gamma, beta, static_cast<__nv_bfloat16*>(proj), static_cast<__nv_bfloat16*>(y),
                     static_cast<__nv_bfloat16*>(patches), mean, rstd, B, H, W, E, eps};
  const long n_pair = (static_cast<long>(B) * (H / 4) * (W / 4) + 1) / 2;
  long blocks = (n_pair + PE_WARPS - 1) / PE_WARPS;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  const size_t sm = (static_cast<size_t>(PE_K) * E + PE_WARPS * 2 * PE_K) * sizeof(float);
  if (E == 96) patch_embed_fwd_kernel<3><<<static_cast<unsigned>(blocks), 32 * PE_WARPS, sm, stream>>>(p);
  else patch_embed_fwd_kernel<4><<<static_cast<unsigned>(blocks), 32 * PE_WARPS, sm, stream>>>(p);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
