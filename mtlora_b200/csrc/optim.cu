// Flat multi-tensor optimizer step for the trainable tensors of the path: gradient norm + GradScaler unscale + clip +
// AdamW over ~200 small tensors in two launches.
//
// Reference semantics: main.py:341-353 -> utils.py:348-369 (NativeScalerWithGradNormCount.__call__): scaler.unscale_,
// torch.nn.utils.clip_grad_norm_(parameters, clip_grad), scaler.step(optimizer) with optimizer = optim.AdamW
// (optimizer.py:58-60). The reference issues several hundred small launches for this (one unscale / norm / mul / Adam
// chain per parameter or per multi-tensor chunk); here the tensors are described by a device-side segment table and the
// work is split into fixed-size chunks, one CTA per chunk.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {
namespace {

constexpr int kChunk = MTL_OPT_CHUNK;   // elements per CTA
constexpr int kThreads = 256;

// segment owning chunk `c`: the last s with prefix[s] <= c
__device__ __forceinline__ int find_seg(const int32_t* __restrict__ prefix, int n_segs, int c) {
  int lo = 0, hi = n_segs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(prefix + mid) <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads)
opt_sqnorm_kernel(const mtl_opt_seg* __restrict__ segs, const int32_t* __restrict__ prefix, int n_segs,
                  float* __restrict__ out) {
  const int c = blockIdx.x;
  const int s = find_seg(prefix, n_segs, c);
  const mtl_opt_seg sg = segs[s];
  float acc = 0.f;
  if (sg.grad != nullptr) {
    const int64_t e0 = static_cast<int64_t>(c - prefix[s]) * kChunk;
    const int64_t e1 = min(e0 + kChunk, sg.numel);
    const float* g = sg.grad;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      for (int64_t i = e0 + 4 * threadIdx.x; i < e1; i += 4 * kThreads) {
        if (i + 4 <= e1) {
          const float4 v = *reinterpret_cast<const float4*>(g + i);
          acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        } else {
          for (int64_t k = i; k < e1; ++k) acc += g[k] * g[k];
        }
      }
    } else {
      for (int64_t i = e0 + threadIdx.x; i < e1; i += kThreads) acc += g[i] * g[i];
    }
  }
  acc = warp_sum(acc);
  __shared__ float part[kThreads / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);   // inf / nan propagate through the sum
  }
}

struct StepArgs {
  mtl_opt_group groups[MTL_OPT_MAX_GROUPS];
  float max_norm;        // <= 0: no clipping
  int adam_w;            // 1: decoupled weight decay (AdamW); 0: L2 added to the gradient (Adam)
};

__global__ void __launch_bounds__(kThreads)
opt_adamw_kernel(const mtl_opt_seg* __restrict__ segs, const int32_t* __restrict__ prefix, int n_segs,
                 float* __restrict__ flat_m, float* __restrict__ flat_v, float* __restrict__ state,
                 const float* __restrict__ grad_scale, const float* __restrict__ found_inf,
                 const float* __restrict__ sqnorm, const __grid_constant__ StepArgs a) {
  // state[0] = CTA completion counter of this launch, state[1] = total gradient norm of this step (after unscaling,
  // before clipping; reported to the caller), state[2 + i] = number of steps taken by segment i (torch.optim keeps one
  // step counter per parameter: a parameter without a gradient does not advance)
  const bool skip = found_inf != nullptr && *found_inf != 0.f;
  const int c = blockIdx.x;
  if (!skip) {
    const int s = find_seg(prefix, n_segs, c);
    const mtl_opt_seg sg = segs[s];
    if (sg.grad != nullptr) {
      const float step_prev = state[2 + s];
      const mtl_opt_group g = a.groups[sg.group];
      const float inv_scale = grad_scale != nullptr ? 1.f / *grad_scale : 1.f;
      float coef = inv_scale;
      if (a.max_norm > 0.f && sqnorm != nullptr) {
        const float total = sqrtf(*sqnorm) * inv_scale;
        coef *= fminf(a.max_norm / (total + 1e-6f), 1.f);   // torch.nn.utils.clip_grad_norm_
      }
      const double t = static_cast<double>(step_prev) + 1.0;
      const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(g.beta1), t));
      const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(g.beta2), t)));
      const float step_size = g.lr / bc1;
      const float decay = 1.f - g.lr * g.weight_decay;
      const int64_t e0 = static_cast<int64_t>(c - prefix[s]) * kChunk;
      const int64_t e1 = min(e0 + kChunk, sg.numel);
      float* __restrict__ p = sg.param;
      const float* __restrict__ gr = sg.grad;
      float* __restrict__ m = flat_m + sg.offset;
      float* __restrict__ v = flat_v + sg.offset;
      for (int64_t i = e0 + threadIdx.x; i < e1; i += kThreads) {
        float gi = gr[i] * coef;
        float pi = p[i];
        if (a.adam_w) pi *= decay; else gi += g.weight_decay * pi;
        const float mi = g.beta1 * m[i] + (1.f - g.beta1) * gi;
        const float vi = g.beta2 * v[i] + (1.f - g.beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * mi / (sqrtf(vi) / bc2_sqrt + g.eps);
      }
    }
  }
  // the last CTA to finish advances the step counters (every CTA has read its own by then) and publishes the norm
  __shared__ int is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const float done = atomicAdd(state, 1.f);
    is_last = done == static_cast<float>(gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (!skip)
      for (int i = threadIdx.x; i < n_segs; i += kThreads)
        if (segs[i].grad != nullptr) state[2 + i] += 1.f;
    if (threadIdx.x == 0) {
      state[0] = 0.f;
      if (sqnorm != nullptr) state[1] = sqrtf(*sqnorm) * (grad_scale != nullptr ? 1.f / *grad_scale : 1.f);
    }
  }
}

}  // namespace

int opt_sqnorm(const mtl_opt_seg* segs, const int32_t* prefix, int n_segs, int n_chunks, float* out,
               cudaStream_t stream) {
  MTL_REQUIRE(n_segs > 0 && n_chunks > 0 && out != nullptr, "adamw: empty segment table");
  MTL_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  opt_sqnorm_kernel<<<n_chunks, kThreads, 0, stream>>>(segs, prefix, n_segs, out);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int opt_adamw(const mtl_opt_seg* segs, const int32_t* prefix, int n_segs, int n_chunks, float* flat_m, float* flat_v,
              float* state, const mtl_opt_group* groups, int n_groups, const float* grad_scale, const float* found_inf,
              const float* sqnorm, float max_norm, int adam_w, cudaStream_t stream) {
  MTL_REQUIRE(n_segs > 0 && n_chunks > 0, "adamw: empty segment table");
  MTL_REQUIRE(n_groups >= 1 && n_groups <= MTL_OPT_MAX_GROUPS, "adamw: %d parameter groups (at most %d)", n_groups,
              MTL_OPT_MAX_GROUPS);
  MTL_REQUIRE(max_norm <= 0.f || sqnorm != nullptr, "adamw: clipping needs the squared norm (mtl_opt_sqnorm)");
  StepArgs a;
  for (int i = 0; i < MTL_OPT_MAX_GROUPS; ++i) a.groups[i] = groups[i < n_groups ? i : 0];
  a.max_norm = max_norm;
  a.adam_w = adam_w;
  opt_adamw_kernel<<<n_chunks, kThreads, 0, stream>>>(segs, prefix, n_segs, flat_m, flat_v, state, grad_scale,
                                                      found_inf, sqnorm, a);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
