// pointwise.cu — the small fused kernels around the GEMMs: action scoring (EltwiseProdScoring, model.py:342-352, in its re-associated form) and
// the per-step tail of the follower rollout (follower.py:476-505).
#include "kernels.h"

namespace sfb {

// ---------------------------------------------------------------- action scoring
// w_out . ((W_h ht + b_h) (.) (W_a u + b_a)) + b_out  ==  u . g + c   with  t' = W_h ht + b_h,
// g = W_a^T (w_out (.) t')  and  c = sum_d b_a[d] w_out[d] t'[d] + b_out   (SURVEY.md §7 hard part 1).
__global__ void __launch_bounds__(256) action_scoring_kernel(const ScoringParams p) {
  extern __shared__ __align__(16) float gs[];   // [E]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = p.E >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(p.g + (size_t)b * p.E);
  for (int j = tid; j < nvec; j += 256) reinterpret_cast<float4*>(gs)[j] = g4[j];
  float c = 0.f;
  for (int d = tid; d < p.D; d += 256) c = fmaf(__ldg(p.b_a + d) * __ldg(p.w_out + d), p.tp[(size_t)b * p.D + d], c);
  c = warp_sum(c);
  if (lane == 0) red[warp] = c;
  __syncthreads();
  float cst = __ldg(p.b_out);
#pragma unroll
  for (int w = 0; w < 8; ++w) cst += red[w];
  for (int a = warp; a < p.A; a += 8) {
    const float4* u4 = reinterpret_cast<const float4*>(p.all_u_t + ((size_t)b * p.A + a) * p.E);
    float acc = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      const float4 u = __ldg(u4 + j);
      const float4 g = reinterpret_cast<const float4*>(gs)[j];
      acc = fmaf(u.x, g.x, acc);
      acc = fmaf(u.y, g.y, acc);
      acc = fmaf(u.z, g.z, acc);
      acc = fmaf(u.w, g.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) p.logit[(size_t)b * p.A + a] = acc + cst;
  }
}

int32_t launch_action_scoring(const ScoringParams& p, cudaStream_t stream) {
  SFB_CHECK_ARG((p.E % 4) == 0, "scoring: E % 4");
  SFB_CHECK_ARG((size_t)p.E * 4 <= 48 * 1024, "scoring: E too large");
  action_scoring_kernel<<<p.B, 256, (size_t)p.E * sizeof(float), stream>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- follower rollout tail (one warp per row)
__global__ void __launch_bounds__(128) follower_tail_kernel(const TailParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= p.B) return;
  float* lg = p.logit + (size_t)b * p.A;
  const float* valid = p.is_valid + (size_t)b * p.A;
  // mask, max / first argmax (torch.max returns the first maximal index)
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int a = lane; a < p.A; a += 32) {
    float v = lg[a];
    if (valid[a] == 0.f) {
      v = -INFINITY;
      lg[a] = v;
    }
    if (v > m) { m = v; am = a; }
  }
  // warp arg-max with lowest-index tie break
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  __syncwarp();
  float z = 0.f;
  for (int a = lane; a < p.A; a += 32) {
    const float v = lg[a];
    if (v != -INFINITY) z += expf(v - m);
  }
  z = warp_sum(z);
  const float lse = m + logf(z);
  int a_t;
  const int tgt = p.target ? p.target[b] : -1;
  if (p.feedback == 0) {
    a_t = tgt < 0 ? 0 : tgt;
  } else if (p.feedback == 1) {
    a_t = am;
  } else {
    // inverse-CDF draw over softmax(logit)*valid (follower.py:491-497), sequential in lane 0 (A is tiny)
    a_t = 0;
    if (lane == 0) {
      const float u = p.sample_u[b];
      float cdf = 0.f;
      int last_valid = 0, pick = -1;
      for (int a = 0; a < p.A; ++a) {
        const float v = lg[a];
        if (v == -INFINITY) continue;
        cdf += expf(v - m) / z;
        last_valid = a;
        if (pick < 0 && !(u > cdf)) pick = a;
      }
      a_t = pick < 0 ? last_valid : pick;
    }
    a_t = __shfl_sync(0xffffffffu, a_t, 0);
  }
  if (lane == 0) {
    p.a_t[b] = a_t;
    if (p.action_score) p.action_score[b] = lg[a_t] - lse;
    if (p.ce) p.ce[b] = tgt < 0 ? 0.f : -(lg[tgt] - lse);
  }
  if (p.u_next) {
    const float4* src = reinterpret_cast<const float4*>(p.all_u_t + ((size_t)b * p.A + a_t) * p.E);
    float4* dst = reinterpret_cast<float4*>(p.u_next + (size_t)b * p.E);
    for (int j = lane; j < (p.E >> 2); j += 32) dst[j] = __ldg(src + j);
  }
}

int32_t launch_follower_tail(const TailParams& p, cudaStream_t stream) {
  follower_tail_kernel<<<(p.B + 3) / 4, 128, 0, stream>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}

}  // namespace sfb
