// tail.cuh — the per-row rollout tail (follower.py:476-505), shared by the scoring kernels.
#pragma once
#include <cuda_bf16.h>

#include "kernels.h"

namespace sfb {

// ---------------------------------------------------------------- rollout tail of one batch row, executed by one warp
// (follower.py:476-505): mask, log-softmax, teacher / argmax / inverse-CDF sample, next-u gather, score and CE terms
// the chosen candidate row -> u_next (fp32) and / or the packed operand blocks of the next step's gate GEMM; `nthr`
// threads (ids t) share the copy
__device__ __forceinline__ void tail_copy_u(const TailParams& p, const int b, const int a_t, const float* rows, int t, int nthr) {
  if (!(p.u_next || p.upk)) return;
  const float4* src = reinterpret_cast<const float4*>(rows + (size_t)a_t * p.E);   // global all_u_t or staged smem rows
  float4* dst = p.u_next ? reinterpret_cast<float4*>(p.u_next + (size_t)b * p.E) : nullptr;
  const size_t half = (size_t)p.upk_NB * 128;
  for (int j = t; j < (p.E >> 2); j += nthr) {
    const float4 o = src[j];
    if (dst) dst[j] = o;
    if (p.upk) {
      const int k = j * 4;
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
      unsigned char* pd = p.upk + (size_t)(k >> 6) * (2 * half) + (size_t)(b >> 3) * 1024 + (size_t)((k & 63) >> 3) * 128 +
                          (size_t)(b & 7) * 16 + (size_t)(k & 7) * 2;
      *reinterpret_cast<uint2*>(pd) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      *reinterpret_cast<uint2*>(pd + half) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
  }
}

// `lgw` (optional): a shared-memory working copy of the row's raw logits — the tail then reads and masks that copy and
// writes the masked row back to p.logit once at the end; `valid_s` (optional): the row's validity flags in shared memory.
// copy_u = false: the caller copies the chosen row itself (tail_copy_u) with more threads.  tgt_pre / u_pre: the row's
// teacher index / uniform draw when the caller has fetched them ahead of time (TAIL_NOT_LOADED / negative: read here).
// Returns a_t.
constexpr int TAIL_NOT_LOADED = -0x7fffffff - 1;
__device__ __forceinline__ int tail_row(const TailParams& p, const int b, const int lane, const float* rows,
                                        float* lgw = nullptr, const float* valid_s = nullptr, bool copy_u = true,
                                        int tgt_pre = TAIL_NOT_LOADED, float u_pre = -1.f) {
  float* lg = lgw ? lgw : p.logit + (size_t)b * p.A;
  const float* valid = valid_s ? valid_s : p.is_valid + (size_t)b * p.A;
  // mask, max / first argmax (torch.max returns the first maximal index)
  // (torch.max treats NaN as the maximum and always returns an in-range index: a row of NaNs / no valid action must
  // not leave an out-of-range index behind, it is used as an address below)
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int a = lane; a < p.A; a += 32) {
    float v = lg[a];
    if (valid[a] == 0.f) {
      v = -INFINITY;
      lg[a] = v;
    }
    if (v != v) v = INFINITY;   // NaN wins, like torch.max
    if (v > m) { m = v; am = a; }
  }
  // warp arg-max with lowest-index tie break
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  if (am >= p.A) am = 0;       // every action masked: index 0 (torch.max over a row of -inf)
  __syncwarp();
  float z = 0.f;
  for (int a = lane; a < p.A; a += 32) {
    const float v = lg[a];
    if (v != -INFINITY) z += expf(v - m);
  }
  z = warp_sum(z);
  const float lse = m + logf(z);
  int a_t;
  int tgt = tgt_pre != TAIL_NOT_LOADED ? tgt_pre : (p.target ? p.target[b] : -1);
  if (tgt >= p.A) tgt = p.A - 1;   // out-of-range teacher index: clamp instead of reading outside the row
  if (p.feedback == 0) {
    a_t = tgt < 0 ? 0 : tgt;
  } else if (p.feedback == 1) {
    a_t = am;
  } else {
    // inverse-CDF draw over softmax(logit)*valid (follower.py:491-497), sequential in lane 0 (A is tiny)
    a_t = 0;
    if (lane == 0) {
      const float u = u_pre >= 0.f ? u_pre : p.sample_u[b];
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
  if (lgw) {
    __syncwarp();
    for (int a = lane; a < p.A; a += 32) p.logit[(size_t)b * p.A + a] = lgw[a];
  }
  if (copy_u) tail_copy_u(p, b, a_t, rows, lane, 32);
  return a_t;
}

}  // namespace sfb
