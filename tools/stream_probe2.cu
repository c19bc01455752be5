// stream_probe2.cu — per-SM streaming limits for the "one CTA per batch element" gather of step_fused.cu.
//   one<NCH,WORK> : B CTAs (one per SM), producer warp streams the element's 36 rows as 9 chunks of 4 rows through a
//                   ring of NCH chunks; WORK = 0 consumers only wait, 1 = 8 warps x 3 float4 (the fused kernel's loop),
//                   2 = 17 warps x 1 float4
//   span<NCH>     : G CTAs split the B*36 rows evenly (contiguous row ranges), no work: does spreading over all SMs help?
//   ldg1024       : B CTAs x 1024 threads, LDG.128 of the element's slab, no shared memory
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sp2 tools/stream_probe2.cu && /tmp/sp2
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}

constexpr int ROWF = 2048, ROW = ROWF * 4, ROWS = 36, RB = 4, CH = RB * ROW, NCHUNK = ROWS / RB;

template <int NCH, int WORK>
__global__ void __launch_bounds__(WORK == 2 ? 576 : 288, 1) k_one(const char* table, const int* idx, float* sink) {
  constexpr int NCW = WORK == 2 ? 17 : 8;          // consumer warps
  constexpr int NT = NCW * 32;
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + NCH * CH);
  uint64_t* empty = full + NCH;
  float* red = reinterpret_cast<float*>(empty + NCH);   // [2][NCW][RB]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NCH; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NCW) {
    if (lane == 0) {
      const char* slab = table + (size_t)idx[b] * ROWS * ROW;
      for (int c = 0; c < NCHUNK; ++c) {
        const int s = c % NCH;
        if (c >= NCH) while (!mb_try(&empty[s], (c / NCH - 1) & 1)) {}
        mb_expect(&full[s], CH);
        bulk(sm + s * CH, slab + (size_t)c * CH, CH, &full[s]);
      }
    }
    return;
  }
  float acc4[WORK == 1 ? 8 : 4] = {0.f};
  float m = -1e30f, Z = 0.f;
  for (int c = 0; c < NCHUNK; ++c) {
    const int s = c % NCH;
    while (!mb_try(&full[s], (c / NCH) & 1)) {}
    if (WORK == 0) {
      acc4[0] += reinterpret_cast<const float*>(sm + s * CH)[tid];
      asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
      if (tid == 0) mb_arrive(&empty[s]);
    } else {
      constexpr int NJ = WORK == 1 ? 2 : 1;        // 2048 floats = 512 float4: 8 warps x 2 or 16 (+1 idle) warps x 1
      const float4* ch4 = reinterpret_cast<const float4*>(sm + s * CH);
      float4 v[RB][NJ];
      float part[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        part[r] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int col = tid + NT * j;
          v[r][j] = col < 512 ? ch4[r * 512 + col] : make_float4(0, 0, 0, 0);
          part[r] += v[r][j].x * 0.01f + v[r][j].y * 0.02f + v[r][j].z * 0.03f + v[r][j].w * 0.04f;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < RB; ++r) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
      float* rb = red + (c & 1) * NCW * RB;
      if (lane == 0)
#pragma unroll
        for (int r = 0; r < RB; ++r) rb[warp * RB + r] = part[r];
      asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
      if (tid == 0) mb_arrive(&empty[s]);
      float sr[RB], mn = m;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        sr[r] = 0.f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) sr[r] += rb[w * RB + r];
        mn = fmaxf(mn, sr[r]);
      }
      const float corr = __expf(m - mn);
      float e[RB], es = 0.f;
#pragma unroll
      for (int r = 0; r < RB; ++r) { e[r] = __expf(sr[r] - mn); es += e[r]; }
      Z = Z * corr + es;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float a0 = acc4[4 * j] * corr, a1 = acc4[4 * j + 1] * corr, a2 = acc4[4 * j + 2] * corr, a3 = acc4[4 * j + 3] * corr;
#pragma unroll
        for (int r = 0; r < RB; ++r) { a0 += e[r] * v[r][j].x; a1 += e[r] * v[r][j].y; a2 += e[r] * v[r][j].z; a3 += e[r] * v[r][j].w; }
        acc4[4 * j] = a0; acc4[4 * j + 1] = a1; acc4[4 * j + 2] = a2; acc4[4 * j + 3] = a3;
      }
      m = mn;
    }
  }
  if (acc4[0] + Z == 123.456f) sink[0] = acc4[0];
}

// G CTAs split the B*36 rows evenly (contiguous ranges of rows, each row from its element's slab), ring of NCH row pairs
template <int NCH>
__global__ void __launch_bounds__(288, 1) k_span(const char* table, const int* idx, int B, float* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int CB = 2 * ROW;   // 2 rows per copy (a range never needs more alignment than that here)
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + NCH * CB);
  uint64_t* empty = full + NCH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = B * ROWS / 2;                                   // row pairs
  const int per = (total + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(total, p0 + per);
  if (tid == 0) {
    for (int s = 0; s < NCH; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    if (lane == 0)
      for (int i = p0; i < p1; ++i) {
        const int c = i - p0, s = c % NCH;
        if (c >= NCH) while (!mb_try(&empty[s], (c / NCH - 1) & 1)) {}
        const int row = 2 * i, b = row / ROWS, r = row - b * ROWS;
        mb_expect(&full[s], CB);
        bulk(sm + s * CB, table + ((size_t)idx[b] * ROWS + r) * ROW, CB, &full[s]);
      }
    return;
  }
  float acc = 0.f;
  for (int i = p0; i < p1; ++i) {
    const int c = i - p0, s = c % NCH;
    while (!mb_try(&full[s], (c / NCH) & 1)) {}
    acc += reinterpret_cast<const float*>(sm + s * CB)[tid];
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) mb_arrive(&empty[s]);
  }
  if (acc == 123.456f) sink[0] = acc;
}

__global__ void __launch_bounds__(1024, 1) k_ldg1024(const char* table, const int* idx, float* sink) {
  const float4* slab = reinterpret_cast<const float4*>(table + (size_t)idx[blockIdx.x] * ROWS * ROW);
  float acc = 0.f;
  constexpr int N4 = ROWS * ROW / 16;   // 18432 float4
#pragma unroll 1
  for (int i0 = threadIdx.x; i0 < N4; i0 += 1024 * 6) {
    float4 v[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) v[u] = (i0 + u * 1024 < N4) ? __ldcs(slab + i0 + u * 1024) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 6; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int NVP = 10000, NL = 200, B = 100;
  char* table; float* sink; int* idx;
  CK(cudaMalloc(&table, (size_t)NVP * ROWS * ROW));
  CK(cudaMemset(table, 0, (size_t)NVP * ROWS * ROW));
  CK(cudaMalloc(&sink, 16));
  std::vector<int> h((size_t)NL * B);
  srand(1);
  for (auto& x : h) x = rand() % NVP;
  CK(cudaMalloc(&idx, h.size() * 4));
  CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double mb = (double)B * ROWS * ROW / 1e6;
  auto time = [&](const char* name, auto launch) {
    for (int i = 0; i < 5; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < NL; ++i) launch(i);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("B=%3d %-34s %7.2f us/launch  %6.0f GB/s\n", B, name, ms * 1e3 / NL, mb / (ms / NL) * 1e3 / 1e3);
  };
#define ONE(NCH, W, name)                                                                                          \
  {                                                                                                                \
    const int smem = NCH * CH + 2 * NCH * 8 + 2 * 17 * RB * 4 + 64;                                                \
    CK(cudaFuncSetAttribute(k_one<NCH, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                    \
    time(name, [&](int i) { k_one<NCH, W><<<B, (W == 2 ? 576 : 288), smem>>>(table, idx + (size_t)i * B, sink); }); \
  }
  ONE(2, 0, "one CTA/elem ring 2x32KB nowork");
  ONE(4, 0, "one CTA/elem ring 4x32KB nowork");
  ONE(6, 0, "one CTA/elem ring 6x32KB nowork");
  ONE(4, 1, "one CTA/elem ring 4 work 8w x2");
  ONE(6, 1, "one CTA/elem ring 6 work 8w x2");
  ONE(4, 2, "one CTA/elem ring 4 work 17w x1");
  ONE(6, 2, "one CTA/elem ring 6 work 17w x1");
#define SPAN(NCH, G, name)                                                                                        \
  {                                                                                                                \
    const int smem = NCH * 2 * ROW + 2 * NCH * 8 + 64;                                                             \
    CK(cudaFuncSetAttribute(k_span<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                      \
    time(name, [&](int i) { k_span<NCH><<<G, 288, smem>>>(table, idx + (size_t)i * B, B, sink); });                \
  }
  SPAN(8, 100, "span 100 CTAs ring 8x16KB");
  SPAN(12, 100, "span 100 CTAs ring 12x16KB");
  SPAN(8, 144, "span 144 CTAs ring 8x16KB");
  SPAN(12, 144, "span 144 CTAs ring 12x16KB");
  SPAN(12, 148, "span 148 CTAs ring 12x16KB");
  SPAN(12, 72, "span 72 CTAs ring 12x16KB");
  SPAN(12, 50, "span 50 CTAs ring 12x16KB");
  time("ldg 100 x 1024 thr", [&](int i) { k_ldg1024<<<B, 1024>>>(table, idx + (size_t)i * B, sink); });
  return 0;
}
