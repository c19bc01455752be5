// stream_probe.cu — what can a 31 MB launch reach on this GPU?  Pure streaming kernels with the attention gather's
// geometry (100 x 36 rows of 8704 B picked from a 3 GB table, 400 CTAs) and no arithmetic worth mentioning:
//   bulk   : cp.async.bulk rows into a shared-memory ring, consumer only waits (the attention kernel's data path)
//   bulk3  : the same bytes as 3 large contiguous copies per CTA instead of 9 row copies
//   ldg    : LDG.128 straight into registers (what a torch reduction kernel does), 400 x 256 threads
//   ldg148 : LDG.128, one persistent 1024-thread CTA per SM
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_probe tools/stream_probe.cu && /tmp/stream_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}

constexpr int ROW = 2176 * 4, ROWS = 36, SPLIT = 4, PER = ROWS / SPLIT;   // 9 rows of 8704 B per CTA

// grid (4, B): CTA streams rows rank, rank+4, ... of slab idx[b]; ring of NST rows; one row per barrier like the real kernel
template <int NST, bool WORK = false, bool CSYNC = false>
__global__ void __launch_bounds__(256, 4) k_bulk(const char* table, const int* idx, float* sink) {
  __shared__ float red[2][8];
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + NST * ROW);
  const int b = blockIdx.y, rank = blockIdx.x, tid = threadIdx.x;
  const char* slab = table + (size_t)idx[b] * ROWS * ROW;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mb_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < NST && i < PER; ++i) { mb_expect(&bar[i], ROW); bulk(sm + i * ROW, slab + (size_t)(rank + SPLIT * i) * ROW, ROW, &bar[i]); }
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = 0; i < PER; ++i) {
    const int s = i % NST;
    while (!mb_try(&bar[s], (i / NST) & 1)) {}
    if (WORK) {   // the attention kernel's per-row work: 3 x LDS.128 + dot, warp + block reduction, exp, FMA
      const float4* r4 = reinterpret_cast<const float4*>(sm + s * ROW);
      float part = 0.f;
      float4 v[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int c = tid + 256 * j;
        v[j] = c < 544 ? r4[c] : make_float4(0, 0, 0, 0);
        part += v[j].x * 0.01f + v[j].y * 0.02f + v[j].z * 0.03f + v[j].w * 0.04f;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((tid & 31) == 0) red[i & 1][tid >> 5] = part;
      __syncthreads();
      float sr = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sr += red[i & 1][w];
      const float e = __expf(sr - 1.0f);
#pragma unroll
      for (int j = 0; j < 3; ++j) acc += e * (v[j].x + v[j].y + v[j].z + v[j].w);
    } else {
      acc += reinterpret_cast<const float*>(sm + s * ROW)[tid];
      __syncthreads();
    }
    if (tid == 0 && i + NST < PER) { mb_expect(&bar[s], ROW); bulk(sm + s * ROW, slab + (size_t)(rank + SPLIT * (i + NST)) * ROW, ROW, &bar[s]); }
  }
  if (CSYNC) {   // the two cluster barriers of the merge
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (acc == 123.456f) sink[0] = acc;
}

// same bytes, contiguous rows [9*rank, 9*rank+9) as 3 copies of 3 rows each, all in flight at once
__global__ void __launch_bounds__(256, 2) k_bulk3(const char* table, const int* idx, float* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + PER * ROW);
  const int b = blockIdx.y, rank = blockIdx.x, tid = threadIdx.x;
  const char* src = table + ((size_t)idx[b] * ROWS + (size_t)rank * PER) * ROW;
  if (tid == 0) {
    for (int s = 0; s < 3; ++s) mb_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < 3; ++s) { mb_expect(&bar[s], 3 * ROW); bulk(sm + s * 3 * ROW, src + (size_t)s * 3 * ROW, 3 * ROW, &bar[s]); }
  }
  __syncthreads();
  float acc = 0.f;
  for (int s = 0; s < 3; ++s) {
    while (!mb_try(&bar[s], 0)) {}
    acc += reinterpret_cast<const float*>(sm + s * 3 * ROW)[tid];
  }
  if (acc == 123.456f) sink[0] = acc;
}

// LDG.128: CTA (rank, b) reads its 9 interleaved rows; 544 float4 per row over 256 threads
__global__ void __launch_bounds__(256) k_ldg(const char* table, const int* idx, float* sink) {
  const int b = blockIdx.y, rank = blockIdx.x, tid = threadIdx.x;
  const float4* slab = reinterpret_cast<const float4*>(table + (size_t)idx[b] * ROWS * ROW);
  float4 v[PER * 3];
#pragma unroll
  for (int i = 0; i < PER; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = tid + 256 * j;
      v[i * 3 + j] = c < 544 ? __ldcs(slab + (size_t)(rank + SPLIT * i) * 544 + c) : make_float4(0, 0, 0, 0);
    }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < PER * 3; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
  if (acc == 123.456f) sink[0] = acc;
}

// persistent: 148 CTAs x 1024 threads, flat float4 index over the B slabs
__global__ void __launch_bounds__(1024) k_ldg148(const char* table, const int* idx, int B, float* sink) {
  const long long per_slab = (long long)ROWS * 544, total = per_slab * B;
  float acc = 0.f;
  for (long long i0 = (long long)blockIdx.x * 1024 * 8 + threadIdx.x; i0 < total; i0 += (long long)gridDim.x * 1024 * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long i = i0 + (long long)u * 1024;
      if (i < total) {
        const int b = (int)(i / per_slab);
        v[u] = __ldcs(reinterpret_cast<const float4*>(table + (size_t)idx[b] * ROWS * ROW) + (i - (long long)b * per_slab));
      } else v[u] = make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  const int NVP = 10000, NL = 200;
  for (int B : {100, 400}) {
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
      printf("B=%3d %-22s %7.2f us/launch  %6.0f GB/s\n", B, name, ms * 1e3 / NL, mb / (ms / NL) * 1e3 / 1e3);
    };
    CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * ROW + 64));
    CK(cudaFuncSetAttribute(k_bulk<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * ROW + 64));
    CK(cudaFuncSetAttribute(k_bulk3, cudaFuncAttributeMaxDynamicSharedMemorySize, PER * ROW + 64));
    time("bulk ring 4 rows", [&](int i) { k_bulk<4><<<dim3(SPLIT, B), 256, 4 * ROW + 64>>>(table, idx + (size_t)i * B, sink); });
    time("bulk ring 6 rows", [&](int i) { k_bulk<6><<<dim3(SPLIT, B), 256, 6 * ROW + 64>>>(table, idx + (size_t)i * B, sink); });
    auto launch_cl = [&](auto kern, int i, bool pdl) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(SPLIT, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 6 * ROW + 64;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = SPLIT; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
      const char* t = table; const int* ix = idx + (size_t)i * B;
      CK(cudaLaunchKernelEx(&cfg, kern, t, ix, sink));
    };
    CK(cudaFuncSetAttribute(k_bulk<6, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * ROW + 64));
    CK(cudaFuncSetAttribute(k_bulk<6, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * ROW + 64));
    CK(cudaFuncSetAttribute(k_bulk<6, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * ROW + 64));
    time("ring6 + row work", [&](int i) { k_bulk<6, true, false><<<dim3(SPLIT, B), 256, 6 * ROW + 64>>>(table, idx + (size_t)i * B, sink); });
    time("ring6 cluster(4)", [&](int i) { launch_cl(k_bulk<6, false, false>, i, false); });
    time("ring6 cluster(4)+2 csync", [&](int i) { launch_cl(k_bulk<6, false, true>, i, false); });
    time("ring6 cl+work+csync", [&](int i) { launch_cl(k_bulk<6, true, true>, i, false); });
    time("ring6 cl+work+csync+pdl", [&](int i) { launch_cl(k_bulk<6, true, true>, i, true); });
    time("bulk 3 x 26 KB", [&](int i) { k_bulk3<<<dim3(SPLIT, B), 256, PER * ROW + 64>>>(table, idx + (size_t)i * B, sink); });
    time("ldg 400x256", [&](int i) { k_ldg<<<dim3(SPLIT, B), 256>>>(table, idx + (size_t)i * B, sink); });
    time("ldg persistent 148x1024", [&](int i) { k_ldg148<<<148, 1024>>>(table, idx + (size_t)i * B, B, sink); });
    time("ldg persistent 296x1024", [&](int i) { k_ldg148<<<296, 1024>>>(table, idx + (size_t)i * B, B, sink); });
    CK(cudaFree(table)); CK(cudaFree(sink)); CK(cudaFree(idx));
  }
  return 0;
}
