// L2 -> SM random-row gather micro-benchmark (evidence for DESIGN.md §6: what can the part deliver to a gather?).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/gather_micro tools/gather_micro.cu -lcuda
//   tools/bin/gather_micro            # prints one line per (variant, table size)
//
// A table of R rows x 512 B (256 bf16, the row of the headline workload) is gathered through a uniform random
// index list of E entries; every gathered byte reaches the SM (registers or shared memory) and is reduced.
//   ldg        warp per 8-row batch, one 16-byte LDG per lane and row, FADD2 accumulate  (= aggregate inner loop)
//   ldg_nop    same loads, XOR instead of unpack+add                                     (= memory path alone)
//   tma_g4     cp.async.bulk.tensor.2d ... tile::gather4 into a 64-stage shared-memory ring (mbarriers), consumer
//              warps accumulate from shared memory                                       (= TMA-staged aggregate)
// Table sizes: 32 MB (L2 resident on a 126 MB L2) and 512 MB (the 1 M-row feature matrix of rmat-1m).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kRowBytes = 512;

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

template <bool kNop>
__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ tab, const uint32_t* __restrict__ idx, int64_t n_edges,
                                                  int edges_per_warp, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t e0 = warp * edges_per_warp;
  if (e0 >= n_edges) return;
  const int n = static_cast<int>(min(static_cast<int64_t>(edges_per_warp), n_edges - e0));
  float2 acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  uint4 x = make_uint4(0, 0, 0, 0);
  for (int base = 0; base < n; base += 32) {
    const uint32_t my = base + lane < n ? __ldg(idx + e0 + base + lane) : 0u;
    const int m = min(32, n - base);
    for (int u0 = 0; u0 + 8 <= m; u0 += 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t s = __shfl_sync(0xffffffffu, my, u0 + u);
        v[u] = __ldg(tab + static_cast<uint64_t>(s) * (kRowBytes / 16) + lane);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (kNop) {
          x.x ^= v[u].x; x.y ^= v[u].y; x.z ^= v[u].z; x.w ^= v[u].w;
        } else {
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[u]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc[i] = add2(acc[i], make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u)));
        }
      }
    }
  }
  float r = acc[0].x + acc[0].y + acc[1].x + acc[1].y + acc[2].x + acc[2].y + acc[3].x + acc[3].y +
            __uint_as_float(x.x ^ x.y ^ x.z ^ x.w);
  if (r == 123.456f) out[0] = r;       // keeps the loads alive
}

// ---- TMA gather4 ring -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_gather4(uint32_t smem, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

constexpr int kStages = 64;                 // 64 x 4 rows x 512 B = 128 KB in flight per CTA
constexpr int kStageBytes = 4 * kRowBytes;
constexpr int kConsWarps = 8;

__global__ void __launch_bounds__(32 * (kConsWarps + 1), 1)
tma_g4_kernel(const __grid_constant__ CUtensorMap map, const uint32_t* __restrict__ idx, int64_t n_edges, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's share of the edge list, in groups of 4 edges
  const int64_t groups = n_edges / 4;
  const int64_t g0 = groups * blockIdx.x / gridDim.x, g1 = groups * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == kConsWarps) {
    // producer: every lane issues the gather of its own stage -- 32 gathers (64 KB) per warp iteration
    for (int64_t g = g0 + lane; g < g1; g += 32) {
      const int64_t k = g - g0;
      const int stage = static_cast<int>(k % kStages);
      const uint32_t phase = static_cast<uint32_t>((k / kStages) & 1);
      const uint4 i4 = __ldg(reinterpret_cast<const uint4*>(idx) + g);
      mbar_wait(smem_u32(empty + stage), phase ^ 1);
      mbar_expect_tx(smem_u32(full + stage), kStageBytes);
      tma_gather4(smem_u32(smem + stage * kStageBytes), &map, smem_u32(full + stage), 0, i4.x, i4.y, i4.z, i4.w);
    }
  } else {
    float2 acc[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int64_t k = warp; k < g1 - g0; k += kConsWarps) {
      const int stage = static_cast<int>(k % kStages);
      const uint32_t phase = static_cast<uint32_t>((k / kStages) & 1);
      mbar_wait(smem_u32(full + stage), phase);
      const uint4* sp = reinterpret_cast<const uint4*>(smem + stage * kStageBytes) + lane;
      uint4 v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = sp[r * (kRowBytes / 16)];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(empty + stage));
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[r]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          acc[i] = add2(acc[i], make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u)));
      }
    }
    float r = acc[0].x + acc[0].y + acc[1].x + acc[1].y + acc[2].x + acc[2].y + acc[3].x + acc[3].y;
    if (r == 123.456f) out[0] = r;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int64_t n_edges = 16ll << 20;       // 16 Mi rows of 512 B = 8.6 GB gathered per launch
  float* out;
  CK(cudaMalloc(&out, 4));
  EncodeTiledFn enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  const size_t smem_bytes = kStages * kStageBytes + 2 * kStages * 8;
  CK(cudaFuncSetAttribute(tma_g4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  for (int64_t rows : {65536ll, 1048576ll}) {
    uint4* tab;
    CK(cudaMalloc(&tab, rows * kRowBytes));
    CK(cudaMemset(tab, 0x3c, rows * kRowBytes));
    std::vector<uint32_t> h(n_edges);
    uint64_t st = 88172645463325252ull;
    for (auto& v : h) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = static_cast<uint32_t>(st % rows); }
    uint32_t* idx;
    CK(cudaMalloc(&idx, n_edges * 4));
    CK(cudaMemcpy(idx, h.data(), n_edges * 4, cudaMemcpyHostToDevice));
    CUtensorMap map;
    {
      cuuint64_t gdim[2] = {256, static_cast<cuuint64_t>(rows)};
      cuuint64_t gstride[1] = {kRowBytes};
      cuuint32_t box[2] = {256, 1};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, tab, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch) {
      for (int i = 0; i < 3; ++i) launch();
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      const int reps = 10;
      for (int i = 0; i < reps; ++i) launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= reps;
      printf("{\"variant\": \"%s\", \"table_mb\": %lld, \"rows_gathered\": %lld, \"ms\": %.4f, \"gather_tb_s\": %.3f}\n", name,
             (long long)(rows * kRowBytes >> 20), (long long)n_edges, ms, n_edges * (double)kRowBytes / ms / 1e9);
      fflush(stdout);
    };
    for (int epw : {64, 512}) {
      const int64_t warps = (n_edges + epw - 1) / epw;
      const unsigned blocks = static_cast<unsigned>((warps + 7) / 8);
      char nm[64];
      snprintf(nm, sizeof nm, "ldg_%d_rows_per_warp", epw);
      run(nm, [&] { ldg_kernel<false><<<blocks, 256>>>(tab, idx, n_edges, epw, out); });
      snprintf(nm, sizeof nm, "ldg_nop_%d_rows_per_warp", epw);
      run(nm, [&] { ldg_kernel<true><<<blocks, 256>>>(tab, idx, n_edges, epw, out); });
    }
    run("tma_gather4_ring64", [&] { tma_g4_kernel<<<sms, 32 * (kConsWarps + 1), smem_bytes>>>(map, idx, n_edges, out); });
    CK(cudaFree(tab));
    CK(cudaFree(idx));
  }
  return 0;
}
