// Weight gradients of the GraphSAGE layer on the 5th-generation tensor cores:
//
//   D[n, k] = sum_s sum_m  G_s[m, n] * X_s[m, k]            n <= 256 outputs, k <= 256 inputs, m = rows (10^5..10^6)
//
// i.e. gW1 = g^T feat[:N_in] and gW2 = g^T ah of the autograd of /root/reference/module/layer.py:51 (one call each),
// with three (G, X) pairs per product for fp32 activations (3xTF32: hi*hi + hi*lo + lo*hi), or six bf16 pairs
// (fp32 = b0 + b1 + b2 in bf16: b0*b0 + b0*b1 + b1*b0 + b0*b2 + b2*b0 + b1*b1).
//
// Both operands are row-major [m, .]: the contraction index m is the SLOW index, so the tiles are MN-major for the
// tensor core (instruction-descriptor bits a_major = b_major = 1).  A TMA box of [128 bytes of n] x [KB rows of m]
// with the 128-byte swizzle is exactly one column of the canonical MN-major SWIZZLE_128B layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: SBO = 8 rows * 128 B, LBO = one box (tf32: the 32-byte-atom
// variant of the swizzle, 4 rows per atom).
// Split-K over m: every CTA owns a contiguous range of rows, accumulates its partial D in TMEM (two M=128
// accumulators when n > 128) and writes it to a workspace; a second kernel sums the partials in split order
// (deterministic).  HBM-bound: every operand row is read once per call.
//   warp 0: TMA producer   warp 1: tcgen05.mma issuer   warp 2: TMEM alloc   warps 4-7: epilogue (tcgen05.ld -> gmem)
#include <algorithm>

#include "tc_common.cuh"

namespace pg {

constexpr int kWgThreads = 256;
constexpr int kWgStages = 3;
constexpr int kWgMaxSrc = 6;

struct WgradMaps {
  CUtensorMap a[kWgMaxSrc];   // G_s: dims {n, m}, box {128 B of n, KB rows}
  CUtensorMap b[kWgMaxSrc];   // X_s: dims {k, m}, box {128 B of k, KB rows}
};

struct WgradArgs {
  int m, n, k;               // rows, outputs (D rows), inputs (D columns)
  int n_src;
  int kb_rows;               // KB: rows of m per pipeline stage (64 bf16, 32 tf32)
  int eb;                    // elements per 128-byte MN block (64 bf16, 32 tf32)
  int a_blocks, b_blocks;    // MN blocks per A / B tile (A padded to 128 or 256 rows of D)
  int n_acc;                 // 1 (n <= 128) or 2 accumulators of M = 128
  int k_pad;                 // UMMA N: k rounded up to 16
  int tmem_cols;
  int rows_per_split;        // multiple of kb_rows
  uint32_t idesc;
  float* partial;            // [splits][n_acc * 128][k_pad]
};

// MN-major descriptors, LBO = bytes between 128-byte MN blocks:
//   16-bit operands: SWIZZLE_128B (layout type 2), atoms of 8 rows of m: SBO = 1024 B
//   32-bit operands (tf32): SWIZZLE_128B_BASE32B (layout type 1; 32-byte chunks swizzled, Swizzle<2,5,2>), atoms of
//   4 rows of m: SBO = 512 B -- the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
template <bool kTf32>
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
  constexpr uint64_t sbo = kTf32 ? 512 : 1024;
  constexpr uint64_t type = kTf32 ? 1 : 2;
  return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((sbo >> 4) << 32) | (1ull << 46) | (type << 61);
}

template <bool kTf32>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ WgradMaps maps, const WgradArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int box_bytes = p.kb_rows * 128;
  const int a_bytes = p.a_blocks * box_bytes, b_bytes = p.b_blocks * box_bytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
  uint64_t* full_bar = bars;                 // [kWgStages]
  uint64_t* empty_bar = bars + kWgStages;    // [kWgStages]
  uint64_t* done_bar = bars + 2 * kWgStages; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_begin = blockIdx.x * p.rows_per_split;
  const int row_end = min(p.m, row_begin + p.rows_per_split);
  const int n_kb = row_end > row_begin ? (row_end - row_begin + p.kb_rows - 1) / p.kb_rows : 0;
  const int total_it = n_kb * p.n_src;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[s])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b[s])) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total_it; ++it) {
        const int kb = it / p.n_src, src = it % p.n_src;
        const int m0 = row_begin + kb * p.kb_rows;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        const uint32_t fb = smem_u32(full_bar + stage);
        mbar_expect_tx(fb, static_cast<uint32_t>(stage_bytes));
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        for (int blk = 0; blk < p.a_blocks; ++blk) tma_load_2d(sa + blk * box_bytes, &maps.a[src], fb, blk * p.eb, m0);
        for (int blk = 0; blk < p.b_blocks; ++blk) tma_load_2d(sb + blk * box_bytes, &maps.b[src], fb, blk * p.eb, m0);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int k_rows = kTf32 ? 8 : 16;                   // rows of m per MMA (32 bytes of K)
      const int blocks_per_acc = 128 / p.eb;               // MN blocks of A per M = 128 accumulator
      for (int it = 0; it < total_it; ++it) {
        mbar_wait(smem_u32(full_bar + stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        for (int ks = 0; ks < p.kb_rows / k_rows; ++ks) {
          const uint64_t bd = smem_desc_mn_sw128<kTf32>(sb + ks * k_rows * 128, box_bytes);
          for (int acc = 0; acc < p.n_acc; ++acc) {
            const uint64_t ad = smem_desc_mn_sw128<kTf32>(sa + acc * blocks_per_acc * box_bytes + ks * k_rows * 128, box_bytes);
            tc_mma<kTf32>(tmem_base + static_cast<uint32_t>(acc * 256), ad, bd, p.idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        tc_commit(smem_u32(empty_bar + stage));
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(smem_u32(done_bar));
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    float* out = p.partial + static_cast<int64_t>(blockIdx.x) * (p.n_acc * 128) * p.k_pad;
    if (total_it > 0) {
      mbar_wait(smem_u32(done_bar), 0);
      tc_fence_after();
    }
    for (int acc = 0; acc < p.n_acc; ++acc) {
      const int row = acc * 128 + q * 32 + lane;           // D row = output index n
      float* op = out + static_cast<int64_t>(row) * p.k_pad;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 256);
      for (int c0 = 0; c0 < p.k_pad; c0 += 16) {
        uint32_t r[16];
        if (total_it > 0) {
          tmem_ld16(tbase + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(op + c0 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// out[n, k] = sum over splits of partial[split][n][k], in split order (eight partials in flight per thread)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int rows_pad, int k_pad, float* __restrict__ out,
                    int64_t ldo, int n, int k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * k) return;
  const int r = idx / k, c = idx % k;
  const float* p = partial + static_cast<int64_t>(r) * k_pad + c;
  const int64_t stride = static_cast<int64_t>(rows_pad) * k_pad;
  float s = 0.f;
  int sp = 0;
  for (; sp + 8 <= splits; sp += 8) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = __ldcs(p + (sp + u) * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += t[u];
  }
  for (; sp < splits; ++sp) s += __ldcs(p + sp * stride);
  out[static_cast<int64_t>(r) * ldo + c] = s;
}

// [rows m, cols] row-major operand, box = [128 bytes of cols, kb_rows], 128B swizzle, zero fill out of bounds
static int make_mn_map(CUtensorMap* map, const void* ptr, int64_t ld, int m, int cols, int kb_rows, bool tf32) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return PG_ERR_UNSUPPORTED; }
  const int es = tf32 ? 4 : 2;
  PG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "pg_wgrad: operand base must be 16-byte aligned");
  PG_REQUIRE((ld * es) % 16 == 0, "pg_wgrad: operand row stride (%lld elements) must be a multiple of 16 bytes", (long long)ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(m)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), static_cast<cuuint32_t>(kb_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   tf32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) m=%d cols=%d ld=%lld", (int)r, m, cols, (long long)ld); return PG_ERR_CUDA; }
  return PG_OK;
}

static int wgrad_splits(int m, int kb_rows, int sm_count, int* rows_per_split) {
  const int n_kb = (m + kb_rows - 1) / kb_rows;
  int splits = std::min(sm_count, std::max(1, n_kb / 8));       // at least 8 k-blocks per CTA
  const int kb_per = (n_kb + splits - 1) / splits;
  splits = (n_kb + kb_per - 1) / kb_per;
  *rows_per_split = kb_per * kb_rows;
  return splits;
}

static int g_wg_sm_count = 0;
static int wg_sm_count() {
  if (g_wg_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_wg_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_wg_sm_count = 148;
  }
  return g_wg_sm_count;
}

}  // namespace pg

extern "C" int64_t pg_wgrad_workspace(int32_t m, int32_t n, int32_t k, int dtype_in) {
  using namespace pg;
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  int rps = 0;
  // the SM count is only known with a device; 148 (B200) bounds the split count from above
  const int splits = wgrad_splits(m, dtype_in == PG_F32 ? 32 : 64, 148, &rps);
  const int64_t rows_pad = n > 128 ? 256 : 128;
  return static_cast<int64_t>(splits) * rows_pad * round_up(k, 16);
}

extern "C" int pg_wgrad(int dtype_in, const pg_gemm_src* srcs, int32_t n_src, float* out, int64_t ldo, int32_t m,
                        int32_t n, int32_t k, float* workspace, int64_t workspace_floats, void* stream) {
  using namespace pg;
  PG_REQUIRE(srcs && out, "pg_wgrad: null operand");
  PG_REQUIRE(n_src >= 1 && n_src <= kWgMaxSrc, "pg_wgrad: 1..%d operand pairs, got %d", kWgMaxSrc, n_src);
  PG_REQUIRE(m >= 0 && n > 0 && n <= 256 && k > 0 && k <= 256, "pg_wgrad: unsupported shape m=%d n=%d k=%d (n, k <= 256)", m, n, k);
  PG_REQUIRE(dtype_in == PG_F32 || dtype_in == PG_BF16, "pg_wgrad: bad input dtype");
  PG_REQUIRE(ldo >= k, "pg_wgrad: ldo < k");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (m == 0) {
    PG_CHECK_CUDA(cudaMemset2DAsync(out, ldo * sizeof(float), 0, k * sizeof(float), n, st));
    return PG_OK;
  }
  const bool tf32 = dtype_in == PG_F32;
  WgradArgs p;
  WgradMaps maps;
  p.m = m; p.n = n; p.k = k; p.n_src = n_src;
  p.kb_rows = tf32 ? 32 : 64;
  p.eb = tf32 ? 32 : 64;
  p.n_acc = n > 128 ? 2 : 1;
  p.a_blocks = p.n_acc * 128 / p.eb;
  p.k_pad = static_cast<int>(round_up(k, 16));
  p.b_blocks = static_cast<int>(round_up(p.k_pad, p.eb) / p.eb);
  int cols = 32;
  while (cols < (p.n_acc == 2 ? 256 + p.k_pad : p.k_pad)) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t fmt = tf32 ? 2u : 1u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
            (static_cast<uint32_t>(p.k_pad >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
  const int splits = wgrad_splits(m, p.kb_rows, std::min(wg_sm_count(), 148), &p.rows_per_split);
  const int64_t rows_pad = p.n_acc * 128;
  PG_REQUIRE(workspace != nullptr && workspace_floats >= static_cast<int64_t>(splits) * rows_pad * p.k_pad,
             "pg_wgrad: workspace too small (%lld floats, need %lld: pg_wgrad_workspace)", (long long)workspace_floats,
             (long long)(static_cast<int64_t>(splits) * rows_pad * p.k_pad));
  p.partial = workspace;
  int rc;
  for (int s = 0; s < kWgMaxSrc; ++s) {
    if (s < n_src) {
      PG_REQUIRE(srcs[s].a && srcs[s].b, "pg_wgrad: operand pair %d is empty", s);
      if ((rc = make_mn_map(&maps.a[s], srcs[s].a, srcs[s].lda, m, n, p.kb_rows, tf32)) != PG_OK) return rc;
      if ((rc = make_mn_map(&maps.b[s], srcs[s].b, srcs[s].ldb, m, k, p.kb_rows, tf32)) != PG_OK) return rc;
    } else {
      maps.a[s] = maps.a[0];
      maps.b[s] = maps.b[0];
    }
  }
  const size_t smem = static_cast<size_t>(kWgStages) * (p.a_blocks + p.b_blocks) * p.kb_rows * 128 + 256 /*barriers*/ +
                      1024 /*align*/;
  static bool attr_set = false;
  if (!attr_set) {
    PG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    PG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  PG_REQUIRE(smem <= 232448, "pg_wgrad: tile does not fit shared memory (%zu bytes)", smem);
  if (tf32) wgrad_tcgen05_kernel<true><<<splits, kWgThreads, smem, st>>>(maps, p);
  else wgrad_tcgen05_kernel<false><<<splits, kWgThreads, smem, st>>>(maps, p);
  PG_LAUNCH_CHECK();
  const int total = n * k;
  wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(workspace, splits, static_cast<int>(rows_pad), p.k_pad, out, ldo, n, k);
  PG_LAUNCH_CHECK();
  return PG_OK;
}
