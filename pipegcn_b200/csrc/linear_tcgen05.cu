// Dense part of the GraphSAGE layer on the 5th-generation tensor cores:
//
//   C[M,N] = sum_s A_s[M,K_s] * B_s[N,K_s]^T  (+ bias[N])  (/ row_div[M])          1 <= s <= 6 operand pairs
//
// i.e. `linear1(feat[:N_in]) + linear2(ah)` of /root/reference/module/layer.py:51 as ONE kernel
// (pair 0 = inner rows x W1, pair 1 = neighbour mean x W2, bias = b1 + b2), with a single pair the dX
// GEMMs of its backward (row_div = in_deg fuses the `/ degs` gradient), and with three pairs per product
// the split-fp32 "3xTF32" evaluation (hi*hi + hi*lo + lo*hi) used for fp32 activations.
//
// Persistent, warp-specialised, sm_100a only:
//   warp 0    TMA producer: 128-byte-swizzled K-major tiles of A (128 rows) and B (n_pad rows)
//             into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1    one thread issues tcgen05.mma (kind::f16 for bf16, kind::tf32 for fp32 inputs),
//             fp32 accumulators live in TMEM, two accumulator stages so that the epilogue of tile
//             i overlaps the MMAs of tile i+1
//   warp 2    allocates / frees TMEM
//   warps 4-7 epilogue: tcgen05.ld (32 lanes x 16 columns), bias / row scale, convert, 128B-swizzled
//             shared-memory slab, TMA store (cp.async.bulk.tensor) -- every global write is a full line
// M is tiled by 128 (TMA zero-fills the tail, the epilogue masks it); N <= 256 is one tile
// (padded to a multiple of 16 with zero-filled weight rows); K tails are zero-filled by TMA.
#include "tc_common.cuh"

namespace pg {

constexpr int kBlockM = 128;
constexpr int kRowBytes = 128;          // one swizzle-128B row: 64 bf16 or 32 tf32 of K
constexpr int kStages = 4;
constexpr int kGemmThreads = 256;
constexpr int kABytes = kBlockM * kRowBytes;   // 16 KB per stage

constexpr int kMaxSrc = 6;              // 2 operands pairs x 3 passes of the split-fp32 (3xTF32) product

struct GemmMaps {
  CUtensorMap a[kMaxSrc];
  CUtensorMap b[kMaxSrc];
  CUtensorMap c;             // output, box = [32 rows, 128 bytes], used by the TMA-store epilogue
};
constexpr int kSlabBytes = 32 * kRowBytes;     // one epilogue warp's staging slab: 32 rows x 128 B

struct GemmArgs {
  int m, n, n_pad;
  int n_src;
  int kb_end[kMaxSrc];       // prefix sums of the k-blocks (of 128 bytes) of every source
  int k_elems;               // elements per k-block
  const float* bias;         // [n] or null
  const float* row_div;      // [m] or null
  void* c;
  int64_t ldc;
  int out_bf16;              // 1: bf16 output, 0: fp32 output
  int tmem_cols;             // power of two >= 2 * n_pad
  uint32_t idesc;
  int is_tf32;
  int tma_store;             // 1: epilogue stages through shared memory and stores with TMA
  int epi_batch;             // 1: the epilogue issues all TMEM loads of a 128-byte chunk before waiting
  int epi_slabs;             // 1 | 2 staging slabs per epilogue warp
  DropArg drop;              // dropout applied to the output as it is written (thresh16 == 0: none)
  int64_t drop_row0;         // row index of c's first row in the tensor the mask is defined on
  int drop_nvec;             // 16-byte vectors per row of that tensor
};

// bias / row scale / convert / store `cnt` (16 or 32) accumulator columns of one row
template <int CNT>
__device__ __forceinline__ void store_chunk(const GemmArgs& p, int row, int c0, const uint32_t* r, float inv_unused,
                                            float dv) {
  float v[CNT];
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    v[i] = __uint_as_float(r[i]);
    if (p.bias != nullptr && c0 + i < p.n) v[i] += __ldg(p.bias + c0 + i);
    if (p.row_div != nullptr) v[i] = v[i] * (1.f / dv);
  }
  const bool full = (c0 + CNT <= p.n);
  if (p.out_bf16) {
    __nv_bfloat16* cp = static_cast<__nv_bfloat16*>(p.c) + static_cast<int64_t>(row) * p.ldc + c0;
    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < CNT; i += 8) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
        *reinterpret_cast<uint4*>(cp + i) = pk;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CNT; ++i)
        if (c0 + i < p.n) cp[i] = __float2bfloat16_rn(v[i]);
    }
  } else {
    float* cp = static_cast<float*>(p.c) + static_cast<int64_t>(row) * p.ldc + c0;
    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < CNT; i += 4) *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < CNT; ++i)
        if (c0 + i < p.n) cp[i] = v[i];
    }
  }
}

template <bool kTf32>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const GemmArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment of the swizzled tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.n_pad * kRowBytes;
  const int stage_bytes = kABytes + b_bytes;
  uint8_t* slabs = smem + kStages * stage_bytes;                          // 4 warps x 2 x 4 KB, 1024-byte aligned
  float* s_bias = reinterpret_cast<float*>(slabs + 8 * kSlabBytes);       // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + 8 * kSlabBytes + 1024);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.m + kBlockM - 1) / kBlockM;
  const int kb_total = p.kb_end[p.n_src - 1];

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[s])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b[s])) : "memory");
    }
  }
  for (int i = threadIdx.x; i < 256; i += kGemmThreads)
    s_bias[i] = (p.bias != nullptr && i < p.n) ? __ldg(p.bias + i) : 0.f;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(tfull_bar + s), 1);
      mbar_init(smem_u32(tempty_bar + s), 4);     // one arrive per epilogue warp
    }
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = tile * kBlockM;
        int src = 0, kb_begin = 0;
        for (int kb = 0; kb < kb_total; ++kb) {
          while (kb >= p.kb_end[src]) { kb_begin = p.kb_end[src]; ++src; }
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t fb = smem_u32(full_bar + stage);
          mbar_expect_tx(fb, static_cast<uint32_t>(stage_bytes));
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + kABytes;
          tma_load_2d(sa, &maps.a[src], fb, (kb - kb_begin) * p.k_elems, m0);
          tma_load_2d(sb, &maps.b[src], fb, (kb - kb_begin) * p.k_elems, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(tempty_bar + as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + static_cast<uint32_t>(as * p.n_pad);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kRowBytes / 32; ++k) {        // 32 bytes of K per MMA (16 bf16 / 8 tf32)
            const uint64_t ad = smem_desc_sw128(sa + k * 32);
            const uint64_t bd = smem_desc_sw128(sb + k * 32);
            tc_mma<kTf32>(tmem_c, ad, bd, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(empty_bar + stage));            // frees the smem stage when the MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(smem_u32(tfull_bar + as));                 // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;                                  // == warp % 4: TMEM lane quarter
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t drop_hi = drop_seed_hi(p.drop);
    if (p.tma_store) {
      // TMEM -> registers -> (bias, row scale, convert) -> 128B-swizzled smem slab -> TMA store:
      // every global write is a full, coalesced 128-byte row segment issued by the copy engine
      // two staging slabs per warp: the TMA store of a chunk reads one while the next chunk is written to the other
      // (epi_slabs == 1: one slab, every chunk waits for the previous store to have read it)
      const int two = p.epi_slabs == 2 ? 1 : 0;
      uint32_t ck = 0;                                       // chunks issued by this warp
      const int cpc = p.out_bf16 ? 64 : 32;                  // columns per 128-byte chunk
      const int xr = lane & 7;                               // swizzle phase of this thread's row
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(tfull_bar + as), aphase);
        tc_fence_after();
        const int row0 = tile * kBlockM + q * 32;
        const int row = row0 + lane;
        const float scale = (p.row_div != nullptr && row < p.m) ? 1.f / __ldg(p.row_div + row) : 1.f;
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.n_pad);
        for (int c0 = 0; c0 < p.n; c0 += cpc, ++ck) {
          uint8_t* slab = slabs + (q * 2 + (two ? (ck & 1u) : 0u)) * kSlabBytes;
          const uint32_t slab_u32 = smem_u32(slab);
          uint8_t* my = slab + lane * kRowBytes;
          // epi_batch: all TMEM loads of the chunk (up to 4 x 16 columns) are issued first and waited for ONCE, in
          // flight while the previous chunk's TMA store finishes reading the slab; otherwise one load + wait per 16
          // columns after the slab is free (pg_set_option("gemm_epi_batch", 0|1))
          const int nsub = (min(cpc, p.n_pad - c0) + 15) >> 4;
          uint32_t r[4][16];
          if (p.epi_batch) {
#pragma unroll
            for (int sb = 0; sb < 4; ++sb)
              if (sb < nsub) tmem_ld16(tbase + c0 + 16 * sb, r[sb]);
          }
          if (lane == 0) {                                   // the slab about to be written is free again
            if (two) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
          if (p.epi_batch) tmem_ld_wait();
#pragma unroll
          for (int sb = 0; sb < 4; ++sb) {
            if (sb >= nsub) continue;
            const int cc = 16 * sb;
            if (!p.epi_batch) {
              tmem_ld16(tbase + c0 + cc, r[sb]);
              tmem_ld_wait();
            }
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + cc + i);
              v[i] = (__uint_as_float(r[sb][i]) + b4.x) * scale;
              v[i + 1] = (__uint_as_float(r[sb][i + 1]) + b4.y) * scale;
              v[i + 2] = (__uint_as_float(r[sb][i + 2]) + b4.z) * scale;
              v[i + 3] = (__uint_as_float(r[sb][i + 3]) + b4.w) * scale;
            }
            if (p.drop.thresh16 != 0u) {
              // mask of the (rounded) output, per 16-byte vector: what pg_dropout_rows would make of c
              const uint64_t ibase = static_cast<uint64_t>(p.drop_row0 + row) * p.drop_nvec;
              if (p.out_bf16) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  float t[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) t[j] = __bfloat162float(__float2bfloat16_rn(v[8 * h + j]));
                  drop_apply<8>(t, ibase + ((c0 + cc) >> 3) + h, p.drop.thresh16, p.drop.scale, p.drop.seed_lo, drop_hi);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * h + j] = t[j];
                }
              } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  float t[4] = {v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]};
                  drop_apply<4>(t, ibase + ((c0 + cc) >> 2) + h, p.drop.thresh16, p.drop.scale, p.drop.seed_lo, drop_hi);
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[4 * h + j] = t[j];
                }
              }
            }
            if (p.out_bf16) {                                // 16 columns = 2 chunks of 16 bytes
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint4 pk;
                __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int j = 0; j < 4; ++j) hp[j] = __floats2bfloat162_rn(v[8 * h + 2 * j], v[8 * h + 2 * j + 1]);
                const int chunk = (cc >> 3) + h;
                *reinterpret_cast<uint4*>(my + ((chunk ^ xr) << 4)) = pk;
              }
            } else {                                         // 16 columns = 4 chunks of 16 bytes
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const int chunk = (cc >> 2) + h;
                *reinterpret_cast<float4*>(my + ((chunk ^ xr) << 4)) =
                    make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (row0 < p.m) tma_store_2d(&maps.c, slab_u32, c0, row0);
            // one group per chunk, empty or not: the slab parity (ck & 1) and the group count stay in step
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty_bar + as));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(tfull_bar + as), aphase);
        tc_fence_after();
        const int row = tile * kBlockM + q * 32 + lane;
        const bool row_ok = row < p.m;
        const float dv = (p.row_div != nullptr && row_ok) ? __ldg(p.row_div + row) : 1.f;
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.n_pad);
        int c0 = 0;
        for (; c0 + 32 <= p.n_pad; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tbase + c0, r);
          tmem_ld_wait();
          if (row_ok && c0 < p.n) store_chunk<32>(p, row, c0, r, 0.f, dv);
        }
        if (c0 < p.n_pad) {
          uint32_t r[16];
          tmem_ld16(tbase + c0, r);
          tmem_ld_wait();
          if (row_ok && c0 < p.n) store_chunk<16>(p, row, c0, r, 0.f, dv);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(tempty_bar + as));
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
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

// ---- host side -----------------------------------------------------------------------------------
// rows x k row-major matrix, box = [box_rows, 128 bytes of k], 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* ptr, int64_t ld, int rows, int k, int box_rows, bool tf32) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return PG_ERR_UNSUPPORTED; }
  const int es = tf32 ? 4 : 2;
  PG_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "pg_linear: operand base must be 16-byte aligned");
  PG_REQUIRE((ld * es) % 16 == 0, "pg_linear: operand row stride (%lld elements) must be a multiple of 16 bytes", (long long)ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kRowBytes / es), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d k=%d ld=%lld", (int)r, rows, k, (long long)ld); return PG_ERR_CUDA; }
  return PG_OK;
}

static int g_sm_count = 0;
int g_gemm_epi_batch = 1;     // pg_set_option("gemm_epi_batch", 0|1)
int g_gemm_epi_slabs = 2;     // pg_set_option("gemm_epi_slabs", 1|2)

}  // namespace pg

extern "C" int pg_linear(int dtype_in, int dtype_out, const pg_gemm_src* srcs, int32_t n_src, const float* bias,
                         const float* row_div, void* c, int64_t ldc, int32_t m, int32_t n, void* stream) {
  return pg_linear_drop(dtype_in, dtype_out, srcs, n_src, bias, row_div, c, ldc, m, n, nullptr, 0, stream);
}

extern "C" int pg_linear_drop(int dtype_in, int dtype_out, const pg_gemm_src* srcs, int32_t n_src, const float* bias,
                              const float* row_div, void* c, int64_t ldc, int32_t m, int32_t n, const pg_drop* drop,
                              int64_t drop_row0, void* stream) {
  using namespace pg;
  PG_REQUIRE(srcs && c, "pg_linear: null operand");
  PG_REQUIRE(n_src >= 1 && n_src <= kMaxSrc, "pg_linear: 1..%d operand pairs, got %d", kMaxSrc, n_src);
  PG_REQUIRE(m >= 0 && n > 0 && n <= 256, "pg_linear: unsupported shape m=%d n=%d (n <= 256)", m, n);
  PG_REQUIRE(dtype_in == PG_F32 || dtype_in == PG_BF16, "pg_linear: bad input dtype");
  PG_REQUIRE(dtype_out == PG_F32 || dtype_out == PG_BF16, "pg_linear: bad output dtype");
  PG_REQUIRE(ldc >= n, "pg_linear: ldc < n");
  if (m == 0) return PG_OK;
  const bool tf32 = dtype_in == PG_F32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmArgs p;
  GemmMaps maps;
  p.m = m; p.n = n; p.n_pad = static_cast<int>(round_up(n, 16));
  p.n_src = n_src;
  p.k_elems = kRowBytes / (tf32 ? 4 : 2);
  int kb = 0, rc;
  for (int s = 0; s < kMaxSrc; ++s) {
    if (s < n_src) {
      PG_REQUIRE(srcs[s].a && srcs[s].b && srcs[s].k > 0, "pg_linear: operand pair %d is empty", s);
      kb += (srcs[s].k + p.k_elems - 1) / p.k_elems;
      if ((rc = make_map(&maps.a[s], srcs[s].a, srcs[s].lda, m, srcs[s].k, kBlockM, tf32)) != PG_OK) return rc;
      if ((rc = make_map(&maps.b[s], srcs[s].b, srcs[s].ldb, n, srcs[s].k, p.n_pad, tf32)) != PG_OK) return rc;
    } else {
      maps.a[s] = maps.a[0];
      maps.b[s] = maps.b[0];
    }
    p.kb_end[s] = kb;
  }
  p.bias = bias; p.row_div = row_div; p.c = c; p.ldc = ldc; p.out_bf16 = dtype_out == PG_BF16; p.is_tf32 = tf32;
  int cols = 32;
  while (cols < 2 * p.n_pad) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t fmt = tf32 ? 2u : 1u;        // UMMA F16F32Format: BF16 = 1, TF32 = 2
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(p.n_pad >> 3) << 17) |
            (static_cast<uint32_t>(kBlockM >> 4) << 24);
  if (g_sm_count == 0) {
    int dev = 0;
    PG_CHECK_CUDA(cudaGetDevice(&dev));
    PG_CHECK_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  {
    // TMA-store epilogue when the output rows are 16-byte aligned (always true for engine buffers)
    const int eso = p.out_bf16 ? 2 : 4;
    p.tma_store = ((reinterpret_cast<uintptr_t>(c) & 15) == 0 && (ldc * eso) % 16 == 0) ? 1 : 0;
    p.epi_batch = g_gemm_epi_batch;
    p.epi_slabs = g_gemm_epi_slabs;
    if (p.tma_store) {
      EncodeTiledFn enc = get_encode();
      cuuint64_t gdim[2] = {static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(m)};
      cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ldc) * eso};
      cuuint32_t box[2] = {static_cast<cuuint32_t>(kRowBytes / eso), 32};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&maps.c, p.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c,
                       gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(output) failed (%d)", (int)r); return PG_ERR_CUDA; }
    } else {
      maps.c = maps.a[0];
    }
  }
  p.drop = make_drop(drop);
  p.drop_row0 = drop_row0;
  p.drop_nvec = static_cast<int>(round_up(n, p.out_bf16 ? 8 : 4) / (p.out_bf16 ? 8 : 4));
  PG_REQUIRE(p.drop.thresh16 == 0u || p.tma_store, "pg_linear_drop: the fused dropout needs 16-byte aligned output rows");
  const int n_tiles = (m + kBlockM - 1) / kBlockM;
  const int grid = n_tiles < g_sm_count ? n_tiles : g_sm_count;
  const size_t smem = static_cast<size_t>(kStages) * (kABytes + p.n_pad * kRowBytes) + 8 * kSlabBytes /*staging*/ +
                      1024 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;
  static bool attr_set = false;       // once: opt in to the full 227 KB of shared memory (not a stream operation)
  if (!attr_set) {
    PG_CHECK_CUDA(cudaFuncSetAttribute(linear_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    PG_CHECK_CUDA(cudaFuncSetAttribute(linear_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  if (tf32) linear_tcgen05_kernel<true><<<grid, kGemmThreads, smem, st>>>(maps, p);
  else linear_tcgen05_kernel<false><<<grid, kGemmThreads, smem, st>>>(maps, p);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// hi = x with the 13 low mantissa bits cleared (exactly a tf32), lo = x - hi (exact in fp32)
namespace pg {
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ hi, float* __restrict__ lo, int64_t ld,
                  int rows, int d) {
  const int64_t total = static_cast<int64_t>(rows) * d;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / d), c = static_cast<int>(i % d);
    const float v = x[static_cast<int64_t>(r) * ldx + c];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[static_cast<int64_t>(r) * ld + c] = h;
    lo[static_cast<int64_t>(r) * ld + c] = v - h;
  }
}
}  // namespace pg

extern "C" int pg_split_tf32(const float* x, int64_t ldx, float* hi, float* lo, int64_t ld, int32_t rows, int32_t d,
                             void* stream) {
  PG_REQUIRE(x && hi && lo, "pg_split_tf32: null argument");
  PG_REQUIRE(rows >= 0 && d > 0 && ldx >= d && ld >= d, "pg_split_tf32: bad sizes");
  const int64_t total = static_cast<int64_t>(rows) * d;
  if (total == 0) return PG_OK;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  pg::split_tf32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, hi, lo, ld, rows, d);
  PG_LAUNCH_CHECK();
  return PG_OK;
}
