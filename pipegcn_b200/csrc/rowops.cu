// Row-wise epilogue kernels of the layer loop, sm_100a, HBM-bound:
//   * LayerNorm + ReLU forward in one pass              (/root/reference/module/model.py:53-56)
//   * its backward, fused with the three column reductions the step needs (d gamma, d beta and the
//     bias gradient of the linear that produced the pre-norm tensor)
//   * summed soft-max cross-entropy forward / backward  (/root/reference/train.py:320,351)
// One warp owns one row; a row of up to 32 lanes x kMaxVec 16-byte vectors stays in registers between
// the statistics pass and the output pass, so every tensor is read once and written once.
// Column reductions are two-stage and deterministic: per-CTA partials, then a fixed-order sum.
#include "common.cuh"

namespace pg {

constexpr int kRowThreads = 256;
constexpr int kMaxVec = 4;          // vectors per lane kept in registers (bf16: d <= 1024, fp32: d <= 512)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// VPL vectors per lane (d <= 32 * VPL * V), R rows per warp iteration: R * VPL 16-byte loads in flight per lane
template <typename T, int VPL, int R>
__global__ void __launch_bounds__(kRowThreads)
ln_relu_fwd_kernel(const T* __restrict__ y, int64_t ldy, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, int relu, T* __restrict__ out, int64_t ldo, float* __restrict__ mean_out,
                   float* __restrict__ rstd_out, int n_rows, int d, T* __restrict__ out_clean, int64_t ldc, DropArg drop) {
  using P = Pack<T, 16>;
  using Raw = typename P::Raw;
  constexpr int V = P::V;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kRowThreads) >> 5;
  const int nvec = d / V;
  const uint32_t seed_hi = drop_seed_hi(drop);
  float gm[VPL][V], bt[VPL][V];
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (lane + j * 32) * V + i;
      gm[j][i] = c < d ? __ldg(gamma + c) : 0.f;
      bt[j][i] = c < d ? __ldg(beta + c) : 0.f;
    }
  for (int row0 = ((blockIdx.x * kRowThreads + threadIdx.x) >> 5) * R; row0 < n_rows; row0 += warps * R) {
    Raw raw[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (row0 + r < n_rows && lane + j * 32 < nvec)
          raw[r][j] = *reinterpret_cast<const Raw*>(y + static_cast<int64_t>(row0 + r) * ldy + static_cast<int64_t>(lane + j * 32) * V);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r;
      if (row >= n_rows) break;
      float x[VPL][V];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (lane + j * 32 < nvec) {
          P::unpack(raw[r][j], x[j]);
#pragma unroll
          for (int i = 0; i < V; ++i) s += x[j][i];
        }
      const float mean = warp_sum(s) / d;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (lane + j * 32 < nvec)
#pragma unroll
          for (int i = 0; i < V; ++i) { const float c = x[j][i] - mean; q += c * c; }
      const float rstd = rsqrtf(warp_sum(q) / d + eps);
      if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (lane + j * 32 < nvec) {
          float o[V];
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float v = (x[j][i] - mean) * rstd * gm[j][i] + bt[j][i];
            o[i] = (relu && v < 0.f) ? 0.f : v;
          }
          Raw packed = P::pack(o);
          // the clean result (source of the halo push) when asked for, and dropout(result) for the next layer --
          // computed from the ROUNDED clean value, i.e. exactly what pg_dropout would make of the clean tensor
          if (out_clean != nullptr)
            st_vec<16>(out_clean + static_cast<int64_t>(row) * ldc + static_cast<int64_t>(lane + j * 32) * V, packed);
          if (drop.thresh16 != 0u) {
            P::unpack(packed, o);
            drop_apply<V>(o, static_cast<uint64_t>(row) * nvec + (lane + j * 32), drop.thresh16, drop.scale, drop.seed_lo, seed_hi);
            packed = P::pack(o);
          }
          st_vec<16>(out + static_cast<int64_t>(row) * ldo + static_cast<int64_t>(lane + j * 32) * V, packed);
        }
    }
  }
}

// g_y = rstd * (gh*gamma - mean_d(gh*gamma) - xhat * mean_d(gh*gamma*xhat)),  gh = g_out * (out > 0)
// partial[blockIdx][0] += gh * xhat (d gamma), [1] += gh (d beta), [2] += g_y (bias gradient upstream)
template <typename T, int VPL, int R>
__global__ void __launch_bounds__(kRowThreads)
ln_relu_bwd_kernel(const T* __restrict__ g_out, int64_t ldg, const T* __restrict__ out, int64_t ldo,
                   const T* __restrict__ y, int64_t ldy, const float* __restrict__ mean, const float* __restrict__ rstd,
                   const float* __restrict__ gamma, int relu, T* __restrict__ g_y, int64_t ldgy,
                   float* __restrict__ partial, int n_rows, int d, const float* __restrict__ beta) {
  using P = Pack<T, 16>;
  using Raw = typename P::Raw;
  constexpr int V = P::V;
  // beta given: the ReLU mask is recomputed from y (same expression as the forward) instead of read from `out` --
  // one [N, d] tensor less to read
  const bool remask = relu && beta != nullptr;
  extern __shared__ float red[];                             // [3][d]
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kRowThreads) >> 5;
  const int nvec = d / V;
  float gm[VPL][V], bt[VPL][V], cg[VPL][V], cb[VPL][V], cy[VPL][V];      // gamma, beta and this lane's column partials
#pragma unroll
  for (int j = 0; j < VPL; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (lane + j * 32) * V + i;
      gm[j][i] = c < d ? __ldg(gamma + c) : 0.f;
      bt[j][i] = (remask && c < d) ? __ldg(beta + c) : 0.f;
      cg[j][i] = cb[j][i] = cy[j][i] = 0.f;
    }

  for (int row0 = ((blockIdx.x * kRowThreads + threadIdx.x) >> 5) * R; row0 < n_rows; row0 += warps * R) {
    Raw rg[R][VPL], ry[R][VPL], ro[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (row0 + r < n_rows && lane + j * 32 < nvec) {
          const int64_t c = static_cast<int64_t>(lane + j * 32) * V;
          rg[r][j] = *reinterpret_cast<const Raw*>(g_out + static_cast<int64_t>(row0 + r) * ldg + c);
          ry[r][j] = *reinterpret_cast<const Raw*>(y + static_cast<int64_t>(row0 + r) * ldy + c);
          if (relu && !remask) ro[r][j] = *reinterpret_cast<const Raw*>(out + static_cast<int64_t>(row0 + r) * ldo + c);
        }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r;
      if (row >= n_rows) break;
      const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
      float gx[VPL][V], xh[VPL][V];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (lane + j * 32 < nvec) {
          float g[V], o[V], yy[V];
          P::unpack(rg[r][j], g);
          P::unpack(ry[r][j], yy);
          if (relu && !remask) P::unpack(ro[r][j], o);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float xx = (yy[i] - mu) * rs;
            const bool dead = remask ? !((yy[i] - mu) * rs * gm[j][i] + bt[j][i] > 0.f) : (relu && !(o[i] > 0.f));
            const float gg = dead ? 0.f : g[i];
            xh[j][i] = xx;
            gx[j][i] = gg * gm[j][i];
            s1 += gx[j][i];
            s2 += gx[j][i] * xx;
            cg[j][i] += gg * xx;
            cb[j][i] += gg;
          }
        }
      const float c1 = warp_sum(s1) / d, c2 = warp_sum(s2) / d;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (lane + j * 32 < nvec) {
          float o[V];
#pragma unroll
          for (int i = 0; i < V; ++i) {
            o[i] = rs * (gx[j][i] - c1 - xh[j][i] * c2);
            // the tensor handed upstream is stored in T: reduce what is stored
            cy[j][i] += (sizeof(T) == 2) ? __bfloat162float(__float2bfloat16_rn(o[i])) : o[i];
          }
          st_vec<16>(g_y + static_cast<int64_t>(row) * ldgy + static_cast<int64_t>(lane + j * 32) * V, P::pack(o));
        }
    }
  }
  // CTA-level reduction of the column partials in a fixed order (warp 0..7), then one partial row per CTA
  for (int i = threadIdx.x; i < 3 * d; i += kRowThreads) red[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < kRowThreads / 32; ++w) {
    if ((threadIdx.x >> 5) == w) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int vi = lane + j * 32;
        if (vi < nvec)
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int c = vi * V + i;
            red[c] += cg[j][i];
            red[d + c] += cb[j][i];
            red[2 * d + c] += cy[j][i];
          }
      }
    }
    __syncthreads();
  }
  float* pp = partial + static_cast<int64_t>(blockIdx.x) * 3 * d;
  for (int i = threadIdx.x; i < 3 * d; i += kRowThreads) pp[i] = red[i];
}

// out[k] = sum over blocks of partial[b][k]: one warp per column, lanes stride over the blocks, fixed order
__global__ void colsum_final_kernel(const float* __restrict__ partial, int n_blocks, int width, float* __restrict__ out0,
                                    float* __restrict__ out1, float* __restrict__ out2, int d) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= width) return;
  float s = 0.f;
  for (int b = lane; b < n_blocks; b += 32) s += partial[static_cast<int64_t>(b) * width + k];
  s = warp_sum(s);
  float* o = (k < d) ? out0 : (k < 2 * d ? out1 : out2);
  if (lane == 0 && o != nullptr) o[k % d] = s;
}

// ---------------------------------------------------------------------------------------------------------
// loss = sum_rows (logsumexp(z) - z[label]);  lse[row] kept for the backward
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
ce_fwd_kernel(const T* __restrict__ z, int64_t ld, const int64_t* __restrict__ labels, int n_rows, int c,
              float* __restrict__ lse, float* __restrict__ partial) {
  __shared__ float wsum[kRowThreads / 32];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kRowThreads) >> 5;
  float acc = 0.f;
  for (int row = (blockIdx.x * kRowThreads + threadIdx.x) >> 5; row < n_rows; row += warps) {
    const T* zp = z + static_cast<int64_t>(row) * ld;
    float m = -INFINITY;
    for (int k = lane; k < c; k += 32) m = fmaxf(m, static_cast<float>(zp[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int k = lane; k < c; k += 32) s += expf(static_cast<float>(zp[k]) - m);
    s = warp_sum(s);
    const float l = m + logf(s);
    if (lane == 0) {
      lse[row] = l;
      acc += l - static_cast<float>(zp[labels[row]]);
    }
  }
  if (lane == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kRowThreads / 32; ++w) s += wsum[w];
    partial[blockIdx.x] = s;
  }
}

// g[row, k] = (exp(z - lse) - [k == label]) * upstream   for row < n_rows, zero for n_rows <= row < n_total
template <typename T>
__global__ void __launch_bounds__(kRowThreads)
ce_bwd_kernel(const T* __restrict__ z, int64_t ld, const int64_t* __restrict__ labels, const float* __restrict__ lse,
              const float* __restrict__ upstream, int n_rows, int n_total, int c, T* __restrict__ g, int64_t ldg,
              float* __restrict__ partial) {
  extern __shared__ float red[];                             // [c]
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * kRowThreads) >> 5;
  const float up = upstream ? __ldg(upstream) : 1.f;
  for (int i = threadIdx.x; i < c; i += kRowThreads) red[i] = 0.f;
  __syncthreads();
  for (int row = (blockIdx.x * kRowThreads + threadIdx.x) >> 5; row < n_total; row += warps) {
    T* gp = g + static_cast<int64_t>(row) * ldg;
    if (row >= n_rows) {
      for (int k = lane; k < c; k += 32) gp[k] = static_cast<T>(0.f);
      continue;
    }
    const T* zp = z + static_cast<int64_t>(row) * ld;
    const float l = lse[row];
    const int lab = static_cast<int>(labels[row]);
    for (int k = lane; k < c; k += 32) {
      const float v = (expf(static_cast<float>(zp[k]) - l) - (k == lab ? 1.f : 0.f)) * up;
      const T t = static_cast<T>(v);
      gp[k] = t;
      atomicAdd(&red[k], static_cast<float>(t));             // shared-memory partial; order fixed within a warp only
    }
  }
  __syncthreads();
  float* pp = partial + static_cast<int64_t>(blockIdx.x) * c;
  for (int i = threadIdx.x; i < c; i += kRowThreads) pp[i] = red[i];
}

// ---------------------------------------------------------------------------------------------------------
// Dropout with a counter-based generator: keep(i) is a pure function of (seed, element index), so the backward
// regenerates the mask instead of storing it.  out = keep ? x / (1 - p) : 0.  16 random bits per element.
template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ out, int64_t ldo, int n_rows, int nvec,
               int64_t row0, DropArg drop) {
  using P = Pack<T, 16>;
  const uint32_t seed_hi = drop_seed_hi(drop);                         // device-side epoch counter (graph replay)
  const uint32_t thresh16 = drop.thresh16, seed_lo = drop.seed_lo;
  const float scale = drop.scale;
  constexpr int V = P::V;
  const int64_t total = static_cast<int64_t>(n_rows) * nvec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / nvec), vi = static_cast<int>(i % nvec);
    float f[V];
    P::unpack(*reinterpret_cast<const typename P::Raw*>(x + static_cast<int64_t>(r) * ldx + static_cast<int64_t>(vi) * V), f);
    drop_apply<V>(f, static_cast<uint64_t>(i) + static_cast<uint64_t>(row0) * nvec, thresh16, scale, seed_lo, seed_hi);
    st_vec<16>(out + static_cast<int64_t>(r) * ldo + static_cast<int64_t>(vi) * V, P::pack(f));
  }
}

static int row_grid(int n_rows) {
  const int need = (n_rows + (kRowThreads / 32) - 1) / (kRowThreads / 32);
  return need < 148 * 8 ? (need > 0 ? need : 1) : 148 * 8;
}

}  // namespace pg

extern "C" int pg_row_grid(int32_t n_rows) { return pg::row_grid(n_rows); }

extern "C" int pg_dropout_rows(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t row0, int32_t n_rows, int32_t d,
                               int dtype, const pg_drop* drop, void* stream) {
  using namespace pg;
  PG_REQUIRE(x && out && drop && drop->p >= 0.f && drop->p < 1.f, "pg_dropout: bad argument");
  const int es = elem_size(dtype), v = 16 / es;
  PG_REQUIRE(d > 0 && round_up(d, v) <= ldx && round_up(d, v) <= ldo && vec_bytes(x, ldx, es) == 16 && vec_bytes(out, ldo, es) == 16,
             "pg_dropout: rows must be 16-byte aligned and padded to the vector width");
  const int nvec = static_cast<int>(round_up(d, v) / v);
  const int64_t total = static_cast<int64_t>(n_rows) * nvec;
  if (total == 0) return PG_OK;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  DropArg a = make_drop(drop);
  if (drop->p == 0.f) { a.thresh16 = 0u; a.scale = 1.f; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == PG_F32)
    dropout_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), ldx, static_cast<float*>(out), ldo, n_rows, nvec, row0, a);
  else
    dropout_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, n_rows, nvec, row0, a);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int pg_dropout(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t n_rows, int32_t d, int dtype,
                          float p, uint64_t seed, const uint32_t* step_dev, void* stream) {
  pg_drop dr{p, seed, step_dev, 0};
  return pg_dropout_rows(x, ldx, out, ldo, 0, n_rows, d, dtype, &dr, stream);
}

extern "C" int pg_ln_relu_fwd(const void* y, int64_t ldy, const float* gamma, const float* beta, float eps, int relu,
                              void* out, int64_t ldo, float* mean, float* rstd, int32_t n_rows, int32_t d, int dtype,
                              void* stream) {
  return pg_ln_relu_drop_fwd(y, ldy, gamma, beta, eps, relu, out, ldo, nullptr, 0, mean, rstd, n_rows, d, dtype, nullptr, stream);
}

extern "C" int pg_ln_relu_drop_fwd(const void* y, int64_t ldy, const float* gamma, const float* beta, float eps, int relu,
                                   void* out, int64_t ldo, void* out_clean, int64_t ldc, float* mean, float* rstd,
                                   int32_t n_rows, int32_t d, int dtype, const pg_drop* drop, void* stream) {
  using namespace pg;
  PG_REQUIRE(y && gamma && beta && out && mean && rstd, "pg_ln_relu_fwd: null argument");
  PG_REQUIRE(out_clean == nullptr || (ldc >= d && vec_bytes(out_clean, ldc, elem_size(dtype)) == 16),
             "pg_ln_relu_drop_fwd: out_clean rows must be 16-byte aligned");
  const DropArg da = make_drop(drop);
  const int v = 16 / elem_size(dtype);
  PG_REQUIRE(d > 0 && d % v == 0 && d / v <= 32 * kMaxVec, "pg_ln_relu_fwd: d=%d must be a multiple of %d and <= %d", d, v, 32 * kMaxVec * v);
  PG_REQUIRE(vec_bytes(y, ldy, elem_size(dtype)) == 16 && vec_bytes(out, ldo, elem_size(dtype)) == 16, "pg_ln_relu_fwd: rows must be 16-byte aligned");
  if (n_rows == 0) return PG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(n_rows);
  const int vpl = (d / v + 31) / 32;
#define PG_LNF(T_, VPL_, R_) ln_relu_fwd_kernel<T_, VPL_, R_><<<grid, kRowThreads, 0, st>>>(static_cast<const T_*>(y), ldy, gamma, beta, eps, relu, static_cast<T_*>(out), ldo, mean, rstd, n_rows, d, static_cast<T_*>(out_clean), ldc, da)
  if (dtype == PG_F32) {
    if (vpl <= 1) PG_LNF(float, 1, 4); else if (vpl <= 2) PG_LNF(float, 2, 2); else PG_LNF(float, 4, 1);
  } else {
    if (vpl <= 1) PG_LNF(__nv_bfloat16, 1, 4); else if (vpl <= 2) PG_LNF(__nv_bfloat16, 2, 2); else PG_LNF(__nv_bfloat16, 4, 1);
  }
#undef PG_LNF
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int pg_ln_relu_bwd(const void* g_out, int64_t ldg, const void* out, int64_t ldo, const void* y, int64_t ldy,
                              const float* mean, const float* rstd, const float* gamma, int relu, void* g_y,
                              int64_t ldgy, float* dgamma, float* dbeta, float* colsum, float* partial,
                              int32_t n_rows, int32_t d, int dtype, void* stream) {
  return pg_ln_relu_bwd2(g_out, ldg, out, ldo, y, ldy, mean, rstd, gamma, nullptr, relu, g_y, ldgy, dgamma, dbeta, colsum,
                         partial, n_rows, d, dtype, stream);
}

extern "C" int pg_ln_relu_bwd2(const void* g_out, int64_t ldg, const void* out, int64_t ldo, const void* y, int64_t ldy,
                               const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                               void* g_y, int64_t ldgy, float* dgamma, float* dbeta, float* colsum, float* partial,
                               int32_t n_rows, int32_t d, int dtype, void* stream) {
  using namespace pg;
  PG_REQUIRE(g_out && y && mean && rstd && gamma && g_y && partial, "pg_ln_relu_bwd: null argument");
  PG_REQUIRE(!relu || out || beta, "pg_ln_relu_bwd: relu needs the forward output or beta");
  const int es = elem_size(dtype), v = 16 / es;
  PG_REQUIRE(d > 0 && d % v == 0 && d / v <= 32 * kMaxVec, "pg_ln_relu_bwd: unsupported d=%d", d);
  PG_REQUIRE(vec_bytes(g_out, ldg, es) == 16 && vec_bytes(y, ldy, es) == 16 && vec_bytes(g_y, ldgy, es) == 16 &&
             (!relu || beta || vec_bytes(out, ldo, es) == 16), "pg_ln_relu_bwd: rows must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(n_rows);
  const size_t smem = 3 * static_cast<size_t>(d) * sizeof(float);
  const int vpl = (d / v + 31) / 32;
#define PG_LNB(T_, VPL_, R_) ln_relu_bwd_kernel<T_, VPL_, R_><<<grid, kRowThreads, smem, st>>>(static_cast<const T_*>(g_out), ldg, static_cast<const T_*>(out), ldo, static_cast<const T_*>(y), ldy, mean, rstd, gamma, relu, static_cast<T_*>(g_y), ldgy, partial, n_rows, d, beta)
  if (dtype == PG_F32) {
    if (vpl <= 1) PG_LNB(float, 1, 2); else if (vpl <= 2) PG_LNB(float, 2, 1); else PG_LNB(float, 4, 1);
  } else {
    if (vpl <= 1) PG_LNB(__nv_bfloat16, 1, 2); else if (vpl <= 2) PG_LNB(__nv_bfloat16, 2, 1); else PG_LNB(__nv_bfloat16, 4, 1);
  }
#undef PG_LNB
  PG_LAUNCH_CHECK();
  colsum_final_kernel<<<(3 * d * 32 + 255) / 256, 256, 0, st>>>(partial, grid, 3 * d, dgamma, dbeta, colsum, d);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int pg_ce_fwd(const void* z, int64_t ld, const int64_t* labels, int32_t n_rows, int32_t c, int dtype,
                         float* lse, float* partial, float* loss, void* stream) {
  using namespace pg;
  PG_REQUIRE(z && labels && lse && partial && loss && c > 0, "pg_ce_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(n_rows);
  if (dtype == PG_F32) ce_fwd_kernel<float><<<grid, kRowThreads, 0, st>>>(static_cast<const float*>(z), ld, labels, n_rows, c, lse, partial);
  else ce_fwd_kernel<__nv_bfloat16><<<grid, kRowThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(z), ld, labels, n_rows, c, lse, partial);
  PG_LAUNCH_CHECK();
  colsum_final_kernel<<<1, 32, 0, st>>>(partial, grid, 1, loss, nullptr, nullptr, 1);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int pg_ce_bwd(const void* z, int64_t ld, const int64_t* labels, const float* lse, const float* upstream,
                         int32_t n_rows, int32_t n_total, int32_t c, int dtype, void* g, int64_t ldg, float* colsum,
                         float* partial, void* stream) {
  using namespace pg;
  PG_REQUIRE(z && labels && lse && g && partial && c > 0 && n_total >= n_rows, "pg_ce_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(n_total);
  const size_t smem = static_cast<size_t>(c) * sizeof(float);
  if (dtype == PG_F32) ce_bwd_kernel<float><<<grid, kRowThreads, smem, st>>>(static_cast<const float*>(z), ld, labels, lse, upstream, n_rows, n_total, c, static_cast<float*>(g), ldg, partial);
  else ce_bwd_kernel<__nv_bfloat16><<<grid, kRowThreads, smem, st>>>(static_cast<const __nv_bfloat16*>(z), ld, labels, lse, upstream, n_rows, n_total, c, static_cast<__nv_bfloat16*>(g), ldg, partial);
  PG_LAUNCH_CHECK();
  if (colsum != nullptr) {
    colsum_final_kernel<<<(c * 32 + 255) / 256, 256, 0, st>>>(partial, grid, c, colsum, nullptr, nullptr, c);
    PG_LAUNCH_CHECK();
  }
  return PG_OK;
}
