// LayerNorm(eps=1e-6) forward / backward (K7, K13; modules/transformer.py:128,132, audio_transformer.py:113).
// One warp per row, row kept in registers (D = 128*NV, NV <= 8), two-pass mean/variance like ATen.
// Strided rows let the same kernels normalise only the CLS row of every sequence (K13).
#include "common.cuh"

namespace atst {

template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, long long x_stride, const float* __restrict__ gamma,
              const float* __restrict__ beta, float* __restrict__ y, long long y_stride, float* __restrict__ mean,
              float* __restrict__ rstd, int rows, float eps, int round_out) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  for (int row = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_cta) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * x_stride);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = xr[lane + 32 * i];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
      q += a * a + b * b + c * c + d * d;
    }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
    float4* yr = reinterpret_cast<float4*>(y + static_cast<long long>(row) * y_stride);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 gm = reinterpret_cast<const float4*>(gamma)[lane + 32 * i];
      const float4 bt = reinterpret_cast<const float4*>(beta)[lane + 32 * i];
      float4 o;
      o.x = (v[i].x - mu) * rs * gm.x + bt.x;
      o.y = (v[i].y - mu) * rs * gm.y + bt.y;
      o.z = (v[i].z - mu) * rs * gm.z + bt.z;
      o.w = (v[i].w - mu) * rs * gm.w + bt.w;
      if (round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      yr[lane + 32 * i] = o;
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (per-lane register partials -> smem -> atomics)
// WPR warps share a row (WPR = 2 for the wide rows: half the per-lane state, so two CTAs fit per SM and twice as
// many loads are in flight; the two warps exchange their partial row sums through shared memory).
template <int NV, int WPR>
__global__ void __launch_bounds__(256, WPR == 2 ? 2 : 1)
ln_bwd_kernel(const float* __restrict__ dy, long long dy_stride, const float* __restrict__ x, long long x_stride,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* __restrict__ dres, long long dres_stride, float* __restrict__ dx, long long dx_stride,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
              float* __restrict__ dys, long long dys_stride, const float* __restrict__ rowscale, int rows_per_seq,
              float* __restrict__ colsum_out) {
  // optional second output for the consumer branch of dx: dys = tf32(rowscale[row / rows_per_seq] * dx) (the
  // DropPath-scaled, GEMM-ready copy) and colsum_out += column sums of dys (= the consumer Linear's bias gradient)
  constexpr int D = NV * 128;
  constexpr int NW = NV / WPR;  // float4 slabs per lane
  static_assert(NV % WPR == 0, "row slabs must split evenly over the warps of a row");
  __shared__ float red[3][8][32 * 4];  // [dgamma|dbeta|colsum][warp][lane*4] for one slab at a time
  __shared__ float xch[2][8][2];       // [parity][warp][s1, s2] partial row sums (WPR == 2)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = warp % WPR, slot = warp / WPR;  // this warp's part of the row / row slot inside the CTA
  constexpr int kRowsPerCta = 8 / WPR;
  float4 pg[NW], pb[NW], pc[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) pg[i] = pb[i] = pc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gm[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) gm[i] = reinterpret_cast<const float4*>(gamma)[lane + 32 * (sub * NW + i)];

  int it = 0;
  // every warp of a row slot runs the same number of iterations (the pair barrier below needs both warps)
  for (int row = blockIdx.x * kRowsPerCta + slot; row < rows; row += gridDim.x * kRowsPerCta, ++it) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * x_stride) + 32 * sub * NW;
    const float4* dr = reinterpret_cast<const float4*>(dy + static_cast<long long>(row) * dy_stride) + 32 * sub * NW;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NW], g[NW], rv[NW];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const float4 xv = xr[lane + 32 * i];
      const float4 dv = dr[lane + 32 * i];
      if (dres != nullptr)
        rv[i] = (reinterpret_cast<const float4*>(dres + static_cast<long long>(row) * dres_stride) + 32 * sub * NW)[lane + 32 * i];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
      s1 += g[i].x + g[i].y + g[i].z + g[i].w;
      s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      pg[i].x += dv.x * xh[i].x; pg[i].y += dv.y * xh[i].y; pg[i].z += dv.z * xh[i].z; pg[i].w += dv.w * xh[i].w;
      pb[i].x += dv.x; pb[i].y += dv.y; pb[i].z += dv.z; pb[i].w += dv.w;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (WPR == 2) {
      float* xc = &xch[it & 1][0][0];
      if (lane == 0) { xc[warp * 2] = s1; xc[warp * 2 + 1] = s2; }
      asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");  // the two warps of this row
      s1 += xc[(warp ^ 1) * 2];
      s2 += xc[(warp ^ 1) * 2 + 1];
    }
    s1 *= (1.0f / D);
    s2 *= (1.0f / D);
    float4* ox = reinterpret_cast<float4*>(dx + static_cast<long long>(row) * dx_stride) + 32 * sub * NW;
    const float sc = (dys != nullptr && rowscale != nullptr) ? rowscale[row / rows_per_seq] : 1.0f;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      float4 o;
      o.x = rs * (g[i].x - s1 - xh[i].x * s2);
      o.y = rs * (g[i].y - s1 - xh[i].y * s2);
      o.z = rs * (g[i].z - s1 - xh[i].z * s2);
      o.w = rs * (g[i].w - s1 - xh[i].w * s2);
      if (dres != nullptr) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
      ox[lane + 32 * i] = o;
      if (dys != nullptr) {
        float4 y = make_float4(round_tf32(sc * o.x), round_tf32(sc * o.y), round_tf32(sc * o.z), round_tf32(sc * o.w));
        (reinterpret_cast<float4*>(dys + static_cast<long long>(row) * dys_stride) + 32 * sub * NW)[lane + 32 * i] = y;
        pc[i].x += y.x; pc[i].y += y.y; pc[i].z += y.z; pc[i].w += y.w;
      }
    }
  }
  // cross-warp reduction of the parameter-gradient partials, one 128-column slab per `sub` at a time
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    *reinterpret_cast<float4*>(&red[0][warp][lane * 4]) = pg[i];
    *reinterpret_cast<float4*>(&red[1][warp][lane * 4]) = pb[i];
    *reinterpret_cast<float4*>(&red[2][warp][lane * 4]) = pc[i];
    __syncthreads();
    {
      const int which = threadIdx.x >> 7, col = threadIdx.x & 127;  // 256 threads: dgamma | dbeta
#pragma unroll
      for (int sb = 0; sb < WPR; ++sb) {
        float acc = 0.f;
        for (int w = sb; w < 8; w += WPR) acc += red[which][w][col];
        // column of this slab: float4 index (lane + 32 (sb NW + i)) -> element (lane*4 + j) + 128 (sb NW + i)
        atomicAdd((which ? dbeta : dgamma) + 128 * (sb * NW + i) + col, acc);
        if (colsum_out != nullptr && which == 0) {
          float c = 0.f;
          for (int w = sb; w < 8; w += WPR) c += red[2][w][col];
          atomicAdd(colsum_out + 128 * (sb * NW + i) + col, c);
        }
      }
    }
    __syncthreads();
  }
}

template <int NV>
static int ln_fwd_launch(const float* x, long long xs, const float* g, const float* b, float* y, long long ys,
                         float* mean, float* rstd, int rows, float eps, int round_out, cudaStream_t st) {
  int grid = (rows + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  ln_fwd_kernel<NV><<<grid, 256, 0, st>>>(x, xs, g, b, y, ys, mean, rstd, rows, eps, round_out);
  return atst_check_launch("ln_fwd_kernel");
}
template <int NV>
static int ln_bwd_launch(const float* dy, long long dys, const float* x, long long xs, const float* mean,
                         const float* rstd, const float* g, const float* dres, long long drs, float* dx,
                         long long dxs, float* dg, float* db, int rows, float* dys2, long long dys2s,
                         const float* rowscale, int rps, float* colsum_out, cudaStream_t st) {
  constexpr int WPR = (NV % 2 == 0 && NV >= 6) ? 2 : 1;
  constexpr int kRowsPerCta = 8 / WPR;
  int grid = (rows + kRowsPerCta - 1) / kRowsPerCta;
  if (grid > 148 * 2) grid = 148 * 2;
  ln_bwd_kernel<NV, WPR><<<grid, 256, 0, st>>>(dy, dys, x, xs, mean, rstd, g, dres, drs, dx, dxs, dg, db, rows, dys2,
                                              dys2s, rowscale, rps > 0 ? rps : 1, colsum_out);
  return atst_check_launch("ln_bwd_kernel");
}

int layernorm_forward(const float* x, long long x_stride, const float* gamma, const float* beta, float* y,
                      long long y_stride, float* mean, float* rstd, int rows, int D, float eps, int round_out,
                      cudaStream_t st) {
  ATST_REQUIRE(rows > 0 && D % 128 == 0 && D <= 1024, "layernorm: D must be a multiple of 128 <= 1024, got %d", D);
  ATST_REQUIRE(x_stride % 4 == 0 && y_stride % 4 == 0, "layernorm: row strides must be multiples of 4");
  switch (D / 128) {
    case 1: return ln_fwd_launch<1>(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, eps, round_out, st);
    case 2: return ln_fwd_launch<2>(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, eps, round_out, st);
    case 3: return ln_fwd_launch<3>(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, eps, round_out, st);
    case 4: return ln_fwd_launch<4>(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, eps, round_out, st);
    case 6: return ln_fwd_launch<6>(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, eps, round_out, st);
    case 8: return ln_fwd_launch<8>(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, eps, round_out, st);
    default: atst_set_error("layernorm: unsupported D=%d", D); return ATST_ERR_ARG;
  }
}

int layernorm_backward(const float* dy, long long dy_stride, const float* x, long long x_stride, const float* mean,
                       const float* rstd, const float* gamma, const float* dres, long long dres_stride, float* dx,
                       long long dx_stride, float* dgamma, float* dbeta, int rows, int D, float* dys,
                       long long dys_stride, const float* rowscale, int rows_per_seq, float* colsum_out,
                       cudaStream_t st) {
  ATST_REQUIRE(rows > 0 && D % 128 == 0 && D <= 1024, "layernorm_backward: unsupported D=%d", D);
  switch (D / 128) {
    case 1: return ln_bwd_launch<1>(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma, dbeta, rows, dys, dys_stride, rowscale, rows_per_seq, colsum_out, st);
    case 2: return ln_bwd_launch<2>(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma, dbeta, rows, dys, dys_stride, rowscale, rows_per_seq, colsum_out, st);
    case 3: return ln_bwd_launch<3>(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma, dbeta, rows, dys, dys_stride, rowscale, rows_per_seq, colsum_out, st);
    case 4: return ln_bwd_launch<4>(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma, dbeta, rows, dys, dys_stride, rowscale, rows_per_seq, colsum_out, st);
    case 6: return ln_bwd_launch<6>(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma, dbeta, rows, dys, dys_stride, rowscale, rows_per_seq, colsum_out, st);
    case 8: return ln_bwd_launch<8>(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma, dbeta, rows, dys, dys_stride, rowscale, rows_per_seq, colsum_out, st);
    default: atst_set_error("layernorm_backward: unsupported D=%d", D); return ATST_ERR_ARG;
  }
}

}  // namespace atst
