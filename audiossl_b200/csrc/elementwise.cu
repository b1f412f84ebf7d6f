// Memory-bound helpers around the GEMMs: patch extraction (K5), token assembly (K6), column sums for bias
// gradients, BatchNorm1d(+ReLU) of the projector/predictor (K14), small copies.
#include "common.cuh"

namespace atst {

// ------------------------------------------------------------------ PatchEmbed_v2 rearrange (audio_transformer.py:64-75)
// mel [S,1,64,T] -> patches [S*P, 256], column p1*4+p2 = mel[s, p1, 4w+p2]; P = T/4 (remainder frames dropped)
__global__ void patchify_kernel(const float* __restrict__ mel, long long clip_stride, int T, int P,
                                float* __restrict__ patches, long long total /* S*P*64 */) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p1 = static_cast<int>(i & 63);
    const long long sw = i >> 6;
    const int w = static_cast<int>(sw % P);
    const long long s = sw / P;
    const float* src = mel + s * clip_stride + static_cast<long long>(p1) * T + 4 * w;
    float4 v = make_float4(round_tf32(src[0]), round_tf32(src[1]), round_tf32(src[2]), round_tf32(src[3]));
    reinterpret_cast<float4*>(patches)[i] = v;  // (sw*256 + p1*4)/4 == i
  }
}

// x[s, 0] = cls + pos[0]; x[s, 1+w] = pe[s*P+w] (or mask_embed where mask[s,w]) + pos[1+w]      (use_cls)
// x[s, w] = pe/mask_embed + pos[1+w]                                                            (frame model)
__global__ void tokens_fwd_kernel(const float* __restrict__ pe, const float* __restrict__ cls,
                                  const float* __restrict__ pos, const float* __restrict__ mask_embed,
                                  const unsigned char* __restrict__ mask, float* __restrict__ x, int S, int P, int D,
                                  int use_cls) {
  const int N = P + (use_cls ? 1 : 0);
  const int d4 = D / 4;
  const long long total = static_cast<long long>(S) * N * d4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d4);
    const long long sn = i / d4;
    const int n = static_cast<int>(sn % N);
    const long long s = sn / N;
    float4 v;
    int posrow;
    if (use_cls && n == 0) {
      v = reinterpret_cast<const float4*>(cls)[c];
      posrow = 0;
    } else {
      const int w = use_cls ? n - 1 : n;
      posrow = w + 1;
      if (mask != nullptr && mask[s * P + w]) v = reinterpret_cast<const float4*>(mask_embed)[c];
      else v = reinterpret_cast<const float4*>(pe)[(s * P + w) * d4 + c];
    }
    const float4 pz = reinterpret_cast<const float4*>(pos)[static_cast<long long>(posrow) * d4 + c];
    v.x += pz.x; v.y += pz.y; v.z += pz.z; v.w += pz.w;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

// dpe[s*P+w] = dx[s, 1+w] (0 where masked); dpos[n'] += sum_s dx[s, n]; dcls += sum_s dx[s,0];
// dmask_embed += sum over masked (s,w) of dx.  One thread per (n, column) pair looping over sequences.
__global__ void tokens_bwd_kernel(const float* __restrict__ dx, const unsigned char* __restrict__ mask,
                                  float* __restrict__ dpe, float* __restrict__ dpos, float* __restrict__ dcls,
                                  float* __restrict__ dmask_embed, int S, int P, int D, int use_cls) {
  const int N = P + (use_cls ? 1 : 0);
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(N) * D) return;
  const int n = static_cast<int>(i / D), d = static_cast<int>(i % D);
  float acc = 0.f, macc = 0.f;
  const bool is_cls = use_cls && n == 0;
  const int w = use_cls ? n - 1 : n;
  for (int s = 0; s < S; ++s) {
    const float g = dx[(static_cast<long long>(s) * N + n) * D + d];
    acc += g;
    if (!is_cls) {
      const bool m = mask != nullptr && mask[static_cast<long long>(s) * P + w];
      dpe[(static_cast<long long>(s) * P + w) * D + d] = m ? 0.f : round_tf32(g);
      if (m) macc += g;
    }
  }
  if (is_cls) {
    atomicAdd(dcls + d, acc);
    atomicAdd(dpos + d, acc);
  } else {
    atomicAdd(dpos + static_cast<long long>(w + 1) * D + d, acc);
    if (mask != nullptr && dmask_embed != nullptr) atomicAdd(dmask_embed + d, macc);
  }
}

// out[c] += sum_r X[r, c]; CTA = 32 columns x 8 row lanes, grid.y splits the rows
__global__ void colsum_kernel(const float* __restrict__ X, long long ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, rows);
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + ry; r < r1; r += 8) acc += X[static_cast<long long>(r) * ld + c];
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][cx];
    atomicAdd(out + c, s);
  }
}

// same, 128 columns per CTA (float4 per lane), four rows in flight per thread; cols and ld multiples of 4
__global__ void __launch_bounds__(256)
colsum4_kernel(const float* __restrict__ X, long long ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = blockIdx.x * 32 + lane;
  const int rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, rows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < cols) {
    for (int r = r0 + warp; r < r1; r += 32) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = r + 8 * k < r1;
        v[k] = ok ? __ldcs(reinterpret_cast<const float4*>(X + static_cast<long long>(r + 8 * k) * ld) + c4)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
  }
  *reinterpret_cast<float4*>(&red[warp][lane * 4]) = acc;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col < cols) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      atomicAdd(out + col, s);
    }
  }
}

// ------------------------------------------------------------------ BatchNorm1d (train mode) + ReLU
// pass 1: per-column mean and M2 = sum (x - mean)^2 over the local rows (two-pass, like ATen)
__global__ void bn_stats_kernel(const float* __restrict__ X, int rows, int cols, float* __restrict__ mean,
                                float* __restrict__ m2) {
  __shared__ float red[8][33];
  __shared__ float smean[32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (c < cols)
    for (int r = ry; r < rows; r += 8) acc += X[static_cast<long long>(r) * cols + c];
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][cx];
    smean[cx] = s / rows;
  }
  __syncthreads();
  const float mu = smean[cx];
  acc = 0.f;
  if (c < cols)
    for (int r = ry; r < rows; r += 8) {
      const float d = X[static_cast<long long>(r) * cols + c] - mu;
      acc += d * d;
    }
  __syncthreads();
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += red[j][cx];
    mean[c] = mu;
    m2[c] = s;
  }
}

// finalize: rstd from (global) mean / M2 / count, running-stat update with momentum (unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ mean, const float* __restrict__ m2, float count,
                                   float eps, float momentum, float* __restrict__ rstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const float var = m2[c] / count;
  rstd[c] = rsqrtf(var + eps);
  if (running_mean != nullptr) {
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean[c];
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (m2[c] / fmaxf(count - 1.0f, 1.0f));
  }
}

// y = relu((x - mean) * rstd * gamma + beta), rounded to tf32 (it feeds the next Linear)
__global__ void bn_relu_fwd_kernel(const float* __restrict__ X, const float* __restrict__ mean,
                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ Y, long long total4, int cols4,
                                   int round_out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cols4);
    const float4 x = reinterpret_cast<const float4*>(X)[i];
    const float4 mu = reinterpret_cast<const float4*>(mean)[c], rs = reinterpret_cast<const float4*>(rstd)[c];
    const float4 g = reinterpret_cast<const float4*>(gamma)[c], b = reinterpret_cast<const float4*>(beta)[c];
    float4 y;
    y.x = fmaxf((x.x - mu.x) * rs.x * g.x + b.x, 0.f);
    y.y = fmaxf((x.y - mu.y) * rs.y * g.y + b.y, 0.f);
    y.z = fmaxf((x.z - mu.z) * rs.z * g.z + b.z, 0.f);
    y.w = fmaxf((x.w - mu.w) * rs.w * g.w + b.w, 0.f);
    if (round_out) {  // the consumer is a plain TF32 GEMM (0: a 3xTF32 GEMM splits the fp32 value itself)
      y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w);
    }
    reinterpret_cast<float4*>(Y)[i] = y;
  }
}

// backward pass 1: s1[c] = sum dyr, s2[c] = sum dyr * xhat with dyr = dy * (bn_out > 0)
__global__ void bn_relu_bwd_stats_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, int rows,
                                         int cols, float* __restrict__ s1, float* __restrict__ s2) {
  __shared__ float red[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a1 = 0.f, a2 = 0.f;
  if (c < cols) {
    const float mu = mean[c], rs = rstd[c], g = gamma[c], b = beta[c];
    for (int r = ry; r < rows; r += 8) {
      const float xh = (X[static_cast<long long>(r) * cols + c] - mu) * rs;
      const float dy = (xh * g + b > 0.f) ? dY[static_cast<long long>(r) * cols + c] : 0.f;
      a1 += dy;
      a2 += dy * xh;
    }
  }
  red[0][ry][cx] = a1;
  red[1][ry][cx] = a2;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1 += red[0][j][cx]; t2 += red[1][j][cx]; }
    s1[c] = t1;
    s2[c] = t2;
  }
}

// backward pass 2: dx = gamma * rstd * (dyr - s1/n - xhat * s2/n), rounded to tf32 (feeds dgrad/wgrad)
__global__ void bn_relu_bwd_apply_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         const float* __restrict__ s1, const float* __restrict__ s2, float inv_n,
                                         float* __restrict__ dX, long long total, int cols, int round_out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cols);
    const float xh = (X[i] - mean[c]) * rstd[c];
    const float dy = (xh * gamma[c] + beta[c] > 0.f) ? dY[i] : 0.f;
    const float dx = gamma[c] * rstd[c] * (dy - s1[c] * inv_n - xh * s2[c] * inv_n);
    dX[i] = round_out ? round_tf32(dx) : dx;
  }
}

// un-fused GELU pair (used when the GEMM epilogue would be the bottleneck): same single-exp formulation as the
// fused epilogues in gemm_tcgen05.cu
__device__ __forceinline__ void gelu_cdf_pdf(float u, float& cdf, float& pdf) {
  const float x = u * 0.70710678118654752f;
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));  // MUFU.RCP, branch-free (keeps the 8 chains interleaved)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = __expf(-ax * ax);
  cdf = 0.5f * (1.0f + copysignf(fmaf(-poly, e, 1.0f), x));
  pdf = 0.3989422804014327f * e;
}
// g = tf32(gelu(u))
__global__ void gelu_fwd_kernel(const float* __restrict__ u, float* __restrict__ g, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(u) + i);
    float c, p;
    float4 o;
    gelu_cdf_pdf(v.x, c, p); o.x = round_tf32(v.x * c);
    gelu_cdf_pdf(v.y, c, p); o.y = round_tf32(v.y * c);
    gelu_cdf_pdf(v.z, c, p); o.z = round_tf32(v.z * c);
    gelu_cdf_pdf(v.w, c, p); o.w = round_tf32(v.w * c);
    reinterpret_cast<float4*>(g)[i] = o;
  }
}
// d = tf32(d * gelu'(u)) in place; optionally colsum_out += column sums of the result (fc1 bias gradient).
// CTA = 128 columns (32 lanes x float4) x a row chunk; warps stride over the rows, 512 B contiguous per warp-row.
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(float* __restrict__ d, const float* __restrict__ u, int rows, int cols, float* __restrict__ colsum_out) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c4 = blockIdx.x * 32 + lane;  // float4 column index
  const int cols4 = cols >> 2;
  const int rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, rows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < cols4) {
    // four rows per trip: eight independent 16-byte loads in flight per thread (the row stride defeats the
    // hardware's sequential prefetch, so memory-level parallelism has to come from the unroll)
    for (int r = r0 + warp; r < r1; r += 32) {
      float4 v[4], o[4];
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ok[k] = r + 8 * k < r1;
        const size_t i = static_cast<size_t>(ok[k] ? r + 8 * k : r) * cols4 + c4;
        v[k] = __ldcs(reinterpret_cast<const float4*>(u) + i);
        o[k] = reinterpret_cast<float4*>(d)[i];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        float c, p;
        gelu_cdf_pdf(v[k].x, c, p); o[k].x = round_tf32(o[k].x * fmaf(v[k].x, p, c));
        gelu_cdf_pdf(v[k].y, c, p); o[k].y = round_tf32(o[k].y * fmaf(v[k].y, p, c));
        gelu_cdf_pdf(v[k].z, c, p); o[k].z = round_tf32(o[k].z * fmaf(v[k].z, p, c));
        gelu_cdf_pdf(v[k].w, c, p); o[k].w = round_tf32(o[k].w * fmaf(v[k].w, p, c));
        reinterpret_cast<float4*>(d)[static_cast<size_t>(r + 8 * k) * cols4 + c4] = o[k];
        acc.x += o[k].x; acc.y += o[k].y; acc.z += o[k].z; acc.w += o[k].w;
      }
    }
  }
  if (colsum_out == nullptr) return;
  *reinterpret_cast<float4*>(&red[warp][lane * 4]) = acc;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col < cols) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
      atomicAdd(colsum_out + col, s);
    }
  }
}

// out[r, :] = x[idx[r], :]   (ATST-Frame: masked valid frames -> compact rows, atstframe/audio_transformer.py:207)
__global__ void gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ out,
                                   long long total4, int d4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / d4;
    const int c = static_cast<int>(i - r * d4);
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(x)[static_cast<long long>(idx[r]) * d4 + c];
  }
}
// dst[idx[r], :] = src[r, :]   (dst zero-filled by the caller; indices are unique)
__global__ void scatter_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst,
                                    long long total4, int d4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / d4;
    const int c = static_cast<int>(i - r * d4);
    reinterpret_cast<float4*>(dst)[static_cast<long long>(idx[r]) * d4 + c] = reinterpret_cast<const float4*>(src)[i];
  }
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(src)[i];
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
}

// 3xTF32 operand split (validation build's GEMMs): dst block b of element (r, c) = part[b] of src[r, c], where
// x = hi + lo, hi = tf32(x), lo = tf32(x - hi); parts = {hi, lo, hi} (pattern 0) or {hi, hi, lo} (pattern 1)
__global__ void split_tf32_kernel(const float* __restrict__ src, long long ld, int rows, int cols,
                                  float* __restrict__ dst, int pattern, int along_rows) {
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const float x = src[r * ld + c];
    const float hi = cvt_rna_tf32(x);
    const float lo = cvt_rna_tf32(x - hi);
    const float p1 = pattern == 0 ? lo : hi, p2 = pattern == 0 ? hi : lo;
    if (along_rows) {
      dst[i] = hi;
      dst[total + i] = p1;
      dst[2 * total + i] = p2;
    } else {
      float* d = dst + r * 3LL * cols + c;
      d[0] = hi;
      d[cols] = p1;
      d[2 * cols] = p2;
    }
  }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] += a * x[i];
}

static inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int patchify(const float* mel, long long clip_stride, int S, int T, float* patches, cudaStream_t st) {
  const int P = T / 4;
  ATST_REQUIRE(S > 0 && P > 0, "patchify: bad shape S=%d T=%d", S, T);
  const long long total = static_cast<long long>(S) * P * 64;
  patchify_kernel<<<grid_for(total), 256, 0, st>>>(mel, clip_stride, T, P, patches, total);
  return atst_check_launch("patchify_kernel");
}

int tokens_forward(const float* pe, const float* cls, const float* pos, const float* mask_embed,
                   const unsigned char* mask, float* x, int S, int P, int D, int use_cls, cudaStream_t st) {
  ATST_REQUIRE(D % 4 == 0, "tokens_forward: D %% 4 != 0");
  const long long total = static_cast<long long>(S) * (P + (use_cls ? 1 : 0)) * (D / 4);
  tokens_fwd_kernel<<<grid_for(total), 256, 0, st>>>(pe, cls, pos, mask_embed, mask, x, S, P, D, use_cls);
  return atst_check_launch("tokens_fwd_kernel");
}

int tokens_backward(const float* dx, const unsigned char* mask, float* dpe, float* dpos, float* dcls,
                    float* dmask_embed, int S, int P, int D, int use_cls, cudaStream_t st) {
  const long long total = static_cast<long long>(P + (use_cls ? 1 : 0)) * D;
  tokens_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(dx, mask, dpe, dpos, dcls, dmask_embed, S, P,
                                                                          D, use_cls);
  return atst_check_launch("tokens_bwd_kernel");
}

int colsum_accumulate(const float* X, long long ld, int rows, int cols, float* out, cudaStream_t st) {
  int gy = (rows + 2047) / 2048;
  if (gy > 64) gy = 64;
  if (gy < 1) gy = 1;
  if (cols % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    const int gx = (cols / 4 + 31) / 32;
    int gy4 = (148 * 8 + gx - 1) / gx;
    if (gy4 > (rows + 63) / 64) gy4 = (rows + 63) / 64;
    if (gy4 < 1) gy4 = 1;
    colsum4_kernel<<<dim3(gx, gy4), 256, 0, st>>>(X, ld, rows, cols, out);
    return atst_check_launch("colsum4_kernel");
  }
  colsum_kernel<<<dim3((cols + 31) / 32, gy), 256, 0, st>>>(X, ld, rows, cols, out);
  return atst_check_launch("colsum_kernel");
}

int bn_stats(const float* X, int rows, int cols, float* mean, float* m2, cudaStream_t st) {
  bn_stats_kernel<<<(cols + 31) / 32, 256, 0, st>>>(X, rows, cols, mean, m2);
  return atst_check_launch("bn_stats_kernel");
}
int bn_finalize(const float* mean, const float* m2, float count, float eps, float momentum, float* rstd,
                float* running_mean, float* running_var, int cols, cudaStream_t st) {
  bn_finalize_kernel<<<(cols + 255) / 256, 256, 0, st>>>(mean, m2, count, eps, momentum, rstd, running_mean,
                                                         running_var, cols);
  return atst_check_launch("bn_finalize_kernel");
}
int bn_relu_forward(const float* X, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    float* Y, int rows, int cols, int round_out, cudaStream_t st) {
  ATST_REQUIRE(cols % 4 == 0, "bn_relu_forward: cols %% 4 != 0");
  const long long total4 = static_cast<long long>(rows) * cols / 4;
  bn_relu_fwd_kernel<<<grid_for(total4), 256, 0, st>>>(X, mean, rstd, gamma, beta, Y, total4, cols / 4, round_out);
  return atst_check_launch("bn_relu_fwd_kernel");
}
int bn_relu_backward_stats(const float* dY, const float* X, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, int rows, int cols, float* s1, float* s2, cudaStream_t st) {
  bn_relu_bwd_stats_kernel<<<(cols + 31) / 32, 256, 0, st>>>(dY, X, mean, rstd, gamma, beta, rows, cols, s1, s2);
  return atst_check_launch("bn_relu_bwd_stats_kernel");
}
int bn_relu_backward_apply(const float* dY, const float* X, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, const float* s1, const float* s2, float count, float* dX, int rows,
                           int cols, int round_out, cudaStream_t st) {
  const long long total = static_cast<long long>(rows) * cols;
  bn_relu_bwd_apply_kernel<<<grid_for(total), 256, 0, st>>>(dY, X, mean, rstd, gamma, beta, s1, s2, 1.0f / count, dX,
                                                           total, cols, round_out);
  return atst_check_launch("bn_relu_bwd_apply_kernel");
}
int gather_rows(const float* x, const int* idx, float* out, int rows, int D, cudaStream_t st) {
  ATST_REQUIRE(D % 4 == 0 && rows >= 0, "gather_rows: D %% 4 != 0");
  if (rows == 0) return ATST_OK;
  const long long total4 = static_cast<long long>(rows) * (D / 4);
  gather_rows_kernel<<<grid_for(total4), 256, 0, st>>>(x, idx, out, total4, D / 4);
  return atst_check_launch("gather_rows_kernel");
}
int scatter_rows(const float* src, const int* idx, float* dst, int rows, int D, cudaStream_t st) {
  ATST_REQUIRE(D % 4 == 0 && rows >= 0, "scatter_rows: D %% 4 != 0");
  if (rows == 0) return ATST_OK;
  const long long total4 = static_cast<long long>(rows) * (D / 4);
  scatter_rows_kernel<<<grid_for(total4), 256, 0, st>>>(src, idx, dst, total4, D / 4);
  return atst_check_launch("scatter_rows_kernel");
}
int gelu_forward(const float* u, float* g, long long n, cudaStream_t st) {
  ATST_REQUIRE(n % 4 == 0, "gelu_forward: n %% 4 != 0");
  gelu_fwd_kernel<<<grid_for(n / 4, 256, 148 * 32), 256, 0, st>>>(u, g, n / 4);
  return atst_check_launch("gelu_fwd_kernel");
}
int gelu_backward(float* d, const float* u, int rows, int cols, float* colsum_out, cudaStream_t st) {
  ATST_REQUIRE(cols % 4 == 0 && rows > 0, "gelu_backward: cols %% 4 != 0");
  const int gx = (cols / 4 + 31) / 32;
  int gy = (148 * 8 + gx - 1) / gx;
  if (gy > (rows + 63) / 64) gy = (rows + 63) / 64;
  if (gy < 1) gy = 1;
  gelu_bwd_kernel<<<dim3(gx, gy), 256, 0, st>>>(d, u, rows, cols, colsum_out);
  return atst_check_launch("gelu_bwd_kernel");
}
int round_tf32_copy(const float* src, float* dst, long long n, cudaStream_t st) {
  ATST_REQUIRE(n % 4 == 0, "round_tf32_copy: n %% 4 != 0");
  round_tf32_kernel<<<grid_for(n / 4), 256, 0, st>>>(src, dst, n / 4);
  return atst_check_launch("round_tf32_kernel");
}
int split_tf32(const float* src, long long ld, int rows, int cols, float* dst, int pattern, int along_rows,
               cudaStream_t st) {
  ATST_REQUIRE(rows > 0 && cols > 0 && ld >= cols && (pattern == 0 || pattern == 1), "split_tf32: bad arguments");
  split_tf32_kernel<<<grid_for(static_cast<long long>(rows) * cols), 256, 0, st>>>(src, ld, rows, cols, dst, pattern,
                                                                                     along_rows);
  return atst_check_launch("split_tf32_kernel");
}
int axpy(float* y, const float* x, float a, long long n, cudaStream_t st) {
  axpy_kernel<<<grid_for(n), 256, 0, st>>>(y, x, a, n);
  return atst_check_launch("axpy_kernel");
}

}  // namespace atst
