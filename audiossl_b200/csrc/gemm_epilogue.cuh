// Epilogue shared by the 1-CTA and the CTA-pair GEMM kernels: one accumulator tile (128 TMEM lanes x BLOCK_N
// columns of this CTA) -> registers -> fused math -> 256-bit global accesses.  See gemm_tcgen05.cu for the design
// notes (no smem staging: the smem port belongs to TMA + UMMA).
#pragma once
#include "common.cuh"
#include "gemm.h"

namespace atst {

// erf by Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below the TF32 operand rounding of the next GEMM):
// one MUFU.RCP + one MUFU.EX2 + 7 FMA instead of libdevice erff's ~25 instructions.  The exp(-u^2/2) factor is
// shared between the cdf and the pdf, so gelu'(u) costs no second exponential.
struct GeluParts { float cdf, pdf; };
__device__ __forceinline__ GeluParts gelu_parts(float u) {
  const float x = u * 0.70710678118654752f;
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));  // MUFU.RCP, branch-free (keeps the 8 chains interleaved)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = __expf(-ax * ax);            // exp(-u^2 / 2)
  const float erf_abs = fmaf(-poly, e, 1.0f);  // erf(|x|)
  GeluParts g;
  g.cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  g.pdf = 0.3989422804014327f * e;
  return g;
}
__device__ __forceinline__ float gelu_exact(float u) { return u * gelu_parts(u).cdf; }
__device__ __forceinline__ float gelu_grad(float u) {
  const GeluParts g = gelu_parts(u);
  return fmaf(u, g.pdf, g.cdf);
}


// Thread `lane` of epilogue warp `ew` owns accumulator row 32*ew + lane.  `release()` is called once, right after
// the last TMEM load of the tile, to hand the accumulator stage back to the MMA issuer.
template <int BLOCK_N, class Release>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, int m0, int n0, bool empty_split, uint32_t taddr,
                                              float* sbias, int ew, int lane, uint64_t* tfull_bar, uint32_t acc_phase,
                                              Release release) {
  const float* side_ptr = (p.epi == EPI_RESID) ? p.resid : ((p.epi == EPI_DGELU) ? p.aux : nullptr);
  const int side_ld = (p.epi == EPI_RESID) ? p.ldr : p.ldaux;
  const int gm = m0 + ew * 32 + lane;
  const bool row_ok = gm < p.M && !empty_split;
  const float rs = (p.rowscale != nullptr && row_ok) ? p.rowscale[gm / p.rows_per_seq] : 1.0f;
  const float* side_row = side_ptr ? side_ptr + static_cast<size_t>(gm) * side_ld : nullptr;
  float* c_row = p.C + static_cast<size_t>(gm) * p.ldc;
  float* aux_row = (p.epi == EPI_GELU) ? p.aux + static_cast<size_t>(gm) * p.ldaux : nullptr;
  // while this tile's main loop is still running: pull the side-input rows into L2 and stage the bias slice
  if (side_row != nullptr && row_ok) {
#pragma unroll
    for (int c = 0; c < BLOCK_N / 32; ++c)
      if (n0 + c * 32 < p.N) prefetch_l2(side_row + n0 + c * 32);
  }
  if (p.bias != nullptr) {
    const int t128 = ew * 32 + lane;
    for (int j = t128; j < BLOCK_N; j += 128) sbias[j] = (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
  }
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  taddr += static_cast<uint32_t>(ew * 32) << 16;

  auto load_side = [&](int c, float (&sd)[32]) {
    const int gn = n0 + c * 32;
    if (side_row == nullptr || !row_ok || gn >= p.N) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (gn + 8 * j < p.N) ld_global_v8(side_row + gn + 8 * j, &sd[8 * j]);
    }
  };
  auto process = [&](int c, const float (&sd)[32]) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    tmem_ld_wait();
    if (c == BLOCK_N / 32 - 1) {
      tc_fence_before();
      __syncwarp();
      release();
    }
    const int gn = n0 + c * 32;
    if (!row_ok || gn >= p.N) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 8 columns at a time
      if (gn + 8 * j >= p.N) break;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]);
      if (p.bias != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(sbias + c * 32 + 8 * j);  // smem broadcast
        const float4 b1 = *reinterpret_cast<const float4*>(sbias + c * 32 + 8 * j + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      switch (p.epi) {
        case EPI_GELU:
          st_global_v8(aux_row + gn + 8 * j, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = gelu_exact(v[e]);
          break;
        case EPI_DGELU:
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= gelu_grad(sd[8 * j + e]);
          break;
        case EPI_RESID:
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaf(rs, v[e], sd[8 * j + e]);
          break;
        case EPI_SCALE:
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= rs;
          break;
        case EPI_RELU:
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
          break;
        default:
          break;
      }
      if (p.round_out) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = round_tf32(v[e]);
      }
      float* cp = c_row + gn + 8 * j;
      if (p.epi == EPI_ATOMIC) {
        red_add_v4(cp, v[0], v[1], v[2], v[3]);
        red_add_v4(cp + 4, v[4], v[5], v[6], v[7]);
      } else {
        st_global_v8(cp, v);
      }
    }
  };
  float side_a[32], side_b[32];
  load_side(0, side_a);
#pragma unroll 1
  for (int c = 0; c < BLOCK_N / 32; c += 2) {
    load_side(c + 1, side_b);
    process(c, side_a);
    if (c + 2 < BLOCK_N / 32) load_side(c + 2, side_a);
    process(c + 1, side_b);
  }
}

}  // namespace atst
