// Epilogue shared by the 1-CTA and the CTA-pair GEMM kernels: one accumulator tile (128 TMEM lanes x BLOCK_N
// columns of this CTA) -> registers -> fused math -> 256-bit global accesses.  See gemm_tcgen05.cu for the design
// notes (no smem staging: the smem port belongs to TMA + UMMA).
#pragma once
#include "common.cuh"
#include "gemm.h"

#ifndef ATST_EPI_X16
#define ATST_EPI_X16 0   // 1: the two-stream epilogues fetch 16 accumulator columns per tcgen05.ld (single buffer)
#endif

namespace atst {

// erf by Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below the TF32 operand rounding of the next GEMM):
// one MUFU.RCP + one MUFU.EX2 + 7 FMA instead of libdevice erff's ~25 instructions.  The exp(-u^2/2) factor is
// shared between the cdf and the pdf, so gelu'(u) costs no second exponential.
struct GeluParts { float cdf, pdf; };
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ GeluParts gelu_parts(float u) {
  const float x = u * 0.70710678118654752f;
  const float ax = fabsf(x);
  // raw MUFU.RCP / MUFU.EX2 (flush-to-zero forms): the argument of the reciprocal is >= 1 and a flushed exp(-x^2)
  // only occurs where erf has long saturated, so the denormal fix-ups of __fdividef / __expf buy nothing here
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);  // exp(-u^2 / 2)
  const float erf_abs = fmaf(-poly, e, 1.0f);                  // erf(|x|)
  GeluParts g;
  g.cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  g.pdf = 0.3989422804014327f * e;
  return g;
}
__device__ __forceinline__ float gelu_exact(float u) { return u * gelu_parts(u).cdf; }
// two fp32 -> one packed f16x2 word (round to nearest even), `lo` in the low half = the lower address; and back
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void unpack_half2(uint32_t w, float& lo, float& hi) {
  asm("{\n\t"
      ".reg .f16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "cvt.f32.f16 %0, l;\n\t"
      "cvt.f32.f16 %1, h;\n\t"
      "}"
      : "=f"(lo), "=f"(hi)
      : "r"(w));
}
__device__ __forceinline__ float gelu_grad(float u) {
  const GeluParts g = gelu_parts(u);
  return fmaf(u, g.pdf, g.cdf);
}


// Thread `lane` of an epilogue warp of TMEM lane quadrant `ew` owns accumulator row 32*ew + lane and walks the
// 8-column groups [g0, g1) of the tile (g1 - g0 a multiple of 4; four warps per quadrant: a 64-column quarter each).
// The loop is deliberately NOT unrolled over the tile: one compact body (TMEM load of the next group and its side
// input in flight while the current group is computed and stored) instead of a copy of every epilogue variant per
// column chunk - the unrolled form thrashed the instruction cache (ncu: 23 % of the stall samples were "no
// instruction" with the GELU epilogue).  `release()` is called once per warp, right after its last TMEM load of the
// tile, to hand the accumulator stage back to the MMA issuer.  EPI_THREADS = epilogue threads of the CTA.
//
// What bounds this epilogue (per-tile clock64 timeline, tools/bringup.py gemm_trace): 128 KB of global
// traffic per tile take ~11 000 cycles whatever the access shape - the SM's link to L2 is already filled by the main
// loop's operand fetch (768 KB per tile) - so one output stream hides under the 13 600-cycle main loop and a second
// one (residual / saved pre-activation in, pre-activation out) does not.  Measured and rejected: 4 lanes per row after
// a quad transpose (same time), a shared-memory staging tile with TMA loads / stores (slower: the shared-memory port
// belongs to TMA + UMMA), 32- or 16-column accumulator loads in flight with an early release of the stage (plain
// shapes +5 %, fused shapes -20 %: the unthrottled main loop takes the link from the epilogue that is the bottleneck).
//
// ECLS: the epilogue class this instantiation is compiled for.  The runtime `switch (p.epi)` form of every epilogue in
// one kernel (ECLS_GENERIC, kept for the rarely used variants) costs all of them its registers (96 with spills) and its
// code: the same fc1 + GELU epilogue without a side stream runs at 0.985 ms per config-2 launch compiled alone
// against 1.098 ms inside the generic kernel (tools/ab_gelu_half.py; plain fc1: 0.962 ms).
enum EpiClass : int {
  ECLS_GENERIC = 0,   // any p.epi at run time (fp32 pre-activation side streams, column sums)
  ECLS_PLAIN = 1,     // one output stream: EPI_STORE / EPI_SCALE / EPI_RELU / EPI_ATOMIC (+ bias, rounding)
  ECLS_RESID = 2,     // EPI_RESID
  ECLS_GELU = 3,      // EPI_GELU without a stored pre-activation (teacher / inference fc1)
  ECLS_HALF_FWD = 4,  // EPI_GELU_H
  ECLS_HALF_BWD = 5,  // EPI_DGELU_H
};

template <int BLOCK_N, int EPI_THREADS, int ECLS = ECLS_GENERIC, class Release>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, int m0, int n0, bool empty_split, uint32_t taddr,
                                              float* sbias, int ew, int lane, int epi_tid, int g0, int g1,
                                              uint64_t* tfull_bar, uint32_t acc_phase, Release release,
                                              long long* trace = nullptr) {
  constexpr bool kGeneric = ECLS == ECLS_GENERIC;
  constexpr bool kHalf = ECLS == ECLS_GELU || ECLS == ECLS_HALF_FWD || ECLS == ECLS_HALF_BWD;
  const float* side_ptr = nullptr;
  int side_ld = 0;
  if constexpr (ECLS == ECLS_RESID) {
    side_ptr = p.resid;
    side_ld = p.ldr;
  } else if constexpr (kGeneric) {
    side_ptr = (p.epi == EPI_RESID) ? p.resid : ((p.epi == EPI_DGELU) ? p.aux : nullptr);
    side_ld = (p.epi == EPI_RESID) ? p.ldr : p.ldaux;
  }
  const int gm = m0 + ew * 32 + lane;
  const bool row_ok = gm < p.M && !empty_split;
  const float rs = (p.rowscale != nullptr && row_ok) ? p.rowscale[gm / p.rows_per_seq] : 1.0f;
  const float* side_row = (side_ptr && row_ok) ? side_ptr + static_cast<size_t>(gm) * side_ld + n0 : nullptr;
  float* c_row = p.C + static_cast<size_t>(gm) * p.ldc + n0;
  float* aux_row = (kGeneric && p.epi == EPI_GELU && p.aux != nullptr) ? p.aux + static_cast<size_t>(gm) * p.ldaux + n0 : nullptr;
  const int ncols = min(BLOCK_N, p.N - n0);  // valid columns of this tile (multiple of 8)
  // while this tile's main loop is still running: pull the side-input rows into L2 and stage the bias slice
  if (side_row != nullptr) {
    for (int g = g0; g < g1; g += 4)
      if (g * 8 < ncols) prefetch_l2(side_row + g * 8);
  }
  if (p.bias != nullptr) {
    for (int j = epi_tid; j < BLOCK_N; j += EPI_THREADS) sbias[j] = (j < ncols) ? __ldg(p.bias + n0 + j) : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");  // epilogue warps only
  }
  if (trace) trace[1] = clock64();  // bias staged, waiting for the accumulator
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  if (trace) trace[2] = clock64();  // accumulator complete
  taddr += static_cast<uint32_t>(ew * 32) << 16;

  auto fetch_acc = [&](int g, uint32_t (&r)[8]) { tmem_ld_32x8(taddr + g * 8, r); };
  auto fetch_side = [&](int g, float (&sd)[8]) {
    if (side_row != nullptr && g < g1 && g * 8 < ncols) ld_global_v8(side_row + g * 8, sd);
  };
  const bool do_colsum = kGeneric && p.colsum != nullptr;  // warp-uniform
  auto finish = [&](int g, const uint32_t (&r)[8], const float (&sd)[8]) {
    if (g * 8 >= ncols || (!row_ok && !do_colsum)) return;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]);
    if (p.bias != nullptr) {
      const float4 b0 = *reinterpret_cast<const float4*>(sbias + g * 8);  // smem broadcast
      const float4 b1 = *reinterpret_cast<const float4*>(sbias + g * 8 + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if constexpr (ECLS == ECLS_RESID) {  // C <- resid + rowscale[seq] * (acc + bias)
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(rs, v[e], sd[e]);
    } else if constexpr (ECLS == ECLS_PLAIN) {
      if (p.epi == EPI_SCALE) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= rs;
      } else if (p.epi == EPI_RELU) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
      }
    } else {
      switch (p.epi) {
      case EPI_GELU:  // aux (nullable: the teacher keeps no pre-activation) <- pre-activation, C <- gelu
        if (aux_row != nullptr) st_global_v8(aux_row + g * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gelu_exact(v[e]);
        break;
      case EPI_DGELU:  // C <- acc * gelu'(aux)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= gelu_grad(sd[e]);
        break;
      case EPI_RESID:  // C <- resid + rowscale[seq] * (acc + bias)
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaf(rs, v[e], sd[e]);
        break;
      case EPI_SCALE:  // C <- rowscale[seq] * acc   (dgrad through droppath)
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
    }
    if (p.round_out) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = round_tf32(v[e]);
    }
    float* cp = c_row + g * 8;
    if (row_ok) {
      if (ECLS != ECLS_RESID && p.epi == EPI_ATOMIC) {
        red_add_v4(cp, v[0], v[1], v[2], v[3]);
        red_add_v4(cp + 4, v[4], v[5], v[6], v[7]);
      } else {
        st_global_v8(cp, v);
      }
    }
    if (do_colsum) {
      // column sums of the stored values over this warp's 32 rows (the consumer Linear's bias gradient): a
      // butterfly per column, then one red.global.add per column from eight different lanes
      float mine = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float s = warp_sum(row_ok ? v[e] : 0.f);
        if (lane == e) mine = s;
      }
      if (lane < 8) atomicAdd(p.colsum + n0 + g * 8 + lane, mine);
    }
  };
  if constexpr (kHalf) {
    // The GELU classes.  ECLS_HALF_FWD / ECLS_HALF_BWD: the student's MLP with an fp16 side stream - the forward epilogue
    // stores gelu'(pre-activation), from the same cdf / pdf as the GELU itself, and the dgrad epilogue multiplies by it.
    // ECLS_GELU: the same forward loop without a side stream (teacher / inference).  Same pipeline as the two-stream
    // loop at the end (8-column accumulator loads one ahead).  The side stream moves as ONE 32-byte sector per PAIR of
    // groups (16 halfs): the even group of a pair loads / the odd group stores the sector, `hp` carries the other
    // group's four packed words in between.  g0 even, N % 16 == 0 (host-checked).
    constexpr bool fwd = ECLS != ECLS_HALF_BWD;
    constexpr bool deriv = ECLS == ECLS_HALF_FWD;  // ECLS_GELU: the same loop without the side stream
    uint16_t* haux = (ECLS != ECLS_GELU && p.aux != nullptr && row_ok)
                         ? reinterpret_cast<uint16_t*>(p.aux) + static_cast<size_t>(gm) * p.ldaux + n0 : nullptr;
    if (!fwd && haux != nullptr) {
      for (int g = g0; g < g1; g += 8)
        if (g * 8 < ncols) prefetch_l2(haux + g * 8);
    }
    uint32_t ra[8], rb[8], hp[4] = {0u, 0u, 0u, 0u};
    float sh[2][8];  // the 16 halfs of a pair of groups, fetched two pairs ahead
    auto fetch_pair = [&](int g, float (&sd)[8]) {  // g even: the halfs of groups g and g + 1
      if (!fwd && haux != nullptr && g < g1 && g * 8 < ncols) ld_global_v8(reinterpret_cast<const float*>(haux + g * 8), sd);
    };
    auto finish_h = [&](int g, const uint32_t (&r)[8], const float (&sd)[8], const bool even) {
      if (g * 8 >= ncols || !row_ok) return;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]);
      if (fwd) {
        if (p.bias != nullptr) {
          const float4 b0 = *reinterpret_cast<const float4*>(sbias + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sbias + g * 8 + 4);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const GeluParts a = gelu_parts(v[2 * i]), b = gelu_parts(v[2 * i + 1]);
          if constexpr (deriv) w[i] = pack_half2(fmaf(v[2 * i], a.pdf, a.cdf), fmaf(v[2 * i + 1], b.pdf, b.cdf));
          v[2 * i] *= a.cdf;
          v[2 * i + 1] *= b.cdf;
        }
        if constexpr (!deriv) {
        } else if (even) {
#pragma unroll
          for (int i = 0; i < 4; ++i) hp[i] = w[i];
        } else if (haux != nullptr) {
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(haux + (g - 1) * 8), "r"(hp[0]),
                       "r"(hp[1]), "r"(hp[2]), "r"(hp[3]), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      } else {
        if (even) {
#pragma unroll
          for (int i = 0; i < 4; ++i) hp[i] = __float_as_uint(sd[4 + i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float lo, hi;
          unpack_half2(even ? __float_as_uint(sd[i]) : hp[i], lo, hi);
          v[2 * i] *= lo;
          v[2 * i + 1] *= hi;
        }
      }
      if (p.round_out) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = round_tf32(v[e]);
      }
      st_global_v8(c_row + g * 8, v);
    };
#if ATST_EPI_X16
    {
      uint32_t r16[16];
      fetch_pair(g0, sh[0]);
      fetch_pair(g0 + 2, sh[1]);
      tmem_ld_32x16(taddr + g0 * 8, r16);
#pragma unroll 1
      for (int g = g0; g < g1; g += 4) {
        tmem_ld_wait();
        finish_h(g, *reinterpret_cast<const uint32_t(*)[8]>(&r16[0]), sh[0], true);
        finish_h(g + 1, *reinterpret_cast<const uint32_t(*)[8]>(&r16[8]), sh[0], false);
        tmem_ld_32x16(taddr + (g + 2) * 8, r16);
        fetch_pair(g + 4, sh[0]);
        tmem_ld_wait();
        if (g + 4 >= g1) {  // this warp's last TMEM load of the tile has completed
          tc_fence_before();
          __syncwarp();
          release();
        }
        finish_h(g + 2, *reinterpret_cast<const uint32_t(*)[8]>(&r16[0]), sh[1], true);
        finish_h(g + 3, *reinterpret_cast<const uint32_t(*)[8]>(&r16[8]), sh[1], false);
        if (g + 4 < g1) tmem_ld_32x16(taddr + (g + 4) * 8, r16);
        fetch_pair(g + 6, sh[1]);
      }
      return;
    }
#endif
    fetch_pair(g0, sh[0]);
    fetch_pair(g0 + 2, sh[1]);
    tmem_ld_32x8(taddr + g0 * 8, ra);
#pragma unroll 1
    for (int g = g0; g < g1; g += 4) {
      tmem_ld_wait();
      tmem_ld_32x8(taddr + (g + 1) * 8, rb);
      finish_h(g, ra, sh[0], true);
      fetch_pair(g + 4, sh[0]);
      tmem_ld_wait();
      tmem_ld_32x8(taddr + (g + 2) * 8, ra);
      finish_h(g + 1, rb, sh[0], false);
      tmem_ld_wait();
      tmem_ld_32x8(taddr + (g + 3) * 8, rb);
      finish_h(g + 2, ra, sh[1], true);
      fetch_pair(g + 6, sh[1]);
      tmem_ld_wait();
      if (g + 4 < g1) {
        tmem_ld_32x8(taddr + (g + 4) * 8, ra);
      } else {  // this warp's last TMEM load of the tile has completed
        tc_fence_before();
        __syncwarp();
        release();
      }
      finish_h(g + 3, rb, sh[1], false);
    }
    return;
  }
  // warp-uniform by construction (parameters only): the two forms below use .sync.aligned TMEM loads and each calls
  // release() once per warp, so a warp must never split between them - with the per-thread side_row in this test, a
  // warp straddling the last valid row of a residual epilogue did (rows past M have no side pointer), which
  // double-released the accumulator stage and hung the kernel for M % 32 != 0
  bool one_stream;
  if constexpr (ECLS == ECLS_PLAIN) one_stream = true;
  else if constexpr (ECLS == ECLS_RESID) one_stream = false;
  else one_stream = side_ptr == nullptr && p.epi != EPI_GELU && p.epi != EPI_DGELU;
  if (one_stream) {
    // Plain epilogues (one global stream, no GELU math): the tile period is the main loop, and what matters is handing
    // the accumulator stage back early - a tcgen05.ld takes ~1500 cycles under a running main loop, so eight dependent
    // 8-column loads hold the stage for ~12 000 cycles and the MMA issuer waits ~2500 cycles per tile for a free one.
    // Two 16-column loads in flight, the next pair issued while this one is stored, release after the last wait
    // (+5 % on these shapes; the fused epilogues below lose 20 % with it and keep the sequential form).
    uint32_t wa[16], wb[16];
    float none[8];
    tmem_ld_32x16(taddr + g0 * 8, wa);
    tmem_ld_32x16(taddr + g0 * 8 + 16, wb);
    tmem_ld_wait();
    if (g0 + 4 >= g1) {
      tc_fence_before();
      __syncwarp();
      release();
    }
#pragma unroll 1
    for (int g = g0; g < g1; g += 4) {
      const bool more = g + 4 < g1;
      finish(g, *reinterpret_cast<const uint32_t(*)[8]>(&wa[0]), none);
      finish(g + 1, *reinterpret_cast<const uint32_t(*)[8]>(&wa[8]), none);
      if (more) tmem_ld_32x16(taddr + (g + 4) * 8, wa);
      finish(g + 2, *reinterpret_cast<const uint32_t(*)[8]>(&wb[0]), none);
      finish(g + 3, *reinterpret_cast<const uint32_t(*)[8]>(&wb[8]), none);
      if (more) {
        tmem_ld_32x16(taddr + (g + 6) * 8, wb);
        tmem_ld_wait();
        if (g + 8 >= g1) {  // this warp's last TMEM load of the tile has completed
          tc_fence_before();
          __syncwarp();
          release();
        }
      }
    }
    return;
  }
#if ATST_EPI_X16
  // 16 accumulator columns per tcgen05.ld, one buffer: a load is ~1500 cycles under a running main loop whatever its
  // width, so four round trips per 64-column slice instead of eight; the next load is issued as soon as the two groups
  // of this one have been taken out of the registers.  Side inputs four groups ahead.  (g1 - g0) % 4 == 0
  {
    uint32_t r16[16];
    float sd[4][8];
    fetch_side(g0, sd[0]);
    fetch_side(g0 + 1, sd[1]);
    fetch_side(g0 + 2, sd[2]);
    fetch_side(g0 + 3, sd[3]);
    tmem_ld_32x16(taddr + g0 * 8, r16);
#pragma unroll 1
    for (int g = g0; g < g1; g += 4) {
      tmem_ld_wait();
      finish(g, *reinterpret_cast<const uint32_t(*)[8]>(&r16[0]), sd[0]);
      finish(g + 1, *reinterpret_cast<const uint32_t(*)[8]>(&r16[8]), sd[1]);
      tmem_ld_32x16(taddr + (g + 2) * 8, r16);
      fetch_side(g + 4, sd[0]);
      fetch_side(g + 5, sd[1]);
      tmem_ld_wait();
      if (g + 4 >= g1) {  // this warp's last TMEM load of the tile has completed
        tc_fence_before();
        __syncwarp();
        release();
      }
      finish(g + 2, *reinterpret_cast<const uint32_t(*)[8]>(&r16[0]), sd[2]);
      finish(g + 3, *reinterpret_cast<const uint32_t(*)[8]>(&r16[8]), sd[3]);
      if (g + 4 < g1) tmem_ld_32x16(taddr + (g + 4) * 8, r16);
      fetch_side(g + 6, sd[2]);
      fetch_side(g + 7, sd[3]);
    }
    return;
  }
#endif
  // accumulator groups are fetched one ahead (tcgen05.wait::ld waits for every outstanding load, so deeper does not
  // help), side inputs three ahead (their L2 / HBM latency is several groups long); (g1 - g0) % 4 == 0
  uint32_t ra[8], rb[8];
  float sd[4][8];
  fetch_side(g0, sd[0]);
  fetch_side(g0 + 1, sd[1]);
  fetch_side(g0 + 2, sd[2]);
  fetch_acc(g0, ra);
#pragma unroll 1
  for (int g = g0; g < g1; g += 4) {
    tmem_ld_wait();
    fetch_acc(g + 1, rb);
    fetch_side(g + 3, sd[3]);
    finish(g, ra, sd[0]);
    tmem_ld_wait();
    fetch_acc(g + 2, ra);
    fetch_side(g + 4, sd[0]);
    finish(g + 1, rb, sd[1]);
    tmem_ld_wait();
    fetch_acc(g + 3, rb);
    fetch_side(g + 5, sd[1]);
    finish(g + 2, ra, sd[2]);
    tmem_ld_wait();
    if (g + 4 < g1) {
      fetch_acc(g + 4, ra);
      fetch_side(g + 6, sd[2]);
    } else {  // this warp's last TMEM load of the tile has completed
      tc_fence_before();
      __syncwarp();
      release();
    }
    finish(g + 3, rb, sd[3]);
  }
}

}  // namespace atst
