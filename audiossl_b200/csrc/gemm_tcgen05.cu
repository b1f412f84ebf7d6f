// TF32 tcgen05 GEMM for the AST transformer.
//
//   NT  ("K-major"):   C[M,N] = epi( A[M,K] . B[N,K]^T )          forward (B = weight [out,in]) and
//                                                                 dgrad   (B = transposed weight copy)
//   TN  ("MN-major"):  C[M,N] += A[T,M]^T . B[T,N]  (split over T) wgrad: both operands are token-major
//                                                                 activations, contraction over tokens
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor -> 128B-swizzled smem ring (kStages x (A 16 KB + B 16/32 KB))
//   warp 1      MMA issuer: one lane issues tcgen05.mma.kind::tf32 (M=128, N=BLOCK_N, K=8) into TMEM,
//               tcgen05.commit releases smem slots / publishes the accumulator
//   warp 2      TMEM allocator (2 accumulator stages x BLOCK_N columns)
//   warps 4-7   epilogue: tcgen05.ld 32x32b -> registers -> per-warp smem transpose -> fused epilogue
//               (bias / exact-erf GELU / GELU' / droppath-scaled residual / split-K atomics) ->
//               coalesced float4 global stores.  The accumulator is double buffered so the epilogue of
//               tile i overlaps the main loop of tile i+1.
// fp32 containers everywhere; the tensor core reads the top 19 bits (TF32).  Producers round-to-nearest
// (cvt.rna.tf32) the activations/weights they hand to a GEMM so operand error is unbiased.
#include "common.cuh"
#include "gemm.h"
#include "gemm_epilogue.cuh"
#include "ops.h"

namespace atst {

constexpr int kBlockM = 128;
constexpr int kBlockKBytes = 128;  // one 128B swizzle row
constexpr int kBlockK = 32;        // tf32 elements per k-block
constexpr int kUmmaK = 8;          // tf32 elements per tcgen05.mma
constexpr int kThreads = 256;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockKBytes;
  static constexpr int kBBytes = BLOCK_N * kBlockKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kEpiBytes = 0;  // epilogue goes TMEM -> registers -> global, no smem staging
  static constexpr int kBiasBytes = 2 * 256 * 4;  // per accumulator stage: the tile's bias slice
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + 256 + kBiasBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 512 or 256: power of two
};

template <int BLOCK_N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // 1024B alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                     // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;     // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages; // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* smem_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kb_total = (p.K + kBlockK - 1) / kBlockK;
  const int splits = p.splits > 0 ? p.splits : 1;
  const int kb_per_split = (kb_total + splits - 1) / splits;
  const int total_tiles = m_tiles * n_tiles * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // A second cursor runs kPrefetchDist k-blocks ahead of the load cursor (across tile boundaries) and pulls the
    // streaming operand tiles into L2, so a smem slot refill sees L2-hit latency instead of HBM latency: with
    // 4 x 48 KB slots only ~1500 cycles of latency are covered by the ring itself.
    if (lane == 0) {
      constexpr int kPrefetchDist = 12;
      const int tiles_mn = m_tiles * n_tiles;
      auto tile_span = [&](int tile, int& m0, int& n0, int& kb0, int& kb1) {
        const int split = tile / tiles_mn;
        const int rem = tile - split * tiles_mn;
        m0 = (rem / n_tiles) * kBlockM;
        n0 = (rem % n_tiles) * BLOCK_N;
        kb0 = split * kb_per_split;
        kb1 = min(kb0 + kb_per_split, kb_total);
      };
      int pf_tile = blockIdx.x, pf_m0 = 0, pf_n0 = 0, pf_kb = 0, pf_kb1 = 0;
      if (pf_tile < total_tiles) tile_span(pf_tile, pf_m0, pf_n0, pf_kb, pf_kb1);
      auto prefetch_step = [&]() {
        while (pf_tile < total_tiles && pf_kb >= pf_kb1) {
          pf_tile += gridDim.x;
          if (pf_tile < total_tiles) tile_span(pf_tile, pf_m0, pf_n0, pf_kb, pf_kb1);
        }
        if (pf_tile >= total_tiles) return;
        if (A_MN) tma_prefetch_3d(&tmA, 0, pf_kb * kBlockK, pf_m0 / 32);
        else      tma_prefetch_2d(&tmA, pf_kb * kBlockK, pf_m0);
        if (A_MN && B_MN) tma_prefetch_3d(&tmB, 0, pf_kb * kBlockK, pf_n0 / 32);  // wgrad: both operands stream
        ++pf_kb;
      };
      if (p.l2_prefetch)
        for (int i = 0; i < kPrefetchDist; ++i) prefetch_step();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m0, n0, kb0, kb1;
        tile_span(tile, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (p.l2_prefetch) prefetch_step();
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          void* sa = smem_a + stage * Cfg::kABytes;
          void* sb = smem_b + stage * Cfg::kBBytes;
          // MN-major operands are [k rows, features]: box {32 features, 32 k rows, BLOCK/32 feature chunks}
          if (A_MN) tma_load_3d(sa, &tmA, &full_bar[stage], 0, kb * kBlockK, m0 / 32);
          else      tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBlockK, m0);
          if (B_MN) tma_load_3d(sb, &tmB, &full_bar[stage], 0, kb * kBlockK, n0 / 32);
          else      tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, n0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(2u, kBlockM, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / (m_tiles * n_tiles);
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, kb_total);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // K-major : [rows][128 B], 8-row swizzle atoms 1024 B apart; one MMA consumes 32 B of each row
            // MN-major: [feature chunk][32 k rows][128 B]; one MMA consumes 8 k rows = p.mn_kstep bytes
            const uint64_t da = A_MN ? make_smem_desc(a_addr + k * p.mn_kstep, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                     : make_smem_desc(a_addr + k * 32, 16, 1024, 2);
            const uint64_t db = B_MN ? make_smem_desc(b_addr + k * p.mn_kstep, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                     : make_smem_desc(b_addr + k * 32, 16, 1024, 2);
            umma_tf32(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (kb1 <= kb0 && lane == 0) umma_commit(&tfull_bar[acc]);  // empty split: nothing accumulated
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    // Thread t of epilogue warp ew owns accumulator row 32*ew + t (TMEM lane) and walks its 32-column chunks:
    // tcgen05.ld -> registers -> fused math -> 256-bit global stores, with the chunk's side input (residual /
    // saved pre-activation) fetched one chunk ahead by 256-bit loads.  No shared-memory staging: the smem port
    // is already saturated by the TMA fill + UMMA operand reads (96 + 96 B/cycle against 128 B/cycle), and every
    // lane still reads/writes whole 32-byte sectors.
    const int ew = warp - 4;  // == warp % 4: TMEM lane quadrant
    int acc = 0;
    uint32_t acc_phase = 0;
    const float* side_ptr = (p.epi == EPI_RESID) ? p.resid : ((p.epi == EPI_DGELU) ? p.aux : nullptr);
    const int side_ld = (p.epi == EPI_RESID) ? p.ldr : p.ldaux;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / (m_tiles * n_tiles);
      const int rem = tile - split * (m_tiles * n_tiles);
      const int m0 = (rem / n_tiles) * kBlockM;
      const int n0 = (rem % n_tiles) * BLOCK_N;
      const int kb0 = split * kb_per_split;
      const bool empty_split = min(kb0 + kb_per_split, kb_total) <= kb0;
      const int gm = m0 + ew * 32 + lane;
      const bool row_ok = gm < p.M && !empty_split;
      const float rs = (p.rowscale != nullptr && row_ok) ? p.rowscale[gm / p.rows_per_seq] : 1.0f;
      const float* side_row = side_ptr ? side_ptr + static_cast<size_t>(gm) * side_ld : nullptr;
      float* c_row = p.C + static_cast<size_t>(gm) * p.ldc;
      float* aux_row = (p.epi == EPI_GELU && p.aux != nullptr) ? p.aux + static_cast<size_t>(gm) * p.ldaux : nullptr;
      // while this tile's main loop is still running: pull the side-input rows into L2 and stage the bias slice
      if (side_row != nullptr && row_ok) {
#pragma unroll
        for (int c = 0; c < BLOCK_N / 32; ++c)
          if (n0 + c * 32 < p.N) prefetch_l2(side_row + n0 + c * 32);
      }
      float* sbias = smem_bias + acc * 256;
      if (p.bias != nullptr) {
        const int t128 = ew * 32 + lane;
        for (int j = t128; j < BLOCK_N; j += 128) sbias[j] = (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);

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
        if (p.epi != EPI_DBG_NOLOAD) {
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (c == BLOCK_N / 32 - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp before the global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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
            case EPI_GELU:  // aux (nullable) <- pre-activation, C <- gelu
              if (aux_row != nullptr) st_global_v8(aux_row + gn + 8 * j, v);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = gelu_exact(v[e]);
              break;
            case EPI_DGELU:  // C <- acc * gelu'(aux)
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] *= gelu_grad(sd[8 * j + e]);
              break;
            case EPI_RESID:  // C <- resid + rowscale[seq] * (acc + bias)
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaf(rs, v[e], sd[8 * j + e]);
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
          if (p.round_out) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = round_tf32(v[e]);
          }
          float* cp = c_row + gn + 8 * j;
          if (p.epi == EPI_ATOMIC) {
            red_add_v4(cp, v[0], v[1], v[2], v[3]);
            red_add_v4(cp + 4, v[4], v[5], v[6], v[7]);
          } else if (p.epi != EPI_DBG_NOSTORE || v[0] == 123.456f) {
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
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  __syncwarp();  // reconverge single-lane roles before the CTA-wide barrier
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// --------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// row-major [rows, cols] fp32 with leading dimension ld (elements); box = {32 cols, box_rows}
static int make_map_kmajor(CUtensorMap* map, const float* ptr, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { atst_set_error("cuTensorMapEncodeTiled entry point not available"); return ATST_ERR_CUDA; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { atst_set_error("cuTensorMapEncodeTiled(K-major) failed: %d", (int)r); return ATST_ERR_CUDA; }
  return ATST_OK;
}

// token-major [tokens, feats] fp32 (ld elements) viewed as {32 feats, tokens, feats/32}; box {32, 32, box_feats/32}
static int make_map_mnmajor(CUtensorMap* map, const float* ptr, int tokens, int feats, int ld, int box_feats,
                            int swizzle_mode) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { atst_set_error("cuTensorMapEncodeTiled entry point not available"); return ATST_ERR_CUDA; }
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(tokens), static_cast<cuuint64_t>((feats + 31) / 32)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, 128};
  cuuint32_t box[3] = {32, 32, static_cast<cuuint32_t>(box_feats / 32)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, static_cast<CUtensorMapSwizzle>(swizzle_mode),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { atst_set_error("cuTensorMapEncodeTiled(MN-major) failed: %d", (int)r); return ATST_ERR_CUDA; }
  return ATST_OK;
}

static int g_cta_pair = 1;  // route 256-wide problems to the CTA-pair kernel (gemm2_tcgen05.cu)
void gemm_set_cta_pair(int on) { g_cta_pair = on; }
int gemm2_launch(int a_mn, int b_mn, const float* A, int lda, int a_rows, int a_cols, const float* B, int ldb, int b_rows,
                 int b_cols, const GemmParams& p, cudaStream_t stream);

static long long* g_gemm_trace = nullptr;
void gemm_set_trace(long long* buf) { g_gemm_trace = buf; }
long long* gemm_trace_ptr() { return g_gemm_trace; }
static int g_l2_prefetch = 0;  // measured: no gain (the ring is L2-bandwidth, not latency, limited)
void gemm_set_l2_prefetch(int on) { g_l2_prefetch = on; }

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int make_map_kmajor_pub(CUtensorMap* map, const float* ptr, int rows, int cols, int ld, int box_rows) {
  return make_map_kmajor(map, ptr, rows, cols, ld, box_rows);
}
int make_map_mnmajor_pub(CUtensorMap* map, const float* ptr, int tokens, int feats, int ld, int box_feats,
                         int swizzle_mode) {
  return make_map_mnmajor(map, ptr, tokens, feats, ld, box_feats, swizzle_mode);
}
int gemm_num_sms() { return num_sms(); }

// generic tensor maps for other kernels (attention_tc.cu): K-major 2-D tiles and token-major 3-D tiles
int make_map_generic_2d(CUtensorMap* map, const float* ptr, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { atst_set_error("cuTensorMapEncodeTiled entry point not available"); return ATST_ERR_CUDA; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { atst_set_error("cuTensorMapEncodeTiled(2d) failed: %d", (int)r); return ATST_ERR_CUDA; }
  return ATST_OK;
}
int make_map_generic_3d(CUtensorMap* map, const float* ptr, long long rows, int feats, int ld, int box_rows,
                        int box_chunks) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { atst_set_error("cuTensorMapEncodeTiled entry point not available"); return ATST_ERR_CUDA; }
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(feats / 32)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, 128};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_chunks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { atst_set_error("cuTensorMapEncodeTiled(3d) failed: %d", (int)r); return ATST_ERR_CUDA; }
  return ATST_OK;
}

// [S, N, feats] fp32 (contiguous) viewed as {32 floats, N rows, feats/32 chunks, S}: a box {32, box_rows, 2, 1} is one
// head's 64 columns of up to box_rows tokens of ONE sequence - rows past N are clipped on store and zero-filled on load,
// so a 128-row tile never touches the next sequence.  Standard 128B swizzle ([chunk][row][128 B] in shared memory).
int make_map_seq4d(CUtensorMap* map, const float* ptr, int S, int N, int feats, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { atst_set_error("cuTensorMapEncodeTiled entry point not available"); return ATST_ERR_CUDA; }
  cuuint64_t dims[4] = {32, static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(feats / 32), static_cast<cuuint64_t>(S)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(feats) * 4, 128, static_cast<cuuint64_t>(N) * feats * 4};
  cuuint32_t box[4] = {32, static_cast<cuuint32_t>(box_rows), 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { atst_set_error("cuTensorMapEncodeTiled(4d) failed: %d", (int)r); return ATST_ERR_CUDA; }
  return ATST_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  static bool configured = false;
  auto kfn = gemm_tf32_kernel<BLOCK_N, A_MN, B_MN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { atst_set_error("cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = m_tiles * n_tiles * (p.splits > 0 ? p.splits : 1);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  GemmParams q = p;
  q.l2_prefetch = g_l2_prefetch;
  kfn<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, q);
  return atst_check_launch("gemm_tf32_kernel");
}

// the epilogue moves 8 fp32 (one 32-byte sector) per lane and instruction
static int check_output_layout(const GemmParams& p, const char* who) {
  ATST_REQUIRE(p.N % 8 == 0 && p.ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 31) == 0,
               "%s: N and ldc must be multiples of 8 and C 32-byte aligned (N=%d ldc=%d)", who, p.N, p.ldc);
  ATST_REQUIRE(p.resid == nullptr || (p.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 31) == 0),
               "%s: resid must be 32-byte aligned with ldr %% 8 == 0", who);
  ATST_REQUIRE(p.aux == nullptr || (p.ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 31) == 0),
               "%s: aux must be 32-byte aligned with ldaux %% 8 == 0", who);
  ATST_REQUIRE(p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, "%s: bias must be 16-byte aligned", who);
  if (p.epi == EPI_GELU_H || p.epi == EPI_DGELU_H) {
    // fp16 side stream: one 32-byte sector per pair of 8-column groups, CTA-pair kernel's epilogue only
    const bool wide = (p.N % 256 == 0) || p.N > 1024;
    ATST_REQUIRE(wide && g_cta_pair, "%s: the fp16 GELU' epilogues need the CTA-pair kernel (N %% 256 == 0 or N > 1024)", who);
    ATST_REQUIRE(p.N % 16 == 0 && (p.aux == nullptr || p.ldaux % 16 == 0),
                 "%s: fp16 aux needs N and ldaux multiples of 16 (N=%d ldaux=%d)", who, p.N, p.ldaux);
  }
  return ATST_OK;
}

int gemm_nt(const float* A, int lda, const float* B, int ldb, GemmParams p, cudaStream_t stream) {
  ATST_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_nt: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  ATST_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && p.K % 4 == 0,
               "gemm_nt: K and operand leading dimensions must be multiples of 4 (K=%d lda=%d ldb=%d)", p.K, lda, ldb);
  ATST_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
               "gemm_nt: operand pointers must be 16-byte aligned");
  int rc0 = check_output_layout(p, "gemm_nt");
  if (rc0) return rc0;
  const bool wide = (p.N % 256 == 0) || p.N > 1024;
  p.splits = 1;
  if (wide && g_cta_pair) return gemm2_launch(0, 0, A, lda, p.M, p.K, B, ldb, p.N, p.K, p, stream);
  CUtensorMap ta, tb;
  int rc = make_map_kmajor(&ta, A, p.M, p.K, lda, kBlockM);
  if (rc) return rc;
  rc = make_map_kmajor(&tb, B, p.N, p.K, ldb, wide ? 256 : 128);
  if (rc) return rc;
  return wide ? launch<256, false, false>(ta, tb, p, stream) : launch<128, false, false>(ta, tb, p, stream);
}

static void mn_defaults(GemmParams& p) {
  if (p.mn_layout == 0 && p.mn_lbo == 0) {  // defaults: 128B swizzle with 32B atoms (the tf32 MN-major layout)
    p.mn_layout = 1;
    p.mn_lbo = 32 * 128;  // bytes between 32-feature chunks: 32 k rows x 128 B
    p.mn_sbo = 4 * 128;   // bytes between 4-row k groups of the 32B-atom swizzle
    p.mn_kstep = 8 * 128; // 8 k rows per MMA
    p.mn_tma_swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  }
}

int gemm_nn(const float* A, int lda, const float* B, int ldb, GemmParams p, cudaStream_t stream) {
  // C[M,N] = epi(A[M,K] . B[K,N]); B row-major [K, N] (a Linear weight [out=K, in=N] used for dgrad)
  ATST_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm_nn: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  ATST_REQUIRE(p.N % 32 == 0 && lda % 4 == 0 && ldb % 4 == 0 && p.K % 4 == 0,
               "gemm_nn: N must be a multiple of 32, K and leading dims multiples of 4 (N=%d K=%d)", p.N, p.K);
  int rc0 = check_output_layout(p, "gemm_nn");
  if (rc0) return rc0;
  mn_defaults(p);
  const bool wide = (p.N % 256 == 0) || p.N > 1024;
  p.splits = 1;
  if (wide && g_cta_pair) return gemm2_launch(0, 1, A, lda, p.M, p.K, B, ldb, p.K, p.N, p, stream);
  float* colsum = p.colsum;  // only the pair kernel's epilogue takes the column sums: separate pass here
  p.colsum = nullptr;
  CUtensorMap ta, tb;
  int rc = make_map_kmajor(&ta, A, p.M, p.K, lda, kBlockM);
  if (rc) return rc;
  rc = make_map_mnmajor(&tb, B, p.K, p.N, ldb, wide ? 256 : 128, p.mn_tma_swizzle);
  if (rc) return rc;
  rc = wide ? launch<256, false, true>(ta, tb, p, stream) : launch<128, false, true>(ta, tb, p, stream);
  if (rc == ATST_OK && colsum != nullptr) rc = colsum_accumulate(p.C, p.ldc, p.M, p.N, colsum, stream);
  return rc;
}

int gemm_tn(const float* A, int lda, const float* B, int ldb, int T, GemmParams p, cudaStream_t stream) {
  // C[M,N] (+)= A[T,M]^T B[T,N]; M, N are feature counts, T tokens.
  ATST_REQUIRE(p.M > 0 && p.N > 0 && T > 0, "gemm_tn: empty problem");
  ATST_REQUIRE(p.M % 32 == 0 && p.N % 32 == 0 && lda % 4 == 0 && ldb % 4 == 0,
               "gemm_tn: feature dims must be multiples of 32 (M=%d N=%d)", p.M, p.N);
  int rc0 = check_output_layout(p, "gemm_tn");
  if (rc0) return rc0;
  p.K = T;
  const bool wide = (p.N % 256 == 0) || p.N > 1024;
  const int bn = wide ? 256 : 128;
  mn_defaults(p);
  const bool pair = wide && g_cta_pair;
  const int bm = pair ? 2 * kBlockM : kBlockM;
  const int m_tiles = (p.M + bm - 1) / bm, n_tiles = (p.N + bn - 1) / bn;
  const int kb_total = (T + kBlockK - 1) / kBlockK;
  if (p.splits <= 0) {
    // pick the split count whose work-unit count fills whole waves of SMs best (>= 8 k-blocks per unit)
    const int t = m_tiles * n_tiles, sms = pair ? num_sms() / 2 : num_sms();
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64 && s * 8 <= kb_total; ++s) {
      const int units = t * s;
      const int waves = (units + sms - 1) / sms;
      const double eff = static_cast<double>(units) / (static_cast<double>(waves) * sms);
      if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    p.splits = best;
  }
  // no empty splits: shrink until every split owns at least one k-block
  while (p.splits > 1 && (p.splits - 1) * ((kb_total + p.splits - 1) / p.splits) >= kb_total) --p.splits;
  p.epi = EPI_ATOMIC;
  if (pair) return gemm2_launch(1, 1, A, lda, T, p.M, B, ldb, T, p.N, p, stream);
  CUtensorMap ta, tb;
  int rc = make_map_mnmajor(&ta, A, T, p.M, lda, kBlockM, p.mn_tma_swizzle);
  if (rc) return rc;
  rc = make_map_mnmajor(&tb, B, T, p.N, ldb, bn, p.mn_tma_swizzle);
  if (rc) return rc;
  return wide ? launch<256, true, true>(ta, tb, p, stream) : launch<128, true, true>(ta, tb, p, stream);
}

}  // namespace atst
