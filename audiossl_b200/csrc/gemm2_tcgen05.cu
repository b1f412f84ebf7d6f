// CTA-pair TF32 GEMM (tcgen05 cta_group::2): 256 x 256 output tile per cluster of two CTAs.
//
// Why: with one CTA per 128x256 tile (gemm_tcgen05.cu) every k-block costs 96 B/clk of UMMA operand reads plus
// 96 B/clk of TMA fill against a 128 B/clk shared-memory port, and 96 B/clk/SM of L2->SM traffic.  As a pair, each
// CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 N rows); the tensor cores of the
// two SMs read the other half from the peer's shared memory.  Per CTA: 64 B/clk of TMA fill, 64 B/clk of UMMA
// reads, 64 B/clk of L2 traffic, and the 32 KB stages make the smem ring 6 deep instead of 4.
//
// Roles per CTA (640 threads): warp 0 TMA producer (both CTAs; transactions complete on the LEADER's full
// barrier), warp 1 MMA issuer (leader only: tcgen05.mma.cta_group::2, M=256; tcgen05.commit multicast frees the
// smem slot / publishes the accumulator in both CTAs), warp 2 TMEM allocator (cta_group::2 alloc in both CTAs),
// warps 4-19 epilogue (own 128 TMEM lanes, four warps per lane quadrant; the accumulator stage is handed back on the leader's barrier, remotely
// from the follower).  Operand majors as in the 1-CTA kernel (K-major 128B swizzle, or token-major tensors via
// the 32B-atom swizzle).
#include <cooperative_groups.h>
#include "common.cuh"
#include "gemm.h"
#include "gemm_epilogue.cuh"

namespace atst {

namespace {

constexpr int kBM = 128;       // rows per CTA (256 per pair)
constexpr int kBN = 256;       // N columns per pair tile
constexpr int kBNHalf = 128;   // B rows staged per CTA
constexpr int kBK = 32;        // tf32 elements per k-block (128 B)
constexpr int kStages2 = 6;
constexpr int kEpiWarps2 = 16;                 // four per TMEM lane quadrant, each walking a 64-column quarter of the tile
constexpr int kThreads2 = 128 + 32 * kEpiWarps2;
constexpr int kA2 = kBM * 128;       // 16 KB
constexpr int kB2 = kBNHalf * 128;   // 16 KB
constexpr int kStage2 = kA2 + kB2;   // 32 KB
constexpr int kSmem2 = 1024 + kStages2 * kStage2 + 256 + 2 * 256 * 4;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// predicated forms for a converged warp (see common.cuh): only the lane with `lead` set issues
__device__ __forceinline__ void umma2_tf32_p(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate, uint32_t lead) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, L;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 L, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@L tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(lead)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc_p(uint64_t* bar, uint32_t lead) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t"
      ".reg .pred L;\n\t"
      "setp.ne.b32 L, %2, 0;\n\t"
      "@L tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(smem_u32(bar)),
      "h"(mask), "r"(lead)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at the same smem offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                   "r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// arrive on the barrier at this smem offset in CTA 0 of the pair (works from either CTA).  Default semantics
// (.release.cta): what is handed over is "this warp's tcgen05.ld of the stage have completed", ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not memory.  The .release.cluster form compiled to
// MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrive, i.e. every epilogue warp waited for its outstanding
// global stores to drain before the MMA issuer got the accumulator stage back (ncu: "membar" was 13 % of the stall
// cycles of the plain epilogue, profiles/r02_ncu_epilogue_classes.txt).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}

}  // namespace

template <bool A_MN, bool B_MN, int ECLS = ECLS_GENERIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
gemm2_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages2 * kA2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStage2);
  uint64_t* full_bar = bars;                    // [kStages2]  (the leader's are the live ones)
  uint64_t* empty_bar = bars + kStages2;        // [kStages2]  per CTA, armed by the leader's multicast commit
  uint64_t* tfull_bar = bars + 2 * kStages2;    // [2]         per CTA, multicast commit
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]         leader's: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* smem_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int m_tiles = (p.M + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (p.N + kBN - 1) / kBN;
  const int kb_total = (p.K + kBK - 1) / kBK;
  const int splits = p.splits > 0 ? p.splits : 1;
  const int kb_per_split = (kb_total + splits - 1) / splits;
  const int total_tiles = m_tiles * n_tiles * splits;
  const int first_tile = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * kEpiWarps2);  // every epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync();  // barriers of both CTAs initialised before any remote arrive / transaction
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int split = tile / (m_tiles * n_tiles);
        const int rem = tile - split * (m_tiles * n_tiles);
        const int m0 = (rem / n_tiles) * 2 * kBM + rank * kBM;   // this CTA's A rows
        const int n0 = (rem % n_tiles) * kBN + rank * kBNHalf;   // this CTA's half of the B rows
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStage2);  // both CTAs' bytes land on this barrier
          void* sa = smem_a + stage * kA2;
          void* sb = smem_b + stage * kB2;
          if (A_MN) tma2_load_3d(sa, &tmA, &full_bar[stage], 0, kb * kBK, m0 / 32);
          else      tma2_load_2d(sa, &tmA, &full_bar[stage], kb * kBK, m0);
          if (B_MN) tma2_load_3d(sb, &tmB, &full_bar[stage], 0, kb * kBK, n0 / 32);
          else      tma2_load_2d(sb, &tmB, &full_bar[stage], kb * kBK, n0);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const uint32_t idesc = make_idesc(2u, 2 * kBM, kBN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      // descriptor halves: the high word is fixed per operand layout, the low word is (address >> 4) | LBO and
      // advances by (bytes >> 4); issue is predicated on one elected lane of the converged warp (uniform datapath,
      // no per-instruction divergent region)
      const uint32_t a_hi = A_MN ? smem_desc_hi(p.mn_lbo, p.mn_sbo, p.mn_layout) : smem_desc_hi(16, 1024, 2);
      const uint32_t b_hi = B_MN ? smem_desc_hi(p.mn_lbo, p.mn_sbo, p.mn_layout) : smem_desc_hi(16, 1024, 2);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem_a), A_MN ? p.mn_lbo : 16);
      const uint32_t b_lo0 = smem_desc_lo(smem_u32(smem_b), B_MN ? p.mn_lbo : 16);
      const uint32_t a_step = (A_MN ? p.mn_kstep : 32) >> 4, b_step = (B_MN ? p.mn_kstep : 32) >> 4;
      const uint32_t lead = elect_one() ? 1u : 0u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int split = tile / (m_tiles * n_tiles);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, kb_total);
        const int tidx = (tile - first_tile) / tile_step;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && tidx >= 8 && tidx < 12;
        if (tr) p.trace[8 * (tidx - 8) + 4] = clock64();  // MMA warp arrives at the tile
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        if (tr) p.trace[8 * (tidx - 8) + 5] = clock64();  // accumulator stage free
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (kA2 >> 4), b_lo = b_lo0 + stage * (kB2 >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma2_tf32_p(d_tmem, a_lo + k * a_step, a_hi, b_lo + k * b_step, b_hi, idesc, (kb > kb0 || k > 0) ? 1u : 0u,
                         lead);
          umma2_commit_mc_p(&empty_bar[stage], lead);
          if (kb == kb1 - 1) umma2_commit_mc_p(&tfull_bar[acc], lead);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (kb1 <= kb0) umma2_commit_mc_p(&tfull_bar[acc], lead);
        if (tr) p.trace[8 * (tidx - 8) + 6] = clock64();  // last MMA of the tile issued
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    // kEpiWarps2 / 4 warps per TMEM lane quadrant, each walking one column slice of the tile
    constexpr int kGroupsPerWarp = (kBN / 8) / (kEpiWarps2 / 4);
    const int ew = (warp - 4) & 3, chalf = (warp - 4) >> 2;
    const int epi_tid = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int split = tile / (m_tiles * n_tiles);
      const int rem = tile - split * (m_tiles * n_tiles);
      const int m0 = (rem / n_tiles) * 2 * kBM + rank * kBM;
      const int n0 = (rem % n_tiles) * kBN;
      const int kb0 = split * kb_per_split;
      const bool empty_split = min(kb0 + kb_per_split, kb_total) <= kb0;
      uint64_t* tempty = &tempty_bar[acc];
      auto release = [&]() {
        if (lane == 0) mbar_arrive_leader(tempty);
      };
      const int tidx = (tile - first_tile) / tile_step;  // this CTA's tile counter
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0 && tidx >= 8 && tidx < 12;
      if (tr) p.trace[8 * (tidx - 8) + 0] = clock64();  // epilogue warp arrives at the tile
      epilogue_tile<kBN, 32 * kEpiWarps2, ECLS>(p, m0, n0, empty_split, tmem_base + acc * kBN, smem_bias + acc * 256, ew, lane,
                                          epi_tid, chalf * kGroupsPerWarp, (chalf + 1) * kGroupsPerWarp, &tfull_bar[acc],
                                          acc_phase, release, tr ? p.trace + 8 * (tidx - 8) : nullptr);
      if (tr) p.trace[8 * (tidx - 8) + 3] = clock64();  // tile stored
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync();  // nobody exits (or frees TMEM) while the peer may still signal / read
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
int make_map_kmajor_pub(CUtensorMap* map, const float* ptr, int rows, int cols, int ld, int box_rows);
int make_map_mnmajor_pub(CUtensorMap* map, const float* ptr, int tokens, int feats, int ld, int box_feats,
                         int swizzle_mode);
int gemm_num_sms();
long long* gemm_trace_ptr();

template <bool A_MN, bool B_MN, int ECLS = ECLS_GENERIC>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  auto kfn = gemm2_tf32_kernel<A_MN, B_MN, ECLS>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2);
    if (e != cudaSuccess) { atst_set_error("cudaFuncSetAttribute(gemm2): %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  const int m_tiles = (p.M + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (p.N + kBN - 1) / kBN;
  const int tiles = m_tiles * n_tiles * (p.splits > 0 ? p.splits : 1);
  const int pairs = gemm_num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  GemmParams q = p;
  q.trace = gemm_trace_ptr();
  kfn<<<grid, kThreads2, kSmem2, stream>>>(ta, tb, q);
  return atst_check_launch("gemm2_tf32_kernel");
}

// same contracts as gemm_nt / gemm_nn / gemm_tn (operands validated by the callers in gemm_tcgen05.cu)
int gemm2_launch(int a_mn, int b_mn, const float* A, int lda, int a_rows, int a_cols, const float* B, int ldb, int b_rows,
                 int b_cols, const GemmParams& p, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_map_mnmajor_pub(&ta, A, a_rows, a_cols, lda, kBM, p.mn_tma_swizzle);
  else      rc = make_map_kmajor_pub(&ta, A, a_rows, a_cols, lda, kBM);
  if (rc) return rc;
  if (b_mn) rc = make_map_mnmajor_pub(&tb, B, b_rows, b_cols, ldb, kBNHalf, p.mn_tma_swizzle);
  else      rc = make_map_kmajor_pub(&tb, B, b_rows, b_cols, ldb, kBNHalf);
  if (rc) return rc;
  // one instantiation per (operand layout, epilogue class) in use; everything else goes to the generic epilogue
  const bool simple = p.colsum == nullptr;
  const bool plain = simple && (p.epi == EPI_STORE || p.epi == EPI_SCALE || p.epi == EPI_RELU || p.epi == EPI_ATOMIC);
  if (!a_mn && !b_mn) {  // NT: forward
    if (p.epi == EPI_GELU_H) return p.aux != nullptr ? launch2<false, false, ECLS_HALF_FWD>(ta, tb, p, stream)
                                                     : launch2<false, false, ECLS_GELU>(ta, tb, p, stream);
    if (p.epi == EPI_GELU && p.aux == nullptr && simple) return launch2<false, false, ECLS_GELU>(ta, tb, p, stream);
    if (p.epi == EPI_RESID && simple) return launch2<false, false, ECLS_RESID>(ta, tb, p, stream);
    if (plain) return launch2<false, false, ECLS_PLAIN>(ta, tb, p, stream);
    if (p.epi == EPI_DGELU_H) { atst_set_error("gemm2: EPI_DGELU_H belongs to the dgrad (NN) GEMM"); return ATST_ERR_ARG; }
    return launch2<false, false>(ta, tb, p, stream);
  }
  if (!a_mn && b_mn) {  // NN: dgrad
    if (p.epi == EPI_DGELU_H) return launch2<false, true, ECLS_HALF_BWD>(ta, tb, p, stream);
    if (p.epi == EPI_GELU_H) { atst_set_error("gemm2: EPI_GELU_H belongs to the forward (NT) GEMM"); return ATST_ERR_ARG; }
    if (plain) return launch2<false, true, ECLS_PLAIN>(ta, tb, p, stream);
    return launch2<false, true>(ta, tb, p, stream);
  }
  // TN: wgrad, always split-K accumulation
  if (!plain) { atst_set_error("gemm2: epilogue %d is not available for the wgrad (TN) GEMM", p.epi); return ATST_ERR_ARG; }
  return launch2<true, true, ECLS_PLAIN>(ta, tb, p, stream);
}

}  // namespace atst
