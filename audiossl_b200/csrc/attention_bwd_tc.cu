// Attention backward on tcgen05 for sequences of up to 256 tokens (every ATST config: N <= 251).  Same math and
// outputs as attn_kernel<1> / attn_kernel<2> in attention.cu (reference: modules/transformer.py:107-121 under
// autograd); TF32 operands, fp32 accumulation in tensor memory, fp32 softmax algebra.
//
// One templated kernel, two launches per layer (scores are recomputed in each, as in the mma.sync version):
//   MODE 0 (dQ)     rows = queries, column blocks = keys.   S = Q K^T, dP = dO V^T, P = exp2(S c - L_row),
//                   dS = P (dP - delta_row),  dQ = scale * dS K
//   MODE 1 (dK dV)  rows = keys, column blocks = queries.   S^T = K Q^T, dP^T = V dO^T, P^T = exp2(S^T c - L_col),
//                   dS^T = P^T (dP^T - delta_col),  dV = P^T dO,  dK = scale * dS^T Q
//
// One CTA per (sequence, head); it walks the 128-row tiles of the "row" operand and, per tile, the 64-wide
// quarters of the "column" operand.  Operands are TMA-loaded ONCE in the token-major 32B-atom 128B swizzle and read
// by the tensor core both K-major (contraction over the head dim: S, dP) and MN-major (contraction over tokens:
// dQ / dK / dV) from the same bytes (profiles/r01_umma_operand_probe.log).  P and dS never touch shared memory: the
// compute warps overwrite the S / dP quarter in tensor memory with tf32(P) / tf32(dS) (tcgen05.st) and the second
// stage MMAs take that as their TMEM A operand.
//
// Shared memory (226 KB): R1, R2 = the row tile's two operands [2 k-chunks][128 rows][128 B]; Y1, Y2 = the column
// operands [4 quarters][2 k-chunks][64 rows][128 B]; a 32 KB staging tile for the TMA store of the results; per-column
// lse / delta for MODE 1.
// Tensor memory (512 columns): accumulators acc1 [0,64) (dQ | dK), acc2 [64,128) (dV); three ring slots of
// {S quarter, dP quarter} at 128 + 128 s.
// Warps: 0 TMA producer, 1 MMA issuer for S / dP, 3 MMA issuer for the second stage, 2 TMEM allocator, 4-11 compute (thread = row = TMEM lane; the two
// warps of a lane quadrant split each quarter's 64 columns), 12-15 epilogue (accumulators -> swizzled staging tile ->
// one TMA store per result, clipped at the sequence end by the 4-D tensor map).
#include "common.cuh"

namespace atst {

namespace {

constexpr int kR1 = 0;
constexpr int kR2 = 32 * 1024;
constexpr int kY1 = 64 * 1024;
constexpr int kY2 = 128 * 1024;
constexpr int kStage = 192 * 1024;         // [2 chunks][128 rows][128 B] result tile for the TMA store
constexpr int kStat = 224 * 1024;          // [2][256] floats
constexpr int kBars = kStat + 2 * 256 * 4;  // mbarriers
constexpr int kSmemBwd = kBars + 256;       // no alignment slack: the dynamic segment is declared 1024-byte aligned
constexpr int kSlots = 3;
constexpr uint32_t kKLbo = 4096, kSbo = 512, kLayout = 1;  // 32B-atom 128B swizzle (see the operand probe)

struct BwdTcParams {
  float* dqkv;         // [S*N, 3D]
  const float* lse;    // [S, H, N] log2 domain
  const float* delta;  // [S, H, N]
  const int* lengths;
  int N, H, D;
  float scale;
  int prefetch_dist, num_items;  // see attention_tc.cu
  long long* trace;  // bring-up: clock64() timeline of one CTA (tools/bringup.py attn_trace), or null
  int trace_seq;
};

}  // namespace

// timeline slots: 0 start | 1+4i.. MMA warp {produce ready, produce issued, consume ready, consume issued} of work
// quarter i | 40+2i.. compute warp 0 {quarter ready, quarter handed back} | 60+2t.. {accumulator ready, tile stored}
#define ATTN_TRACE(idx)                                       \
  do {                                                        \
    if (tracing && lane == 0) p.trace[(idx)] = clock64();     \
  } while (0)

template <int MODE>
__global__ void __launch_bounds__(512, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQkvR, const __grid_constant__ CUtensorMap tmQkvY,
                   const __grid_constant__ CUtensorMap tmDoR, const __grid_constant__ CUtensorMap tmDoY,
                   const __grid_constant__ CUtensorMap tmOut, BwdTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled tiles need 1024-byte alignment
  float* sL = reinterpret_cast<float*>(smem + kStat);
  float* sD = sL + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBars);
  uint64_t* bar_r = bars + 0;      // R1 (first row operand) loaded
  uint64_t* bar_rfree = bars + 1;  // every S MMA of the tile has read R1
  uint64_t* bar_acc = bars + 2;
  uint64_t* bar_accfree = bars + 3;
  uint64_t* bar_y = bars + 4;       // [4]
  uint64_t* bar_full = bars + 8;    // [3]
  uint64_t* bar_ds = bars + 11;     // [3]
  uint64_t* bar_free = bars + 14;   // [3]
  uint64_t* bar_yfree = bars + 17;  // [4]
  uint64_t* bar_r2 = bars + 21;      // R2 (second row operand) loaded
  uint64_t* bar_r2free = bars + 22;  // every dP MMA of the tile has read R2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, D = p.D, H = p.H;
  const int tiles = (N + 127) >> 7;
  // Persistent CTA: items (sequence, head) blockIdx.x, blockIdx.x + gridDim.x, ...  Every role walks the same item
  // sequence with running counters (work quarters gi -> ring slot / use count, row tiles gt, per-quarter parity
  // bits), so the next item's operands stream in while this item's last quarters are still being computed.
  auto item_len = [&](int s) {
    int len = p.lengths ? p.lengths[s] : N;
    return (len <= 0 || len > N) ? N : len;  // see attention.cu
  };
  auto item_nq = [&](int len) { return (((MODE == 0) ? len : N) + 63) >> 6; };  // dQ only needs keys < len

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQkvR);
    tma_prefetch_desc(&tmQkvY);
    tma_prefetch_desc(&tmDoR);
    tma_prefetch_desc(&tmDoY);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_r, 1);
    mbar_init(bar_rfree, 1);
    mbar_init(bar_r2, 1);
    mbar_init(bar_r2free, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_accfree, 128);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bar_y[i], 1);
      mbar_init(&bar_yfree[i], 1);
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_ds[i], 256);
      mbar_init(&bar_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc1 = tmem_base, tm_acc2 = tmem_base + 64;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t gt = 0, ypar = 0, yused = 0;  // row tiles loaded so far; per column quarter: load-count parity / loaded before
      int iter = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
        const bool tracing = p.trace != nullptr && blockIdx.x == 0 && iter == p.trace_seq;
        const int s = item / H, h = item - s * H;
        const int row0 = s * N, nq = item_nq(item_len(s));
        const int cq = (h * 64) >> 5, ck = cq + (D >> 5), cv = cq + (D >> 4), cdo = cq;
        for (int t = 0; t < tiles; ++t, ++gt) {
          // the two row operands are refilled separately: R1 as soon as the previous tile's last S MMAs are done,
          // R2 after its last dP MMAs (the reload sits on the tile-boundary critical path)
          const int r = row0 + t * 128;
          if (gt > 0) mbar_wait(bar_rfree, (gt - 1) & 1);
          ATTN_TRACE(70 + t);
          mbar_expect_tx(bar_r, 32 * 1024);
          tma_load_3d(smem + kR1, &tmQkvR, bar_r, 0, r, MODE == 0 ? cq : ck);
          if (gt > 0) mbar_wait(bar_r2free, (gt - 1) & 1);
          mbar_expect_tx(bar_r2, 32 * 1024);
          if (MODE == 0) tma_load_3d(smem + kR2, &tmDoR, bar_r2, 0, r, cdo);
          else           tma_load_3d(smem + kR2, &tmQkvR, bar_r2, 0, r, cv);
          // the loads that follow are known exactly (next row tile, then the next item's columns): pull them into
          // L2 now so that they hit when the buffers free up (their latency sits on the tile-boundary critical path)
          if (p.prefetch_dist) {
            int ns = s, nh = h, nt = t + 1;
            if (nt == tiles) {
              const int nitem = item + gridDim.x;
              nt = nitem < p.num_items ? 0 : -1;
              ns = nitem / H;
              nh = nitem - ns * H;
            }
            if (nt >= 0) {
              const int nr = ns * N + nt * 128, nc = (nh * 64) >> 5;
              tma_prefetch_3d(&tmQkvR, 0, nr, MODE == 0 ? nc : nc + (D >> 5));
              if (MODE == 0) tma_prefetch_3d(&tmDoR, 0, nr, nc);
              else           tma_prefetch_3d(&tmQkvR, 0, nr, nc + (D >> 4));
              if (nt == 0) {
                for (int q = 0; q < ((N + 63) >> 6); ++q) {
                  tma_prefetch_3d(&tmQkvY, 0, ns * N + q * 64, MODE == 0 ? nc + (D >> 5) : nc);
                  if (MODE == 0) tma_prefetch_3d(&tmQkvY, 0, ns * N + q * 64, nc + (D >> 4));
                  else           tma_prefetch_3d(&tmDoY, 0, ns * N + q * 64, nc);
                }
              }
            }
          }
          if (t != 0) continue;
          for (int q = 0; q < nq; ++q) {
            // k-th load of this quarter: the (k-1)-th use's last second-stage MMAs (previous item) have completed
            if ((yused >> q) & 1) mbar_wait(&bar_yfree[q], ((ypar >> q) & 1) ^ 1);
            mbar_expect_tx(&bar_y[q], 32 * 1024);
            const int ry = row0 + q * 64;
            if (MODE == 0) {
              tma_load_3d(smem + kY1 + q * 16384, &tmQkvY, &bar_y[q], 0, ry, ck);
              tma_load_3d(smem + kY2 + q * 16384, &tmQkvY, &bar_y[q], 0, ry, cv);
            } else {
              tma_load_3d(smem + kY1 + q * 16384, &tmQkvY, &bar_y[q], 0, ry, cq);
              tma_load_3d(smem + kY2 + q * 16384, &tmDoY, &bar_y[q], 0, ry, cdo);
            }
            ypar ^= 1u << q;
            yused |= 1u << q;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer 1: S / dP quarters into the ring
    const uint32_t idesc_s = make_idesc(2u, 128, 64, 0u, 0u);  // A, B K-major
    const uint32_t hi = smem_desc_hi(kKLbo, kSbo, kLayout);
    const uint32_t r1 = smem_desc_lo(smem_u32(smem + kR1), kKLbo), r2 = smem_desc_lo(smem_u32(smem + kR2), kKLbo);
    const uint32_t y1 = smem_desc_lo(smem_u32(smem + kY1), kKLbo), y2 = smem_desc_lo(smem_u32(smem + kY2), kKLbo);
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t gi = 0, gt = 0, ypar = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int nq = item_nq(item_len(item / H));
      const bool tracing = p.trace != nullptr && blockIdx.x == 0 && iter == p.trace_seq;
      ATTN_TRACE(0);
      for (int t = 0; t < tiles; ++t, ++gt) {
        for (int q = 0; q < nq; ++q, ++gi) {
          const int i = t * nq + q;
          const uint32_t slot = gi % kSlots, u = gi / kSlots;
          if (q == 0) mbar_wait(bar_r, gt & 1);
          if (t == 0) {
            mbar_wait(&bar_y[q], (ypar >> q) & 1);
            ypar ^= 1u << q;
          }
          if (u > 0) mbar_wait(&bar_free[slot], (u - 1) & 1);  // the slot's second-stage MMAs have read it
          tc_fence_after();
          ATTN_TRACE(1 + 4 * i);
          const uint32_t tm_s = tmem_base + 128 + slot * 128, tm_dp = tm_s + 64;
          const uint32_t yq = q * (16384 >> 4);
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ss_p(tm_s, r1 + ((kc * 16384 + k * 32) >> 4), hi, y1 + yq + ((kc * 8192 + k * 32) >> 4), hi,
                             idesc_s, (kc | k) ? 1u : 0u, leader);
          if (q == nq - 1) umma_commit_p(bar_rfree, leader);  // R1 may be refilled
          if (q == 0) {
            mbar_wait(bar_r2, gt & 1);
            tc_fence_after();
          }
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ss_p(tm_dp, r2 + ((kc * 16384 + k * 32) >> 4), hi, y2 + yq + ((kc * 8192 + k * 32) >> 4), hi,
                             idesc_s, (kc | k) ? 1u : 0u, leader);
          umma_commit_p(&bar_full[slot], leader);
          if (q == nq - 1) umma_commit_p(bar_r2free, leader);
          ATTN_TRACE(2 + 4 * i);
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ MMA issuer 2: second-stage MMAs (A operand in TMEM)
    const uint32_t idesc_acc = make_idesc(2u, 128, 64, 0u, 1u);  // A in TMEM (K-major), B MN-major
    const uint32_t hi = smem_desc_hi(8192, kSbo, kLayout);
    const uint32_t y1 = smem_desc_lo(smem_u32(smem + kY1), 8192), y2 = smem_desc_lo(smem_u32(smem + kY2), 8192);
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t gi = 0, gt = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int nq = item_nq(item_len(item / H));
      const bool tracing = p.trace != nullptr && blockIdx.x == 0 && iter == p.trace_seq;
      for (int t = 0; t < tiles; ++t, ++gt) {
        for (int q = 0; q < nq; ++q, ++gi) {
          const int i = t * nq + q;
          const uint32_t slot = gi % kSlots, u = gi / kSlots;
          mbar_wait(&bar_ds[slot], u & 1);
          if (q == 0 && gt > 0) mbar_wait(bar_accfree, (gt - 1) & 1);  // the previous tile's accumulators were read out
          tc_fence_after();
          ATTN_TRACE(3 + 4 * i);
          const uint32_t tm_s = tmem_base + 128 + slot * 128, tm_dp = tm_s + 64;
          const uint32_t yq = q * (16384 >> 4);
          const uint32_t acc0 = q ? 1u : 0u;
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            if (MODE == 1)  // dV += P^T dO
              umma_tf32_ts_p(tm_acc2, tm_s + k8 * 8, y2 + yq + ((k8 * 1024) >> 4), hi, idesc_acc, k8 ? 1u : acc0, leader);
            // dQ += dS K   |   dK += dS^T Q
            umma_tf32_ts_p(tm_acc1, tm_dp + k8 * 8, y1 + yq + ((k8 * 1024) >> 4), hi, idesc_acc, k8 ? 1u : acc0, leader);
          }
          umma_commit_p(&bar_free[slot], leader);
          if (t == tiles - 1) umma_commit_p(&bar_yfree[q], leader);  // the item's last reader of this column quarter
          if (q == nq - 1) umma_commit_p(bar_acc, leader);
          ATTN_TRACE(4 + 4 * i);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------ compute warps (thread = row = TMEM lane)
    const int cw = warp - 4;
    const int quad = cw & 3, half = cw >> 2;
    const int rt = quad * 32 + lane;  // row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    uint32_t gi = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int s = item / H, h = item - s * H;
      const int len = item_len(s), nq = item_nq(len);
      const bool tracing = p.trace != nullptr && blockIdx.x == 0 && iter == p.trace_seq;
      const size_t stat_base = (static_cast<size_t>(s) * H + h) * N;
      if (MODE == 1) {
        const int i = threadIdx.x - 128;  // 0..255
        const float lv = i < N ? p.lse[stat_base + i] : 0.f, dv0 = i < N ? p.delta[stat_base + i] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // every compute warp is done with the previous item's stats
        sL[i] = lv;
        sD[i] = dv0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int t = 0; t < tiles; ++t) {
        const int row = t * 128 + rt;
        float Lr = 0.f, dr = 0.f;
        if (MODE == 0 && row < N) {
          Lr = p.lse[stat_base + row];
          dr = p.delta[stat_base + row];
        }
        const bool row_ok = (MODE == 0) ? (row < N) : (row < len);
        for (int q = 0; q < nq; ++q, ++gi) {
          const int i = t * nq + q;
          const uint32_t slot = gi % kSlots, u = gi / kSlots;
          mbar_wait(&bar_full[slot], u & 1);
          tc_fence_after();
          if (cw == 0) ATTN_TRACE(40 + 2 * i);
          const uint32_t tm_s = tmem_base + 128 + slot * 128 + lane_addr + half * 32, tm_dp = tm_s + 64;
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(tm_s, sv);
          tmem_ld_32x32(tm_dp, dv);
          tmem_ld_wait();
          const int col0 = q * 64 + half * 32;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            // MODE 1: per-column lse / delta, four columns per 128-bit shared-memory broadcast load
            float4 L4 = make_float4(Lr, Lr, Lr, Lr), d4 = make_float4(dr, dr, dr, dr);
            if (MODE == 1) {
              L4 = *reinterpret_cast<const float4*>(sL + col0 + 4 * j4);
              d4 = *reinterpret_cast<const float4*>(sD + col0 + 4 * j4);
            }
            const float Lc4[4] = {L4.x, L4.y, L4.z, L4.w}, dc4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int j = 4 * j4 + i4, col = col0 + j;
              const bool ok = row_ok && col < ((MODE == 0) ? len : N);
              const float pv = ok ? ex2_approx(fmaf(__uint_as_float(sv[j]), c, -Lc4[i4])) : 0.f;
              const float ds = pv * (__uint_as_float(dv[j]) - dc4[i4]);
              if (MODE == 1) sv[j] = __float_as_uint(round_tf32(pv));
              dv[j] = __float_as_uint(round_tf32(ds));
            }
          }
          if (MODE == 1) tmem_st_32x32(tm_s, sv);
          tmem_st_32x32(tm_dp, dv);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bar_ds[slot]);
          if (cw == 0) ATTN_TRACE(41 + 2 * i);
        }
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------ epilogue warps: accumulators -> staging -> TMA store
    const int quad = warp & 3;
    const int rt = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t stage = smem_u32(smem + kStage);
    const bool leader = threadIdx.x == 12 * 32;
    // v: the thread's 64 result columns -> staging row rt, 128B-swizzled 16-byte units, both 32-column chunks
    auto stage_row = [&](const uint32_t (&v0)[32], const uint32_t (&v1)[32], float scale) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const uint32_t dst = stage + ch * 16384 + rt * 128;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint32_t* v = ch ? v1 : v0;
          st_shared_v4(dst + ((j4 ^ (rt & 7)) << 4), round_tf32(__uint_as_float(v[4 * j4]) * scale),
                       round_tf32(__uint_as_float(v[4 * j4 + 1]) * scale),
                       round_tf32(__uint_as_float(v[4 * j4 + 2]) * scale),
                       round_tf32(__uint_as_float(v[4 * j4 + 3]) * scale));
        }
      }
    };
    bool pending = false;  // a bulk store may still be reading the staging tile
    uint32_t gt = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int s = item / H, h = item - s * H;
      const int cq = (h * 64) >> 5, ck = cq + (D >> 5), cv = cq + (D >> 4);
      const bool tracing = p.trace != nullptr && blockIdx.x == 0 && iter == p.trace_seq;
      for (int t = 0; t < tiles; ++t, ++gt) {
        mbar_wait(bar_acc, gt & 1);
        tc_fence_after();
        if (warp == 12) ATTN_TRACE(60 + 2 * t);
        for (int r = 0; r < (MODE == 1 ? 2 : 1); ++r) {
          uint32_t a0[32], a1[32];
          const uint32_t tm_acc = (r == 0 ? tm_acc1 : tm_acc2) + lane_addr;
          tmem_ld_32x32(tm_acc, a0);
          tmem_ld_32x32(tm_acc + 32, a1);
          tmem_ld_wait();
          if (r == (MODE == 1 ? 1 : 0)) {
            tc_fence_before();
            mbar_arrive(bar_accfree);  // the next tile's second-stage MMAs may overwrite the accumulators
          }
          if (pending) {
            if (leader) tma_store_wait_read();
            asm volatile("bar.sync 2, 128;" ::: "memory");
          }
          stage_row(a0, a1, r == 0 ? p.scale : 1.0f);  // dQ | dK, then dV
          fence_proxy_async_smem();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (leader) {
            const int chunk = (MODE == 0) ? cq : (r == 0 ? ck : cv);
            tma_store_4d(&tmOut, smem + kStage, 0, t * 128, chunk, s);
            tma_store_commit();
          }
          pending = true;
        }
        if (warp == 12) ATTN_TRACE(61 + 2 * t);
      }
    }
    if (leader) tma_store_wait_read();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
int make_map_generic_3d(CUtensorMap* map, const float* ptr, long long rows, int feats, int ld, int box_rows,
                        int box_chunks);
int make_map_seq4d(CUtensorMap* map, const float* ptr, int S, int N, int feats, int box_rows);
int attention_l2_prefetch_enabled();
int gemm_num_sms();
int attention_delta(const float* o, const float* d_o, float* delta, int S, int N, int H, cudaStream_t stream);

static long long* g_trace = nullptr;
static int g_trace_seq = 0, g_trace_mode = 0;
void attention_set_trace(long long* buf, int seq, int mode) { g_trace = buf; g_trace_seq = seq; g_trace_mode = mode; }

int attention_backward_tc(const float* qkv, const float* o, const float* d_o, const float* lse, float* delta_ws,
                          float* dqkv, const int* lengths, int S, int N, int H, cudaStream_t stream) {
  const int D = H * 64;
  ATST_REQUIRE(N <= 256, "attention_backward_tc: N=%d > 256", N);
  ATST_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_o) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dqkv) & 31) == 0,
               "attention_backward_tc: qkv / d_o must be 16-byte and dqkv 32-byte aligned");
  int rc = attention_delta(o, d_o, delta_ws, S, N, H, stream);
  if (rc) return rc;
  const long long rows = static_cast<long long>(S) * N;
  CUtensorMap tqr, tqy, tdr, tdy, tout;
  if ((rc = make_map_seq4d(&tout, dqkv, S, N, 3 * D, 128))) return rc;
  if ((rc = make_map_generic_3d(&tqr, qkv, rows, 3 * D, 3 * D, 128, 2))) return rc;
  if ((rc = make_map_generic_3d(&tqy, qkv, rows, 3 * D, 3 * D, 64, 2))) return rc;
  if ((rc = make_map_generic_3d(&tdr, d_o, rows, D, D, 128, 2))) return rc;
  if ((rc = make_map_generic_3d(&tdy, d_o, rows, D, D, 64, 2))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd);
    if (e != cudaSuccess) { atst_set_error("attn_bwd_tc smem attr: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  BwdTcParams p{};
  p.dqkv = dqkv; p.lse = lse; p.delta = delta_ws; p.lengths = lengths; p.N = N; p.H = H; p.D = D; p.scale = 0.125f;
  p.prefetch_dist = attention_l2_prefetch_enabled();  // L2 prefetch of the CTA's next loads
  p.num_items = S * H;
  const int grid = p.num_items < gemm_num_sms() ? p.num_items : gemm_num_sms();  // persistent: one CTA per SM
  p.trace_seq = g_trace_seq;
  p.trace = g_trace_mode == 0 ? g_trace : nullptr;
  attn_bwd_tc_kernel<0><<<grid, 512, kSmemBwd, stream>>>(tqr, tqy, tdr, tdy, tout, p);
  rc = atst_check_launch("attn_bwd_tc_kernel<dQ>");
  if (rc) return rc;
  p.trace = g_trace_mode == 1 ? g_trace : nullptr;
  attn_bwd_tc_kernel<1><<<grid, 512, kSmemBwd, stream>>>(tqr, tqy, tdr, tdy, tout, p);
  return atst_check_launch("attn_bwd_tc_kernel<dKdV>");
}

}  // namespace atst
