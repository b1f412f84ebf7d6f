// Attention forward on tcgen05 (5th-gen tensor cores) for sequences of up to 256 tokens (every ATST config:
// N <= 251).  Same math and outputs as attn_kernel<0> in attention.cu: o = softmax(q k^T / 8 + key-padding) v,
// lse in the log2 domain; TF32 operands, fp32 accumulation in tensor memory, fp32 softmax.
//
// Persistent CTAs (one per SM) walk the (sequence, head) items; an item is both 128-query tiles, K and V are loaded
// once per item and the next item's operands stream in as buffers free up.  Operands arrive by TMA in the
// token-major 32B-atom 128B swizzle: Q and K are read K-major (S = Q K^T), V MN-major (O = P V).  The probabilities
// never touch shared memory: the compute warps overwrite each 64-key quarter of S in tensor memory with tf32(P)
// (tcgen05.st) and the P V MMAs take it as their TMEM A operand.
//
// Tensor memory (512 columns): a ring of six 64-column quarter slots [0,384) - a tile's S occupies up to four, so the
// next tile's first quarters are produced while this tile's softmax runs - and the two tiles' O accumulators
// [384,448), [448,512).
// The row maximum is taken over the whole row first (S is read twice from tensor memory), so P needs no rescaling.
// Warps: 0 TMA producer, 1 MMA issuer for S, 3 MMA issuer for P V, 2 TMEM allocator, 4-11 softmax (thread = query row
// = TMEM lane; the two warps of a lane quadrant split each quarter's 64 columns and exchange their partial row
// max / row sum through shared memory), 12-15 epilogue (O / l -> the dead second Q buffer, 128B-swizzled -> one TMA
// store per tile, clipped at the sequence end by the 4-D tensor map; lse).
#include "common.cuh"

namespace atst {

namespace {

constexpr int kQ = 0;                    // [2 tiles][2 k-chunks][128 rows][128 B]
constexpr int kK = 64 * 1024;            // [4 quarters][2 k-chunks][64 rows][128 B]
constexpr int kV = 128 * 1024;           // same, read MN-major
constexpr int kStatF = 192 * 1024;       // sXm [2 tiles][2 halves][128], sXl [2][2][128]
constexpr int kBarsF = kStatF + 2 * 2 * 2 * 128 * 4;
constexpr int kSmemFwd = kBarsF + 512;
constexpr int kRing = 6;
constexpr uint32_t kLboK = 4096, kSboF = 512, kLayoutF = 1;  // 32B-atom 128B swizzle (see the operand probe)

struct AttnTcParams {
  float* lse;
  const int* lengths;
  int N, H, D;
  float scale;
  int prefetch_dist;  // non-zero: L2-prefetch the CTA's next item
  int num_items;      // S * H
};

}  // namespace

__global__ void __launch_bounds__(512, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmY,
                   const __grid_constant__ CUtensorMap tmOut, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled tiles need 1024-byte alignment
  float* sXm = reinterpret_cast<float*>(smem + kStatF);  // [tile][half][row]
  float* sXl = sXm + 2 * 2 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarsF);
  uint64_t* bar_q = bars + 0;        // [2]  Q tile loaded
  uint64_t* bar_k = bars + 2;        // [4]  K quarter loaded
  uint64_t* bar_v = bars + 6;        // [4]  V quarter loaded
  uint64_t* bar_full = bars + 10;    // [6]  S quarter in the ring slot
  uint64_t* bar_p = bars + 16;       // [6]  P written over it
  uint64_t* bar_free = bars + 22;    // [6]  P V MMAs have read it
  uint64_t* bar_o = bars + 28;       // [2]  O accumulator of tile t complete
  uint64_t* bar_stats = bars + 30;   // [2]  row max / row sum of tile t published
  uint64_t* bar_ofree = bars + 32;   // [2]  epilogue has read O and the stats of tile t
  uint64_t* bar_kfree = bars + 34;   // [4]  last S MMA of the item on K quarter q done
  uint64_t* bar_vfree = bars + 38;   // [4]  last P V MMA of the item on V quarter q done
  uint64_t* bar_q0free = bars + 42;  //      S MMAs of tile 0 done: Q buffer 0 may be refilled
  uint64_t* bar_q1dead = bars + 43;  //      S MMAs of tile 1 done: Q buffer 1 may serve as the output staging tile
  uint64_t* bar_q1free = bars + 44;  //      the item's last output store has left Q buffer 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 45);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, D = p.D, H = p.H;
  const int tiles = (N + 127) >> 7;
  // Persistent CTA over the items (sequence, head) blockIdx.x, blockIdx.x + gridDim.x, ...: every role walks the
  // same sequence with running counters, so the next item's Q / K / V stream in while this item is still computed.
  auto item_len = [&](int s) {
    int len = p.lengths ? p.lengths[s] : N;
    return (len <= 0 || len > N) ? N : len;  // see attention.cu
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q[i], 1);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_stats[i], 256);
      mbar_init(&bar_ofree[i], 128);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bar_k[i], 1);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_kfree[i], 1);
      mbar_init(&bar_vfree[i], 1);
    }
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_p[i], 256);
      mbar_init(&bar_free[i], 1);
    }
    mbar_init(bar_q0free, 1);
    mbar_init(bar_q1dead, 1);
    mbar_init(bar_q1free, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_o = tmem_base + 384;  // + 64 * tile

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: Q0, K quarters, V quarters, Q1
    if (lane == 0) {
      uint32_t kpar = 0, kused = 0, vpar = 0, vused = 0;  // per quarter: load-count parity / loaded before
      int iter = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
        const int s = item / H, h = item - s * H;
        const int row0 = s * N, nq = (item_len(s) + 63) >> 6;
        const int cq = (h * 64) >> 5, ck = cq + (D >> 5), cv = cq + (D >> 4);
        if (iter > 0) mbar_wait(bar_q0free, (iter - 1) & 1);
        mbar_expect_tx(&bar_q[0], 32 * 1024);
        tma_load_3d(smem + kQ, &tmR, &bar_q[0], 0, row0, cq);
        for (int q = 0; q < nq; ++q) {
          if ((kused >> q) & 1) mbar_wait(&bar_kfree[q], ((kpar >> q) & 1) ^ 1);
          mbar_expect_tx(&bar_k[q], 16 * 1024);
          tma_load_3d(smem + kK + q * 16384, &tmY, &bar_k[q], 0, row0 + q * 64, ck);
          kpar ^= 1u << q;
          kused |= 1u << q;
        }
        for (int q = 0; q < nq; ++q) {
          if ((vused >> q) & 1) mbar_wait(&bar_vfree[q], ((vpar >> q) & 1) ^ 1);
          mbar_expect_tx(&bar_v[q], 16 * 1024);
          tma_load_3d(smem + kV + q * 16384, &tmY, &bar_v[q], 0, row0 + q * 64, cv);
          vpar ^= 1u << q;
          vused |= 1u << q;
        }
        if (tiles > 1) {
          if (iter > 0) mbar_wait(bar_q1free, (iter - 1) & 1);  // the previous item's output staging has drained
          mbar_expect_tx(&bar_q[1], 32 * 1024);
          tma_load_3d(smem + kQ + 32768, &tmR, &bar_q[1], 0, row0 + 128, cq);
        }
        // the next item's operands are known exactly: pull them into L2 while this one is computed
        const int nitem = item + gridDim.x;
        if (p.prefetch_dist && nitem < p.num_items) {
          const int s2 = nitem / H, h2 = nitem - s2 * H;
          const int r2 = s2 * N, c2 = (h2 * 64) >> 5;
          for (int t = 0; t < tiles; ++t) tma_prefetch_3d(&tmR, 0, r2 + t * 128, c2);
          for (int q = 0; q < ((N + 63) >> 6); ++q) {
            tma_prefetch_3d(&tmY, 0, r2 + q * 64, c2 + (D >> 5));
            tma_prefetch_3d(&tmY, 0, r2 + q * 64, c2 + (D >> 4));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer 1: S quarters into the ring
    const uint32_t idesc_s = make_idesc(2u, 128, 64, 0u, 0u);
    const uint32_t hi = smem_desc_hi(kLboK, kSboF, kLayoutF);
    const uint32_t qd = smem_desc_lo(smem_u32(smem + kQ), kLboK), kd = smem_desc_lo(smem_u32(smem + kK), kLboK);
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t gi = 0, kpar = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int nq = (item_len(item / H) + 63) >> 6;
      for (int t = 0; t < tiles; ++t) {
        for (int q = 0; q < nq; ++q, ++gi) {
          const uint32_t slot = gi % kRing, u = gi / kRing;
          if (q == 0) mbar_wait(&bar_q[t], iter & 1);
          if (t == 0) {
            mbar_wait(&bar_k[q], (kpar >> q) & 1);
            kpar ^= 1u << q;
          }
          if (u > 0) mbar_wait(&bar_free[slot], (u - 1) & 1);  // the slot's P V MMAs have read it
          tc_fence_after();
          const uint32_t tm_s = tmem_base + slot * 64;
          const uint32_t qa = qd + t * (32768 >> 4), ka = kd + q * (16384 >> 4);
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ss_p(tm_s, qa + ((kc * 16384 + k * 32) >> 4), hi, ka + ((kc * 8192 + k * 32) >> 4), hi, idesc_s,
                             (kc | k) ? 1u : 0u, leader);
          umma_commit_p(&bar_full[slot], leader);
          if (t == tiles - 1) umma_commit_p(&bar_kfree[q], leader);
          if (q == nq - 1) {
            if (t == 0) umma_commit_p(bar_q0free, leader);
            if (t == 1) umma_commit_p(bar_q1dead, leader);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ MMA issuer 2: O += P V (P in tensor memory)
    const uint32_t idesc_pv = make_idesc(2u, 128, 64, 0u, 1u);
    const uint32_t hi = smem_desc_hi(8192, kSboF, kLayoutF);
    const uint32_t vd = smem_desc_lo(smem_u32(smem + kV), 8192);
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t gi = 0, vpar = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int nq = (item_len(item / H) + 63) >> 6;
      for (int t = 0; t < tiles; ++t) {
        for (int q = 0; q < nq; ++q, ++gi) {
          const uint32_t slot = gi % kRing, u = gi / kRing;
          mbar_wait(&bar_p[slot], u & 1);
          if (t == 0) {
            mbar_wait(&bar_v[q], (vpar >> q) & 1);
            vpar ^= 1u << q;
          }
          if (q == 0 && iter > 0) mbar_wait(&bar_ofree[t], (iter - 1) & 1);  // the previous item's O was read out
          tc_fence_after();
          const uint32_t tm_p = tmem_base + slot * 64;
          const uint32_t va = vd + q * (16384 >> 4);
          const uint32_t acc0 = q ? 1u : 0u;
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8)
            umma_tf32_ts_p(tm_o + t * 64, tm_p + k8 * 8, va + ((k8 * 1024) >> 4), hi, idesc_pv, k8 ? 1u : acc0, leader);
          umma_commit_p(&bar_free[slot], leader);
          if (t == tiles - 1) umma_commit_p(&bar_vfree[q], leader);
          if (q == nq - 1) umma_commit_p(&bar_o[t], leader);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------ softmax warps (thread = query row = TMEM lane)
    const int cw = warp - 4;
    const int quad = cw & 3, half = cw >> 2;
    const int rt = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    uint32_t gi = 0;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int len = item_len(item / H), nq = (len + 63) >> 6;
      for (int t = 0; t < tiles; ++t) {
        // pass 1: row maximum over this warp's half of every quarter
        float m = -INFINITY;
        for (int q = 0; q < nq; ++q) {
          const uint32_t g = gi + q, slot = g % kRing, u = g / kRing;
          mbar_wait(&bar_full[slot], u & 1);
          tc_fence_after();
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + slot * 64 + half * 32 + lane_addr, v);
          tmem_ld_wait();
          const int col0 = q * 64 + half * 32;
          if (col0 + 32 <= len) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < len) m = fmaxf(m, __uint_as_float(v[j]));
          }
        }
        if (iter > 0) mbar_wait(&bar_ofree[t], (iter - 1) & 1);  // the epilogue is done with the previous item's stats
        float* xm = sXm + t * 256;
        xm[half * 128 + rt] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(xm[rt], xm[128 + rt]);  // at least one valid key per row (len >= 1) lies in half 0 of quarter 0
        const float mc = m * c;
        // pass 2: probabilities, written back over S as the TMEM operand of the P V MMAs
        float l = 0.f;
        for (int q = 0; q < nq; ++q) {
          const uint32_t slot = (gi + q) % kRing;
          const uint32_t ta = tmem_base + slot * 64 + half * 32 + lane_addr;
          uint32_t v[32];
          tmem_ld_32x32(ta, v);
          tmem_ld_wait();
          const int col0 = q * 64 + half * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pv = (col0 + j < len) ? ex2_approx(fmaf(__uint_as_float(v[j]), c, -mc)) : 0.f;
            l += pv;
            v[j] = __float_as_uint(round_tf32(pv));
          }
          tmem_st_32x32(ta, v);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bar_p[slot]);
        }
        gi += nq;
        sXl[t * 256 + half * 128 + rt] = l;
        mbar_arrive(&bar_stats[t]);  // publishes this tile's sXm / sXl to the epilogue warps
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------ epilogue warps: O / l -> staging -> TMA store, lse
    const int quad = warp & 3;
    const int rt = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const bool leader = threadIdx.x == 12 * 32;
    const float c = p.scale * 1.4426950408889634f;
    // staging = Q buffer 1 (both tiles): dead once the S MMAs of tile 1 have completed
    const uint32_t stage = smem_u32(smem + kQ + 32768);
    bool pending = false;
    int iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++iter) {
      const int s = item / H, h = item - s * H;
      const int cq = (h * 64) >> 5;
      for (int t = 0; t < tiles; ++t) {
        mbar_wait(&bar_stats[t], iter & 1);
        const float m = fmaxf(sXm[t * 256 + rt], sXm[t * 256 + 128 + rt]);
        const float l = sXl[t * 256 + rt] + sXl[t * 256 + 128 + rt];
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        const int row = t * 128 + rt;
        if (row < N) p.lse[(static_cast<size_t>(s) * H + h) * N + row] = m * c + log2f(l);
        mbar_wait(&bar_o[t], iter & 1);
        tc_fence_after();
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(tm_o + t * 64 + lane_addr, a0);
        tmem_ld_32x32(tm_o + t * 64 + lane_addr + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_ofree[t]);  // the next item may overwrite O and the stats of this tile
        if (t == 0 && tiles > 1) mbar_wait(bar_q1dead, iter & 1);
        if (pending) {
          if (leader) tma_store_wait_read();
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const uint32_t dst = stage + ch * 16384 + rt * 128;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const uint32_t* v = ch ? a1 : a0;
            st_shared_v4(dst + ((j4 ^ (rt & 7)) << 4), round_tf32(__uint_as_float(v[4 * j4]) * inv),
                         round_tf32(__uint_as_float(v[4 * j4 + 1]) * inv),
                         round_tf32(__uint_as_float(v[4 * j4 + 2]) * inv),
                         round_tf32(__uint_as_float(v[4 * j4 + 3]) * inv));
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (leader) {
          tma_store_4d(&tmOut, smem + kQ + 32768, 0, t * 128, cq, s);
          tma_store_commit();
          if (t == tiles - 1 && tiles > 1) {  // Q buffer 1 may be refilled for the next item
            tma_store_wait_read();
            mbar_arrive(bar_q1free);
          }
        }
        pending = true;
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

int gemm_num_sms();
static int g_attn_pf = 1;  // L2 prefetch of a persistent CTA's next loads (2-5 % on the attention kernels)
void attention_set_l2_prefetch(int on) { g_attn_pf = on; }
int attention_l2_prefetch_enabled() { return g_attn_pf; }
static int g_attn_tc = 3;  // bit 0: forward, bit 1: backward on tcgen05 (N <= 256); 0 = the mma.sync kernels
void attention_set_tc(int on) { g_attn_tc = on; }
int attention_tc_enabled() { return g_attn_tc; }

int attention_forward_tc(const float* qkv, float* o, float* lse, const int* lengths, int S, int N, int H,
                         cudaStream_t stream) {
  const int D = H * 64;
  ATST_REQUIRE(N <= 256, "attention_forward_tc: N=%d > 256", N);
  ATST_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0,
               "attention_forward_tc: qkv and o must be 16-byte aligned");
  CUtensorMap tr, ty, tout;
  const long long rows = static_cast<long long>(S) * N;
  int rc;
  if ((rc = make_map_generic_3d(&tr, qkv, rows, 3 * D, 3 * D, 128, 2))) return rc;
  if ((rc = make_map_generic_3d(&ty, qkv, rows, 3 * D, 3 * D, 64, 2))) return rc;
  if ((rc = make_map_seq4d(&tout, o, S, N, D, 128))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd);
    if (e != cudaSuccess) { atst_set_error("attn_fwd_tc smem attr: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  AttnTcParams p{};
  p.lse = lse; p.lengths = lengths; p.N = N; p.H = H; p.D = D; p.scale = 0.125f;
  p.prefetch_dist = attention_l2_prefetch_enabled();
  p.num_items = S * H;
  const int grid = p.num_items < gemm_num_sms() ? p.num_items : gemm_num_sms();  // persistent: one CTA per SM
  attn_fwd_tc_kernel<<<grid, 512, kSmemFwd, stream>>>(tr, ty, tout, p);
  return atst_check_launch("attn_fwd_tc_kernel");
}

}  // namespace atst
