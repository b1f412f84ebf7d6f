// Attention forward on tcgen05 (5th-gen tensor cores) for sequences of up to 256 tokens (every ATST config:
// N <= 251).  Same math and outputs as attn_kernel<0> in attention.cu: o = softmax(q k^T / 8 + key-padding) v,
// lse in the log2 domain; TF32 operands, fp32 accumulation in TMEM, fp32 softmax.
//
// One CTA = one (sequence, head, 128-query tile); 224 KB of shared memory, all operands resident:
//   Q  [2 k-chunks][128 rows][128 B]   K-major, 128B swizzle   (TMA, box {32 floats, 128 rows})
//   K  [2 k-chunks][256 rows][128 B]   K-major, 128B swizzle
//   V  [2 key halves][2 dh-chunks][128 keys][128 B]  token-major ("MN-major" B operand), 128B swizzle / 32B atoms
//   P  [2 buffers][2 key-chunks][128 rows][128 B]    K-major, written by the softmax threads with the swizzle applied
// TMEM: S = Q K^T in columns [0,256) (one UMMA N=256), O in columns [256,320).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 softmax + epilogue
// (thread = query row = TMEM lane, so row max / row sum need no shuffles).  The row maximum is taken over the whole
// row first (S stays in TMEM and is read twice), so P needs no rescaling: the four 64-key quarters of P stream
// through two smem buffers while the previous quarter's P V MMAs run.
#include "common.cuh"

namespace atst {

namespace {

constexpr int kQOff = 0;
constexpr int kKOff = 32 * 1024;
constexpr int kVOff = 96 * 1024;
constexpr int kPOff = 160 * 1024;
constexpr int kBarOff = 224 * 1024;
constexpr int kSmemTc = 1024 + kBarOff + 256;

struct AttnTcParams {
  float* o;
  float* lse;
  const int* lengths;
  int N, H, D;
  float scale;
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(256, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV, AttnTcParams p) {
  extern __shared__ uint8_t smem_raw_tc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;      // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;      // [2]
  uint64_t* bar_pfree = bars + 6;  // [2]
  uint64_t* bar_o = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, s = blockIdx.z;
  const int N = p.N, D = p.D;
  int len = p.lengths ? p.lengths[s] : N;
  if (len <= 0 || len > N) len = N;  // see attention.cu
  const int nq = (len + 63) >> 6;  // 64-key quarters that contain valid keys
  const int row0 = s * N;          // first token row of this sequence in the [S*N, 3D] tensor

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(&bar_v[0], 1);
    mbar_init(&bar_v[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(&bar_p[0], 128);
    mbar_init(&bar_p[1], 128);
    mbar_init(&bar_pfree[0], 1);
    mbar_init(&bar_pfree[1], 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      // Q (2 boxes) + K (4 boxes of 128 rows) on one barrier, the two key halves of V on their own
      mbar_expect_tx(bar_qk, 96 * 1024);
      for (int kc = 0; kc < 2; ++kc) {
        tma_load_2d(smem + kQOff + kc * 16384, &tmQK, bar_qk, h * 64 + kc * 32, row0 + q0);
        tma_load_2d(smem + kKOff + kc * 32768, &tmQK, bar_qk, D + h * 64 + kc * 32, row0);
        tma_load_2d(smem + kKOff + kc * 32768 + 16384, &tmQK, bar_qk, D + h * 64 + kc * 32, row0 + 128);
      }
      for (int hf = 0; hf < 2; ++hf) {
        if (hf * 2 >= nq) break;
        mbar_expect_tx(&bar_v[hf], 32 * 1024);
        tma_load_3d(smem + kVOff + hf * 32768, &tmV, &bar_v[hf], 0, row0 + hf * 128, (2 * D + h * 64) / 32);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = make_idesc(2u, 128, 256, 0u, 0u);
    const uint32_t idesc_pv = make_idesc(2u, 128, 64, 0u, 1u);
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t qa = smem_u32(smem + kQOff), ka = smem_u32(smem + kKOff);
#pragma unroll
      for (int kc = 0; kc < 2; ++kc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_s, make_smem_desc(qa + kc * 16384 + k * 32, 16, 1024, 2),
                    make_smem_desc(ka + kc * 32768 + k * 32, 16, 1024, 2), idesc_s, (kc | k) ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    for (int q = 0; q < nq; ++q) {
      const int b = q & 1, hf = q >> 1;
      if ((q & 1) == 0) {
        mbar_wait(&bar_v[hf], 0);
      }
      mbar_wait(&bar_p[b], (q >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pa = smem_u32(smem + kPOff + b * 32768);
        const uint32_t va = smem_u32(smem + kVOff + hf * 32768) + (q & 1) * 64 * 128;  // 64 key rows into the half
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_o, make_smem_desc(pa + kc * 16384 + k * 32, 16, 1024, 2),
                      make_smem_desc(va + (kc * 32 + k * 8) * 128, 16384, 512, 1), idesc_pv,
                      (q | kc | k) ? 1u : 0u);
        umma_commit(&bar_pfree[b]);
        if (q == nq - 1) umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax + epilogue (thread = query row)
    const int ew = warp - 4;
    const int r = ew * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(ew * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // pass 1: row maximum over the valid keys
    float m = -INFINITY;
    const int nchunks = (len + 31) >> 5;
    for (int ch = 0; ch < nchunks; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_s + lane_addr + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (ch * 32 + j < len) m = fmaxf(m, __uint_as_float(v[j]));
    }
    // pass 2: probabilities, quarter by quarter through the two P buffers
    float l = 0.f;
    const uint32_t p_base = smem_u32(smem + kPOff);
    for (int q = 0; q < nq; ++q) {
      const int b = q & 1;
      if (q >= 2) {
        mbar_wait(&bar_pfree[b], 0);  // the P V MMAs of quarter q-2 have consumed this buffer
      }
#pragma unroll
      for (int c32 = 0; c32 < 2; ++c32) {
        uint32_t v[32];
        const int col0 = q * 64 + c32 * 32;
        tmem_ld_32x32(tmem_s + lane_addr + col0, v);
        tmem_ld_wait();
        const uint32_t dst = p_base + b * 32768 + c32 * 16384 + r * 128;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float e[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = col0 + j4 * 4 + i;
            const float pv = col < len ? exp2f((__uint_as_float(v[j4 * 4 + i]) - m) * c) : 0.f;
            l += pv;
            e[i] = round_tf32(pv);
          }
          st_shared_v4(dst + ((j4 ^ (r & 7)) << 4), e[0], e[1], e[2], e[3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&bar_p[b]);
    }
    // epilogue: O / l -> global, lse
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int qrow = q0 + r;
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    float* op = p.o + (static_cast<size_t>(row0) + qrow) * D + h * 64;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_o + lane_addr + ch * 32, v);
      tmem_ld_wait();
      if (qrow < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = round_tf32(__uint_as_float(v[8 * j + e]) * inv);
          st_global_v8(op + ch * 32 + 8 * j, o8);
        }
      }
    }
    if (qrow < N) p.lse[(static_cast<size_t>(s) * p.H + h) * N + qrow] = m * c + log2f(l);
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
int make_map_generic_2d(CUtensorMap* map, const float* ptr, long long rows, int cols, int ld, int box_rows);
int make_map_generic_3d(CUtensorMap* map, const float* ptr, long long rows, int feats, int ld, int box_rows,
                        int box_chunks);

static int g_attn_tc = 0;
void attention_set_tc(int on) { g_attn_tc = on; }
int attention_tc_enabled() { return g_attn_tc; }

int attention_forward_tc(const float* qkv, float* o, float* lse, const int* lengths, int S, int N, int H,
                         cudaStream_t stream) {
  const int D = H * 64;
  ATST_REQUIRE(N <= 256, "attention_forward_tc: N=%d > 256", N);
  ATST_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 31) == 0,
               "attention_forward_tc: qkv must be 16-byte and o 32-byte aligned");
  CUtensorMap tq, tv;
  int rc = make_map_generic_2d(&tq, qkv, static_cast<long long>(S) * N, 3 * D, 3 * D, 128);
  if (rc) return rc;
  rc = make_map_generic_3d(&tv, qkv, static_cast<long long>(S) * N, 3 * D, 3 * D, 128, 2);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc);
    if (e != cudaSuccess) { atst_set_error("attn_fwd_tc smem attr: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  AttnTcParams p{};
  p.o = o; p.lse = lse; p.lengths = lengths; p.N = N; p.H = H; p.D = D; p.scale = 0.125f;
  dim3 grid((N + 127) / 128, H, S);
  attn_fwd_tc_kernel<<<grid, 256, kSmemTc, stream>>>(tq, tv, p);
  return atst_check_launch("attn_fwd_tc_kernel");
}

}  // namespace atst
