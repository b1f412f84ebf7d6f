// Bring-up probe for tcgen05 operand forms that the attention kernels rely on (tools/bringup.py umma_probe):
//   mode 0  D[128,128] = A[128,64] . B[128,64]^T, both operands TMA-loaded with the 32B-atom 128B swizzle (the
//           token-major 3-D map) but read K-major with a caller-supplied descriptor (layout / LBO / SBO / k step)
//   mode 1  same product, A stored to TMEM by the threads (tcgen05.st) and passed as a TMEM operand, B K-major with
//           the standard 128B swizzle
//   mode 2  D[128,128] = A[128,64] . B[64,128]: A in TMEM, B token-major ("MN-major", 32B-atom swizzle) as V is in P.V
// One CTA of 128 threads; thread r owns row r (TMEM lane r).
#include "common.cuh"

namespace atst {

int make_map_generic_2d(CUtensorMap* map, const float* ptr, long long rows, int cols, int ld, int box_rows);
int make_map_generic_3d(CUtensorMap* map, const float* ptr, long long rows, int feats, int ld, int box_rows,
                        int box_chunks);

struct ProbeParams {
  const float* A;
  float* D;
  int mode;
  uint32_t layout, lbo, sbo, kstep;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ProbeParams p) {
  extern __shared__ uint8_t smem_raw_pb[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_pb) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 32 KB
  uint8_t* sB = smem + 32 * 1024;  // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;

  if (p.mode != 0) {  // A rows -> TMEM columns [128, 192)
    uint32_t v[32];
    for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(p.A[r * 64 + ch * 32 + j]);
      tmem_st_32x32(tmem_base + lane_addr + 128 + ch * 32, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    if (p.mode == 0) {
      mbar_expect_tx(&bars[0], 64 * 1024);
      tma_load_3d(sA, &tmA, &bars[0], 0, 0, 0);
      tma_load_3d(sB, &tmB, &bars[0], 0, 0, 0);
    } else if (p.mode == 1) {
      mbar_expect_tx(&bars[0], 32 * 1024);
      tma_load_2d(sB, &tmB, &bars[0], 0, 0);
      tma_load_2d(sB + 16384, &tmB, &bars[0], 32, 0);
    } else {
      mbar_expect_tx(&bars[0], 32 * 1024);
      tma_load_3d(sB, &tmB, &bars[0], 0, 0, 0);  // [4 chunks][64 rows][128 B]
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    if (p.mode == 0) {
      const uint32_t idesc = make_idesc(2u, 128, 128, 0u, 0u);
      for (int kc = 0; kc < 2; ++kc)
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_base, make_smem_desc(a_addr + kc * 16384 + k * p.kstep, p.lbo, p.sbo, p.layout),
                    make_smem_desc(b_addr + kc * 16384 + k * p.kstep, p.lbo, p.sbo, p.layout), idesc, (kc | k) ? 1u : 0u);
    } else if (p.mode == 1) {
      const uint32_t idesc = make_idesc(2u, 128, 128, 0u, 0u);
      for (int kc = 0; kc < 2; ++kc)
        for (int k = 0; k < 4; ++k)
          umma_tf32_ts(tmem_base, tmem_base + 128 + kc * 32 + k * 8,
                       make_smem_desc(b_addr + kc * 16384 + k * 32, 16, 1024, 2), idesc, (kc | k) ? 1u : 0u);
    } else {
      const uint32_t idesc = make_idesc(2u, 128, 128, 0u, 1u);
      for (int k = 0; k < 8; ++k)
        umma_tf32_ts(tmem_base, tmem_base + 128 + k * 8, make_smem_desc(b_addr + k * 8 * 128, 64 * 128, 512, 1), idesc,
                     k ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + lane_addr + ch * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) p.D[r * 128 + ch * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- DRAM access-pattern probe: copy [rows, cols] fp32 the way the GEMM epilogue touches memory (mode 0: a warp owns
// 32 rows x 64 columns, lane = row, eight 32-byte accesses per lane) or fully coalesced (mode 1: float4 grid-stride)
__global__ void __launch_bounds__(512) copy_pattern_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int rows, int cols, int mode) {
  if (mode == 1) {
    const long long n4 = static_cast<long long>(rows) * cols / 4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode == 2) {
    // same tiles, but a warp instruction covers 8 rows x 128 contiguous bytes (4 lanes per row) instead of 32 rows x 32 B
    const int m_tiles2 = (rows + 127) / 128, n_tiles2 = cols / 256;
    for (int tile = blockIdx.x; tile < m_tiles2 * n_tiles2; tile += gridDim.x) {
      const int m0 = (tile / n_tiles2) * 128, n0 = (tile % n_tiles2) * 256;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {      // 4 x 8 rows of this warp's 32
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {    // 2 x 32 columns of this warp's 64
          const int row = m0 + (warp & 3) * 32 + rr * 8 + (lane >> 2);
          if (row >= rows) continue;
          const size_t off = static_cast<size_t>(row) * cols + n0 + (warp >> 2) * 64 + cc * 32 + (lane & 3) * 8;
          float v[8];
          ld_global_v8(src + off, v);
          st_global_v8(dst + off, v);
        }
      }
    }
    return;
  }
  // tiles of 128 rows x 256 columns, 16 warps per tile: quadrant = warp & 3 (32 rows), slice = warp >> 2 (64 columns)
  const int m_tiles = (rows + 127) / 128, n_tiles = cols / 256;
  for (int tile = blockIdx.x; tile < m_tiles * n_tiles; tile += gridDim.x) {
    const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * 256;
    const int row = m0 + (warp & 3) * 32 + lane;
    if (row >= rows) continue;
    const size_t off = static_cast<size_t>(row) * cols + n0 + (warp >> 2) * 64;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float v[8];
      ld_global_v8(src + off + g * 8, v);
      st_global_v8(dst + off + g * 8, v);
    }
  }
}

int copy_pattern(const float* src, float* dst, int rows, int cols, int mode, cudaStream_t stream) {
  ATST_REQUIRE(cols % 256 == 0, "copy_pattern: cols %% 256 != 0");
  copy_pattern_kernel<<<148 * (mode == 1 ? 4 : 1), 512, 0, stream>>>(src, dst, rows, cols, mode);
  return atst_check_launch("copy_pattern_kernel");
}

int umma_probe(int mode, const float* A, const float* B, float* D, unsigned layout, unsigned lbo, unsigned sbo,
               unsigned kstep, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = make_map_generic_3d(&ta, A, 128, 64, 64, 128, 2);
  if (rc) return rc;
  if (mode == 0) rc = make_map_generic_3d(&tb, B, 128, 64, 64, 128, 2);
  else if (mode == 1) rc = make_map_generic_2d(&tb, B, 128, 64, 64, 128);
  else rc = make_map_generic_3d(&tb, B, 64, 128, 128, 64, 4);
  if (rc) return rc;
  const int smem = 1024 + 64 * 1024 + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { atst_set_error("umma_probe smem attr: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  ProbeParams p{A, D, mode, layout, lbo, sbo, kstep};
  umma_probe_kernel<<<1, 128, smem, stream>>>(ta, tb, p);
  return atst_check_launch("umma_probe_kernel");
}

}  // namespace atst
