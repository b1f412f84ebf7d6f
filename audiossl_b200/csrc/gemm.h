// Internal GEMM launcher interface shared by the C-ABI layer and the encoder driver.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace atst {

enum GemmEpilogue : int {
  EPI_STORE = 0,   // C = acc (+ bias)
  EPI_GELU = 1,    // aux = acc + bias ; C = gelu_erf(aux)
  EPI_DGELU = 2,   // C = acc * gelu'(aux)
  EPI_RESID = 3,   // C = resid + rowscale[row / rows_per_seq] * (acc + bias)
  EPI_SCALE = 4,   // C = rowscale[row / rows_per_seq] * acc
  EPI_RELU = 5,    // C = max(acc + bias, 0)
  EPI_ATOMIC = 6,  // C += acc   (split-K partial sums, red.global.add)
  EPI_GELU_H = 10,   // aux (fp16, ldaux in halfs) = gelu'(acc + bias) ; C = gelu_erf(acc + bias)   (pair kernel only)
  EPI_DGELU_H = 11,  // C = acc * aux (fp16: the derivative stored by EPI_GELU_H)                   (pair kernel only)
  EPI_DBG_NOSTORE = 8,  // profiling aid: full epilogue without the global stores
  EPI_DBG_NOLOAD = 9,   // profiling aid: epilogue without the TMEM loads (stores zeros)
};

struct GemmParams {
  int M = 0, N = 0, K = 0;
  float* C = nullptr;
  int ldc = 0;
  const float* bias = nullptr;      // [N] or null
  const float* resid = nullptr;     // [M, ldr]
  int ldr = 0;
  float* aux = nullptr;             // [M, ldaux] pre-activation (written by EPI_GELU, read by EPI_DGELU); the _H
                                    // epilogues keep gelu'(pre-activation) there instead, as fp16 (a __half*)
  int ldaux = 0;
  const float* rowscale = nullptr;  // per-sequence DropPath scale (mask / keep_prob) or null
  int rows_per_seq = 1;
  int epi = EPI_STORE;
  int round_out = 0;                // round C to tf32 (it feeds another GEMM)
  long long* trace = nullptr;       // bring-up: clock64() timeline of one epilogue warp / the MMA warp of CTA 0
  float* colsum = nullptr;          // [N] or null: colsum[n] += sum over rows of the stored C[:, n] (a bias gradient)
  int splits = 0;                   // TN only: split-K factor (0 = auto)
  int l2_prefetch = 1;              // producer prefetches streaming operand tiles into L2 ahead of the smem ring
  // TN (token-major operands) shared-memory descriptor fields; 0 = defaults
  uint32_t mn_lbo = 0, mn_sbo = 0, mn_kstep = 0, mn_layout = 0;
  int mn_tma_swizzle = 0;
};

// C[M,N] = epi(A[M,K] . B[N,K]^T); A, B row-major with leading dims lda, ldb (elements)
int gemm_nt(const float* A, int lda, const float* B, int ldb, GemmParams p, cudaStream_t stream);
// C[M,N] = epi(A[M,K] . B[K,N]); B row-major [K,N] (dgrad against a Linear weight [out=K, in=N])
int gemm_nn(const float* A, int lda, const float* B, int ldb, GemmParams p, cudaStream_t stream);
// C[M,N] += A[T,M]^T . B[T,N]; contraction over the T rows (tokens); p.K ignored
int gemm_tn(const float* A, int lda, const float* B, int ldb, int T, GemmParams p, cudaStream_t stream);

void gemm_set_l2_prefetch(int on);
void gemm_set_cta_pair(int on);

}  // namespace atst
