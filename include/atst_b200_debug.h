/* Bring-up / profiling entry points.  NOT part of the drop-in boundary (include/atst_b200.h): they are compiled only
 * into libatst_b200_debug.so (python -m audiossl_b200.build --debug, -DATST_DEBUG_ABI) and used by tools/bringup.py. */
#ifndef ATST_B200_DEBUG_H_
#define ATST_B200_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif

/* debug/bring-up variant of atst_gemm_tn / atst_gemm_nn with explicit shared-memory descriptor fields */
int atst_gemm_mn_debug(int nn, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                       unsigned lbo, unsigned sbo, unsigned kstep, unsigned layout, int tma_swizzle, int splits,
                       void* stream);
/* bring-up probe of tcgen05 operand forms (K-major reads of 32B-atom-swizzled tiles, A operand in tensor memory):
 * mode 0/1: D[128,128] = A[128,64] . B[128,64]^T ; mode 2: D[128,128] = A[128,64] . B[64,128] */
int atst_umma_probe(int mode, const float* A, const float* B, float* D, unsigned layout, unsigned lbo, unsigned sbo,
                    unsigned kstep, void* stream);

/* bring-up: clock64() timeline (32 slots, device buffer) of tiles 8-11 of CTA 0 of the CTA-pair GEMM on subsequent
 * GEMM calls: per tile {epilogue warp arrives, bias staged, accumulator complete, tile stored, MMA warp arrives,
 * accumulator stage free, last MMA issued}; buf = NULL switches it off */
int atst_gemm_trace(long long* buf);
/* bring-up: copy [rows, cols] fp32 with the GEMM epilogue's access pattern (mode 0: lane = row, 32-byte accesses) or
 * fully coalesced (mode 1), to measure what each pattern reaches in DRAM bandwidth */
int atst_copy_pattern(const float* src, float* dst, int rows, int cols, int mode, void* stream);
/* bring-up: record a clock64() timeline (80 slots, device buffer) of the CTA of head 0 / sequence seq in the tcgen05
 * backward kernel `mode` (0 dQ, 1 dK dV) on subsequent atst_attention_backward calls; buf = NULL switches it off */
int atst_attention_trace(long long* buf, int seq, int mode);

#ifdef __cplusplus
}
#endif
#endif /* ATST_B200_DEBUG_H_ */
