// extern "C" surface declared in include/atst_b200.h: thin forwarding to the launchers.
#include "../../include/atst_b200.h"
#ifdef ATST_DEBUG_ABI
#include "../../include/atst_b200_debug.h"
#endif
#include "common.cuh"
#include "gemm.h"
#include "ops.h"
#include <string.h>

namespace atst {
void attention_set_tc(int on);
void attention_set_l2_prefetch(int on);
void attention_set_trace(long long* buf, int seq, int mode);
void gemm_set_trace(long long* buf);
#ifdef ATST_DEBUG_ABI
int copy_pattern(const float* src, float* dst, int rows, int cols, int mode, cudaStream_t stream);
int umma_probe(int mode, const float* A, const float* B, float* D, unsigned layout, unsigned lbo, unsigned sbo,
               unsigned kstep, cudaStream_t stream);
#endif
}
#define ST(s) reinterpret_cast<cudaStream_t>(s)
using namespace atst;

extern "C" {

int atst_version(void) { return 100; }

int atst_set_option(const char* name, int value) {
  if (name != nullptr && strcmp(name, "gemm_l2_prefetch") == 0) { gemm_set_l2_prefetch(value); return ATST_OK; }
  if (name != nullptr && strcmp(name, "gemm_cta_pair") == 0) { gemm_set_cta_pair(value); return ATST_OK; }
  if (name != nullptr && strcmp(name, "attn_tcgen05") == 0) { attention_set_tc(value); return ATST_OK; }
  if (name != nullptr && strcmp(name, "attn_l2_prefetch") == 0) { attention_set_l2_prefetch(value); return ATST_OK; }
  atst_set_error("atst_set_option: unknown option '%s'", name ? name : "(null)");
  return ATST_ERR_ARG;
}

int atst_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { atst_set_error("atst_init: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { atst_set_error("atst_init: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
  if (prop.major != 10) {
    atst_set_error("atst_init: device %s is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.name,
                   prop.major, prop.minor);
    return ATST_ERR_ARCH;
  }
  return ATST_OK;
}

int atst_mel_forward(const float* wav, int B, int n, long long wav_stride, const long long* clip_start,
                     int win_length, float* out, long long out_stride, unsigned int* clip_ws, int normalize,
                     void* stream) {
  return mel_forward(wav, B, n, wav_stride, clip_start, win_length, out, out_stride, clip_ws, normalize, ST(stream));
}

int atst_gemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 const float* bias, int epi, const float* resid, int ldr, float* aux, int ldaux,
                 const float* rowscale, int rows_per_seq, int round_out, void* stream) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias; p.epi = epi; p.resid = resid; p.ldr = ldr;
  p.aux = aux; p.ldaux = ldaux; p.rowscale = rowscale; p.rows_per_seq = rows_per_seq > 0 ? rows_per_seq : 1;
  p.round_out = round_out;
  ATST_REQUIRE((epi >= EPI_STORE && epi <= EPI_RELU) || epi == EPI_GELU_H || epi == EPI_DBG_NOSTORE ||
                   epi == EPI_DBG_NOLOAD,
               "atst_gemm_nt: bad epilogue %d", epi);
  ATST_REQUIRE(!(epi == EPI_RESID && resid == nullptr), "atst_gemm_nt: EPI_RESID needs resid");
  ATST_REQUIRE(!(epi == EPI_DGELU && aux == nullptr), "atst_gemm_nt: EPI_DGELU needs aux");
  return gemm_nt(A, lda, B, ldb, p, ST(stream));
}

int atst_gemm_nn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, int epi,
                 float* aux, int ldaux, const float* rowscale, int rows_per_seq, int round_out, float* colsum_out,
                 void* stream) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.epi = epi; p.aux = aux; p.ldaux = ldaux;
  p.rowscale = rowscale; p.rows_per_seq = rows_per_seq > 0 ? rows_per_seq : 1; p.round_out = round_out;
  p.colsum = colsum_out;
  ATST_REQUIRE(epi == EPI_STORE || epi == EPI_DGELU || epi == EPI_DGELU_H || epi == EPI_SCALE,
               "atst_gemm_nn: bad epilogue %d", epi);
  ATST_REQUIRE(!((epi == EPI_DGELU || epi == EPI_DGELU_H) && aux == nullptr), "atst_gemm_nn: EPI_DGELU needs aux");
  return gemm_nn(A, lda, B, ldb, p, ST(stream));
}

int atst_gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int T,
                 void* stream) {
  GemmParams p;
  p.M = M; p.N = N; p.C = C; p.ldc = ldc;
  return gemm_tn(A, lda, B, ldb, T, p, ST(stream));
}

#ifdef ATST_DEBUG_ABI  // bring-up entry points: libatst_b200_debug.so only (include/atst_b200_debug.h)
int atst_gemm_mn_debug(int nn, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                       unsigned lbo, unsigned sbo, unsigned kstep, unsigned layout, int tma_swizzle, int splits,
                       void* stream) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
  p.mn_lbo = lbo; p.mn_sbo = sbo; p.mn_kstep = kstep; p.mn_layout = layout; p.mn_tma_swizzle = tma_swizzle;
  p.splits = splits;
  if (nn) return gemm_nn(A, lda, B, ldb, p, ST(stream));
  return gemm_tn(A, lda, B, ldb, K, p, ST(stream));
}

int atst_umma_probe(int mode, const float* A, const float* B, float* D, unsigned layout, unsigned lbo, unsigned sbo,
                    unsigned kstep, void* stream) {
  return umma_probe(mode, A, B, D, layout, lbo, sbo, kstep, ST(stream));
}

int atst_gemm_trace(long long* buf) {
  gemm_set_trace(buf);
  return ATST_OK;
}
int atst_copy_pattern(const float* src, float* dst, int rows, int cols, int mode, void* stream) {
  return copy_pattern(src, dst, rows, cols, mode, ST(stream));
}
int atst_attention_trace(long long* buf, int seq, int mode) {
  attention_set_trace(buf, seq, mode);
  return ATST_OK;
}
#endif

int atst_layernorm_forward(const float* x, long long x_stride, const float* gamma, const float* beta, float* y,
                           long long y_stride, float* mean, float* rstd, int rows, int D, float eps, int round_out,
                           void* stream) {
  return layernorm_forward(x, x_stride, gamma, beta, y, y_stride, mean, rstd, rows, D, eps, round_out, ST(stream));
}
int atst_layernorm_backward(const float* dy, long long dy_stride, const float* x, long long x_stride,
                            const float* mean, const float* rstd, const float* gamma, const float* dres,
                            long long dres_stride, float* dx, long long dx_stride, float* dgamma, float* dbeta,
                            int rows, int D, float* dys, long long dys_stride, const float* rowscale,
                            int rows_per_seq, float* colsum_out, void* stream) {
  return layernorm_backward(dy, dy_stride, x, x_stride, mean, rstd, gamma, dres, dres_stride, dx, dx_stride, dgamma,
                            dbeta, rows, D, dys, dys_stride, rowscale, rows_per_seq, colsum_out, ST(stream));
}
int atst_attention_forward(const float* qkv, float* o, float* lse, const int* lengths, int S, int N, int H,
                           void* stream) {
  return attention_forward(qkv, o, lse, lengths, S, N, H, ST(stream));
}
int atst_attention_backward(const float* qkv, const float* o, const float* d_o, const float* lse, float* delta_ws,
                            float* dqkv, const int* lengths, int S, int N, int H, void* stream) {
  return attention_backward(qkv, o, d_o, lse, delta_ws, dqkv, lengths, S, N, H, ST(stream));
}
int atst_patchify(const float* mel, long long clip_stride, int S, int T, float* patches, void* stream) {
  return patchify(mel, clip_stride, S, T, patches, ST(stream));
}
int atst_tokens_forward(const float* pe, const float* cls, const float* pos, const float* mask_embed,
                        const unsigned char* mask, float* x, int S, int P, int D, int use_cls, void* stream) {
  return tokens_forward(pe, cls, pos, mask_embed, mask, x, S, P, D, use_cls, ST(stream));
}
int atst_tokens_backward(const float* dx, const unsigned char* mask, float* dpe, float* dpos, float* dcls,
                         float* dmask_embed, int S, int P, int D, int use_cls, void* stream) {
  return tokens_backward(dx, mask, dpe, dpos, dcls, dmask_embed, S, P, D, use_cls, ST(stream));
}
int atst_colsum_accumulate(const float* X, long long ld, int rows, int cols, float* out, void* stream) {
  return colsum_accumulate(X, ld, rows, cols, out, ST(stream));
}
int atst_bn_stats(const float* X, int rows, int cols, float* mean, float* m2, void* stream) {
  return bn_stats(X, rows, cols, mean, m2, ST(stream));
}
int atst_bn_finalize(const float* mean, const float* m2, float count, float eps, float momentum, float* rstd,
                     float* running_mean, float* running_var, int cols, void* stream) {
  return bn_finalize(mean, m2, count, eps, momentum, rstd, running_mean, running_var, cols, ST(stream));
}
int atst_bn_relu_forward(const float* X, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         float* Y, int rows, int cols, int round_out, void* stream) {
  return bn_relu_forward(X, mean, rstd, gamma, beta, Y, rows, cols, round_out, ST(stream));
}
int atst_bn_relu_backward_stats(const float* dY, const float* X, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, int rows, int cols, float* s1, float* s2,
                                void* stream) {
  return bn_relu_backward_stats(dY, X, mean, rstd, gamma, beta, rows, cols, s1, s2, ST(stream));
}
int atst_bn_relu_backward_apply(const float* dY, const float* X, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, const float* s1, const float* s2, float count,
                                float* dX, int rows, int cols, int round_out, void* stream) {
  return bn_relu_backward_apply(dY, X, mean, rstd, gamma, beta, s1, s2, count, dX, rows, cols, round_out, ST(stream));
}
int atst_byol_loss(const float* student, const float* teacher, int ncrops, int B, float* dstudent, float* acc_ws,
                   void* stream) {
  return byol_loss(student, teacher, ncrops, B, dstudent, acc_ws, ST(stream));
}
int atst_byol_finalize(const float* acc_ws, float n_student_rows, float n_teacher_rows, int ncrops, int B,
                       float* out3, void* stream) {
  return byol_finalize(acc_ws, n_student_rows, n_teacher_rows, ncrops, B, out3, ST(stream));
}
int atst_ema_update(float* k, const float* q, float m, const float* m_dev, long long n, void* stream) {
  return ema_update(k, q, m, m_dev, n, ST(stream));
}
int atst_adamw_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float wd,
                    float beta1, float beta2, float eps, float grad_scale, const float* dyn, void* stream) {
  return adamw_step(p, g, m, v, n, step, lr, wd, beta1, beta2, eps, grad_scale, dyn, ST(stream));
}
int atst_mixup_forward(const float* x, int x_T, const float* bank, int bank_T, const int* idx, const int* zlen,
                       const int* start, const float* alpha, float* out, int Hm, int B, void* stream) {
  return mixup_forward(x, x_T, bank, bank_T, idx, zlen, start, alpha, out, Hm, B, ST(stream));
}
int atst_resize_crop_forward(const float* lms, const int* rect, float* out, int B, int Hm, int T, int canvas_h,
                             int canvas_w, void* stream) {
  return resize_crop_forward(lms, rect, out, B, Hm, T, canvas_h, canvas_w, ST(stream));
}
int atst_gather_rows(const float* x, const int* idx, float* out, int rows, int D, void* stream) {
  return gather_rows(x, idx, out, rows, D, ST(stream));
}
int atst_scatter_rows(const float* src, const int* idx, float* dst, int rows, int D, void* stream) {
  return scatter_rows(src, idx, dst, rows, D, ST(stream));
}
int atst_gelu_forward(const float* u, float* g, long long n, void* stream) { return gelu_forward(u, g, n, ST(stream)); }
int atst_gelu_backward(float* d, const float* u, int rows, int cols, float* colsum_out, void* stream) {
  return gelu_backward(d, u, rows, cols, colsum_out, ST(stream));
}
int atst_round_tf32(const float* src, float* dst, long long n, void* stream) {
  return round_tf32_copy(src, dst, n, ST(stream));
}
int atst_is_precise(void) {
#ifdef ATST_PRECISE
  return 1;
#else
  return 0;
#endif
}
int atst_split_tf32(const float* src, long long ld, int rows, int cols, float* dst, int pattern, int along_rows,
                    void* stream) {
  return split_tf32(src, ld, rows, cols, dst, pattern, along_rows, ST(stream));
}
int atst_axpy(float* y, const float* x, float a, long long n, void* stream) { return axpy(y, x, a, n, ST(stream)); }

}  // extern "C"
