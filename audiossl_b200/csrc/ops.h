// internal launcher prototypes (namespace atst), one per kernel family
#pragma once
#include <cuda_runtime.h>
namespace atst {
int mel_forward(const float* wav, int B, int n, long long wav_stride, const long long* clip_start, int win_length,
                float* out, long long out_stride, unsigned int* clip_ws, int normalize, cudaStream_t stream);
int layernorm_forward(const float* x, long long x_stride, const float* gamma, const float* beta, float* y,
                      long long y_stride, float* mean, float* rstd, int rows, int D, float eps, int round_out,
                      cudaStream_t st);
int layernorm_backward(const float* dy, long long dy_stride, const float* x, long long x_stride, const float* mean,
                       const float* rstd, const float* gamma, const float* dres, long long dres_stride, float* dx,
                       long long dx_stride, float* dgamma, float* dbeta, int rows, int D, float* dys,
                       long long dys_stride, const float* rowscale, int rows_per_seq, float* colsum_out,
                       cudaStream_t st);
int attention_forward(const float* qkv, float* o, float* lse, const int* lengths, int S, int N, int H, cudaStream_t);
int attention_backward(const float* qkv, const float* o, const float* d_o, const float* lse, float* delta_ws,
                       float* dqkv, const int* lengths, int S, int N, int H, cudaStream_t);
int patchify(const float* mel, long long clip_stride, int S, int T, float* patches, cudaStream_t st);
int tokens_forward(const float* pe, const float* cls, const float* pos, const float* mask_embed,
                   const unsigned char* mask, float* x, int S, int P, int D, int use_cls, cudaStream_t st);
int tokens_backward(const float* dx, const unsigned char* mask, float* dpe, float* dpos, float* dcls,
                    float* dmask_embed, int S, int P, int D, int use_cls, cudaStream_t st);
int colsum_accumulate(const float* X, long long ld, int rows, int cols, float* out, cudaStream_t st);
int bn_stats(const float* X, int rows, int cols, float* mean, float* m2, cudaStream_t st);
int bn_finalize(const float* mean, const float* m2, float count, float eps, float momentum, float* rstd,
                float* running_mean, float* running_var, int cols, cudaStream_t st);
int bn_relu_forward(const float* X, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    float* Y, int rows, int cols, int round_out, cudaStream_t st);
int bn_relu_backward_stats(const float* dY, const float* X, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, int rows, int cols, float* s1, float* s2, cudaStream_t st);
int bn_relu_backward_apply(const float* dY, const float* X, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, const float* s1, const float* s2, float count, float* dX, int rows,
                           int cols, int round_out, cudaStream_t st);
int mixup_forward(const float* x, int x_T, const float* bank, int bank_T, const int* idx, const int* zlen,
                  const int* start, const float* alpha, float* out, int Hm, int B, cudaStream_t st);
int resize_crop_forward(const float* lms, const int* rect, float* out, int B, int Hm, int T, int canvas_h,
                        int canvas_w, cudaStream_t st);
int gather_rows(const float* x, const int* idx, float* out, int rows, int D, cudaStream_t st);
int scatter_rows(const float* src, const int* idx, float* dst, int rows, int D, cudaStream_t st);
int gelu_forward(const float* u, float* g, long long n, cudaStream_t st);
int gelu_backward(float* d, const float* u, int rows, int cols, float* colsum_out, cudaStream_t st);
int round_tf32_copy(const float* src, float* dst, long long n, cudaStream_t st);
int split_tf32(const float* src, long long ld, int rows, int cols, float* dst, int pattern, int along_rows,
               cudaStream_t st);
int axpy(float* y, const float* x, float a, long long n, cudaStream_t st);
int byol_loss(const float* student, const float* teacher, int ncrops, int B, float* dstudent, float* acc_ws,
              cudaStream_t st);
int byol_finalize(const float* acc_ws, float n_student_rows, float n_teacher_rows, int ncrops, int B, float* out3,
                  cudaStream_t st);
int ema_update(float* k, const float* q, float m, const float* m_dev, long long n, cudaStream_t st);
int adamw_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float wd, float b1,
               float b2, float eps, float grad_scale, const float* dyn, cudaStream_t st);
}  // namespace atst
