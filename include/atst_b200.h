/* atst_b200 - C ABI of the B200 (sm_100a) ATST pre-training hot path.
 *
 * The reference (Audio-WestlakeU/audiossl @ ec3a14d) has no FFI layer: its hot path is Python over
 * torch/torchaudio library calls.  These entry points are what a binding for that path
 * needs; each cites the reference code it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error (-1 bad argument, -2 CUDA error, -3 wrong arch);
 *     atst_last_error() returns the message of the last failure on the calling thread.
 *   - all pointers are DEVICE pointers to fp32 data owned by the caller (e.g. tensor.data_ptr()),
 *     16-byte aligned, row-major, dense unless a stride is given.  The library never allocates or frees
 *     device memory and keeps no pointer after a call returns.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously, no host syncs.
 *   - no CPU fallback: on a non-sm_100 device atst_init() fails.
 */
#ifndef ATST_B200_H_
#define ATST_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

int atst_version(void);
const char* atst_last_error(void);
/* checks the current device is sm_100 and warms the driver entry points */
int atst_init(void);
/* tuning switches (bring-up / profiling): "gemm_l2_prefetch" = 0|1, "gemm_cta_pair" = 0|1 (cta_group::2 kernel) */
int atst_set_option(const char* name, int value);

/* ---- mel front-end: torchaudio MelSpectrogram(16000,n_fft=1024,hop=160,win=1024|640,f_min=60,f_max=7800,
 *      n_mels=64) -> AmplitudeToDB("power",top_db=80) -> MinMax(-79.6482,50.6842)
 *      replaces audiossl/methods/atst/transform.py:14-29 (mel_feature), audiossl/transforms/common.py:97-110,
 *      audiossl/methods/atstframe/transform.py:16-42.
 *   wav [B, >= n] (row stride wav_stride) -> out [B, 64, n/160+1] (clip stride out_stride); clip b is the n samples
 *   starting at wav + b*wav_stride + clip_start[b] (clip_start NULL: 0) - the RandomCrop of the train transform
 *   (audiossl/transforms/common.py:63-74) without a copy; reflect padding is relative to the window;
 *   clip_ws: 2*B uint32 scratch; normalize=0 stops after the dB stage (no clamp, no MinMax). */
int atst_mel_forward(const float* wav, int B, int n, long long wav_stride, const long long* clip_start,
                     int win_length, float* out, long long out_stride, unsigned int* clip_ws, int normalize,
                     void* stream);

/* ---- GEMMs (tcgen05, TF32 operands, fp32 accumulate).  Replace nn.Linear forward/backward in
 *      audiossl/modules/transformer.py:86-92,102-119, audiossl/models/atst/audio_transformer.py:60,68,
 *      audiossl/models/atst/byol.py:6-22.
 *   epi: 0 store(+bias) | 1 bias+GELU (aux <- pre-activation) | 2 acc*gelu'(aux) | 3 resid + rowscale*(acc+bias)
 *        | 4 rowscale*acc | 5 relu(acc+bias)
 *        | 10 bias+GELU with aux <- gelu'(pre-activation) stored as fp16 (aux is a __half*, ldaux in halfs)
 *        | 11 acc*aux with that fp16 derivative (10 / 11: N, ldaux multiples of 16; N % 256 == 0 or N > 1024)
 *   rowscale: per-sequence DropPath scale (mask/keep_prob) indexed by row / rows_per_seq, or NULL. */
int atst_gemm_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 const float* bias, int epi, const float* resid, int ldr, float* aux, int ldaux,
                 const float* rowscale, int rows_per_seq, int round_out, void* stream);
/* C[M,N] = epi(A[M,K] . B[K,N]) : input-gradient of a Linear whose weight is B = W[out=K, in=N].
 * colsum_out (nullable, [N]) += column sums of the stored C: the bias gradient of the Linear that produced the
 * forward input of this one (fc1.bias from the GELU' epilogue), taken inside the epilogue. */
int atst_gemm_nn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, int epi,
                 float* aux, int ldaux, const float* rowscale, int rows_per_seq, int round_out, float* colsum_out,
                 void* stream);
/* C[M,N] += A[T,M]^T . B[T,N] : weight-gradient (split over T, atomically accumulated into C) */
int atst_gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int T,
                 void* stream);
/* ---- LayerNorm(eps) forward/backward, row strides in elements (audiossl/modules/transformer.py:128,132;
 *      final norm on the CLS row only: audiossl/models/atst/audio_transformer.py:201,210) */
int atst_layernorm_forward(const float* x, long long x_stride, const float* gamma, const float* beta, float* y,
                           long long y_stride, float* mean, float* rstd, int rows, int D, float eps, int round_out,
                           void* stream);
/* backward: dx = dres + LN'(dy); dgamma/dbeta accumulate.  Optional GEMM-ready copy for the branch that consumes dx:
 * dys = tf32(rowscale[row / rows_per_seq] * dx) (DropPath backward) and colsum_out += column sums of dys
 * (the consumer Linear's bias gradient).  Pass NULLs to skip. */
int atst_layernorm_backward(const float* dy, long long dy_stride, const float* x, long long x_stride,
                            const float* mean, const float* rstd, const float* gamma, const float* dres,
                            long long dres_stride, float* dx, long long dx_stride, float* dgamma, float* dbeta,
                            int rows, int D, float* dys, long long dys_stride, const float* rowscale,
                            int rows_per_seq, float* colsum_out, void* stream);

/* ---- attention core with key padding by length (audiossl/modules/transformer.py:107-121,152-159).
 *   qkv [S*N, 3*H*64] as written by the qkv Linear; o [S*N, H*64]; lse [S,H,N]; lengths int32 [S] or NULL */
int atst_attention_forward(const float* qkv, float* o, float* lse, const int* lengths, int S, int N, int H,
                           void* stream);
int atst_attention_backward(const float* qkv, const float* o, const float* d_o, const float* lse, float* delta_ws,
                            float* dqkv, const int* lengths, int S, int N, int H, void* stream);

/* ---- patch embedding plumbing (audiossl/models/atst/audio_transformer.py:56-75,153-186;
 *      frame model: audiossl/methods/atstframe/audio_transformer.py:161-181) */
int atst_patchify(const float* mel, long long clip_stride, int S, int T, float* patches, void* stream);
int atst_tokens_forward(const float* pe, const float* cls, const float* pos, const float* mask_embed,
                        const unsigned char* mask, float* x, int S, int P, int D, int use_cls, void* stream);
int atst_tokens_backward(const float* dx, const unsigned char* mask, float* dpe, float* dpos, float* dcls,
                         float* dmask_embed, int S, int P, int D, int use_cls, void* stream);
int atst_colsum_accumulate(const float* X, long long ld, int rows, int cols, float* out, void* stream);

/* ---- projector / predictor BatchNorm1d(train) + ReLU (audiossl/models/atst/byol.py:6-22) */
int atst_bn_stats(const float* X, int rows, int cols, float* mean, float* m2, void* stream);
int atst_bn_finalize(const float* mean, const float* m2, float count, float eps, float momentum, float* rstd,
                     float* running_mean, float* running_var, int cols, void* stream);
int atst_bn_relu_forward(const float* X, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         float* Y, int rows, int cols, int round_out, void* stream);
int atst_bn_relu_backward_stats(const float* dY, const float* X, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, int rows, int cols, float* s1, float* s2,
                                void* stream);
int atst_bn_relu_backward_apply(const float* dY, const float* X, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, const float* s1, const float* s2, float count,
                                float* dX, int rows, int cols, int round_out, void* stream);

/* ---- BYOL loss + compute_var statistics (audiossl/models/atst/byol.py:24-78).
 *   student [ncrops*B,256], teacher [2*B,256]; dstudent = d loss / d student;
 *   acc_ws [1+4*256]: raw sums (all-reduce these across ranks before finalize); out3 = loss, std_s, std_t */
int atst_byol_loss(const float* student, const float* teacher, int ncrops, int B, float* dstudent, float* acc_ws,
                   void* stream);
int atst_byol_finalize(const float* acc_ws, float n_student_rows, float n_teacher_rows, int ncrops, int B,
                       float* out3, void* stream);

/* ---- teacher EMA (audiossl/models/atst/atst.py:29-34) and HF-semantics AdamW
 *      (audiossl/methods/atst/model.py:44-48; transformers 4.x AdamW, eps inside, decay after update)
 *      m_dev / dyn (nullable device pointers): the per-step scalars - EMA momentum; {lr*sqrt(1-b2^t)/(1-b1^t), lr*wd} -
 *      read from device memory instead of the arguments, so a captured CUDA graph of the step replays with new values */
int atst_ema_update(float* k, const float* q, float m, const float* m_dev, long long n, void* stream);
int atst_adamw_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float wd,
                    float beta1, float beta2, float eps, float grad_scale, const float* dyn, void* stream);

/* ---- device-batched augmentations between mel and encoder (audiossl/transforms/byol_a.py:7-49,61-115):
 *   mixup: out[b] = log((1-alpha[b]) e^x[b] + alpha[b] e^bank[idx[b]] + eps), idx[b] < 0 copies x[b]; x [B,Hm,x_T],
 *          bank entries [Hm,bank_T] with zlen[b] valid frames (NULL: x_T) - a longer entry is read from frame
 *          start[b], a shorter one is mixed into frames [start[b], start[b]+zlen[b]) of x (log_mixup_exp's branches)
 *   resize_crop: crop rect[b] = (i, j, h, w) of the zero canvas [canvas_h, canvas_w] holding lms[b] centred,
 *                bicubic (align_corners=True, A=-0.75) resize back to [Hm, T] */
int atst_mixup_forward(const float* x, int x_T, const float* bank, int bank_T, const int* idx, const int* zlen,
                       const int* start, const float* alpha, float* out, int Hm, int B, void* stream);
int atst_resize_crop_forward(const float* lms, const int* rect, float* out, int B, int Hm, int T, int canvas_h,
                             int canvas_w, void* stream);

/* ---- ATST-Frame row selection (audiossl/methods/atstframe/audio_transformer.py:187-207: frame_repr[mask & valid]):
 *      out[r] = x[idx[r]] and its adjoint dst[idx[r]] = src[r] (dst pre-zeroed, unique indices) */
int atst_gather_rows(const float* x, const int* idx, float* out, int rows, int D, void* stream);
int atst_scatter_rows(const float* src, const int* idx, float* dst, int rows, int D, void* stream);

/* ---- exact-erf GELU as separate passes (audiossl/modules/transformer.py:78,88): g = gelu(u); d *= gelu'(u) */
int atst_gelu_forward(const float* u, float* g, long long n, void* stream);
/* d[rows, cols] *= gelu'(u) in place (tf32-rounded); colsum_out (nullable) += column sums of the result */
int atst_gelu_backward(float* d, const float* u, int rows, int cols, float* colsum_out, void* stream);

/* ---- misc */
/* producer rounding of a GEMM operand (cvt.rna.tf32); the 3xTF32 validation build copies instead */
int atst_round_tf32(const float* src, float* dst, long long n, void* stream);
/* 1 if this library is the 3xTF32 validation build (libatst_b200_precise.so, -DATST_PRECISE), else 0 */
int atst_is_precise(void);
/* error-compensated operand split for 3xTF32 products: x = hi + lo, hi = tf32(x), lo = tf32(x - hi).
 * src [rows, cols] (row stride ld) -> dst, three blocks side by side ([rows, 3*cols], along_rows = 0) or stacked
 * ([3*rows, cols], along_rows = 1) holding hi|lo|hi (pattern 0) or hi|hi|lo (pattern 1): contracting a pattern-0
 * operand with a pattern-1 operand over the tripled dimension gives hi*hi + lo*hi + hi*lo */
int atst_split_tf32(const float* src, long long ld, int rows, int cols, float* dst, int pattern, int along_rows,
                    void* stream);
int atst_axpy(float* y, const float* x, float a, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATST_B200_H_ */
