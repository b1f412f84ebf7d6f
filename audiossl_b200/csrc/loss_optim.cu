// BYOL loss + logged std statistics (K15; models/atst/byol.py:24-78), multi-tensor EMA (K17; atst.py:29-34),
// HF-semantics AdamW over flat buffers (K18; methods/atst/model.py:44-48).
#include "common.cuh"

namespace atst {

constexpr int kOut = 256;  // projector / predictor output width (byol.py:97-101)

// One warp per student row (crop iv, sample b).  For every teacher chunk iq != iv:
//   term = 2 - 2 * <t^, s^> ;  loss = mean over rows and pairs.
// d loss / d s = -(2 / (n_terms * B)) * (T - s^ <s^, T>) / max(|s|, eps),  T = sum_{iq != iv} t^_{iq,b}
// Also accumulates sum / sum-of-squares per output dim of the normalised rows for compute_var (logged only).
__global__ void __launch_bounds__(256)
byol_loss_kernel(const float* __restrict__ student, const float* __restrict__ teacher, int ncrops, int B,
                 float* __restrict__ dstudent, float* __restrict__ loss_acc /* [1] sum of <t^,s^> */,
                 float* __restrict__ stats /* [4][256]: s_sum, s_sq, t_sum, t_sq */, int symmetric_pairs) {
  __shared__ float sstat[4][kOut];
  for (int i = threadIdx.x; i < 4 * kOut; i += blockDim.x) (&sstat[0][0])[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int rows = ncrops * B;
  const float eps = 1e-12f;
  float dot_acc = 0.f;
  const int n_terms = 2 * ncrops - 2;
  const float coef = -2.0f / (static_cast<float>(n_terms) * B);
  for (int row = blockIdx.x * wpc + warp; row < rows; row += gridDim.x * wpc) {
    const int iv = row / B, b = row - iv * B;
    float s[8];
    const float4* sp = reinterpret_cast<const float4*>(student + static_cast<size_t>(row) * kOut);
    float4 a = sp[lane], c = sp[lane + 32];
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = c.x; s[5] = c.y; s[6] = c.z; s[7] = c.w;
    float nn = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) nn += s[j] * s[j];
    const float inv_s = 1.0f / fmaxf(sqrtf(warp_sum(nn)), eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] *= inv_s;  // s^
    float T[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int iq = 0; iq < 2; ++iq) {
      const float4* tp = reinterpret_cast<const float4*>(teacher + static_cast<size_t>(iq * B + b) * kOut);
      float tv[8];
      a = tp[lane]; c = tp[lane + 32];
      tv[0] = a.x; tv[1] = a.y; tv[2] = a.z; tv[3] = a.w; tv[4] = c.x; tv[5] = c.y; tv[6] = c.z; tv[7] = c.w;
      float tn = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) tn += tv[j] * tv[j];
      const float inv_t = 1.0f / fmaxf(sqrtf(warp_sum(tn)), eps);
      if (iv == iq) {  // each teacher row is visited exactly once through its own-index student row: teacher stats
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = (j < 4) ? lane * 4 + j : 128 + lane * 4 + (j - 4);
          const float y = tv[j] * inv_t;
          atomicAdd(&sstat[2][col], y);
          atomicAdd(&sstat[3][col], y * y);
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) T[j] += tv[j] * inv_t;
    }
    float st = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) st += s[j] * T[j];
    st = warp_sum(st);  // <s^, T> = sum over pairs of <t^, s^>
    dot_acc += st;
    float4 g0, g1;
    g0.x = coef * (T[0] - s[0] * st) * inv_s; g0.y = coef * (T[1] - s[1] * st) * inv_s;
    g0.z = coef * (T[2] - s[2] * st) * inv_s; g0.w = coef * (T[3] - s[3] * st) * inv_s;
    g1.x = coef * (T[4] - s[4] * st) * inv_s; g1.y = coef * (T[5] - s[5] * st) * inv_s;
    g1.z = coef * (T[6] - s[6] * st) * inv_s; g1.w = coef * (T[7] - s[7] * st) * inv_s;
    float4* gp = reinterpret_cast<float4*>(dstudent + static_cast<size_t>(row) * kOut);
    gp[lane] = g0;
    gp[lane + 32] = g1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = (j < 4) ? lane * 4 + j : 128 + lane * 4 + (j - 4);
      atomicAdd(&sstat[0][col], s[j]);
      atomicAdd(&sstat[1][col], s[j] * s[j]);
    }
  }
  if (lane == 0 && dot_acc != 0.f) atomicAdd(loss_acc, dot_acc);
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * kOut; i += blockDim.x) {
    const float v = (&sstat[0][0])[i];
    if (v != 0.f) atomicAdd(stats + i, v);
  }
}

// loss = 2 - 2 * dot_sum / (n_terms * B); std = mean_d sqrt(ss/(n-1) - s^2/(n(n-1)) + 1e-6)
__global__ void byol_finalize_kernel(const float* __restrict__ loss_acc, const float* __restrict__ stats, float n_s,
                                     float n_t, float pairs_times_b, float* __restrict__ out /* loss,std_s,std_t */) {
  __shared__ float red[2][8];
  const int d = threadIdx.x;  // 256 threads
  const float ss = stats[d], sq = stats[kOut + d], ts = stats[2 * kOut + d], tq = stats[3 * kOut + d];
  float vs = sqrtf(sq / (n_s - 1.f) - ss * ss / (n_s * (n_s - 1.f)) + 1e-6f);
  float vt = sqrtf(tq / (n_t - 1.f) - ts * ts / (n_t * (n_t - 1.f)) + 1e-6f);
  vs = warp_sum(vs);
  vt = warp_sum(vt);
  if ((d & 31) == 0) { red[0][d >> 5] = vs; red[1][d >> 5] = vt; }
  __syncthreads();
  if (d == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    out[0] = 2.0f - 2.0f * loss_acc[0] / pairs_times_b;
    out[1] = a / kOut;
    out[2] = b / kOut;
  }
}

// k = m*k + (1-m)*q over a flat buffer
// (m_dev: the momentum read from device memory, so that a CUDA graph of the step can be replayed with a new value)
__global__ void ema_kernel(float* __restrict__ k, const float* __restrict__ q, float m, const float* __restrict__ m_dev,
                           long long n4) {
  if (m_dev != nullptr) m = *m_dev;
  const float om = 1.0f - m;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 kv = reinterpret_cast<float4*>(k)[i];
    const float4 qv = reinterpret_cast<const float4*>(q)[i];
    kv.x = kv.x * m + om * qv.x; kv.y = kv.y * m + om * qv.y; kv.z = kv.z * m + om * qv.z; kv.w = kv.w * m + om * qv.w;
    reinterpret_cast<float4*>(k)[i] = kv;
  }
}

// transformers-4.x AdamW: m,v update; p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps); then p -= lr*wd*p
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float step_size, float lr_wd,
                             const float* __restrict__ dyn, float b1, float b2, float eps, float grad_scale) {
  if (dyn != nullptr) {  // {step_size, lr * wd} of this step from device memory (CUDA-graph replay)
    step_size = dyn[0];
    lr_wd = dyn[1];
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gr = g[i] * grad_scale;
    const float mi = m[i] * b1 + (1.0f - b1) * gr;
    const float vi = v[i] * b2 + (1.0f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    float pi = p[i] - step_size * (mi / (sqrtf(vi) + eps));
    pi -= lr_wd * pi;
    p[i] = pi;
  }
}

int byol_loss(const float* student, const float* teacher, int ncrops, int B, float* dstudent, float* acc_ws,
              cudaStream_t st) {
  // acc_ws: [1 + 4*256] floats, zeroed here; stats stay un-finalised so DDP can all-reduce them first
  ATST_REQUIRE(ncrops >= 2 && B > 0, "byol_loss: need ncrops >= 2, B > 0");
  cudaError_t e = cudaMemsetAsync(acc_ws, 0, (1 + 4 * kOut) * sizeof(float), st);
  if (e != cudaSuccess) { atst_set_error("byol_loss memset: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
  int grid = (ncrops * B + 7) / 8;
  if (grid > 148 * 4) grid = 148 * 4;
  byol_loss_kernel<<<grid, 256, 0, st>>>(student, teacher, ncrops, B, dstudent, acc_ws, acc_ws + 1, 0);
  return atst_check_launch("byol_loss_kernel");
}
int byol_finalize(const float* acc_ws, float n_student_rows, float n_teacher_rows, int ncrops, int B, float* out3,
                  cudaStream_t st) {
  byol_finalize_kernel<<<1, 256, 0, st>>>(acc_ws, acc_ws + 1, n_student_rows, n_teacher_rows,
                                          static_cast<float>(2 * ncrops - 2) * B, out3);
  return atst_check_launch("byol_finalize_kernel");
}
int ema_update(float* k, const float* q, float m, const float* m_dev, long long n, cudaStream_t st) {
  ATST_REQUIRE(n % 4 == 0, "ema_update: n %% 4 != 0");
  long long g = (n / 4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  ema_kernel<<<static_cast<int>(g), 256, 0, st>>>(k, q, m, m_dev, n / 4);
  return atst_check_launch("ema_kernel");
}
int adamw_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float wd, float b1,
               float b2, float eps, float grad_scale, const float* dyn, cudaStream_t st) {
  ATST_REQUIRE(step >= 1 || dyn != nullptr, "adamw_step: step counts from 1");
  if (step < 1) step = 1;
  const double bc1 = 1.0 - pow(static_cast<double>(b1), step), bc2 = 1.0 - pow(static_cast<double>(b2), step);
  const float step_size = static_cast<float>(lr * sqrt(bc2) / bc1);
  long long gr = (n + 255) / 256;
  if (gr > 148 * 16) gr = 148 * 16;
  adamw_kernel<<<static_cast<int>(gr), 256, 0, st>>>(p, g, m, v, n, step_size, lr * wd, dyn, b1, b2, eps, grad_scale);
  return atst_check_launch("adamw_kernel");
}

}  // namespace atst
