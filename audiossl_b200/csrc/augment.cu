// Device-batched BYOL-A style augmentations that sit between the mel kernel and the encoder in the training
// recipe (audiossl/methods/atst/transform.py:35-46,68-73):
//   * log-mixup-exp against a device-resident memory bank (audiossl/transforms/byol_a.py:61-82,85-115)
//   * RandomResizeCrop: zero "virtual canvas" + random crop + bicubic(align_corners=True) resize back
//     (audiossl/transforms/byol_a.py:7-49; torch upsample_bicubic2d semantics, A = -0.75, border taps clamped
//     to the crop rectangle)
// Random draws are made by the host wrapper; the kernels are pure functions of (input, parameters).
#include "common.cuh"

namespace atst {

// out[b] = log((1 - a_b) * exp(x[b]) + a_b * exp(z_b) + eps), z_b = bank entry idx_b; idx_b < 0 => out = x (empty bank).
// x [B, Hm, x_T]; bank entries are [Hm, bank_T] with zlen_b valid frames.  Length mismatch as log_mixup_exp
// (byol_a.py:61-82): a longer z is read from the window [start_b, start_b + x_T); a shorter z is mixed into the window
// [start_b, start_b + zlen_b) of x and the frames outside it become log(exp(x) + eps).
__global__ void mixup_kernel(const float* __restrict__ x, int x_T, const float* __restrict__ bank, int bank_T,
                             const int* __restrict__ idx, const int* __restrict__ zlen, const int* __restrict__ start,
                             const float* __restrict__ alpha, float* __restrict__ out, int Hm, int B) {
  const int b = blockIdx.y;
  const int j = idx[b];
  const float a = alpha[b];
  const int lb = zlen ? zlen[b] : x_T;
  const int s0 = start ? start[b] : 0;
  const long long per_clip = static_cast<long long>(Hm) * x_T;
  const float* xb = x + b * per_clip;
  const float* zb = j >= 0 ? bank + static_cast<long long>(j) * Hm * bank_T : nullptr;
  float* ob = out + b * per_clip;
  const float eps = 1.1920928955078125e-07f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_clip;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float xv = xb[i];
    if (!zb) {
      ob[i] = xv;
      continue;
    }
    const int r = static_cast<int>(i / x_T), t = static_cast<int>(i - static_cast<long long>(r) * x_T);
    int tz = t;            // frame of z mixed into frame t of x (-1: none)
    if (x_T < lb) tz = s0 + t;
    else if (x_T > lb) tz = (t >= s0 && t < s0 + lb) ? t - s0 : -1;
    ob[i] = tz >= 0 ? logf((1.0f - a) * expf(xv) + a * expf(zb[static_cast<long long>(r) * bank_T + tz]) + eps)
                    : logf(expf(xv) + eps);
  }
}

__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

// lms [B, Hm, T] -> out [B, Hm, T]; crop rectangle (i, j, h, w) per clip on a canvas [canvas_h, canvas_w] that holds
// the input at offset (y0, x0) and zeros elsewhere
__global__ void resize_crop_kernel(const float* __restrict__ lms, const int* __restrict__ rect, float* __restrict__ out,
                                   int Hm, int T, int canvas_h, int canvas_w, int y0, int x0) {
  const int b = blockIdx.z;
  const int oy = blockIdx.y;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (ox >= T) return;
  const int ci = rect[4 * b], cj = rect[4 * b + 1], ch = rect[4 * b + 2], cw = rect[4 * b + 3];
  const float sy = Hm > 1 ? static_cast<float>(ch - 1) / (Hm - 1) : 0.f;
  const float sx = T > 1 ? static_cast<float>(cw - 1) / (T - 1) : 0.f;
  const float ry = sy * oy, rx = sx * ox;
  const int iy = static_cast<int>(floorf(ry)), ix = static_cast<int>(floorf(rx));
  float wy[4], wx[4];
  cubic_coeffs(ry - iy, wy);
  cubic_coeffs(rx - ix, wx);
  const float* src = lms + static_cast<long long>(b) * Hm * T;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), ch - 1) + ci;  // canvas row (taps clamped to the crop)
    float row = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xx = min(max(ix - 1 + c, 0), cw - 1) + cj;  // canvas column
      const int sy_ = yy - y0, sx_ = xx - x0;
      const float v = (sy_ >= 0 && sy_ < Hm && sx_ >= 0 && sx_ < T) ? src[sy_ * T + sx_] : 0.f;
      row = fmaf(wx[c], v, row);
    }
    acc = fmaf(wy[a], row, acc);
  }
  out[(static_cast<long long>(b) * Hm + oy) * T + ox] = acc;
}

int mixup_forward(const float* x, int x_T, const float* bank, int bank_T, const int* idx, const int* zlen,
                  const int* start, const float* alpha, float* out, int Hm, int B, cudaStream_t st) {
  ATST_REQUIRE(B > 0 && Hm > 0 && x_T > 0 && bank_T > 0, "mixup_forward: empty batch");
  const long long per_clip = static_cast<long long>(Hm) * x_T;
  int gx = static_cast<int>((per_clip + 255) / 256);
  if (gx > 64) gx = 64;
  mixup_kernel<<<dim3(gx, B), 256, 0, st>>>(x, x_T, bank, bank_T, idx, zlen, start, alpha, out, Hm, B);
  return atst_check_launch("mixup_kernel");
}

int resize_crop_forward(const float* lms, const int* rect, float* out, int B, int Hm, int T, int canvas_h,
                        int canvas_w, cudaStream_t st) {
  ATST_REQUIRE(B > 0 && Hm > 0 && T > 0 && canvas_h >= Hm && canvas_w >= T, "resize_crop_forward: bad shape");
  const int y0 = (canvas_h - Hm) / 2, x0 = (canvas_w - T) / 2;
  resize_crop_kernel<<<dim3((T + 127) / 128, Hm, B), 128, 0, st>>>(lms, rect, out, Hm, T, canvas_h, canvas_w, y0, x0);
  return atst_check_launch("resize_crop_kernel");
}

}  // namespace atst
