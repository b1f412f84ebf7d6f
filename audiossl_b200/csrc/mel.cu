// Fused log-mel front-end (reference call site audiossl/methods/atst/transform.py:14-29,
// arithmetic = torchaudio MelSpectrogram(16000, n_fft 1024, hop 160, win 1024|640, f 60..7800, 64 HTK mels,
// power 2) -> AmplitudeToDB("power", top_db 80, per clip) -> MinMax(-79.6482, 50.6842)).
//
// Kernel 1 (mel_db_kernel): one CTA = 32 consecutive frames of one clip.  The 5984-sample waveform segment
// is staged once in shared memory with coalesced float4 loads (6.4x frame overlap is served from smem, the
// waveform is read from HBM once), reflect padding is applied at the clip edges.  Four frames are transformed
// at a time, 64 threads per frame: real 1024-point FFT = 512-point complex Stockham FFT (3 radix-8 passes in
// registers, padded smem exchange) + split-radix untangle; |X|^2; banded mel dot (each of the 64 threads owns
// one triangular band, <= 39 bins); 10*log10(max(.,1e-10)).  Output tile [64 mel][32 frames] is transposed in
// smem and written with 128-byte row segments.  Per-clip max of the dB values via warp-reduce + atomicMax.
// Kernel 2 (mel_norm_kernel): clamp at (clip max - 80 dB) and MinMax-normalise in place.
#include <math.h>
#include "common.cuh"

namespace atst {

constexpr int kNfft = 1024;
constexpr int kHop = 160;
constexpr int kMels = 64;
constexpr int kBins = 513;
constexpr int kFramesPerCta = 32;
constexpr int kSeg = (kFramesPerCta - 1) * kHop + kNfft;  // 5984 samples
constexpr int kMaxW = 1024;                                // upper bound on non-zero mel weights (970 used)

struct MelTables {
  float2 tw512[512];    // exp(-2 pi i m / 512)
  float2 tw1024[512];   // exp(-2 pi i k / 1024), k < 512
  float window[2][kNfft];  // [0]: win_length 1024, [1]: win_length 640 (zero padded, centred)
  float w[kMaxW];       // band weights, band-major
  int band_start[kMels];   // first fft bin of band m
  int band_len[kMels];
  int band_off[kMels];     // offset of band m in w[]
};

__device__ MelTables g_mel_tables;

// XOR swizzle of the FFT exchange buffers (index < 512): conflict-free for all four access patterns of the three
// radix-8 passes (stride-8 scatter, stride-1 gather, the 64a + k + 8c scatter of pass 2, stride-64 scatter); the
// former i + (i >> 5) padding left the pass-2 scatter 4-way conflicted (ncu: 46 % of the kernel's shared wavefronts)
__device__ __forceinline__ int padi(int i) { return i ^ ((i >> 3) & 31); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// in-register forward DFT-8 (decimation in frequency); result r is left in slot kRev[r]
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float s = 0.70710678118654752f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = v[i];
    v[i] = make_float2(t.x + v[i + 4].x, t.y + v[i + 4].y);
    v[i + 4] = make_float2(t.x - v[i + 4].x, t.y - v[i + 4].y);
  }
  v[5] = make_float2(s * (v[5].x + v[5].y), s * (v[5].y - v[5].x));    // * (1 - i)/sqrt2
  v[6] = make_float2(v[6].y, -v[6].x);                                 // * (-i)
  v[7] = make_float2(s * (v[7].y - v[7].x), -s * (v[7].x + v[7].y));   // * (-1 - i)/sqrt2
#pragma unroll
  for (int h = 0; h < 8; h += 4) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float2 t = v[h + i];
      v[h + i] = make_float2(t.x + v[h + i + 2].x, t.y + v[h + i + 2].y);
      v[h + i + 2] = make_float2(t.x - v[h + i + 2].x, t.y - v[h + i + 2].y);
    }
    v[h + 3] = make_float2(v[h + 3].y, -v[h + 3].x);
  }
#pragma unroll
  for (int h = 0; h < 8; h += 2) {
    float2 t = v[h];
    v[h] = make_float2(t.x + v[h + 1].x, t.y + v[h + 1].y);
    v[h + 1] = make_float2(t.x - v[h + 1].x, t.y - v[h + 1].y);
  }
}

__global__ void __launch_bounds__(256)
mel_db_kernel(const float* __restrict__ wav, int n, long long wav_stride, int T, int win_idx,
              float* __restrict__ out, long long out_stride, unsigned int* __restrict__ clip_max_bits) {
  extern __shared__ float sm[];
  float* seg = sm;                       // [kSeg]
  float* zre = seg + kSeg;               // [4][528]
  float* zim = zre + 4 * 528;            // [4][528]
  float* pw = zim + 4 * 528;             // [4][520]
  float* tile = pw + 4 * 520;            // [64][33]
  float* s_win = tile + kMels * 33;      // [1024]
  float2* s_tw = reinterpret_cast<float2*>(s_win + kNfft);  // [512]
  float2* s_tw2 = s_tw + 512;            // [512]
  float* s_w = reinterpret_cast<float*>(s_tw2 + 512);       // [kMaxW]

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerCta;
  const int tid = threadIdx.x;
  const float* x = wav + static_cast<long long>(b) * wav_stride;

  // tables -> smem
  for (int i = tid; i < kNfft; i += 256) s_win[i] = g_mel_tables.window[win_idx][i];
  for (int i = tid; i < 512; i += 256) {
    s_tw[i] = g_mel_tables.tw512[i];
    s_tw2[i] = g_mel_tables.tw1024[i];
  }
  for (int i = tid; i < kMaxW; i += 256) s_w[i] = g_mel_tables.w[i];

  // waveform segment: sample index s = t0*160 - 512 + i ; reflect (no edge repeat) outside [0, n)
  const long long s0 = static_cast<long long>(t0) * kHop - kNfft / 2;
  if (s0 >= 0 && s0 + kSeg <= n && ((reinterpret_cast<uintptr_t>(x + s0) & 15) == 0)) {
    const float4* src = reinterpret_cast<const float4*>(x + s0);
    for (int i = tid; i < kSeg / 4; i += 256) reinterpret_cast<float4*>(seg)[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < kSeg; i += 256) {
      long long s = s0 + i;
      if (s < 0) s = -s;
      if (s >= n) s = 2LL * (n - 1) - s;
      float v = 0.f;
      if (s >= 0 && s < n) v = __ldg(x + s);
      seg[i] = v;
    }
  }
  __syncthreads();

  const int slot = tid >> 6;  // frame slot 0..3
  const int q = tid & 63;
  float* re = zre + slot * 528;
  float* im = zim + slot * 528;
  float* P = pw + slot * 520;
  const int band_start = g_mel_tables.band_start[q];
  const int band_len = g_mel_tables.band_len[q];
  const int band_off = g_mel_tables.band_off[q];
  float local_max = -INFINITY;

  for (int it = 0; it < kFramesPerCta / 4; ++it) {
    const int f = it * 4 + slot;  // frame within the CTA
    const float* fr = seg + f * kHop;
    float2 v[8];
    // ---- pass 1 (Ns = 1): inputs z[q + 64 r] = (w x)[2j], (w x)[2j+1]; no twiddles
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int j = q + 64 * r;
      const float2 xx = *reinterpret_cast<const float2*>(fr + 2 * j);
      const float2 ww = *reinterpret_cast<const float2*>(s_win + 2 * j);
      v[r] = make_float2(xx.x * ww.x, xx.y * ww.y);
    }
    dft8(v);
    {
      const int d = q * 8;
      re[padi(d + 0)] = v[0].x; im[padi(d + 0)] = v[0].y;
      re[padi(d + 4)] = v[1].x; im[padi(d + 4)] = v[1].y;
      re[padi(d + 2)] = v[2].x; im[padi(d + 2)] = v[2].y;
      re[padi(d + 6)] = v[3].x; im[padi(d + 6)] = v[3].y;
      re[padi(d + 1)] = v[4].x; im[padi(d + 1)] = v[4].y;
      re[padi(d + 5)] = v[5].x; im[padi(d + 5)] = v[5].y;
      re[padi(d + 3)] = v[6].x; im[padi(d + 3)] = v[6].y;
      re[padi(d + 7)] = v[7].x; im[padi(d + 7)] = v[7].y;
    }
    __syncthreads();
    // ---- pass 2 (Ns = 8)
    {
      const int k = q & 7;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = padi(q + 64 * r);
        v[r] = make_float2(re[j], im[j]);
        if (r > 0) v[r] = cmul(v[r], s_tw[r * k * 8]);
      }
      dft8(v);
      __syncthreads();
      const int d = (q >> 3) * 64 + k;
      re[padi(d + 0 * 8)] = v[0].x; im[padi(d + 0 * 8)] = v[0].y;
      re[padi(d + 4 * 8)] = v[1].x; im[padi(d + 4 * 8)] = v[1].y;
      re[padi(d + 2 * 8)] = v[2].x; im[padi(d + 2 * 8)] = v[2].y;
      re[padi(d + 6 * 8)] = v[3].x; im[padi(d + 6 * 8)] = v[3].y;
      re[padi(d + 1 * 8)] = v[4].x; im[padi(d + 1 * 8)] = v[4].y;
      re[padi(d + 5 * 8)] = v[5].x; im[padi(d + 5 * 8)] = v[5].y;
      re[padi(d + 3 * 8)] = v[6].x; im[padi(d + 3 * 8)] = v[6].y;
      re[padi(d + 7 * 8)] = v[7].x; im[padi(d + 7 * 8)] = v[7].y;
    }
    __syncthreads();
    // ---- pass 3 (Ns = 64)
    {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = padi(q + 64 * r);
        v[r] = make_float2(re[j], im[j]);
        if (r > 0) v[r] = cmul(v[r], s_tw[r * q]);
      }
      dft8(v);
      __syncthreads();
      re[padi(q + 0 * 64)] = v[0].x; im[padi(q + 0 * 64)] = v[0].y;
      re[padi(q + 4 * 64)] = v[1].x; im[padi(q + 4 * 64)] = v[1].y;
      re[padi(q + 2 * 64)] = v[2].x; im[padi(q + 2 * 64)] = v[2].y;
      re[padi(q + 6 * 64)] = v[3].x; im[padi(q + 6 * 64)] = v[3].y;
      re[padi(q + 1 * 64)] = v[4].x; im[padi(q + 1 * 64)] = v[4].y;
      re[padi(q + 5 * 64)] = v[5].x; im[padi(q + 5 * 64)] = v[5].y;
      re[padi(q + 3 * 64)] = v[6].x; im[padi(q + 3 * 64)] = v[6].y;
      re[padi(q + 7 * 64)] = v[7].x; im[padi(q + 7 * 64)] = v[7].y;
    }
    __syncthreads();
    // ---- untangle the packed real FFT and take |X[k]|^2, k = 0..512
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int k = q + 64 * s;
      const int kk = (512 - k) & 511;
      const float2 a = make_float2(re[padi(k)], im[padi(k)]);
      const float2 c = make_float2(re[padi(kk)], -im[padi(kk)]);  // conj(Z[512-k])
      const float2 ze = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y + c.y));
      const float2 d = make_float2(0.5f * (a.x - c.x), 0.5f * (a.y - c.y));
      const float2 zo = make_float2(d.y, -d.x);  // d / i
      const float2 t = cmul(zo, s_tw2[k]);
      const float xr = ze.x + t.x, xi = ze.y + t.y;
      P[k] = xr * xr + xi * xi;
      if (k == 0) {  // Nyquist bin 512: Ze[0] - Zo[0]
        const float yr = ze.x - zo.x, yi = ze.y - zo.y;
        P[512] = yr * yr + yi * yi;
      }
    }
    __syncthreads();
    // ---- banded mel dot + dB
    {
      float acc = 0.f;
      for (int i = 0; i < band_len; ++i) acc = fmaf(s_w[band_off + i], P[band_start + i], acc);
      const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
      tile[q * 33 + f] = db;
      if (t0 + f < T) local_max = fmaxf(local_max, db);
    }
    // (the next iteration's first write to re/im happens after its own dft8; P is rewritten only after
    //  two more __syncthreads, so no barrier is needed here)
  }
  __syncthreads();
  // ---- coalesced store of the [64][32] tile: each warp writes 8 mel rows, 32 consecutive frames per row
  {
    const int wrp = tid >> 5, ln = tid & 31;
    float* o = out + static_cast<long long>(b) * out_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = wrp * 8 + i;
      if (t0 + ln < T) o[static_cast<long long>(m) * T + t0 + ln] = tile[m * 33 + ln];
    }
  }
  // ---- per-clip max of the dB values (order-preserving uint encoding of floats)
  local_max = warp_max(local_max);
  if ((tid & 31) == 0 && local_max > -INFINITY) {
    unsigned int u = __float_as_uint(local_max);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    atomicMax(clip_max_bits + b, u);
  }
}

__global__ void mel_norm_kernel(float* __restrict__ mel, long long per_clip, long long out_stride,
                                const unsigned int* __restrict__ clip_max_bits, float top_db, float mn, float range) {
  const int b = blockIdx.y;
  unsigned int u = clip_max_bits[b];
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  const float floor_db = __uint_as_float(u) - top_db;
  float* p = mel + static_cast<long long>(b) * out_stride;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_clip;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float db = fmaxf(p[i], floor_db);
    p[i] = (db - mn) / range * 2.0f - 1.0f;
  }
}

__global__ void fill_u32_kernel(unsigned int* p, int n, unsigned int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// --------------------------------------------------------------------------- host tables
static bool g_tables_ready = false;

static int init_tables(cudaStream_t stream) {
  if (g_tables_ready) return ATST_OK;
  static MelTables h;  // static: 20 KB, keep off the stack
  const double pi = 3.14159265358979323846;
  for (int m = 0; m < 512; ++m) {
    h.tw512[m] = make_float2((float)cos(2.0 * pi * m / 512.0), (float)-sin(2.0 * pi * m / 512.0));
    h.tw1024[m] = make_float2((float)cos(2.0 * pi * m / 1024.0), (float)-sin(2.0 * pi * m / 1024.0));
  }
  const int wins[2] = {1024, 640};
  for (int wi = 0; wi < 2; ++wi) {
    const int wl = wins[wi], left = (kNfft - wl) / 2;
    for (int i = 0; i < kNfft; ++i) h.window[wi][i] = 0.f;
    for (int i = 0; i < wl; ++i) h.window[wi][left + i] = (float)(0.5 - 0.5 * cos(2.0 * pi * i / wl));
  }
  // HTK mel filterbank, norm=None (torchaudio.functional.melscale_fbanks): fp32 edge frequencies
  const double f_min = 60.0, f_max = 7800.0, sr2 = 8000.0;
  const double m_min = 2595.0 * log10(1.0 + f_min / 700.0), m_max = 2595.0 * log10(1.0 + f_max / 700.0);
  float f_pts[kMels + 2];
  for (int i = 0; i < kMels + 2; ++i) {
    const double mp = m_min + (m_max - m_min) * i / (kMels + 1);
    f_pts[i] = (float)(700.0 * (pow(10.0, mp / 2595.0) - 1.0));
  }
  int off = 0;
  for (int m = 0; m < kMels; ++m) {
    int start = -1, last = -1;
    static float wrow[kBins];
    for (int k = 0; k < kBins; ++k) {
      const float fk = (float)(sr2 * k / (kBins - 1));
      const float down = (fk - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const float up = (f_pts[m + 2] - fk) / (f_pts[m + 2] - f_pts[m + 1]);
      wrow[k] = fmaxf(0.f, fminf(down, up));
      if (wrow[k] > 0.f) {
        if (start < 0) start = k;
        last = k;
      }
    }
    const int len = start < 0 ? 0 : last - start + 1;
    if (off + len > kMaxW) { atst_set_error("mel: weight table overflow"); return ATST_ERR_ARG; }
    for (int i = 0; i < len; ++i) h.w[off + i] = wrow[start + i];
    h.band_start[m] = start < 0 ? 0 : start;
    h.band_len[m] = len;
    h.band_off[m] = off;
    off += len;
  }
  cudaError_t e = cudaMemcpyToSymbolAsync(g_mel_tables, &h, sizeof(MelTables), 0, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { atst_set_error("mel tables upload: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
  e = cudaStreamSynchronize(stream);  // one-time: the host buffer must stay valid until the copy is done
  if (e != cudaSuccess) { atst_set_error("mel tables upload: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
  g_tables_ready = true;
  return ATST_OK;
}

constexpr int kMelSmemBytes =
    (kSeg + 8 * 528 + 4 * 520 + kMels * 33 + kNfft + 2 * 512 * 2 + kMaxW) * 4;

int mel_forward(const float* wav, int B, int n, long long wav_stride, int win_length, float* out,
                long long out_stride, unsigned int* clip_max_ws, int normalize, cudaStream_t stream) {
  ATST_REQUIRE(B > 0 && n > kNfft / 2, "mel: need B > 0 and n > 512 (reflect padding), got B=%d n=%d", B, n);
  ATST_REQUIRE(win_length == 1024 || win_length == 640, "mel: win_length must be 1024 or 640, got %d", win_length);
  int rc = init_tables(stream);
  if (rc) return rc;
  const int T = n / kHop + 1;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mel_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMelSmemBytes);
    if (e != cudaSuccess) { atst_set_error("mel smem attr: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  fill_u32_kernel<<<(B + 255) / 256, 256, 0, stream>>>(clip_max_ws, B, 0u);
  dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
  mel_db_kernel<<<grid, 256, kMelSmemBytes, stream>>>(wav, n, wav_stride, T, win_length == 1024 ? 0 : 1, out,
                                                      out_stride, clip_max_ws);
  rc = atst_check_launch("mel_db_kernel");
  if (rc) return rc;
  if (normalize) {
    const long long per_clip = static_cast<long long>(kMels) * T;
    int gx = static_cast<int>((per_clip + 1023) / 1024);
    if (gx > 64) gx = 64;
    mel_norm_kernel<<<dim3(gx, B), 256, 0, stream>>>(out, per_clip, out_stride, clip_max_ws, 80.0f, -79.6482f,
                                                    50.6842f - (-79.6482f));
    rc = atst_check_launch("mel_norm_kernel");
  }
  return rc;
}

}  // namespace atst
