// Fused log-mel front-end (reference call site audiossl/methods/atst/transform.py:14-29,
// arithmetic = torchaudio MelSpectrogram(16000, n_fft 1024, hop 160, win 1024|640, f 60..7800, 64 HTK mels,
// power 2) -> AmplitudeToDB("power", top_db 80, per clip) -> MinMax(-79.6482, 50.6842)).
//
// mel_kernel: one CTA = 32 consecutive frames of one clip, 8 warps.  The 5984-sample waveform segment is staged once
// in shared memory - by ONE 1-D TMA bulk copy (cp.async.bulk, 23.9 KB, completion on an mbarrier) for interior CTAs,
// by a reflect-padding loop at the clip edges - so the 6.4x frame overlap is served from smem and the waveform is
// read from HBM once.  Each WARP then owns whole frames (4 per warp) and never meets the other warps again until the
// tile is stored: real 1024-point FFT = 512-point complex Stockham FFT (3 radix-8 passes in registers, every lane
// working two of the 64 butterfly columns, XOR-swizzled warp-private exchange buffers, __syncwarp only) + split-radix
// untangle; |X|^2; banded mel dot (each lane owns two triangular bands, <= 39 bins); 10*log10(max(.,1e-10)).  The
// window, twiddle and filterbank tables (16 KB) are read through L1 (__ldg) instead of being copied into every CTA's
// shared memory.  Output tile [64 mel][32 frames] is transposed in smem and written with 128-byte row segments.
// Per-clip max of the dB values via warp-reduce + atomicMax; the LAST CTA of a clip to finish (atomic ticket) applies
// the top_db clamp and the MinMax normalisation to the whole clip in place (256 KB, L2-resident), so there is no
// second kernel and no second pass over HBM.
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

constexpr int kWarps = 8;
constexpr int kFftBuf = 528;  // floats per re / im exchange buffer (512 swizzled + the Nyquist bin of the power spectrum)

// smem layout: mbarrier (16 B) | seg[kSeg] | per warp {re[528], im[528]} | tile[64][33]
constexpr int kMelSmemBytes = 16 + (kSeg + kWarps * 2 * kFftBuf + kMels * 33) * 4;

// dft8 leaves result r in slot kRev[r] = {0,4,2,6,1,5,3,7}
__device__ __forceinline__ void store8(float* re, float* im, const float2 (&v)[8], int d, int stride) {
  re[padi(d + 0 * stride)] = v[0].x; im[padi(d + 0 * stride)] = v[0].y;
  re[padi(d + 4 * stride)] = v[1].x; im[padi(d + 4 * stride)] = v[1].y;
  re[padi(d + 2 * stride)] = v[2].x; im[padi(d + 2 * stride)] = v[2].y;
  re[padi(d + 6 * stride)] = v[3].x; im[padi(d + 6 * stride)] = v[3].y;
  re[padi(d + 1 * stride)] = v[4].x; im[padi(d + 1 * stride)] = v[4].y;
  re[padi(d + 5 * stride)] = v[5].x; im[padi(d + 5 * stride)] = v[5].y;
  re[padi(d + 3 * stride)] = v[6].x; im[padi(d + 3 * stride)] = v[6].y;
  re[padi(d + 7 * stride)] = v[7].x; im[padi(d + 7 * stride)] = v[7].y;
}

__global__ void __launch_bounds__(256, 3)
mel_kernel(const float* __restrict__ wav, int n, long long wav_stride, const long long* __restrict__ clip_start,
           int T, int win_idx, float* __restrict__ out, long long out_stride, unsigned int* __restrict__ clip_ws,
           int B, int normalize, float top_db, float mn, float range) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float* seg = reinterpret_cast<float*>(smem_raw + 16);   // [kSeg], 16-byte aligned for the bulk copy
  float* fft = seg + kSeg;                                  // [kWarps][2][kFftBuf]
  float* tile = fft + kWarps * 2 * kFftBuf;                 // [64][33]
  __shared__ int s_ticket;

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerCta;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const float* x = wav + static_cast<long long>(b) * wav_stride + (clip_start ? clip_start[b] : 0);

  // waveform segment: sample index s = t0*160 - 512 + i ; reflect (no edge repeat) outside [0, n)
  const long long s0 = static_cast<long long>(t0) * kHop - kNfft / 2;
  const bool bulk = s0 >= 0 && s0 + kSeg <= n && ((reinterpret_cast<uintptr_t>(x + s0) & 15) == 0);
  if (bulk) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
      mbar_expect_tx(bar, kSeg * 4);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(smem_u32(seg)), "l"(x + s0), "r"(kSeg * 4), "r"(smem_u32(bar)) : "memory");
    }
    __syncthreads();  // the barrier is initialised before anyone waits on it
    mbar_wait(bar, 0);
  } else {
    for (int i = tid; i < kSeg; i += 256) {
      long long s = s0 + i;
      if (s < 0) s = -s;
      if (s >= n) s = 2LL * (n - 1) - s;
      float v = 0.f;
      if (s >= 0 && s < n) v = __ldg(x + s);
      seg[i] = v;
    }
    __syncthreads();
  }

  float* re = fft + warp * 2 * kFftBuf;
  float* im = re + kFftBuf;
  const float* g_win = g_mel_tables.window[win_idx];
  const float2* g_tw = g_mel_tables.tw512;
  const float2* g_tw2 = g_mel_tables.tw1024;
  const float* g_w = g_mel_tables.w;
  float local_max = -INFINITY;

  for (int f = warp; f < kFramesPerCta; f += kWarps) {
    const float* fr = seg + f * kHop;
    float2 v[2][8];
    // ---- pass 1 (Ns = 1): inputs z[q + 64 r] = (w x)[2j], (w x)[2j+1]; no twiddles.  q = lane, lane + 32
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = lane + 32 * h;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = q + 64 * r;
        const float2 xx = *reinterpret_cast<const float2*>(fr + 2 * j);
        const float2 ww = __ldg(reinterpret_cast<const float2*>(g_win + 2 * j));
        v[h][r] = make_float2(xx.x * ww.x, xx.y * ww.y);
      }
      dft8(v[h]);
    }
    __syncwarp();  // the previous frame's band sums have been read out of `re`
#pragma unroll
    for (int h = 0; h < 2; ++h) store8(re, im, v[h], (lane + 32 * h) * 8, 1);
    __syncwarp();
    // ---- pass 2 (Ns = 8)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = lane + 32 * h, k = q & 7;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = padi(q + 64 * r);
        v[h][r] = make_float2(re[j], im[j]);
        if (r > 0) v[h][r] = cmul(v[h][r], __ldg(g_tw + r * k * 8));
      }
      dft8(v[h]);
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = lane + 32 * h;
      store8(re, im, v[h], (q >> 3) * 64 + (q & 7), 8);
    }
    __syncwarp();
    // ---- pass 3 (Ns = 64)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = lane + 32 * h;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int j = padi(q + 64 * r);
        v[h][r] = make_float2(re[j], im[j]);
        if (r > 0) v[h][r] = cmul(v[h][r], __ldg(g_tw + r * q));
      }
      dft8(v[h]);
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) store8(re, im, v[h], lane + 32 * h, 64);
    __syncwarp();
    // ---- untangle the packed real FFT and take |X[k]|^2, k = 0..512 (into registers, then over `re`).
    // Bins k and 512 - k share everything but a sign: with Ze = (Z[k] + conj Z[512-k]) / 2, Zo = (Z[k] - conj Z[512-k]) / 2i
    // and t = W^k Zo,  X[k] = Ze + t  and  X[512-k] = conj(Ze - t)  (W^(512-k) = -conj W^k), so one lane takes the
    // pair; k = 0 pairs with the Nyquist bin 512 the same way, k = 256 is its own partner (lane 0).
    float pk_lo[8], pk_hi[8], mid = 0.f;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int k = lane + 32 * s;       // 0 .. 255
      const int kk = (512 - k) & 511;
      const float2 a = make_float2(re[padi(k)], im[padi(k)]);
      const float2 c = make_float2(re[padi(kk)], -im[padi(kk)]);  // conj(Z[512-k])
      const float2 ze = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y + c.y));
      const float2 d = make_float2(0.5f * (a.x - c.x), 0.5f * (a.y - c.y));
      const float2 zo = make_float2(d.y, -d.x);  // d / i
      const float2 t = cmul(zo, __ldg(g_tw2 + k));
      const float xr = ze.x + t.x, xi = ze.y + t.y;
      const float yr = ze.x - t.x, yi = ze.y - t.y;
      pk_lo[s] = xr * xr + xi * xi;
      pk_hi[s] = yr * yr + yi * yi;
    }
    if (lane == 0) {
      const float2 a = make_float2(re[padi(256)], im[padi(256)]);
      const float2 t = cmul(make_float2(a.y, 0.f), __ldg(g_tw2 + 256));  // Ze = (a.x, 0), Zo = (a.y, 0)
      const float xr = a.x + t.x, xi = t.y;
      mid = xr * xr + xi * xi;
    }
    __syncwarp();
    float* P = re;  // power spectrum, plain index 0..512
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      P[lane + 32 * s] = pk_lo[s];
      P[512 - (lane + 32 * s)] = pk_hi[s];
    }
    if (lane == 0) P[256] = mid;
    __syncwarp();
    // ---- banded mel dot + dB: bands lane and lane + 32
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      const int band_start = __ldg(g_mel_tables.band_start + m);
      const int band_len = __ldg(g_mel_tables.band_len + m);
      const float* wrow = g_w + __ldg(g_mel_tables.band_off + m);
      float acc = 0.f;
      for (int i = 0; i < band_len; ++i) acc = fmaf(__ldg(wrow + i), P[band_start + i], acc);
      const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
      tile[m * 33 + f] = db;
      if (t0 + f < T) local_max = fmaxf(local_max, db);
    }
  }
  __syncthreads();
  // ---- coalesced store of the [64][32] tile: each warp writes 8 mel rows, 32 consecutive frames per row
  float* o = out + static_cast<long long>(b) * out_stride;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = warp * 8 + i;
    if (t0 + lane < T) o[static_cast<long long>(m) * T + t0 + lane] = tile[m * 33 + lane];
  }
  // ---- per-clip max of the dB values (order-preserving uint encoding of floats)
  local_max = warp_max(local_max);
  if (lane == 0 && local_max > -INFINITY) {
    unsigned int u = __float_as_uint(local_max);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    atomicMax(clip_ws + b, u);
  }
  if (!normalize) return;
  // ---- the last CTA of the clip clamps at (clip max - top_db) and MinMax-normalises the whole clip in place
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = static_cast<int>(atomicAdd(clip_ws + B + b, 1u));
  __syncthreads();
  if (s_ticket != static_cast<int>(gridDim.x) - 1) return;
  __threadfence();
  unsigned int u = *reinterpret_cast<volatile unsigned int*>(clip_ws + b);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  const float floor_db = __uint_as_float(u) - top_db;
  const int per_clip = kMels * T;
  for (int i = tid; i < per_clip; i += 256) {
    const float db = fmaxf(__ldcg(o + i), floor_db);
    o[i] = (db - mn) / range * 2.0f - 1.0f;
  }
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


int mel_forward(const float* wav, int B, int n, long long wav_stride, const long long* clip_start, int win_length,
                float* out, long long out_stride, unsigned int* clip_ws, int normalize, cudaStream_t stream) {
  ATST_REQUIRE(B > 0 && n > kNfft / 2, "mel: need B > 0 and n > 512 (reflect padding), got B=%d n=%d", B, n);
  ATST_REQUIRE(win_length == 1024 || win_length == 640, "mel: win_length must be 1024 or 640, got %d", win_length);
  ATST_REQUIRE(B <= 65535, "mel: at most 65535 clips per call, got %d", B);
  int rc = init_tables(stream);
  if (rc) return rc;
  const int T = n / kHop + 1;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMelSmemBytes);
    if (e != cudaSuccess) { atst_set_error("mel smem attr: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
    configured = true;
  }
  // workspace: [B] per-clip max (order-preserving bits, 0 = below every float) | [B] finished-CTA tickets
  cudaError_t e = cudaMemsetAsync(clip_ws, 0, sizeof(unsigned int) * 2 * B, stream);
  if (e != cudaSuccess) { atst_set_error("mel workspace memset: %s", cudaGetErrorString(e)); return ATST_ERR_CUDA; }
  dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
  mel_kernel<<<grid, 256, kMelSmemBytes, stream>>>(wav, n, wav_stride, clip_start, T, win_length == 1024 ? 0 : 1, out,
                                                   out_stride, clip_ws, B, normalize, 80.0f, -79.6482f,
                                                   50.6842f - (-79.6482f));
  return atst_check_launch("mel_kernel");
}

}  // namespace atst
