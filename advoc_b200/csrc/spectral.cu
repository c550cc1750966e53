// Framed STFT / mel feature kernels.
//
// One CTA transforms TWO consecutive frames of one (batch, channel) signal with a
// single complex Stockham FFT (frame a -> real part, frame b -> imaginary part),
// separates the two real spectra, and -- in the fused variant -- applies the mel
// filterbank, 20*log10 and the [0,1] clip before anything leaves the SM.  The
// waveform is read once (frames overlap 4x, served from L1/L2), the only HBM
// write is the final feature.  Replaces advoc/spectral.py:11-41,60-83,98-227.
#include "common.cuh"

#include <math.h>

namespace advoc {

namespace {

struct StftArgs {
  const float* wav;     // [batch, nsamps, 1, nch]
  const float* window;  // [nfft]
  const float2* tw;     // [nfft] exp(-2 pi i j / nfft)
  float* out_c64;       // [batch, frames, bins, nch, 2] or null
  float* out_mag;       // [batch, frames, bins, nch] or null
  const float* mel_fb;  // [nmels, bins] or null (fused mel mode)
  const int* mel_range; // [nmels, 2] nonzero span per filter, or null (dense)
  float* out_mel;       // [batch, frames, nmels, nch]
  int batch, nsamps, nch, nfft, nhop, frames, bins, nmels;
  float min_level, min_db, ref_db;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Complex in-smem Stockham autosort FFT, radix 4 with one trailing radix-2 pass when
// log2(N) is odd.  `a` holds the input, result pointer is returned (a or b).
template <bool kInv>
__device__ float2* stockham_fft(float2* a, float2* b, const float2* __restrict__ tw, int N) {
  int n = N, s = 1;
  while (n >= 4) {
    const int n1 = n >> 2;
    const int tws = N / n;
    for (int t = threadIdx.x; t < (N >> 2); t += blockDim.x) {
      const int p = t / s, q = t - p * s;
      const float2 A = a[q + s * p];
      const float2 B = a[q + s * (p + n1)];
      const float2 C = a[q + s * (p + 2 * n1)];
      const float2 D = a[q + s * (p + 3 * n1)];
      const float2 apc = make_float2(A.x + C.x, A.y + C.y);
      const float2 amc = make_float2(A.x - C.x, A.y - C.y);
      const float2 bpd = make_float2(B.x + D.x, B.y + D.y);
      const float2 jbmd = make_float2(-(B.y - D.y), B.x - D.x);  // i * (B - D)
      float2 w1 = __ldg(tw + p * tws);
      float2 w2 = __ldg(tw + 2 * p * tws);
      float2 w3 = __ldg(tw + 3 * p * tws);
      if (kInv) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
      const int o = q + s * 4 * p;
      const float sg = kInv ? -1.f : 1.f;   // inverse transform: +i instead of -i rotations
      b[o] = make_float2(apc.x + bpd.x, apc.y + bpd.y);
      b[o + s] = cmul(w1, make_float2(amc.x - sg * jbmd.x, amc.y - sg * jbmd.y));
      b[o + 2 * s] = cmul(w2, make_float2(apc.x - bpd.x, apc.y - bpd.y));
      b[o + 3 * s] = cmul(w3, make_float2(amc.x + sg * jbmd.x, amc.y + sg * jbmd.y));
    }
    __syncthreads();
    float2* t2 = a; a = b; b = t2;
    n >>= 2;
    s <<= 2;
  }
  if (n == 2) {
    for (int q = threadIdx.x; q < s; q += blockDim.x) {
      const float2 A = a[q], B = a[q + s];
      b[q] = make_float2(A.x + B.x, A.y + B.y);
      b[q + s] = make_float2(A.x - B.x, A.y - B.y);
    }
    __syncthreads();
    float2* t2 = a; a = b; b = t2;
  }
  return a;
}

// smem: [2*nfft float2 ping-pong][2*bins float magnitudes (mel mode)]
template <bool kPow2>
__global__ void __launch_bounds__(256) stft_pair_kernel(const StftArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* bufa = reinterpret_cast<float2*>(smem_raw);
  float2* bufb = bufa + g.nfft;
  float* mag = reinterpret_cast<float*>(bufb + g.nfft);  // [2][bins], mel mode only

  const int pairs = (g.frames + 1) >> 1;
  const int pair = blockIdx.x % pairs;
  const int bc = blockIdx.x / pairs;
  const int c = bc % g.nch, b = bc / g.nch;
  const int f0 = pair * 2;
  const bool has_b = (f0 + 1) < g.frames;
  const float* x = g.wav + ((size_t)b * g.nsamps) * g.nch + c;

  // load + window; samples beyond the signal are the reference's zero tail padding
  for (int i = threadIdx.x; i < g.nfft; i += blockDim.x) {
    const float w = __ldg(g.window + i);
    const long ia = (long)f0 * g.nhop + i;
    const long ib = ia + g.nhop;
    const float va = ia < g.nsamps ? __ldg(x + ia * g.nch) * w : 0.f;
    const float vb = (has_b && ib < g.nsamps) ? __ldg(x + ib * g.nch) * w : 0.f;
    bufa[i] = make_float2(va, vb);
  }
  __syncthreads();

  const float2* Z;
  if (kPow2) {
    Z = stockham_fft<false>(bufa, bufb, g.tw, g.nfft);
  } else {
    // generic length: direct DFT (tacotron2 preset nfft=1200); O(N^2), correctness path
    for (int k = threadIdx.x; k < g.nfft; k += blockDim.x) {
      float2 acc = make_float2(0.f, 0.f);
      int idx = 0;
      for (int n = 0; n < g.nfft; ++n) {
        const float2 w = __ldg(g.tw + idx);
        const float2 v = bufa[n];
        acc.x += v.x * w.x - v.y * w.y;
        acc.y += v.x * w.y + v.y * w.x;
        idx += k;
        if (idx >= g.nfft) idx -= g.nfft;
      }
      bufb[k] = acc;
    }
    __syncthreads();
    Z = bufb;
  }

  // split the packed transform into the two real-input spectra
  for (int k = threadIdx.x; k < g.bins; k += blockDim.x) {
    const float2 zk = Z[k];
    const float2 zn = Z[k == 0 ? 0 : g.nfft - k];
    const float2 xa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
    const float2 xb = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
    const size_t oa = (((size_t)b * g.frames + f0) * g.bins + k) * g.nch + c;
    const size_t ob = oa + (size_t)g.bins * g.nch;
    if (g.out_c64) {
      reinterpret_cast<float2*>(g.out_c64)[oa] = xa;
      if (has_b) reinterpret_cast<float2*>(g.out_c64)[ob] = xb;
    }
    const float ma = sqrtf(xa.x * xa.x + xa.y * xa.y);
    const float mb = sqrtf(xb.x * xb.x + xb.y * xb.y);
    if (g.out_mag) {
      g.out_mag[oa] = ma;
      if (has_b) g.out_mag[ob] = mb;
    }
    if (g.mel_fb) {
      mag[k] = ma;
      mag[g.bins + k] = mb;
    }
  }
  if (!g.mel_fb) return;
  __syncthreads();

  // mel filterbank (sparse triangular rows) + dB normalisation, both frames
  for (int t = threadIdx.x; t < 2 * g.nmels; t += blockDim.x) {
    const int f = t / g.nmels, m = t - f * g.nmels;
    if (f == 1 && !has_b) continue;
    int lo = 0, hi = g.bins;
    if (g.mel_range) {
      lo = __ldg(g.mel_range + 2 * m);
      hi = __ldg(g.mel_range + 2 * m + 1);
    }
    const float* wrow = g.mel_fb + (size_t)m * g.bins;
    const float* mrow = mag + f * g.bins;
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(mrow[k], __ldg(wrow + k), acc);
    const float db = 20.f * log10f(fmaxf(g.min_level, acc)) - g.ref_db;
    const float v = fminf(fmaxf((db - g.min_db) / -g.min_db, 0.f), 1.f);
    g.out_mel[(((size_t)b * g.frames + f0 + f) * g.nmels + m) * g.nch + c] = v;
  }
}

__global__ void mel_ranges_kernel(const float* __restrict__ fb, int nmels, int bins, int* out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nmels) return;
  int lo = bins, hi = 0;
  for (int k = 0; k < bins; ++k) {
    if (fb[(size_t)m * bins + k] != 0.f) {
      if (k < lo) lo = k;
      hi = k + 1;
    }
  }
  if (hi == 0) lo = 0;
  out[2 * m] = lo;
  out[2 * m + 1] = hi;
}

// y[r,n] = sum_k f(x[r,k]) w[n,k]; 64x64 output tile, 16-wide k slab, 4x4 per thread
__global__ void __launch_bounds__(256) matmul_lastdim_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             float* __restrict__ y, long rows,
                                                             int K, int N, int pow10, float db_scale,
                                                             float db_offset) {
  __shared__ float xs[16][64 + 4];
  __shared__ float ws[16][64 + 4];
  const long r0 = (long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      float v = 0.f;
      if (r0 + rr < rows && k0 + kk < K) {
        v = __ldg(x + (r0 + rr) * K + k0 + kk);
        if (pow10) v = exp10f((v * db_scale + db_offset) * 0.05f);
      }
      xs[kk][rr] = v;
      float u = 0.f;
      if (n0 + rr < N && k0 + kk < K) u = __ldg(w + (size_t)(n0 + rr) * K + k0 + kk);
      ws[kk][rr] = u;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = xs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = ws[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long r = r0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N) y[r * N + n] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Inversion: windowed inverse frames of two consecutive frames per CTA (z = Xa + i Xb, both
// Hermitian, so ifft(z) = xa + i xb), optionally preceded by a Griffin-Lim phase update
//   S = stft(wave) ; angle = S / |S| ; X = mag * angle            (advoc/spectral.py:304-307)
// and a separate atomics-free overlap-add.  lws istft with perfectrec=False: synthesis window ==
// analysis window, output length (frames-1)*hop + nfft (advoc/spectral.py:300-311).
// ---------------------------------------------------------------------------------------------
struct IstftArgs {
  const float2* spec;   // [batch, frames, bins] complex, or null (Griffin-Lim mode)
  const float* mag;     // [batch, frames, bins] (Griffin-Lim mode)
  const float* wave;    // [batch, nwave] current estimate (Griffin-Lim mode)
  const float* window;
  const float2* tw;
  float* frames_out;    // [batch, frames, nfft]
  int batch, frames, bins, nfft, nhop;
  long nwave;
};

__global__ void __launch_bounds__(256) istft_pair_kernel(const IstftArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* bufa = reinterpret_cast<float2*>(smem_raw);
  float2* bufb = bufa + g.nfft;
  const int pairs = (g.frames + 1) >> 1;
  const int pair = blockIdx.x % pairs;
  const int b = blockIdx.x / pairs;
  const int f0 = pair * 2;
  const bool has_b = (f0 + 1) < g.frames;
  const size_t sbase = ((size_t)b * g.frames + f0) * g.bins;

  if (g.spec == nullptr) {
    // Griffin-Lim: analysis of the current waveform estimate (lws stft: no end padding needed,
    // the estimate is exactly (frames-1)*hop + nfft long)
    const float* x = g.wave + (size_t)b * g.nwave;
    for (int i = threadIdx.x; i < g.nfft; i += blockDim.x) {
      const float w = __ldg(g.window + i);
      const long ia = (long)f0 * g.nhop + i, ib = ia + g.nhop;
      const float va = ia < g.nwave ? __ldg(x + ia) * w : 0.f;
      const float vb = (has_b && ib < g.nwave) ? __ldg(x + ib) * w : 0.f;
      bufa[i] = make_float2(va, vb);
    }
    __syncthreads();
    float2* Z = stockham_fft<false>(bufa, bufb, g.tw, g.nfft);
    float2* O = (Z == bufa) ? bufb : bufa;
    // split the two spectra, replace the magnitude, re-pack z = Xa + i Xb over the full circle
    for (int k = threadIdx.x; k < g.bins; k += blockDim.x) {
      const float2 zk = Z[k];
      const float2 zn = Z[k == 0 ? 0 : g.nfft - k];
      float2 xa = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      float2 xb = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
      const float na = sqrtf(xa.x * xa.x + xa.y * xa.y), nb = sqrtf(xb.x * xb.x + xb.y * xb.y);
      const float ma = __ldg(g.mag + sbase + k);
      const float mb = has_b ? __ldg(g.mag + sbase + g.bins + k) : 0.f;
      xa = na > 0.f ? make_float2(xa.x / na * ma, xa.y / na * ma) : make_float2(ma, 0.f);   // angle(0) = 0
      xb = nb > 0.f ? make_float2(xb.x / nb * mb, xb.y / nb * mb) : make_float2(mb, 0.f);
      if (k == 0 || k == g.nfft / 2) { xa.y = 0.f; xb.y = 0.f; }   // irfft ignores these imaginary parts
      O[k] = make_float2(xa.x - xb.y, xa.y + xb.x);
      if (k != 0 && k != g.nfft / 2) O[g.nfft - k] = make_float2(xa.x + xb.y, -xa.y + xb.x);  // conj(Xa) + i conj(Xb)
    }
    __syncthreads();
    float2* other = (O == bufa) ? bufb : bufa;
    bufa = O; bufb = other;
  } else {
    for (int k = threadIdx.x; k < g.bins; k += blockDim.x) {
      float2 xa = __ldg(g.spec + sbase + k);
      float2 xb = has_b ? __ldg(g.spec + sbase + g.bins + k) : make_float2(0.f, 0.f);
      if (k == 0 || k == g.nfft / 2) { xa.y = 0.f; xb.y = 0.f; }
      bufa[k] = make_float2(xa.x - xb.y, xa.y + xb.x);
      if (k != 0 && k != g.nfft / 2) bufa[g.nfft - k] = make_float2(xa.x + xb.y, -xa.y + xb.x);
    }
    __syncthreads();
  }
  const float2* Y = stockham_fft<true>(bufa, bufb, g.tw, g.nfft);
  const float scale = 1.f / (float)g.nfft;
  float* oa = g.frames_out + ((size_t)b * g.frames + f0) * g.nfft;
  for (int i = threadIdx.x; i < g.nfft; i += blockDim.x) {
    const float w = __ldg(g.window + i) * scale;
    oa[i] = Y[i].x * w;
    if (has_b) oa[g.nfft + i] = Y[i].y * w;
  }
}

// out[b, i] = sum over frames m with 0 <= i - m*hop < nfft of frames[b, m, i - m*hop]
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ fr, float* __restrict__ out,
                                                          int batch, int frames, int nfft, int nhop, long nout) {
  const long total = (long)batch * nout;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long b = t / nout, i = t - b * nout;
    long m_hi = i / nhop;
    if (m_hi > frames - 1) m_hi = frames - 1;
    long m_lo = (i - nfft + nhop) / nhop;   // smallest m with i - m*hop < nfft
    if (i - nfft + 1 <= 0) m_lo = 0;
    float s = 0.f;
    for (long m = m_lo; m <= m_hi; ++m) {
      const long j = i - m * nhop;
      if (j >= 0 && j < nfft) s += __ldg(fr + ((size_t)b * frames + m) * nfft + j);
    }
    out[t] = s;
  }
}

int launch_stft(StftArgs& g, void* stream) {
  if (g.frames == 0 || g.batch == 0 || g.nch == 0) return ADVOC_OK;
  const bool pow2 = (g.nfft & (g.nfft - 1)) == 0;
  const size_t smem = (size_t)2 * g.nfft * sizeof(float2) + (g.mel_fb ? 2 * g.bins * sizeof(float) : 0);
  ADVOC_REQUIRE(smem <= 200 * 1024, ADVOC_UNSUPPORTED, "nfft %d too large for the smem FFT", g.nfft);
  const long grid = (long)g.batch * g.nch * ((g.frames + 1) / 2);
  ADVOC_REQUIRE(grid < 2147483647L, ADVOC_BAD_SHAPE, "too many frames for one launch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pow2) {
    if (smem > 48 * 1024)
      ADVOC_CHECK_CUDA(cudaFuncSetAttribute(stft_pair_kernel<true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_pair_kernel<true><<<(unsigned)grid, 256, smem, st>>>(g);
  } else {
    if (smem > 48 * 1024)
      ADVOC_CHECK_CUDA(cudaFuncSetAttribute(stft_pair_kernel<false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_pair_kernel<false><<<(unsigned)grid, 256, smem, st>>>(g);
  }
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace
}  // namespace advoc

using namespace advoc;

extern "C" int advoc_num_frames(int nsamps, int nfft, int nhop, int pad_end) {
  if (nsamps <= 0 || nhop <= 0 || nfft <= 0) return 0;
  if (pad_end == 1) return (nsamps + nhop - 1) / nhop;
  if (pad_end == 2) return nsamps >= nfft ? (nsamps - nfft) / nhop + 1 : 0;   // tf.contrib.signal.stft(pad_end=False)
  int d = nsamps - nfft;
  int m = (d > 0 ? (d + nhop - 1) / nhop : -((-d) / nhop)) + 1;
  return m < 1 ? 1 : m;
}

static int check_stft_common(const float* d_wav, int batch, int nsamps, int nch, int nfft, int nhop,
                             const float* d_window, const float* d_twiddle) {
  ADVOC_REQUIRE(batch >= 0 && nsamps >= 0 && nch >= 1, ADVOC_BAD_SHAPE, "bad wav shape [%d,%d,1,%d]",
                batch, nsamps, nch);
  ADVOC_REQUIRE(nfft >= 8 && (nfft % 2) == 0 && nhop >= 1, ADVOC_BAD_ARG, "bad nfft/nhop %d/%d", nfft,
                nhop);
  ADVOC_REQUIRE(d_window && d_twiddle, ADVOC_BAD_ARG, "window/twiddle pointer is NULL");
  ADVOC_REQUIRE(d_wav || batch * nsamps == 0, ADVOC_BAD_ARG, "wav pointer is NULL");
  return ADVOC_OK;
}

extern "C" int advoc_stft_f32(const float* d_wav, int batch, int nsamps, int nch, int nfft, int nhop,
                              int pad_end, const float* d_window, const float* d_twiddle,
                              float* d_out_c64, float* d_out_mag, void* stream) {
  int st = check_stft_common(d_wav, batch, nsamps, nch, nfft, nhop, d_window, d_twiddle);
  if (st) return st;
  ADVOC_REQUIRE(d_out_c64 || d_out_mag, ADVOC_BAD_ARG, "no output requested");
  StftArgs g = {};
  g.wav = d_wav; g.window = d_window; g.tw = reinterpret_cast<const float2*>(d_twiddle);
  g.out_c64 = d_out_c64; g.out_mag = d_out_mag;
  g.batch = batch; g.nsamps = nsamps; g.nch = nch; g.nfft = nfft; g.nhop = nhop;
  g.frames = advoc_num_frames(nsamps, nfft, nhop, pad_end);
  g.bins = nfft / 2 + 1;
  return launch_stft(g, stream);
}

extern "C" int advoc_mel_ranges(const float* d_mel_fb, int nmels, int bins, int* d_ranges,
                                void* stream) {
  ADVOC_REQUIRE(d_mel_fb && d_ranges && nmels > 0 && bins > 0, ADVOC_BAD_ARG, "bad mel_ranges args");
  mel_ranges_kernel<<<(nmels + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_mel_fb, nmels, bins, d_ranges);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_melspec_f32(const float* d_wav, int batch, int nsamps, int nch, int nfft,
                                 int nhop, const float* d_window, const float* d_twiddle,
                                 const float* d_mel_fb, const int* d_mel_ranges, int nmels,
                                 float min_level_db, float ref_level_db, float* d_out,
                                 void* stream) {
  int st = check_stft_common(d_wav, batch, nsamps, nch, nfft, nhop, d_window, d_twiddle);
  if (st) return st;
  ADVOC_REQUIRE(d_mel_fb && d_out && nmels >= 1, ADVOC_BAD_ARG, "mel filterbank/out is NULL");
  ADVOC_REQUIRE(min_level_db < 0.f, ADVOC_BAD_ARG, "min_level_db must be negative");
  StftArgs g = {};
  g.wav = d_wav; g.window = d_window; g.tw = reinterpret_cast<const float2*>(d_twiddle);
  g.mel_fb = d_mel_fb; g.mel_range = d_mel_ranges; g.out_mel = d_out; g.nmels = nmels;
  g.batch = batch; g.nsamps = nsamps; g.nch = nch; g.nfft = nfft; g.nhop = nhop;
  g.frames = advoc_num_frames(nsamps, nfft, nhop, 1);
  g.bins = nfft / 2 + 1;
  g.min_db = min_level_db; g.ref_db = ref_level_db;
  g.min_level = (float)exp((double)min_level_db / 20.0 * log(10.0));
  return launch_stft(g, stream);
}

extern "C" int advoc_matmul_lastdim_f32(const float* d_x, const float* d_w, float* d_y, long rows,
                                        int K, int N, int pow10_scale, float min_level_db, float ref_level_db,
                                        void* stream) {
  ADVOC_REQUIRE(rows >= 0 && K >= 1 && N >= 1, ADVOC_BAD_SHAPE, "bad matmul shape");
  if (rows == 0) return ADVOC_OK;
  ADVOC_REQUIRE(d_x && d_w && d_y, ADVOC_BAD_ARG, "NULL pointer");
  dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((N + 63) / 64));
  // x in [0,1] -> dB: x * (-min_db) + min_db + ref_db   (advoc/spectral.py:367-369)
  matmul_lastdim_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_x, d_w, d_y, rows, K, N, pow10_scale, -min_level_db, min_level_db + ref_level_db);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

static int check_inv_common(int batch, int frames, int nfft, int nhop, const float* d_window, const float* d_twiddle,
                            float* d_frames) {
  ADVOC_REQUIRE(batch >= 0 && frames >= 0, ADVOC_BAD_SHAPE, "bad spectrogram shape [%d,%d]", batch, frames);
  ADVOC_REQUIRE(nfft >= 8 && (nfft & (nfft - 1)) == 0 && nhop >= 1 && nhop <= nfft, ADVOC_UNSUPPORTED,
                "inverse transform needs a power-of-two nfft (got %d) and 1 <= nhop <= nfft", nfft);
  ADVOC_REQUIRE(d_window && d_twiddle && d_frames, ADVOC_BAD_ARG, "NULL window / twiddle / frames pointer");
  ADVOC_REQUIRE((size_t)2 * nfft * sizeof(float2) <= 200 * 1024, ADVOC_UNSUPPORTED, "nfft too large");
  return ADVOC_OK;
}

static int launch_istft(IstftArgs& g, void* stream) {
  if (g.batch == 0 || g.frames == 0) return ADVOC_OK;
  const size_t smem = (size_t)2 * g.nfft * sizeof(float2);
  if (smem > 48 * 1024)
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(istft_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long grid = (long)g.batch * ((g.frames + 1) / 2);
  ADVOC_REQUIRE(grid < 2147483647L, ADVOC_BAD_SHAPE, "too many frames for one launch");
  istft_pair_kernel<<<(unsigned)grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(g);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_istft_frames_f32(const float* d_spec_c64, int batch, int frames, int nfft, int nhop,
                                      const float* d_window, const float* d_twiddle, float* d_frames,
                                      void* stream) {
  int st = check_inv_common(batch, frames, nfft, nhop, d_window, d_twiddle, d_frames);
  if (st) return st;
  ADVOC_REQUIRE(d_spec_c64 || batch * frames == 0, ADVOC_BAD_ARG, "spectrum pointer is NULL");
  IstftArgs g = {};
  g.spec = reinterpret_cast<const float2*>(d_spec_c64);
  g.window = d_window; g.tw = reinterpret_cast<const float2*>(d_twiddle); g.frames_out = d_frames;
  g.batch = batch; g.frames = frames; g.bins = nfft / 2 + 1; g.nfft = nfft; g.nhop = nhop;
  return launch_istft(g, stream);
}

extern "C" int advoc_griffin_lim_iter_f32(const float* d_wave, long nwave, const float* d_mag, int batch,
                                          int frames, int nfft, int nhop, const float* d_window,
                                          const float* d_twiddle, float* d_frames, void* stream) {
  int st = check_inv_common(batch, frames, nfft, nhop, d_window, d_twiddle, d_frames);
  if (st) return st;
  ADVOC_REQUIRE((d_wave && d_mag) || batch * frames == 0, ADVOC_BAD_ARG, "wave / mag pointer is NULL");
  ADVOC_REQUIRE(nwave == (long)(frames - 1) * nhop + nfft || frames == 0, ADVOC_BAD_SHAPE,
                "waveform estimate must be (frames-1)*hop + nfft samples long");
  IstftArgs g = {};
  g.mag = d_mag; g.wave = d_wave; g.nwave = nwave;
  g.window = d_window; g.tw = reinterpret_cast<const float2*>(d_twiddle); g.frames_out = d_frames;
  g.batch = batch; g.frames = frames; g.bins = nfft / 2 + 1; g.nfft = nfft; g.nhop = nhop;
  return launch_istft(g, stream);
}

extern "C" int advoc_overlap_add_f32(const float* d_frames, int batch, int frames, int nfft, int nhop,
                                     float* d_out, void* stream) {
  ADVOC_REQUIRE(batch >= 0 && frames >= 0 && nfft >= 1 && nhop >= 1, ADVOC_BAD_SHAPE, "bad overlap-add shape");
  if (batch == 0 || frames == 0) return ADVOC_OK;
  ADVOC_REQUIRE(d_frames && d_out, ADVOC_BAD_ARG, "NULL pointer");
  const long nout = (long)(frames - 1) * nhop + nfft;
  const long total = (long)batch * nout;
  const long blocks = (total + 255) / 256;
  overlap_add_kernel<<<(unsigned)(blocks < 65535 * 16 ? blocks : 65535 * 16), 256, 0,
                       reinterpret_cast<cudaStream_t>(stream)>>>(d_frames, d_out, batch, frames, nfft, nhop, nout);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}
