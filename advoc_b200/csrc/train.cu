// Backward-pass / optimiser kernels of the train step that are not convolutions-as-GEMM on the
// tensor cores: filter gradients (CUDA-core split-K GEMM + the thin-end special cases), bias
// gradients, the PatchGAN head (conv to one channel), losses with their gradient seeds, TF1 Adam.
// replaces: the gradient graph `opt.minimize` builds for models/advoc/advoc_model.py:238-257.
#include "epilogue.cuh"

namespace advoc {

int check_conv_desc(const advoc_conv_desc* d);
bool wgrad_tc_eligible(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small);
int wgrad_tc(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small, float* dw,
             void* stream);
bool wgrad_thin_tc_eligible(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small);
int wgrad_thin_tc(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small, float* dw,
                  void* stream);
int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* out);

namespace {

// ---------------------------------------------------------------------------------------------
// generic filter gradient: dW[tap][cb][cs] += sum_pix big[pix*s - pad + tap, cb] * small[pix, cs]
// 64x64 tile of (cb, cs) per CTA for one tap and one slice of the pixel range (split-K), fp32.
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* big;    // [N, H, W, ldb]
  const float* small;  // [N, Ho, Wo, lds]
  float* dw;           // [taps][Cb][Cs]
  int N, H, W, ldb, Cb, Ho, Wo, lds, Cs;
  int kh, kw, sh, sw, pt, pl;
  long chunk;          // pixels per split
};

constexpr int WT = 64, WK = 16;

__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradArgs a) {
  __shared__ float As[WK][WT + 4];
  __shared__ float Bs[WK][WT + 4];
  const int cb0 = blockIdx.x * WT, cs0 = blockIdx.y * WT;
  const int taps = a.kh * a.kw;
  const int tap = blockIdx.z % taps;
  const long split = blockIdx.z / taps;
  const int kh = tap / a.kw, kw = tap - kh * a.kw;
  const long M = (long)a.N * a.Ho * a.Wo;
  const long p0 = split * a.chunk;
  const long p1 = p0 + a.chunk < M ? p0 + a.chunk : M;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (long pb = p0; pb < p1; pb += WK) {
    // 16 pixels x 64 channels of each operand; thread -> (pixel = tid/16, 4 channels = (tid%16)*4)
    const int pp = threadIdx.x >> 4, cq = (threadIdx.x & 15) * 4;
    const long pix = pb + pp;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (pix < p1) {
      const int ow = (int)(pix % a.Wo);
      const long r = pix / a.Wo;
      const int oh = (int)(r % a.Ho);
      const long img = r / a.Ho;
      const int ih = oh * a.sh - a.pt + kh, iw = ow * a.sw - a.pl + kw;
      if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) {
        const float* bp = a.big + (((size_t)img * a.H + ih) * a.W + iw) * a.ldb + cb0 + cq;
        if (cb0 + cq + 3 < a.Cb) av = __ldg(reinterpret_cast<const float4*>(bp));
        else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j) if (cb0 + cq + j < a.Cb) t[j] = __ldg(bp + j);
          av = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      const float* sp = a.small + (size_t)pix * a.lds + cs0 + cq;
      if (cs0 + cq + 3 < a.Cs) bv = __ldg(reinterpret_cast<const float4*>(sp));
      else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j) if (cs0 + cq + j < a.Cs) t[j] = __ldg(sp + j);
        bv = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    As[pp][cq] = av.x; As[pp][cq + 1] = av.y; As[pp][cq + 2] = av.z; As[pp][cq + 3] = av.w;
    Bs[pp][cq] = bv.x; Bs[pp][cq + 1] = bv.y; Bs[pp][cq + 2] = bv.z; Bs[pp][cq + 3] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      float x[4], y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cb = cb0 + ty * 4 + i;
    if (cb >= a.Cb) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cs = cs0 + tx + 16 * j;
      if (cs < a.Cs) atomicAdd(a.dw + ((size_t)tap * a.Cb + cb) * a.Cs + cs, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// thin filter gradient (k4): one side of the layer has CT in {1,2} channels.
//   kWideOnBig = false: dW[tap][ct][c] += sum_o thin[o*s - pad + tap, ct] * wide[o, c]
//                       (encoder_1, discriminator layer_1, decoder_1: thin = conv-input side)
//   kWideOnBig = true : dW[tap][c][ct] += sum_i wide[i, c] * thin[(i + pad - tap)/s, ct]
//                       (discriminator layer_5: thin = conv-output side), CT == 1 only
// one thread = one pixel of the wide tensor x 4 channels, register accumulators over a
// grid-stride loop, then shared -> global atomics.
// ---------------------------------------------------------------------------------------------
struct WgradThinArgs {
  const float* thin;
  const float* wide;
  float* dw;
  int N, Ht, Wt, ldt;       // thin tensor
  int Hw, Ww, ldw, C;       // wide tensor
  int sh, sw, pt, pl;
};

template <int CT, bool kWideOnBig>
__global__ void __launch_bounds__(256) wgrad_thin_kernel(const WgradThinArgs a) {
  extern __shared__ float dws[];  // [16][CT][C] (or [16][C] when kWideOnBig)
  for (int i = threadIdx.x; i < 16 * CT * a.C; i += blockDim.x) dws[i] = 0.f;
  __syncthreads();
  const int groups = a.C >> 2;
  const int ppb = 256 / groups;
  const int cg = threadIdx.x % groups, n = cg * 4;
  const unsigned pix_in_img = a.Hw * a.Ww;
  const long npix = (long)a.N * pix_in_img;
  float acc[16][CT][4];
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][c][j] = 0.f;
  for (long pix = (long)blockIdx.x * ppb + threadIdx.x / groups; pix < npix; pix += (long)gridDim.x * ppb) {
    const unsigned img = (unsigned)(pix / pix_in_img);
    const unsigned rem = (unsigned)(pix - (long)img * pix_in_img);
    const int ph = rem / a.Ww, pw = rem - ph * a.Ww;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(a.wide + (size_t)pix * a.ldw + n));
    const float* tb = a.thin + (size_t)img * a.Ht * a.Wt * a.ldt;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      int th;
      if (!kWideOnBig) {
        th = ph * a.sh - a.pt + kh;
      } else {
        const int q = ph + a.pt - kh;
        if (a.sh == 1) th = q;                                   // stride 1 (PatchGAN head): no division
        else th = (q >= 0 && q % a.sh == 0) ? q / a.sh : -1;
      }
      if (th < 0 || th >= a.Ht) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        int tw;
        if (!kWideOnBig) {
          tw = pw * a.sw - a.pl + kw;
        } else {
          const int q = pw + a.pl - kw;
          if (a.sw == 1) tw = q;
          else tw = (q >= 0 && q % a.sw == 0) ? q / a.sw : -1;
        }
        if (tw < 0 || tw >= a.Wt) continue;
        const float* tp = tb + ((size_t)th * a.Wt + tw) * a.ldt;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const float tv = __ldg(tp + c);
          acc[kh * 4 + kw][c][0] = fmaf(tv, wv.x, acc[kh * 4 + kw][c][0]);
          acc[kh * 4 + kw][c][1] = fmaf(tv, wv.y, acc[kh * 4 + kw][c][1]);
          acc[kh * 4 + kw][c][2] = fmaf(tv, wv.z, acc[kh * 4 + kw][c][2]);
          acc[kh * 4 + kw][c][3] = fmaf(tv, wv.w, acc[kh * 4 + kw][c][3]);
        }
      }
    }
  }
  // lanes l, l+groups, l+2*groups, ... of a warp hold the same channel group: fold them with
  // shuffles first so that only `groups` lanes per warp touch the shared accumulators
  const bool fold = groups < 32 && (groups & (groups - 1)) == 0;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[t][c][j];
        if (fold) {
          for (int o = 16; o >= groups; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        if (!fold || lane < groups) atomicAdd(&dws[(t * CT + c) * a.C + n + j], v);
      }
  __syncthreads();
  // dW layouts: thin-on-big  -> [tap][ct][C]   (== dws)
  //             wide-on-big  -> [tap][C][ct=1] (== dws for CT == 1)
  for (int i = threadIdx.x; i < 16 * CT * a.C; i += blockDim.x) atomicAdd(a.dw + i, dws[i]);
}

// ---------------------------------------------------------------------------------------------
// thin filter gradient, shared-memory tiled (thin side = conv input with CT in {1,2} channels):
//   dW[tap][ct][c] += sum_o thin[o*s - pad + tap, ct] * wide[o, c]
// A block walks row segments of 64 output pixels: the segment of the wide tensor ([64][C]) and the
// 4 input rows it touches are staged in shared memory with coalesced loads, every thread owns one
// (tap, ct, 4-channel) accumulator (or two when there are 512 of them) over all the segments of
// the block and adds it to dW once at the end.  The register-tiled kernel above issues ~17
// dependent global loads per pixel and thread and runs 10-15x off the memory roofline.
// ---------------------------------------------------------------------------------------------
constexpr int WT2_PX = 64;

template <int CT, int C4, int KS = 4>
__global__ void __launch_bounds__(256) wgrad_thin_tiled_kernel(const WgradThinArgs a) {
  constexpr int ITEMS = KS * KS * CT * C4;             // (tap, ct, channel quad) accumulators
  constexpr int PER = ITEMS >= 256 ? (ITEMS + 255) / 256 : 1;  // accumulators per thread
  constexpr int SPLIT = ITEMS >= 256 ? 1 : 256 / ITEMS;  // pixel splits when there are fewer items than threads
  constexpr int TW = (WT2_PX - 1) * 2 + KS;            // thin columns per segment at stride <= 2
  __shared__ float4 wide_s[WT2_PX * C4];
  __shared__ float thin_s[KS][TW][CT];
  const int segs = (a.Ww + WT2_PX - 1) / WT2_PX;
  const long units = (long)a.N * a.Hw * segs;
  float4 acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int item0 = threadIdx.x % (ITEMS < 256 ? ITEMS : 256);
  const int split = threadIdx.x / (ITEMS < 256 ? ITEMS : 256);
  for (long u = blockIdx.x; u < units; u += gridDim.x) {
    const int seg = (int)(u % segs);
    const long r = u / segs;
    const int oh = (int)(r % a.Hw);
    const long img = r / a.Hw;
    const int ow0 = seg * WT2_PX;
    __syncthreads();   // the previous segment has been consumed
    // wide segment, zero beyond the row end
    for (int i = threadIdx.x; i < WT2_PX * C4; i += 256) {
      const int p = i / C4, q = i - p * C4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ow0 + p < a.Ww)
        v = __ldg(reinterpret_cast<const float4*>(a.wide + (((size_t)img * a.Hw + oh) * a.Ww + ow0 + p) * a.ldw) + q);
      wide_s[i] = v;
    }
    // the 4 thin rows, zero outside the image
    const int tw0 = ow0 * a.sw - a.pl;
    for (int i = threadIdx.x; i < KS * TW * CT; i += 256) {
      const int kh = i / (TW * CT);
      const int rem = i - kh * (TW * CT);
      const int col = rem / CT, ct = rem - col * CT;
      const int th = oh * a.sh - a.pt + kh, tw = tw0 + col;
      float v = 0.f;
      if (th >= 0 && th < a.Ht && tw >= 0 && tw < a.Wt)
        v = __ldg(a.thin + (((size_t)img * a.Ht + th) * a.Wt + tw) * a.ldt + ct);
      thin_s[kh][col][ct] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int item = item0 + 256 * k;
      if (item >= ITEMS) continue;
      const int q = item % C4;
      const int tc = item / C4;          // tap * CT + ct
      const int ct = tc % CT, tap = tc / CT;
      const int kh = tap / KS, kw = tap - kh * KS;
      const int p0 = split * (WT2_PX / SPLIT), p1 = p0 + WT2_PX / SPLIT;
      float4 s = acc[k];
#pragma unroll 8
      for (int p = p0; p < p1; ++p) {
        const float tv = thin_s[kh][p * a.sw + kw][ct];
        const float4 wv = wide_s[p * C4 + q];
        s.x = fmaf(tv, wv.x, s.x);
        s.y = fmaf(tv, wv.y, s.y);
        s.z = fmaf(tv, wv.z, s.z);
        s.w = fmaf(tv, wv.w, s.w);
      }
      acc[k] = s;
    }
  }
  // dW layout [tap][ct][C]: item index * 4 is the flat offset
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    if (item0 + 256 * k >= ITEMS) continue;
    float* dst = a.dw + (size_t)(item0 + 256 * k) * 4;
    atomicAdd(dst, acc[k].x);
    atomicAdd(dst + 1, acc[k].y);
    atomicAdd(dst + 2, acc[k].z);
    atomicAdd(dst + 3, acc[k].w);
  }
}

// ---------------------------------------------------------------------------------------------
// bias gradient: db[c] += sum_pix dy[pix, c]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, int ld, long pixels, int C,
                                                     float* __restrict__ db, int vec) {
  extern __shared__ float part[];  // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  if (vec) {
    const int groups = C >> 2;
    const int ppb = blockDim.x / groups > 0 ? blockDim.x / groups : 1;
    const int cg = threadIdx.x % groups;
    if ((int)(threadIdx.x / groups) < ppb) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (long p = (long)blockIdx.x * ppb + threadIdx.x / groups; p < pixels; p += (long)gridDim.x * ppb) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dy + (size_t)p * ld + cg * 4));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      atomicAdd(&part[cg * 4], s.x); atomicAdd(&part[cg * 4 + 1], s.y);
      atomicAdd(&part[cg * 4 + 2], s.z); atomicAdd(&part[cg * 4 + 3], s.w);
    }
  } else {
    for (int c = 0; c < C; ++c) {
      float s = 0.f;
      for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long)gridDim.x * blockDim.x)
        s += __ldg(dy + (size_t)p * ld + c);
      atomicAdd(&part[c], s);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(db + i, part[i]);
}

// ---------------------------------------------------------------------------------------------
// conv to ONE output channel (PatchGAN head, advoc_model.py:199-202): a warp per output pixel,
// lanes across input channels (16-byte coalesced loads), butterfly reduction.
// ---------------------------------------------------------------------------------------------
struct ToOneConvArgs {
  const float* x;
  const float* w;  // HWIO [taps][Cin][1]
  int N, H, W, ldx, Cin, Ho, Wo, kh, kw, sh, sw, pt, pl;
  EpiDev epi;
};

__global__ void __launch_bounds__(256) conv_to_one_kernel(const ToOneConvArgs a) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long npix = (long)a.N * a.Ho * a.Wo;
  if (warp >= npix) return;
  const int ow = (int)(warp % a.Wo);
  const long r = warp / a.Wo;
  const int oh = (int)(r % a.Ho);
  const long img = r / a.Ho;
  float acc = 0.f;
  for (int kh = 0; kh < a.kh; ++kh) {
    const int ih = oh * a.sh - a.pt + kh;
    if (ih < 0 || ih >= a.H) continue;
    for (int kw = 0; kw < a.kw; ++kw) {
      const int iw = ow * a.sw - a.pl + kw;
      if (iw < 0 || iw >= a.W) continue;
      const float* xp = a.x + (((size_t)img * a.H + ih) * a.W + iw) * a.ldx;
      const float* wp = a.w + (size_t)(kh * a.kw + kw) * a.Cin;
      for (int c = lane * 4; c < a.Cin; c += 128) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + c));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + c));
        acc = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc))));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) epi_store(a.epi, (size_t)warp, 0, acc);
}

// ---------------------------------------------------------------------------------------------
// losses (advoc_model.py:238-245) with the gradient seeds autodiff would produce
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;
}

constexpr float kEps = 1e-12f;  // advoc_model.py:8

__global__ void __launch_bounds__(256) gan_logloss_kernel(const float* __restrict__ p_real,
                                                          const float* __restrict__ p_fake, long n, int mode,
                                                          float weight, float* loss, float* dz_real,
                                                          float* dz_fake) {
  __shared__ float red[8];
  const float inv_n = 1.f / (float)n;
  float s = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float pf = __ldg(p_fake + i);
    if (mode == 0) {
      const float pr = __ldg(p_real + i);
      s -= logf(pr + kEps) + logf(1.f - pf + kEps);
      if (dz_real) dz_real[i] = -(pr * (1.f - pr)) / (pr + kEps) * inv_n;
      if (dz_fake) dz_fake[i] = (pf * (1.f - pf)) / (1.f - pf + kEps) * inv_n;
    } else {
      s -= logf(pf + kEps);
      if (dz_fake) dz_fake[i] = -(pf * (1.f - pf)) / (pf + kEps) * inv_n * weight;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, s * inv_n * (mode == 0 ? 1.f : weight));
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ gen, int ldg, int coffg,
                                                      const float* __restrict__ target, long n, float weight,
                                                      float* loss, float* dgen, int accumulate) {
  __shared__ float red[8];
  const float inv_n = 1.f / (float)n;
  float s = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float d = __ldg(gen + (size_t)i * ldg + coffg) - __ldg(target + i);
    s += fabsf(d);
    if (dgen) {
      const float g = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * weight * inv_n;
      dgen[i] = accumulate ? dgen[i] + g : g;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, s * inv_n * weight);
}

// TF1 Adam: epsilon OUTSIDE the bias correction (SURVEY appendix B rule 5)
__global__ void __launch_bounds__(256) adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v, long n,
                                                      float lr_t, float b1, float b2, float eps, float gscale) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// same update with the bias-corrected step size read from device memory, so that a captured CUDA graph of
// the optimiser step stays valid while t advances (the host refreshes the scalar before each replay)
__global__ void __launch_bounds__(256) adam_tf_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v, long n,
                                                          const float* __restrict__ lr_t_ptr, float b1, float b2,
                                                          float eps, float gscale) {
  const float lr_t = __ldg(lr_t_ptr);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int grid_for(long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  const long cap = (long)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

// used by conv_fwd_simt for Cout == 1
bool conv_to_one_eligible(const advoc_conv_desc* d, const float* x, int ldx, const advoc_epilogue* ep) {
  return d->Cout == 1 && d->Cin % 4 == 0 && ldx % 4 == 0 && aligned16(x) && ep->keep_prob >= 1.f &&
         ep->d_out1 == nullptr && ep->store_w == 0;
}

int conv_to_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                void* stream) {
  ToOneConvArgs a = {};
  int st = lower_epilogue(ep, d->Ho, d->Wo, 1, &a.epi);
  if (st) return st;
  a.x = x; a.w = w; a.N = d->N; a.H = d->H; a.W = d->W; a.ldx = ldx; a.Cin = d->Cin; a.Ho = d->Ho; a.Wo = d->Wo;
  a.kh = d->kh; a.kw = d->kw; a.sh = d->sh; a.sw = d->sw; a.pt = d->pad_t; a.pl = d->pad_l;
  const long npix = (long)d->N * d->Ho * d->Wo;
  if (npix == 0) return ADVOC_OK;
  const long blocks = (npix * 32 + 255) / 256;
  ADVOC_REQUIRE(blocks < 2147483647L, ADVOC_BAD_SHAPE, "too many pixels");
  conv_to_one_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc

using namespace advoc;

extern "C" int advoc_conv2d_wgrad(const advoc_conv_desc* d, const float* d_big, int ld_big,
                                  const float* d_small, int ld_small, float* d_dw, void* stream) {
  int st = check_conv_desc(d);
  if (st) return st;
  ADVOC_REQUIRE(d_big && d_small && d_dw, ADVOC_BAD_ARG, "NULL pointer");
  ADVOC_REQUIRE(ld_big >= d->Cin && ld_small >= d->Cout, ADVOC_BAD_SHAPE, "ld smaller than the channel count");
  const long M = (long)d->N * d->Ho * d->Wo;
  if (M == 0) return ADVOC_OK;
  if (wgrad_tc_eligible(d, d_big, ld_big, d_small, ld_small))
    return wgrad_tc(d, d_big, ld_big, d_small, ld_small, d_dw, stream);
  if (wgrad_thin_tc_eligible(d, d_big, ld_big, d_small, ld_small)) {
    st = wgrad_thin_tc(d, d_big, ld_big, d_small, ld_small, d_dw, stream);
    if (st != ADVOC_UNSUPPORTED) return st;     // (no workspace during a graph capture: CUDA-core kernels below)
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool k4 = d->kh == 4 && d->kw == 4;
  if (d->kh == 5 && d->kw == 5 && d->Cin == 1 && d->Cout == 64 && d->sw <= 2 && ld_small % 4 == 0 && aligned16(d_small)) {
    // MelspecGAN conv_0 / upconv_4 filter gradients (one-channel side = conv input)
    WgradThinArgs a = {d_big, d_small, d_dw, d->N, d->H, d->W, ld_big, d->Ho, d->Wo, ld_small, d->Cout,
                       d->sh, d->sw, d->pad_t, d->pad_l};
    const long units = (long)d->N * d->Ho * ((d->Wo + WT2_PX - 1) / WT2_PX);
    const int blocks = (int)(units < (long)sm_count() * 4 ? units : (long)sm_count() * 4);
    wgrad_thin_tiled_kernel<1, 16, 5><<<blocks, 256, 0, s>>>(a);
    count_launch();
    ADVOC_CHECK_CUDA(cudaGetLastError());
    return ADVOC_OK;
  }
  const bool vec_small = d->Cout % 4 == 0 && ld_small % 4 == 0 && aligned16(d_small) && d->Cout <= 256 &&
                         256 % (d->Cout / 4 > 0 ? d->Cout / 4 : 1) == 0;
  const bool vec_big = d->Cin % 4 == 0 && ld_big % 4 == 0 && aligned16(d_big) && d->Cin <= 512 &&
                       256 % (d->Cin / 4 > 0 ? d->Cin / 4 : 1) == 0;
  if (k4 && (d->Cin == 1 || d->Cin == 2) && vec_small) {
    // thin side = conv input (big), wide = conv output (small)
    WgradThinArgs a = {d_big, d_small, d_dw, d->N, d->H, d->W, ld_big, d->Ho, d->Wo, ld_small, d->Cout,
                       d->sh, d->sw, d->pad_t, d->pad_l};
    if ((d->Cout == 32 || d->Cout == 64 || (d->Cout == 128 && d->Cin == 1)) && d->sw <= 2) {
      const long units = (long)d->N * d->Ho * ((d->Wo + WT2_PX - 1) / WT2_PX);
      const int blocks = (int)(units < (long)sm_count() * 4 ? units : (long)sm_count() * 4);
      if (d->Cin == 1 && d->Cout == 32) wgrad_thin_tiled_kernel<1, 8><<<blocks, 256, 0, s>>>(a);
      else if (d->Cin == 1 && d->Cout == 64) wgrad_thin_tiled_kernel<1, 16><<<blocks, 256, 0, s>>>(a);
      else if (d->Cin == 1) wgrad_thin_tiled_kernel<1, 32><<<blocks, 256, 0, s>>>(a);   // regular decoder_1
      else if (d->Cout == 32) wgrad_thin_tiled_kernel<2, 8><<<blocks, 256, 0, s>>>(a);
      else wgrad_thin_tiled_kernel<2, 16><<<blocks, 256, 0, s>>>(a);
    } else {
      const int ppb = 256 / (d->Cout / 4);
      const int blocks = grid_for(M, ppb * 8) < sm_count() * 2 ? grid_for(M, ppb * 8) : sm_count() * 2;
      const size_t smem = (size_t)16 * d->Cin * d->Cout * sizeof(float);
      if (d->Cin == 1) wgrad_thin_kernel<1, false><<<blocks, 256, smem, s>>>(a);
      else wgrad_thin_kernel<2, false><<<blocks, 256, smem, s>>>(a);
    }
  } else if (k4 && d->Cout == 1 && vec_big) {
    // thin side = conv output (small), wide = conv input (big)
    WgradThinArgs a = {d_small, d_big, d_dw, d->N, d->Ho, d->Wo, ld_small, d->H, d->W, ld_big, d->Cin,
                       d->sh, d->sw, d->pad_t, d->pad_l};
    const long Mb = (long)d->N * d->H * d->W;
    const int ppb = 256 / (d->Cin / 4);
    const int blocks = grid_for(Mb, ppb * 8) < sm_count() * 2 ? grid_for(Mb, ppb * 8) : sm_count() * 2;
    const size_t smem = (size_t)16 * d->Cin * sizeof(float);
    wgrad_thin_kernel<1, true><<<blocks, 256, smem, s>>>(a);
  } else {
    WgradArgs a = {d_big, d_small, d_dw, d->N, d->H, d->W, ld_big, d->Cin, d->Ho, d->Wo, ld_small, d->Cout,
                   d->kh, d->kw, d->sh, d->sw, d->pad_t, d->pad_l, 0};
    const int taps = d->kh * d->kw;
    const int tiles = ((d->Cin + WT - 1) / WT) * ((d->Cout + WT - 1) / WT) * taps;
    long splits = ((long)sm_count() * 8 + tiles - 1) / tiles;
    const long max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.chunk = ((M + splits - 1) / splits + WK - 1) / WK * WK;
    splits = (M + a.chunk - 1) / a.chunk;
    ADVOC_REQUIRE((long)taps * splits < 65536, ADVOC_BAD_SHAPE, "too many wgrad splits");
    dim3 grid((d->Cin + WT - 1) / WT, (d->Cout + WT - 1) / WT, (unsigned)(taps * splits));
    wgrad_simt_kernel<<<grid, 256, 0, s>>>(a);
  }
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bias_grad(const float* d_dy, int ld_dy, long pixels, int channels, float* d_dbias,
                               void* stream) {
  ADVOC_REQUIRE(d_dy && d_dbias && channels > 0 && ld_dy >= channels && pixels >= 0, ADVOC_BAD_ARG,
                "bad bias_grad arguments");
  if (pixels == 0) return ADVOC_OK;
  ADVOC_REQUIRE(channels <= 4096, ADVOC_BAD_SHAPE, "too many channels");
  const bool vec = (channels & 3) == 0 && (ld_dy & 3) == 0 && aligned16(d_dy) && channels <= 1024 &&
                   256 % (channels / 4) == 0;
  const int ppb = vec ? 256 / (channels / 4) : 256;
  colsum_kernel<<<grid_for(pixels, ppb * 16), 256, channels * sizeof(float),
                  reinterpret_cast<cudaStream_t>(stream)>>>(d_dy, ld_dy, pixels, channels, d_dbias, vec ? 1 : 0);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_gan_logloss(const float* d_p_real, const float* d_p_fake, long n, int mode, float weight,
                                 float* d_loss, float* d_dlogit_real, float* d_dlogit_fake, void* stream) {
  ADVOC_REQUIRE(d_p_fake && n > 0 && (mode == 0 || mode == 1), ADVOC_BAD_ARG, "bad logloss arguments");
  ADVOC_REQUIRE(mode == 1 || d_p_real, ADVOC_BAD_ARG, "discriminator loss needs p_real");
  gan_logloss_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_p_real, d_p_fake, n, mode, weight, d_loss, d_dlogit_real, d_dlogit_fake);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_l1_loss(const float* d_gen, int ld_gen, int c_off_gen, const float* d_target, long n,
                             float weight, float* d_loss, float* d_dgen, int accumulate, void* stream) {
  ADVOC_REQUIRE(d_gen && d_target && n > 0 && ld_gen >= 1 && c_off_gen >= 0 && c_off_gen < ld_gen, ADVOC_BAD_ARG,
                "bad l1 arguments");
  l1_loss_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_gen, ld_gen, c_off_gen, d_target, n, weight, d_loss, d_dgen, accumulate);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_adam_tf_step_dev(float* d_p, const float* d_g, float* d_m, float* d_v, long n,
                                      const float* d_lr_t, float beta1, float beta2, float eps, float grad_scale,
                                      void* stream) {
  ADVOC_REQUIRE(d_p && d_g && d_m && d_v && d_lr_t && n >= 0, ADVOC_BAD_ARG, "bad adam arguments");
  if (n == 0) return ADVOC_OK;
  adam_tf_dev_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_p, d_g, d_m, d_v, n, d_lr_t, beta1, beta2, eps, grad_scale);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_adam_tf_step(float* d_p, const float* d_g, float* d_m, float* d_v, long n, float lr,
                                  float beta1, float beta2, float eps, long t, float grad_scale, void* stream) {
  ADVOC_REQUIRE(d_p && d_g && d_m && d_v && n >= 0 && t >= 1, ADVOC_BAD_ARG, "bad adam arguments");
  if (n == 0) return ADVOC_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  adam_tf_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_p, d_g, d_m, d_v, n, (float)lr_t, beta1, beta2, eps, grad_scale);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}
