// CUDA-core kernels for the two HBM-bound ends of the generator / discriminator, where a
// 128xN MMA tile cannot be filled:
//   * conv_thin_in_kernel  : Cin in {1,2} -> Cout (generator encoder_1, advoc_model.py:91-94;
//                            discriminator layer_1, :184-187).  K = 16*Cin is tiny, the layer
//                            is bound by WRITING [N,Ho,Wo,Cout] (twice for encoder_1: lrelu for
//                            encoder_2 and relu into the decoder_1 concat slice).
//   * deconv_to_one_kernel : Cin -> 1 channel, k4 s2 (generator decoder_1, :153-158).  Bound by
//                            READING the [N,H/2,W/2,Cin] concat buffer once.
// Both are exact fp32.
#include "epilogue.cuh"

namespace advoc {

namespace {

// ---------------------------------------------------------------------------------------------
// thin-input conv (k4): one thread = one output pixel x 4 consecutive output channels; the 8..128
// threads of a pixel are adjacent, so every store instruction of a warp writes whole 128-byte
// lines.  All index arithmetic is 32-bit with the pixel decomposition done once per thread.
// ---------------------------------------------------------------------------------------------
template <int CIN>
struct ThinArgs {
  const float* x;
  const float* w;  // HWIO [16][CIN][Cout]
  int N, H, W, ldx, Ho, Wo, Cout;
  int sh, sw, pt, pl;
  EpiDev epi;
};

// kT: transposed form y[h,w,:] = sum_taps x[(h+pt-kh)/sh, (w+pl-kw)/sw] * w[tap][:]  (CIN == 1;
// the input gradient of a conv to one channel, e.g. the PatchGAN head)
template <int CIN, bool kT>
__global__ void __launch_bounds__(256) conv_thin_in_kernel(const ThinArgs<CIN> a) {
  extern __shared__ float ws[];  // [16*CIN][Cout]
  for (int i = threadIdx.x; i < 16 * CIN * a.Cout; i += blockDim.x) ws[i] = __ldg(a.w + i);
  __syncthreads();
  const int groups = a.Cout >> 2;
  const int ppb = 256 / groups;                       // pixels per block (groups divides 256)
  const int cg = threadIdx.x % groups;
  const unsigned pix_in_img = a.Ho * a.Wo;
  const long npix = (long)a.N * pix_in_img;
  const EpiDev& e = a.epi;
  const int n = cg * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = __ldg(e.bias + n + j);
  }
  for (long pix = (long)blockIdx.x * ppb + threadIdx.x / groups; pix < npix; pix += (long)gridDim.x * ppb) {
    const unsigned img = (unsigned)(pix / pix_in_img);
    const unsigned rem = (unsigned)(pix - (long)img * pix_in_img);
    const int oh = rem / a.Wo, ow = rem - oh * a.Wo;
    const float* xb = a.x + (size_t)img * a.H * a.W * a.ldx;
    const int ih0 = oh * a.sh - a.pt, iw0 = ow * a.sw - a.pl;
    float acc[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      int ih;
      if (!kT) {
        ih = ih0 + kh;
      } else {
        const int q = oh + a.pt - kh;
        ih = (q >= 0 && q % a.sh == 0) ? q / a.sh : -1;
      }
      if (ih < 0 || ih >= a.H) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        int iw;
        if (!kT) {
          iw = iw0 + kw;
        } else {
          const int q = ow + a.pl - kw;
          iw = (q >= 0 && q % a.sw == 0) ? q / a.sw : -1;
        }
        if (iw < 0 || iw >= a.W) continue;
        const float* xp = xb + ((size_t)ih * a.W + iw) * a.ldx;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float xv = __ldg(xp + c);
          const float4 wv = *reinterpret_cast<const float4*>(ws + ((kh * 4 + kw) * CIN + c) * a.Cout + n);
          acc[0] = fmaf(xv, wv.x, acc[0]);
          acc[1] = fmaf(xv, wv.y, acc[1]);
          acc[2] = fmaf(xv, wv.z, acc[2]);
          acc[3] = fmaf(xv, wv.w, acc[3]);
        }
      }
    }
    float y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = apply_act(acc[j], e.act0, e.alpha);
    if (e.gate) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] *= gate_factor(e, (size_t)pix, n + j);
    }
    float4* dst0 = reinterpret_cast<float4*>(e.out0 + (size_t)pix * e.ld0 + e.coff0 + n);
    if (e.accumulate) {
      const float4 o = *dst0;
      y[0] += o.x; y[1] += o.y; y[2] += o.z; y[3] += o.w;
    }
    if (e.round) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = round_tf32(y[j]);
    }
    *dst0 = make_float4(y[0], y[1], y[2], y[3]);
    if (e.out1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[j] = apply_act(acc[j], e.act1, e.alpha);
        if (e.round) y[j] = round_tf32(y[j]);
      }
      *reinterpret_cast<float4*>(e.out1 + (size_t)pix * e.ld1 + e.coff1 + n) =
          make_float4(y[0], y[1], y[2], y[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k4 s2 transposed conv to ONE channel, two phases per CTA tile of 16 x 32 input positions:
//  (1) per position, the 16 tap dot products  t[kh,kw] = <x[a,b,:], w[kh,kw,0,:]>  -- a half-warp
//      per position, lanes across channels (coalesced 16-byte loads; x is read ONCE), the 16
//      partial sums folded across the 16 lanes with a butterfly -> smem;
//  (2) col2im without atomics: out[2a+ph, 2b+pw] = sum_{dr,dc} t(a+dr, b+dc)[ph+1-2dr, pw+1-2dc]
//      for the positions whose 3x3 neighbourhood lies inside the tile (14 x 30 interior).
// ---------------------------------------------------------------------------------------------
struct ToOneArgs {
  const float* x;  // [N, Hs, Ws, ldx], Cs channels used
  const float* w;  // HWOI [16][1][Cs]
  int N, Hs, Ws, ldx, Cs;
  int tiles_h, tiles_w;
  EpiDev epi;      // stored extent 2*Hs x epi.Ws
};

constexpr int T1_TH = 16, T1_TW = 32;   // tile incl. halo
constexpr int T1_IH = T1_TH - 2, T1_IW = T1_TW - 2;

template <bool kRegWeights>  // Cs == 64: each lane keeps its 16 x 4 filter values in registers
__global__ void __launch_bounds__(256) deconv_to_one_kernel(const ToOneArgs a) {
  __shared__ float tsm[T1_TH * T1_TW][17];  // +1 pad: phase 2 reads down a column of positions
  const int tw_i = blockIdx.x % a.tiles_w;
  const int th_i = (blockIdx.x / a.tiles_w) % a.tiles_h;
  const int img = blockIdx.x / (a.tiles_w * a.tiles_h);
  const int a0 = th_i * T1_IH - 1, b0 = tw_i * T1_IW - 1;  // tile origin (halo included)
  const float* xb = a.x + (size_t)img * a.Hs * a.Ws * a.ldx;
  const int hl = threadIdx.x & 15;   // lane within the half-warp
  const int hw = threadIdx.x >> 4;   // half-warp id, 0..15

  float4 wreg[16];
  if (kRegWeights) {
#pragma unroll
    for (int k = 0; k < 16; ++k) wreg[k] = __ldg(reinterpret_cast<const float4*>(a.w + k * a.Cs + hl * 4));
  }

  // phase 1
  for (int pos = hw; pos < T1_TH * T1_TW; pos += 16) {
    const int ar = a0 + pos / T1_TW, bc = b0 + pos % T1_TW;
    float t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) t[k] = 0.f;
    if (ar >= 0 && ar < a.Hs && bc >= 0 && bc < a.Ws) {
      const float* xp = xb + ((size_t)ar * a.Ws + bc) * a.ldx;
      if (kRegWeights) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + hl * 4));
#pragma unroll
        for (int k = 0; k < 16; ++k)
          t[k] = fmaf(xv.x, wreg[k].x, fmaf(xv.y, wreg[k].y, fmaf(xv.z, wreg[k].z, xv.w * wreg[k].w)));
      } else
      for (int c = hl * 4; c < a.Cs; c += 64) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + c));
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(a.w + k * a.Cs + c));
          t[k] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, t[k]))));
        }
      }
    }
    // fold 16 values across 16 lanes: after the butterfly lane l holds the full sum of tap l
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float keep = (hl & 8) ? t[k + 8] : t[k];
      const float send = (hl & 8) ? t[k] : t[k + 8];
      t[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float keep = (hl & 4) ? t[k + 4] : t[k];
      const float send = (hl & 4) ? t[k] : t[k + 4];
      t[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float keep = (hl & 2) ? t[k + 2] : t[k];
      const float send = (hl & 2) ? t[k] : t[k + 2];
      t[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
      const float keep = (hl & 1) ? t[1] : t[0];
      const float send = (hl & 1) ? t[0] : t[1];
      t[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    tsm[pos][hl] = t[0];  // lane hl ends up with tap index hl (bit order 8,4,2,1 matches)
  }
  __syncthreads();

  // phase 2: 2*T1_IH x 2*T1_IW outputs
  const int OWT = 2 * T1_IW;
  for (int o = threadIdx.x; o < 2 * T1_IH * OWT; o += 256) {
    const int orow = o / OWT, ocol = o - orow * OWT;
    const int ai = 1 + (orow >> 1), bi = 1 + (ocol >> 1);  // position inside the tile
    const int ph = orow & 1, pw = ocol & 1;
    const int ar = a0 + ai, bc = b0 + bi;
    const int oh = 2 * ar + ph, ow = 2 * bc + pw;
    if (ar >= a.Hs || bc >= a.Ws || ow >= a.epi.Ws) continue;
    float acc = 0.f;
#pragma unroll
    for (int jr = 0; jr < 2; ++jr)
#pragma unroll
      for (int jc = 0; jc < 2; ++jc) {
        const int dr = ph - 1 + jr, dc = pw - 1 + jc;
        const int kh = ph + 1 - 2 * dr, kw = pw + 1 - 2 * dc;
        acc += tsm[(ai + dr) * T1_TW + (bi + dc)][kh * 4 + kw];
      }
    const size_t pix = ((size_t)img * a.epi.Hs + oh) * a.epi.Ws + ow;
    epi_store(a.epi, pix, 0, acc);
  }
}

}  // namespace

bool conv_thin_eligible(const advoc_conv_desc* d, const advoc_epilogue* ep) {
  auto ok = [](const float* p, int ld, int co) { return aligned16(p) && ld % 4 == 0 && co % 4 == 0; };
  return (d->Cin == 1 || d->Cin == 2) && d->kh == 4 && d->kw == 4 && d->Cout % 4 == 0 && d->Cout <= 256 &&
         256 % (d->Cout / 4) == 0 && (long)d->Ho * d->Wo < 2147483647L &&
         ep->keep_prob >= 1.f && ep->store_w == 0 && ok(ep->d_out0, ep->ld0, ep->c_off0) &&
         (!ep->d_out1 || ok(ep->d_out1, ep->ld1, ep->c_off1));
}

int conv_thin(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
              void* stream) {
  EpiDev e;
  int st = lower_epilogue(ep, d->Ho, d->Wo, d->Cout, &e);
  if (st) return st;
  const long npix = (long)d->N * d->Ho * d->Wo;
  if (npix == 0) return ADVOC_OK;
  const int ppb = 256 / (d->Cout / 4);
  const long want = (npix + ppb - 1) / ppb;
  const int blocks = (int)(want < (long)sm_count() * 32 ? want : (long)sm_count() * 32);
  const size_t smem = (size_t)16 * d->Cin * d->Cout * sizeof(float);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d->Cin == 1) {
    ThinArgs<1> a = {x, w, d->N, d->H, d->W, ldx, d->Ho, d->Wo, d->Cout, d->sh, d->sw, d->pad_t, d->pad_l, e};
    conv_thin_in_kernel<1, false><<<blocks, 256, smem, s>>>(a);
  } else {
    ThinArgs<2> a = {x, w, d->N, d->H, d->W, ldx, d->Ho, d->Wo, d->Cout, d->sh, d->sw, d->pad_t, d->pad_l, e};
    conv_thin_in_kernel<2, false><<<blocks, 256, smem, s>>>(a);
  }
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

// transposed conv FROM one channel: x [N,Ho,Wo,1] -> y [N,H,W,Cin(desc)]; w [16][Cin(desc)][1]
bool deconv_from_one_eligible(const advoc_conv_desc* d, const advoc_epilogue* ep) {
  auto ok = [](const float* p, int ld, int co) { return aligned16(p) && ld % 4 == 0 && co % 4 == 0; };
  return d->Cout == 1 && d->kh == 4 && d->kw == 4 && d->Cin % 4 == 0 && d->Cin <= 256 &&
         256 % (d->Cin / 4) == 0 && ep->keep_prob >= 1.f && ep->store_w == 0 && ep->d_out1 == nullptr &&
         ok(ep->d_out0, ep->ld0, ep->c_off0);
}

int deconv_from_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                    void* stream) {
  EpiDev e;
  int st = lower_epilogue(ep, d->H, d->W, d->Cin, &e);
  if (st) return st;
  const long npix = (long)d->N * d->H * d->W;
  if (npix == 0) return ADVOC_OK;
  const int ppb = 256 / (d->Cin / 4);
  const long want = (npix + ppb - 1) / ppb;
  const int blocks = (int)(want < (long)sm_count() * 32 ? want : (long)sm_count() * 32);
  // the kernel's "input" is the small side, its "output" the big side
  ThinArgs<1> a = {x, w, d->N, d->Ho, d->Wo, ldx, d->H, d->W, d->Cin, d->sh, d->sw, d->pad_t, d->pad_l, e};
  conv_thin_in_kernel<1, true><<<blocks, 256, (size_t)16 * d->Cin * sizeof(float),
                                 reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

bool deconv_to_one_eligible(const advoc_conv_desc* d, const float* x, int ldx, const advoc_epilogue* ep) {
  return d->Cin == 1 && d->kh == 4 && d->kw == 4 && d->sh == 2 && d->sw == 2 && d->pad_t == 1 && d->pad_l == 1 &&
         d->H == 2 * d->Ho && (d->W == 2 * d->Wo || (ep->accumulate && d->W > 2 * d->Wo)) && d->Cout % 4 == 0 && ldx % 4 == 0 && aligned16(x) && ep->keep_prob >= 1.f && ep->d_out1 == nullptr;
}

int deconv_to_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                  void* stream) {
  ToOneArgs a = {};
  int st = lower_epilogue(ep, d->H, d->W, 1, &a.epi);
  if (st) return st;
  a.x = x; a.w = w; a.N = d->N; a.Hs = d->Ho; a.Ws = d->Wo; a.ldx = ldx; a.Cs = d->Cout;
  a.tiles_h = (a.Hs + T1_IH - 1) / T1_IH;
  a.tiles_w = (a.Ws + T1_IW - 1) / T1_IW;
  const long ctas = (long)a.N * a.tiles_h * a.tiles_w;
  if (ctas == 0) return ADVOC_OK;
  ADVOC_REQUIRE(ctas < 2147483647L, ADVOC_BAD_SHAPE, "too many tiles");
  if (a.Cs == 64)
    deconv_to_one_kernel<true><<<(unsigned)ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  else
    deconv_to_one_kernel<false><<<(unsigned)ctas, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc
