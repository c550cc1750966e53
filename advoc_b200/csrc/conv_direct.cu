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
        if (a.sh == 1) ih = q;                                   // stride 1 (PatchGAN head): no division
        else ih = (q >= 0 && q % a.sh == 0) ? q / a.sh : -1;
      }
      if (ih < 0 || ih >= a.H) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        int iw;
        if (!kT) {
          iw = iw0 + kw;
        } else {
          const int q = ow + a.pl - kw;
          if (a.sw == 1) iw = q;
          else iw = (q >= 0 && q % a.sw == 0) ? q / a.sw : -1;
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
// thin-input conv, forward fast path: one thread = one output pixel x ALL COUT channels (register
// accumulators), filters as float4 broadcasts from smem.  ~4x fewer instructions per pixel than
// the channel-split kernel above; used when the epilogue is the plain forward one
// (bias + slope activation on one or two outputs).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_slope(int act, float alpha) {
  return act == ADVOC_ACT_LRELU ? alpha : (act == ADVOC_ACT_RELU ? 0.f : 1.f);
}

// PPT = output pixels per thread (two pixels share each filter word).  Measured r02: PPT = 2 changes nothing
// with fp16 outputs (79 us) and costs 40 % with fp32 outputs (116 against 82 us: 140 registers, twice the
// store instructions per thread) -- the kernel is bound by instruction issue (address / select / convert),
// not by the filter reads -- so every launch uses PPT = 1; the parameter stays for the next attempt.
template <int CIN, int COUT, int KS = 4, int PPT = 1>
__global__ void __launch_bounds__(128) conv_thin_px_kernel(const ThinArgs<CIN> a) {
  __shared__ float4 ws4[KS * KS * CIN * COUT / 4];
  __shared__ float4 bs4[COUT / 4];
  for (int i = threadIdx.x; i < KS * KS * CIN * COUT / 4; i += blockDim.x)
    ws4[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);
  const EpiDev& e = a.epi;
  if (threadIdx.x < COUT / 4)
    bs4[threadIdx.x] = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias) + threadIdx.x)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  constexpr int PXB = 128 * PPT;     // pixels per block
  const unsigned pix_in_img = a.Ho * a.Wo;
  const long npix = (long)a.N * pix_in_img;
  const long pix0 = (long)blockIdx.x * PXB;
  float4 acc[PPT][COUT / 4];
  const float* row0[PPT];            // image base of each pixel
  int ih0[PPT], iw0[PPT];
  bool valid[PPT];
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    const long pix = pix0 + threadIdx.x + 128 * u;     // pixel u of this thread: 128 apart, so that the
    valid[u] = pix < npix;                              // threads of a warp read neighbouring input columns
    const long pp = valid[u] ? pix : 0;
    const unsigned img = (unsigned)(pp / pix_in_img);
    const unsigned rem = (unsigned)(pp - (long)img * pix_in_img);
    const int oh = rem / a.Wo, ow = rem - oh * a.Wo;
    row0[u] = a.x + (size_t)img * a.H * a.W * a.ldx;
    ih0[u] = oh * a.sh - a.pt;
    iw0[u] = ow * a.sw - a.pl;
#pragma unroll
    for (int j = 0; j < COUT / 4; ++j) acc[u][j] = bs4[j];
  }
#pragma unroll
  for (int kh = 0; kh < KS; ++kh) {
#pragma unroll
    for (int kw = 0; kw < KS; ++kw) {
      float xv[PPT][CIN];
#pragma unroll
      for (int u = 0; u < PPT; ++u) {
        const int ih = ih0[u] + kh, iw = iw0[u] + kw;
        const bool in = valid[u] && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
#pragma unroll
        for (int c = 0; c < CIN; ++c)
          xv[u][c] = in ? __ldg(row0[u] + ((size_t)ih * a.W + iw) * a.ldx + c) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float4* wp = ws4 + ((kh * KS + kw) * CIN + c) * (COUT / 4);
#pragma unroll
        for (int j = 0; j < COUT / 4; ++j) {
          const float4 w = wp[j];
#pragma unroll
          for (int u = 0; u < PPT; ++u) {
            acc[u][j].x = fmaf(xv[u][c], w.x, acc[u][j].x);
            acc[u][j].y = fmaf(xv[u][c], w.y, acc[u][j].y);
            acc[u][j].z = fmaf(xv[u][c], w.z, acc[u][j].z);
            acc[u][j].w = fmaf(xv[u][c], w.w, acc[u][j].w);
          }
        }
      }
    }
  }
  // Stores: a thread-per-pixel store touches 32 different 128-byte lines per instruction.  Stage the
  // raw accumulators in shared memory (16-byte chunk j of pixel p at chunk j ^ (p & 7): conflict
  // free both ways) and write them out cooperatively, consecutive lanes along the channels of a
  // pixel, so every store instruction covers whole lines; both activations are applied on the way out.
  {
    __shared__ float4 stage[PXB * (COUT / 4)];
    constexpr int Q = COUT / 4;
    static_assert(Q == 8 || Q == 16, "staging layout assumes 32 or 64 output channels");
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      const int p = threadIdx.x + 128 * u;
      if (valid[u]) {
#pragma unroll
        for (int j = 0; j < Q; ++j) stage[p * Q + (j ^ (p & 7))] = acc[u][j];
      }
    }
    __syncthreads();
    const float s0 = act_slope(e.act0, e.alpha), s1 = act_slope(e.act1, e.alpha);
    auto act4 = [](const float4& v, float s, float (&y)[4]) {
      y[0] = v.x > 0.f ? v.x : s * v.x; y[1] = v.y > 0.f ? v.y : s * v.y;
      y[2] = v.z > 0.f ? v.z : s * v.z; y[3] = v.w > 0.f ? v.w : s * v.w;
    };
    if (e.h0 && (!e.out1 || e.h1)) {
      // fp16 destinations: eight channels (two staged chunks) per lane = one 16-byte store
      constexpr int Q2 = Q / 2;
      __half* o0 = reinterpret_cast<__half*>(e.out0);
      __half* o1 = reinterpret_cast<__half*>(e.out1);
#pragma unroll
      for (int k = 0; k < Q2 * PPT; ++k) {
        const int i = threadIdx.x + 128 * k;
        const int p = i / Q2, j = (i % Q2) * 2;
        const long gp = pix0 + p;
        if (gp >= npix) continue;
        const float4 va = stage[p * Q + (j ^ (p & 7))], vb = stage[p * Q + ((j + 1) ^ (p & 7))];
        float ya[4], yb[4];
        act4(va, s0, ya); act4(vb, s0, yb);
        const size_t gp0 = e.row_pad0 ? (size_t)gp + (size_t)(gp / a.Wo) * e.row_pad0 : (size_t)gp;
        *reinterpret_cast<uint4*>(o0 + gp0 * e.ld0 + e.coff0 + 4 * j) =
            make_uint4(pack_half2(ya[0], ya[1]), pack_half2(ya[2], ya[3]), pack_half2(yb[0], yb[1]), pack_half2(yb[2], yb[3]));
        if (e.out1) {
          act4(va, s1, ya); act4(vb, s1, yb);
          *reinterpret_cast<uint4*>(o1 + (size_t)gp * e.ld1 + e.coff1 + 4 * j) =
              make_uint4(pack_half2(ya[0], ya[1]), pack_half2(ya[2], ya[3]), pack_half2(yb[0], yb[1]), pack_half2(yb[2], yb[3]));
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < Q * PPT; ++k) {
        const int i = threadIdx.x + 128 * k;
        const int p = i / Q, j = i % Q;
        const long gp = pix0 + p;
        if (gp >= npix) continue;
        const float4 v = stage[p * Q + (j ^ (p & 7))];
        float y[4];
        act4(v, s0, y);
        if (e.round && !e.h0) {
#pragma unroll
          for (int u = 0; u < 4; ++u) y[u] = round_tf32(y[u]);
        }
        // fp32 or fp16 destination; rows of out0 may carry e.row_pad0 extra pixels
        const size_t gp0 = e.row_pad0 ? (size_t)gp + (size_t)(gp / a.Wo) * e.row_pad0 : (size_t)gp;
        store4(e.out0, e.h0, gp0 * e.ld0 + e.coff0 + 4 * j, y);
        if (e.out1) {
          float z[4];
          act4(v, s1, z);
          if (e.round && !e.h1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) z[u] = round_tf32(z[u]);
          }
          store4(e.out1, e.h1, (size_t)gp * e.ld1 + e.coff1 + 4 * j, z);
        }
      }
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

__global__ void __launch_bounds__(256) deconv_to_one_kernel(const ToOneArgs a) {
  __shared__ float tsm[T1_TH * T1_TW][17];  // +1 pad: conflict-free column access
  extern __shared__ float4 wsm[];           // [16][Cs/4] filter
  const int tw_i = blockIdx.x % a.tiles_w;
  const int th_i = (blockIdx.x / a.tiles_w) % a.tiles_h;
  const int img = blockIdx.x / (a.tiles_w * a.tiles_h);
  const int a0 = th_i * T1_IH - 1, b0 = tw_i * T1_IW - 1;  // tile origin (halo included)
  const float* xb = a.x + (size_t)img * a.Hs * a.Ws * a.ldx;
  const int c4 = a.Cs >> 2;
  for (int i = threadIdx.x; i < 16 * c4; i += blockDim.x) wsm[i] = __ldg(reinterpret_cast<const float4*>(a.w) + i);
  __syncthreads();

  // phase 1: thread t owns positions t and t + 256 (tile rows r and r + 8, same column)
  {

    const int col = threadIdx.x & (T1_TW - 1), row = threadIdx.x >> 5;
    const int bc = b0 + col, ar0 = a0 + row, ar1 = ar0 + T1_TH / 2;
    const bool okc = bc >= 0 && bc < a.Ws;
    const bool ok0 = okc && ar0 >= 0 && ar0 < a.Hs, ok1 = okc && ar1 >= 0 && ar1 < a.Hs;
    const float4* x0 = reinterpret_cast<const float4*>(xb + ((size_t)(ok0 ? ar0 : 0) * a.Ws + (okc ? bc : 0)) * a.ldx);
    const float4* x1 = reinterpret_cast<const float4*>(xb + ((size_t)(ok1 ? ar1 : 0) * a.Ws + (okc ? bc : 0)) * a.ldx);
    float t0[16], t1[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { t0[k] = 0.f; t1[k] = 0.f; }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < c4; ++c) {
      const float4 u = ok0 ? __ldg(x0 + c) : zero;
      const float4 v = ok1 ? __ldg(x1 + c) : zero;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 w = wsm[k * c4 + c];
        t0[k] = fmaf(u.x, w.x, fmaf(u.y, w.y, fmaf(u.z, w.z, fmaf(u.w, w.w, t0[k]))));
        t1[k] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, t1[k]))));
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      tsm[threadIdx.x][k] = t0[k];
      tsm[threadIdx.x + 256][k] = t1[k];
    }
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

static bool thin_plain(const advoc_epilogue* ep) {
  auto slope_act = [](int a) { return a == ADVOC_ACT_NONE || a == ADVOC_ACT_LRELU || a == ADVOC_ACT_RELU; };
  return !ep->d_gate && !ep->accumulate && slope_act(ep->act0) && slope_act(ep->act1) &&
         (!ep->d_bias || aligned16(ep->d_bias));
}

bool conv_thin_eligible(const advoc_conv_desc* d, const advoc_epilogue* ep) {
  auto ok = [](const float* p, int ld, int co) { return aligned16(p) && ld % 4 == 0 && co % 4 == 0; };
  // 5x5 (MelspecGAN conv_0, models/melspecgan/conv2d.py:182-184): the per-pixel fast path only
  const bool k5 = d->kh == 5 && d->kw == 5 && d->Cin == 1 && d->Cout == 64 && thin_plain(ep);
  return (d->Cin == 1 || d->Cin == 2) && ((d->kh == 4 && d->kw == 4) || k5) && d->Cout % 4 == 0 && d->Cout <= 256 &&
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
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto slope_act = [](int a) { return a == ADVOC_ACT_NONE || a == ADVOC_ACT_LRELU || a == ADVOC_ACT_RELU; };
  const bool plain = !ep->d_gate && !ep->accumulate && slope_act(ep->act0) && slope_act(ep->act1) &&
                     (!ep->d_bias || aligned16(ep->d_bias)) && aligned16(w) && (npix + 127) / 128 < 2147483647L;
  if (plain && (d->Cout == 32 || d->Cout == 64)) {
    const unsigned blocks = (unsigned)((npix + 127) / 128);
    if (d->Cin == 1) {
      ThinArgs<1> a = {x, w, d->N, d->H, d->W, ldx, d->Ho, d->Wo, d->Cout, d->sh, d->sw, d->pad_t, d->pad_l, e};
      if (d->kh == 5) conv_thin_px_kernel<1, 64, 5><<<blocks, 128, 0, s>>>(a);
      else if (d->Cout == 32) conv_thin_px_kernel<1, 32><<<blocks, 128, 0, s>>>(a);
      else conv_thin_px_kernel<1, 64><<<blocks, 128, 0, s>>>(a);
    } else {
      ThinArgs<2> a = {x, w, d->N, d->H, d->W, ldx, d->Ho, d->Wo, d->Cout, d->sh, d->sw, d->pad_t, d->pad_l, e};
      if (d->Cout == 32) conv_thin_px_kernel<2, 32><<<blocks, 128, 0, s>>>(a);
      else conv_thin_px_kernel<2, 64><<<blocks, 128, 0, s>>>(a);
    }
    count_launch();
    ADVOC_CHECK_CUDA(cudaGetLastError());
    return ADVOC_OK;
  }
  ADVOC_REQUIRE(!e.h0 && !e.h1 && e.row_pad0 == 0, ADVOC_UNSUPPORTED,
                "fp16 destinations / padded rows need the plain 32/64-channel thin conv");
  const int ppb = 256 / (d->Cout / 4);
  const long want = (npix + ppb - 1) / ppb;
  const int blocks = (int)(want < (long)sm_count() * 32 ? want : (long)sm_count() * 32);
  const size_t smem = (size_t)16 * d->Cin * d->Cout * sizeof(float);
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
  return d->Cout == 1 && d->kh == 4 && d->kw == 4 && d->Cin % 4 == 0 && d->Cin <= 512 &&
         256 % (d->Cin / 4) == 0 && ep->keep_prob >= 1.f && ep->store_w == 0 && ep->d_out1 == nullptr &&
         ok(ep->d_out0, ep->ld0, ep->c_off0);
}

int deconv_from_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                    void* stream) {
  EpiDev e;
  int st = lower_epilogue(ep, d->H, d->W, d->Cin, &e);
  if (st) return st;
  ADVOC_REQUIRE(!e.h0, ADVOC_UNSUPPORTED, "fp16 destination not supported by the from-one-channel kernel");
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
         d->H == 2 * d->Ho && (d->W == 2 * d->Wo || (ep->accumulate && d->W > 2 * d->Wo)) && d->Cout % 4 == 0 && d->Cout <= 192 && ldx % 4 == 0 && aligned16(x) && ep->keep_prob >= 1.f && ep->d_out1 == nullptr;
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
  const size_t smem = (size_t)16 * a.Cs * sizeof(float);
  ADVOC_REQUIRE(smem <= 12 * 1024, ADVOC_UNSUPPORTED, "too many input channels for deconv_to_one");
  deconv_to_one_kernel<<<(unsigned)ctas, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc
