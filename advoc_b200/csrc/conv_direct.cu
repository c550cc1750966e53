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
// thin-input conv: one thread = one output pixel x 4 consecutive output channels
// ---------------------------------------------------------------------------------------------
template <int CIN>
struct ThinArgs {
  const float* x;
  const float* w;  // HWIO [kh*kw][CIN][Cout]
  int N, H, W, ldx, Ho, Wo, Cout;
  int kh, kw, sh, sw, pt, pl;
  EpiDev epi;
};

template <int CIN, int TAPS>
__global__ void __launch_bounds__(256) conv_thin_in_kernel(const ThinArgs<CIN> a) {
  extern __shared__ float ws[];  // [TAPS*CIN][Cout]
  for (int i = threadIdx.x; i < TAPS * CIN * a.Cout; i += blockDim.x) ws[i] = __ldg(a.w + i);
  __syncthreads();
  const int groups = a.Cout >> 2;
  const long total = (long)a.N * a.Ho * a.Wo * groups;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % groups);
    const long pix = t / groups;
    const int ow = (int)(pix % a.Wo);
    const long r = pix / a.Wo;
    const int oh = (int)(r % a.Ho);
    const long img = r / a.Ho;
    const float* xb = a.x + (size_t)img * a.H * a.W * a.ldx;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < TAPS; ++tap) {
      const int ih = oh * a.sh - a.pt + tap / a.kw;
      const int iw = ow * a.sw - a.pl + tap % a.kw;
      if (ih < 0 || ih >= a.H || iw < 0 || iw >= a.W) continue;
      const float* xp = xb + ((size_t)ih * a.W + iw) * a.ldx;
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float xv = __ldg(xp + c);
        const float4 wv = *reinterpret_cast<const float4*>(ws + (tap * CIN + c) * a.Cout + cg * 4);
        acc[0] = fmaf(xv, wv.x, acc[0]);
        acc[1] = fmaf(xv, wv.y, acc[1]);
        acc[2] = fmaf(xv, wv.z, acc[2]);
        acc[3] = fmaf(xv, wv.w, acc[3]);
      }
    }
    const EpiDev& e = a.epi;
    const int n = cg * 4;
    float v[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[j] + (e.bias ? __ldg(e.bias + n + j) : 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      y[j] = apply_act(v[j], e.act0, e.alpha);
      if (e.round) y[j] = round_tf32(y[j]);
    }
    *reinterpret_cast<float4*>(e.out0 + (size_t)pix * e.ld0 + e.coff0 + n) = make_float4(y[0], y[1], y[2], y[3]);
    if (e.out1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[j] = apply_act(v[j], e.act1, e.alpha);
        if (e.round) y[j] = round_tf32(y[j]);
      }
      *reinterpret_cast<float4*>(e.out1 + (size_t)pix * e.ld1 + e.coff1 + n) =
          make_float4(y[0], y[1], y[2], y[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k4 s2 transposed conv to ONE channel: one thread = one 2x2 output block (the four parity
// classes of input position (a,b)); it reads the 3x3 input neighbourhood, 4 channels at a time.
//   out[2a+ph, 2b+pw] = sum_{dr,dc} dot(x[a+dr, b+dc, :], w[kh = ph+1-2dr, kw = pw+1-2dc, 0, :])
// ---------------------------------------------------------------------------------------------
struct ToOneArgs {
  const float* x;  // [N, Hs, Ws, ldx], Cs channels used
  const float* w;  // HWOI [16][1][Cs]
  int N, Hs, Ws, ldx, Cs;
  EpiDev epi;      // stored extent Hs*2 x epi.Ws
};

__global__ void __launch_bounds__(128) deconv_to_one_kernel(const ToOneArgs a) {
  extern __shared__ float ws[];  // [16][Cs]
  for (int i = threadIdx.x; i < 16 * a.Cs; i += blockDim.x) ws[i] = __ldg(a.w + i);
  __syncthreads();
  const long total = (long)a.N * a.Hs * a.Ws;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int b = (int)(t % a.Ws);
  const long r = t / a.Ws;
  const int ar = (int)(r % a.Hs);
  const long img = r / a.Hs;
  const float* xb = a.x + (size_t)img * a.Hs * a.Ws * a.ldx;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int c = 0; c < a.Cs; c += 4) {
    float4 xv[3][3];
#pragma unroll
    for (int dr = 0; dr < 3; ++dr)
#pragma unroll
      for (int dc = 0; dc < 3; ++dc) {
        const int ih = ar + dr - 1, iw = b + dc - 1;
        xv[dr][dc] = (ih >= 0 && ih < a.Hs && iw >= 0 && iw < a.Ws)
                         ? __ldg(reinterpret_cast<const float4*>(xb + ((size_t)ih * a.Ws + iw) * a.ldx + c))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw)
#pragma unroll
        for (int jr = 0; jr < 2; ++jr)
#pragma unroll
          for (int jc = 0; jc < 2; ++jc) {
            const int dr = ph - 1 + jr + 1, dc = pw - 1 + jc + 1;  // index into xv (offset by +1)
            const int kh = ph + 1 - 2 * (dr - 1), kw = pw + 1 - 2 * (dc - 1);
            const float4 wv = *reinterpret_cast<const float4*>(ws + (kh * 4 + kw) * a.Cs + c);
            const float4 x4 = xv[dr][dc];
            acc[ph][pw] = fmaf(x4.x, wv.x, fmaf(x4.y, wv.y, fmaf(x4.z, wv.z, fmaf(x4.w, wv.w, acc[ph][pw]))));
          }
  }
#pragma unroll
  for (int ph = 0; ph < 2; ++ph)
#pragma unroll
    for (int pw = 0; pw < 2; ++pw) {
      const int ow = 2 * b + pw;
      if (ow >= a.epi.Ws) continue;
      const size_t pix = ((size_t)img * a.epi.Hs + (2 * ar + ph)) * a.epi.Ws + ow;
      epi_store(a.epi, pix, 0, acc[ph][pw]);
    }
}

}  // namespace

bool conv_thin_eligible(const advoc_conv_desc* d, const advoc_epilogue* ep) {
  auto ok = [](const float* p, int ld, int co) { return aligned16(p) && ld % 4 == 0 && co % 4 == 0; };
  return (d->Cin == 1 || d->Cin == 2) && d->kh == 4 && d->kw == 4 && d->Cout % 4 == 0 && d->Cout <= 512 &&
         ep->keep_prob >= 1.f && ep->store_w == 0 && ok(ep->d_out0, ep->ld0, ep->c_off0) &&
         (!ep->d_out1 || ok(ep->d_out1, ep->ld1, ep->c_off1));
}

int conv_thin(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
              void* stream) {
  EpiDev e;
  int st = lower_epilogue(ep, d->Ho, d->Wo, d->Cout, &e);
  if (st) return st;
  const long total = (long)d->N * d->Ho * d->Wo * (d->Cout / 4);
  if (total == 0) return ADVOC_OK;
  const int blocks = (int)((total + 255) / 256 < (long)sm_count() * 16 ? (total + 255) / 256 : sm_count() * 16);
  const size_t smem = (size_t)16 * d->Cin * d->Cout * sizeof(float);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (d->Cin == 1) {
    ThinArgs<1> a = {x, w, d->N, d->H, d->W, ldx, d->Ho, d->Wo, d->Cout, d->kh, d->kw, d->sh, d->sw,
                     d->pad_t, d->pad_l, e};
    conv_thin_in_kernel<1, 16><<<blocks, 256, smem, s>>>(a);
  } else {
    ThinArgs<2> a = {x, w, d->N, d->H, d->W, ldx, d->Ho, d->Wo, d->Cout, d->kh, d->kw, d->sh, d->sw,
                     d->pad_t, d->pad_l, e};
    static bool cfg = false;
    if (!cfg && smem > 48 * 1024) {
      ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_thin_in_kernel<2, 16>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      cfg = true;
    }
    conv_thin_in_kernel<2, 16><<<blocks, 256, smem, s>>>(a);
  }
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

bool deconv_to_one_eligible(const advoc_conv_desc* d, const float* x, int ldx, const advoc_epilogue* ep) {
  return d->Cin == 1 && d->kh == 4 && d->kw == 4 && d->sh == 2 && d->sw == 2 && d->pad_t == 1 && d->pad_l == 1 &&
         d->H == 2 * d->Ho && d->W == 2 * d->Wo && d->Cout % 4 == 0 && d->Cout <= 2048 && ldx % 4 == 0 &&
         aligned16(x) && ep->keep_prob >= 1.f && ep->d_out1 == nullptr;
}

int deconv_to_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                  void* stream) {
  ToOneArgs a = {};
  int st = lower_epilogue(ep, d->H, d->W, 1, &a.epi);
  if (st) return st;
  a.x = x; a.w = w; a.N = d->N; a.Hs = d->Ho; a.Ws = d->Wo; a.ldx = ldx; a.Cs = d->Cout;
  const long total = (long)a.N * a.Hs * a.Ws;
  if (total == 0) return ADVOC_OK;
  const size_t smem = (size_t)16 * a.Cs * sizeof(float);
  static bool cfg = false;
  if (!cfg && smem > 48 * 1024) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(deconv_to_one_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          128 * 1024));
    cfg = true;
  }
  deconv_to_one_kernel<<<(unsigned)((total + 127) / 128), 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc
