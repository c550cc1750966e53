// CUDA-core fp32 convolution kernels.
//
// (1) conv_simt_kernel: shape-generic implicit GEMM (64 pixels x 64 channels per CTA,
//     4x4 register tile, 16-deep k slab = one filter tap x 16 channels).  It is the
//     exact-fp32 math mode (ADVOC_MATH_FP32) used to cross-check the tensor-core path on
//     the GPU and the kernel for the degenerate layers a 128xN MMA tile cannot fill:
//     Cin in {1,2} (encoder_1, discriminator layer_1) and Cout = 1 (decoder_1, layer_5).
//     kTransposed = false : y[b,oh,ow,co]  = sum x[b, oh*sh-pt+kh, ow*sw-pl+kw, ci] w
//     kTransposed = true  : y[b,h,w,cb]    = sum x[b,(h+pt-kh)/sh,(w+pl-kw)/sw, cs] w
//     (the transposed form is tf conv2d_transpose and also the input gradient of a conv).
// Replaces tf.layers.conv2d / conv2d_transpose, models/advoc/advoc_model.py:27-32,46-51,65-69.
#include "epilogue.cuh"

namespace advoc {

int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* e) {
  ADVOC_REQUIRE(ep != nullptr, ADVOC_BAD_ARG, "epilogue is NULL");
  ADVOC_REQUIRE(ep->d_out0 != nullptr, ADVOC_BAD_ARG, "epilogue.d_out0 is NULL");
  ADVOC_REQUIRE(ep->store_w >= 0 && ep->store_w <= Wfull, ADVOC_BAD_SHAPE,
                "store_w %d outside [0,%d]", ep->store_w, Wfull);
  ADVOC_REQUIRE(ep->ld0 >= ep->c_off0 + Cout && ep->c_off0 >= 0, ADVOC_BAD_SHAPE,
                "out0 channel window [%d,%d) does not fit ld %d", ep->c_off0, ep->c_off0 + Cout,
                ep->ld0);
  if (ep->d_out1)
    ADVOC_REQUIRE(ep->ld1 >= ep->c_off1 + Cout && ep->c_off1 >= 0, ADVOC_BAD_SHAPE,
                  "out1 channel window [%d,%d) does not fit ld %d", ep->c_off1, ep->c_off1 + Cout,
                  ep->ld1);
  ADVOC_REQUIRE(ep->keep_prob > 0.f && ep->keep_prob <= 1.f, ADVOC_BAD_ARG, "keep_prob %f not in (0,1]",
                (double)ep->keep_prob);
  ADVOC_REQUIRE(ep->act0 >= 0 && ep->act0 <= ADVOC_ACT_TANH && ep->act1 >= 0 &&
                    ep->act1 <= ADVOC_ACT_TANH,
                ADVOC_BAD_ARG, "unknown activation");
  e->bias = ep->d_bias;
  e->out0 = ep->d_out0;
  e->out1 = ep->d_out1;
  e->mask = ep->d_dropout_mask;
  e->seed = ep->seed;
  e->act0 = ep->act0;
  e->act1 = ep->act1;
  e->ld0 = ep->ld0; e->coff0 = ep->c_off0;
  e->ld1 = ep->ld1; e->coff1 = ep->c_off1;
  e->Hs = Hs;
  e->Ws = ep->store_w ? ep->store_w : Wfull;
  e->Cout = Cout;
  e->alpha = ep->alpha;
  e->keep_prob = ep->keep_prob;
  e->round = ep->round_tf32;
  e->gate = ep->d_gate;
  e->ldg = ep->ld_gate; e->coffg = ep->c_off_gate;
  e->gate_act = ep->gate_act; e->gate_split = ep->gate_split;
  e->gscale0 = ep->gate_scale0; e->gscale1 = ep->gate_scale1;
  e->accumulate = ep->accumulate;
  e->seed_ptr = reinterpret_cast<const unsigned long long*>(ep->d_seed);
  ADVOC_REQUIRE((ep->out0_dtype == ADVOC_DT_F32 || ep->out0_dtype == ADVOC_DT_F16) &&
                    (ep->out1_dtype == ADVOC_DT_F32 || ep->out1_dtype == ADVOC_DT_F16),
                ADVOC_BAD_ARG, "unknown output dtype");
  ADVOC_REQUIRE(ep->out0_row_pad >= 0, ADVOC_BAD_ARG, "out0_row_pad must be >= 0");
  e->row_pad0 = ep->out0_row_pad;
  e->h0 = ep->out0_dtype == ADVOC_DT_F16;
  e->h1 = ep->d_out1 != nullptr && ep->out1_dtype == ADVOC_DT_F16;
  if (e->h0 || e->h1)
    ADVOC_REQUIRE(!ep->accumulate && !ep->d_gate, ADVOC_UNSUPPORTED,
                  "fp16 destinations are forward-pass only (no gate / accumulate)");
  if (ep->d_gate) {
    ADVOC_REQUIRE(ep->ld_gate >= ep->c_off_gate + Cout && ep->c_off_gate >= 0, ADVOC_BAD_SHAPE,
                  "gate channel window does not fit its ld");
    ADVOC_REQUIRE(ep->gate_act == ADVOC_ACT_LRELU || ep->gate_act == ADVOC_ACT_RELU, ADVOC_BAD_ARG,
                  "gate_act must be lrelu or relu");
    ADVOC_REQUIRE(ep->d_out1 == nullptr, ADVOC_BAD_ARG, "gate is not supported with a second output");
  }
  return ADVOC_OK;
}

namespace {

struct SimtArgs {
  const float* x;  // contraction-side activations [N, Hin, Win, ldx]
  const float* w;
  int N;
  int Hin, Win, ldx, Ck;       // input spatial, pixel stride, contraction channels
  int Hout, Wout, Cn;          // produced spatial extent and channels
  int kh, kw, sh, sw, pt, pl;
  long wt, wk, wn;             // filter strides: tap, contraction channel, produced channel
  EpiDev epi;
};

constexpr int TM = 64, TN = 64, TK = 16;

template <bool kTransposed>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtArgs a) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int pb[TM], ph[TM], pw_[TM];

  const long m0 = (long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const long M = (long)a.N * a.Hout * a.epi.Ws;  // only stored columns are produced
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  if (threadIdx.x < TM) {
    long m = m0 + threadIdx.x;
    if (m < M) {
      const int w = (int)(m % a.epi.Ws);
      const long r = m / a.epi.Ws;
      pw_[threadIdx.x] = w;
      ph[threadIdx.x] = (int)(r % a.Hout);
      pb[threadIdx.x] = (int)(r / a.Hout);
    } else {
      pb[threadIdx.x] = -1;
      ph[threadIdx.x] = 0;
      pw_[threadIdx.x] = 0;
    }
  }
  __syncthreads();

  float acc[4][4] = {};
  const int ntaps = a.kh * a.kw;
  for (int tap = 0; tap < ntaps; ++tap) {
    const int kh = tap / a.kw, kw = tap - kh * a.kw;
    for (int c0 = 0; c0 < a.Ck; c0 += TK) {
      // A slab: 64 pixels x 16 channels
      for (int i = threadIdx.x; i < TM * TK; i += 256) {
        const int mm = i / TK, kk = i - mm * TK;
        float v = 0.f;
        const int b = pb[mm];
        if (b >= 0 && c0 + kk < a.Ck) {
          int ih, iw;
          bool ok;
          if (!kTransposed) {
            ih = ph[mm] * a.sh - a.pt + kh;
            iw = pw_[mm] * a.sw - a.pl + kw;
            ok = ih >= 0 && ih < a.Hin && iw >= 0 && iw < a.Win;
          } else {
            const int th = ph[mm] + a.pt - kh, tw = pw_[mm] + a.pl - kw;
            ok = th >= 0 && tw >= 0 && (th % a.sh) == 0 && (tw % a.sw) == 0;
            ih = th / a.sh;
            iw = tw / a.sw;
            ok = ok && ih < a.Hin && iw < a.Win;
          }
          if (ok) v = __ldg(a.x + (((size_t)b * a.Hin + ih) * a.Win + iw) * a.ldx + c0 + kk);
        }
        As[kk][mm] = v;
      }
      // B slab: 16 channels x 64 produced channels
      for (int i = threadIdx.x; i < TN * TK; i += 256) {
        int nn, kk;
        if (a.wn == 1) { nn = i & (TN - 1); kk = i >> 6; } else { kk = i & (TK - 1); nn = i >> 4; }
        float v = 0.f;
        if (n0 + nn < a.Cn && c0 + kk < a.Ck)
          v = __ldg(a.w + tap * a.wt + (long)(c0 + kk) * a.wk + (long)(n0 + nn) * a.wn);
        Bs[kk][nn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < a.Cn) epi_store(a.epi, (size_t)m, n, acc[i][j]);
    }
  }
}

}  // namespace

int check_conv_desc(const advoc_conv_desc* d) {
  ADVOC_REQUIRE(d != nullptr, ADVOC_BAD_ARG, "conv desc is NULL");
  ADVOC_REQUIRE(d->N >= 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, ADVOC_BAD_SHAPE,
                "bad conv shape N%d H%d W%d Cin%d Cout%d", d->N, d->H, d->W, d->Cin, d->Cout);
  ADVOC_REQUIRE(d->kh > 0 && d->kw > 0 && d->kh <= 8 && d->kw <= 8 && d->sh > 0 && d->sw > 0,
                ADVOC_BAD_SHAPE, "bad kernel/stride");
  ADVOC_REQUIRE(d->pad_t >= 0 && d->pad_l >= 0 && d->pad_t < d->kh && d->pad_l < d->kw, ADVOC_BAD_SHAPE,
                "bad padding");
  ADVOC_REQUIRE(d->Ho > 0 && d->Wo > 0, ADVOC_BAD_SHAPE, "bad output size");
  // every tap of every output pixel must start inside the leading pad
  ADVOC_REQUIRE((d->Ho - 1) * d->sh - d->pad_t < d->H && (d->Wo - 1) * d->sw - d->pad_l < d->W,
                ADVOC_BAD_SHAPE, "output %dx%d reaches past the input %dx%d", d->Ho, d->Wo, d->H, d->W);
  return ADVOC_OK;
}

// conv_direct.cu
bool conv_thin_eligible(const advoc_conv_desc* d, const advoc_epilogue* ep);
int conv_thin(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
              void* stream);
bool deconv_to_one_eligible(const advoc_conv_desc* d, const float* x, int ldx, const advoc_epilogue* ep);
int deconv_to_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                  void* stream);

bool deconv_from_one_eligible(const advoc_conv_desc* d, const advoc_epilogue* ep);
int deconv_from_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                    void* stream);
// train.cu
bool conv_to_one_eligible(const advoc_conv_desc* d, const float* x, int ldx, const advoc_epilogue* ep);
int conv_to_one(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                void* stream);

int conv_fwd_simt(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                  const advoc_epilogue* ep, void* stream) {
  if (ep && ep->d_out0 && conv_thin_eligible(d, ep)) return conv_thin(d, x, ldx, w, ep, stream);
  ADVOC_REQUIRE(!ep || ep->out0_row_pad == 0, ADVOC_UNSUPPORTED,
                "out0_row_pad is only supported by the thin-input convolution");
  if (ep && ep->d_out0 && conv_to_one_eligible(d, x, ldx, ep)) return conv_to_one(d, x, ldx, w, ep, stream);
  SimtArgs a = {};
  int st = lower_epilogue(ep, d->Ho, d->Wo, d->Cout, &a.epi);
  if (st) return st;
  a.x = x; a.w = w; a.N = d->N;
  a.Hin = d->H; a.Win = d->W; a.ldx = ldx; a.Ck = d->Cin;
  a.Hout = d->Ho; a.Wout = d->Wo; a.Cn = d->Cout;
  a.kh = d->kh; a.kw = d->kw; a.sh = d->sh; a.sw = d->sw; a.pt = d->pad_t; a.pl = d->pad_l;
  a.wt = (long)d->Cin * d->Cout; a.wk = d->Cout; a.wn = 1;  // HWIO
  const long M = (long)a.N * a.Hout * a.epi.Ws;
  if (M == 0) return ADVOC_OK;
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((a.Cn + TN - 1) / TN));
  conv_simt_kernel<false><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

// w_is_hwoi: filter is [kh,kw,Cbig,Csmall] (tf conv2d_transpose kernel), else HWIO
// [kh,kw,Cbig,Csmall]... both are indexed [tap][big][small]; what differs between the two
// callers (deconv forward / conv dgrad) is only the naming, so one stride set serves both.
int conv_transposed_simt(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                         const advoc_epilogue* ep, void* stream) {
  if (ep && ep->d_out0 && deconv_to_one_eligible(d, x, ldx, ep)) return deconv_to_one(d, x, ldx, w, ep, stream);
  if (ep && ep->d_out0 && deconv_from_one_eligible(d, ep)) return deconv_from_one(d, x, ldx, w, ep, stream);
  SimtArgs a = {};
  int st = lower_epilogue(ep, d->H, d->W, d->Cin, &a.epi);
  if (st) return st;
  a.x = x; a.w = w; a.N = d->N;
  a.Hin = d->Ho; a.Win = d->Wo; a.ldx = ldx; a.Ck = d->Cout;
  a.Hout = d->H; a.Wout = d->W; a.Cn = d->Cin;
  a.kh = d->kh; a.kw = d->kw; a.sh = d->sh; a.sw = d->sw; a.pt = d->pad_t; a.pl = d->pad_l;
  a.wt = (long)d->Cin * d->Cout; a.wk = 1; a.wn = d->Cout;  // [tap][big][small]
  const long M = (long)a.N * a.Hout * a.epi.Ws;
  if (M == 0) return ADVOC_OK;
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((a.Cn + TN - 1) / TN));
  conv_simt_kernel<true><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc
