/*
 * advoc_b200 -- C-ABI of the B200-native adversarial-vocoder hot path.
 *
 * The reference (paarthneekhara/advoc) has no FFI boundary: its hot path is a
 * TensorFlow-1 graph built from Python.  This header is the boundary a
 * maintainer binds instead (ctypes stub in INTEGRATION.md); every entry point
 * names the reference call site it replaces as `path:line` under the reference
 * repository root.
 *
 * Conventions
 *   - every function returns an advoc_status (0 == ADVOC_OK); the message of the
 *     last failure on the calling thread is read with advoc_last_error().
 *   - all pointers named d_* are DEVICE pointers owned by the caller, 16-byte
 *     aligned and densely packed in the stated layout; the library allocates
 *     nothing on the hot path and never synchronises the device.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - activations are NHWC float32 [batch, time(H), freq(W), channels] -- the
 *     reference's layout.  Convolution kernels are given in the TensorFlow
 *     layouts (conv HWIO [kh,kw,Cin,Cout]; conv_transpose HWOI [kh,kw,Cout,Cin])
 *     and are re-packed by advoc_pack_filter into the K-major TF32 operand
 *     layout the tensor-core kernels read.
 *   - nothing here takes or returns a torch type.
 */
#ifndef ADVOC_B200_H_
#define ADVOC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVOC_B200_VERSION 100

#if defined(__GNUC__)
#define ADVOC_API __attribute__((visibility("default")))
#else
#define ADVOC_API
#endif

typedef enum advoc_status {
  ADVOC_OK = 0,
  ADVOC_BAD_ARG = 1,          /* -> ValueError           */
  ADVOC_BAD_SHAPE = 2,        /* -> ValueError           */
  ADVOC_BAD_ALIGN = 3,        /* -> ValueError           */
  ADVOC_UNSUPPORTED = 4,      /* -> NotImplementedError  */
  ADVOC_WORKSPACE_TOO_SMALL = 5,
  ADVOC_CUDA_ERROR = 6        /* -> RuntimeError         */
} advoc_status;

typedef enum advoc_act {
  ADVOC_ACT_NONE = 0,
  ADVOC_ACT_LRELU = 1,   /* max(alpha*x, x)  models/advoc/advoc_model.py:86-87 */
  ADVOC_ACT_RELU = 2,    /* models/advoc/advoc_model.py:138                   */
  ADVOC_ACT_SIGMOID = 3, /* models/advoc/advoc_model.py:201                   */
  ADVOC_ACT_TANH = 4     /* models/melspecgan/conv2d.py:141                   */
} advoc_act;

typedef enum advoc_math {
  ADVOC_MATH_AUTO = 0,  /* tensor cores (TF32, fp32 accumulate) where the layer is eligible */
  ADVOC_MATH_FP32 = 1,  /* CUDA-core fp32 direct kernels (exact-fp32 cross-check path)      */
  ADVOC_MATH_TF32 = 2,  /* force the tcgen05 path; ADVOC_UNSUPPORTED if not eligible        */
  /* fp16 OPERANDS (d_x and d_w hold IEEE half: the same 10-bit mantissa as TF32, half the bytes;
   * tcgen05 kind::f16, K = 16 per instruction), fp32 accumulate.  Forward pass only.  Needs the
   * contraction channel count % 64 == 0, ld_x % 8 == 0 and the filter packed by advoc_pack_filter with
   * mode 2; ADVOC_UNSUPPORTED otherwise (ask advoc_conv2d_path first).  Values must stay inside
   * fp16's range (|v| < 65504): true for magnitude spectra and N(0, 0.02)-initialised stacks, and the
   * producing epilogues saturate instead of overflowing to inf. */
  ADVOC_MATH_F16 = 3
} advoc_math;

typedef enum advoc_dtype {
  ADVOC_DT_F32 = 0,
  ADVOC_DT_F16 = 1
} advoc_dtype;

ADVOC_API int advoc_version(void);
/* Copies the calling thread's last error text (NUL terminated) into buf. */
ADVOC_API int advoc_last_error(char* buf, size_t buf_len);
/* Number of kernels this library has launched in this process (bench.py gpu_launches). */
ADVOC_API unsigned long long advoc_launch_count(void);
/* Compute capability of the current device as major*10+minor (100 on B200). */
ADVOC_API int advoc_device_arch(int* arch);

/* ------------------------------------------------------------------------- *
 * Spectral features
 * ------------------------------------------------------------------------- */

/* Frame count of the reference framing rules.  pad_end = 1: ceil(n / hop) (advoc/spectral.py:32-39 and
 * tf.contrib.signal.stft(pad_end=True), :75-81); 0: lws' own rule, ceil((n - nfft) / hop) + 1, at least 1,
 * the last partial frame zero-padded (tests/test_spectral.py:35-36); 2: tf.contrib.signal.stft(pad_end=
 * False), floor((n - nfft) / hop) + 1 whole frames, possibly 0.  The same codes are accepted as the
 * `pad_end` argument of advoc_stft_f32. */
ADVOC_API int advoc_num_frames(int nsamps, int nfft, int nhop, int pad_end);

/* Framed, windowed real FFT.  d_wav [batch, nsamps, 1, nch] f32 ->
 * d_out_c64 [batch, frames, nfft/2+1, nch] interleaved (re,im) f32 and/or
 * d_out_mag [same] f32 magnitude (either may be NULL).  d_window [nfft] f32.
 * d_twiddle [nfft] (cos,-sin) pairs made by advoc_fill_twiddle_host / any caller.
 * replaces: advoc/spectral.py:11-41 (`stft`, lws C++), :60-83 (`stft_tf`). */
ADVOC_API int advoc_stft_f32(const float* d_wav, int batch, int nsamps, int nch, int nfft, int nhop,
                   int pad_end, const float* d_window, const float* d_twiddle,
                   float* d_out_c64, float* d_out_mag, void* stream);

/* Fused waveform -> dB-normalised mel: STFT -> |.| -> mel filterbank -> 20log10 -> clip.
 * d_mel_fb [nmels, nfft/2+1] f32 row-major.  d_out [batch, frames, nmels, nch] f32.
 * replaces: advoc/spectral.py:98-154 and :158-227 (`waveform_to_melspec[_tf]`). */
ADVOC_API int advoc_melspec_f32(const float* d_wav, int batch, int nsamps, int nch, int nfft, int nhop,
                      const float* d_window, const float* d_twiddle, const float* d_mel_fb,
                      const int* d_mel_ranges, int nmels, float min_level_db,
                      float ref_level_db, float* d_out, void* stream);

/* d_ranges[2m], d_ranges[2m+1] = [first, last+1) non-zero bin of filter m; lets the fused
 * kernel skip the zeros of the triangular filters (optional argument above; NULL = dense). */
ADVOC_API int advoc_mel_ranges(const float* d_mel_fb, int nmels, int bins, int* d_ranges, void* stream);

/* y[r, n] = sum_k x[r, k] * w[n, k]   (x [rows,K], w [N,K], y [rows,N], all f32).
 * replaces: models/advoc/spectral_util.py:29-32 (mag -> linear mel, w = mel_fb) and
 * :34-43 (linear mel -> mag, w = pinv(mel_fb)); scripts/spectrogram_advoc.py:21.
 * If pow10_scale != 0 the input is first mapped x -> 10^((x*(-min_db) + min_db + ref_db)/20): the dB
 * de-normalisation of advoc/spectral.py:367-369 (defaults -100 / 20: scripts/spectrogram_advoc.py:19-20;
 * tacotron2 features use -40 / 20). */
ADVOC_API int advoc_matmul_lastdim_f32(const float* d_x, const float* d_w, float* d_y, long rows, int K,
                             int N, int pow10_scale, float min_level_db, float ref_level_db, void* stream);

/* Spectrogram inversion ("next" rows of SURVEY.md section 8(f)).  lws istft with
 * perfectrec=False: out = overlap-add( irfft(X[m]) * window ), length (frames-1)*hop + nfft.
 * Step 1: windowed inverse frames, d_spec_c64 [batch, frames, nfft/2+1] interleaved (re,im) ->
 * d_frames [batch, frames, nfft].  nfft must be a power of two.
 * replaces: lws istft, advoc/spectral.py:303,307,322. */
ADVOC_API int advoc_istft_frames_f32(const float* d_spec_c64, int batch, int frames, int nfft, int nhop,
                                     const float* d_window, const float* d_twiddle, float* d_frames,
                                     void* stream);
/* Step 2: atomics-free overlap-add, d_out [batch, (frames-1)*nhop + nfft]. */
ADVOC_API int advoc_overlap_add_f32(const float* d_frames, int batch, int frames, int nfft, int nhop,
                                    float* d_out, void* stream);
/* One Griffin-Lim iteration fused in one kernel: S = stft(d_wave); X = d_mag * S/|S|; windowed
 * inverse frames of X -> d_frames (follow with advoc_overlap_add_f32 to get the next estimate).
 * d_wave [batch, nwave], nwave == (frames-1)*nhop + nfft; d_mag [batch, frames, nfft/2+1].
 * replaces: the loop body of advoc/spectral.py:304-307. */
ADVOC_API int advoc_griffin_lim_iter_f32(const float* d_wave, long nwave, const float* d_mag, int batch,
                                         int frames, int nfft, int nhop, const float* d_window,
                                         const float* d_twiddle, float* d_frames, void* stream);

/* ------------------------------------------------------------------------- *
 * Convolution stacks
 * ------------------------------------------------------------------------- */

/* Geometry of one 2-D convolution on NHWC data.  For a transposed convolution the
 * struct describes the *forward* convolution it is the input-gradient of
 * (SURVEY appendix B rule 2): (N,H,W,Cin) is the big (output-of-deconv) side. */
typedef struct advoc_conv_desc {
  int N, H, W, Cin;     /* conv input                                          */
  int Cout;             /* conv output channels                                */
  int kh, kw;           /* 4x4 (AdVoc) or 5x5 (MelspecGAN)                     */
  int sh, sw;           /* strides                                             */
  int pad_t, pad_l;     /* leading zero padding (trailing is implied by Ho/Wo) */
  int Ho, Wo;           /* conv output spatial size                            */
  int math;             /* advoc_math                                          */
} advoc_conv_desc;

/* What happens to an output tile before it is stored.
 *   y0 = act0(acc + bias) [* dropout]  -> d_out0 at channel offset c_off0 of a
 *                                         buffer with ld0 channels per pixel
 *   y1 = act1(acc + bias) [* dropout]  -> d_out1 (optional second consumer: the
 *                                         reference applies lrelu for the next
 *                                         encoder and relu for the decoder skip to
 *                                         the same tensor, advoc_model.py:109,138)
 * Only the first `store_w` output columns are written (the reference drops the last
 * column of every decoder output, advoc_model.py:137,154,156); 0 = all.
 * d_dropout_mask (optional, uint8 0/1, dense [N,Ho,Wo,Cout] of the *stored* extent)
 * multiplies by mask/keep_prob (tf.nn.dropout, advoc_model.py:144-149); if NULL and
 * keep_prob < 1 a counter-based generator keyed by (seed, element index) is used.
 * round_tf32: round stored values to TF32 (RNA) so the consuming tensor-core layer
 * reads exactly-representable operands.
 * Backward-pass extensions (apply to out0 only; leave zero in the forward pass):
 *   d_gate != NULL: y0 *= act'(g) * (n < gate_split ? gate_scale0 : gate_scale1), where g is the
 *     STORED ACTIVATED value of the tensor whose pre-activation gradient is being produced
 *     (gate_act = LRELU: g > 0 ? 1 : alpha; RELU: g > 0 ? 1 : 0) -- the ReluGrad / Maximum-grad /
 *     dropout-grad nodes autodiff adds for advoc_model.py:87,138,146-149 fused into the
 *     input-gradient GEMM.  A dropped element stored 0, so relu's gate also gates dropout and
 *     gate_scale0 = 1/keep_prob restores its scale on the decoder slice [0, gate_split).
 *   accumulate != 0: out0 += y0 (second gradient contribution of a skip connection). */
typedef struct advoc_epilogue {
  const float* d_bias;       /* [Cout] or NULL */
  int act0, act1;
  float alpha;               /* lrelu slope */
  float* d_out0; int ld0, c_off0;
  float* d_out1; int ld1, c_off1;   /* d_out1 may be NULL */
  int store_w;
  const uint8_t* d_dropout_mask;
  float keep_prob;           /* 1.0 = no dropout */
  uint64_t seed;
  int round_tf32;
  int accumulate;
  const float* d_gate; int ld_gate, c_off_gate;
  int gate_act, gate_split;
  float gate_scale0, gate_scale1;
  /* Optional dropout step counter in DEVICE memory.  When non-NULL the generator is keyed by
   * (*d_seed) * 0x9E3779B1 + seed instead of seed (seed then is the layer's salt), so a captured CUDA
   * graph draws a fresh mask on every replay once the caller bumps the counter (the reference
   * resamples dropout on every sess.run, advoc_model.py:144-149). */
  const uint64_t* d_seed;
  /* advoc_dtype of d_out0 / d_out1: ADVOC_DT_F16 stores IEEE half (round-to-nearest-even); ld / c_off
   * stay in elements.  Forward pass only (no gate / accumulate). */
  int out0_dtype, out1_dtype;
  /* Extra (never written) pixels at the end of every image row of d_out0: pixel (h, w) of image b is
   * stored at pixel index (b*Ho + h) * (Wo + out0_row_pad) + w.  The fp16 generator pads encoder_1's
   * odd-width output by one zero pixel so that the next layer can read PAIRS of 32-channel pixels as
   * 64-channel k-blocks (advoc_b200/nets.py).  Supported by the thin-input convolution only; 0 elsewhere. */
  int out0_row_pad;
  int reserved0;
} advoc_epilogue;

/* Re-pack a TF-layout filter [kh*kw, A, B] (A,B = Cin,Cout for conv; Cout,Cin for
 * conv_transpose) into [kh*kw, B, A] (transpose != 0) or copy it, optionally rounding every
 * value to TF32 (round-to-nearest) so the tensor-core path reads exactly representable
 * operands.  replaces: nothing in the reference (TF/cuDNN pick their own filter layouts). */
/* mode: 0 copy, 1 round to TF32, 2 convert to fp16 (d_packed then holds IEEE half, for
 * ADVOC_MATH_F16 layers). */
ADVOC_API int advoc_pack_filter(const float* d_w, void* d_packed, int taps, int A, int B,
                                int transpose, int mode, void* stream);

/* Reads and clears the library's device-side debug word: non-zero means a pipeline barrier of a
 * tcgen05 kernel timed out (1 producer, 2 MMA issuer, 3 epilogue).  Synchronises the device. */
ADVOC_API int advoc_debug_flags(unsigned int* out);

/* The same code read from its pinned host mirror: no CUDA call, no synchronisation, does not
 * clear.  The host engines (infer.MelToMag, train.TrainEngine, melspecgan.MelspecGAN, bench.py) call
 * it at their own synchronisation points and raise RuntimeError when it is non-zero, so that a timed
 * out launch can never pass stale buffers on as results. */
ADVOC_API int advoc_debug_peek(unsigned int* out);

/* y = conv2d(x, w) ; x [N,H,W,Cin] (pixel stride ld_x >= Cin, channel offset 0).
 * d_w: on the CUDA-core path (advoc_conv2d_path == ADVOC_MATH_FP32) the TF layout HWIO;
 * on the tcgen05 path the K-major pack [taps, Cout, Cin] made by advoc_pack_filter.
 * replaces: tf.layers.conv2d call sites models/advoc/advoc_model.py:27-32, :46-51. */
/* d_x / d_w are float, or IEEE half when d->math == ADVOC_MATH_F16 (ld_x in elements either way). */
ADVOC_API int advoc_conv2d_fwd(const advoc_conv_desc* d, const void* d_x, int ld_x, const void* d_w,
                     const advoc_epilogue* ep, void* stream);

/* y = conv2d_transpose(x, w); x [N,Ho,Wo,Cout] (ld_x), output [N,H,W,Cin] in desc naming.
 * d_w: TF layout HWOI [kh,kw,Cin(desc),Cout(desc)] on both paths (it already is K-major for
 * this GEMM; pass a TF32-rounded copy on the tcgen05 path).
 * replaces: tf.layers.conv2d_transpose models/advoc/advoc_model.py:65-69. */
ADVOC_API int advoc_conv2d_transpose_fwd(const advoc_conv_desc* d, const void* d_x, int ld_x,
                               const void* d_w, const advoc_epilogue* ep, void* stream);

/* ------------------------------------------------------------------------- *
 * Backward pass and optimiser (the gradient graph of advoc_model.py:238-257)
 *
 * Input gradients reuse the two forward entry points: the input gradient of a conv is
 * advoc_conv2d_transpose_fwd with the same desc and the conv's own HWIO filter (already
 * K-major for that GEMM); the input gradient of a conv_transpose is advoc_conv2d_fwd over the
 * big side.  The activation / dropout derivative and the skip-connection sum are fused into
 * their epilogues (advoc_epilogue.d_gate / accumulate).
 * ------------------------------------------------------------------------- */

/* Filter gradient.  d_dw [kh*kw][Cin][Cout] (+=, caller zeroes it): for a conv this IS the
 * TF layout HWIO; for a conv_transpose described by `d` (big side = its output) it is HWOI.
 *   d_dw[tap][cb][cs] += sum_{n,oh,ow} big[n, oh*sh-pad_t+kh, ow*sw-pad_l+kw, cb] * small[n,oh,ow,cs]
 * d_big [N,H,W,*] (pixel stride ld_big, Cin channels), d_small [N,Ho,Wo,*] (ld_small, Cout).
 * Tensor-core paths (wgrad_tc.cu, wgrad_thin_tc.cu: every layer of the AdVoc / MelspecGAN nets unless the desc asks
 * for ADVOC_MATH_FP32) reduce their pixel splits in a fixed order: bit-identical from run to run.  The CUDA-core
 * fallbacks accumulate with fp32 atomics (sum order not fixed).
 * replaces: Conv2DBackpropFilter built by opt.minimize, advoc_model.py:254-257. */
ADVOC_API int advoc_conv2d_wgrad(const advoc_conv_desc* d, const float* d_big, int ld_big,
                                 const float* d_small, int ld_small, float* d_dw, void* stream);

/* d_dbias[c] += sum over pixels of d_dy[pix*ld_dy + c]   (BiasAddGrad). */
ADVOC_API int advoc_bias_grad(const float* d_dy, int ld_dy, long pixels, int channels,
                              float* d_dbias, void* stream);

/* GAN log-loss on sigmoid probabilities with gradient seeds wrt the LOGITS (eps = 1e-12).
 * mode 0: discriminator loss  mean(-(log(p_real+eps)+log(1-p_fake+eps)))
 * mode 1: generator GAN loss  mean(-log(p_fake+eps)) * weight
 * d_loss[0] += loss (atomic).  p_real and the gradient outputs may be NULL.
 * replaces: models/advoc/advoc_model.py:201,238-239 and their autodiff. */
ADVOC_API int advoc_gan_logloss(const float* d_p_real, const float* d_p_fake, long n, int mode,
                                float weight, float* d_loss, float* d_dlogit_real,
                                float* d_dlogit_fake, void* stream);

/* d_loss[0] += weight*mean|target-gen| ; d_dgen[i] (=|+=) weight*sign(gen-target)/n.
 * gen is read with pixel stride ld_gen at channel c_off_gen (it lives in the discriminator's
 * 2-channel input buffer).  replaces: models/advoc/advoc_model.py:240,243. */
ADVOC_API int advoc_l1_loss(const float* d_gen, int ld_gen, int c_off_gen, const float* d_target,
                            long n, float weight, float* d_loss, float* d_dgen, int accumulate,
                            void* stream);

/* TF1 Adam over one flat fp32 buffer (epsilon outside the bias correction):
 *   m=b1*m+(1-b1)g ; v=b2*v+(1-b2)g^2 ; p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
 * grad_scale multiplies g first (1/world_size after a sum-allreduce).
 * replaces: tf.train.AdamOptimizer(0.0002, 0.5).minimize, advoc_model.py:250-257. */
ADVOC_API int advoc_adam_tf_step(float* d_p, const float* d_g, float* d_m, float* d_v, long n,
                                 float lr, float beta1, float beta2, float eps, long t,
                                 float grad_scale, void* stream);

/* The same update with the bias-corrected step size lr*sqrt(1-b2^t)/(1-b1^t) read from device memory
 * (one float at d_lr_t), so that a captured CUDA graph of the optimiser step stays valid while t advances.
 * replaces: the AdamOptimizer train ops of models/melspecgan/train.py:117-139 when replayed from a graph. */
ADVOC_API int advoc_adam_tf_step_dev(float* d_p, const float* d_g, float* d_m, float* d_v, long n,
                                     const float* d_lr_t, float beta1, float beta2, float eps,
                                     float grad_scale, void* stream);

/* Which kernel family a call with this geometry takes: ADVOC_MATH_TF32 (tcgen05 implicit
 * GEMM) or ADVOC_MATH_FP32 (CUDA-core kernel).  transposed != 0 asks about
 * advoc_conv2d_transpose_fwd.  Pure host logic. */
ADVOC_API int advoc_conv2d_path(const advoc_conv_desc* d, int ld_x, int transposed);

/* Which kernel a FORWARD call (no gate / accumulate in the epilogue) with this geometry launches -- input-gradient
 * calls may take the patch kernel where a forward call takes the per-tap one: 0 CUDA-core fp32 (conv_simt.cu /
 * conv_direct.cu), 1 tcgen05 per-tap im2col kernel (conv_tc.cu), 2 tcgen05 persistent patch kernel
 * (conv_p2d.cu), 3 tcgen05 transposed conv to one channel (deconv_one_tc.cu), 4 tcgen05 conv from one channel
 * (conv_one_in_tc.cu; taken when the epilogue is a plain forward one, else 0).  store_w as in advoc_epilogue (0 = full width).  Host-side query, used by bench.py
 * to attribute time to kernels.  replaces: nothing in the reference. */
ADVOC_API int advoc_conv2d_kernel(const advoc_conv_desc* d, int ld_x, int transposed, int store_w);

/* Accumulator tile width N (the BN template parameter: conv_tc_kernel<BN,...> / conv_p2d_kernel<BN>)
 * of the kernel advoc_conv2d_kernel names for this geometry; 16 for the to-one-channel kernel, 0 for
 * the CUDA-core kernels.  Host-side query: the parity tests use it to assert which instantiation
 * they covered.  replaces: nothing in the reference. */
ADVOC_API int advoc_conv2d_tile_n(const advoc_conv_desc* d, int ld_x, int transposed, int store_w);

/* ------------------------------------------------------------------------- *
 * MelspecGAN building blocks (models/melspecgan/conv2d.py, models/melspecgan/train.py).
 * The 5x5 stride-2 convolutions go through advoc_conv2d_fwd / _transpose_fwd / _wgrad.
 * ------------------------------------------------------------------------- */

/* C[M,N] (+)= op(A) * op(B) (+ bias[N]); row-major fp32, op = identity or transpose (trans_* != 0
 * means the operand is stored transposed: A as [K,M], B as [N,K]).
 * replaces: tf.matmul + tf.nn.bias_add of dense_layer (conv2d.py:4-14) and their gradients. */
ADVOC_API int advoc_gemm_f32(const float* d_a, int lda, const float* d_b, int ldb, float* d_c, int ldc,
                             int M, int N, int K, int trans_a, int trans_b, int accumulate,
                             const float* d_bias, void* stream);

/* d_stats[0:C] += sum_pixels x, d_stats[C:2C] += sum_pixels x^2 (caller zeroes d_stats).
 * replaces: the batch moments of tf.layers.batch_normalization(training=True), conv2d.py:108,178. */
ADVOC_API int advoc_bn_stats(const float* d_x, int ld, long pixels, int C, float* d_stats, void* stream);

/* y = act(gamma * (x - mean) / sqrt(var + eps) + beta), act in {none, lrelu, relu}; mean / biased
 * var from d_stats.  replaces: tf.layers.batch_normalization + tf.nn.relu / leaky_relu,
 * conv2d.py:116-137,186-201. */
ADVOC_API int advoc_bn_apply(const float* d_x, int ldx, long pixels, int C, const float* d_stats,
                             const float* d_gamma, const float* d_beta, float eps, int act, float alpha,
                             float* d_y, int ldy, int round_tf32, void* stream);

/* Inference-mode batch norm (training=False): y = act(gamma * (x - mean) / sqrt(var + eps) + beta) with the
 * given per-channel moving averages.  replaces: tf.layers.batch_normalization(training=False) + relu of
 * the generator's inference graph, conv2d.py:108,116-137 as built by models/melspecgan/infer.py:14-17. */
ADVOC_API int advoc_bn_inference(const float* d_x, int ldx, long pixels, int C, const float* d_mean,
                                 const float* d_var, const float* d_gamma, const float* d_beta, float eps,
                                 int act, float alpha, float* d_y, int ldy, int round_tf32, void* stream);

/* moving_mean / moving_var <- moving - (moving - batch) * (1 - momentum) from the (sum, sum of squares) in
 * d_stats; the batch variance carries Bessel's correction pixels / (pixels - 1) like TF's fused op.
 * replaces: the UPDATE_OPS of tf.layers.batch_normalization run under control_dependencies,
 * conv2d.py:143-148 (momentum 0.99). */
ADVOC_API int advoc_bn_moving_update(const float* d_stats, long pixels, int C, float momentum,
                                     float* d_moving_mean, float* d_moving_var, void* stream);

/* Backward of advoc_bn_apply: g = dy * act'(y); d_red[0:C] += sum g (dbeta), d_red[C:2C] += sum g*xhat
 * (dgamma) (caller zeroes d_red); dx = gamma/sqrt(var+eps) * (g - dbeta/M - xhat*dgamma/M).
 * replaces: the autodiff of the same ops (opt.minimize, train.py:137-139). */
ADVOC_API int advoc_bn_backward(const float* d_dy, int lddy, const float* d_y, int ldy, const float* d_x,
                                int ldx, long pixels, int C, const float* d_stats, const float* d_gamma,
                                float eps, int act, float alpha, float* d_red, float* d_dx, int lddx,
                                int round_tf32, void* stream);

/* dx = dy * (1 - y^2).  replaces: the gradient of tf.nn.tanh, conv2d.py:141. */
ADVOC_API int advoc_tanh_backward(const float* d_dy, const float* d_y, float* d_dx, long n, void* stream);

/* Losses on the critic's logits with gradient seeds (mean over n):
 *   mode 0 dcgan D: (xent(fake,0) + xent(real,1)) / 2     mode 1 dcgan G: xent(fake,1)
 *   mode 2 wgan  D: mean(fake) - mean(real) (penalty term not included)   mode 3 wgan G: -mean(fake)
 * replaces: models/melspecgan/train.py:76-97. */
ADVOC_API int advoc_gan_logit_loss(const float* d_real, const float* d_fake, int n, int mode, float* d_loss,
                                   float* d_dreal, float* d_dfake, void* stream);

/* WGAN-GP seed (models/melspecgan/train.py:105-109): per sample s = ||g_b||_2 over n values,
 * d_loss[0] += lambda/B sum (s-1)^2, d_u = d penalty / d g.  d_g is the critic's input gradient. */
ADVOC_API int advoc_gp_seed(const float* d_g, int batch, long n, float lambda, float* d_loss, float* d_u,
                            int round_tf32, void* stream);

/* Second-order terms of a batch-normalised leaky-ReLU layer for the penalty's parameter gradient
 * (the double backward TF builds for tf.gradients inside the loss, train.py:106).  d_v = adjoint of
 * the first backward's BN-input gradient, d_dy / d_y / d_x = the first backward's incoming gradient,
 * the layer output and the BN input, all dense [pixels, C].  Outputs: d_sums [5C] (zeroed by the
 * caller; sum v, sum a, sum v a, sum v xhat, sum a xhat with a = dy * act'(y)), d_vz = adjoint of
 * d_dy (masked, feeds the next layer's forward conv), d_xbar = adjoint of d_x. */
ADVOC_API int advoc_bn_gp(const float* d_v, const float* d_dy, const float* d_y, const float* d_x, long pixels,
                          int C, const float* d_stats, const float* d_gamma, float eps, float alpha,
                          float* d_sums, float* d_vz, float* d_xbar, int round_tf32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADVOC_B200_H_ */
