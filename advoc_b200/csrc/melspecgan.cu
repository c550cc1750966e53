// Kernels the MelspecGAN stacks need beyond the convolutions (models/melspecgan/conv2d.py,
// models/melspecgan/train.py): dense layers, batch normalisation with batch statistics (forward
// and backward, the ReLU / leaky-ReLU that follows it fused in), tanh backward and the GAN losses
// on the critic's logits.  All exact fp32 on CUDA cores: these ops are small next to the 5x5
// convolutions ([64, 4..32, 5..40, 64..512] activations).
#include "common.cuh"

namespace advoc {

namespace {

int grid_for(long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  const long cap = (long)sm_count() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// C[M,N] (+)= op(A) * op(B) (+ bias[N]);  row-major, op = identity or transpose.
// 64x64 tile, 16-deep slabs, 4x4 register tile per thread.
// ---------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                   int ldb, float* __restrict__ C, int ldc, int M, int N, int K,
                                                   int tA, int tB, int accumulate, const float* __restrict__ bias) {
  __shared__ float As[GK][GT + 4];
  __shared__ float Bs[GK][GT + 4];
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int i = threadIdx.x; i < GT * GK; i += 256) {
      // A slab: element (m, k)
      const int k = tA ? i / GT : i % GK, m = tA ? i % GT : i / GK;
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = tA ? __ldg(A + (size_t)gk * lda + gm) : __ldg(A + (size_t)gm * lda + gk);
      As[k][m] = v;
    }
    for (int i = threadIdx.x; i < GT * GK; i += 256) {
      // B slab: element (k, n)
      const int k = tB ? i % GK : i / GT, n = tB ? i / GK : i % GT;
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = tB ? __ldg(B + (size_t)gn * ldb + gk) : __ldg(B + (size_t)gk * ldb + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      float* dst = C + (size_t)m * ldc + n;
      if (accumulate) v += *dst;
      *dst = v;
    }
  }
}

// N == 1 (the critic's output layer): one warp per row of A, lanes across K.
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ A, int lda, const float* __restrict__ b,
                                                   int ldb, float* __restrict__ c, int ldc, int M, int K,
                                                   int accumulate, const float* __restrict__ bias) {
  const int lane = threadIdx.x & 31;
  const int m = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (m >= M) return;
  const float* a = A + (size_t)m * lda;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(__ldg(a + k), __ldg(b + (size_t)k * ldb), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    float v = s + (bias ? __ldg(bias) : 0.f);
    if (accumulate) v += c[(size_t)m * ldc];
    c[(size_t)m * ldc] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// batch normalisation over [pixels, C] (pixel stride ld), batch statistics
// ---------------------------------------------------------------------------------------------
// stats[0:C] += sum x, stats[C:2C] += sum x^2   (caller zeroes stats)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int ld, long pixels, int C,
                                                       float* __restrict__ stats) {
  extern __shared__ float part[];  // [2C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const int lanes = C < 256 ? C : 256;            // threads per pixel row
  const int ppb = 256 / lanes;
  const int c0 = threadIdx.x % lanes;
  if ((int)(threadIdx.x / lanes) < ppb) {
    for (int c = c0; c < C; c += lanes) {
      float s = 0.f, q = 0.f;
      for (long p = (long)blockIdx.x * ppb + threadIdx.x / lanes; p < pixels; p += (long)gridDim.x * ppb) {
        const float v = __ldg(x + (size_t)p * ld + c);
        s += v;
        q = fmaf(v, v, q);
      }
      atomicAdd(&part[c], s);
      atomicAdd(&part[C + c], q);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(stats + i, part[i]);
}

__device__ __forceinline__ void bn_moments(const float* stats, int C, int c, float inv_m, float eps, float* mean,
                                           float* invstd) {
  const float m = __ldg(stats + c) * inv_m;
  const float var = fmaxf(__ldg(stats + C + c) * inv_m - m * m, 0.f);
  *mean = m;
  *invstd = rsqrtf(var + eps);
}

// y = act(gamma * (x - mean) * invstd + beta)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, int ldx, long pixels, int C,
                                                       const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, int act, float alpha,
                                                       float* __restrict__ y, int ldy, int round) {
  const long total = pixels * C;
  const float inv_m = 1.f / (float)pixels;
  const ActLin a = act_linear(act, alpha);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long p = i / C;
    const int c = (int)(i - p * C);
    float mean, invstd;
    bn_moments(stats, C, c, inv_m, eps, &mean, &invstd);
    float v = (__ldg(x + (size_t)p * ldx + c) - mean) * invstd * __ldg(gamma + c) + __ldg(beta + c);
    v = apply_lin(v, a);
    if (round) v = round_tf32(v);
    y[(size_t)p * ldy + c] = v;
  }
}

// inference-mode batch norm (training=False): per-channel mean / variance are given (the moving averages)
__global__ void __launch_bounds__(256) bn_infer_kernel(const float* __restrict__ x, int ldx, long pixels, int C,
                                                       const float* __restrict__ mean, const float* __restrict__ var,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int act, float alpha, float* __restrict__ y, int ldy,
                                                       int round) {
  const long total = pixels * C;
  const ActLin a = act_linear(act, alpha);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long p = i / C;
    const int c = (int)(i - p * C);
    const float scale = rsqrtf(__ldg(var + c) + eps) * __ldg(gamma + c);
    float v = (__ldg(x + (size_t)p * ldx + c) - __ldg(mean + c)) * scale + __ldg(beta + c);
    v = apply_lin(v, a);
    if (round) v = round_tf32(v);
    y[(size_t)p * ldy + c] = v;
  }
}

// moving <- moving - (moving - batch) * (1 - momentum); the batch variance carries Bessel's correction
// M / (M - 1), as the fused batch-norm op reports it to the moving-average update
__global__ void __launch_bounds__(256) bn_moving_update_kernel(const float* __restrict__ stats, long pixels, int C,
                                                               float momentum, float* __restrict__ moving_mean,
                                                               float* __restrict__ moving_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv_m = 1.f / (float)pixels;
  const float m = stats[c] * inv_m;
  float var = fmaxf(stats[C + c] * inv_m - m * m, 0.f);
  if (pixels > 1) var *= (float)pixels / (float)(pixels - 1);
  const float d = 1.f - momentum;
  moving_mean[c] -= (moving_mean[c] - m) * d;
  moving_var[c] -= (moving_var[c] - var) * d;
}

// g = dy * act'(y);  red[0:C] += sum g (= dbeta), red[C:2C] += sum g * xhat (= dgamma)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, int lddy,
                                                            const float* __restrict__ y, int ldy,
                                                            const float* __restrict__ x, int ldx, long pixels, int C,
                                                            const float* __restrict__ stats, float eps, int act,
                                                            float alpha, float* __restrict__ red) {
  extern __shared__ float part[];  // [2C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const float inv_m = 1.f / (float)pixels;
  const float neg = act == ADVOC_ACT_LRELU ? alpha : (act == ADVOC_ACT_RELU ? 0.f : 1.f);
  const int lanes = C < 256 ? C : 256;
  const int ppb = 256 / lanes;
  const int c0 = threadIdx.x % lanes;
  if ((int)(threadIdx.x / lanes) < ppb) {
    for (int c = c0; c < C; c += lanes) {
      float mean, invstd;
      bn_moments(stats, C, c, inv_m, eps, &mean, &invstd);
      float s = 0.f, q = 0.f;
      for (long p = (long)blockIdx.x * ppb + threadIdx.x / lanes; p < pixels; p += (long)gridDim.x * ppb) {
        float g = __ldg(dy + (size_t)p * lddy + c);
        if (!(__ldg(y + (size_t)p * ldy + c) > 0.f)) g *= neg;
        const float xh = (__ldg(x + (size_t)p * ldx + c) - mean) * invstd;
        s += g;
        q = fmaf(g, xh, q);
      }
      atomicAdd(&part[c], s);
      atomicAdd(&part[C + c], q);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(red + i, part[i]);
}

// dx = gamma * invstd * (g - dbeta/M - xhat * dgamma/M)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, int lddy,
                                                           const float* __restrict__ y, int ldy,
                                                           const float* __restrict__ x, int ldx, long pixels, int C,
                                                           const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, float eps, int act,
                                                           float alpha, const float* __restrict__ red,
                                                           float* __restrict__ dx, int lddx, int round) {
  const long total = pixels * C;
  const float inv_m = 1.f / (float)pixels;
  const float neg = act == ADVOC_ACT_LRELU ? alpha : (act == ADVOC_ACT_RELU ? 0.f : 1.f);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long p = i / C;
    const int c = (int)(i - p * C);
    float mean, invstd;
    bn_moments(stats, C, c, inv_m, eps, &mean, &invstd);
    float g = __ldg(dy + (size_t)p * lddy + c);
    if (!(__ldg(y + (size_t)p * ldy + c) > 0.f)) g *= neg;
    const float xh = (__ldg(x + (size_t)p * ldx + c) - mean) * invstd;
    float v = __ldg(gamma + c) * invstd * (g - __ldg(red + c) * inv_m - xh * __ldg(red + C + c) * inv_m);
    if (round) v = round_tf32(v);
    dx[(size_t)p * lddx + c] = v;
  }
}

__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                       float* __restrict__ dx, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float t = __ldg(y + i);
    dx[i] = __ldg(dy + i) * (1.f - t * t);
  }
}

// losses on logits (one block).  mode 0 dcgan D, 1 dcgan G, 2 wgan D (without the penalty), 3 wgan G
__device__ __forceinline__ float softplus(float v) { return fmaxf(v, 0.f) + log1pf(__expf(-fabsf(v))); }

__global__ void __launch_bounds__(256) gan_logit_loss_kernel(const float* __restrict__ real,
                                                             const float* __restrict__ fake, int n, int mode,
                                                             float* __restrict__ loss, float* __restrict__ dreal,
                                                             float* __restrict__ dfake) {
  __shared__ float red[256];
  float s = 0.f;
  const float inv = 1.f / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float f = __ldg(fake + i);
    const float sf = 1.f / (1.f + __expf(-f));
    if (mode == 0) {
      // (xent(fake, 0) + xent(real, 1)) / 2 ;  xent(l, 0) = softplus(l), xent(l, 1) = softplus(-l)
      const float r = __ldg(real + i);
      const float sr = 1.f / (1.f + __expf(-r));
      s += 0.5f * (softplus(f) + softplus(-r));
      dfake[i] = 0.5f * sf * inv;
      dreal[i] = 0.5f * (sr - 1.f) * inv;
    } else if (mode == 1) {
      s += softplus(-f);
      dfake[i] = (sf - 1.f) * inv;
    } else if (mode == 2) {
      const float r = __ldg(real + i);
      s += f - r;
      dfake[i] = inv;
      dreal[i] = -inv;
    } else {
      s -= f;
      dfake[i] = -inv;
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] * inv;
}

// ---------------------------------------------------------------------------------------------
// WGAN-GP (models/melspecgan/train.py:99-109): gradient penalty seed and the second-order terms
// of batch normalisation that its parameter gradient needs (double backward).
// ---------------------------------------------------------------------------------------------
// One block per sample: s = ||g_b||_2, loss += lambda/B (s-1)^2, u_b = lambda * 2/B * (s-1)/s * g_b
__global__ void __launch_bounds__(256) gp_seed_kernel(const float* __restrict__ g, int B, long n, float lambda,
                                                      float* __restrict__ loss, float* __restrict__ u, int round) {
  __shared__ float red[256];
  const float* gb = g + (size_t)blockIdx.x * n;
  float s = 0.f;
  for (long i = threadIdx.x; i < n; i += blockDim.x) { const float v = __ldg(gb + i); s = fmaf(v, v, s); }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float norm = sqrtf(red[0]);
  const float coef = lambda * 2.f / (float)B * (norm - 1.f) / fmaxf(norm, 1e-30f);
  if (threadIdx.x == 0) atomicAdd(loss, lambda / (float)B * (norm - 1.f) * (norm - 1.f));
  float* ub = u + (size_t)blockIdx.x * n;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    float v = coef * __ldg(gb + i);
    if (round) v = round_tf32(v);
    ub[i] = v;
  }
}

// Per channel, with a = dy * act'(y) (the first backward's gradient at the BN output) and
// v = the adjoint of the first backward's BN input gradient:
//   sums[0:C] = sum v, [C:2C] = sum a, [2C:3C] = sum v a, [3C:4C] = sum v xhat, [4C:5C] = sum a xhat
__global__ void __launch_bounds__(256) bn_gp_reduce_kernel(const float* __restrict__ v, const float* __restrict__ dy,
                                                           const float* __restrict__ y, const float* __restrict__ x,
                                                           long pixels, int C, const float* __restrict__ stats,
                                                           float eps, float alpha, float* __restrict__ sums) {
  extern __shared__ float part[];  // [5C]
  for (int i = threadIdx.x; i < 5 * C; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const float inv_m = 1.f / (float)pixels;
  const int lanes = C < 256 ? C : 256;
  const int ppb = 256 / lanes;
  const int c0 = threadIdx.x % lanes;
  if ((int)(threadIdx.x / lanes) < ppb) {
    for (int c = c0; c < C; c += lanes) {
      float mean, invstd;
      bn_moments(stats, C, c, inv_m, eps, &mean, &invstd);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
      for (long p = (long)blockIdx.x * ppb + threadIdx.x / lanes; p < pixels; p += (long)gridDim.x * ppb) {
        const size_t o = (size_t)p * C + c;
        const float vv = __ldg(v + o);
        float a = __ldg(dy + o);
        if (!(__ldg(y + o) > 0.f)) a *= alpha;
        const float xh = (__ldg(x + o) - mean) * invstd;
        s0 += vv; s1 += a; s2 = fmaf(vv, a, s2); s3 = fmaf(vv, xh, s3); s4 = fmaf(a, xh, s4);
      }
      atomicAdd(&part[c], s0); atomicAdd(&part[C + c], s1); atomicAdd(&part[2 * C + c], s2);
      atomicAdd(&part[3 * C + c], s3); atomicAdd(&part[4 * C + c], s4);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 5 * C; i += blockDim.x) atomicAdd(sums + i, part[i]);
}

// vz   = act'(y) * gamma/sigma * (v - mean(v) - xhat mean(v xhat))     (adjoint of dy; feeds the next conv)
// xbar = 1/sigma (G - mean(G) - xhat mean(G xhat)) + Gs xhat / M        (adjoint of the forward x)
//   with G = -gamma/sigma (v A + a V), A = mean(a xhat), V = mean(v xhat), Gs = -S/sigma,
//   S = gamma/sigma (sum v a - M mean(v) mean(a) - M V A)
__global__ void __launch_bounds__(256) bn_gp_apply_kernel(const float* __restrict__ v, const float* __restrict__ dy,
                                                          const float* __restrict__ y, const float* __restrict__ x,
                                                          long pixels, int C, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma, float eps, float alpha,
                                                          const float* __restrict__ sums, float* __restrict__ vz,
                                                          float* __restrict__ xbar, int round) {
  const long total = pixels * C;
  const float M = (float)pixels, inv_m = 1.f / M;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float mean, invstd;
    bn_moments(stats, C, c, inv_m, eps, &mean, &invstd);
    const float gs = __ldg(gamma + c) * invstd;
    const float mv = __ldg(sums + c) * inv_m, ma = __ldg(sums + C + c) * inv_m, sva = __ldg(sums + 2 * C + c);
    const float V = __ldg(sums + 3 * C + c) * inv_m, A = __ldg(sums + 4 * C + c) * inv_m;
    const float vv = __ldg(v + i);
    const bool pos = __ldg(y + i) > 0.f;
    const float a = __ldg(dy + i) * (pos ? 1.f : alpha);
    const float xh = (__ldg(x + i) - mean) * invstd;
    float z = gs * (vv - mv - xh * V) * (pos ? 1.f : alpha);
    if (round) z = round_tf32(z);
    vz[i] = z;
    const float S = gs * (sva - M * mv * ma - M * V * A);
    const float G = -gs * (vv * A + a * V);
    const float mG = -gs * (A * mv + V * ma), mGx = -2.f * gs * A * V;
    float xb = invstd * (G - mG - xh * mGx) + (-S * invstd) * xh * inv_m;
    xbar[i] = xb;
  }
}

}  // namespace
}  // namespace advoc

using namespace advoc;

extern "C" int advoc_gemm_f32(const float* d_a, int lda, const float* d_b, int ldb, float* d_c, int ldc, int M,
                              int N, int K, int trans_a, int trans_b, int accumulate, const float* d_bias,
                              void* stream) {
  ADVOC_REQUIRE(d_a && d_b && d_c, ADVOC_BAD_ARG, "NULL matrix");
  ADVOC_REQUIRE(M > 0 && N > 0 && K > 0 && lda > 0 && ldb > 0 && ldc >= N, ADVOC_BAD_SHAPE, "bad gemm shape");
  if (N == 1 && !trans_a && K >= 256) {
    // b[k] = B[k][0] (ldb) or, stored transposed [1][K], B[0][k]
    gemv_kernel<<<(M * 32 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_a, lda, d_b, trans_b ? 1 : ldb, d_c, ldc, M, K, accumulate, d_bias);
    count_launch();
    ADVOC_CHECK_CUDA(cudaGetLastError());
    return ADVOC_OK;
  }
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, 1);
  gemm_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_a, lda, d_b, ldb, d_c, ldc, M, N, K,
                                                                        trans_a, trans_b, accumulate, d_bias);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bn_stats(const float* d_x, int ld, long pixels, int C, float* d_stats, void* stream) {
  ADVOC_REQUIRE(d_x && d_stats && pixels > 0 && C > 0 && C <= 4096 && ld >= C, ADVOC_BAD_ARG, "bad bn_stats arguments");
  const int ppb = 256 / (C < 256 ? C : 256);
  int blocks = grid_for(pixels, ppb * 8);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  bn_stats_kernel<<<blocks, 256, 2 * C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(d_x, ld, pixels, C,
                                                                                                 d_stats);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bn_apply(const float* d_x, int ldx, long pixels, int C, const float* d_stats,
                              const float* d_gamma, const float* d_beta, float eps, int act, float alpha, float* d_y,
                              int ldy, int round_tf32, void* stream) {
  ADVOC_REQUIRE(d_x && d_stats && d_gamma && d_beta && d_y && pixels > 0 && C > 0, ADVOC_BAD_ARG, "bad bn_apply arguments");
  ADVOC_REQUIRE(act >= ADVOC_ACT_NONE && act <= ADVOC_ACT_RELU, ADVOC_BAD_ARG, "bn_apply fuses none / lrelu / relu only");
  int blocks = grid_for(pixels * C, 256 * 4);
  bn_apply_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_x, ldx, pixels, C, d_stats, d_gamma,
                                                                             d_beta, eps, act, alpha, d_y, ldy,
                                                                             round_tf32);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bn_inference(const float* d_x, int ldx, long pixels, int C, const float* d_mean,
                                  const float* d_var, const float* d_gamma, const float* d_beta, float eps, int act,
                                  float alpha, float* d_y, int ldy, int round_tf32, void* stream) {
  ADVOC_REQUIRE(d_x && d_mean && d_var && d_gamma && d_beta && d_y && pixels > 0 && C > 0, ADVOC_BAD_ARG,
                "bad bn_inference arguments");
  ADVOC_REQUIRE(act >= ADVOC_ACT_NONE && act <= ADVOC_ACT_RELU, ADVOC_BAD_ARG, "bn_inference fuses none / lrelu / relu only");
  bn_infer_kernel<<<grid_for(pixels * C, 256 * 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_x, ldx, pixels, C, d_mean, d_var, d_gamma, d_beta, eps, act, alpha, d_y, ldy, round_tf32);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bn_moving_update(const float* d_stats, long pixels, int C, float momentum, float* d_moving_mean,
                                      float* d_moving_var, void* stream) {
  ADVOC_REQUIRE(d_stats && d_moving_mean && d_moving_var && pixels > 0 && C > 0, ADVOC_BAD_ARG,
                "bad bn_moving_update arguments");
  ADVOC_REQUIRE(momentum >= 0.f && momentum <= 1.f, ADVOC_BAD_ARG, "momentum must lie in [0, 1]");
  bn_moving_update_kernel<<<(C + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_stats, pixels, C, momentum, d_moving_mean, d_moving_var);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bn_backward(const float* d_dy, int lddy, const float* d_y, int ldy, const float* d_x, int ldx,
                                 long pixels, int C, const float* d_stats, const float* d_gamma, float eps, int act,
                                 float alpha, float* d_red, float* d_dx, int lddx, int round_tf32, void* stream) {
  ADVOC_REQUIRE(d_dy && d_y && d_x && d_stats && d_gamma && d_red && d_dx && pixels > 0 && C > 0 && C <= 4096,
                ADVOC_BAD_ARG, "bad bn_backward arguments");
  ADVOC_REQUIRE(act >= ADVOC_ACT_NONE && act <= ADVOC_ACT_RELU, ADVOC_BAD_ARG, "bn_backward handles none / lrelu / relu");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ppb = 256 / (C < 256 ? C : 256);
  int blocks = grid_for(pixels, ppb * 8);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  bn_bwd_reduce_kernel<<<blocks, 256, 2 * C * sizeof(float), s>>>(d_dy, lddy, d_y, ldy, d_x, ldx, pixels, C, d_stats,
                                                                  eps, act, alpha, d_red);
  bn_bwd_apply_kernel<<<grid_for(pixels * C, 256 * 4), 256, 0, s>>>(d_dy, lddy, d_y, ldy, d_x, ldx, pixels, C, d_stats,
                                                                    d_gamma, eps, act, alpha, d_red, d_dx, lddx,
                                                                    round_tf32);
  count_launch(2);
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_tanh_backward(const float* d_dy, const float* d_y, float* d_dx, long n, void* stream) {
  ADVOC_REQUIRE(d_dy && d_y && d_dx && n > 0, ADVOC_BAD_ARG, "bad tanh_backward arguments");
  tanh_bwd_kernel<<<grid_for(n, 256 * 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_dy, d_y, d_dx, n);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_gan_logit_loss(const float* d_real, const float* d_fake, int n, int mode, float* d_loss,
                                    float* d_dreal, float* d_dfake, void* stream) {
  ADVOC_REQUIRE(d_fake && d_loss && d_dfake && n > 0 && mode >= 0 && mode <= 3, ADVOC_BAD_ARG, "bad loss arguments");
  ADVOC_REQUIRE((mode & 1) || (d_real && d_dreal), ADVOC_BAD_ARG, "the critic losses need the real logits");
  gan_logit_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_real, d_fake, n, mode, d_loss,
                                                                              d_dreal, d_dfake);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_gp_seed(const float* d_g, int batch, long n, float lambda, float* d_loss, float* d_u,
                             int round_tf32, void* stream) {
  ADVOC_REQUIRE(d_g && d_loss && d_u && batch > 0 && n > 0, ADVOC_BAD_ARG, "bad gp_seed arguments");
  gp_seed_kernel<<<batch, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_g, batch, n, lambda, d_loss, d_u, round_tf32);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

extern "C" int advoc_bn_gp(const float* d_v, const float* d_dy, const float* d_y, const float* d_x, long pixels, int C,
                           const float* d_stats, const float* d_gamma, float eps, float alpha, float* d_sums,
                           float* d_vz, float* d_xbar, int round_tf32, void* stream) {
  ADVOC_REQUIRE(d_v && d_dy && d_y && d_x && d_stats && d_gamma && d_sums && d_vz && d_xbar && pixels > 0 && C > 0 &&
                    C <= 2048,
                ADVOC_BAD_ARG, "bad bn_gp arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ppb = 256 / (C < 256 ? C : 256);
  int blocks = grid_for(pixels, ppb * 8);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  bn_gp_reduce_kernel<<<blocks, 256, 5 * C * sizeof(float), s>>>(d_v, d_dy, d_y, d_x, pixels, C, d_stats, eps, alpha,
                                                                 d_sums);
  bn_gp_apply_kernel<<<grid_for(pixels * C, 256 * 4), 256, 0, s>>>(d_v, d_dy, d_y, d_x, pixels, C, d_stats, d_gamma, eps,
                                                                   alpha, d_sums, d_vz, d_xbar, round_tf32);
  count_launch(2);
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}
