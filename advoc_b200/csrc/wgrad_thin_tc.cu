// Filter gradients of the THIN ends on the tensor cores (TF32 x split-fp32 operands, fp32 accumulate in TMEM):
//   dW[k][c] += sum over pixels p of  P[p][k] * S[p][c]
// where S is the wide side exactly as it sits in its NHWC buffer and P[p][.] are the <= 32 thin-side values that
// meet pixel p under the filter taps:
//   * MODE 0 / 1 -- conv from one / two input channels (generator encoder_1 and, as the conv over the big side,
//     decoder_1: models/advoc/advoc_model.py:91-94,153-158; discriminator layer_1: :184-187): p runs over the conv's
//     OUTPUT pixels, S = d loss / d output [p][Cout], P[p][tap * Cin + ci] = x[p * s - pad + tap][ci];
//   * MODE 2 -- stride-1 conv TO one channel (PatchGAN head, :196-199): p runs over the conv's INPUT pixels,
//     S = the input activation [p][Cin], P[p][tap] = dy[p + pad - tap].
// replaces: Conv2DBackpropFilter built by opt.minimize, advoc_model.py:254-257, for these layers.
//
// The CUDA-core kernels for these gradients (train.cu: wgrad_thin_tiled_kernel / wgrad_thin_kernel) ran at
// 256-473 us per layer on the regular model, 6-14x the time HBM needs to deliver S once
// (profiles/r02A_launches_train_regular.csv).  Here
//   * S tiles come in by TMA as the MN-major A operand (128 channels on the TMEM lanes, 128 pixels per stage:
//     [32 channels x 128 pixels] boxes in the 128B / 32B-atom swizzle, as in wgrad_tc.cu);
//   * the four producer warps gather P (one pixel per thread, the loads of the next stage in flight) and write it
//     TRANSPOSED into a K-major, 128B-swizzled B tile per warp: row k holds the 32 pixels of the warp, "hi" rows
//     (tf32(v)) followed by "lo" rows (tf32(v - hi)), so the thin side enters the product with ~21 mantissa bits;
//   * one warp issues 16 M=128, N = 2K, K=8 MMAs per stage; the accumulator [128 x 2K] lives in TMEM for the CTA's
//     whole pixel range;
//   * every CTA stores its partial [K][C] tile to the library workspace and wgrad_thin_reduce_kernel adds the
//     partials in a fixed order: the result is deterministic (the kernels it replaces used fp32 atomics).
// Channels beyond 128 run as blockIdx.y chunks.  With C = 32 / 64 only one / two boxes are loaded and the other
// A rows multiply whatever the ring holds there: their accumulator lanes are never read.
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace advoc {

float* wgrad_workspace(size_t bytes, cudaStream_t st);   // wgrad_tc.cu

namespace {

using namespace tc;

constexpr int T_BK = 128;                       // pixels per stage = producer threads
constexpr int T_THREADS = 160;                  // warps 0-3: gather producers, then epilogue; warp 4: MMA issuer
constexpr uint32_t T_BOX = T_BK * 128;          // one [128 pixels][32 channels] box
constexpr int T_STAGES = 2;

enum { W_CONV1 = 0, W_CONV2 = 1, W_TOONE = 2, W_CONV1_K5 = 3 };   // 3: 5x5 from one channel (MelspecGAN conv_0)

struct alignas(64) ThinWgParams {
  CUtensorMap tmS;      // wide side as a [pixels, channels] matrix
  const float* thin;    // gathered side
  int N, Ht, Wt, ldt;   // its extent and pixel stride
  int Hp, Wp;           // pixel grid p runs over
  int sh, sw, pt, pl;
  int C;                // all wide channels (row length of dW)
  long P, chunk;        // pixels, pixels per CTA (multiple of T_BK)
  float* ws;            // [gridDim.x][K][C] partials
  unsigned int* dbg;
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo_bytes) {   // see wgrad_tc.cu: make_mn_desc
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void mbar_arrive_cta_(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// CB = 32-channel boxes loaded per stage (1, 2 or 4)
template <int MODE, int CB>
__global__ void __launch_bounds__(T_THREADS, 1) wgrad_thin_tc_kernel(const __grid_constant__ ThinWgParams p) {
  constexpr int CIN = MODE == W_CONV2 ? 2 : 1;
  constexpr int KS = MODE == W_CONV1_K5 ? 5 : 4;
  constexpr int KREAL = KS * KS * CIN;                      // rows of dW
  constexpr int KV = KREAL <= 16 ? 16 : 32;                 // 25 -> 32: zero rows
  constexpr int NCOL = 2 * KV;                              // hi rows, lo rows
  constexpr uint32_t A_BYTES = 4 * T_BOX;                   // the MMA always spans 128 lanes (see the header)
  constexpr uint32_t BT_BYTES = NCOL * 128;                 // one warp's K-major tile: NCOL rows x 32 pixels
  constexpr uint32_t STAGE_BYTES = CB * T_BOX + 4 * BT_BYTES;
  // kind::tf32, D fp32, A MN-major (bit 15), B K-major, M = 128, N = NCOL
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(NCOL >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[T_STAGES], empty_bar[T_STAGES], acc_bar;
  __shared__ uint32_t tmem_base_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int m0 = (int)blockIdx.y * 128;
  const long p0 = (long)blockIdx.x * p.chunk;
  const long p1 = p0 + p.chunk < p.P ? p0 + p.chunk : p.P;
  const int iters = p0 < p.P ? (int)((p1 - p0 + T_BK - 1) / T_BK) : 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tmS);
    for (int s = 0; s < T_STAGES; ++s) { mbar_init(&full_bar[s], 5); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"((uint32_t)NCOL)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  const bool aborted = p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0;

  if (aborted) {
  } else if (warp < 4) {
    // ===== gather producers =====
    const int m = (int)threadIdx.x;
    const long pix_in_img = (long)p.Hp * p.Wp;
    auto gather = [&](int it, float (&v)[KV]) {
#pragma unroll
      for (int j = 0; j < KV; ++j) v[j] = 0.f;
      const long pix = p0 + (long)it * T_BK + m;
      if (it >= iters || pix >= p1) return;
      const int img = (int)(pix / pix_in_img);
      const int rem = (int)(pix - (long)img * pix_in_img);
      const int ph = rem / p.Wp, pw = rem - ph * p.Wp;
      // conv: tap (kh, kw) meets thin pixel (ph sh - pt + kh, pw sw - pl + kw); to-one: dy pixel (ph + pt - kh, pw + pl - kw)
      const int ih0 = MODE == W_TOONE ? ph + p.pt : ph * p.sh - p.pt;
      const int iw0 = MODE == W_TOONE ? pw + p.pl : pw * p.sw - p.pl;
      constexpr int DIR = MODE == W_TOONE ? -1 : 1;
      const float* x00 = p.thin + ((size_t)img * p.Ht * p.Wt + (long)ih0 * p.Wt + iw0) * p.ldt;
      bool cok[KS];
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) cok[kw] = (unsigned)(iw0 + DIR * kw) < (unsigned)p.Wt;
      const long rstride = (long)p.Wt * p.ldt;
#pragma unroll
      for (int kh = 0; kh < KS; ++kh) {
        const bool rok = (unsigned)(ih0 + DIR * kh) < (unsigned)p.Ht;
        const float* xr = x00 + DIR * kh * rstride;
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          if (rok && cok[kw]) {
            if (CIN == 1) {
              v[kh * KS + kw] = __ldg(xr + DIR * kw * p.ldt);
            } else {
              const float2 xv = __ldg(reinterpret_cast<const float2*>(xr + kw * p.ldt));
              v[(kh * KS + kw) * CIN] = xv.x;
              v[(kh * KS + kw) * CIN + CIN - 1] = xv.y;
            }
          }
        }
      }
    };
    float v[KV], vn[KV];
    gather(0, v);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      gather(it + 1, vn);
      if (threadIdx.x == 0) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.dbg, 61u);
        mbar_expect_tx(&full_bar[stage], CB * T_BOX);
        const int row = (int)(p0 + (long)it * T_BK);       // rows beyond P read as zero
#pragma unroll
        for (int b = 0; b < CB; ++b)
          tma_load_2d(&p.tmS, &full_bar[stage], smem_raw + (ring - smem_u32(smem_raw)) + stage * STAGE_BYTES + b * T_BOX,
                      m0 + 32 * b, row);
      } else {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.dbg, 61u);
      }
      // column `lane` of this warp's tile, rows k (hi) and KV + k (lo): 16-byte chunk (lane / 4) ^ (row % 8)
      const uint32_t tile = ring + (uint32_t)stage * STAGE_BYTES + CB * T_BOX + (uint32_t)warp * BT_BYTES;
      const uint32_t col = ((uint32_t)lane & 3u) << 2;
      const uint32_t chunk = (uint32_t)lane >> 2;
#pragma unroll
      for (int k = 0; k < KV; ++k) {
        const float hi = __uint_as_float(__float_as_uint(v[k]) & 0xFFFFE000u);   // truncated; lo = x - hi exactly, and
        const float lo = v[k] - hi;                                               // the tensor core truncates lo itself
        const uint32_t off = ((chunk ^ ((uint32_t)k & 7u)) << 4) + col;
        sts_f32(tile + (uint32_t)k * 128u + off, hi);
        sts_f32(tile + (uint32_t)(KV + k) * 128u + off, lo);   // (KV + k) % 8 == k % 8
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cta_(&full_bar[stage]);
      if (++stage == T_STAGES) { stage = 0; phase ^= 1u; }
#pragma unroll
      for (int j = 0; j < KV; ++j) v[j] = vn[j];
    }
    // ===== epilogue: channel c = m0 + 32 warp + lane on TMEM lane 32 warp + lane =====
    const int c = m0 + warp * 32 + lane;
    float* dst = p.ws + (size_t)blockIdx.x * KREAL * p.C;
    if (iters > 0) {
      mbar_wait(&acc_bar, 0u, p.dbg, 63u);
      tc_fence_after();
    }
    {
      uint32_t a[32], b[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = b[j] = 0u;
      if (iters > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), a);                       // KV = 16: hi 0-15 | lo 0-15
        if (KV == 32) tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u, b);   // KV = 32: a = hi, b = lo
        tmem_ld_wait();
      }
      if (c < p.C && warp < CB) {
#pragma unroll
        for (int j = 0; j < KREAL; ++j) {
          const float lo = KV == 16 ? __uint_as_float(a[(16 + j) & 31]) : __uint_as_float(b[j & 31]);
          dst[(size_t)j * p.C + c] = __uint_as_float(a[j]) + lo;
        }
      }
    }
  } else {
    // ===== MMA issuer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase, p.dbg, 62u);
        tc_fence_after();
        const uint32_t a_addr = ring + (uint32_t)stage * STAGE_BYTES;
        const uint64_t da = mn_desc(a_addr, T_BOX);
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
          const uint64_t db = make_smem_desc(a_addr + CB * T_BOX + (uint32_t)kt * BT_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)      // 8 pixels: 1024 bytes along K in every A box, 32 bytes in the B rows
            umma_tf32(tmem_base, da + (uint64_t)(64 * (kt * 4 + ks)), db + (uint64_t)(2 * ks), idesc,
                      (it | kt | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == T_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (iters > 0) umma_commit(&acc_bar);
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)NCOL)
                 : "memory");
  }
  (void)A_BYTES;
}

// dw[i] += sum over splits of ws[s][i], always in the same order: 8 interleaved partial sums per element,
// combined 0..7
__global__ void __launch_bounds__(256) wgrad_thin_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                                int nelem, int nsplits) {
  __shared__ float part[8][32];
  const int e = blockIdx.x * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;
  float acc = 0.f;
  if (e < nelem)
    for (int s = sl; s < nsplits; s += 8) acc += __ldg(ws + (size_t)s * nelem + e);
  part[sl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sl == 0 && e < nelem) {
    float t = part[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 8; ++j) t += part[j][threadIdx.x];
    dw[e] += t;
  }
}

template <int MODE, int CB>
int launch_thin_wg(const ThinWgParams& p, int splits, int chunks, float* dw, cudaStream_t st) {
  constexpr int KREAL = MODE == W_CONV1_K5 ? 25 : (MODE == W_CONV2 ? 32 : 16);
  constexpr int KV = KREAL <= 16 ? 16 : 32;
  // the A descriptor always spans four boxes: keep what lies behind a narrower stage inside the allocation
  constexpr int smem = T_STAGES * (CB * (int)T_BOX + 4 * 2 * KV * 128) + (4 - CB) * (int)T_BOX + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_thin_tc_kernel<MODE, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  wgrad_thin_tc_kernel<MODE, CB><<<dim3((unsigned)splits, (unsigned)chunks), T_THREADS, smem, st>>>(p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  const int nelem = KREAL * p.C;
  wgrad_thin_reduce_kernel<<<(nelem + 31) / 32, 256, 0, st>>>(p.ws, dw, nelem, splits);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

bool wide_ok(int C, const float* s, int ld) {
  return (C == 32 || C == 64 || (C >= 128 && C <= 1024 && C % 128 == 0)) && ld % 4 == 0 && aligned16(s);
}

}  // namespace

bool wgrad_thin_tc_eligible(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small) {
  static const bool disabled = getenv("ADVOC_NO_THIN_WGRAD_TC") != nullptr;   // A/B switch
  if (disabled || d->math == ADVOC_MATH_FP32 || !tc::tma_ok() || device_arch() != 100) return false;
  if ((long)d->N * d->H * d->W >= 2147483647L) return false;
  if (d->kh == 5 && d->kw == 5) return d->Cin == 1 && wide_ok(d->Cout, small, ld_small);
  if (d->kh != 4 || d->kw != 4) return false;
  if (d->Cin == 1 || d->Cin == 2)
    return wide_ok(d->Cout, small, ld_small) &&
           (d->Cin == 1 || (ld_big % 2 == 0 && (reinterpret_cast<uintptr_t>(big) & 7u) == 0));
  return d->Cout == 1 && d->sh == 1 && d->sw == 1 && wide_ok(d->Cin, big, ld_big);
}

int wgrad_thin_tc(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small, float* dw,
                  void* stream) {
  ThinWgParams p = {};
  const bool to_one = d->Cout == 1 && d->Cin > 2;
  const int mode = to_one ? W_TOONE : (d->kh == 5 ? W_CONV1_K5 : (d->Cin == 2 ? W_CONV2 : W_CONV1));
  const float* wide = to_one ? big : small;
  const int ldw = to_one ? ld_big : ld_small;
  p.C = to_one ? d->Cin : d->Cout;
  p.thin = to_one ? small : big;
  p.ldt = to_one ? ld_small : ld_big;
  p.N = d->N;
  p.Ht = to_one ? d->Ho : d->H; p.Wt = to_one ? d->Wo : d->W;
  p.Hp = to_one ? d->H : d->Ho; p.Wp = to_one ? d->W : d->Wo;
  p.sh = d->sh; p.sw = d->sw; p.pt = d->pad_t; p.pl = d->pad_l;
  p.P = (long)p.N * p.Hp * p.Wp;
  if (p.P == 0) return ADVOC_OK;
  int st = encode_tiled2d(&p.tmS, wide, p.C, p.P, (size_t)ldw * 4, 32, T_BK, true);
  if (st) return st;
  const int cb = p.C >= 128 ? 4 : p.C / 32;
  const int chunks = p.C >= 128 ? p.C / 128 : 1;
  long splits = ((long)sm_count() + chunks - 1) / chunks;      // one CTA per SM
  const long max_splits = (p.P + T_BK - 1) / T_BK;
  if (splits > max_splits) splits = max_splits;
  p.chunk = ((p.P + splits - 1) / splits + T_BK - 1) / T_BK * T_BK;
  splits = (p.P + p.chunk - 1) / p.chunk;
  cudaStream_t cst = reinterpret_cast<cudaStream_t>(stream);
  const size_t kv = mode == W_CONV2 || mode == W_CONV1_K5 ? 32 : 16;
  p.ws = wgrad_workspace((size_t)splits * kv * p.C * sizeof(float), cst);
  if (p.ws == nullptr) return ADVOC_UNSUPPORTED;   // workspace cannot grow inside a graph capture: the caller falls back
  p.dbg = tc::debug_word();
#define ADVOC_THIN_WG(MODE)                                                     \
  switch (cb) {                                                                 \
    case 1: return launch_thin_wg<MODE, 1>(p, (int)splits, chunks, dw, cst);    \
    case 2: return launch_thin_wg<MODE, 2>(p, (int)splits, chunks, dw, cst);    \
    default: return launch_thin_wg<MODE, 4>(p, (int)splits, chunks, dw, cst);   \
  }
  if (mode == W_CONV1) { ADVOC_THIN_WG(W_CONV1) }
  if (mode == W_CONV2) { ADVOC_THIN_WG(W_CONV2) }
  if (mode == W_CONV1_K5) { ADVOC_THIN_WG(W_CONV1_K5) }
  ADVOC_THIN_WG(W_TOONE)
#undef ADVOC_THIN_WG
}

}  // namespace advoc
