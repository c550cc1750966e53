// tcgen05 filter gradient (TF32 operands, fp32 accumulate in TMEM):
//   dW[tap][cb][cs] += sum over output pixels p of  big[p*s - pad + tap, cb] * small[p, cs]
// replaces: Conv2DBackpropFilter built by opt.minimize, models/advoc/advoc_model.py:254-257.
//
// GEMM view per filter tap: D[m, n] = sum_p X[p, m] * Y[p, n] with the CONTRACTION over pixels, so
// both operands are "MN-major" (channels contiguous) exactly as they sit in the NHWC buffers:
//   * big side : TMA im2col-mode boxes {32 channels x BK pixels} at the tap's offset (padding and
//                stride live in the tensor map, same as the forward conv),
//   * small side: plain 2-D TMA boxes {32 channels x BK pixels} of the [pixels, channels] matrix.
// Each box lands as BK rows of 128 B in the 128B/32B-atom swizzle = one column block of the
// canonical MN-major UMMA layout for 32-bit operands (4-pixel atoms 512 B apart, 32-channel
// blocks BK*128 B apart).
// The small side sits on the 128 TMEM lanes (M); the N axis is the concatenation (tap, big channel):
// WNB boxes = 32 * WNB accumulator columns per CTA, so the small-side tile is fetched once for WNB
// tap/channel blocks.  Work = (M tile, N tile, pixel split).  The pixel splits are reduced
// DETERMINISTICALLY (round 2; round 1 used fp32 atomics): every split stores its partial tile to a
// library-owned workspace [split][tap][cb][cs] and wgrad_reduce_kernel adds the splits to dW in split
// order; a layer with a single split adds its tile to dW directly (one owner per element).
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace advoc {

int check_conv_desc(const advoc_conv_desc* d);
float* wgrad_workspace(size_t bytes, cudaStream_t st);   // shared with wgrad_thin_tc.cu

namespace {

using namespace tc;

constexpr int WM = 128;       // UMMA_M: small-side channels on the TMEM lanes
// Pixels per pipeline stage (template WBK).  Every (channel block, WBK pixels) box is one TMA operation and
// the kernel issues 12 of them per stage; with 2 KB boxes (WBK = 16, round 1) the per-operation cost of
// the TMA unit, not bytes, paced the ring (ncu: 3.5 TB/s L2->SM at 22 % tensor activity, long-scoreboard
// waits), so round 2 moves 32 or 64 pixels per box.
// 32-channel blocks of the (tap, big-channel) axis per CTA.  Round 1 used 16 (512 TMEM columns, one
// CTA per SM); ncu showed the kernel latency-bound that way (tensor pipe 20-23 % active, L2 8-13 %,
// issue slots 5-8 %: profiles/r02k_ncu_full_wgrad_tc_regular.csv), so round 2 halves the tile to 8
// blocks = 256 columns and 96 KB of ring and runs TWO CTAs per SM whose TMA latencies overlap.
constexpr int WNB = 8;
constexpr int W_NPROD = 2;        // TMA producer warps: warp 0 and warps 6 .. 4 + W_NPROD (r02: 1 -> 1420, 2 -> 1460, 4 -> 1465 samples/s)
constexpr int W_THREADS = 192 + 32 * (W_NPROD - 1);   // + warp 1: MMA issuer, warps 2-5: epilogue

struct alignas(64) WgParams {
  CUtensorMap tmBig;    // im2col
  CUtensorMap tmSmall;  // 2-D
  float* dw;
  int Cb, Cs;
  int cbb;              // 32-channel blocks of the big side
  int nboxes;           // taps * cbb: (tap, channel-block) boxes; the last N tile may be partial
  int mtiles, ntiles, kw;
  int Ho, Wo, sh, sw, lower_h, lower_w;
  long P, chunk;        // output pixels, pixels per split (multiple of WBK)
  float* ws;            // partial tiles [nsplits][taps * Cb * Cs], or nullptr: accumulate into dw (see `atomic`)
  long ws_stride;
  int atomic;           // 1: fp32 atomics into dw (fallback when the workspace cannot be grown)
  unsigned int* dbg;
};

// MN-major 32-bit operand.  tcgen05 accepts exactly one smem layout for it: SWIZZLE_128B_BASE32B
// (128-byte rows, 32-byte chunks XOR-swizzled over 4-row atoms) -- what TMA produces with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  32-channel blocks are `lbo` bytes apart, 4-pixel atoms
// 512 bytes apart.
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

// kind::tf32, D fp32, A and B MN-major, M = 128, N = 256
constexpr uint32_t kIdescMN = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                              ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);

// D[cs, (tap, cb)] += sum_p small[p, cs] * big[p @ tap, cb]: the small-side tile is fetched once
// per stage and shared by all 16 (tap, channel-block) boxes of the big side, which sit back to
// back in smem so that ONE N=256 MMA spans eight of them (LBO = box size).
template <int WBK, int WSTAGES>
__global__ void __launch_bounds__(W_THREADS, 2) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  constexpr int BLK_BYTES = WBK * 128;                  // one [WBK pixels][32 channels] box
  constexpr int A_BYTES = (WM / 32) * BLK_BYTES;        // small side, 128 channels
  constexpr int B_BYTES = WNB * BLK_BYTES;              // big side, WNB (tap, channel-block) boxes
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WSTAGES];
  __shared__ __align__(8) uint64_t empty_bar[WSTAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  // work decode: N tile fastest, then M tile, then pixel split (CTAs of one split share L2 lines)
  unsigned id = blockIdx.x;
  const int nt = id % p.ntiles; id /= p.ntiles;
  const int mt = id % p.mtiles; id /= p.mtiles;
  const unsigned split = id;
  const long p0 = (long)id * p.chunk;
  if (p0 >= p.P) return;
  const long p1 = p0 + p.chunk < p.P ? p0 + p.chunk : p.P;
  const int iters = (int)((p1 - p0 + WBK - 1) / WBK);   // (the last stage may run past p1: TMA zero-fills rows beyond P,
                                                        //  and chunk is a multiple of WBK so splits never overlap)
  if (p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = mt * WM;
  const int gb0 = nt * WNB;   // first (tap, channel-block) index of this CTA

  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  constexpr uint32_t TMEM_COLS = 32 * WNB;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmBig);
    prefetch_tmap(&p.tmSmall);
#pragma unroll
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(&full_bar[s], W_NPROD);    // one arrival (+ its transaction bytes) per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  if (warp == 0 || warp >= 6) {
    // ===== TMA producers: the 12 boxes of a stage (4 small-side, 8 big-side) are split over W_NPROD warps -- issuing one TMA load costs its thread ~150 cycles and the lanes of one warp only partly
    // overlap (profiles/r02r_tma_issue_probe.txt), so two warps halve the time a stage spends being issued =====
    constexpr int PER_WARP = (WM / 32 + WNB) / W_NPROD;
    static_assert(PER_WARP * W_NPROD == WM / 32 + WNB, "boxes must split evenly over the producer warps");
    const int pw = warp == 0 ? 0 : warp - 5;
    const int box = lane + PER_WARP * pw;          // box index of this lane (valid for lane < PER_WARP)
    const bool mine = lane < PER_WARP;
    const bool is_small = mine && box < WM / 32;
    int my_c = 0;
    uint16_t my_kw = 0, my_kh = 0;
    bool have = mine;
    if (is_small) {
      my_c = m0 + 32 * box;
    } else if (mine) {
      const int gb = gb0 + (box - WM / 32);
      have = gb < p.nboxes;                        // boxes of a partial last tile are not loaded
      const int tap = gb / p.cbb, cblk = gb - tap * p.cbb;
      my_c = 32 * cblk;
      my_kh = (uint16_t)(tap / p.kw);
      my_kw = (uint16_t)(tap - (tap / p.kw) * p.kw);
    }
    const unsigned loaded = __ballot_sync(0xffffffffu, have);
    const uint32_t my_bytes = (uint32_t)__popc(loaded) * BLK_BYTES;
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      const long pix = p0 + (long)it * WBK;
      const int ow = (int)(pix % p.Wo);
      const long r = pix / p.Wo;
      const int oh = (int)(r % p.Ho);
      const int img = (int)(r / p.Ho);
      if (lane == 0) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.dbg, 11u);
        mbar_expect_tx(&full_bar[stage], my_bytes);
      }
      __syncwarp();
      uint8_t* dst = ring_ptr + stage * STAGE_BYTES + box * BLK_BYTES;
      if (is_small) {
        tma_load_2d(&p.tmSmall, &full_bar[stage], dst, my_c, (int)pix);
      } else if (have) {
        tma_load_im2col_4d(&p.tmBig, &full_bar[stage], dst, my_c, ow * p.sw + p.lower_w, oh * p.sh + p.lower_h,
                           img, my_kw, my_kh);
      }
      if (++stage == WSTAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // one thread waits and issues, inside a single elect.sync region (see tc::elect_one)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase, p.dbg, 12u);
        const uint32_t a_addr = ring + stage * STAGE_BYTES;
        const uint64_t da = make_mn_desc(a_addr, BLK_BYTES);
        const uint64_t db0 = make_mn_desc(a_addr + A_BYTES, BLK_BYTES);
#pragma unroll
        for (int k = 0; k < WBK / 8; ++k) {
          // 8 pixels = 1024 bytes along K inside every box; one N = 256 MMA spans the 8 big-side boxes
          const uint32_t acc = (it | k) != 0 ? 1u : 0u;
          umma_tf32(tmem_base, da + (uint64_t)(64 * k), db0 + (uint64_t)(64 * k), kIdescMN, acc);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == WSTAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&tmem_full_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cs = m0 + q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0u, p.dbg, 13u);
    tc_fence_after();
#pragma unroll 1
    for (int b = 0; b < WNB; ++b) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 32), v);
      tmem_ld_wait();
      const int gb = gb0 + b;
      const int tap = gb / p.cbb, cblk = gb - tap * p.cbb;
      if (cs < p.Cs && gb < p.nboxes) {
        float* dst = p.ws ? p.ws + (size_t)split * p.ws_stride : p.dw;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int cb = cblk * 32 + j;
          if (cb >= p.Cb) continue;
          float* q = dst + ((size_t)tap * p.Cb + cb) * p.Cs + cs;   // 32 lanes = 32 consecutive cs: one 128-byte line
          const float x = __uint_as_float(v[j]);
          if (p.ws) *q = x;
          else if (p.atomic) atomicAdd(q, x);
          else *q += x;                                             // single split: this thread owns the element
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// dw[i] += sum over splits (in split order) of ws[s][i]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                           long nelem4, int nsplits, long stride) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nelem4; i += (long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < nsplits; ++sp) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ws + (size_t)sp * stride) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4* d = reinterpret_cast<float4*>(dw) + i;
    float4 o = *d;
    o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
    *d = o;
  }
}

}  // namespace

// Grow-only workspace for the split partials (one stream at a time uses the library's wgrad / split-K convs; the
// partials of a launch are consumed by the reduce kernel queued right behind it).  Cannot grow while
// the stream is being captured into a CUDA graph: the caller's eager warm-up pass sizes it.
float* wgrad_workspace(size_t bytes, cudaStream_t st) {
  static float* buf = nullptr;
  static size_t cap = 0;
  if (bytes <= cap) return buf;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  float* nb = nullptr;
  const size_t want = bytes < ((size_t)64 << 20) ? ((size_t)64 << 20) : bytes + (bytes >> 2);
  if (cudaMalloc(&nb, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  // The outgrown buffer is NOT freed: CUDA graphs captured earlier (an inference engine's split-K layers, a train
  // engine's filter gradients) keep launching kernels that point at it.  Growth happens a handful of times per
  // process (sizes are per layer geometry), so the retired buffers cost at most a few hundred MB.
  buf = nb; cap = want;
  return buf;
}

bool wgrad_tc_eligible(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small) {
  const int cbb = (d->Cin + 31) / 32;
  return d->math != ADVOC_MATH_FP32 && tc::tma_ok() && device_arch() == 100 && d->Cin >= 32 && d->Cout >= 32 &&
         d->Cin % 4 == 0 && d->Cout % 4 == 0 && ld_big % 4 == 0 && ld_small % 4 == 0 && aligned16(big) &&
         aligned16(small) && d->kh * d->kw <= 255 && d->pad_t <= 127 && d->pad_l <= 127 && cbb > 0;
}

template <int WBK, int WSTAGES>
int wgrad_tc_impl(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small, float* dw,
                  void* stream) {
  constexpr int STAGE_BYTES = (WM / 32 + WNB) * WBK * 128;
  WgParams p = {};
  const long P = (long)d->N * d->Ho * d->Wo;
  if (P == 0) return ADVOC_OK;
  ADVOC_REQUIRE(P < 2147483647L, ADVOC_BAD_SHAPE, "too many pixels");
  const int lower_h = -d->pad_t, lower_w = -d->pad_l;
  const int upper_h = lower_h + (d->Ho - 1) * d->sh - (d->H - 1);
  const int upper_w = lower_w + (d->Wo - 1) * d->sw - (d->W - 1);
  ADVOC_REQUIRE(upper_h >= -128 && upper_h <= 127 && upper_w >= -128 && upper_w <= 127, ADVOC_UNSUPPORTED,
                "im2col corner out of range");
  int st = encode_im2col(&p.tmBig, big, d->N, d->H, d->W, ld_big, d->Cin, lower_h, lower_w, upper_h, upper_w,
                         d->sh, d->sw, 32, WBK, true);
  if (st) return st;
  st = encode_tiled2d(&p.tmSmall, small, d->Cout, P, (size_t)ld_small * 4, 32, WBK, true);
  if (st) return st;
  p.dw = dw; p.Cb = d->Cin; p.Cs = d->Cout;
  p.cbb = (d->Cin + 31) / 32;
  p.mtiles = (d->Cout + WM - 1) / WM;
  p.nboxes = d->kh * d->kw * p.cbb;
  p.ntiles = (p.nboxes + WNB - 1) / WNB;
  p.kw = d->kw;
  p.Ho = d->Ho; p.Wo = d->Wo; p.sh = d->sh; p.sw = d->sw; p.lower_h = lower_h; p.lower_w = lower_w;
  p.P = P;
  const long tiles = (long)p.mtiles * p.ntiles;
  // pixel splits fill the chip's 2 x #SM CTA slots once; a layer that already has about that many
  // (M, N) tiles runs unsplit (one owner per gradient element, no partials to write and reduce)
  static const long waves = getenv("ADVOC_WGRAD_WAVES") ? atol(getenv("ADVOC_WGRAD_WAVES")) : 4;   // r02 sweep: 2 -> 1178, 3 -> 1242, 4 -> 1243, 6 -> 1233, 8 -> 1235 samples/s (regular step)
  long splits = ((long)sm_count() * waves + tiles - 1) / tiles;
  if (tiles * 10 >= (long)sm_count() * waves * 7) splits = 1;
  const long max_splits = (P + 8 * WBK - 1) / (8 * WBK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.chunk = ((P + splits - 1) / splits + WBK - 1) / WBK * WBK;
  splits = (P + p.chunk - 1) / p.chunk;
  const long ctas = tiles * splits;
  ADVOC_REQUIRE(ctas < 2147483647L, ADVOC_BAD_SHAPE, "too many wgrad CTAs");
  cudaStream_t cst = reinterpret_cast<cudaStream_t>(stream);
  const long nelem = (long)d->kh * d->kw * d->Cin * d->Cout;
  static const bool force_atomic = getenv("ADVOC_WGRAD_ATOMIC") != nullptr;   // A/B switch (round-1 behaviour)
  p.ws = nullptr; p.ws_stride = nelem; p.atomic = 0;
  if (force_atomic) {
    p.atomic = 1;
  } else if (splits > 1) {
    p.ws = (nelem % 4 == 0) ? wgrad_workspace((size_t)splits * nelem * sizeof(float), cst) : nullptr;
    if (!p.ws) p.atomic = 1;               // workspace unavailable (capture in progress / odd size): atomics
  }
  p.dbg = debug_word();
  constexpr int smem = WSTAGES * STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<WBK, WSTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem));
    configured = true;
  }
  wgrad_tc_kernel<WBK, WSTAGES><<<(unsigned)ctas, W_THREADS, smem, cst>>>(p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  if (p.ws) {
    const long n4 = nelem / 4;
    const long want = (n4 + 255) / 256;
    const int blocks = (int)(want < (long)sm_count() * 8 ? want : (long)sm_count() * 8);
    wgrad_reduce_kernel<<<blocks, 256, 0, cst>>>(p.ws, dw, n4, (int)splits, nelem);
    count_launch();
    ADVOC_CHECK_CUDA(cudaGetLastError());
  }
  return ADVOC_OK;
}

int wgrad_tc(const advoc_conv_desc* d, const float* big, int ld_big, const float* small, int ld_small, float* dw,
             void* stream) {
  // pixels per TMA box / ring depth (two CTAs per SM up to 96 KB of ring): ADVOC_WGRAD_BK = 16 | 32 | 64
  static const int bk = getenv("ADVOC_WGRAD_BK") ? atoi(getenv("ADVOC_WGRAD_BK")) : 32;   // r02: 16 -> 1247, 32 -> 1399, 64 -> 1370 samples/s
  if (bk == 16) return wgrad_tc_impl<16, 4>(d, big, ld_big, small, ld_small, dw, stream);
  if (bk == 64) return wgrad_tc_impl<64, 2>(d, big, ld_big, small, ld_small, dw, stream);
  return wgrad_tc_impl<32, 2>(d, big, ld_big, small, ld_small, dw, stream);
}

}  // namespace advoc
