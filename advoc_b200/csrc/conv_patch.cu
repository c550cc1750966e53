// tcgen05 transposed convolution with a shared-memory activation PATCH (sm_100a).
//
// conv_tc.cu fetches the im2col operand once per filter tap, so a k4 s2 transposed convolution
// pulls every input pixel through L2 -> smem 16 times and is bound by that traffic.  Here the CTA
// loads, per 32-channel block, ONE contiguous run of zero-padded input pixels ("patch", a single
// im2col-mode TMA box of up to 1024 pixels whose pixel-box corners include the halo) and presents
// every filter tap to the tensor core as a ROW-SHIFTED view of that patch: the UMMA descriptor's
// start address is advanced by (dr*Wp + dc) rows of 128 bytes.  The 128B swizzle is a function of
// the absolute shared-memory address, so the shifted view de-swizzles correctly with
// base_offset = 0 (probed on B200: csrc/selftest.cu, scripts/dev_desc_shift.py).
//
//   out[2a+ph, 2b+pw, n] = sum_{dr,dc,c} x[a+dr, b+dc, c] * W[kh=ph+1-2dr, kw=pw+1-2dc][n][c]
//
// GEMM rows are padded-linear positions q = (img*Hp + h')*Wp + w' (halo positions are computed and
// thrown away: 5-28 % of the rows), all s_h*s_w parity classes of a position are accumulated side
// by side in TMEM (class z in columns [z*BN, z*BN+BN)), the filter taps stream through a ring of
// [BN x 32] K-major tiles.  Warp roles / barriers as in conv_tc.cu, with separate rings for the
// patch (2 stages) and the filter tiles.
// replaces: tf.layers.conv2d_transpose (models/advoc/advoc_model.py:65-69) and the input
// gradients of the stride-2 convolutions.
#include "epilogue.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace advoc {

int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* out);

namespace {

using namespace tc;

constexpr int PM = 128;          // positions per CTA (UMMA_M)
constexpr int PK = 32;           // channels per k-block (one 128-byte swizzle row)
constexpr int P_MAXT = 36;       // filter taps
constexpr int P_MAXC = 4;        // parity classes
constexpr int P_ASTAGES = 2;
constexpr int P_BOX = 64;         // pixels per patch TMA box (several boxes in flight per stage)
constexpr int P_MAXG = 4;        // filter-group ring depth
constexpr int P_THREADS = 192;

struct alignas(64) PatchParams {
  CUtensorMap tmA;               // im2col mode, box {32 channels, R pixels}
  CUtensorMap tmB;               // filter [taps*Cn][Ck], box {32, BN}
  int ntaps, ncls;
  short tap_cls[P_MAXT];
  short tap_wrow[P_MAXT];
  int tap_shift[P_MAXT];         // rows from the patch start (>= 0)
  int cls_ph[P_MAXC], cls_pw[P_MAXC];
  int Hs, Ws, Hp, Wp, lo_h, lo_w, Nimg;
  int lead;                      // rows between patch start and the first position of the tile
  int R;                         // patch rows (multiple of P_BOX)
  int G;                         // filter taps per group (one barrier per group)
  long Q;                        // padded positions in the batch
  int Cn, kblocks, osh, osw;
  int bstages;                   // groups in the filter ring
  EpiDev epi;
  unsigned int* dbg;
};

__device__ __forceinline__ void epi_store4p(const EpiDev& e, size_t pix, int n, float4 acc);

template <int BN>
__global__ void __launch_bounds__(P_THREADS) conv_patch_kernel(const __grid_constant__ PatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[P_ASTAGES], a_empty[P_ASTAGES];
  __shared__ __align__(8) uint64_t b_full[P_MAXG], b_empty[P_MAXG];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  const int n_ntiles = p.Cn / BN;
  const int n_tile = (int)(blockIdx.x % n_ntiles);
  const long m_tile = blockIdx.x / n_ntiles;
  const long q0 = (long)p.lead + m_tile * PM;     // first position of this tile
  if (q0 >= p.Q) return;
  if (p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0) return;
  const int n0 = n_tile * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  const uint32_t a_bytes = ((uint32_t)p.R * 128u + 1023u) & ~1023u;   // per patch stage, 1 KB aligned
  const uint32_t b_off = P_ASTAGES * a_bytes;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = (P_MAXC * BN) <= 512 ? (P_MAXC * BN < 32 ? 32 : P_MAXC * BN) : 512;
  const int bstages = p.bstages;
  const uint32_t g_bytes = (uint32_t)p.G * B_BYTES;   // one filter group

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    for (int s = 0; s < P_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < bstages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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

  if (warp == 0) {
    // ===== TMA producer: the whole warp issues (one box / one filter tile per lane) =====
    const long qs = q0 - p.lead;                   // patch start (>= 0)
    const int nboxes = p.R / P_BOX;
    int bw = 0, bh = 0, bimg = 0;                  // start coordinates of this lane's patch box
    if (lane < nboxes) {
      const long qb = qs + (long)lane * P_BOX;
      bw = (int)(qb % p.Wp) - p.lo_w;
      const long r = qb / p.Wp;
      bh = (int)(r % p.Hp) - p.lo_h;
      bimg = (int)(r / p.Hp);
    }
    const int ngroups = (p.ntaps + p.G - 1) / p.G;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int kb = 0; kb < p.kblocks; ++kb) {
      if (lane == 0) {
        mbar_wait(&a_empty[as], aph ^ 1u, p.dbg, 21u);
        mbar_expect_tx(&a_full[as], (uint32_t)p.R * 128u);
      }
      __syncwarp();
      if (lane < nboxes)
        tma_load_im2col_4d(&p.tmA, &a_full[as], ring_ptr + as * a_bytes + lane * (P_BOX * 128), kb * PK, bw, bh, bimg,
                           0, 0);
      if (++as == P_ASTAGES) { as = 0; aph ^= 1u; }
      for (int g = 0; g < ngroups; ++g) {
        const int t0 = g * p.G;
        const int cnt = p.ntaps - t0 < p.G ? p.ntaps - t0 : p.G;
        if (lane == 0) {
          mbar_wait(&b_empty[bs], bph ^ 1u, p.dbg, 22u);
          mbar_expect_tx(&b_full[bs], (uint32_t)cnt * B_BYTES);
        }
        __syncwarp();
        if (lane < cnt)
          tma_load_2d(&p.tmB, &b_full[bs], ring_ptr + b_off + bs * g_bytes + lane * B_BYTES, kb * PK,
                      (int)p.tap_wrow[t0 + lane] * p.Cn + n0);
        if (++bs == bstages) { bs = 0; bph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(PM >> 4) << 24);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, started = 0;
      const int ngroups = (p.ntaps + p.G - 1) / p.G;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&a_full[as], aph, p.dbg, 23u);
        const uint32_t a_addr = ring + as * a_bytes;
        for (int g = 0; g < ngroups; ++g) {
          const int t0 = g * p.G;
          const int cnt = p.ntaps - t0 < p.G ? p.ntaps - t0 : p.G;
          mbar_wait(&b_full[bs], bph, p.dbg, 24u);
          tc_fence_after();
          for (int i = 0; i < cnt; ++i) {
            const int t = t0 + i;
            const int cls = p.tap_cls[t];
            // the tap = the patch viewed from `tap_shift` rows further down
            const uint64_t da = make_smem_desc(a_addr + (uint32_t)p.tap_shift[t] * 128u);
            const uint64_t db = make_smem_desc(ring + b_off + bs * g_bytes + i * B_BYTES);
            const uint32_t d_tmem = tmem_base + (uint32_t)(cls * BN);
#pragma unroll
            for (int k = 0; k < PK / 8; ++k)
              umma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                        (((started >> cls) & 1u) | (uint32_t)(k != 0)));
            started |= 1u << cls;
          }
          umma_commit(&b_empty[bs]);
          if (++bs == bstages) { bs = 0; bph ^= 1u; }
        }
        umma_commit(&a_empty[as]);
        if (++as == P_ASTAGES) { as = 0; aph ^= 1u; }
      }
      umma_commit(&tmem_full_bar);
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;
    const long pos = q0 + q * 32 + lane;
    bool interior = false;
    int hh = 0, ww = 0;
    long img = 0;
    if (pos < p.Q) {
      const int wq = (int)(pos % p.Wp);
      const long r = pos / p.Wp;
      const int hq = (int)(r % p.Hp);
      img = r / p.Hp;
      hh = hq - p.lo_h;
      ww = wq - p.lo_w;
      interior = hh >= 0 && hh < p.Hs && ww >= 0 && ww < p.Ws && img < p.Nimg;
    }
    mbar_wait(&tmem_full_bar, 0u, p.dbg, 25u);
    tc_fence_after();
#pragma unroll 1
    for (int z = 0; z < p.ncls; ++z) {
      const int oh = hh * p.osh + p.cls_ph[z], ow = ww * p.osw + p.cls_pw[z];
      const bool valid = interior && oh < p.epi.Hs && ow < p.epi.Ws;
      const size_t pix = valid ? ((size_t)img * p.epi.Hs + oh) * p.epi.Ws + ow : 0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(z * BN + c0), v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            epi_store4p(p.epi, pix, n0 + c0 + j,
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                    __uint_as_float(v[j + 3])));
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

// four consecutive channels of one stored pixel (same contract as conv_tc.cu's epilogue)
__device__ __forceinline__ void epi_store4p(const EpiDev& e, size_t pix, int n, float4 acc) {
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  if (e.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  float sc[4] = {1.f, 1.f, 1.f, 1.f};
  if (e.keep_prob < 1.f) {
    const size_t idx = pix * e.Cout + n;
    const float inv = 1.f / e.keep_prob;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool keep = e.mask ? (__ldg(e.mask + idx + j) != 0) : dropout_keep(e.seed, idx + j, e.keep_prob);
      sc[j] = keep ? inv : 0.f;
    }
  }
  float y[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) y[j] = apply_act(v[j], e.act0, e.alpha) * sc[j];
  if (e.gate) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(e.gate + pix * e.ldg + e.coffg + n));
    const float gv[4] = {g.x, g.y, g.z, g.w};
    const float neg = e.gate_act == ADVOC_ACT_LRELU ? e.alpha : 0.f;
    const float s = n < e.gate_split ? e.gscale0 : e.gscale1;
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] *= (gv[j] > 0.f ? 1.f : neg) * s;
  }
  float4* dst = reinterpret_cast<float4*>(e.out0 + pix * e.ld0 + e.coff0 + n);
  if (e.accumulate) {
    const float4 o = *dst;
    y[0] += o.x; y[1] += o.y; y[2] += o.z; y[3] += o.w;
  }
  if (e.round) {
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = round_tf32(y[j]);
  }
  *dst = make_float4(y[0], y[1], y[2], y[3]);
  if (e.out1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      y[j] = apply_act(v[j], e.act1, e.alpha) * sc[j];
      if (e.round) y[j] = round_tf32(y[j]);
    }
    *reinterpret_cast<float4*>(e.out1 + pix * e.ld1 + e.coff1 + n) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

struct PatchPlan {
  PatchParams p;
  int bn;
  size_t smem;
  long ctas;
};

// Geometry of y = conv_transpose(x): x [N,Ho,Wo,Cout(desc)] -> y [N,H,Ws,Cin(desc)].
// Returns false if the layer does not fit this kernel (caller falls back to conv_tc.cu).
bool plan_patch(const advoc_conv_desc* d, int Wstored, PatchPlan* pl) {
  PatchParams& p = pl->p;
  // position grid = the class grid of the OUTPUT: a in [0, ceil(H/sh)), b in [0, ceil(Wstored/sw));
  // input rows/columns outside [0,Ho) x [0,Wo) are zero-filled by the TMA box corners
  const int Hs = (d->H + d->sh - 1) / d->sh, Ws = (Wstored + d->sw - 1) / d->sw;
  if (d->sh * d->sw > P_MAXC || d->kh * d->kw > P_MAXT) return false;
  int lo_h = 0, hi_h = 0, lo_w = 0, hi_w = 0, nt = 0, nc = 0;
  struct T { int cls, wrow, dr, dc; } taps[P_MAXT];
  for (int ph = 0; ph < d->sh; ++ph)
    for (int pw = 0; pw < d->sw; ++pw) {
      // does this class produce any stored output?
      if (ph >= d->H || pw >= Wstored) continue;
      const int z = nc++;
      p.cls_ph[z] = ph; p.cls_pw[z] = pw;
      for (int kh = 0; kh < d->kh; ++kh) {
        if ((ph + d->pad_t - kh) % d->sh != 0) continue;
        const int dr = (ph + d->pad_t - kh) / d->sh;   // exact (may be negative): C++ division of a multiple
        for (int kw = 0; kw < d->kw; ++kw) {
          if ((pw + d->pad_l - kw) % d->sw != 0) continue;
          const int dc = (pw + d->pad_l - kw) / d->sw;
          if (nt >= P_MAXT) return false;
          taps[nt++] = {z, kh * d->kw + kw, dr, dc};
          if (-dr > lo_h) lo_h = -dr;
          if (dr > hi_h) hi_h = dr;
          if (-dc > lo_w) lo_w = -dc;
          if (dc > hi_w) hi_w = dc;
        }
      }
    }
  if (nc == 0 || nt == 0) return false;
  const int Cn = d->Cin, Ck = d->Cout;
  int bn = Cn % 128 == 0 ? 128 : (Cn % 64 == 0 ? 64 : 32);
  while (nc * bn > 512) bn >>= 1;
  if (bn < 32 || Cn % bn != 0 || Ck % PK != 0) return false;
  p.ntaps = nt; p.ncls = nc;
  p.Hs = Hs; p.Ws = Ws; p.Hp = Hs + lo_h + hi_h; p.Wp = Ws + lo_w + hi_w; p.lo_h = lo_h; p.lo_w = lo_w;
  p.Nimg = d->N;
  p.lead = lo_h * p.Wp + lo_w;
  const int tail = hi_h * p.Wp + hi_w;
  p.R = (PM + p.lead + tail + P_BOX - 1) / P_BOX * P_BOX;
  if (p.R / P_BOX > 32) return false;
  const int up_h = p.Hp - lo_h - d->Ho, up_w = p.Wp - lo_w - d->Wo;   // TMA upper corners
  if (p.R > 1024 || lo_h > 127 || lo_w > 127 || up_h > 127 || up_w > 127 || up_h < -128 || up_w < -128) return false;
  for (int t = 0; t < nt; ++t) {
    p.tap_cls[t] = (short)taps[t].cls;
    p.tap_wrow[t] = (short)taps[t].wrow;
    p.tap_shift[t] = p.lead + taps[t].dr * p.Wp + taps[t].dc;
  }
  p.Q = (long)d->N * p.Hp * p.Wp;
  p.Cn = Cn; p.kblocks = Ck / PK; p.osh = d->sh; p.osw = d->sw;
  const size_t a_bytes = ((size_t)p.R * 128 + 1023) & ~(size_t)1023;
  const size_t budget = 220 * 1024;
  const size_t tile = (size_t)bn * 128;
  int G = (int)((bn >= 64 ? 64 * 1024 : 32 * 1024) / tile);      // 8, 8, 4 taps for BN = 32, 64, 128
  if (G > nt) G = nt;
  if (G > 32) G = 32;
  if (P_ASTAGES * a_bytes + 2 * G * tile + 1024 > budget) return false;
  long stages = (long)((budget - 1024 - P_ASTAGES * a_bytes) / (G * tile));
  if (stages > P_MAXG) stages = P_MAXG;
  p.G = G;
  p.bstages = (int)stages;
  pl->bn = bn;
  pl->smem = P_ASTAGES * a_bytes + (size_t)p.bstages * G * tile + 1024;
  const long mtiles = (p.Q - p.lead + PM - 1) / PM;
  pl->ctas = mtiles * (Cn / bn);
  return pl->ctas > 0 && pl->ctas < 2147483647L;
}

template <int BN>
int launch_patch(const PatchPlan& pl, cudaStream_t st) {
  static size_t configured = 0;
  if (pl.smem > configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(220 * 1024)));
    configured = 220 * 1024;
  }
  conv_patch_kernel<BN><<<(unsigned)pl.ctas, P_THREADS, pl.smem, st>>>(pl.p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace

bool conv_transposed_patch_eligible(const advoc_conv_desc* d, int ldx, int store_w) {
  static const bool disabled = getenv("ADVOC_NO_PATCH") != nullptr;   // A/B switch for benchmarking
  if (disabled || !(tc::tma_ok() && device_arch() == 100 && ldx % 4 == 0)) return false;
  PatchPlan pl = {};
  return plan_patch(d, store_w ? store_w : d->W, &pl);
}

int conv_transposed_patch(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                          const advoc_epilogue* ep, void* stream) {
  PatchPlan pl = {};
  int st = lower_epilogue(ep, d->H, d->W, d->Cin, &pl.p.epi);
  if (st) return st;
  ADVOC_REQUIRE(plan_patch(d, pl.p.epi.Ws, &pl), ADVOC_UNSUPPORTED, "layer does not fit the patch kernel");
  PatchParams& p = pl.p;
  st = encode_im2col(&p.tmA, x, d->N, d->Ho, d->Wo, ldx, d->Cout, -p.lo_h, -p.lo_w, p.Hp - p.lo_h - d->Ho,
                     p.Wp - p.lo_w - d->Wo, 1, 1, PK, P_BOX);
  if (st) return st;
  st = encode_tiled2d(&p.tmB, w, d->Cout, (long)d->kh * d->kw * d->Cin, (size_t)d->Cout * 4, PK, pl.bn);
  if (st) return st;
  p.dbg = debug_word();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (pl.bn) {
    case 128: return launch_patch<128>(pl, s);
    case 64: return launch_patch<64>(pl, s);
    default: return launch_patch<32>(pl, s);
  }
}

}  // namespace advoc
