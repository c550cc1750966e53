// tcgen05 / TMA implicit-GEMM convolution for sm_100a (TF32 operands, fp32 accumulate in TMEM).
//
// One kernel serves the forward convolution (tf.layers.conv2d, advoc_model.py:27-32,46-51) and
// the transposed convolution (tf.layers.conv2d_transpose, advoc_model.py:65-69; also every
// input-gradient).  GEMM view:  D[pixel, n] = sum_{tap, c} A[pixel, tap, c] * B[tap, n, c]
//   A  activations, NHWC fp32, fetched by TMA in IM2COL mode straight into the 128B-swizzled
//      K-major operand layout -- the tensor map holds the padding (pixel-box corners) and the
//      traversal stride, the filter tap is the per-load offset; no unfold buffer exists.
//   B  filters, K-major [tap][n][c] fp32, fetched by a tiled TMA box.
//   D  128 x BN fp32 accumulator in tensor memory, written by tcgen05.mma.kind::tf32.
// A transposed convolution with stride s is run as s_h*s_w independent "parity classes"
// (interleaved in launch order): every output pixel of a class sees the same dense sub-filter (2x2 taps for
// k4 s2), so no MAC is spent on the zeros a zero-insertion formulation would multiply.
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 TMEM allocator + single-thread MMA issuer,
// warps 2-5 epilogue (TMEM -> registers -> bias / activation / dropout / TF32 rounding ->
// one or two NHWC destinations, channel-offset and width-cropped: advoc_epilogue).
// smem ring of kStages {A 16 KB, B BN*128 B} slots with full/empty mbarriers.
#include "epilogue.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace advoc {

float* wgrad_workspace(size_t bytes, cudaStream_t st);   // wgrad_tc.cu: grow-only library workspace (one stream)

namespace {

using namespace tc;

constexpr int BM = 128;        // UMMA_M (cta_group::1)
constexpr int BK_BYTES = 128;  // one swizzle row of the K-major operands: 32 tf32 or 64 fp16 channels
constexpr int MAX_CLASSES = 4;
constexpr int MAX_TAPS = 25;
constexpr int NUM_THREADS = 224;   // warp 0: activation producer, 1: MMA issuer, 2-5: epilogue, 6: filter producer

struct alignas(64) TcParams {
  CUtensorMap tmA[MAX_CLASSES];
  CUtensorMap tmB;
  int Ah[MAX_CLASSES], Aw[MAX_CLASSES];          // output positions per image, per class
  int ph[MAX_CLASSES], pw[MAX_CLASSES];          // output offset of the class
  int base_h[MAX_CLASSES], base_w[MAX_CLASSES];  // im2col lower corner
  int ntaps[MAX_CLASSES];
  unsigned short tap_off[MAX_CLASSES][MAX_TAPS];   // (offset_h << 8) | offset_w
  unsigned short tap_wrow[MAX_CLASSES][MAX_TAPS];  // filter tap index (row block of B)
  int Nimg, nclasses;
  int trav_h, trav_w;  // im2col traversal stride
  int osh, osw;        // output position stride (1 for conv, s for transposed conv)
  int Cn;              // produced channels (rows of B per tap)
  int kblocks;         // contraction channels / channels per 128-byte row (32 tf32, 64 fp16)
  long total_tiles;    // M tiles (of the largest class) x classes x N tiles x K splits
  // split-K (round 2: the bottleneck layers of the regular model have 8-96 output tiles for 296 CTA slots and
  // stream a 4-17 MB filter through them): every tile's tap / channel-block loop is cut into `ksplit` equal
  // ranges run by different CTAs; each stores its raw fp32 accumulator to ws[split][stored pixel][channel] and
  // splitk_finalize_kernel adds the splits in order and applies the epilogue.  1 = off.
  int ksplit;
  float* ws;
  long ws_stride;
  // merged parity classes (conv_tc_merged_kernel): shift s = (dh + 1) * 3 + (dw + 1) of the 3 x 3 neighbourhood
  // feeds m_cnt[s] (class, tap) pairs: accumulator m_cls[s][j] (= 2 ph + pw) with filter tap m_wrow[s][j]
  // the 9 entries are stored in ISSUE order (centre shift first: it reaches all four classes, so one N = 4 BN MMA
  // initialises every accumulator); m_off = (dh + 1) << 2 | (dw + 1); m_span[s] = how many consecutive pairs of the
  // entry form one MMA (their classes are adjacent accumulators and their filter tiles adjacent in the stage)
  unsigned char m_cnt[9], m_cls[9][4], m_wrow[9][4], m_off[9], m_span[9];
  EpiDev epi;
  unsigned int* dbg;   // [0] != 0 after a barrier wait timed out
};

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK_BYTES;  // 16 KB
  static constexpr int B_BYTES = BN * BK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

// loop invariants of an epilogue (r02: with dropout on, reloading the seed word, re-dividing and hashing with 64-bit
// multiplies made the epilogue the slowest stage of the kernel: profiles/r02K_conv_tc_small_stall_samples.txt)
struct TcEpiInv {
  uint64_t seed;
  float inv_keep;
  bool plain;       // bias + none / relu / lrelu [+ dropout] to one or two fp32 / fp16 destinations
  bool gated;       // input-gradient epilogue: gate (+ accumulate) to one fp32 destination
  ActLin a0, a1;
};
__device__ __forceinline__ TcEpiInv tc_epi_invariants(const EpiDev& e) {
  TcEpiInv q;
  q.seed = e.keep_prob < 1.f ? epi_seed(e) : 0ull;
  q.inv_keep = 1.f / e.keep_prob;
  q.plain = act_is_linear(e.act0) && act_is_linear(e.act1) && !e.gate && !e.accumulate && !e.mask;
  q.gated = e.gate && act_is_linear(e.act0) && !e.mask && e.keep_prob >= 1.f && !e.h0 && !e.out1;
  q.a0 = act_linear(e.act0, e.alpha);
  q.a1 = act_linear(e.act1, e.alpha);
  return q;
}

// 32 consecutive channels [n, n + 32) of stored pixel `pix`: accumulator values v + bias values bv -> destination(s)
// GATED: compile the input-gradient fast path (fp32-operand kernels only: it costs ~25 registers, which the fp16
// forward kernels should not pay -- r02: small forward 19.2 -> 18.7 M mel-frames/s with it compiled in everywhere)
template <bool GATED>
__device__ __forceinline__ void tc_store_chunk(const EpiDev& e, const TcEpiInv& q, size_t pix, int n,
                                               const uint32_t (&v)[32], const float4 (&bv)[8]) {
  if (q.plain) {
          // forward fast path (bias + none / relu / lrelu [+ dropout] to one or two fp32 / fp16 destinations):
          // 32 values of one pixel row, 16-byte stores
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = bv[j >> 2];
            x[j] = __uint_as_float(v[j]) + b.x; x[j + 1] = __uint_as_float(v[j + 1]) + b.y;
            x[j + 2] = __uint_as_float(v[j + 2]) + b.z; x[j + 3] = __uint_as_float(v[j + 3]) + b.w;
          }
          float sc[32];
          if (e.keep_prob < 1.f) {
            const size_t idx = pix * e.Cout + n;
#pragma unroll
            for (int j = 0; j < 32; ++j) sc[j] = dropout_keep(q.seed, idx + j, e.keep_prob) ? q.inv_keep : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sc[j] = 1.f;
          }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            if (o == 1 && !e.out1) break;
            const ActLin a = o == 0 ? q.a0 : q.a1;
            float* base = o == 0 ? e.out0 : e.out1;
            const size_t el = pix * (size_t)(o == 0 ? e.ld0 : e.ld1) + (size_t)((o == 0 ? e.coff0 : e.coff1) + n);
            const bool half = o == 0 ? e.h0 : e.h1;
            if (half) {
              __half* dst = reinterpret_cast<__half*>(base) + el;
              if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8)
                  *reinterpret_cast<uint4*>(dst + j) =
                      make_uint4(pack_half2(apply_lin(x[j], a) * sc[j], apply_lin(x[j + 1], a) * sc[j + 1]),
                                 pack_half2(apply_lin(x[j + 2], a) * sc[j + 2], apply_lin(x[j + 3], a) * sc[j + 3]),
                                 pack_half2(apply_lin(x[j + 4], a) * sc[j + 4], apply_lin(x[j + 5], a) * sc[j + 5]),
                                 pack_half2(apply_lin(x[j + 6], a) * sc[j + 6], apply_lin(x[j + 7], a) * sc[j + 7]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<uint2*>(dst + j) =
                      make_uint2(pack_half2(apply_lin(x[j], a) * sc[j], apply_lin(x[j + 1], a) * sc[j + 1]),
                                 pack_half2(apply_lin(x[j + 2], a) * sc[j + 2], apply_lin(x[j + 3], a) * sc[j + 3]));
              }
            } else {
              float* dst = base + el;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float y[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  y[u] = apply_lin(x[j + u], a) * sc[j + u];
                  if (e.round) y[u] = round_tf32(y[u]);
                }
                *reinterpret_cast<float4*>(dst + j) = make_float4(y[0], y[1], y[2], y[3]);
              }
            }
          }
  } else if (GATED && q.gated) {
    // backward fast path (input gradients): v * act'(gate) * scale[channel < split] (+ old value) -> fp32, in two
    // halves of 16 channels so that the gate / destination vectors in flight stay at 8 registers each
    const float neg = e.gate_act == ADVOC_ACT_LRELU ? e.alpha : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int nn = n + 16 * h;
      const float4* gp = reinterpret_cast<const float4*>(e.gate + pix * (size_t)e.ldg + (size_t)(e.coffg + nn));
      float4* dp = reinterpret_cast<float4*>(e.out0 + pix * (size_t)e.ld0 + (size_t)(e.coff0 + nn));
      float4 gv[4], ov[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = __ldg(gp + j);
#pragma unroll
      for (int j = 0; j < 4; ++j) ov[j] = e.accumulate ? dp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b = bv[4 * h + j];
        const int c = 16 * h + 4 * j;
        const float sgs = (nn + 4 * j) < e.gate_split ? e.gscale0 : e.gscale1;
        float y[4] = {apply_lin(__uint_as_float(v[c]) + b.x, q.a0), apply_lin(__uint_as_float(v[c + 1]) + b.y, q.a0),
                      apply_lin(__uint_as_float(v[c + 2]) + b.z, q.a0), apply_lin(__uint_as_float(v[c + 3]) + b.w, q.a0)};
        const float g[4] = {gv[j].x, gv[j].y, gv[j].z, gv[j].w};
        const float o[4] = {ov[j].x, ov[j].y, ov[j].z, ov[j].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          y[u] *= (g[u] > 0.f ? 1.f : neg) * sgs;
          if (e.accumulate) y[u] += o[u];
          if (e.round) y[u] = round_tf32(y[u]);
        }
        dp[j] = make_float4(y[0], y[1], y[2], y[3]);
      }
    }
  } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = bv[j >> 2];
            epi_store_vec4_core(e, pix, n + j,
                                make_float4(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y,
                                            __uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w),
                                q.seed, q.inv_keep);
          }
  }
}

// HALF: fp16 operands (kind::f16, 64 channels per 128-byte row) instead of tf32 (32 channels); the
// byte layout of the ring, the descriptors and the four 32-byte K steps per row are the same.
//
// PERSISTENT (round 2): the grid is one or two CTAs per SM and every CTA walks the tile list
// t = blockIdx.x, blockIdx.x + gridDim.x, ...; the producer warp keeps streaming operand tiles across
// tile boundaries and the accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of
// tile i (TMEM -> registers -> global) overlaps the main loop of tile i + 1 and the pipeline never
// drains between tiles.  (Round 1 ran one tile per CTA: with the K loops of the fp16 path only 16-64
// ring slots long, pipeline fill + epilogue + TMEM alloc were a large share of every CTA's life:
// profiles/r02c_ncu_fwd_f16.csv, tensor pipe 8-17 % active.)
// Tile order: N tile fastest, then parity class, then M tile, so that the CTAs running at the same
// time share their A tiles through L2.
struct TcTile {
  int n0, cls, iters, it0, ks;
  long m0;
  bool valid;
};

template <int BN>
__device__ __forceinline__ TcTile tc_decode(const TcParams& p, long t, int n_ntiles) {
  TcTile x;
  unsigned tt = (unsigned)t;                        // total tiles < 2^31 (checked on the host)
  x.ks = 0;
  if (p.ksplit > 1) {                               // K split fastest: the CTAs of one tile share its A tiles in L2
    x.ks = (int)(tt % (unsigned)p.ksplit);
    tt /= (unsigned)p.ksplit;
  }
  const unsigned per_m = (unsigned)(n_ntiles * p.nclasses);
  const unsigned m_tile = tt / per_m;
  const unsigned rem = tt - m_tile * per_m;
  x.cls = (int)(rem / (unsigned)n_ntiles);
  x.n0 = (int)(rem - (unsigned)x.cls * (unsigned)n_ntiles) * BN;
  x.m0 = (long)m_tile * BM;
  x.valid = x.m0 < (long)p.Nimg * p.Ah[x.cls] * p.Aw[x.cls];   // small classes have fewer M tiles
  x.iters = p.ntaps[x.cls] * p.kblocks / p.ksplit;   // ksplit divides every class's loop (host)
  x.it0 = x.ks * x.iters;
  return x;
}

template <int BN, int STAGES, bool HALF>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int BK = HALF ? 64 : 32;   // channels per k-block
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_holder;

  const int n_ntiles = p.Cn / BN;
  const long ntl = p.total_tiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // operand ring, 1024-byte aligned (SWIZZLE_128B atom)
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulators

  if (warp == 0 && lane == 0) {
    for (int c = 0; c < p.nclasses; ++c) prefetch_tmap(&p.tmA[c]);
    prefetch_tmap(&p.tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);    // the activation producer and the filter producer each arrive with their bytes
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
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
  const bool aborted = p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0;

  if (aborted) {
    // a previous launch timed out: do nothing (the host raises at its next sync point)
  } else if (warp == 0 || warp == 6) {
    // ===== TMA producers: warp 0 streams the activation (im2col) tiles, warp 6 the filter tiles, each through
    // one elected lane.  One thread spends ~540 cycles per {wait, expect_tx, issue} chain and ~150 more per extra
    // TMA issue (profiles/r02r_tma_issue_probe.txt); with both loads on one thread a ring slot cost more issue
    // time than the 192-512 tensor cycles it holds. =====
    const bool filt = warp == 6;
    int stage = 0;
    uint32_t phase = 0;
    for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
      const TcTile tl = tc_decode<BN>(p, t, n_ntiles);
      if (!tl.valid) continue;
      const int cls = tl.cls;
      const int Ah = p.Ah[cls], Aw = p.Aw[cls];
      const int a_w = (int)(tl.m0 % Aw);
      const long r = tl.m0 / Aw;
      const int a_h = (int)(r % Ah);
      const int img = (int)(r / Ah);
      const int cw = a_w * p.trav_w + p.base_w[cls];
      const int ch = a_h * p.trav_h + p.base_h[cls];
      int tap = tl.it0 / p.kblocks, kb = tl.it0 - tap * p.kblocks;
      for (int it = 0; it < tl.iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.dbg, 1u);
        uint8_t* a_dst = ring_ptr + stage * L::STAGE_BYTES;
        const unsigned off = p.tap_off[cls][tap];
        const int wrow = (int)p.tap_wrow[cls][tap] * p.Cn + tl.n0;
        __syncwarp();
        if (elect_one()) {
          if (filt) {
            mbar_expect_tx(&full_bar[stage], L::B_BYTES);
            tma_load_2d(&p.tmB, &full_bar[stage], a_dst + L::A_BYTES, kb * BK, wrow);
          } else {
            mbar_expect_tx(&full_bar[stage], L::A_BYTES);
            tma_load_im2col_4d(&p.tmA[cls], &full_bar[stage], a_dst, kb * BK, cw, ch, img, (uint16_t)(off & 0xFF),
                               (uint16_t)(off >> 8));
          }
        }
        __syncwarp();
        if (++kb == p.kblocks) { kb = 0; ++tap; }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, tcgen05.mma issued by one elected lane (a plain
    // `if (lane == 0)` region makes the compiler wrap every UTCMMA in an ELECT/BRA.U.ANY loop that
    // costs ~140 cycles per instruction: scripts/dev_mma_rate.py) =====
    constexpr uint32_t idesc = umma_idesc<HALF>(BM, BN);
    if (elect_one()) {   // one thread waits and issues; entered through elect.sync (see tc::elect_one)
      int stage = 0;
      uint32_t phase = 0;
      uint32_t i = 0;   // tiles this CTA has started
      for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
        const TcTile tl = tc_decode<BN>(p, t, n_ntiles);
        if (!tl.valid) continue;
        const uint32_t buf = i & 1u;
        mbar_wait(&acc_empty[buf], ((i >> 1) & 1u) ^ 1u, p.dbg, 4u);   // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)BN;
        for (int it = 0; it < tl.iters; ++it) {
          mbar_wait(&full_bar[stage], phase, p.dbg, 2u);
          const uint32_t a_addr = ring + stage * L::STAGE_BYTES;
          const uint64_t da = make_smem_desc(a_addr);
          const uint64_t db = make_smem_desc(a_addr + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // advance 32 bytes (= 2 x 16-byte units) along K inside the swizzle atom
            umma_op<HALF>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&acc_full[buf]);       // accumulator complete -> epilogue
        ++i;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp q may only touch TMEM lanes [32q, 32q+32) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t i = 0;
    // loop invariants of the epilogue (r02: with dropout on, reloading the seed word, re-dividing and hashing with
    // 64-bit multiplies made this the slowest stage of the kernel: profiles/r02K_conv_tc_small_stall_samples.txt)
    const TcEpiInv inv = tc_epi_invariants(p.epi);
    const float* bias = p.epi.bias;
    for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
      const TcTile tl = tc_decode<BN>(p, t, n_ntiles);
      if (!tl.valid) continue;
      const int cls = tl.cls;
      const int Ah = p.Ah[cls], Aw = p.Aw[cls];
      const long m = tl.m0 + row;
      const bool valid = m < (long)p.Nimg * Ah * Aw;
      size_t pix = 0;
      if (valid) {
        const int a_w = (int)(m % Aw);
        const long r = m / Aw;
        const int a_h = (int)(r % Ah);
        const long img = r / Ah;
        pix = ((size_t)img * p.epi.Hs + (size_t)(a_h * p.osh + p.ph[cls])) * p.epi.Ws +
              (size_t)(a_w * p.osw + p.pw[cls]);
      }
      const uint32_t buf = i & 1u;
      mbar_wait(&acc_full[buf], (i >> 1) & 1u, p.dbg, 3u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        float4 bv[8];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)BN + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)      // the chunk's 32 bias values, in flight together with the TMEM load
          bv[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + tl.n0 + c0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_ld_wait();
        if (valid && p.ksplit > 1) {
          float4* dst = reinterpret_cast<float4*>(p.ws + (size_t)tl.ks * p.ws_stride + pix * p.Cn + tl.n0 + c0);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            dst[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                      __uint_as_float(v[j + 3]));
        } else if (valid) {
          tc_store_chunk<!HALF>(p.epi, inv, pix, tl.n0 + c0, v, bv);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
      }
      ++i;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// k4 s2 transposed convolution with the FOUR PARITY CLASSES MERGED (round 2, end).  The per-class schedule above
// fetches every input pixel once per (class, tap): 16 im2col tiles per 128 positions and k-block, and at N <= 64
// that L2 -> SM stream paces the layer (AdVoc-small decoder_2 on the per-class schedule: 1.3 GB in 90 us).  Here a
// tile is 128 positions (a, b) of the INPUT grid and owns all four outputs (2a + ph, 2b + pw): the nine shifted
// views (a + dh, b + dw), dh, dw in {-1, 0, 1}, are fetched once each and every view feeds the 1, 2 or 4 (class,
// tap) pairs that read it -- class (ph, pw) sees tap kh at dh = (ph + 1 - kh) / 2 -- into four accumulators that
// sit side by side in TMEM.  9 A tiles + 16 filter tiles per k-block instead of 16 + 16.
// ---------------------------------------------------------------------------------------------
// ACC_BUFS = 1: single accumulator set (64-channel tiles: 256 TMEM columns, so that two CTAs still share an SM and
// overlap each other's epilogues)
template <int BN, int STAGES, bool HALF, int ACC_BUFS>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_merged_kernel(const __grid_constant__ TcParams p) {
  constexpr int BK = HALF ? 64 : 32;
  constexpr int A_BYTES = BM * BK_BYTES, B_BYTES = BN * BK_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + 4 * B_BYTES;
  constexpr uint32_t ACC_COLS = 4 * BN;                       // four class accumulators
  constexpr uint32_t TMEM_COLS = ACC_BUFS * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_holder;

  const int n_ntiles = p.Cn / BN;
  const long ntl = p.total_tiles;                             // M tiles x N tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ah = p.Ah[0], Aw = p.Aw[0];                       // the input-grid extent the tiles walk
  const long Mtot = (long)p.Nimg * Ah * Aw;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[0]);
    prefetch_tmap(&p.tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
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
  const bool aborted = p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0;

  if (aborted) {
  } else if (warp == 0 || warp == 6) {
    // ===== TMA producers: warp 0 the nine shifted activation tiles, warp 6 the filter tiles of their pairs =====
    const bool filt = warp == 6;
    int stage = 0;
    uint32_t phase = 0;
    for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
      const unsigned mt = (unsigned)t / (unsigned)n_ntiles;
      const int n0 = (int)((unsigned)t - mt * (unsigned)n_ntiles) * BN;
      const long m0 = (long)mt * BM;
      const int b = (int)(m0 % Aw);
      const long r = m0 / Aw;
      const int a = (int)(r % Ah);
      const int img = (int)(r / Ah);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        for (int s = 0; s < 9; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, p.dbg, 1u);
          uint8_t* dst = ring_ptr + stage * STAGE_BYTES;
          const int cnt = p.m_cnt[s];
          __syncwarp();
          if (elect_one()) {
            if (filt) {
              mbar_expect_tx(&full_bar[stage], (uint32_t)cnt * B_BYTES);
              for (int j = 0; j < cnt; ++j)
                tma_load_2d(&p.tmB, &full_bar[stage], dst + A_BYTES + j * B_BYTES, kb * BK, (int)p.m_wrow[s][j] * p.Cn + n0);
            } else {
              mbar_expect_tx(&full_bar[stage], A_BYTES);
              tma_load_im2col_4d(&p.tmA[0], &full_bar[stage], dst, kb * BK, b + p.base_w[0], a + p.base_h[0], img,
                                 (uint16_t)(p.m_off[s] & 3), (uint16_t)(p.m_off[s] >> 2));
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc<HALF>(BM, BN);
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, i = 0;
      for (long t = blockIdx.x; t < ntl; t += gridDim.x, ++i) {
        const uint32_t buf = i % ACC_BUFS, use = i / ACC_BUFS;
        mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u, p.dbg, 4u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * ACC_COLS;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          for (int s = 0; s < 9; ++s) {
            mbar_wait(&full_bar[stage], phase, p.dbg, 2u);
            const uint32_t a_addr = ring + stage * STAGE_BYTES;
            const uint64_t da = make_smem_desc(a_addr);
            const int cnt = p.m_cnt[s], span = p.m_span[s];
            // (kb, s) = (0, 0) is the centre shift: one MMA over all four classes zeroes the accumulators
            const uint32_t keep = (kb | s) != 0 ? 1u : 0u;
            for (int j = 0; j < cnt; j += span) {
              const uint32_t cls = p.m_cls[s][j];
              const uint64_t db = make_smem_desc(a_addr + A_BYTES + j * B_BYTES);
              const uint32_t idesc = span == 4 ? umma_idesc<HALF>(BM, 4 * BN)
                                               : (span == 2 ? umma_idesc<HALF>(BM, 2 * BN) : umma_idesc<HALF>(BM, BN));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_op<HALF>(d0 + cls * (uint32_t)BN, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                              (k != 0 || keep) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: row = one input-grid position, four output pixels =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t i = 0;
    const TcEpiInv inv = tc_epi_invariants(p.epi);
    const float* bias = p.epi.bias;
    for (long t = blockIdx.x; t < ntl; t += gridDim.x, ++i) {
      const unsigned mt = (unsigned)t / (unsigned)n_ntiles;
      const int n0 = (int)((unsigned)t - mt * (unsigned)n_ntiles) * BN;
      const long m = (long)mt * BM + row;
      int a = 0, b = 0;
      long img = 0;
      if (m < Mtot) {
        b = (int)(m % Aw);
        const long r = m / Aw;
        a = (int)(r % Ah);
        img = r / Ah;
      }
      const uint32_t buf = i % ACC_BUFS, use = i / ACC_BUFS;
      mbar_wait(&acc_full[buf], use & 1u, p.dbg, 3u);
      tc_fence_after();
#pragma unroll 1
      for (int cls = 0; cls < 4; ++cls) {
        const int oh = 2 * a + (cls >> 1), ow = 2 * b + (cls & 1);
        const bool valid = m < Mtot && oh < p.epi.Hs && ow < p.epi.Ws;
        const size_t pix = ((size_t)img * p.epi.Hs + (size_t)oh) * p.epi.Ws + (size_t)ow;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          float4 bv[8];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + (uint32_t)(cls * BN + c0), v);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bv[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + n0 + c0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          if (valid) tc_store_chunk<false>(p.epi, inv, pix, n0 + c0, v, bv);   // (transposed forward layers only)
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
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

template <int BN, int STAGES, bool HALF, int ACC_BUFS>
int launch_merged(const TcParams& p, long m_tiles, cudaStream_t st) {
  constexpr int smem = STAGES * (BM * BK_BYTES + 4 * BN * BK_BYTES) + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_merged_kernel<BN, STAGES, HALF, ACC_BUFS>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  TcParams q = p;
  q.ksplit = 1;
  q.total_tiles = m_tiles * (p.Cn / BN);
  ADVOC_REQUIRE(q.total_tiles < 2147483647L, ADVOC_BAD_SHAPE, "too many output tiles");
  const int per_sm = (2 * smem <= 220 * 1024 && 2 * ACC_BUFS * 4 * BN <= 512) ? 2 : 1;
  const long slots = (long)sm_count() * per_sm;
  conv_tc_merged_kernel<BN, STAGES, HALF, ACC_BUFS><<<(unsigned)(q.total_tiles < slots ? q.total_tiles : slots), NUM_THREADS, smem, st>>>(q);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

// sum of the K splits (in split order: deterministic) + the layer's epilogue, four channels per thread
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, long ws_stride, int ksplit,
                                                              long npix, int Cn, const EpiDev e) {
  const int quads = Cn >> 2;
  const long total = npix * quads;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / quads;
    const int n = (int)(i - pix * quads) << 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ws + (size_t)s * ws_stride + (size_t)pix * Cn + n));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    epi_store_vec4(e, (size_t)pix, n, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: driver entry points (no link-time libcuda dependency), tensor maps, launch
// ---------------------------------------------------------------------------------------------
struct ClassGeom {
  int Ah, Aw, ph, pw, lower_h, lower_w, upper_h, upper_w, ntaps;
  unsigned short off[MAX_TAPS], wrow[MAX_TAPS];
};

inline int bk_of(int half) { return half ? 64 : 32; }

// im2col map over activations [Nimg, Hin, Win, ld] reading `Ck` channels per pixel
int encode_A(CUtensorMap* tm, const void* x, int Nimg, int Hin, int Win, int ld, int Ck, const ClassGeom& g,
             int trav_h, int trav_w, int half) {
  return encode_im2col(tm, x, Nimg, Hin, Win, ld, Ck, g.lower_h, g.lower_w, g.upper_h, g.upper_w, trav_h, trav_w,
                       bk_of(half), BM, false, half);
}

int encode_B(CUtensorMap* tm, const void* w, int rows, int Ck, int bn, int half) {
  return encode_tiled2d(tm, w, Ck, rows, (size_t)Ck * (half ? 2 : 4), bk_of(half), bn, false, half);
}

template <int BN, int STAGES, bool HALF>
int launch(const TcParams& p, int nclasses, long max_tiles, cudaStream_t st) {
  constexpr int smem = STAGES * SmemLayout<BN>::STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem));
    configured = true;
  }
  long tiles = max_tiles * (p.Cn / BN) * nclasses;
  ADVOC_REQUIRE(tiles < 2147483647L, ADVOC_BAD_SHAPE, "too many output tiles");
  TcParams q = p;
  // persistent: as many CTAs as fit on the chip at once (two per SM while two rings and two pairs of
  // accumulators fit: 2 x smem <= 227 KB and 2 x 2 x BN <= 512 TMEM columns)
  const int per_sm = (2 * smem <= 220 * 1024 && 4 * BN <= 512) ? 2 : 1;
  const long slots = (long)sm_count() * per_sm;
  // split-K when the layer cannot fill half of the slots: the largest power of two that divides every class's
  // loop, leaves >= 8 ring iterations per split and keeps tiles x splits within the slots
  q.ksplit = 1;
  static const bool no_split = getenv("ADVOC_TC_NO_SPLITK") != nullptr;   // A/B switch
  const long npix = (long)p.Nimg * p.epi.Hs * p.epi.Ws;
  if (!no_split && tiles * 2 <= slots) {
    int ks = 1;
    for (;;) {
      const int next = ks * 2;
      bool ok = tiles * next <= slots && (size_t)next * npix * p.Cn * sizeof(float) <= ((size_t)48 << 20);
      for (int c = 0; c < nclasses && ok; ++c) {
        const int total = p.ntaps[c] * p.kblocks;
        ok = total % next == 0 && total / next >= 8;
      }
      if (!ok) break;
      ks = next;
    }
    if (ks > 1) {
      q.ws = wgrad_workspace((size_t)ks * npix * p.Cn * sizeof(float), st);
      if (q.ws) { q.ksplit = ks; q.ws_stride = npix * p.Cn; }
    }
  }
  tiles *= q.ksplit;
  q.total_tiles = tiles;
  dim3 grid((unsigned)(tiles < slots ? tiles : slots), 1, 1);
  conv_tc_kernel<BN, STAGES, HALF><<<grid, NUM_THREADS, smem, st>>>(q);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  if (q.ksplit > 1) {
    const long work = npix * (p.Cn / 4);
    const long blocks = (work + 255) / 256;
    splitk_finalize_kernel<<<(unsigned)(blocks < 4L * sm_count() ? blocks : 4L * sm_count()), 256, 0, st>>>(
        q.ws, q.ws_stride, q.ksplit, npix, p.Cn, p.epi);
    count_launch();
    ADVOC_CHECK_CUDA(cudaGetLastError());
  }
  return ADVOC_OK;
}

// Widest N tile that still leaves about two waves of CTAs: a layer with few output pixels and
// many channels (encoder_5: 34 M tiles x 256 channels) otherwise runs on a fraction of the SMs
// with one long serial K loop per CTA.  `m_ctas` = M tiles x parity classes.
int pick_bn(int Cn, long m_ctas) {
  static const bool fixed = getenv("ADVOC_TC_WIDE_N") != nullptr;   // A/B switch: always the widest tile
  int bn = Cn % 256 == 0 ? 256 : (Cn % 128 == 0 ? 128 : (Cn % 64 == 0 ? 64 : 32));
  if (fixed) return bn;
  while (bn > 64 && m_ctas * (Cn / bn) < 2L * sm_count()) bn >>= 1;
  return bn;
}

long m_ctas_of(int Nimg, const ClassGeom* g, int nclasses) {
  long tiles = 0;
  for (int c = 0; c < nclasses; ++c) {
    const long t = ((long)Nimg * g[c].Ah * g[c].Aw + BM - 1) / BM;
    if (t > tiles) tiles = t;
  }
  return tiles * nclasses;
}

int run(TcParams& p, int nclasses, const ClassGeom* g, void* stream, int half) {
  long max_tiles = 0;
  for (int c = 0; c < nclasses; ++c) {
    p.Ah[c] = g[c].Ah; p.Aw[c] = g[c].Aw; p.ph[c] = g[c].ph; p.pw[c] = g[c].pw;
    p.base_h[c] = g[c].lower_h; p.base_w[c] = g[c].lower_w; p.ntaps[c] = g[c].ntaps;
    for (int t = 0; t < g[c].ntaps; ++t) { p.tap_off[c][t] = g[c].off[t]; p.tap_wrow[c][t] = g[c].wrow[t]; }
    const long tiles = ((long)p.Nimg * g[c].Ah * g[c].Aw + BM - 1) / BM;
    if (tiles > max_tiles) max_tiles = tiles;
  }
  if (max_tiles == 0) return ADVOC_OK;
  ADVOC_REQUIRE(max_tiles < 2147483647L, ADVOC_BAD_SHAPE, "too many output tiles");
  p.dbg = debug_word();
  p.nclasses = nclasses;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bn = pick_bn(p.Cn, m_ctas_of(p.Nimg, g, nclasses));
  if (half) {
    switch (bn) {
      case 256: return launch<256, 4, true>(p, nclasses, max_tiles, st);
      case 128: return launch<128, 3, true>(p, nclasses, max_tiles, st);
      case 64: return launch<64, 4, true>(p, nclasses, max_tiles, st);
      default: return launch<32, 4, true>(p, nclasses, max_tiles, st);
    }
  }
  switch (bn) {
    case 256: return launch<256, 4, false>(p, nclasses, max_tiles, st);
    case 128: return launch<128, 3, false>(p, nclasses, max_tiles, st);
    case 64: return launch<64, 4, false>(p, nclasses, max_tiles, st);
    default: return launch<32, 4, false>(p, nclasses, max_tiles, st);
  }
}

// fp16 operands: 64 channels per k-block, pixel stride a multiple of 16 bytes
bool common_eligible(int Ck, int Cn, int ldx, int half) {
  return tma_ok() && device_arch() == 100 && Ck % bk_of(half) == 0 && Cn % 32 == 0 && ldx % (half ? 8 : 4) == 0;
}

bool epilogue_vector_ok(const advoc_epilogue* ep) {
  if (!ep) return true;
  // vector stores of 4 channels: 16 bytes (fp32) or 8 bytes (fp16) -- the same element multiples
  auto ok = [](const float* p, int ld, int co) { return aligned16(p) && ld % 4 == 0 && co % 4 == 0; };
  if (!ok(ep->d_out0, ep->ld0, ep->c_off0)) return false;
  if (ep->d_out1 && !ok(ep->d_out1, ep->ld1, ep->c_off1)) return false;
  if (ep->d_bias && !aligned16(ep->d_bias)) return false;
  if (ep->d_gate && !(ok(ep->d_gate, ep->ld_gate, ep->c_off_gate) && ep->gate_split % 4 == 0)) return false;
  return true;
}

}  // namespace

bool tc_epilogue_ok(const advoc_epilogue* ep) { return epilogue_vector_ok(ep); }

// N tile (template BN) a per-tap launch with this geometry would use; mirrors conv_fwd_tc /
// conv_transposed_tc (host-side query for tests and bench attribution)
int conv_tc_tile_n(const advoc_conv_desc* d, int transposed, int store_w) {
  if (!transposed) {
    ClassGeom g = {};
    g.Ah = d->Ho; g.Aw = d->Wo;
    return pick_bn(d->Cout, m_ctas_of(d->N, &g, 1));
  }
  const int Hout = d->H, Wout = store_w > 0 ? store_w : d->W;
  ClassGeom g[MAX_CLASSES] = {};
  int nc = 0;
  for (int ph = 0; ph < d->sh; ++ph)
    for (int pw = 0; pw < d->sw; ++pw) {
      if (nc >= MAX_CLASSES) return 0;
      g[nc].Ah = Hout > ph ? (Hout - ph + d->sh - 1) / d->sh : 0;
      g[nc].Aw = Wout > pw ? (Wout - pw + d->sw - 1) / d->sw : 0;
      if (g[nc].Ah == 0 || g[nc].Aw == 0) continue;
      ++nc;
    }
  return pick_bn(d->Cin, m_ctas_of(d->N, g, nc));
}

bool conv_fwd_tc_eligible(const advoc_conv_desc* d, int ldx) {
  return common_eligible(d->Cin, d->Cout, ldx, d->math == ADVOC_MATH_F16) && d->kh * d->kw <= MAX_TAPS &&
         d->pad_t <= 127 && d->pad_l <= 127;
}

bool conv_transposed_tc_eligible(const advoc_conv_desc* d, int ldx) {
  return common_eligible(d->Cout, d->Cin, ldx, d->math == ADVOC_MATH_F16) && d->sh * d->sw <= MAX_CLASSES &&
         d->kh * d->kw <= MAX_TAPS;
}

// y = conv(x, w): w packed K-major [tap][Cout][Cin]
int conv_fwd_tc(const advoc_conv_desc* d, const void* x, int ldx, const void* w, const advoc_epilogue* ep,
                void* stream) {
  const int half = d->math == ADVOC_MATH_F16;
  ADVOC_REQUIRE(aligned16(x) && aligned16(w), ADVOC_BAD_ALIGN, "x / w must be 16-byte aligned");
  ADVOC_REQUIRE(epilogue_vector_ok(ep), ADVOC_BAD_ALIGN,
                "tcgen05 path needs 16-byte aligned outputs with ld and channel offset multiples of 4");
  TcParams p = {};
  int st = lower_epilogue(ep, d->Ho, d->Wo, d->Cout, &p.epi);
  if (st) return st;
  ADVOC_REQUIRE(p.epi.Ws == d->Wo, ADVOC_UNSUPPORTED, "store_w crop is only supported on conv_transpose");
  ClassGeom g = {};
  g.Ah = d->Ho; g.Aw = d->Wo; g.ph = 0; g.pw = 0;
  g.lower_h = -d->pad_t; g.lower_w = -d->pad_l;
  // last traversed base position = lower + (A-1)*stride = (size - 1) + upper
  g.upper_h = g.lower_h + (d->Ho - 1) * d->sh - (d->H - 1);
  g.upper_w = g.lower_w + (d->Wo - 1) * d->sw - (d->W - 1);
  g.ntaps = d->kh * d->kw;
  for (int t = 0; t < g.ntaps; ++t) {
    g.off[t] = (unsigned short)(((t / d->kw) << 8) | (t % d->kw));
    g.wrow[t] = (unsigned short)t;
  }
  ADVOC_REQUIRE(g.upper_h >= -128 && g.upper_h <= 127 && g.upper_w >= -128 && g.upper_w <= 127, ADVOC_UNSUPPORTED,
                "im2col corner out of range");
  st = encode_A(&p.tmA[0], x, d->N, d->H, d->W, ldx, d->Cin, g, d->sh, d->sw, half);
  if (st) return st;
  st = encode_B(&p.tmB, w, g.ntaps * d->Cout, d->Cin, pick_bn(d->Cout, m_ctas_of(d->N, &g, 1)), half);
  if (st) return st;
  p.Nimg = d->N; p.trav_h = d->sh; p.trav_w = d->sw; p.osh = 1; p.osw = 1;
  p.Cn = d->Cout; p.kblocks = d->Cin / bk_of(half);
  return run(p, 1, &g, stream, half);
}

// y = conv_transpose(x, w): x [N,Ho,Wo,Cout] (small side), y [N,H,Ws,Cin]; w K-major [tap][Cin][Cout]
// (= the TF conv2d_transpose layout HWOI, and = HWIO of a conv whose input gradient this is)
int conv_transposed_tc(const advoc_conv_desc* d, const void* x, int ldx, const void* w,
                       const advoc_epilogue* ep, void* stream) {
  const int half = d->math == ADVOC_MATH_F16;
  ADVOC_REQUIRE(aligned16(x) && aligned16(w), ADVOC_BAD_ALIGN, "x / w must be 16-byte aligned");
  ADVOC_REQUIRE(epilogue_vector_ok(ep), ADVOC_BAD_ALIGN,
                "tcgen05 path needs 16-byte aligned outputs with ld and channel offset multiples of 4");
  TcParams p = {};
  int st = lower_epilogue(ep, d->H, d->W, d->Cin, &p.epi);
  if (st) return st;
  const int Hout = d->H, Wout = p.epi.Ws;  // stored extent
  ClassGeom g[MAX_CLASSES] = {};
  int nc = 0;
  for (int ph = 0; ph < d->sh; ++ph) {
    for (int pw = 0; pw < d->sw; ++pw) {
      ClassGeom& c = g[nc];
      const int rh = (ph + d->pad_t) % d->sh, rw = (pw + d->pad_l) % d->sw;
      const int Jh = (d->kh - rh + d->sh - 1) / d->sh, Jw = (d->kw - rw + d->sw - 1) / d->sw;
      const int q0h = (ph + d->pad_t - rh) / d->sh, q0w = (pw + d->pad_l - rw) / d->sw;
      c.ph = ph; c.pw = pw;
      c.Ah = Hout > ph ? (Hout - ph + d->sh - 1) / d->sh : 0;
      c.Aw = Wout > pw ? (Wout - pw + d->sw - 1) / d->sw : 0;
      if (c.Ah == 0 || c.Aw == 0 || Jh <= 0 || Jw <= 0) {
        ADVOC_REQUIRE(Jh > 0 && Jw > 0, ADVOC_UNSUPPORTED, "parity class without filter taps");
        continue;
      }
      c.lower_h = q0h - (Jh - 1); c.lower_w = q0w - (Jw - 1);
      c.upper_h = c.Ah + c.lower_h - d->Ho; c.upper_w = c.Aw + c.lower_w - d->Wo;
      c.ntaps = Jh * Jw;
      for (int oh = 0; oh < Jh; ++oh)
        for (int ow = 0; ow < Jw; ++ow) {
          const int kh = rh + (Jh - 1 - oh) * d->sh, kw = rw + (Jw - 1 - ow) * d->sw;
          c.off[oh * Jw + ow] = (unsigned short)((oh << 8) | ow);
          c.wrow[oh * Jw + ow] = (unsigned short)(kh * d->kw + kw);
        }
      ADVOC_REQUIRE(c.lower_h >= -128 && c.upper_h <= 127 && c.lower_w >= -128 && c.upper_w <= 127 &&
                        c.upper_h >= -128 && c.upper_w >= -128,
                    ADVOC_UNSUPPORTED, "im2col corner out of range");
      st = encode_A(&p.tmA[nc], x, d->N, d->Ho, d->Wo, ldx, d->Cout, c, 1, 1, half);
      if (st) return st;
      ++nc;
    }
  }
  const int bn = pick_bn(d->Cin, m_ctas_of(d->N, g, nc));
  p.Nimg = d->N; p.trav_h = 1; p.trav_w = 1; p.osh = d->sh; p.osw = d->sw;
  p.Cn = d->Cin; p.kblocks = d->Cout / bk_of(half);
  // merged parity classes (conv_tc_merged_kernel) for the k4 s2 layers with narrow N tiles and many positions
  static const bool no_merge = getenv("ADVOC_TC_NO_MERGE") != nullptr;   // A/B switch
  // (r02R-r02S: BN = 64 is slower than the per-class schedule both with two accumulator sets -- all 512 TMEM columns,
  //  one CTA per SM: AdVoc-small decoder_3 64-73 us against 54, regular decoder_2 199-235 against 196 -- and with one
  //  set at two CTAs per SM and a two-slot ring: 66 against 56, 213 against 193; BN = 32: decoder_2 71 against 91)
  static const int merge_max_bn = getenv("ADVOC_TC_MERGE_MAX_BN") ? atoi(getenv("ADVOC_TC_MERGE_MAX_BN")) : 32;
  const int mAh = (Hout + 1) / 2, mAw = (Wout + 1) / 2;
  const long m_tiles = ((long)d->N * mAh * mAw + BM - 1) / BM;
  if (!no_merge && nc == 4 && d->kh == 4 && d->kw == 4 && d->sh == 2 && d->sw == 2 && d->pad_t == 1 && d->pad_l == 1 &&
      (bn == 32 || bn == 64) && bn <= merge_max_bn && m_tiles * (d->Cin / bn) >= 2L * sm_count() &&
      mAh - 1 - d->Ho >= -128 && mAw - 1 - d->Wo >= -128 && mAh - 1 - d->Ho <= 127 && mAw - 1 - d->Wo <= 127) {
    ClassGeom m = {};
    m.Ah = mAh; m.Aw = mAw; m.lower_h = -1; m.lower_w = -1;
    m.upper_h = mAh - 1 - d->Ho; m.upper_w = mAw - 1 - d->Wo;
    st = encode_A(&p.tmA[0], x, d->N, d->Ho, d->Wo, ldx, d->Cout, m, 1, 1, half);
    if (st) return st;
    st = encode_B(&p.tmB, w, d->kh * d->kw * d->Cin, d->Cout, bn, half);
    if (st) return st;
    p.Ah[0] = mAh; p.Aw[0] = mAw; p.base_h[0] = -1; p.base_w[0] = -1;
    // entry e of the issue order <- shift (dh, dw); the centre first
    static const int order[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
    for (int e = 0; e < 9; ++e) {
      const int s9 = order[e], dh = s9 / 3 - 1, dw = s9 % 3 - 1;
      p.m_cnt[e] = 0;
      p.m_off[e] = (unsigned char)(((dh + 1) << 2) | (dw + 1));
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw)
          for (int kh = 0; kh < 4; ++kh)
            for (int kw = 0; kw < 4; ++kw) {
              if (((ph + 1 - kh) & 1) || ((pw + 1 - kw) & 1)) continue;     // tap of the other parity
              if ((ph + 1 - kh) / 2 != dh || (pw + 1 - kw) / 2 != dw) continue;
              const int j = p.m_cnt[e]++;
              p.m_cls[e][j] = (unsigned char)(2 * ph + pw);
              p.m_wrow[e][j] = (unsigned char)(kh * 4 + kw);
            }
      // pairs in class order: 4 pairs = classes 0..3, 2 pairs = classes (c, c + 1) when dw == 0, (c, c + 2) otherwise
      p.m_span[e] = 1;
      if (p.m_cnt[e] == 4) p.m_span[e] = 4;
      else if (p.m_cnt[e] == 2 && p.m_cls[e][1] == p.m_cls[e][0] + 1) p.m_span[e] = 2;
    }
    p.dbg = debug_word();
    p.nclasses = 1;
    cudaStream_t cst = reinterpret_cast<cudaStream_t>(stream);
    if (half) return bn == 32 ? launch_merged<32, 3, true, 2>(p, m_tiles, cst) : launch_merged<64, 2, true, 1>(p, m_tiles, cst);
    return bn == 32 ? launch_merged<32, 3, false, 2>(p, m_tiles, cst) : launch_merged<64, 2, false, 1>(p, m_tiles, cst);
  }
  st = encode_B(&p.tmB, w, d->kh * d->kw * d->Cin, d->Cout, bn, half);
  if (st) return st;
  return run(p, nc, g, stream, half);
}

// ---------------------------------------------------------------------------------------------
// Convolution TO ONE channel (PatchGAN head, advoc_model.py:196-199) as a tensor-core GEMM + a gather:
//   T[q][tap] = sum_c x[q][c] w[tap][c]        1x1 "convolution" to 32 channels on conv_tc_kernel (taps 0..15
//                                               are the filter taps, the other rows of the packed filter are 0)
//   y[p] = act(bias + sum_taps T[p s - pad + tap][tap])
// The CUDA-core kernel it replaces (train.cu: conv_to_one_kernel, one warp per output pixel) re-reads every
// input pixel once per tap through L2: 166 us for the 128 MB input of the regular model's head at B = 32.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) pack_to_one_filter_kernel(const float* __restrict__ w, float* __restrict__ wq,
                                                                 int taps, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * Cin) return;
  const int n = i / Cin;
  wq[i] = n < taps ? round_tf32(__ldg(w + i)) : 0.f;
}

struct GatherArgs {
  const float* T;   // [N, H, W, 32]
  int N, H, W, Ho, Wo, kh, kw, sh, sw, pt, pl;
  EpiDev epi;
};

__global__ void __launch_bounds__(256) taps_gather_to_one_kernel(const GatherArgs a) {
  const long npix = (long)a.N * a.Ho * a.Wo;
  for (long pix = (long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long)gridDim.x * blockDim.x) {
    const int ow = (int)(pix % a.Wo);
    const long r = pix / a.Wo;
    const int oh = (int)(r % a.Ho);
    const long img = r / a.Ho;
    float acc = 0.f;
    for (int kh = 0; kh < a.kh; ++kh) {
      const int ih = oh * a.sh - a.pt + kh;
      if (ih < 0 || ih >= a.H) continue;
      for (int kw = 0; kw < a.kw; ++kw) {
        const int iw = ow * a.sw - a.pl + kw;
        if (iw < 0 || iw >= a.W) continue;
        acc += __ldg(a.T + (((size_t)img * a.H + ih) * a.W + iw) * 32 + kh * a.kw + kw);
      }
    }
    epi_store(a.epi, (size_t)pix, 0, acc);
  }
}

// grow-only scratch (T and the packed filter); cannot grow during a graph capture: the eager warm-up pass sizes it
float* to_one_scratch(size_t bytes, cudaStream_t st) {
  static float* buf = nullptr;
  static size_t cap = 0;
  if (bytes <= cap) return buf;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  float* nb = nullptr;
  const size_t want = bytes + (bytes >> 2);
  if (cudaMalloc(&nb, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  buf = nb; cap = want;          // (the outgrown buffer stays allocated: captured graphs may still point at it)
  return buf;
}

}  // namespace

bool conv_to_one_tc_eligible(const advoc_conv_desc* d, const void* x, int ldx, const void* w, const advoc_epilogue* ep) {
  static const bool disabled = getenv("ADVOC_NO_TO_ONE_TC") != nullptr;   // A/B switch
  return !disabled && d->math != ADVOC_MATH_FP32 && d->math != ADVOC_MATH_F16 && d->Cout == 1 && d->kh * d->kw <= 32 &&
         common_eligible(d->Cin, 32, ldx, 0) && aligned16(x) && aligned16(w) && ep && ep->d_out0 && !ep->d_out1 &&
         !ep->d_gate && ep->keep_prob >= 1.f && ep->store_w == 0 && ep->out0_dtype == ADVOC_DT_F32 &&
         ep->out0_row_pad == 0 && (long)d->N * d->H * d->W < 2147483647L / 32;
}

// returns ADVOC_UNSUPPORTED (without an error message) when the scratch cannot be grown right now: the
// caller then takes the CUDA-core kernel
int conv_to_one_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                   void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long pin = (long)d->N * d->H * d->W;
  if (pin == 0 || (long)d->N * d->Ho * d->Wo == 0) return ADVOC_OK;
  const size_t t_floats = (size_t)pin * 32, w_floats = (size_t)32 * d->Cin;
  float* scratch = to_one_scratch((t_floats + w_floats) * sizeof(float), st);
  if (!scratch) return ADVOC_UNSUPPORTED;
  float* T = scratch;
  float* wq = scratch + t_floats;
  pack_to_one_filter_kernel<<<(32 * d->Cin + 255) / 256, 256, 0, st>>>(w, wq, d->kh * d->kw, d->Cin);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  advoc_conv_desc d1 = *d;
  d1.Cout = 32; d1.kh = 1; d1.kw = 1; d1.sh = 1; d1.sw = 1; d1.pad_t = 0; d1.pad_l = 0; d1.Ho = d->H; d1.Wo = d->W;
  advoc_epilogue e1 = {};
  e1.d_out0 = T; e1.ld0 = 32; e1.alpha = 0.f; e1.keep_prob = 1.f;
  e1.gate_scale0 = e1.gate_scale1 = 1.f;
  int rc = conv_fwd_tc(&d1, x, ldx, wq, &e1, stream);
  if (rc) return rc;
  GatherArgs g = {};
  rc = lower_epilogue(ep, d->Ho, d->Wo, 1, &g.epi);
  if (rc) return rc;
  g.T = T; g.N = d->N; g.H = d->H; g.W = d->W; g.Ho = d->Ho; g.Wo = d->Wo; g.kh = d->kh; g.kw = d->kw;
  g.sh = d->sh; g.sw = d->sw; g.pt = d->pad_t; g.pl = d->pad_l;
  const long npix = (long)d->N * d->Ho * d->Wo;
  const long blocks = (npix + 255) / 256;
  taps_gather_to_one_kernel<<<(unsigned)(blocks < 8L * sm_count() ? blocks : 8L * sm_count()), 256, 0, st>>>(g);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc

extern "C" int advoc_debug_flags(unsigned int* out) {
  using namespace advoc;
  ADVOC_REQUIRE(out != nullptr, ADVOC_BAD_ARG, "out is NULL");
  unsigned int* w = debug_word();
  ADVOC_REQUIRE(w != nullptr, ADVOC_CUDA_ERROR, "no debug word");
  ADVOC_CHECK_CUDA(cudaMemcpy(out, w, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  ADVOC_CHECK_CUDA(cudaMemset(w, 0, sizeof(unsigned int)));
  volatile unsigned int* h = tc::debug_host_word();
  if (h) *h = 0;
  return ADVOC_OK;
}

extern "C" int advoc_debug_peek(unsigned int* out) {
  using namespace advoc;
  ADVOC_REQUIRE(out != nullptr, ADVOC_BAD_ARG, "out is NULL");
  volatile unsigned int* h = tc::debug_host_word();
  *out = h ? *h : 0u;
  return ADVOC_OK;
}
