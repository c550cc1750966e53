// k4 convolution FROM ONE input channel on the tensor cores (generator encoder_1,
// models/advoc/advoc_model.py:91-94: x [N, 256, 513, 1] -> [N, 128, 257, ngf], SAME, stride 2).
//
// The CUDA-core kernel for this layer (conv_direct.cu: one thread per output pixel) is bound by
// instruction issue -- ~1840 warp instructions per 32 pixels, 70 % of the issue slots busy, the FMA pipe
// 28 % (profiles/r02k_ncu_full_generator_forward.csv) -- not by the 150-270 MB it writes.  Here the
// contraction over the 16 taps runs as tcgen05 MMAs (M = 128 pixels, N = Cout, K = 16) and the threads only
// gather the im2col rows and convert / store the results:
//   * producers (4 warps): thread m gathers the 16 taps of pixel m of the tile and writes row m of the
//     K-major, 128B-swizzled A tile in shared memory (no TMA: one fp32 per tap has nothing to box);
//   * the row holds the taps TWICE, as tf32 "hi" parts (chunks 0-3) and "lo" parts x - hi (chunks 4-7), and the
//     filter sits in two resident B tiles [w_hi | w_hi] and [w_lo | 0]: six K = 8 MMAs per tile compute
//     x_hi w_hi + x_lo w_hi + x_hi w_lo, i.e. the fp32 product to ~2^-21 -- encoder_1 stays as exact as the
//     CUDA-core kernel it replaces (the 1e-3 parity budget of the stack has no room for a tf32 first layer:
//     scripts emulation 7.9e-4 -> 8.6e-4 at the regular model's worst activation), and the tensor time of a
//     K = 16 layer is negligible either way;
//   * one MMA-issuing warp, accumulators double-buffered in TMEM, four epilogue warps (TMEM -> registers ->
//     bias, one or two slope activations, fp32 / fp16 conversion -> 64-128 contiguous bytes per pixel).
// Tiles are 8 x 16 output patches (4-D TMA stores clip at the image edge and honour padded rows).  Persistent
// CTAs, two per SM.
//
// The same pipeline serves the other thin ends of the train step (round 2; the CUDA-core kernels they replace
// ran 5-10x above their HBM time, profiles/r02A_launches_train_regular.csv):
//   * MODE 1: conv from TWO input channels (discriminator layer_1, advoc_model.py:184-187): K = 32, the hi and lo
//     parts are two 128-byte A tiles and two B tiles, twelve K = 8 MMAs per tile;
//   * MODE 2: stride-1 TRANSPOSED conv from one channel (input gradient of the PatchGAN head, :196-199):
//     row m gathers x[oh + pt - kh, ow + pl - kw];
//   * output channels beyond 128 run as blockIdx.y chunks of 128 with their own resident filter tiles;
//   * backward epilogue (decoder_1's input gradient = conv from the one-channel d loss / d generated to the
//     128-channel concat gradient; the head's input gradient): the activation-derivative gate tile comes in by
//     TMA (double-buffered, one chunk ahead) and the value leaves as  v * gate'(g) * scale[channel < split].
#include "epilogue.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace advoc {

int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* out);

namespace {

using namespace tc;

constexpr int I_TH = 8, I_TW = 16;     // output patch of a tile: 8 rows x 16 columns = the 128 GEMM rows
constexpr int I_THREADS = 288;         // warps 0-3 gather producers, warp 4 MMA issuer, warps 5-8 epilogue
constexpr uint32_t I_A_BYTES = 128 * 128;
constexpr uint32_t I_STAGE_BYTES = 128 * 128;   // one staged output chunk: 128 pixels x <= 128 bytes

enum { M_CONV1 = 0, M_CONV2 = 1, M_TRANS1 = 2, M_CONV1_K5 = 3 };

struct alignas(64) OneInParams {
  CUtensorMap tmO[2];   // output stores, box {CW channels, 16, 8, 1}
  CUtensorMap tmG;      // gate loads of the backward epilogue, same box
  const float* x;
  const float* w;   // [16 * CIN][ldw]: HWIO of a conv, [tap][C][1] of a conv to one channel (its input gradient)
  int N, H, W, ldx, Ho, Wo, sh, sw, pt, pl;   // H, W: gathered (input) side; Ho, Wo: produced side
  int ldw;          // floats between two k rows of the filter = all output channels
  int tiles_h, tiles_w, n_out, cw, out_half, has_gate;
  long tiles;
  EpiDev epi;
  unsigned int* dbg;
};

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_4d_(const CUtensorMap* tm, uint32_t src, int c, int w, int h, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}

struct OneTile { int img, oh0, ow0; };
__device__ __forceinline__ OneTile one_tile(const OneInParams& p, long t) {
  OneTile x;
  unsigned s = (unsigned)t;                      // tiles < 2^31 (checked on the host)
  x.ow0 = (int)(s % (unsigned)p.tiles_w) * I_TW;
  s /= (unsigned)p.tiles_w;
  x.oh0 = (int)(s % (unsigned)p.tiles_h) * I_TH;
  x.img = (int)(s / (unsigned)p.tiles_h);
  return x;
}

__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_load_4d_(const CUtensorMap* tm, uint64_t* bar, uint32_t dst, int c, int w, int h,
                                             int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

template <int COUT, int STAGES, int MODE>
// (COUT = 32 from one channel -- AdVoc-small encoder_1 -- fits 72 registers and, with a two-slot ring, 73 KB of shared
//  memory: three CTAs per SM for the latency chain gather -> MMA -> epilogue -> TMA store, profiles/README.md r02P)
__global__ void __launch_bounds__(I_THREADS, (COUT == 32 && MODE == M_CONV1) ? 3 : 2)
    conv_one_in_tc_kernel(const __grid_constant__ OneInParams p) {
  constexpr int CIN = MODE == M_CONV2 ? 2 : 1;
  constexpr int KS = MODE == M_CONV1_K5 ? 5 : 4;            // filter side (5: MelspecGAN conv_0, models/melspecgan/conv2d.py:182-184)
  constexpr int KREAL = KS * KS * CIN;                      // taps x input channels
  constexpr int KV = KREAL <= 16 ? 16 : 32;                 // K of the GEMM (25 -> 32: zero rows / columns)
  constexpr uint32_t A_STAGE = (uint32_t)(KV / 16) * I_A_BYTES;   // K = 16: [hi | lo] in one tile; K = 32: hi tile, lo tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[STAGES], a_empty[STAGES], acc_full[2], acc_empty[2], gate_full[2];
  __shared__ uint32_t tmem_base_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t B_BYTES = COUT * 128;
  const uint32_t b1 = ring + STAGES * A_STAGE, b2 = b1 + B_BYTES;
  const uint32_t stage_base = b2 + B_BYTES;      // one staging buffer per output, 1024-byte aligned
  // backward epilogue: the two buffers hold the gate tiles (double-buffered TMA loads) and every thread overwrites
  // its own gate row with the result, which the TMA store then reads: no separate staging buffer
  const uint32_t gate_base = stage_base;
  constexpr uint32_t TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;
  const int n0 = (int)blockIdx.y * COUT;         // first output channel of this CTA's chunk

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); mbar_init(&gate_full[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&p.tmO[0]);
    if (p.n_out == 2) prefetch_tmap(&p.tmO[1]);
    if (p.has_gate) prefetch_tmap(&p.tmG);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident filter tiles, K-major rows of 128 B.  K = 16: B1[n] = [w_hi(16) | w_hi(16)], B2[n] = [w_lo | 0];
  // K = 32: B1[n] = w_hi(32), B2[n] = w_lo(32)
  for (int i = threadIdx.x; i < COUT * 8; i += I_THREADS) {
    const int n = i >> 3, c = i & 7;           // row, 16-byte chunk (4 K values)
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = KV == 16 ? (((c & 3) << 2) + j) : ((c << 2) + j);
      const float v = k < KREAL ? __ldg(p.w + (size_t)k * p.ldw + n0 + n) : 0.f;
      hi[j] = round_tf32(v);
      lo[j] = round_tf32(v - hi[j]);
    }
    const uint32_t off = (uint32_t)n * 128u + (((uint32_t)c ^ ((uint32_t)n & 7u)) << 4);
    sts_v4(b1 + off, hi[0], hi[1], hi[2], hi[3]);
    if (KV == 32 || c < 4) sts_v4(b2 + off, lo[0], lo[1], lo[2], lo[3]);
    else sts_v4(b2 + off, 0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  const bool aborted = p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0;

  if (aborted) {
  } else if (warp < 4) {
    // ===== gather producers: one output pixel (patch position r, c) per thread and tile; the loads of the
    // NEXT tile are issued before this tile's row is converted and written =====
    const int m = (int)threadIdx.x;
    const int r = m / I_TW, c = m % I_TW;
    int stage = 0;
    uint32_t phase = 0;
    // (the kernel is bound by instruction issue, so the gather keeps one row pointer per filter row, immediate
    // column offsets and precomputed row / column predicates: ~4 instructions per tap)
    const bool unit_ld = p.ldx == 1;      // inference: the lifted magnitude buffer [B, T, 513, 1] (immediate offsets)
    auto gather = [&](long t, float (&v)[KV]) {
#pragma unroll
      for (int j = 0; j < KV; ++j) v[j] = 0.f;
      if (t >= p.tiles) return;
      const OneTile tl = one_tile(p, t);
      const int oh = tl.oh0 + r, ow = tl.ow0 + c;
      if (oh >= p.Ho || ow >= p.Wo) return;
      // conv: tap (kh, kw) reads (oh sh - pt + kh, ow sw - pl + kw); stride-1 transposed: (oh + pt - kh, ow + pl - kw)
      const int ih0 = MODE == M_TRANS1 ? oh + p.pt : oh * p.sh - p.pt;
      const int iw0 = MODE == M_TRANS1 ? ow + p.pl : ow * p.sw - p.pl;
      constexpr int DIR = MODE == M_TRANS1 ? -1 : 1;
      const float* x00 = p.x + ((size_t)tl.img * p.H * p.W + (long)ih0 * p.W + iw0) * p.ldx;
      bool cok[KS];
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) cok[kw] = (unsigned)(iw0 + DIR * kw) < (unsigned)p.W;
      const long rstride = (long)p.W * p.ldx;
#pragma unroll
      for (int kh = 0; kh < KS; ++kh) {
        const bool rok = (unsigned)(ih0 + DIR * kh) < (unsigned)p.H;
        const float* xr = x00 + DIR * kh * rstride;
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          if (rok && cok[kw]) {
            if (CIN == 1) {
              v[kh * KS + kw] = __ldg(unit_ld ? xr + DIR * kw : xr + DIR * kw * p.ldx);
            } else {                        // two channels of a pixel: one 8-byte load (ldx even, x 8-byte aligned)
              const float2 xv = __ldg(reinterpret_cast<const float2*>(xr + kw * p.ldx));
              v[(kh * KS + kw) * CIN] = xv.x;
              v[(kh * KS + kw) * CIN + CIN - 1] = xv.y;
            }
          }
        }
      }
    };
    float v[KV], vn[KV];
    gather(blockIdx.x, v);
    for (long t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      gather(t + gridDim.x, vn);
      mbar_wait(&a_empty[stage], phase ^ 1u, p.dbg, 51u);
      const uint32_t row = ring + (uint32_t)stage * A_STAGE + (uint32_t)m * 128u;
      const uint32_t sx = (uint32_t)m & 7u;
#pragma unroll
      for (int cc = 0; cc < KV / 4; ++cc) {
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // hi = x truncated to tf32 (one LOP3), lo = x - hi exactly; the tensor core truncates lo itself.  (The
          // round-to-nearest split used before cost 4 instructions per cvt.rna -- 128 of the producer's 440 warp
          // instructions per tile, profiles/r02P -- and is no more accurate: both leave ~2^-21 of x.)
          hi[j] = __uint_as_float(__float_as_uint(v[4 * cc + j]) & 0xFFFFE000u);
          lo[j] = v[4 * cc + j] - hi[j];
        }
        if (KV == 16) {
          sts_v4(row + (((uint32_t)cc ^ sx) << 4), hi[0], hi[1], hi[2], hi[3]);
          sts_v4(row + (((uint32_t)(cc + 4) ^ sx) << 4), lo[0], lo[1], lo[2], lo[3]);
        } else {
          sts_v4(row + (((uint32_t)cc ^ sx) << 4), hi[0], hi[1], hi[2], hi[3]);
          sts_v4(row + I_A_BYTES + (((uint32_t)cc ^ sx) << 4), lo[0], lo[1], lo[2], lo[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&a_full[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
#pragma unroll
      for (int j = 0; j < KV; ++j) v[j] = vn[j];
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc<false>(128, COUT);
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, i = 0;
      const uint64_t db1 = make_smem_desc(b1), db2 = make_smem_desc(b2);
      for (long t = blockIdx.x; t < p.tiles; t += gridDim.x, ++i) {
        const uint32_t buf = i & 1u;
        mbar_wait(&acc_empty[buf], ((i >> 1) & 1u) ^ 1u, p.dbg, 52u);
        mbar_wait(&a_full[stage], phase, p.dbg, 53u);
        tc_fence_after();
        const uint32_t d = tmem_base + buf * (uint32_t)COUT;
        const uint64_t da = make_smem_desc(ring + (uint32_t)stage * A_STAGE);
        if (KV == 16) {
#pragma unroll
          for (int k = 0; k < 4; ++k)      // [x_hi | x_lo] . [w_hi | w_hi]
            umma_tf32(d, da + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 2; ++k)      // x_hi . w_lo
            umma_tf32(d, da + (uint64_t)(2 * k), db2 + (uint64_t)(2 * k), idesc, 1u);
        } else {
          const uint64_t dal = make_smem_desc(ring + (uint32_t)stage * A_STAGE + I_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)      // x_hi . w_hi
            umma_tf32(d, da + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)      // x_lo . w_hi
            umma_tf32(d, dal + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)      // x_hi . w_lo
            umma_tf32(d, da + (uint64_t)(2 * k), db2 + (uint64_t)(2 * k), idesc, 1u);
        }
        umma_commit(&a_empty[stage]);
        umma_commit(&acc_full[buf]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes [32 (w % 4), +32).  Rows are staged in swizzled shared
    // memory (piece j of row m at j ^ sx: the pattern of the store tensor map) and leave through one TMA tensor
    // store per destination and chunk: scattered per-thread stores kept the LSU 60 % busy on their own. =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const EpiDev& e = p.epi;
    // none / relu / lrelu as y = max(v, s * v) with s = 1 / 0 / alpha (alpha <= 1): two instructions per value
    const float s0 = e.act0 == ADVOC_ACT_LRELU ? e.alpha : (e.act0 == ADVOC_ACT_RELU ? 0.f : 1.f);
    const float s1 = e.act1 == ADVOC_ACT_LRELU ? e.alpha : (e.act1 == ADVOC_ACT_RELU ? 0.f : 1.f);
    const float gneg = e.gate_act == ADVOC_ACT_LRELU ? e.alpha : 0.f;   // gate'(g) for g <= 0
    const bool has_gate = p.has_gate != 0;
    const bool issuer = warp == 5 && lane == 0;
    const int CW = p.cw;                                   // channels per staged chunk
    const uint32_t row_bytes = (uint32_t)CW * (p.out_half ? 2u : 4u);   // 128 or 64
    const uint32_t srow = stage_base + (uint32_t)m * row_bytes;
    const bool dbl = row_bytes == 64u && !has_gate;
    const uint32_t sx = row_bytes == 128u ? ((uint32_t)m & 7u) : (((uint32_t)m >> 1) & 3u);
    uint32_t i = 0, g = 0;                                 // tile and chunk counters
    if (has_gate && issuer && (long)blockIdx.x < p.tiles) {
      const OneTile tl = one_tile(p, blockIdx.x);
      mbar_expect_tx(&gate_full[0], I_STAGE_BYTES);
      tma_load_4d_(&p.tmG, &gate_full[0], gate_base, n0, tl.ow0, tl.oh0, tl.img);
    }
    for (long t = blockIdx.x; t < p.tiles; t += gridDim.x, ++i) {
      const uint32_t buf = i & 1u;
      const OneTile tl = one_tile(p, t);
      mbar_wait(&acc_full[buf], (i >> 1) & 1u, p.dbg, 54u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += CW, ++g) {
        // staging buffers free again.  64-byte rows (fp16, 32 channels) use half of a 16 KB slot, so the two halves
        // alternate and only the store before the last one must have been read (the epilogue warps otherwise idle
        // for the store's shared-memory read: 17-30 % of their samples in profiles/r02P)
        if (issuer) {
          if (dbl) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        const uint32_t half_off = dbl ? (g & 1u) * (I_STAGE_BYTES / 2) : 0u;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (has_gate) {
          if (issuer) {                                    // the NEXT chunk's gate tile, one chunk ahead
            long tn = t;
            int cn = c0 + CW;
            if (cn >= COUT) { cn = 0; tn = t + gridDim.x; }
            if (tn < p.tiles) {
              const OneTile nt = one_tile(p, tn);
              const uint32_t nb = (g + 1u) & 1u;
              mbar_expect_tx(&gate_full[nb], I_STAGE_BYTES);
              tma_load_4d_(&p.tmG, &gate_full[nb], gate_base + nb * I_STAGE_BYTES, n0 + cn, nt.ow0, nt.oh0, nt.img);
            }
          }
          mbar_wait(&gate_full[g & 1u], (g >> 1) & 1u, p.dbg, 55u);
        }
        const uint32_t grow = gate_base + (g & 1u) * I_STAGE_BYTES + (uint32_t)m * 128u;
#pragma unroll 1
        for (int sub = 0; sub < CW; sub += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)COUT + (uint32_t)(c0 + sub), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float x[8], y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __uint_as_float(v[j + u]);
            if (e.bias) {
              const float4 ba = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + c0 + sub + j));
              const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + c0 + sub + j) + 1);
              x[0] += ba.x; x[1] += ba.y; x[2] += ba.z; x[3] += ba.w;
              x[4] += bb.x; x[5] += bb.y; x[6] += bb.z; x[7] += bb.w;
            }
            if (has_gate) {   // backward: one fp32 output, 128-byte rows (CW = 32)
              const uint32_t piece = (uint32_t)(sub + j) >> 2;
              const float4 ga = lds_v4(grow + ((piece ^ sx) << 4));
              const float4 gb = lds_v4(grow + (((piece + 1u) ^ sx) << 4));
              const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
              const float sc = (n0 + c0 + sub + j) < e.gate_split ? e.gscale0 : e.gscale1;   // split % 8 == 0
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                y[u] = fmaxf(x[u], s0 * x[u]) * ((gv[u] > 0.f ? 1.f : gneg) * sc);
                if (e.round) y[u] = round_tf32(y[u]);
              }
              sts_v4(grow + ((piece ^ sx) << 4), y[0], y[1], y[2], y[3]);
              sts_v4(grow + (((piece + 1u) ^ sx) << 4), y[4], y[5], y[6], y[7]);
              continue;
            }
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              if (o == 1 && p.n_out != 2) break;
#pragma unroll
              for (int u = 0; u < 8; ++u) y[u] = fmaxf(x[u], (o == 0 ? s0 : s1) * x[u]);
              const uint32_t dst = srow + (uint32_t)o * I_STAGE_BYTES + half_off;
              if (p.out_half) {
                const uint32_t piece = (uint32_t)(sub + j) >> 3;     // 8 halves = 16 bytes
                sts_v4u(dst + ((piece ^ sx) << 4), pack_half2(y[0], y[1]), pack_half2(y[2], y[3]),
                        pack_half2(y[4], y[5]), pack_half2(y[6], y[7]));
              } else {
                if (e.round) {
#pragma unroll
                  for (int u = 0; u < 8; ++u) y[u] = round_tf32(y[u]);
                }
                const uint32_t piece = (uint32_t)(sub + j) >> 2;     // 4 floats = 16 bytes
                sts_v4(dst + ((piece ^ sx) << 4), y[0], y[1], y[2], y[3]);
                sts_v4(dst + (((piece + 1u) ^ sx) << 4), y[4], y[5], y[6], y[7]);
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (issuer) {
          tma_store_4d_(&p.tmO[0], has_gate ? gate_base + (g & 1u) * I_STAGE_BYTES : stage_base + half_off, n0 + c0, tl.ow0,
                        tl.oh0, tl.img);
          if (p.n_out == 2)
            tma_store_4d_(&p.tmO[1], stage_base + I_STAGE_BYTES + half_off, n0 + c0, tl.ow0, tl.oh0, tl.img);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&acc_empty[buf]);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

template <int COUT, int STAGES, int MODE>
int launch_one_in(const OneInParams& p, int chunks, cudaStream_t st) {
  constexpr int ATILES = (MODE == M_CONV2 || MODE == M_CONV1_K5) ? 2 : 1;
  constexpr int max_smem = STAGES * ATILES * (int)I_A_BYTES + 2 * COUT * 128 + 2 * (int)I_STAGE_BYTES + 1024;
  const int smem = STAGES * ATILES * (int)I_A_BYTES + 2 * COUT * 128 + (p.has_gate ? 2 : p.n_out) * (int)I_STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_one_in_tc_kernel<COUT, STAGES, MODE>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  const int per_sm = (COUT == 32 && MODE == M_CONV1 && 3 * (smem + 1024) <= 227 * 1024) ? 3 : 2;
  const long slots = ((long)sm_count() * per_sm + chunks - 1) / chunks;
  const long ctas = p.tiles < slots ? p.tiles : slots;
  conv_one_in_tc_kernel<COUT, STAGES, MODE><<<dim3((unsigned)ctas, (unsigned)chunks), I_THREADS, smem, st>>>(p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

bool epilogue_fits(const advoc_epilogue* ep) {
  auto slope = [](int a) { return a == ADVOC_ACT_NONE || a == ADVOC_ACT_LRELU || a == ADVOC_ACT_RELU; };
  auto ok = [](const void* p, int ld, int co) { return aligned16(p) && ld % 8 == 0 && co % 8 == 0; };
  if (!(ep && ep->d_out0 && !ep->accumulate && ep->keep_prob >= 1.f && ep->store_w == 0 && slope(ep->act0) &&
        slope(ep->act1) && ep->alpha >= 0.f && ep->alpha <= 1.f && (!ep->d_bias || aligned16(ep->d_bias)) &&
        ok(ep->d_out0, ep->ld0, ep->c_off0) && (!ep->d_out1 || ok(ep->d_out1, ep->ld1, ep->c_off1)) &&
        (!ep->d_out1 || ep->out0_dtype == ep->out1_dtype)))      // the staged stores write one element type
    return false;
  if (ep->d_gate)   // backward epilogue: one fp32 output, gate tile by TMA
    return !ep->d_out1 && ep->out0_dtype == ADVOC_DT_F32 && ep->out0_row_pad == 0 &&
           ok(ep->d_gate, ep->ld_gate, ep->c_off_gate) && ep->gate_split % 8 == 0 &&
           (ep->gate_act == ADVOC_ACT_LRELU || ep->gate_act == ADVOC_ACT_RELU);
  return true;
}

bool enabled() {
  static const bool disabled = getenv("ADVOC_NO_ONE_IN_TC") != nullptr;   // A/B switch for benchmarking
  return !disabled && tc::tma_ok() && device_arch() == 100;
}

// Hp x Wp x C: the produced tensor; (Hin, Win): the gathered side
int run_one_in(int mode, int Nimg, int Hin, int Win, int Hp, int Wp, int C, int sh, int sw, int pt, int pl,
               const float* x, int ldx, const float* w, const advoc_epilogue* ep, void* stream) {
  OneInParams p = {};
  int st = lower_epilogue(ep, Hp, Wp, C, &p.epi);
  if (st) return st;
  p.x = x; p.w = w; p.N = Nimg; p.H = Hin; p.W = Win; p.ldx = ldx; p.Ho = Hp; p.Wo = Wp;
  p.sh = sh; p.sw = sw; p.pt = pt; p.pl = pl; p.ldw = C;
  p.tiles_h = (Hp + I_TH - 1) / I_TH;
  p.tiles_w = (Wp + I_TW - 1) / I_TW;
  p.tiles = (long)Nimg * p.tiles_h * p.tiles_w;
  if (p.tiles == 0) return ADVOC_OK;
  const EpiDev& e = p.epi;
  const int cout = C > 128 ? 128 : C, chunks = C / cout;
  p.n_out = e.out1 ? 2 : 1;
  p.out_half = e.h0;
  p.has_gate = e.gate != nullptr;
  p.cw = e.h0 ? (cout >= 64 ? 64 : 32) : 32;
  const int es = e.h0 ? 2 : 4;
  for (int o = 0; o < p.n_out; ++o) {
    const long ld = o == 0 ? e.ld0 : e.ld1;
    const long roww = (long)Wp + (o == 0 ? e.row_pad0 : 0);     // pixels per stored row of this destination
    const char* base = reinterpret_cast<const char*>(o == 0 ? e.out0 : e.out1) + (size_t)(o == 0 ? e.coff0 : e.coff1) * es;
    st = tc::encode_tiled4d(&p.tmO[o], base, C, Wp, Hp, Nimg, ld, roww * ld, (long)Hp * roww * ld, p.cw, I_TW, I_TH, e.h0);
    if (st) return st;
  }
  if (p.has_gate) {
    st = tc::encode_tiled4d(&p.tmG, e.gate + e.coffg, C, Wp, Hp, Nimg, e.ldg, (long)Wp * e.ldg, (long)Hp * Wp * e.ldg, 32,
                            I_TW, I_TH, 0);
    if (st) return st;
  }
  p.dbg = tc::debug_word();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (mode == M_CONV1) {
    switch (cout) {
      case 32: return launch_one_in<32, 2, M_CONV1>(p, chunks, s);
      case 64: return launch_one_in<64, 3, M_CONV1>(p, chunks, s);
      default: return launch_one_in<128, 2, M_CONV1>(p, chunks, s);
    }
  }
  if (mode == M_CONV2) {
    switch (cout) {
      case 32: return launch_one_in<32, 2, M_CONV2>(p, chunks, s);
      case 64: return launch_one_in<64, 2, M_CONV2>(p, chunks, s);
      default: return launch_one_in<128, 2, M_CONV2>(p, chunks, s);
    }
  }
  if (mode == M_CONV1_K5) {
    switch (cout) {
      case 32: return launch_one_in<32, 2, M_CONV1_K5>(p, chunks, s);
      case 64: return launch_one_in<64, 2, M_CONV1_K5>(p, chunks, s);
      default: return launch_one_in<128, 2, M_CONV1_K5>(p, chunks, s);
    }
  }
  switch (cout) {
    case 32: return launch_one_in<32, 3, M_TRANS1>(p, chunks, s);
    case 64: return launch_one_in<64, 3, M_TRANS1>(p, chunks, s);
    default: return launch_one_in<128, 2, M_TRANS1>(p, chunks, s);
  }
}

bool channels_fit(int C) { return C == 32 || C == 64 || (C >= 128 && C <= 1024 && C % 128 == 0); }

}  // namespace

// k4 conv from one or two input channels; forward epilogue (bias + none / relu / lrelu on one or two fp32 / fp16
// outputs) or backward epilogue (gate)
bool conv_one_in_tc_eligible(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                             const advoc_epilogue* ep) {
  return enabled() && d->math != ADVOC_MATH_FP32 && d->math != ADVOC_MATH_F16 &&
         ((d->kh == 4 && d->kw == 4 && (d->Cin == 1 || d->Cin == 2)) || (d->kh == 5 && d->kw == 5 && d->Cin == 1)) &&
         channels_fit(d->Cout) && x && w && epilogue_fits(ep) &&
         (d->Cin == 1 || (ldx % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 7u) == 0)) &&
         (long)d->N * ((d->Ho + I_TH - 1) / I_TH) * ((d->Wo + I_TW - 1) / I_TW) < 2147483647L;
}

int conv_one_in_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                   void* stream) {
  return run_one_in(d->kh == 5 ? M_CONV1_K5 : (d->Cin == 2 ? M_CONV2 : M_CONV1), d->N, d->H, d->W, d->Ho, d->Wo, d->Cout, d->sh, d->sw, d->pad_t,
                    d->pad_l, x, ldx, w, ep, stream);
}

// stride-1 k4 transposed conv FROM one channel (desc: Cout == 1): x [N,Ho,Wo,1] -> y [N,H,W,Cin]; w [16][Cin][1]
bool deconv_from_one_tc_eligible(const advoc_conv_desc* d, const float* x, const float* w, const advoc_epilogue* ep) {
  return enabled() && d->math != ADVOC_MATH_FP32 && d->math != ADVOC_MATH_F16 && d->Cout == 1 && d->kh == 4 &&
         d->kw == 4 && d->sh == 1 && d->sw == 1 && channels_fit(d->Cin) && x && w && epilogue_fits(ep) &&
         ep->out0_row_pad == 0 &&
         (long)d->N * ((d->H + I_TH - 1) / I_TH) * ((d->W + I_TW - 1) / I_TW) < 2147483647L;
}

int deconv_from_one_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                       void* stream) {
  return run_one_in(M_TRANS1, d->N, d->Ho, d->Wo, d->H, d->W, d->Cin, 1, 1, d->pad_t, d->pad_l, x, ldx, w, ep, stream);
}

}  // namespace advoc
