// Persistent tcgen05 convolution / transposed convolution over 2-D shared-memory PATCHES (sm_100a,
// TF32 operands, fp32 accumulators in TMEM).
//
// conv_tc.cu re-fetches the im2col operand from L2 once per filter tap (16x for k4) and runs one
// output tile per CTA with no overlap between tiles; ncu shows it bound by neither DRAM, L2 nor the
// tensor pipe (profiles/r01b_*).  Here:
//   * the CTA is persistent (grid = #SMs, tiles strided over CTAs); TMA producers, the MMA issuer and
//     the epilogue warps run decoupled through mbarrier rings, and the accumulator is double-buffered
//     in TMEM whenever it fits, so tile i's epilogue overlaps tile i+1's main loop;
//   * the activation operand of a tile is ONE tiled-mode TMA box per 32-channel block and input
//     "plane": a (th + halo) x (tw + halo) patch of pixels, zero-filled outside the image by TMA.
//     GEMM row m = r*PW + c is patch position (r, c); every filter tap is the same patch seen through
//     a UMMA descriptor whose start address is advanced by (dr*PW + dc) rows of 128 bytes (the 128B
//     swizzle is a function of the absolute smem address, so a row-shifted view de-swizzles
//     correctly; probed on B200 by csrc/selftest.cu);
//   * stride-s convolution: the input is split into s*s parity planes (one strided tensor map each),
//     tap (kh, kw) reads plane ((kh-pad) mod s, (kw-pad) mod s) at unit shift floor((k-pad)/s);
//   * stride-s transposed convolution: the s*s output parity classes of a position share the patch
//     and accumulate side by side in TMEM (class z in columns [z*BN, z*BN+BN)); each class sees a
//     dense sub-filter, no MAC is spent on inserted zeros.
// Patch traffic is ~1.3-1.6x the activation bytes instead of 16x; filter tiles stream through their
// own ring, or stay resident in smem when the whole filter of the layer fits.
//
// replaces: tf.layers.conv2d / conv2d_transpose (models/advoc/advoc_model.py:27-32,46-51,65-69) and,
// in the backward pass, their input gradients.
#include "epilogue.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

#include <mutex>
#include <vector>

namespace advoc {

int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* out);

namespace {

using namespace tc;

constexpr int QM = 128;         // UMMA_M: patch positions per tile
inline int qk_of(int half) { return half ? 64 : 32; }   // channels per k-block (one 128-byte swizzle row)
constexpr int Q_MAXT = 36;      // filter taps
constexpr int Q_MAXP = 4;       // input planes
constexpr int Q_MAXC = 4;       // output parity classes
constexpr int Q_MAXA = 8;       // patch ring depth
constexpr int Q_MAXB = 32;      // filter-tile ring depth
constexpr int Q_THREADS = 224;  // warp 0 patch producer, 1 MMA issuer, 2 filter producer, 3-6 epilogue
constexpr size_t Q_SMEM_MAX = 216 * 1024;      // operand rings + output staging, one CTA per SM
constexpr uint32_t Q_STAGE_BYTES = QM * 128;   // one staged output chunk: <=128 pixels x 32 channels

struct alignas(64) P2dParams {
  CUtensorMap tmA[Q_MAXP];  // tiled 4-D {c, w, h, n}, box {32, PW, PH, 1}
  CUtensorMap tmB;          // filter [taps * Cn][Ck], box {32, BN}
  CUtensorMap tmO[2][Q_MAXC];  // output stores per destination and parity class, box {32, tw, th, 1}
  CUtensorMap tmG[Q_MAXC];     // backward pass: the activation whose derivative gates the output, same boxes
  int tma_store, n_out, stage_bufs;
  int out_half;                // staged stores write fp16 (both destinations)
  int use_gate, reduce_add;    // gate tiles come in by TMA; accumulate = TMA reduce-add store
  uint32_t stage_off, gate_off;
  int nplanes, ntaps, ncls;
  int plane_tap0[Q_MAXP + 1];  // taps are sorted by plane: plane p owns [plane_tap0[p], plane_tap0[p+1])
  int plane_oh[Q_MAXP], plane_ow[Q_MAXP];  // patch origin relative to the tile origin (plane coordinates)
  unsigned char tap_cls[Q_MAXT];
  unsigned short tap_wrow[Q_MAXT];   // row block of the filter matrix
  unsigned short tap_shift[Q_MAXT];  // rows from the patch start
  // lean issue loop (lean_issue): per tap, bits 0-15 = descriptor-field offset of the tap's window into the
  // patch (tap_shift * 128 >> 4), bits 16-30 = TMEM column offset of its class (cls * BN), bit 31 = first
  // tap of its class in issue order (that MMA overwrites the accumulator in the first K block)
  unsigned int tap_word[Q_MAXT + 1];   // + 1: the loop reads one entry ahead
  int cls_ph[Q_MAXC], cls_pw[Q_MAXC];
  int PW, PH, th, tw;          // patch box and useful tile extent (positions)
  int tiles_h, tiles_w, n_ntiles;
  long total_tiles;
  int Hc, Wc;                  // position grid per image
  int osh, osw;                // output pixel = position * os + class offset
  int Cn, kblocks;
  uint32_t a_slot_bytes, a_box_bytes, tmem_cols;
  int a_stages, b_stages, acc_bufs, b_resident;
  int G;                       // filter taps per ring slot (one barrier pair per group; divides every plane's tap count)
  EpiDev epi;
  unsigned int* dbg;
  unsigned long long* prof;   // optional [gridDim.x][16] cycle counters (ADVOC_P2D_PROFILE)
};

// profiling wait: like mbar_wait, adds the stalled cycles to `acc` when profiling
__device__ __forceinline__ void mbar_wait_p(uint64_t* bar, uint32_t parity, unsigned int* dbg, unsigned code,
                                            const unsigned long long* prof, unsigned long long& acc) {
  if (prof == nullptr) { mbar_wait(bar, parity, dbg, code); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity, dbg, code);
  acc += (unsigned long long)(clock64() - t0);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c, int w, int h, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* tm, uint32_t src, int c, int w, int h, int n) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c), "r"(w), "r"(h), "r"(n)
               : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One staged chunk: 32 accumulator columns of this thread's pixel -> bias, dropout, activation(s),
// TF32 rounding -> swizzled staging row(s).  kLin: both activations are none / relu / lrelu.
// Backward-pass chunk: out = acc * act'(gate) * scale, gate row read from its swizzled smem tile.
__device__ __forceinline__ void stage_chunk_gated(const EpiDev& e, const uint32_t (&v)[32], uint32_t dst_row,
                                                  uint32_t gate_row, uint32_t sx, int n) {
  const float neg = e.gate_act == ADVOC_ACT_LRELU ? e.alpha : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t off = ((uint32_t)j ^ sx) << 4;
    const float4 g = ld_shared_v4(gate_row + off);
    const float s = (n + 4 * j) < e.gate_split ? e.gscale0 : e.gscale1;   // split is a multiple of 4
    float y[4] = {__uint_as_float(v[4 * j]) * ((g.x > 0.f ? 1.f : neg) * s),
                  __uint_as_float(v[4 * j + 1]) * ((g.y > 0.f ? 1.f : neg) * s),
                  __uint_as_float(v[4 * j + 2]) * ((g.z > 0.f ? 1.f : neg) * s),
                  __uint_as_float(v[4 * j + 3]) * ((g.w > 0.f ? 1.f : neg) * s)};
    if (e.round) {
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = round_tf32(y[u]);
    }
    st_shared_v4(dst_row + off, y[0], y[1], y[2], y[3]);
  }
}

template <bool kLin, bool kDrop>
__device__ __forceinline__ void stage_chunk(const EpiDev& e, const uint32_t (&v)[32], const float4 (&bias4)[8],
                                            uint32_t dst_row, uint32_t sx, int n_out, size_t idx0, uint64_t seed) {
  const ActLin a0 = act_linear(e.act0, e.alpha), a1 = act_linear(e.act1, e.alpha);
  const float inv_keep = 1.f / e.keep_prob;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x[4] = {__uint_as_float(v[4 * j]) + bias4[j].x, __uint_as_float(v[4 * j + 1]) + bias4[j].y,
                        __uint_as_float(v[4 * j + 2]) + bias4[j].z, __uint_as_float(v[4 * j + 3]) + bias4[j].w};
    float sc[4] = {1.f, 1.f, 1.f, 1.f};
    if (kDrop) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t idx = idx0 + 4 * j + u;
        const bool keep = e.mask ? (__ldg(e.mask + idx) != 0) : dropout_keep(seed, idx, e.keep_prob);
        sc[u] = keep ? inv_keep : 0.f;
      }
    }
    float y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      y[u] = (kLin ? apply_lin(x[u], a0) : apply_act(x[u], e.act0, e.alpha)) * sc[u];
      if (e.round) y[u] = round_tf32(y[u]);
    }
    const uint32_t dst = dst_row + (((uint32_t)j ^ sx) << 4);
    st_shared_v4(dst, y[0], y[1], y[2], y[3]);
    if (n_out == 2) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        y[u] = (kLin ? apply_lin(x[u], a1) : apply_act(x[u], e.act1, e.alpha)) * sc[u];
        if (e.round) y[u] = round_tf32(y[u]);
      }
      st_shared_v4(dst + Q_STAGE_BYTES, y[0], y[1], y[2], y[3]);
    }
  }
}

// fp16 destinations: the same 32 accumulator columns become four 16-byte pieces (8 halves each) at
// pieces piece0 .. piece0 + 3 of the staging row (piece index XOR sx = the store map's swizzle).
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <bool kLin, bool kDrop>
__device__ __forceinline__ void stage_chunk_h(const EpiDev& e, const uint32_t (&v)[32], const float4 (&bias4)[8],
                                              uint32_t dst_row, uint32_t sx, uint32_t piece0, int n_out, size_t idx0,
                                              uint64_t seed) {
  const ActLin a0 = act_linear(e.act0, e.alpha), a1 = act_linear(e.act1, e.alpha);
  const float inv_keep = 1.f / e.keep_prob;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x[8], sc[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 b4 = bias4[2 * j + h];
      x[4 * h] = __uint_as_float(v[8 * j + 4 * h]) + b4.x;
      x[4 * h + 1] = __uint_as_float(v[8 * j + 4 * h + 1]) + b4.y;
      x[4 * h + 2] = __uint_as_float(v[8 * j + 4 * h + 2]) + b4.z;
      x[4 * h + 3] = __uint_as_float(v[8 * j + 4 * h + 3]) + b4.w;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      sc[u] = 1.f;
      if (kDrop) {
        const size_t idx = idx0 + 8 * j + u;
        const bool keep = e.mask ? (__ldg(e.mask + idx) != 0) : dropout_keep(seed, idx, e.keep_prob);
        sc[u] = keep ? inv_keep : 0.f;
      }
    }
    float y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) y[u] = (kLin ? apply_lin(x[u], a0) : apply_act(x[u], e.act0, e.alpha)) * sc[u];
    const uint32_t dst = dst_row + (((piece0 + (uint32_t)j) ^ sx) << 4);
    st_shared_v4u(dst, pack_half2(y[0], y[1]), pack_half2(y[2], y[3]), pack_half2(y[4], y[5]), pack_half2(y[6], y[7]));
    if (n_out == 2) {
#pragma unroll
      for (int u = 0; u < 8; ++u) y[u] = (kLin ? apply_lin(x[u], a1) : apply_act(x[u], e.act1, e.alpha)) * sc[u];
      st_shared_v4u(dst + Q_STAGE_BYTES, pack_half2(y[0], y[1]), pack_half2(y[2], y[3]), pack_half2(y[4], y[5]),
                    pack_half2(y[6], y[7]));
    }
  }
}

struct TileCoord {
  int n_tile, img, h0, w0;
};

__device__ __forceinline__ TileCoord decode_tile(const P2dParams& p, long t) {
  // total_tiles < 2^31 (checked on the host): 32-bit divisions
  TileCoord c;
  unsigned s = (unsigned)t;
  const unsigned nn = (unsigned)p.n_ntiles, tw = (unsigned)p.tiles_w, th = (unsigned)p.tiles_h;
  c.n_tile = (int)(s % nn);
  s /= nn;
  c.w0 = (int)(s % tw) * p.tw;
  s /= tw;
  c.h0 = (int)(s % th) * p.th;
  c.img = (int)(s / th);
  return c;
}

// HALF: fp16 operands (tcgen05 kind::f16, 64 channels per 128-byte patch / filter row) instead of tf32
// (32 channels per row); every byte offset, descriptor and barrier of the pipeline is the same.
template <int BN, bool HALF>
__global__ void __launch_bounds__(Q_THREADS, 2) conv_p2d_kernel(const __grid_constant__ P2dParams p) {
  constexpr int KCH = HALF ? 64 : 32;   // channels per k-block
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[Q_MAXA], a_empty[Q_MAXA];
  __shared__ __align__(8) uint64_t b_full[Q_MAXB], b_empty[Q_MAXB];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], g_full[2];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  const uint32_t b_off = (uint32_t)p.a_stages * p.a_slot_bytes;
  constexpr uint32_t B_BYTES = BN * 128;
  const uint32_t acc_cols = (uint32_t)p.ncls * BN;
  const long ntl = p.total_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); mbar_init(&g_full[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nplanes; ++i) prefetch_tmap(&p.tmA[i]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"(p.tmem_cols)
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
  } else if (warp == 0) {
    // ===== patch producer: warp-uniform loop, TMA issued by one elected lane =====
    int as = 0;
    uint32_t aph = 0;
    unsigned long long w_a = 0;
    const long long t_start = clock64();
    for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
      const TileCoord tc_ = decode_tile(p, t);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        for (int pl = 0; pl < p.nplanes; ++pl) {
          mbar_wait_p(&a_empty[as], aph ^ 1u, p.dbg, 31u, p.prof, w_a);
          __syncwarp();
          if (elect_one()) {
            mbar_expect_tx(&a_full[as], p.a_box_bytes);
            tma_load_4d(&p.tmA[pl], &a_full[as], ring_ptr + (size_t)as * p.a_slot_bytes, kb * KCH,
                        tc_.w0 + p.plane_ow[pl], tc_.h0 + p.plane_oh[pl], tc_.img);
          }
          __syncwarp();
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
    if (p.prof && lane == 0) {
      p.prof[blockIdx.x * 16 + 0] = (unsigned long long)(clock64() - t_start);
      p.prof[blockIdx.x * 16 + 1] = w_a;
    }
  } else if (warp == 2) {
    // ===== filter producer: one ring slot = G consecutive taps (G loads, one barrier) =====
    const int G = p.G;
    const uint32_t slot_bytes = (uint32_t)G * B_BYTES;
    if (p.b_resident) {
      if ((long)blockIdx.x < ntl) {
        // everything once: lane l loads tap tile l, l+32, ... of slot (kb, group)
        const int ngroups = p.ntaps / G;
        for (int s = 0; s < p.kblocks * ngroups; ++s) {
          const int kb = s / ngroups, g = s - kb * ngroups;
          if (lane == 0) mbar_expect_tx(&b_full[s], slot_bytes);
          __syncwarp();
          if (lane < G)
            tma_load_2d(&p.tmB, &b_full[s], ring_ptr + b_off + (size_t)s * slot_bytes + (size_t)lane * B_BYTES, kb * KCH,
                        (int)p.tap_wrow[g * G + lane] * p.Cn);
        }
      }
    } else {
      int bs = 0;
      uint32_t bph = 0;
      unsigned long long w_b = 0;
      const long long t_start = clock64();
      for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
        const int n0 = (int)(t % p.n_ntiles) * BN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          for (int t0 = 0; t0 < p.ntaps; t0 += G) {
            mbar_wait_p(&b_empty[bs], bph ^ 1u, p.dbg, 32u, p.prof, w_b);
            if (lane == 0) mbar_expect_tx(&b_full[bs], slot_bytes);
            __syncwarp();
            if (lane < G)
              tma_load_2d(&p.tmB, &b_full[bs], ring_ptr + b_off + (size_t)bs * slot_bytes + (size_t)lane * B_BYTES,
                          kb * KCH, (int)p.tap_wrow[t0 + lane] * p.Cn + n0);
            __syncwarp();
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
          }
        }
      }
      if (p.prof && lane == 0) {
        p.prof[blockIdx.x * 16 + 2] = (unsigned long long)(clock64() - t_start);
        p.prof[blockIdx.x * 16 + 3] = w_b;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole loop runs inside ONE elect.sync region -- a single thread waits and
    // issues, and because the region is entered through elect.sync the compiler keeps the descriptors
    // in uniform registers (no per-instruction ELECT/BRA.U.ANY wrapper, no warp-wide polling).
    // The loop is the "lean" one of round 1 (profiles/r01c_mma_issue_stall_samples.txt; r02: +5 % on the
    // forward, all parity suites green): descriptors are 64-bit adds onto a per-slot base descriptor
    // (patch rows and filter tiles are 16-byte-granular offsets inside the 14-bit address field, which
    // cannot carry: shared memory ends below 2^18), the per-tap constants come as one host-packed word,
    // and the tap loop is kept rolled so that each tap's few scalar instructions sit between its four
    // MMAs and the next tap's. =====
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc<HALF>(QM, BN);
      const int G = p.G;
      const uint32_t slot_bytes = (uint32_t)G * B_BYTES;
      const int ngroups = p.ntaps / G;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      long i = 0;
      unsigned long long w_acc = 0, w_a = 0, w_b = 0;
      const long long t_start = clock64();
      for (long t = blockIdx.x; t < ntl; t += gridDim.x, ++i) {
        const int buf = (int)(i % p.acc_bufs);
        const uint32_t use = (uint32_t)(i / p.acc_bufs);
        mbar_wait_p(&acc_empty[buf], (use & 1u) ^ 1u, p.dbg, 33u, p.prof, w_acc);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)buf * acc_cols;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          const uint32_t later_kb = kb != 0 ? 1u : 0u;
          int tp = 0;
          for (int pl = 0; pl < p.nplanes; ++pl) {
            mbar_wait_p(&a_full[as], aph, p.dbg, 34u, p.prof, w_a);
            const uint64_t a_desc = make_smem_desc(ring + (uint32_t)as * p.a_slot_bytes);
            const int t_end = p.plane_tap0[pl + 1];
            for (; tp < t_end; tp += G) {
              uint64_t db;
              if (p.b_resident) {
                const int slot = kb * ngroups + tp / G;
                if (i == 0) mbar_wait(&b_full[slot], 0u, p.dbg, 35u);
                db = make_smem_desc(ring + b_off + (uint32_t)slot * slot_bytes);
              } else {
                mbar_wait_p(&b_full[bs], bph, p.dbg, 35u, p.prof, w_b);
                db = make_smem_desc(ring + b_off + (uint32_t)bs * slot_bytes);
              }
              // no tcgen05.fence here: the operands were written by TMA (async proxy) and their
              // arrival is ordered by the mbarrier; the fence is only needed for the TMEM hand-off above
              uint32_t w = p.tap_word[tp];
#pragma unroll 1
              for (int u = 0; u < G; ++u) {
                const uint32_t w_next = p.tap_word[tp + u + 1];   // constant-bank latency off the critical path
                const uint64_t da = a_desc + (uint64_t)(w & 0xffffu);
                const uint32_t d_tmem = d_base + ((w >> 16) & 0x7fffu);
                const uint32_t acc0 = later_kb | ((w >> 31) ^ 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // four 32-byte K steps per 128-byte row (8 tf32 / 16 fp16 channels each)
                  umma_op<HALF>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc0 | (uint32_t)(k != 0));
                db += (uint64_t)(B_BYTES >> 4);
                w = w_next;
              }
              if (!p.b_resident) {
                umma_commit(&b_empty[bs]);
                if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
              }
            }
            umma_commit(&a_empty[as]);
            if (++as == p.a_stages) { as = 0; aph ^= 1u; }
          }
        }
        umma_commit(&acc_full[buf]);
      }
      if (p.prof) {
        p.prof[blockIdx.x * 16 + 4] = (unsigned long long)(clock64() - t_start);
        p.prof[blockIdx.x * 16 + 5] = w_acc;
        p.prof[blockIdx.x * 16 + 6] = w_a;
        p.prof[blockIdx.x * 16 + 7] = w_b;
        p.prof[blockIdx.x * 16 + 8] = (unsigned long long)i;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes [32*(w%4), +32) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m / p.PW, c = m - r * p.PW;
    const bool in_tile = r < p.th && c < p.tw;
    const EpiDev& e = p.epi;
    const bool issuer = warp == 3 && lane == 0;
    // One staged chunk = CW channels of every position of the tile = rows of `row_bytes` in the dense
    // {tw x th} store box: 32 fp32 channels (128 B), 64 fp16 channels (128 B) or, for BN = 32, 32 fp16
    // channels (64 B).  16-byte piece j of a row lives at piece j ^ sx, the SWIZZLE_128B (sx = row % 8)
    // or SWIZZLE_64B (sx = (row / 2) % 4) pattern the store tensor map expects -- and conflict-free
    // for the 8 rows of a quarter-warp.
    const bool out_half = p.out_half != 0;
    const int CW = out_half ? (BN >= 64 ? 64 : 32) : 32;
    const uint32_t row_bytes = out_half ? (uint32_t)CW * 2u : 128u;
    const uint32_t srow = (uint32_t)(r * p.tw + c);
    const uint32_t s_row_addr = ring + p.stage_off + srow * row_bytes;
    const uint32_t sx = row_bytes == 128u ? (srow & 7u) : ((srow >> 1) & 3u);
    const bool lin_acts = act_is_linear(e.act0) && act_is_linear(e.act1);
    const uint64_t seed = e.keep_prob < 1.f ? epi_seed(e) : 0ull;
    // gate tiles: chunk g of this CTA (tile-major, then class, then 32-channel block) lands in gate
    // buffer g & 1; the issuer keeps the loads two chunks ahead of the math (fp32 backward pass only)
    constexpr int CPC = BN / 32;
    const int cpt = p.ncls * CPC;
    auto gate_load = [&](uint32_t g) {
      const long gi = g / cpt;
      const int rem = (int)(g - gi * cpt);
      const long gt = blockIdx.x + gi * (long)gridDim.x;
      if (gt >= ntl) return;
      const int gz = rem / CPC, gc0 = (rem - gz * CPC) * 32;
      const TileCoord gc = decode_tile(p, gt);
      mbar_expect_tx(&g_full[g & 1u], (uint32_t)p.th * p.tw * 128u);
      tma_load_4d(&p.tmG[gz], &g_full[g & 1u], ring_ptr + p.gate_off + (g & 1u) * Q_STAGE_BYTES, gc.n_tile * BN + gc0,
                  gc.w0, gc.h0, gc.img);
    };
    if (p.use_gate && issuer) { gate_load(0); gate_load(1); }
    uint32_t chunk_ctr = 0;
    long i = 0;
    unsigned long long w_full = 0, w_bar = 0, w_ld = 0, w_math = 0, w_fence = 0, w_issue = 0;
    const long long t_start = clock64();
    for (long t = blockIdx.x; t < ntl; t += gridDim.x, ++i) {
      const int buf = (int)(i % p.acc_bufs);
      const uint32_t use = (uint32_t)(i / p.acc_bufs);
      const TileCoord tc_ = decode_tile(p, t);
      const int a = tc_.h0 + r, b = tc_.w0 + c;
      const bool row_ok = in_tile && a < p.Hc && b < p.Wc;
      const int n0 = tc_.n_tile * BN;
      mbar_wait_p(&acc_full[buf], use & 1u, p.dbg, 36u, p.prof, w_full);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * acc_cols;
#pragma unroll 1
      for (int z = 0; z < p.ncls; ++z) {
        const int oh = a * p.osh + p.cls_ph[z], ow = b * p.osw + p.cls_pw[z];
        const bool valid = row_ok && oh < e.Hs && ow < e.Ws;
        const size_t pix = valid ? ((size_t)tc_.img * e.Hs + oh) * e.Ws + ow : 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CW) {
          const uint32_t sbuf = (p.stage_bufs == 2 ? (chunk_ctr & 1u) : 0u) * (uint32_t)p.n_out * Q_STAGE_BYTES;
#pragma unroll 1
          for (int sub = 0; sub < CW; sub += 32) {
            uint32_t v[32];
            const long long tl0 = p.prof ? clock64() : 0;
            tmem_ld32(t_base + (uint32_t)(z * BN + c0 + sub), v);
            if (!p.tma_store) {
              tmem_ld_wait();
              if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  epi_store_vec4(e, pix, n0 + c0 + sub + j,
                                 make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                             __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
              }
              continue;
            }
            // ---- staged path: registers -> swizzled smem -> one TMA store per destination ----
            float4 bias4[8];
            if (e.bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) bias4[j] = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + c0 + sub) + j);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) bias4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            tmem_ld_wait();
            if (p.prof) w_ld += (unsigned long long)(clock64() - tl0);
            if (sub == 0) {
              const long long tb0 = p.prof ? clock64() : 0;
              if (issuer) {   // the store that last read this staging buffer is done with it
                if (p.stage_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
              }
              epi_bar(1);
              if (p.prof) w_bar += (unsigned long long)(clock64() - tb0);
              if (p.use_gate) mbar_wait(&g_full[chunk_ctr & 1u], (chunk_ctr >> 1) & 1u, p.dbg, 37u);
            }
            const long long tm0 = p.prof ? clock64() : 0;
            if (in_tile) {
              const uint32_t dst_row = s_row_addr + sbuf;
              const bool drop = e.keep_prob < 1.f && valid;
              const size_t idx0 = pix * e.Cout + n0 + c0 + sub;
              if (p.use_gate) {
                stage_chunk_gated(e, v, dst_row, ring + p.gate_off + (chunk_ctr & 1u) * Q_STAGE_BYTES + srow * 128u, sx,
                                  n0 + c0);
              } else if (out_half) {
                const uint32_t piece0 = (uint32_t)sub >> 3;   // 4 x 16-byte pieces per 32 channels
                if (lin_acts) {
                  if (drop) stage_chunk_h<true, true>(e, v, bias4, dst_row, sx, piece0, p.n_out, idx0, seed);
                  else stage_chunk_h<true, false>(e, v, bias4, dst_row, sx, piece0, p.n_out, idx0, seed);
                } else {
                  if (drop) stage_chunk_h<false, true>(e, v, bias4, dst_row, sx, piece0, p.n_out, idx0, seed);
                  else stage_chunk_h<false, false>(e, v, bias4, dst_row, sx, piece0, p.n_out, idx0, seed);
                }
              } else if (lin_acts) {
                if (drop) stage_chunk<true, true>(e, v, bias4, dst_row, sx, p.n_out, idx0, seed);
                else stage_chunk<true, false>(e, v, bias4, dst_row, sx, p.n_out, idx0, seed);
              } else {
                if (drop) stage_chunk<false, true>(e, v, bias4, dst_row, sx, p.n_out, idx0, seed);
                else stage_chunk<false, false>(e, v, bias4, dst_row, sx, p.n_out, idx0, seed);
              }
            }
            if (p.prof) w_math += (unsigned long long)(clock64() - tm0);
          }
          if (!p.tma_store) continue;
          const long long tf0 = p.prof ? clock64() : 0;
          fence_async_smem();
          epi_bar(2);
          const long long ti0 = p.prof ? clock64() : 0;
          if (p.prof) w_fence += (unsigned long long)(ti0 - tf0);
          if (issuer) {
            if (p.reduce_add)
              tma_reduce_add_4d(&p.tmO[0][z], ring + p.stage_off + sbuf, n0 + c0, tc_.w0, tc_.h0, tc_.img);
            else
              tma_store_4d(&p.tmO[0][z], ring + p.stage_off + sbuf, n0 + c0, tc_.w0, tc_.h0, tc_.img);
            if (p.n_out == 2)
              tma_store_4d(&p.tmO[1][z], ring + p.stage_off + sbuf + Q_STAGE_BYTES, n0 + c0, tc_.w0, tc_.h0, tc_.img);
            bulk_commit();
            if (p.use_gate) gate_load(chunk_ctr + 2);   // this chunk's gate buffer was consumed before bar 2
          }
          if (p.prof) w_issue += (unsigned long long)(clock64() - ti0);
          ++chunk_ctr;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (issuer) bulk_wait_all();
    if (p.prof && issuer) {
      p.prof[blockIdx.x * 16 + 9] = (unsigned long long)(clock64() - t_start);
      p.prof[blockIdx.x * 16 + 10] = w_full;
      p.prof[blockIdx.x * 16 + 11] = w_bar;
      p.prof[blockIdx.x * 16 + 12] = w_ld;
      p.prof[blockIdx.x * 16 + 13] = w_math;
      p.prof[blockIdx.x * 16 + 14] = w_fence;
      p.prof[blockIdx.x * 16 + 15] = w_issue;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side: geometry
// ---------------------------------------------------------------------------------------------
struct Plan {
  P2dParams p;
  int bn;
  size_t smem;
  double efficiency;  // useful positions / GEMM rows
  size_t filter_bytes_per_tile;  // filter bytes one output tile streams through shared memory
  int ctas_per_sm;
  // plane views of the contraction-side tensor
  int plane_ph[Q_MAXP], plane_pw[Q_MAXP];
  int plane_sh, plane_sw;
};

struct TapT { int plane, cls, wrow, fh, fw; };

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// transposed == false: y[N,Ho,Wo,Cout] = conv(x[N,H,W,Cin]);  Cn = Cout, Ck = Cin
// transposed == true : y[N,H,Wstored,Cin] = conv_transpose(x[N,Ho,Wo,Cout]);  Cn = Cin, Ck = Cout
bool make_plan_budget(const advoc_conv_desc* d, bool transposed, int Wstored, int n_out, bool tma_store, bool gate,
                      Plan* pl, size_t Q_SMEM_BUDGET, int max_tmem, size_t max_b_slot);
// operand type of a layer: fp16 (64 channels per k-block) when the descriptor says ADVOC_MATH_F16
inline int half_of(const advoc_conv_desc* d) { return d->math == ADVOC_MATH_F16 ? 1 : 0; }

// Two CTAs per SM when the rings fit in half the shared memory and the accumulators in half the
// tensor memory: the second CTA's MMA stream fills the issue gaps of the first (measured: the
// in-kernel MMA issue rate of one CTA is ~2x below the tensor pipe's).  Otherwise one CTA per SM.
bool make_plan_uncached(const advoc_conv_desc* d, bool transposed, int Wstored, int n_out, bool tma_store, bool gate,
                        Plan* pl);

// The geometry half of a launch depends only on the layer shape and the epilogue kind; it is asked
// for at least twice per launch (eligibility, then the launch itself), so keep it.  The caller fills in
// the pointer-dependent half (tensor maps, epilogue) on its own copy.
struct PlanKey {
  advoc_conv_desc d;
  int transposed, Wstored, n_out, tma_store, gate;
};
struct PlanEntry { PlanKey key; Plan plan; bool ok; };

bool make_plan(const advoc_conv_desc* d, bool transposed, int Wstored, int n_out, bool tma_store, bool gate,
               Plan* pl) {
  static std::mutex mu;
  static std::vector<PlanEntry>* cache = new std::vector<PlanEntry>();
  PlanKey key;
  memset(&key, 0, sizeof(key));
  key.d = *d;
  key.d.math = half_of(d) ? ADVOC_MATH_F16 : 0;
  key.transposed = transposed; key.Wstored = Wstored; key.n_out = n_out; key.tma_store = tma_store; key.gate = gate;
  const EpiDev epi = pl->p.epi;   // lowered by the caller before planning; not part of the cached geometry
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const PlanEntry& e : *cache)
      if (memcmp(&e.key, &key, sizeof(key)) == 0) {
        *pl = e.plan;
        pl->p.epi = epi;
        return e.ok;
      }
  }
  const bool ok = make_plan_uncached(d, transposed, Wstored, n_out, tma_store, gate, pl);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache->size() < 4096) cache->push_back(PlanEntry{key, *pl, ok});
  }
  return ok;
}

bool make_plan_uncached(const advoc_conv_desc* d, bool transposed, int Wstored, int n_out, bool tma_store, bool gate,
                        Plan* pl) {
  static const int ctas_per_sm = getenv("ADVOC_P2D_CTAS") ? atoi(getenv("ADVOC_P2D_CTAS")) : 0;   // 0 = choose
  Plan one = *pl, two = *pl;
  const bool ok1 = make_plan_budget(d, transposed, Wstored, n_out, tma_store, gate, &one, 216 * 1024, 512, 32 * 1024);
  const bool ok2 = ctas_per_sm != 1 &&
                   make_plan_budget(d, transposed, Wstored, n_out, tma_store, gate, &two, 106 * 1024, 256, 8 * 1024) &&
                   two.p.total_tiles >= 2L * sm_count();
  bool use2 = ok2;
  if (ok1 && ok2 && ctas_per_sm == 0) {
    // Rounds of tiles each configuration needs; two co-resident CTAs share one SM's pipes, so a round of theirs
    // lasts ~1.7x a single CTA's (measured r02: decoder_3, 640 tiles: 102 us as 3 rounds of pairs, 77 us as
    // 5 single rounds with the larger filter slots; decoder_2 / encoder_2, 2464 tiles: 91 / 68 us paired,
    // 100 / 79 us single).
    const long slots = sm_count();
    const double t1 = (double)((one.p.total_tiles + slots - 1) / slots);
    const double t2 = 1.7 * (double)((two.p.total_tiles + 2 * slots - 1) / (2 * slots));
    use2 = t2 < t1;
  }
  if (use2) {
    two.ctas_per_sm = 2;
    *pl = two;
    return true;
  }
  one.ctas_per_sm = 1;
  *pl = one;
  return ok1;
}

bool make_plan_budget(const advoc_conv_desc* d, bool transposed, int Wstored, int n_out, bool tma_store, bool gate,
                      Plan* pl, size_t Q_SMEM_BUDGET, int max_tmem, size_t max_b_slot) {
  P2dParams& p = pl->p;
  TapT taps[Q_MAXT];
  int nt = 0, ncls = 0, nplanes = 0;
  int smin_h[Q_MAXP], smax_h[Q_MAXP], smin_w[Q_MAXP], smax_w[Q_MAXP];
  bool used[Q_MAXP] = {false, false, false, false};
  int Hc, Wc;
  if (!transposed) {
    if (d->sh * d->sw > Q_MAXP || d->kh * d->kw > Q_MAXT) return false;
    Hc = d->Ho; Wc = d->Wo;
    ncls = 1;
    p.cls_ph[0] = 0; p.cls_pw[0] = 0;
    p.osh = 1; p.osw = 1;
    pl->plane_sh = d->sh; pl->plane_sw = d->sw;
    for (int kh = 0; kh < d->kh; ++kh) {
      const int eh = kh - d->pad_t, ph = ((eh % d->sh) + d->sh) % d->sh, fh = floordiv(eh - ph, d->sh);
      for (int kw = 0; kw < d->kw; ++kw) {
        const int ew = kw - d->pad_l, pw = ((ew % d->sw) + d->sw) % d->sw, fw = floordiv(ew - pw, d->sw);
        taps[nt++] = {ph * d->sw + pw, 0, kh * d->kw + kw, fh, fw};
      }
    }
  } else {
    if (d->sh * d->sw > Q_MAXC || d->kh * d->kw > Q_MAXT) return false;
    Hc = (d->H + d->sh - 1) / d->sh;
    Wc = (Wstored + d->sw - 1) / d->sw;
    p.osh = d->sh; p.osw = d->sw;
    pl->plane_sh = 1; pl->plane_sw = 1;
    for (int ph = 0; ph < d->sh; ++ph)
      for (int pw = 0; pw < d->sw; ++pw) {
        if (ph >= d->H || pw >= Wstored) continue;
        const int z = ncls++;
        p.cls_ph[z] = ph; p.cls_pw[z] = pw;
        for (int kh = 0; kh < d->kh; ++kh) {
          if ((ph + d->pad_t - kh) % d->sh != 0) continue;
          const int dr = (ph + d->pad_t - kh) / d->sh;
          for (int kw = 0; kw < d->kw; ++kw) {
            if ((pw + d->pad_l - kw) % d->sw != 0) continue;
            const int dc = (pw + d->pad_l - kw) / d->sw;
            if (nt >= Q_MAXT) return false;
            taps[nt++] = {0, z, kh * d->kw + kw, dr, dc};
          }
        }
      }
  }
  if (nt == 0 || ncls == 0) return false;
  for (int t = 0; t < nt; ++t) {
    const int q = taps[t].plane;
    if (!used[q]) { used[q] = true; smin_h[q] = smax_h[q] = taps[t].fh; smin_w[q] = smax_w[q] = taps[t].fw; }
    if (taps[t].fh < smin_h[q]) smin_h[q] = taps[t].fh;
    if (taps[t].fh > smax_h[q]) smax_h[q] = taps[t].fh;
    if (taps[t].fw < smin_w[q]) smin_w[q] = taps[t].fw;
    if (taps[t].fw > smax_w[q]) smax_w[q] = taps[t].fw;
  }
  int halo_h = 0, halo_w = 0, remap[Q_MAXP];
  for (int q = 0; q < Q_MAXP; ++q) {
    remap[q] = -1;
    if (!used[q]) continue;
    remap[q] = nplanes;
    pl->plane_ph[nplanes] = q / (transposed ? 1 : d->sw);
    pl->plane_pw[nplanes] = q % (transposed ? 1 : d->sw);
    p.plane_oh[nplanes] = smin_h[q];
    p.plane_ow[nplanes] = smin_w[q];
    if (smax_h[q] - smin_h[q] > halo_h) halo_h = smax_h[q] - smin_h[q];
    if (smax_w[q] - smin_w[q] > halo_w) halo_w = smax_w[q] - smin_w[q];
    ++nplanes;
  }
  // tile shape: fewest tiles, then smallest patch
  long best_tiles = -1;
  int best_tw = 0, best_th = 0;
  for (int tw = 1; tw <= Wc && tw + halo_w <= QM; ++tw) {
    const int PW = tw + halo_w;
    int th = QM / PW;
    if (th > Hc) th = Hc;
    if (th < 1 || PW > 256 || th + halo_h > 256) continue;
    const long tiles = (long)((Hc + th - 1) / th) * ((Wc + tw - 1) / tw);
    if (best_tiles < 0 || tiles < best_tiles ||
        (tiles == best_tiles && (long)(th + halo_h) * PW < (long)(best_th + halo_h) * (best_tw + halo_w))) {
      best_tiles = tiles; best_tw = tw; best_th = th;
    }
  }
  if (best_tiles < 0) return false;
  p.th = best_th; p.tw = best_tw;
  p.PW = best_tw + halo_w; p.PH = best_th + halo_h;
  p.tiles_h = (Hc + p.th - 1) / p.th;
  p.tiles_w = (Wc + p.tw - 1) / p.tw;
  p.Hc = Hc; p.Wc = Wc;
  pl->efficiency = (double)Hc * Wc / ((double)best_tiles * QM);
  // taps sorted by plane
  int k = 0;
  for (int q = 0; q < nplanes; ++q) {
    p.plane_tap0[q] = k;
    for (int t = 0; t < nt; ++t) {
      if (remap[taps[t].plane] != q) continue;
      const int orig = taps[t].plane;
      p.tap_cls[k] = (unsigned char)taps[t].cls;
      p.tap_wrow[k] = (unsigned short)taps[t].wrow;
      p.tap_shift[k] = (unsigned short)((taps[t].fh - smin_h[orig]) * p.PW + (taps[t].fw - smin_w[orig]));
      ++k;
    }
  }
  p.plane_tap0[nplanes] = k;
  p.nplanes = nplanes; p.ntaps = nt; p.ncls = ncls;
  const int Cn = transposed ? d->Cin : d->Cout, Ck = transposed ? d->Cout : d->Cin;
  const int QK = qk_of(half_of(d));
  if (Ck % QK != 0 || Cn % 32 != 0) return false;
  int bn = Cn % 256 == 0 ? 256 : (Cn % 128 == 0 ? 128 : (Cn % 64 == 0 ? 64 : 32));
  while (ncls * bn > max_tmem) bn >>= 1;
  {  // experiment knob: narrower N tiles = more tiles and room for a second accumulator (default: no cap)
    static const int max_bn = getenv("ADVOC_P2D_MAX_BN") ? atoi(getenv("ADVOC_P2D_MAX_BN")) : 256;
    while (bn > max_bn && bn > 32) bn >>= 1;
  }
  if (bn < 32) return false;
  pl->bn = bn;
  {
    unsigned seen = 0;
    for (int t = 0; t < nt; ++t) {
      const unsigned cls = p.tap_cls[t];
      const unsigned first = (seen >> cls) & 1u ? 0u : 1u;
      seen |= 1u << cls;
      p.tap_word[t] = ((unsigned)p.tap_shift[t] * 8u) | ((cls * (unsigned)bn) << 16) | (first << 31);
    }
    p.tap_word[nt] = 0;
  }
  p.Cn = Cn; p.kblocks = Ck / QK;
  p.n_ntiles = Cn / bn;
  p.total_tiles = (long)d->N * best_tiles * p.n_ntiles;
  p.acc_bufs = (2 * ncls * bn <= max_tmem) ? 2 : 1;
  uint32_t cols = (uint32_t)(p.acc_bufs * ncls * bn), pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  // shared memory: patch slots hold the box and every 128-row shifted window into it
  const int max_shift = halo_h * p.PW + halo_w;
  int rows = p.PH * p.PW;
  if (max_shift + QM > rows) rows = max_shift + QM;
  p.a_box_bytes = (uint32_t)p.PH * p.PW * 128u;
  p.a_slot_bytes = ((uint32_t)rows * 128u + 1023u) & ~1023u;
  // filter ring slot = G taps (<= 32 KB), G divides every plane's tap count
  int G = 1;
  for (int g = 2; g <= 8 && (size_t)g * bn * 128 <= max_b_slot; g *= 2) {
    bool ok = true;
    for (int q = 0; q < nplanes; ++q) ok = ok && ((p.plane_tap0[q + 1] - p.plane_tap0[q]) % g == 0);
    if (ok) G = g;
  }
  p.G = G;
  pl->filter_bytes_per_tile = (size_t)nt * p.kblocks * bn * 128;
  const size_t b_tile = (size_t)G * bn * 128;
  const int nslots_all = nt / G * p.kblocks;
  const size_t all_b = (size_t)nslots_all * b_tile;
  p.tma_store = tma_store ? 1 : 0;
  p.n_out = n_out;
  const size_t gate_bytes = (tma_store && gate) ? (size_t)2 * Q_STAGE_BYTES : 0;
  const size_t stage1 = (tma_store ? (size_t)n_out * Q_STAGE_BYTES : 0) + gate_bytes;
  int a_need = nplanes + 1 > 3 ? nplanes + 1 : 3;   // patches in flight cover the TMA latency
  while (a_need > 2 && (size_t)a_need * p.a_slot_bytes + 4 * b_tile + stage1 > Q_SMEM_BUDGET) --a_need;
  if ((size_t)a_need * p.a_slot_bytes + 3 * b_tile + stage1 > Q_SMEM_BUDGET) return false;
  size_t left = Q_SMEM_BUDGET - stage1 - (size_t)a_need * p.a_slot_bytes;
  p.a_stages = a_need;
  if (p.n_ntiles == 1 && nslots_all <= Q_MAXB && all_b <= left) {
    p.b_resident = 1;
    p.b_stages = nslots_all;
    left -= all_b;
  } else {
    p.b_resident = 0;
    long bs = (long)(left / b_tile);
    const long want = (long)((96 * 1024) / b_tile) > 4 ? (long)((96 * 1024) / b_tile) : 4;   // ~96 KB of filter tiles in flight
    if (bs > want) bs = want;
    if (bs > Q_MAXB) bs = Q_MAXB;
    p.b_stages = (int)bs;
    left -= (size_t)bs * b_tile;
  }
  p.stage_bufs = 1;
  const size_t stage_only = stage1 - gate_bytes;
  if (tma_store && left >= stage_only) { p.stage_bufs = 2; left -= stage_only; }
  while (p.a_stages < Q_MAXA && left >= p.a_slot_bytes) { ++p.a_stages; left -= p.a_slot_bytes; }
  const size_t staging = stage_only * p.stage_bufs + gate_bytes;
  p.stage_off = (uint32_t)((size_t)p.a_stages * p.a_slot_bytes + (size_t)p.b_stages * b_tile);
  p.gate_off = p.stage_off + (uint32_t)(stage_only * p.stage_bufs);
  p.use_gate = gate_bytes ? 1 : 0;
  pl->smem = (size_t)p.stage_off + staging + 1024;
  if (max_tmem > 256 && pl->smem < 120 * 1024) pl->smem = 120 * 1024;  // one CTA per SM when it may own > half of TMEM
  return p.total_tiles > 0 && p.total_tiles < 2147483647L;
}

template <int BN>
int launch_p2d(const Plan& pl, int half, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_p2d_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(Q_SMEM_MAX + 1024)));
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_p2d_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(Q_SMEM_MAX + 1024)));
    configured = true;
  }
  const long slots = (long)sm_count() * pl.ctas_per_sm;
  long ctas = pl.p.total_tiles < slots ? pl.p.total_tiles : slots;
  if (half)
    conv_p2d_kernel<BN, true><<<(unsigned)ctas, Q_THREADS, pl.smem, st>>>(pl.p);
  else
    conv_p2d_kernel<BN, false><<<(unsigned)ctas, Q_THREADS, pl.smem, st>>>(pl.p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

unsigned long long* prof_buffer() {
  static unsigned long long* buf = [] {
    unsigned long long* b = nullptr;
    if (getenv("ADVOC_P2D_PROFILE") == nullptr) return b;
    if (cudaMalloc(&b, 512 * 16 * sizeof(unsigned long long)) != cudaSuccess) return (unsigned long long*)nullptr;
    cudaMemset(b, 0, 512 * 16 * sizeof(unsigned long long));
    return b;
  }();
  return buf;
}

double min_efficiency() {
  static const double v = [] {
    const char* s = getenv("ADVOC_P2D_MIN_EFF");
    return s ? atof(s) : 0.45;
  }();
  return v;
}

}  // namespace

bool conv_p2d_eligible(const advoc_conv_desc* d, int ldx, int transposed, int store_w) {
  static const bool disabled = getenv("ADVOC_NO_P2D") != nullptr;  // A/B switch for benchmarking
  if (disabled || !(tc::tma_ok() && device_arch() == 100 && ldx % (half_of(d) ? 8 : 4) == 0)) return false;
  Plan pl = {};
  if (!make_plan(d, transposed != 0, store_w ? store_w : d->W, 2, true, true, &pl)) return false;
  if (pl.efficiency < min_efficiency()) return false;
  // Measured on B200 (profiles/README.md, r01c): the patch kernel wins where the layer has several
  // tiles per SM and an output tile streams little filter data (it re-reads the filter once per
  // 128-position tile); deep, filter-heavy layers stay on the per-tap kernel of conv_tc.cu.
  static const bool force = getenv("ADVOC_P2D_FORCE") != nullptr;
  if (force) return true;
  // both thresholds can be swept without a rebuild (defaults = the measured rule above)
  static const long max_filter_kb = getenv("ADVOC_P2D_MAX_FILTER_KB") ? atol(getenv("ADVOC_P2D_MAX_FILTER_KB")) : 512;
  static const long min_tiles_per_sm =
      getenv("ADVOC_P2D_MIN_TILES_PER_SM") ? atol(getenv("ADVOC_P2D_MIN_TILES_PER_SM")) : 3;
  // Round 2 (one CTA per SM when that needs fewer tile rounds, fp16 operands): a FORWARD convolution has one
  // output class, so even an N = 256 tile keeps two accumulators in TMEM and the filter stream overlaps
  // the epilogue -- the patch kernel then wins from one tile per SM on, filter-heavy or not (AdVoc-small
  // encoder_4: 46 us against 56 on the per-tap kernel; regular encoder_3: 90 against 204, encoder_4: 96
  // against 106).  A transposed convolution owns 4 x BN columns per tile (single-buffered from BN = 128):
  // there the round-1 rule stands (regular decoder_5: 164 us here against 86 per tap).
  static const long conv_max_filter_kb =
      getenv("ADVOC_P2D_CONV_MAX_FILTER_KB") ? atol(getenv("ADVOC_P2D_CONV_MAX_FILTER_KB")) : 4096;
  if (!transposed)
    return pl.p.total_tiles >= sm_count() && (long)pl.filter_bytes_per_tile <= conv_max_filter_kb * 1024;
  return pl.p.total_tiles >= min_tiles_per_sm * sm_count() &&
         (long)pl.filter_bytes_per_tile <= max_filter_kb * 1024;
}

// Which of the two tcgen05 kernels a call takes.  Until the end of round 2 the patch kernel also ran the forward
// layers conv_p2d_eligible admits; then the per-tap kernel's forward epilogue was rewritten (conv_tc.cu: stall
// samples had shown its four epilogue warps, not the tensor pipe, pacing it) and it now wins or ties on every
// forward layer of both generators (profiles/README.md, r02P: AdVoc-small 0.527 -> 0.466 ms, regular 1.167 -> 1.031 ms
// with the patch kernel off).  The patch kernel keeps the BACKWARD epilogues -- gate tiles by TMA load, skip sums by
// TMA reduce-add, which the per-tap kernel does element-wise (regular train step 1812 samples/s against 1761 without
// it).  ep == nullptr asks about a forward call.  ADVOC_P2D_FORWARD restores the earlier routing for A/B runs.
bool conv_p2d_preferred(const advoc_conv_desc* d, int ldx, int transposed, int store_w, const advoc_epilogue* ep) {
  if (!conv_p2d_eligible(d, ldx, transposed, store_w)) return false;
  static const bool force = getenv("ADVOC_P2D_FORCE") != nullptr;
  static const bool forward_too = getenv("ADVOC_P2D_FORWARD") != nullptr;
  return force || forward_too || (ep != nullptr && (ep->d_gate != nullptr || ep->accumulate != 0));
}

// N tile (template BN) of the patch-kernel launch for this geometry (host-side query)
int conv_p2d_tile_n(const advoc_conv_desc* d, int transposed, int store_w) {
  Plan pl = {};
  if (!make_plan(d, transposed != 0, store_w ? store_w : d->W, 2, true, true, &pl)) return 0;
  return pl.bn;
}

int conv_p2d(const advoc_conv_desc* d, int transposed, const void* x, int ldx, const void* w,
             const advoc_epilogue* ep, void* stream) {
  Plan pl = {};
  const int half = half_of(d);
  const int es = half ? 2 : 4;   // operand element size
  int st = transposed ? lower_epilogue(ep, d->H, d->W, d->Cin, &pl.p.epi)
                      : lower_epilogue(ep, d->Ho, d->Wo, d->Cout, &pl.p.epi);
  if (st) return st;
  if (!transposed)
    ADVOC_REQUIRE(pl.p.epi.Ws == d->Wo, ADVOC_UNSUPPORTED, "store_w crop is only supported on conv_transpose");
  static const bool no_tma_store = getenv("ADVOC_P2D_DIRECT_STORE") != nullptr;   // A/B switch
  // backward-pass epilogues: the gate tile arrives by TMA and the skip-connection sum is a TMA
  // reduce-add; they carry no bias, activation, dropout or second output
  const bool bwd = ep->d_gate || ep->accumulate;
  const bool bwd_ok = !ep->d_bias && ep->act0 == ADVOC_ACT_NONE && ep->keep_prob >= 1.f && !ep->d_out1;
  // the staged stores write one element type: mixed fp32 / fp16 destinations take the direct-store path
  const bool out_half = pl.p.epi.h0 != 0;
  const bool same_dtype = !ep->d_out1 || (pl.p.epi.h0 == pl.p.epi.h1);
  const bool tma_store = !no_tma_store && (!bwd || bwd_ok) && same_dtype;
  ADVOC_REQUIRE(make_plan(d, transposed != 0, pl.p.epi.Ws, ep->d_out1 ? 2 : 1, tma_store, tma_store && ep->d_gate,
                          &pl),
                ADVOC_UNSUPPORTED, "layer does not fit the patch kernel");
  P2dParams& p = pl.p;
  p.reduce_add = (tma_store && ep->accumulate) ? 1 : 0;
  p.out_half = (tma_store && out_half) ? 1 : 0;
  if (tma_store) {
    const EpiDev& e = p.epi;
    const int ocw = out_half ? (pl.bn >= 64 ? 64 : 32) : 32;   // channels per staged chunk (kernel: CW)
    const int oes = out_half ? 2 : 4;
    if (p.use_gate) {
      for (int z = 0; z < p.ncls; ++z) {
        const int ph = p.cls_ph[z], pw = p.cls_pw[z];
        const int Hz = (e.Hs - ph + p.osh - 1) / p.osh, Wz = (e.Ws - pw + p.osw - 1) / p.osw;
        st = encode_tiled4d(&p.tmG[z], e.gate + e.coffg + ((size_t)ph * e.Ws + pw) * e.ldg, p.Cn, Wz, Hz, d->N,
                            (long)p.osw * e.ldg, (long)p.osh * e.Ws * e.ldg, (long)e.Hs * e.Ws * e.ldg, 32, p.tw, p.th);
        if (st) return st;
      }
    }
    for (int o = 0; o < p.n_out; ++o) {
      const long ld = o == 0 ? e.ld0 : e.ld1;
      const char* base = reinterpret_cast<const char*>(o == 0 ? e.out0 : e.out1) + (size_t)(o == 0 ? e.coff0 : e.coff1) * oes;
      for (int z = 0; z < p.ncls; ++z) {
        const int ph = p.cls_ph[z], pw = p.cls_pw[z];
        const int Hz = (e.Hs - ph + p.osh - 1) / p.osh, Wz = (e.Ws - pw + p.osw - 1) / p.osw;
        st = encode_tiled4d(&p.tmO[o][z], base + ((size_t)ph * e.Ws + pw) * ld * oes, p.Cn, Wz, Hz, d->N, p.osw * ld,
                            (long)p.osh * e.Ws * ld, (long)e.Hs * e.Ws * ld, ocw, p.tw, p.th, out_half);
        if (st) return st;
      }
    }
  }
  const int Hin = transposed ? d->Ho : d->H, Win = transposed ? d->Wo : d->W;
  const int Ck = transposed ? d->Cout : d->Cin;
  for (int q = 0; q < p.nplanes; ++q) {
    const int ph = pl.plane_ph[q], pw = pl.plane_pw[q], sh = pl.plane_sh, sw = pl.plane_sw;
    const int Hp = Hin > ph ? (Hin - ph + sh - 1) / sh : 0, Wp = Win > pw ? (Win - pw + sw - 1) / sw : 0;
    ADVOC_REQUIRE(Hp > 0 && Wp > 0, ADVOC_UNSUPPORTED, "empty input plane");
    st = encode_tiled4d(&p.tmA[q], reinterpret_cast<const char*>(x) + ((size_t)ph * Win + pw) * ldx * es, Ck, Wp, Hp,
                        d->N, (long)sw * ldx, (long)sh * Win * ldx, (long)Hin * Win * ldx, qk_of(half), p.PW, p.PH, half);
    if (st) return st;
  }
  st = encode_tiled2d(&p.tmB, w, Ck, (long)d->kh * d->kw * p.Cn, (size_t)Ck * es, qk_of(half), pl.bn, false, half);
  if (st) return st;
  p.dbg = debug_word();
  p.prof = prof_buffer();
  static const bool verbose = getenv("ADVOC_P2D_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr,
            "p2d %s N%d %dx%dx%d->%d k%d s%d | tile %dx%d patch %dx%d planes %d taps %d cls %d | bn %d ntiles %d "
            "tiles %ld eff %.2f | A %d x %u B %d x %d %s stage %d x %d smem %zu tmem %u ctas/sm %d\n",
            transposed ? "deconv" : "conv", d->N, Hin, Win, Ck, p.Cn, d->kh, d->sh, p.th, p.tw, p.PH, p.PW, p.nplanes,
            p.ntaps, p.ncls, pl.bn, p.n_ntiles, p.total_tiles, pl.efficiency, p.a_stages, p.a_slot_bytes, p.b_stages,
            p.G * pl.bn * 128, p.b_resident ? "resident" : "ring", p.stage_bufs, p.n_out, pl.smem, p.tmem_cols, pl.ctas_per_sm);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (pl.bn) {
    case 256: return launch_p2d<256>(pl, half, s);
    case 128: return launch_p2d<128>(pl, half, s);
    case 64: return launch_p2d<64>(pl, half, s);
    default: return launch_p2d<32>(pl, half, s);
  }
}

}  // namespace advoc

// Developer hook (ADVOC_P2D_PROFILE=1): copies the per-CTA cycle counters of the last patch-kernel
// launches to the host ([512][16] u64: up to 3 CTAs on each of 148 SMs) and clears them.  Synchronises the device.
extern "C" __attribute__((visibility("default"))) int advoc_p2d_profile_read(unsigned long long* out) {
  using namespace advoc;
  unsigned long long* b = prof_buffer();
  ADVOC_REQUIRE(b != nullptr && out != nullptr, ADVOC_UNSUPPORTED, "profiling is off (set ADVOC_P2D_PROFILE=1)");
  ADVOC_CHECK_CUDA(cudaMemcpy(out, b, 512 * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  ADVOC_CHECK_CUDA(cudaMemset(b, 0, 512 * 16 * sizeof(unsigned long long)));
  return ADVOC_OK;
}
