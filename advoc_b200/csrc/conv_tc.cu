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

#include <cuda.h>

namespace advoc {

namespace {

constexpr int BM = 128;        // UMMA_M (cta_group::1)
constexpr int BK = 32;         // fp32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;      // kind::tf32
constexpr int MAX_CLASSES = 4;
constexpr int MAX_TAPS = 25;
constexpr int NUM_THREADS = 192;
constexpr unsigned long long WAIT_TIMEOUT_CYCLES = 2000000000ull;  // ~1 s: a bug must not hang the box

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
  int kblocks;         // contraction channels / 32
  EpiDev epi;
  unsigned int* dbg;   // [0] != 0 after a barrier wait timed out
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: on timeout raise the debug flag and fall through (results are then garbage but
// the kernel terminates; tests read the flag).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, unsigned code) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      // another CTA already timed out: the launch is lost, drain quickly
      if (dbg && *reinterpret_cast<volatile unsigned int*>(dbg) != 0) return;
      if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
        if (dbg) atomicExch(dbg, code);
        return;
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// instruction descriptor, kind::tf32: D fp32, A/B tf32, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Four consecutive channels of one stored pixel (the vector form of epi_store).
__device__ __forceinline__ void epi_store4(const EpiDev& e, size_t pix, int n, float4 acc) {
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
    const float s = n < e.gate_split ? e.gscale0 : e.gscale1;  // split is a multiple of 4
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

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 4;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_kernel(const __grid_constant__ TcParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  // CTAs that share an A tile (N tiles, then parity classes) are adjacent in launch order so
  // the activations are fetched from HBM once and re-served from L2
  const int n_ntiles = p.Cn / BN;
  const int n_tile = (int)(blockIdx.x % n_ntiles);
  const int cls = (int)((blockIdx.x / n_ntiles) % p.nclasses);
  const long m_tile = blockIdx.x / (n_ntiles * p.nclasses);
  const int Ah = p.Ah[cls], Aw = p.Aw[cls];
  const long Mc = (long)p.Nimg * Ah * Aw;
  const long m0 = m_tile * BM;
  if (m0 >= Mc) return;  // uniform per CTA, before any barrier or allocation
  if (p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0) return;  // aborted launch
  const int n0 = n_tile * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // operand ring, 1024-byte aligned (SWIZZLE_128B atom)
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[cls]);
    prefetch_tmap(&p.tmB);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
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

  const int iters = p.ntaps[cls] * p.kblocks;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int a_w = (int)(m0 % Aw);
      const long r = m0 / Aw;
      const int a_h = (int)(r % Ah);
      const int img = (int)(r / Ah);
      const int cw = a_w * p.trav_w + p.base_w[cls];
      const int ch = a_h * p.trav_h + p.base_h[cls];
      int stage = 0;
      uint32_t phase = 0;
      int tap = 0, kb = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, p.dbg, 1u);
        uint8_t* a_dst = ring_ptr + stage * L::STAGE_BYTES;
        uint8_t* b_dst = a_dst + L::A_BYTES;
        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
        const unsigned off = p.tap_off[cls][tap];
        tma_load_im2col_4d(&p.tmA[cls], &full_bar[stage], a_dst, kb * BK, cw, ch, img, (uint16_t)(off & 0xFF),
                           (uint16_t)(off >> 8));
        tma_load_2d(&p.tmB, &full_bar[stage], b_dst, kb * BK, (int)p.tap_wrow[cls][tap] * p.Cn + n0);
        if (++kb == p.kblocks) { kb = 0; ++tap; }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[stage], phase, p.dbg, 2u);
        tc_fence_after();
        const uint32_t a_addr = ring + stage * L::STAGE_BYTES;
        const uint64_t da = make_smem_desc(a_addr);
        const uint64_t db = make_smem_desc(a_addr + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 32 bytes (= 2 x 16-byte units) along K inside the swizzle atom
          umma_tf32(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&tmem_full_bar);       // accumulator complete -> epilogue
    }
  } else {
    // ===== epilogue: warp q may only touch TMEM lanes [32q, 32q+32) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const long m = m0 + row;
    const bool valid = m < Mc;
    size_t pix = 0;
    if (valid) {
      const int a_w = (int)(m % Aw);
      const long r = m / Aw;
      const int a_h = (int)(r % Ah);
      const long img = r / Ah;
      pix = ((size_t)img * p.epi.Hs + (size_t)(a_h * p.osh + p.ph[cls])) * p.epi.Ws +
            (size_t)(a_w * p.osw + p.pw[cls]);
    }
    mbar_wait(&tmem_full_bar, 0u, p.dbg, 3u);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          epi_store4(p.epi, pix, n0 + c0 + j,
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

// ---------------------------------------------------------------------------------------------
// host side: driver entry points (no link-time libcuda dependency), tensor maps, launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Driver {
  EncodeTiledFn tiled = nullptr;
  EncodeIm2colFn im2col = nullptr;
  int version = 0;
  bool ok = false;
};

const Driver& driver() {
  static Driver d = [] {
    Driver r;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      r.tiled = reinterpret_cast<EncodeTiledFn>(f);
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      r.im2col = reinterpret_cast<EncodeIm2colFn>(f);
    cudaDriverGetVersion(&r.version);
    cudaGetLastError();
    r.ok = r.tiled && r.im2col;
    return r;
  }();
  return d;
}

unsigned int* debug_word() {
  static unsigned int* w = [] {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, 64) != cudaSuccess) return (unsigned int*)nullptr;
    cudaMemset(p, 0, 64);
    return p;
  }();
  return w;
}

struct ClassGeom {
  int Ah, Aw, ph, pw, lower_h, lower_w, upper_h, upper_w, ntaps;
  unsigned short off[MAX_TAPS], wrow[MAX_TAPS];
};

// im2col map over activations [Nimg, Hin, Win, ld] reading `Ck` channels per pixel
int encode_A(CUtensorMap* tm, const float* x, int Nimg, int Hin, int Win, int ld, int Ck, const ClassGeom& g,
             int trav_h, int trav_w) {
  const Driver& drv = driver();
  cuuint64_t dims[4] = {(cuuint64_t)Ck, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)Win * ld * 4, (cuuint64_t)Hin * Win * ld * 4};
  int lower[2] = {g.lower_w, g.lower_h};
  int upper[2] = {g.upper_w, g.upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)trav_w, (cuuint32_t)trav_h, 1};
  CUresult r = drv.im2col(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, lower,
                          upper, (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADVOC_REQUIRE(r == CUDA_SUCCESS, ADVOC_CUDA_ERROR,
                "cuTensorMapEncodeIm2col failed (%d) dims %d,%d,%d,%d ld %d corners (%d,%d)-(%d,%d)", (int)r, Ck,
                Win, Hin, Nimg, ld, g.lower_w, g.lower_h, g.upper_w, g.upper_h);
  // Known driver issue (<= 13.1): the im2col encoder mis-sets a descriptor bit for tensors smaller
  // than 128 KiB; NVIDIA's own CUTLASS applies the same correction.
  if (drv.version <= 13010 && (size_t)Nimg * Hin * Win * ld * 4 < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return ADVOC_OK;
}

int encode_B(CUtensorMap* tm, const float* w, int rows, int Ck, int bn) {
  const Driver& drv = driver();
  cuuint64_t dims[2] = {(cuuint64_t)Ck, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Ck * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = drv.tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADVOC_REQUIRE(r == CUDA_SUCCESS, ADVOC_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d) rows %d Ck %d", (int)r,
                rows, Ck);
  return ADVOC_OK;
}

template <int BN, int STAGES>
int launch(const TcParams& p, int nclasses, long max_tiles, cudaStream_t st) {
  constexpr int smem = STAGES * SmemLayout<BN>::STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem));
    configured = true;
  }
  const long ctas = max_tiles * (p.Cn / BN) * nclasses;
  ADVOC_REQUIRE(ctas < 2147483647L, ADVOC_BAD_SHAPE, "too many output tiles");
  dim3 grid((unsigned)ctas, 1, 1);
  conv_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, smem, st>>>(p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

int pick_bn(int Cn) {
  if (Cn % 256 == 0) return 256;
  if (Cn % 128 == 0) return 128;
  if (Cn % 64 == 0) return 64;
  return 32;
}

int run(TcParams& p, int nclasses, const ClassGeom* g, void* stream) {
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
  switch (pick_bn(p.Cn)) {
    case 256: return launch<256, 2>(p, nclasses, max_tiles, st);
    case 128: return launch<128, 3>(p, nclasses, max_tiles, st);
    case 64: return launch<64, 4>(p, nclasses, max_tiles, st);
    default: return launch<32, 4>(p, nclasses, max_tiles, st);
  }
}

bool common_eligible(int Ck, int Cn, int ldx) {
  return driver().ok && device_arch() == 100 && Ck % BK == 0 && Cn % 32 == 0 && ldx % 4 == 0;
}

bool epilogue_vector_ok(const advoc_epilogue* ep) {
  if (!ep) return true;
  auto ok = [](const float* p, int ld, int co) { return aligned16(p) && ld % 4 == 0 && co % 4 == 0; };
  if (!ok(ep->d_out0, ep->ld0, ep->c_off0)) return false;
  if (ep->d_out1 && !ok(ep->d_out1, ep->ld1, ep->c_off1)) return false;
  if (ep->d_bias && !aligned16(ep->d_bias)) return false;
  if (ep->d_gate && !(ok(ep->d_gate, ep->ld_gate, ep->c_off_gate) && ep->gate_split % 4 == 0)) return false;
  return true;
}

}  // namespace

bool conv_fwd_tc_eligible(const advoc_conv_desc* d, int ldx) {
  return common_eligible(d->Cin, d->Cout, ldx) && d->kh * d->kw <= MAX_TAPS && d->pad_t <= 127 && d->pad_l <= 127;
}

bool conv_transposed_tc_eligible(const advoc_conv_desc* d, int ldx) {
  return common_eligible(d->Cout, d->Cin, ldx) && d->sh * d->sw <= MAX_CLASSES && d->kh * d->kw <= MAX_TAPS;
}

// y = conv(x, w): w packed K-major [tap][Cout][Cin]
int conv_fwd_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                void* stream) {
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
  st = encode_A(&p.tmA[0], x, d->N, d->H, d->W, ldx, d->Cin, g, d->sh, d->sw);
  if (st) return st;
  st = encode_B(&p.tmB, w, g.ntaps * d->Cout, d->Cin, pick_bn(d->Cout));
  if (st) return st;
  p.Nimg = d->N; p.trav_h = d->sh; p.trav_w = d->sw; p.osh = 1; p.osw = 1;
  p.Cn = d->Cout; p.kblocks = d->Cin / BK;
  return run(p, 1, &g, stream);
}

// y = conv_transpose(x, w): x [N,Ho,Wo,Cout] (small side), y [N,H,Ws,Cin]; w K-major [tap][Cin][Cout]
// (= the TF conv2d_transpose layout HWOI, and = HWIO of a conv whose input gradient this is)
int conv_transposed_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                       const advoc_epilogue* ep, void* stream) {
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
      st = encode_A(&p.tmA[nc], x, d->N, d->Ho, d->Wo, ldx, d->Cout, c, 1, 1);
      if (st) return st;
      ++nc;
    }
  }
  st = encode_B(&p.tmB, w, d->kh * d->kw * d->Cin, d->Cout, pick_bn(d->Cin));
  if (st) return st;
  p.Nimg = d->N; p.trav_h = 1; p.trav_w = 1; p.osh = d->sh; p.osw = d->sw;
  p.Cn = d->Cin; p.kblocks = d->Cout / BK;
  return run(p, nc, g, stream);
}

}  // namespace advoc

extern "C" int advoc_debug_flags(unsigned int* out) {
  using namespace advoc;
  ADVOC_REQUIRE(out != nullptr, ADVOC_BAD_ARG, "out is NULL");
  unsigned int* w = debug_word();
  ADVOC_REQUIRE(w != nullptr, ADVOC_CUDA_ERROR, "no debug word");
  ADVOC_CHECK_CUDA(cudaMemcpy(out, w, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  ADVOC_CHECK_CUDA(cudaMemset(w, 0, sizeof(unsigned int)));
  return ADVOC_OK;
}
