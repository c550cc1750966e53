// k4 s2 transposed convolution to ONE output channel on the tensor cores (generator decoder_1,
// models/advoc/advoc_model.py:153-158; also the input gradient of the discriminator's layer_1).
//
// The layer is bound by reading its [N, Hs, Ws, Cs] input once (269 MB at B = 32 for AdVoc-small).
// The CUDA-core kernel (conv_direct.cu) reads it with thread-per-pixel 16-byte loads (1.8 TB/s); here
// TMA streams 8 x 16-pixel patches (halo included) into 128B-swizzled shared memory and one
// M = 128, N = 16 tcgen05.mma chain per patch computes the 16 tap dot products of every position,
//     T[pos, tap] = <x[pos, :], w[tap, :]>          (TF32 operands, fp32 accumulate in TMEM),
// which the epilogue warps fold into output pixels without atomics (col2im in shared memory):
//     out[2a+ph, 2b+pw] = sum_{dr,dc} T(a+dr, b+dc)[ph+1-2dr, pw+1-2dc]      for the 6 x 14 interior.
// Persistent CTAs; warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue; the accumulator
// and the T staging buffer are double-buffered so the epilogue of a patch overlaps the next.
#include "epilogue.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace advoc {

int lower_epilogue(const advoc_epilogue* ep, int Hs, int Wfull, int Cout, EpiDev* out);

namespace {

using namespace tc;

unsigned long long* g_one_prof = nullptr;

constexpr int O_PH = 8, O_PW = 16;            // patch (GEMM rows = O_PH * O_PW = 128)
constexpr int O_IH = O_PH - 2, O_IW = O_PW - 2;
constexpr int O_STAGES = 8;                   // most 16 KB patch slots a CTA may own (OneParams::stages are used)
constexpr int O_THREADS = 320;               // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int O_MAXKB = 8;                    // Cs <= 256

struct alignas(64) OneParams {
  CUtensorMap tmA;   // tiled 4-D {c, w, h, n}, box {32, 16, 8, 1}
  CUtensorMap tmB;   // taps [16][Cs], box {32, 16}
  int N, Hs, Ws, kblocks, tiles_h, tiles_w;
  int stages;        // patch ring depth of this launch (<= O_STAGES): fewer slots = more CTAs per SM
  long total_tiles;
  EpiDev epi;
  unsigned int* dbg;
  unsigned long long* prof;   // optional [grid][8] cycle counters (ADVOC_ONE_PROFILE)
};

__device__ __forceinline__ void wait_p(uint64_t* bar, uint32_t parity, unsigned int* dbg, unsigned code,
                                       const unsigned long long* prof, unsigned long long& acc) {
  if (prof == nullptr) { mbar_wait(bar, parity, dbg, code); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity, dbg, code);
  acc += (unsigned long long)(clock64() - t0);
}

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

// KS = filter size (4: AdVoc decoder_1; 5: MelspecGAN upconv_4, models/melspecgan/conv2d.py:139-141);
// the tap axis is padded to NP = 16 / 32 accumulator columns (TMA zero-fills the missing filter rows)
// HALF: fp16 operands (64 channels per 128-byte row, kind::f16) instead of tf32 (32 per row)
template <int KS, bool HALF>
__global__ void __launch_bounds__(O_THREADS, KS == 4 ? 3 : 1) deconv_one_tc_kernel(const __grid_constant__ OneParams p) {
  constexpr int NP = KS == 4 ? 16 : 32;
  constexpr int KCH = HALF ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[O_STAGES], a_empty[O_STAGES];
  __shared__ __align__(8) uint64_t b_full, acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_holder;
  __shared__ float tsm[2][O_PH * O_PW][NP + 1];   // T[pos][tap], +1 pad: conflict-free column access

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  constexpr uint32_t A_BYTES = O_PH * O_PW * 128;   // one k-block of one patch
  constexpr uint32_t B_BYTES = NP * 128;            // one k-block of the taps
  const int nstages = p.stages;
  const uint32_t b_off = (uint32_t)nstages * A_BYTES;
  const long ntl = p.total_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(&b_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_holder)),
                 "r"((uint32_t)(2 * NP))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  const bool aborted = p.dbg && *reinterpret_cast<volatile unsigned int*>(p.dbg) != 0;
  const int per_img = p.tiles_h * p.tiles_w;

  if (aborted) {
  } else if (warp == 0) {
    // ===== TMA producer: the resident tap matrix once, then one box per (patch, k-block) =====
    if ((long)blockIdx.x < ntl) {
      if (elect_one()) {
        mbar_expect_tx(&b_full, (uint32_t)p.kblocks * B_BYTES);
        for (int kb = 0; kb < p.kblocks; ++kb)
          tma_load_2d(&p.tmB, &b_full, ring_ptr + b_off + (size_t)kb * B_BYTES, kb * KCH, 0);
      }
      __syncwarp();
    }
    if (elect_one()) {
      int as = 0;
      uint32_t aph = 0;
      unsigned long long w0 = 0;
      const long long ts = clock64();
      for (long t = blockIdx.x; t < ntl; t += gridDim.x) {
        const int img = (int)((unsigned)t / (unsigned)per_img);   // total_tiles < 2^31 (checked on the host)
        const int r = (int)((unsigned)t - (unsigned)img * (unsigned)per_img);
        const int a0 = (r / p.tiles_w) * O_IH - 1, b0 = (r % p.tiles_w) * O_IW - 1;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          wait_p(&a_empty[as], aph ^ 1u, p.dbg, 41u, p.prof, w0);
          mbar_expect_tx(&a_full[as], A_BYTES);
          tma_load_4d(&p.tmA, &a_full[as], ring_ptr + (size_t)as * A_BYTES, kb * KCH, b0, a0, img);
          if (++as == nstages) { as = 0; aph ^= 1u; }
        }
      }
      if (p.prof) { p.prof[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - ts); p.prof[blockIdx.x * 8 + 1] = w0; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: M = 128 positions, N = NP taps, K = Cs =====
    constexpr uint32_t idesc = umma_idesc<HALF>(128, NP);
    if (elect_one()) {
      int as = 0;
      uint32_t aph = 0;
      long i = 0;
      unsigned long long w1 = 0, w2 = 0;
      const long long ts = clock64();
      if ((long)blockIdx.x < ntl) mbar_wait(&b_full, 0u, p.dbg, 42u);
      for (long t = blockIdx.x; t < ntl; t += gridDim.x, ++i) {
        const int buf = (int)(i & 1);
        const uint32_t use = (uint32_t)(i >> 1);
        wait_p(&acc_empty[buf], (use & 1u) ^ 1u, p.dbg, 43u, p.prof, w1);
        tc_fence_after();
        for (int kb = 0; kb < p.kblocks; ++kb) {
          wait_p(&a_full[as], aph, p.dbg, 44u, p.prof, w2);
          const uint64_t da = make_smem_desc(ring + (uint32_t)as * A_BYTES);
          const uint64_t db = make_smem_desc(ring + b_off + (uint32_t)kb * B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_op<HALF>(tmem_base + (uint32_t)buf * NP, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                          (kb | k) != 0 ? 1u : 0u);
          umma_commit(&a_empty[as]);
          if (kb == p.kblocks - 1) umma_commit(&acc_full[buf]);
          if (++as == nstages) { as = 0; aph ^= 1u; }
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 8 + 2] = (unsigned long long)(clock64() - ts); p.prof[blockIdx.x * 8 + 3] = w1;
        p.prof[blockIdx.x * 8 + 4] = w2; p.prof[blockIdx.x * 8 + 7] = (unsigned long long)i;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM -> T[pos][tap] in smem -> col2im -> output pixels =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tid = (warp - 2) * 32 + lane;    // 0..255
    const int half = (warp - 2) >> 2;          // which half of the tap columns this warp reads
    // the (up to O_NOUT) output pixels of this thread inside a patch are the same for every patch
    constexpr int OWT = 2 * O_IW;
    constexpr int O_NOUT = (2 * O_IH * OWT + 255) / 256;
    int o_ai[O_NOUT], o_bi[O_NOUT], o_ph[O_NOUT], o_pw[O_NOUT];
    bool o_ok[O_NOUT];
#pragma unroll
    for (int j = 0; j < O_NOUT; ++j) {
      const int o = tid + 256 * j;
      o_ok[j] = o < 2 * O_IH * OWT;
      const int orow = o / OWT, ocol = o - orow * OWT;
      o_ai[j] = 1 + (orow >> 1); o_bi[j] = 1 + (ocol >> 1);
      o_ph[j] = orow & 1; o_pw[j] = ocol & 1;
    }
    long i = 0;
    unsigned long long w3 = 0;
    const long long ts = clock64();
    // plain forward store: bias + none / relu / lrelu to one fp32 destination
    const bool plain = act_is_linear(p.epi.act0) && p.epi.keep_prob >= 1.f && !p.epi.gate && !p.epi.accumulate &&
                       !p.epi.out1 && !p.epi.h0 && !p.epi.round && !p.epi.mask;
    const float bias0 = p.epi.bias ? __ldg(p.epi.bias) : 0.f;
    const ActLin act0 = act_linear(p.epi.act0, p.epi.alpha);
    for (long t = blockIdx.x; t < ntl; t += gridDim.x, ++i) {
      const int buf = (int)(i & 1);
      const uint32_t use = (uint32_t)(i >> 1);
      const int img = (int)((unsigned)t / (unsigned)per_img);   // total_tiles < 2^31 (checked on the host)
      const int r = (int)((unsigned)t - (unsigned)img * (unsigned)per_img);
      const int a0 = (r / p.tiles_w) * O_IH - 1, b0 = (r % p.tiles_w) * O_IW - 1;
      wait_p(&acc_full[buf], use & 1u, p.dbg, 45u, p.prof, w3);
      tc_fence_after();
      constexpr int HC = NP / 2;               // columns per warp
      uint32_t v[HC];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NP + half * HC);
      if (HC == 8) tmem_ld8(taddr, v);
      else tmem_ld16(taddr, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive1(&acc_empty[buf]);
#pragma unroll
      for (int k = 0; k < HC; ++k)
        if (half * HC + k < KS * KS) tsm[buf][row][half * HC + k] = __uint_as_float(v[k]);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // T of this patch complete (the buffer of two
                                                      // patches ago was consumed before this point)
#pragma unroll
      for (int j = 0; j < O_NOUT; ++j) {
        if (!o_ok[j]) continue;
        const int ar = a0 + o_ai[j], bc = b0 + o_bi[j];
        const int ow = 2 * bc + o_pw[j];
        if (ar >= p.Hs || bc >= p.Ws || ow >= p.epi.Ws) continue;
        const int ph = o_ph[j], pw = o_pw[j], ai = o_ai[j], bi = o_bi[j];
        float acc = 0.f;
        const int kh0 = (ph + 1) & 1, kw0 = (pw + 1) & 1;   // taps of the right parity only
#pragma unroll
        for (int jh = 0; jh < (KS + 1) / 2; ++jh) {
          const int kh = kh0 + 2 * jh;
          if (kh >= KS) break;
          const int dr = (ph + 1 - kh) >> 1;          // -1, 0 or +1
#pragma unroll
          for (int jw = 0; jw < (KS + 1) / 2; ++jw) {
            const int kw = kw0 + 2 * jw;
            if (kw >= KS) break;
            const int dc = (pw + 1 - kw) >> 1;
            acc += tsm[buf][(ai + dr) * O_PW + (bi + dc)][kh * KS + kw];
          }
        }
        const size_t pix = ((size_t)img * p.epi.Hs + (2 * ar + ph)) * p.epi.Ws + ow;
        if (plain) p.epi.out0[pix * (size_t)p.epi.ld0 + p.epi.coff0] = apply_lin(acc + bias0, act0);   // (the col2im epilogue is instruction-bound: profiles/r02P)
        else epi_store(p.epi, pix, 0, acc);
      }
    }
    if (p.prof && tid == 0) { p.prof[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - ts); p.prof[blockIdx.x * 8 + 6] = w3; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * NP))
                 : "memory");
  }
}

}  // namespace

// geometry test (no epilogue): the same layer family as conv_direct.cu's deconv_to_one
bool deconv_one_tc_geometry(const advoc_conv_desc* d, int ldx) {
  static const bool disabled = getenv("ADVOC_NO_ONE_TC") != nullptr;   // A/B switch for benchmarking
  return !disabled && d->math != ADVOC_MATH_FP32 && tc::tma_ok() && device_arch() == 100 && d->Cin == 1 &&
         (d->kh == 4 || d->kh == 5) && d->kw == d->kh && d->sh == 2 && d->sw == 2 && d->pad_t == 1 && d->pad_l == 1 && d->H == 2 * d->Ho &&
         d->W == 2 * d->Wo &&
         (d->math == ADVOC_MATH_F16 ? (d->Cout % 64 == 0 && d->Cout <= 64 * O_MAXKB && ldx % 8 == 0)
                                    : (d->Cout % 32 == 0 && d->Cout <= 32 * O_MAXKB && ldx % 4 == 0));
}

bool deconv_one_tc_eligible(const advoc_conv_desc* d, const void* x, int ldx, const advoc_epilogue* ep) {
  return deconv_one_tc_geometry(d, ldx) && aligned16(x) && ep->keep_prob >= 1.f && ep->d_out1 == nullptr;
}

// w: [16 taps][Cs] fp32 (the HWOI filter of a conv_transpose to one channel), TF32-rounded by the caller
int deconv_one_tc(const advoc_conv_desc* d, const void* x, int ldx, const void* w, const advoc_epilogue* ep,
                  void* stream) {
  OneParams p = {};
  const int half = d->math == ADVOC_MATH_F16;
  const int kch = half ? 64 : 32;
  int st = lower_epilogue(ep, d->H, d->W, 1, &p.epi);
  if (st) return st;
  ADVOC_REQUIRE(aligned16(w), ADVOC_BAD_ALIGN, "filter must be 16-byte aligned");
  p.N = d->N; p.Hs = d->Ho; p.Ws = d->Wo; p.kblocks = d->Cout / kch;
  p.tiles_h = (p.Hs + O_IH - 1) / O_IH;
  p.tiles_w = (p.Ws + O_IW - 1) / O_IW;
  p.total_tiles = (long)p.N * p.tiles_h * p.tiles_w;
  if (p.total_tiles == 0) return ADVOC_OK;
  ADVOC_REQUIRE(p.total_tiles < 2147483647L, ADVOC_BAD_SHAPE, "too many tiles");
  st = encode_tiled4d(&p.tmA, x, d->Cout, p.Ws, p.Hs, p.N, ldx, (long)p.Ws * ldx, (long)p.Hs * p.Ws * ldx, kch, O_PW,
                      O_PH, half);
  if (st) return st;
  const int np = d->kh == 4 ? 16 : 32;
  st = encode_tiled2d(&p.tmB, w, d->Cout, d->kh * d->kw, (size_t)d->Cout * (half ? 2 : 4), kch, np, false, half);
  if (st) return st;
  p.dbg = debug_word();
  static unsigned long long* prof = [] {
    unsigned long long* b = nullptr;
    if (getenv("ADVOC_ONE_PROFILE") == nullptr) return b;
    if (cudaMalloc(&b, 512 * 8 * sizeof(unsigned long long)) != cudaSuccess) return (unsigned long long*)nullptr;
    cudaMemset(b, 0, 512 * 8 * sizeof(unsigned long long));
    return b;
  }();
  p.prof = prof;
  if (prof) g_one_prof = prof;
  // The layer is latency-bound per CTA (one patch -> TMEM -> col2im -> stores is a serial chain that the
  // two accumulator buffers only half hide: r02 measured 97 us for 134 MB of fp16 input), so run up to
  // three CTAs per SM: a ring of max(3, 2 k-blocks) patch slots each instead of eight.
  const int max_smem = O_STAGES * O_PH * O_PW * 128 + O_MAXKB * 32 * 128 + 1024;
  static bool configured = false;
  if (!configured) {
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(deconv_one_tc_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(deconv_one_tc_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(deconv_one_tc_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  static const int max_ctas = getenv("ADVOC_ONE_CTAS") ? atoi(getenv("ADVOC_ONE_CTAS")) : 3;   // A/B switch
  p.stages = max_ctas <= 1 ? O_STAGES : (2 * p.kblocks > 3 ? (2 * p.kblocks < O_STAGES ? 2 * p.kblocks : O_STAGES) : 3);
  const int b_bytes = p.kblocks * np * 128;
  const int smem = p.stages * O_PH * O_PW * 128 + b_bytes + 1024;
  const int static_smem = 2 * O_PH * O_PW * (np + 1) * 4 + 256;
  int per_sm = (220 * 1024) / (smem + static_smem);
  if (per_sm > max_ctas) per_sm = max_ctas;
  if (per_sm < 1) per_sm = 1;
  const long slots = (long)sm_count() * per_sm;
  const long ctas = p.total_tiles < slots ? p.total_tiles : slots;
  ADVOC_REQUIRE(!half || d->kh == 4, ADVOC_UNSUPPORTED, "fp16 operands: k4 only");
  if (half)
    deconv_one_tc_kernel<4, true><<<(unsigned)ctas, O_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  else if (d->kh == 4)
    deconv_one_tc_kernel<4, false><<<(unsigned)ctas, O_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  else
    deconv_one_tc_kernel<5, false><<<(unsigned)ctas, O_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}

}  // namespace advoc

// Developer hook (ADVOC_ONE_PROFILE=1): per-CTA cycle counters of the last deconv_one_tc launch, [512][8] u64.
extern "C" __attribute__((visibility("default"))) int advoc_one_profile_read(unsigned long long* out) {
  using namespace advoc;
  ADVOC_REQUIRE(g_one_prof != nullptr && out != nullptr, ADVOC_UNSUPPORTED, "profiling is off (set ADVOC_ONE_PROFILE=1)");
  ADVOC_CHECK_CUDA(cudaMemcpy(out, g_one_prof, 512 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  ADVOC_CHECK_CUDA(cudaMemset(g_one_prof, 0, 512 * 8 * sizeof(unsigned long long)));
  return ADVOC_OK;
}
