// Inline-PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, tcgen05.mma/ld/commit).
#pragma once
#include "common.cuh"

#include <cuda.h>

namespace advoc {
namespace tc {

constexpr unsigned long long WAIT_TIMEOUT_CYCLES = 2000000000ull;  // ~1 s: a bug must not hang the box

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
// Bounded wait: on timeout raise the debug flag -- in device memory (later launches return at once)
// and in the pinned host word the host side polls at its sync points (tc_host.cu: debug_word) --
// and fall through: the kernel terminates instead of hanging the box, its results are garbage, and
// the host raises RuntimeError at the next advoc_debug_peek / advoc_debug_flags.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, unsigned code) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = clock64();
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      // another CTA already timed out: the launch is lost, drain quickly
      if (dbg && *reinterpret_cast<volatile unsigned int*>(dbg) != 0) return;
      if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
        if (dbg) {
          atomicExch(dbg, code);
          unsigned int* host = *reinterpret_cast<unsigned int* volatile*>(dbg + 2);
          if (host) {
            *reinterpret_cast<volatile unsigned int*>(host) = code;
            __threadfence_system();
          }
        }
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
// tiled-mode 4-D box {c, w, h, n}; out-of-range coordinates (negative included) read as zero
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c, int w, int h,
                                            int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n)
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
// kind::f16 with fp16 A/B (10-bit mantissa like tf32, half the operand bytes, K = 16 per instruction)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
template <bool HALF>
__device__ __forceinline__ void umma_op(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (HALF) umma_f16(tmem_d, da, db, idesc, acc);
  else umma_tf32(tmem_d, da, db, idesc, acc);
}
// instruction descriptor: D fp32; A/B tf32 (kind::tf32) or fp16 (kind::f16), both K-major; M x N tile
template <bool HALF>
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (HALF ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// One lane of a converged warp (elect.sync).  Issuing tcgen05 / TMA instructions under this predicate
// from warp-uniform code lets the compiler keep their operands in uniform registers; a plain
// `if (lane == 0)` region makes it wrap every UTCMMA in an ELECT/BRA.U.ANY loop (~140 cycles each).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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


// ---- host side (tc_host.cu): driver entry points fetched at run time, tensor-map encoders ----
bool tma_ok();                 // cuTensorMapEncode{Tiled,Im2col} available
unsigned int* debug_word();    // device word raised by a timed-out pipeline wait
volatile unsigned int* debug_host_word();   // pinned host mirror of it (nullptr if unavailable)

// im2col-mode map over NHWC activations [Nimg, Hin, Win, ld] exposing C channels; box = box_c
// channels x box_pix pixels, 128-byte swizzle (box_c * 4 must be 128).
// `half` != 0 in every encoder: the tensor holds fp16 elements (ld / strides / box_c stay in ELEMENTS)
int encode_im2col(CUtensorMap* tm, const void* x, int Nimg, int Hin, int Win, int ld, int C, int lower_h,
                  int lower_w, int upper_h, int upper_w, int trav_h, int trav_w, int box_c, int box_pix,
                  bool atom32 = false, int half = 0);
// plain 2-D map: `rows` rows of `inner` floats, row stride in bytes; box = box_inner x box_rows
int encode_tiled2d(CUtensorMap* tm, const void* p, int inner, long rows, size_t row_stride_bytes, int box_inner,
                   int box_rows, bool atom32 = false, int half = 0);
// tiled 4-D map over a strided NHWC view: dims {C, W, H, N} with pixel / row / image strides in
// elements; box {box_c, box_w, box_h, 1}; box rows of 128 bytes use the 128-byte swizzle, rows of 64
// bytes the 64-byte swizzle
int encode_tiled4d(CUtensorMap* tm, const void* x, int C, int W, int H, int Nimg, long stride_w, long stride_h,
                   long stride_n, int box_c, int box_w, int box_h, int half = 0);
// atom32: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks swizzled over 4 rows) -- the only
// layout tcgen05 accepts for MN-major 32-bit (tf32) operands; default is SWIZZLE_128B.

}  // namespace tc
}  // namespace advoc
