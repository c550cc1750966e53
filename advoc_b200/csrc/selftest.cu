// Hardware probe (not on any product path): does a K-major SWIZZLE_128B operand descriptor
// whose start address is shifted by a whole number of 128-byte rows (not a multiple of 8) read
// the rows it points at when `base_offset` carries (shift mod 8)?  The answer decides whether a
// convolution can keep ONE activation patch in shared memory and present every filter tap as a
// row-shifted view of it.  D = A[shift : shift+128, :] * I  (M=128, N=32, K=32, tf32).
#include "tc_ptx.cuh"

namespace advoc {
namespace {
using namespace tc;

__global__ void __launch_bounds__(128) desc_shift_probe(float* out, int shift, int mode) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* ptr = raw + (base - smem_u32(raw));
  constexpr int ROWS = 288;
  float* A = reinterpret_cast<float*>(ptr);                    // [ROWS][32] swizzled by absolute address
  float* B = reinterpret_cast<float*>(ptr + ROWS * 128);       // [32][32] identity, swizzled (ROWS*128 % 1024 == 0)
  for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) {
    const int r = i >> 5, k = i & 31;
    const int off = r * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2));
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(A) + off) = (float)((r & 63) * 32 + k);
  }
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int r = i >> 5, k = i & 31;
    const int off = r * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2));
    *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(B) + off) = (r == k) ? 1.f : 0.f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (MMA)
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"(32u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = base + (uint32_t)shift * 128u;
    uint64_t da = make_smem_desc(a_addr);
    if (mode == 1) da |= (uint64_t)(shift & 7) << 49;          // base_offset = (addr >> 7) & 7
    const uint64_t db = make_smem_desc(base + ROWS * 128);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) umma_tf32(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0u, nullptr, 0u);
  tc_fence_after();
  uint32_t v[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}
}  // namespace
}  // namespace advoc

// d_out [128*32] floats.  Returns the rows the MMA actually read as out[m*32+n] = (row & 63)*32 + n.
extern "C" __attribute__((visibility("default"))) int advoc_selftest_desc_shift(float* d_out, int shift, int mode) {
  using namespace advoc;
  ADVOC_REQUIRE(d_out && shift >= 0 && shift <= 150, ADVOC_BAD_ARG, "bad selftest arguments");
  const int smem = 288 * 128 + 32 * 128 + 1024;
  desc_shift_probe<<<1, 128, smem>>>(d_out, shift, mode);
  ADVOC_CHECK_CUDA(cudaGetLastError());
  ADVOC_CHECK_CUDA(cudaDeviceSynchronize());
  return ADVOC_OK;
}
