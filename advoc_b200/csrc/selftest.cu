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
// Tensor-pipe issue-rate probe: one thread issues `reps` MMAs (M=128, N=n, K=8 tf32 or K=16 f16)
// over uninitialised shared memory and reports cycles from first issue to completion.
//   mode 0: tf32, one accumulator      mode 1: tf32, accumulators alternate between two TMEM regions
//   mode 2: kind::f16 (same bytes per operand row)   mode 3: tf32 + tcgen05.commit after every 4 MMAs
//   mode 4: tf32, A descriptor start shifted by 3 rows (un-aligned view)
//   mode 5: tf32 issued from a warp-uniform loop under elect.sync (instead of `if (thread == 0)`)
//   mode >= 16: bit flags, see the kernel (64 concurrent bulk copies into smem, 128 concurrent
//   st.shared stream, 256 concurrent tcgen05.ld stream)
__global__ void __launch_bounds__(256) mma_rate_probe(unsigned long long* out, int n, int reps, int mode,
                                                      const float* gsrc) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t alloc_cols = (mode >= 16 && ((mode - 16) & 4096)) ? 128u : 512u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                 "r"(alloc_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (mode >= 16 && ((mode - 16) & 512)) {
    // random operand data instead of whatever the shared memory held (zeros after a fresh launch)
    uint32_t h = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
    float* sm = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) {
      h = h * 1664525u + 1013904223u;
      sm[i] = (float)((int)(h >> 8) - (1 << 23)) * (1.0f / (1 << 23));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (mode >= 16) {
    // bit flags under elect.sync issue: 1 shifted A view, 2 four accumulators in rotation, 4 commit after
    // every 4 MMAs, 8 the other three warps poll an mbarrier meanwhile, 32 per-step descriptor rebuild
    const int fl = mode - 16;
    const int issue_warp = (fl & 1024) ? 1 : ((fl & 2048) ? 3 : 0);
    if ((int)(threadIdx.x >> 5) == issue_warp) {
      const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const long long t0 = clock64();
      for (int i = 0; i < reps; i += 4) {
        const int step = i >> 2;
        const uint32_t a_addr = base + ((fl & 1) ? (uint32_t)(1 + (step % 5) * 13) * 128u : 0u);
        const uint64_t da = make_smem_desc(a_addr);
        const uint64_t db = make_smem_desc(base + 64 * 1024 + ((fl & 32) ? (uint32_t)(step & 3) * 8192u : 0u));
        const uint32_t d = tmem + ((fl & 2) ? (uint32_t)(step & 3) * (uint32_t)(n > 128 ? 128 : n) : 0u);
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_tf32, (i | k) ? 1u : 0u);
          if (fl & 4) umma_commit(&bar2);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0u, nullptr, 0u);
      const long long t2 = clock64();
      if ((threadIdx.x & 31) == 0) {
        out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
        out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
      }
    } else if (fl & 8) {
      while (!mbar_try_wait(&bar, 0u)) {}
    } else if ((fl & 64) && threadIdx.x < 64) {
      // warp 1: bulk copies global -> shared (16 KB each, like TMA operand fills) until the MMAs are done
      if (threadIdx.x == 32) {
        uint32_t ph = 0;
        while (!mbar_try_wait(&bar, 0u)) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(16384u) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           base + 96 * 1024),
                       "l"(gsrc + (size_t)blockIdx.x * 4096), "r"(16384u), "r"(smem_u32(&bar2))
                       : "memory");
          while (!mbar_try_wait(&bar2, ph)) {}
          ph ^= 1u;
        }
      }
    } else if ((fl & 128) && threadIdx.x >= 64) {
      // warps 2-3: st.shared.v4 stream (like the epilogue staging)
      const uint32_t a = base + 112 * 1024 + (threadIdx.x - 64) * 16;
      while (!mbar_try_wait(&bar, 0u)) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(a + r * 1024), "f"(1.0f) : "memory");
      }
    } else if ((fl & 256) && threadIdx.x >= 64) {
      // warps 2-3: tcgen05.ld stream from accumulator columns the MMAs do not touch
      uint32_t v[32];
      float sink = 0.f;
      while (!mbar_try_wait(&bar, 0u)) {
        tmem_ld32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 384u, v);
        tmem_ld_wait();
        sink += __uint_as_float(v[0]);
      }
      if (sink == 123.456f) out[0] = 1;
    }
  } else if (mode == 6) {
    // the whole issue loop inside ONE elect.sync region (a single thread also does the waiting)
    if (threadIdx.x < 32) {
      if (elect_one()) {
        const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const long long t0 = clock64();
        for (int i = 0; i < reps; i += 4) {
          const int step = i >> 2;
          const uint64_t da = make_smem_desc(base + (uint32_t)(1 + (step % 5) * 13) * 128u);
          const uint64_t db = make_smem_desc(base + 64 * 1024 + (uint32_t)(step & 3) * 8192u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_tf32, (i | k) ? 1u : 0u);
          umma_commit(&bar2);
          if ((step & 7) == 7) mbar_wait(&bar2, 0u, nullptr, 0u);   // a wait inside the region (already complete)
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0u, nullptr, 0u);
        const long long t2 = clock64();
        out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
        out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
      }
    }
  } else if (mode == 5) {
    if (threadIdx.x < 32) {
      const uint64_t da = make_smem_desc(base);
      const uint64_t db = make_smem_desc(base + 64 * 1024);
      const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint64_t adv = (uint64_t)(2 * (i & 3));
        if (elect_one()) umma_tf32(tmem, da + adv, db + adv, idesc_tf32, i ? 1u : 0u);
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0u, nullptr, 0u);
      const long long t2 = clock64();
      if (threadIdx.x == 0) {
        out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
        out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
      }
    }
  } else if (threadIdx.x == 0) {
    const uint32_t a_addr = base + (mode == 4 ? 3u * 128u : 0u);
    const uint64_t da = make_smem_desc(a_addr);
    const uint64_t db = make_smem_desc(base + 64 * 1024);
    const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_f16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint64_t adv = (uint64_t)(2 * (i & 3));
      const uint32_t d = tmem + ((mode == 1 && (i & 4)) ? 256u : 0u);
      if (mode == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
            "l"(da + adv), "l"(db + adv), "r"(idesc_f16), "r"(i ? 1u : 0u)
            : "memory");
      } else {
        umma_tf32(d, da + adv, db + adv, idesc_tf32, i ? 1u : 0u);
      }
      if (mode == 3 && (i & 3) == 3) umma_commit(&bar2);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0u, nullptr, 0u);
    const long long t2 = clock64();
    out[blockIdx.x * 2 + 0] = (unsigned long long)(t1 - t0);
    out[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(alloc_cols) : "memory");
}
}  // namespace
}  // namespace advoc

// h_out [ctas*2] u64: cycles to issue / to complete `reps` MMAs on each CTA.
extern "C" __attribute__((visibility("default"))) int advoc_selftest_mma_rate(unsigned long long* h_out, int ctas,
                                                                              int n, int reps, int mode) {
  using namespace advoc;
  ADVOC_REQUIRE(h_out && ctas > 0 && ctas <= 1024 && n >= 8 && n <= 256 && n % 8 == 0, ADVOC_BAD_ARG, "bad arguments");
  unsigned long long* d = nullptr;
  ADVOC_CHECK_CUDA(cudaMalloc(&d, ctas * 2 * sizeof(unsigned long long)));
  const int smem = 64 * 1024 + 64 * 1024 + 64 * 1024 + 1024;
  ADVOC_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float* gsrc = nullptr;
  ADVOC_CHECK_CUDA(cudaMalloc(&gsrc, (size_t)ctas * 16384 + 65536));
  ADVOC_CHECK_CUDA(cudaMemset(gsrc, 0, (size_t)ctas * 16384 + 65536));
  const int fl = mode >= 16 ? mode - 16 : 0;
  const int threads = (fl & 8192) ? 224 : 128;
  const int smem_launch = (fl & 16384) ? 216 * 1024 : smem;
  if (fl & 16384)
    ADVOC_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_launch));
  mma_rate_probe<<<ctas, threads, smem_launch>>>(d, n, reps, mode, gsrc);
  ADVOC_CHECK_CUDA(cudaGetLastError());
  ADVOC_CHECK_CUDA(cudaDeviceSynchronize());
  ADVOC_CHECK_CUDA(cudaMemcpy(h_out, d, ctas * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(d);
  cudaFree(gsrc);
  return ADVOC_OK;
}

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

// ---------------------------------------------------------------------------------------------
// TMA operation-rate probe (round 2): how much does one cp.async.bulk.tensor cost as a function of its
// box size?  One elected thread per CTA keeps `depth` 2-D tiled loads of {128 B x rows} in flight from an
// L2-resident source into a ring of smem slots and times `reps` of them.  wgrad_tc (2 KB im2col boxes)
// turned out to be paced by the operation count, not by bytes; this gives the curve.
// ---------------------------------------------------------------------------------------------
namespace advoc {
namespace {
using namespace tc;
struct alignas(64) TmaProbeParams {
  CUtensorMap tm;
  unsigned long long* out;
  int rows, reps, depth, src_rows, lanes;
};
__global__ void __launch_bounds__(64) tma_rate_probe(const __grid_constant__ TmaProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8 * 8];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring_ptr = smem_raw + (ring - smem_u32(smem_raw));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&p.tm);
  }
  __syncthreads();
  // `lanes` threads of warp 0 (1: entered through elect.sync) each run their own stream of loads
  const bool active = p.lanes == 1 ? (threadIdx.x < 32 && elect_one()) : ((int)threadIdx.x < p.lanes);
  if (active) {
    const int me = p.lanes == 1 ? 0 : (int)threadIdx.x;
    const uint32_t bytes = (uint32_t)p.rows * 128u;
    const int nslots = p.depth;
    const int span = p.src_rows / p.rows;             // boxes in this CTA's source window
    const int row0 = (int)blockIdx.x * p.src_rows;
    uint64_t* mybar = bars + me * 8;
    uint8_t* myring = ring_ptr + (size_t)me * nslots * bytes;
    if (p.depth == 1) {
      // pure issue cost: 64 loads into distinct slots against ONE barrier, no waits in between
      const int n = 64 * 1024 / (int)bytes < 64 ? 64 * 1024 / (int)bytes : 64;
      const long long t0 = clock64();
      mbar_expect_tx(&mybar[0], (uint32_t)n * bytes);
      for (int i = 0; i < n; ++i)
        tma_load_2d(&p.tm, &mybar[0], ring_ptr + (size_t)i * bytes, 0, row0 + (i % span) * p.rows);
      const long long t1 = clock64();
      mbar_wait(&mybar[0], 0u, nullptr, 0u);
      const long long t2 = clock64();
      p.out[blockIdx.x] = ((unsigned long long)(t1 - t0) << 32) | (unsigned long long)(t2 - t0);
      return;
    }
    const long long t0 = clock64();
    for (int i = 0; i < p.reps + nslots; ++i) {
      const int s = i % nslots;
      if (i >= nslots) mbar_wait(&mybar[s], (uint32_t)((i / nslots - 1) & 1), nullptr, 0u);
      if (i < p.reps) {
        mbar_expect_tx(&mybar[s], bytes);
        tma_load_2d(&p.tm, &mybar[s], myring + (size_t)s * bytes, 0, row0 + ((i * p.lanes + me) % span) * p.rows);
      }
    }
    if (me == 0) p.out[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}
}  // namespace
}  // namespace advoc

// h_out[ctas] u64: cycles each CTA needed for `reps` loads of {128 B x rows} with `depth` (<= 8) in flight.
extern "C" __attribute__((visibility("default"))) int advoc_selftest_tma_rate(unsigned long long* h_out, int ctas,
                                                                              int rows, int reps, int depth, int lanes) {
  using namespace advoc;
  ADVOC_REQUIRE(h_out && ctas > 0 && ctas <= 1024 && rows >= 8 && rows <= 256 && depth >= 1 && depth <= 8, ADVOC_BAD_ARG,
                "bad arguments");
  ADVOC_REQUIRE(lanes >= 1 && lanes <= 8 && (size_t)lanes * depth * rows * 128 <= 200 * 1024, ADVOC_BAD_ARG,
                "ring does not fit");
  TmaProbeParams p = {};
  p.rows = rows; p.reps = reps; p.depth = depth; p.lanes = lanes;
  p.src_rows = 4096;                                   // 512 KB of source per CTA: L2 resident after the first pass
  float* src = nullptr;
  const size_t total_rows = (size_t)ctas * p.src_rows;
  ADVOC_CHECK_CUDA(cudaMalloc(&src, total_rows * 128));
  ADVOC_CHECK_CUDA(cudaMemset(src, 0, total_rows * 128));
  ADVOC_CHECK_CUDA(cudaMalloc(&p.out, ctas * sizeof(unsigned long long)));
  int st = tc::encode_tiled2d(&p.tm, src, 32, (long)total_rows, 128, 32, rows);
  if (st) return st;
  const int smem = (depth == 1 ? 64 * 1024 : lanes * depth * rows * 128) + 1024;
  ADVOC_CHECK_CUDA(cudaFuncSetAttribute(tma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  for (int pass = 0; pass < 2; ++pass) {               // first pass warms L2
    tma_rate_probe<<<ctas, 64, smem>>>(p);
    ADVOC_CHECK_CUDA(cudaGetLastError());
    ADVOC_CHECK_CUDA(cudaDeviceSynchronize());
  }
  ADVOC_CHECK_CUDA(cudaMemcpy(h_out, p.out, ctas * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(src);
  cudaFree(p.out);
  return ADVOC_OK;
}
