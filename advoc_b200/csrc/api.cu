// C-ABI glue: error plumbing, device queries and the convolution dispatchers.
// No exception or abort crosses this boundary; see include/advoc_b200.h.
#include "common.cuh"

#include <cuda_fp16.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>

// the ctypes mirror in advoc_b200/_native.py relies on these layouts (tests/test_boundary.py)
static_assert(sizeof(advoc_conv_desc) == 56, "advoc_conv_desc layout");
static_assert(sizeof(advoc_epilogue) == 152 && offsetof(advoc_epilogue, out0_row_pad) == 144 && offsetof(advoc_epilogue, d_seed) == 128 &&
                  offsetof(advoc_epilogue, out0_dtype) == 136 && offsetof(advoc_epilogue, d_gate) == 96 &&
                  offsetof(advoc_epilogue, gate_scale1) == 124 && offsetof(advoc_epilogue, d_out0) == 24 &&
                  offsetof(advoc_epilogue, d_out1) == 40 &&
                  offsetof(advoc_epilogue, d_dropout_mask) == 64 &&
                  offsetof(advoc_epilogue, seed) == 80,
              "advoc_epilogue layout");

namespace advoc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(advoc_status st, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return (int)st;
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

int device_arch() {
  static int arch = -1;
  if (arch < 0) {
    int dev = 0, ma = 0, mi = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
    arch = ma * 10 + mi;
  }
  return arch;
}

// defined in conv_simt.cu / conv_tc.cu
int check_conv_desc(const advoc_conv_desc* d);
int conv_fwd_simt(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                  const advoc_epilogue* ep, void* stream);
int conv_transposed_simt(const advoc_conv_desc* d, const float* x, int ldx, const float* w,
                         const advoc_epilogue* ep, void* stream);
bool conv_fwd_tc_eligible(const advoc_conv_desc* d, int ldx);
bool conv_transposed_tc_eligible(const advoc_conv_desc* d, int ldx);
int conv_fwd_tc(const advoc_conv_desc* d, const void* x, int ldx, const void* w,
                const advoc_epilogue* ep, void* stream);
int conv_transposed_tc(const advoc_conv_desc* d, const void* x, int ldx, const void* w,
                       const advoc_epilogue* ep, void* stream);
bool conv_p2d_eligible(const advoc_conv_desc* d, int ldx, int transposed, int store_w);
bool conv_p2d_preferred(const advoc_conv_desc* d, int ldx, int transposed, int store_w, const advoc_epilogue* ep);
int conv_p2d(const advoc_conv_desc* d, int transposed, const void* x, int ldx, const void* w,
             const advoc_epilogue* ep, void* stream);
bool tc_epilogue_ok(const advoc_epilogue* ep);
bool conv_to_one_tc_eligible(const advoc_conv_desc* d, const void* x, int ldx, const void* w, const advoc_epilogue* ep);
int conv_to_one_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                   void* stream);
bool conv_one_in_tc_eligible(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep);
bool deconv_from_one_tc_eligible(const advoc_conv_desc* d, const float* x, const float* w, const advoc_epilogue* ep);
int deconv_from_one_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                       void* stream);
int conv_one_in_tc(const advoc_conv_desc* d, const float* x, int ldx, const float* w, const advoc_epilogue* ep,
                   void* stream);
int conv_tc_tile_n(const advoc_conv_desc* d, int transposed, int store_w);
int conv_p2d_tile_n(const advoc_conv_desc* d, int transposed, int store_w);
bool deconv_one_tc_geometry(const advoc_conv_desc* d, int ldx);
bool deconv_one_tc_eligible(const advoc_conv_desc* d, const void* x, int ldx, const advoc_epilogue* ep);
int deconv_one_tc(const advoc_conv_desc* d, const void* x, int ldx, const void* w, const advoc_epilogue* ep,
                  void* stream);

}  // namespace advoc

using namespace advoc;

extern "C" int advoc_version(void) { return ADVOC_B200_VERSION; }

extern "C" int advoc_last_error(char* buf, size_t buf_len) {
  if (!buf || buf_len == 0) return ADVOC_BAD_ARG;
  strncpy(buf, g_err, buf_len - 1);
  buf[buf_len - 1] = 0;
  return ADVOC_OK;
}

extern "C" unsigned long long advoc_launch_count(void) {
  return __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
}

extern "C" int advoc_device_arch(int* arch) {
  ADVOC_REQUIRE(arch != nullptr, ADVOC_BAD_ARG, "arch is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(ADVOC_CUDA_ERROR, "no CUDA device: %s", cudaGetErrorString(e));
  }
  *arch = device_arch();
  return ADVOC_OK;
}

static int check_conv_io(const advoc_conv_desc* d, const void* x, int ldx, int cx, const void* w,
                         const advoc_epilogue* ep) {
  int st = check_conv_desc(d);
  if (st) return st;
  ADVOC_REQUIRE(x && w && ep, ADVOC_BAD_ARG, "NULL x/w/epilogue");
  ADVOC_REQUIRE(ldx >= cx, ADVOC_BAD_SHAPE, "ld_x %d smaller than the channel count %d", ldx, cx);
  ADVOC_REQUIRE(d->math >= ADVOC_MATH_AUTO && d->math <= ADVOC_MATH_F16, ADVOC_BAD_ARG,
                "unknown math mode %d", d->math);
  return ADVOC_OK;
}

extern "C" int advoc_conv2d_fwd(const advoc_conv_desc* d, const void* d_xv, int ld_x,
                                const void* d_wv, const advoc_epilogue* ep, void* stream) {
  int st = check_conv_io(d, d_xv, ld_x, d ? d->Cin : 0, d_wv, ep);
  if (st) return st;
  const float* d_x = static_cast<const float*>(d_xv);   // half data when math == ADVOC_MATH_F16: only the
  const float* d_w = static_cast<const float*>(d_wv);   // tcgen05 kernels below ever see those pointers
  if (conv_one_in_tc_eligible(d, d_x, ld_x, d_w, ep)) return conv_one_in_tc(d, d_x, ld_x, d_w, ep, stream);
  if (conv_to_one_tc_eligible(d, d_x, ld_x, d_w, ep)) {
    st = conv_to_one_tc(d, d_x, ld_x, d_w, ep, stream);
    if (st != ADVOC_UNSUPPORTED) return st;      // scratch not available inside a graph capture: CUDA-core kernel
  }
  if (ep->out0_row_pad != 0) {
    ADVOC_REQUIRE(d->Cin <= 2 && d->math != ADVOC_MATH_F16, ADVOC_UNSUPPORTED,
                  "out0_row_pad is only supported by the thin-input convolution");
    return conv_fwd_simt(d, d_x, ld_x, d_w, ep, stream);
  }
  if (d->math == ADVOC_MATH_FP32) return conv_fwd_simt(d, d_x, ld_x, d_w, ep, stream);
  const bool ok = conv_fwd_tc_eligible(d, ld_x);
  if (d->math == ADVOC_MATH_TF32 || d->math == ADVOC_MATH_F16 || ok) {
    ADVOC_REQUIRE(ok, ADVOC_UNSUPPORTED, "conv shape not eligible for the tcgen05 path");
    if (conv_p2d_preferred(d, ld_x, 0, 0, ep)) {
      ADVOC_REQUIRE(aligned16(d_x) && aligned16(d_w), ADVOC_BAD_ALIGN, "x / w must be 16-byte aligned");
      ADVOC_REQUIRE(tc_epilogue_ok(ep), ADVOC_BAD_ALIGN,
                    "tcgen05 path needs 16-byte aligned outputs with ld and channel offset multiples of 4");
      return conv_p2d(d, 0, d_x, ld_x, d_w, ep, stream);
    }
    return conv_fwd_tc(d, d_x, ld_x, d_w, ep, stream);
  }
  return conv_fwd_simt(d, d_x, ld_x, d_w, ep, stream);
}

extern "C" int advoc_conv2d_transpose_fwd(const advoc_conv_desc* d, const void* d_xv, int ld_x,
                                          const void* d_wv, const advoc_epilogue* ep,
                                          void* stream) {
  int st = check_conv_io(d, d_xv, ld_x, d ? d->Cout : 0, d_wv, ep);
  if (st) return st;
  const float* d_x = static_cast<const float*>(d_xv);
  const float* d_w = static_cast<const float*>(d_wv);
  ADVOC_REQUIRE(ep->out0_row_pad == 0, ADVOC_UNSUPPORTED, "out0_row_pad is only supported by the thin-input convolution");
  if (d->math == ADVOC_MATH_FP32) return conv_transposed_simt(d, d_x, ld_x, d_w, ep, stream);
  if (deconv_from_one_tc_eligible(d, d_x, d_w, ep)) return deconv_from_one_tc(d, d_x, ld_x, d_w, ep, stream);
  if (deconv_one_tc_eligible(d, d_x, ld_x, ep)) return deconv_one_tc(d, d_x, ld_x, d_w, ep, stream);
  const bool ok = conv_transposed_tc_eligible(d, ld_x);
  if (d->math == ADVOC_MATH_TF32 || d->math == ADVOC_MATH_F16 || ok) {
    ADVOC_REQUIRE(ok, ADVOC_UNSUPPORTED, "conv_transpose shape not eligible for the tcgen05 path");
    if (conv_p2d_preferred(d, ld_x, 1, ep->store_w, ep)) {
      ADVOC_REQUIRE(aligned16(d_x) && aligned16(d_w), ADVOC_BAD_ALIGN, "x / w must be 16-byte aligned");
      ADVOC_REQUIRE(tc_epilogue_ok(ep), ADVOC_BAD_ALIGN,
                    "tcgen05 path needs 16-byte aligned outputs with ld and channel offset multiples of 4");
      return conv_p2d(d, 1, d_x, ld_x, d_w, ep, stream);
    }
    return conv_transposed_tc(d, d_x, ld_x, d_w, ep, stream);
  }
  return conv_transposed_simt(d, d_x, ld_x, d_w, ep, stream);
}

extern "C" int advoc_conv2d_path(const advoc_conv_desc* d, int ld_x, int transposed) {
  if (!d || d->math == ADVOC_MATH_FP32) return ADVOC_MATH_FP32;
  const int tensor = d->math == ADVOC_MATH_F16 ? ADVOC_MATH_F16 : ADVOC_MATH_TF32;
  if (transposed && deconv_one_tc_geometry(d, ld_x)) return tensor;
  const bool ok = transposed ? conv_transposed_tc_eligible(d, ld_x) : conv_fwd_tc_eligible(d, ld_x);
  return ok ? tensor : ADVOC_MATH_FP32;
}

extern "C" int advoc_conv2d_kernel(const advoc_conv_desc* d, int ld_x, int transposed, int store_w) {
  // the one-input-channel forward conv runs on the tensor cores with fp32-exact (3 x tf32) products; it keeps
  // the TF-layout filter, so advoc_conv2d_path still reports ADVOC_MATH_FP32 for it (no packed copy)
  if (d && !transposed && d->math != ADVOC_MATH_FP32 && d->math != ADVOC_MATH_F16 &&
      ((d->kh == 4 && d->kw == 4 && (d->Cin == 1 || d->Cin == 2)) || (d->kh == 5 && d->kw == 5 && d->Cin == 1)) &&
      (d->Cout == 32 || d->Cout == 64 || d->Cout % 128 == 0) && device_arch() == 100 &&
      getenv("ADVOC_NO_ONE_IN_TC") == nullptr)
    return 4;
  if (advoc_conv2d_path(d, ld_x, transposed) == ADVOC_MATH_FP32) return 0;
  if (transposed && deconv_one_tc_geometry(d, ld_x)) return 3;
  return conv_p2d_preferred(d, ld_x, transposed, store_w, nullptr) ? 2 : 1;   // (a forward call: see conv_p2d_preferred)
}

extern "C" int advoc_conv2d_tile_n(const advoc_conv_desc* d, int ld_x, int transposed, int store_w) {
  switch (advoc_conv2d_kernel(d, ld_x, transposed, store_w)) {
    case 1: return conv_tc_tile_n(d, transposed, store_w);
    case 2: return conv_p2d_tile_n(d, transposed, store_w);
    case 3: return 16;
    case 4: return d->Cout;
    default: return 0;
  }
}

// ---------------------------------------------------------------------------------------------
// filter re-pack: [taps, A, B] -> [taps, B, A] (transpose != 0) or copy, optionally TF32-rounded
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void pack_filter_kernel(const float* __restrict__ in, float* __restrict__ out, int taps,
                                   int A, int B, int transpose, int mode) {
  const long total = (long)taps * A * B;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    // i indexes the OUTPUT so that writes are coalesced
    long src = i;
    if (transpose) {
      const int a = (int)(i % A);
      const long r = i / A;
      const int b = (int)(r % B);
      const long t = r / B;
      src = (t * A + a) * B + b;
    }
    float v = __ldg(in + src);
    if (mode == 2) {
      reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
    } else {
      if (mode == 1) v = advoc::round_tf32(v);
      out[i] = v;
    }
  }
}
// transpose through a 32 x 33 shared-memory tile: reads coalesced along B, writes along A (the element-wise kernel
// above reads with stride B: 16 us per layer on the regular model, ~3x its HBM time)
__global__ void __launch_bounds__(256) pack_filter_transpose_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                    int A, int B, int mode) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const size_t base = (size_t)blockIdx.z * A * B;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int a = a0 + ty + k, b = b0 + tx;
    tile[ty + k][tx] = (a < A && b < B) ? __ldg(in + base + (size_t)a * B + b) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int b = b0 + ty + k, a = a0 + tx;
    if (a < A && b < B) {
      const float v = tile[tx][ty + k];
      const size_t o = base + (size_t)b * A + a;
      if (mode == 2) reinterpret_cast<__half*>(out)[o] = __float2half_rn(v);
      else out[o] = mode == 1 ? advoc::round_tf32(v) : v;
    }
  }
}
}  // namespace

extern "C" int advoc_pack_filter(const float* d_w, void* d_packed, int taps, int A, int B,
                                 int transpose, int mode, void* stream) {
  ADVOC_REQUIRE(d_w && d_packed, ADVOC_BAD_ARG, "NULL filter pointer");
  ADVOC_REQUIRE(mode >= 0 && mode <= 2, ADVOC_BAD_ARG, "mode must be 0 (copy), 1 (tf32) or 2 (fp16)");
  ADVOC_REQUIRE(taps > 0 && A > 0 && B > 0, ADVOC_BAD_SHAPE, "bad filter shape");
  const long total = (long)taps * A * B;
  if (transpose && A >= 8 && B >= 8 && taps <= 65535 && (A + 31) / 32 <= 65535) {
    dim3 grid((unsigned)((B + 31) / 32), (unsigned)((A + 31) / 32), (unsigned)taps);
    pack_filter_transpose_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        d_w, static_cast<float*>(d_packed), A, B, mode);
    count_launch();
    ADVOC_CHECK_CUDA(cudaGetLastError());
    return ADVOC_OK;
  }
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_filter_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_w, static_cast<float*>(d_packed), taps, A, B, transpose, mode);
  count_launch();
  ADVOC_CHECK_CUDA(cudaGetLastError());
  return ADVOC_OK;
}
