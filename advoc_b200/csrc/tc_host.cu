// Host half of the tcgen05 kernels: TMA tensor-map encoders through driver entry points fetched at
// run time (the library has no link-time libcuda dependency and loads on a box without a driver).
#include "tc_ptx.cuh"

namespace advoc {
namespace tc {

namespace {
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
}  // namespace

bool tma_ok() { return driver().ok; }

// Device word [0] = code of the first pipeline wait that timed out (later launches see it and return
// at once); bytes 8..15 hold the device-visible address of a pinned, mapped HOST copy of the same
// code, so the host side can notice an aborted launch at any of its own sync points without a CUDA
// call (advoc_debug_peek: MelToMag / TrainEngine / bench.py raise on it).
static volatile unsigned int* g_host_flag = nullptr;

unsigned int* debug_word() {
  static unsigned int* w = [] {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, 64) != cudaSuccess) return (unsigned int*)nullptr;
    cudaMemset(p, 0, 64);
    unsigned int* h = nullptr;
    unsigned int* h_dev = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(&h_dev, h, 0) == cudaSuccess) {
      memset(h, 0, 64);
      g_host_flag = h;
      cudaMemcpy(reinterpret_cast<char*>(p) + 8, &h_dev, sizeof(h_dev), cudaMemcpyHostToDevice);
    } else {
      cudaGetLastError();
    }
    return p;
  }();
  return w;
}

volatile unsigned int* debug_host_word() {
  debug_word();
  return g_host_flag;
}

namespace {
inline CUtensorMapDataType map_dtype(int half) {
  return half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}
}  // namespace

int encode_im2col(CUtensorMap* tm, const void* x, int Nimg, int Hin, int Win, int ld, int C, int lower_h,
                  int lower_w, int upper_h, int upper_w, int trav_h, int trav_w, int box_c, int box_pix, bool atom32,
                  int half) {
  const Driver& drv = driver();
  ADVOC_REQUIRE(drv.ok, ADVOC_UNSUPPORTED, "TMA tensor-map encoders unavailable");
  const cuuint64_t es = half ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)Win * ld * es, (cuuint64_t)Hin * Win * ld * es};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)trav_w, (cuuint32_t)trav_h, 1};
  CUresult r = drv.im2col(tm, map_dtype(half), 4, const_cast<void*>(x), dims, strides, lower,
                          upper, (cuuint32_t)box_c, (cuuint32_t)box_pix, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADVOC_REQUIRE(r == CUDA_SUCCESS, ADVOC_CUDA_ERROR,
                "cuTensorMapEncodeIm2col failed (%d) dims %d,%d,%d,%d ld %d corners (%d,%d)-(%d,%d)", (int)r, C,
                Win, Hin, Nimg, ld, lower_w, lower_h, upper_w, upper_h);
  // Known driver issue (<= 13.1): the im2col encoder mis-sets a descriptor bit for tensors smaller
  // than 128 KiB; NVIDIA's own CUTLASS applies the same correction.
  if (drv.version <= 13010 && (size_t)Nimg * Hin * Win * ld * es < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return ADVOC_OK;
}

int encode_tiled2d(CUtensorMap* tm, const void* p, int inner, long rows, size_t row_stride_bytes, int box_inner,
                   int box_rows, bool atom32, int half) {
  const Driver& drv = driver();
  ADVOC_REQUIRE(drv.ok, ADVOC_UNSUPPORTED, "TMA tensor-map encoders unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = drv.tiled(tm, map_dtype(half), 2, const_cast<void*>(p), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADVOC_REQUIRE(r == CUDA_SUCCESS, ADVOC_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d) inner %d rows %ld", (int)r,
                inner, rows);
  return ADVOC_OK;
}

int encode_tiled4d(CUtensorMap* tm, const void* x, int C, int W, int H, int Nimg, long stride_w, long stride_h,
                   long stride_n, int box_c, int box_w, int box_h, int half) {
  const Driver& drv = driver();
  ADVOC_REQUIRE(drv.ok, ADVOC_UNSUPPORTED, "TMA tensor-map encoders unavailable");
  const cuuint64_t es = half ? 2 : 4;
  const size_t inner_bytes = (size_t)box_c * es;
  ADVOC_REQUIRE(inner_bytes == 128 || inner_bytes == 64, ADVOC_UNSUPPORTED, "box rows must be 64 or 128 bytes");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
  cuuint64_t strides[3] = {(cuuint64_t)stride_w * es, (cuuint64_t)stride_h * es, (cuuint64_t)stride_n * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = drv.tiled(tm, map_dtype(half), 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADVOC_REQUIRE(r == CUDA_SUCCESS, ADVOC_CUDA_ERROR,
                "cuTensorMapEncodeTiled(4d) failed (%d) dims %d,%d,%d,%d strides %ld,%ld,%ld box %d,%d,%d", (int)r, C,
                W, H, Nimg, stride_w, stride_h, stride_n, box_c, box_w, box_h);
  return ADVOC_OK;
}

}  // namespace tc
}  // namespace advoc
