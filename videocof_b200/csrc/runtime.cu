// runtime.cu — host-side plumbing of libvcof: last-error slot, TMA descriptor
// encoding through the driver entry point, device queries.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_last_error("CUDA error %d (%s) at %s", int(e), cudaGetErrorString(e), what);
  return -2;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static int encode(CUtensorMap* map, const void* gptr, uint32_t rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  auto fn = encode_fn();
  VCOF_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  VCOF_REQUIRE((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, "TMA base pointer not 16B aligned");
  for (uint32_t i = 0; i + 1 < rank; ++i)
    VCOF_REQUIRE((gstr[i] & 15) == 0, "TMA stride %u (%llu B) not a multiple of 16", i,
                 (unsigned long long)gstr[i]);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(gptr), gdim, gstr,
                  bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VCOF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %u)",
               int(r), rank);
  return 0;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* gptr, uint64_t inner, uint64_t outer,
                      uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t str[1] = {outer_stride_bytes};
  uint32_t box[2] = {box_inner, box_outer};
  return encode(map, gptr, 2, dims, str, box, 128);
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* gptr, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1,
                      uint32_t b2) {
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t str[2] = {stride1_bytes, stride2_bytes};
  uint32_t box[3] = {b0, b1, b2};
  return encode(map, gptr, 3, dims, str, box, 128);
}

int make_tmap_nd_bf16(CUtensorMap* map, const void* gptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return encode(map, gptr, (uint32_t)rank, dims, strides_bytes, box, swizzle_bytes);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

}  // namespace vcof

extern "C" const char* vcof_last_error(void) { return vcof::g_last_error; }
extern "C" int vcof_abi_version(void) { return VCOF_ABI_VERSION; }
