// umma_probe.cu — single-tile diagnostic for the tcgen05 operand layouts libvcof relies on.
// No TMA: threads fill shared memory with the 128B-swizzle formula by hand, so a failure
// here isolates descriptor / instruction-descriptor / TMEM-operand conventions from the
// TMA + pipeline logic of the production kernels.  Exercised by tools/gpu_probe_umma.py
// and tests/test_kernels_gpu.py.
//
//   D[128,128] (fp32) = A[128,64] (bf16) * B[128,64]^T (bf16)
//   mode bit0: B operand staged MN-major (as the attention V tile) instead of K-major
//   mode bit1: A operand fed from TMEM (as the attention P tile) instead of smem
//   mode bits 4-6: A is a [136, 64] matrix staged with the swizzle of its ABSOLUTE smem rows and the MMA reads the
//             128-row window starting `shift` rows in (a row-shifted view of a resident tile, as an im2col-free
//             convolution would take one per filter tap); bit 7: put (start_address >> 7) & 7 into the descriptor's
//             base-offset field (bits 49-51)
#include "../../../videocof_b200/csrc/vcof_common.cuh"
#include "../../../include/vcof.h"
#include "vcof_probes.h"

namespace vcof {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, float* __restrict__ d,
                  int mode) {
  __shared__ __align__(1024) uint8_t sA[136 * 128];   // 128 (+8 for the shifted-view mode) rows x 64 bf16
  __shared__ __align__(1024) uint8_t sB[128 * 128];   // K-major: 128 n-rows x 64 k | MN-major: 2 x [64 k-rows x 64 n]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const bool b_mn = mode & 1, a_tmem = mode & 2;
  const int shift = (mode >> 4) & 7;
  const bool use_bo = mode & 128;
  const int a_rows = shift ? 136 : 128;
  const int tid = threadIdx.x, warp = tid >> 5;

  // ---- stage operands (swizzle: 16-byte chunk index ^= row & 7) ----
  for (int i = tid; i < a_rows * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<bf16*>(sA + off) = a[r * 64 + c];
  }
  for (int i = tid; i < 128 * 64; i += 128) {
    const int n = i / 64, k = i % 64;  // b is [n][k]
    int off;
    if (!b_mn) {
      off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    } else {
      const int half = n >> 6, nn = n & 63;
      off = half * 8192 + k * 128 + (((nn >> 3) ^ (k & 7)) << 4) + (nn & 7) * 2;
    }
    *reinterpret_cast<bf16*>(sB + off) = b[n * 64 + k];
  }
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor (async) proxy
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t tD = tm, tA = tm + 128;

  if (a_tmem) {
    // thread i <-> lane i: its A row as 32 packed bf16x2 columns
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const __nv_bfloat162 v = __halves2bfloat162(a[tid * 64 + 2 * c], a[tid * 64 + 2 * c + 1]);
      pk[c] = *reinterpret_cast<const uint32_t*>(&v);
    }
    tmem_st32(tA + ((warp * 32u) << 16), pk);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128, false, b_mn);
    for (int k = 0; k < 4; ++k) {
      const uint64_t bd = b_mn ? make_desc_mnmajor_sw128(smem_u32(sB) + k * 2048, 8192, 1024)
                               : make_desc_kmajor_sw128(smem_u32(sB) + k * 32);
      if (a_tmem)
        umma_ts(tD, tA + k * 8, bd, idesc, k != 0);
      else
      {
        const uint32_t a_start = smem_u32(sA) + shift * 128;
        uint64_t ad = make_desc_kmajor_sw128(a_start + k * 32);
        if (use_bo) ad |= uint64_t((a_start >> 7) & 7) << 49;
        umma_ss(tD, ad, bd, idesc, k != 0);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld32(tD + ((warp * 32u) << 16) + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) d[tid * 128 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

}  // namespace vcof

extern "C" int vcof_debug_umma_probe(const void* a, const void* b, float* d, int mode,
                                     void* stream) {
  vcof::umma_probe_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const vcof::bf16*>(a), reinterpret_cast<const vcof::bf16*>(b), d, mode);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
