// tma_probe.cu — diagnostic: sustained TMA tile-load throughput per SM for the box shapes libvcof uses.
// 1-4 producer lanes (different warps) each issue box loads into their own shared-memory ring, one consumer lane
// per producer waits for them and frees the slots; nothing else runs.  Reports achieved bytes/clk/SM so a kernel's feed rate can be compared
// with what the TMA unit can deliver for that box geometry (rows of 64 B vs 128 B, 2-D vs 5-D views).
// Used by tools/tma_probe.py; results in profiles/.
#include "../../../videocof_b200/csrc/vcof_common.cuh"
#include "../../../include/vcof.h"
#include "vcof_probes.h"

namespace vcof {

__device__ __forceinline__ void probe_wait(uint32_t bar, uint32_t parity, int poll) {
  if (!poll) { mbar_wait(bar, parity); return; }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

struct TmaProbeArgs {
  int rank;             // tensor-map rank (2 or 5)
  int box_bytes;        // bytes per box
  int iters;            // boxes per CTA
  int c[5];             // base coordinates
  int step_dim, step;   // coordinate advanced per iteration (wraps at `wrap`)
  int wrap;
  int stages;           // ring depth per producer: ~192 KB in flight in total regardless of the box size
  int producers;        // independent producer/consumer lane pairs (one pair per 64 threads), each with its own ring
  int per_iter;         // boxes issued per barrier round trip (one mbarrier covers all of them)
  int two_maps;         // alternate between two copies of the tensor map
  int poll;             // wait with non-blocking mbarrier.test_wait polling instead of try_wait
  unsigned long long* cycles;  // [gridDim.x]
};

__global__ void __launch_bounds__(256, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm2, TmaProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_all[4 * 24], empty_all[4 * 24];
  const int S = p.stages;
  const int box_stride = (p.box_bytes + 1023) / 1024 * 1024;
  const int stage_bytes = box_stride * p.per_iter;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * 24; ++i) { mbar_init(smem_u32(&full_all[i]), 1); mbar_init(smem_u32(&empty_all[i]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int pid = threadIdx.x >> 6;
  uint64_t* full = full_all + pid * 24;
  uint64_t* empty = empty_all + pid * 24;
  uint8_t* ring = smem + pid * S * stage_bytes;
  const unsigned long long t0 = clock64();
  if ((threadIdx.x & 63) == 0 && pid < p.producers) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.iters; ++i) {
      probe_wait(smem_u32(&empty[s]), ph ^ 1, p.poll);
      mbar_expect_tx(smem_u32(&full[s]), p.box_bytes * p.per_iter);
      for (int b = 0; b < p.per_iter; ++b) {
        int c[5] = {p.c[0], p.c[1], p.c[2], p.c[3], p.c[4]};
        c[p.step_dim] += (((i * p.per_iter + b) * p.step) + (blockIdx.x * 4 + pid) * 7) % p.wrap;
        const CUtensorMap* m = (p.two_maps && (b & 1)) ? &tm2 : &tm;
        const uint32_t dst = smem_u32(ring + s * stage_bytes + b * box_stride);
        if (p.rank == 2) tma_load_2d(dst, m, smem_u32(&full[s]), c[0], c[1]);
        else tma_load_5d(dst, m, smem_u32(&full[s]), c[0], c[1], c[2], c[3], c[4]);
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if ((threadIdx.x & 63) == 32 && pid < p.producers) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.iters; ++i) {
      probe_wait(smem_u32(&full[s]), ph, p.poll);
      mbar_arrive(smem_u32(&empty[s]));
      if (++s == S) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
}

}  // namespace vcof

using namespace vcof;

extern "C" int vcof_debug_tma_probe(const void* base, int rank, const long long* dims, const long long* strides,
                                    const int* box, int swizzle_bytes, int iters, const int* coords, int step_dim,
                                    int step, int wrap, unsigned long long* cycles, int grid, int producers,
                                    int flags, void* stream) {
  CUtensorMap tm;
  uint64_t d[5], s[4];
  uint32_t b[5];
  long long box_elems = 1;
  for (int i = 0; i < rank; ++i) { d[i] = (uint64_t)dims[i]; b[i] = (uint32_t)box[i]; box_elems *= box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = (uint64_t)strides[i] * 2;
  int rc = make_tmap_nd_bf16(&tm, base, rank, d, s, b, swizzle_bytes);
  if (rc) return rc;
  TmaProbeArgs a;
  a.rank = rank;
  a.box_bytes = (int)(box_elems * 2);
  a.iters = iters;
  for (int i = 0; i < 5; ++i) a.c[i] = i < rank ? coords[i] : 0;
  a.step_dim = step_dim; a.step = step; a.wrap = wrap;
  a.cycles = cycles;
  const int stage_bytes = (a.box_bytes + 1023) / 1024 * 1024;
  VCOF_REQUIRE(producers >= 1 && producers <= 4, "vcof_debug_tma_probe: producers must be 1..4");
  a.producers = producers;
  a.per_iter = (flags & 15) ? (flags & 15) : 1;
  a.two_maps = (flags >> 4) & 1;
  a.poll = (flags >> 5) & 1;
  a.stages = (192 * 1024) / (stage_bytes * a.per_iter) / producers;
  if (a.stages > 24) a.stages = 24;
  VCOF_REQUIRE(a.stages >= 2, "vcof_debug_tma_probe: box too large");
  const int smem = producers * a.stages * stage_bytes * a.per_iter + 1024;
  VCOF_CHECK_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_probe_kernel<<<grid, 64 * producers, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm, tm, a);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
