/* vcof_probes.h — hardware probes used while designing libvcof (descriptor conventions, TMA feed rates).
 * They are measurement tools, not product: built into tests/native/libvcof_probes.so (tests/native/build.sh), which
 * links against libvcof.so for the runtime helpers (tensor-map builders, vcof_last_error).  Nothing in libvcof.so or
 * videocof_b200/ references them. */
#ifndef VCOF_PROBES_H_
#define VCOF_PROBES_H_
#ifdef __cplusplus
extern "C" {
#endif

/* One 128x128x64 tcgen05 tile with hand-swizzled operands (no TMA): d_f32[128,128] =
 * a_bf16[128,64] x b_bf16[128,64]^T.  mode bit0: B staged MN-major; bit1: A fed from TMEM.
 * Pins the descriptor conventions the GEMM / attention kernels depend on. */
int vcof_debug_umma_probe(const void* a, const void* b, float* d, int mode, void* stream);

/* Sustained TMA box-load rate: in every CTA `producers` (1..4) lanes of different warps each stream `iters` boxes of
 * the given rank-2 / rank-5 bf16 view through their own shared-memory ring (~192 KB in flight in total, no compute);
 * cycles[cta] receives the elapsed SM clocks.  flags: bits 0-3 boxes per barrier round trip (0 = 1), bit 4 alternate two
 * copies of the tensor map, bit 5 poll with mbarrier.test_wait instead of try_wait.  Explains the feed-rate
 * ceilings quoted in profiles/ (rows of 64 B vs 128 B, strided pixel slices vs contiguous rows). */
int vcof_debug_tma_probe(const void* base, int rank, const long long* dims, const long long* strides, const int* box,
                         int swizzle_bytes, int iters, const int* coords, int step_dim, int step, int wrap,
                         unsigned long long* cycles, int grid, int producers, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif
