// kbench.cu — Python-free timing of the two tensor-core kernels that dominate a denoising step, over the C ABI:
//   kbench gemm M N K [epilogue=0] [reps=20]      D[M,N] = A[M,K] x W[N,K]^T   (vcof_gemm_bf16)
//   kbench attn Lq Lk heads [reps=5]              softmax(Q K^T / sqrt(128)) V  (vcof_attn_fwd, head_dim 128)
// CUDA events on the launching stream, 3 warm-up launches, then `reps` timed launches; one JSON line per run,
// appended to $KBENCH_OUT when set.  The library's own knobs apply (VCOF_GEMM_2CTA, VCOF_GEMM_PRODUCERS,
// VCOF_ATTN_EMU, …), so an A/B is two invocations.  A gpurun call with this binary costs ~30 s of box time against
// minutes for the torch-based tools/kbench.py, and `ncu … tests/native/kbench attn 75600 75600 40 1` profiles the
// C2 self-attention launch without a Python start-up.  Inputs are random bf16 (values matter for power draw).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/vcof.h"

typedef __nv_bfloat16 bf16;

static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static float frand() {  // cheap symmetric noise in (-1, 1)
  g_seed ^= g_seed >> 12; g_seed ^= g_seed << 25; g_seed ^= g_seed >> 27;
  return (float)((int64_t)((g_seed * 0x2545F4914F6CDD1Dull) >> 40) - (1 << 23)) * (1.0f / (1 << 23));
}
static bf16* dev_noise(size_t n, float scale) {
  const size_t chunk = (size_t)4 << 20;
  static std::vector<bf16> host;
  if (host.empty()) {
    host.resize(chunk + 8192);
    for (auto& v : host) v = __float2bfloat16_rn(frand());
  }
  bf16* d = nullptr;
  if (cudaMalloc(&d, n * 2) != cudaSuccess) { fprintf(stderr, "cudaMalloc(%zu) failed\n", n * 2); exit(2); }
  std::vector<bf16> scaled(chunk);
  for (size_t off = 0, rot = 0; off < n; off += chunk, rot = (rot + 997) % 8192) {
    const size_t m = n - off < chunk ? n - off : chunk;
    for (size_t i = 0; i < m; ++i) scaled[i] = __float2bfloat16_rn(__bfloat162float(host[i + rot]) * scale);
    cudaMemcpy(d + off, scaled.data(), m * 2, cudaMemcpyHostToDevice);
  }
  return d;
}
static void emit(const char* line) {
  puts(line);
  const char* path = getenv("KBENCH_OUT");
  if (path) { FILE* f = fopen(path, "a"); if (f) { fputs(line, f); fputc('\n', f); fclose(f); } }
}
static const char* env_or(const char* k) { const char* e = getenv(k); return e ? e : ""; }

static int bench_gemm(int M, int N, int K, int epi, int reps) {
  bf16* a = dev_noise((size_t)M * K, 1.0f);
  bf16* w = dev_noise((size_t)N * K, 1.0f / sqrtf((float)K));
  bf16* bias = dev_noise(N, 0.1f);
  const int base = epi & 0xff;
  const bool f32_out = base == VCOF_EPI_BIAS_GATE_RES_F32 || base == VCOF_EPI_BIAS_F32 || base == VCOF_EPI_RAW_F32;
  void* out = nullptr;
  cudaMalloc(&out, (size_t)M * N * (f32_out ? 4 : 2));
  cudaMemset(out, 0, (size_t)M * N * (f32_out ? 4 : 2));
  std::vector<float> gate_h(N, 0.5f);
  float* gate = nullptr;
  cudaMalloc(&gate, N * 4);
  cudaMemcpy(gate, gate_h.data(), N * 4, cudaMemcpyHostToDevice);
  const bool needs_gate = base == VCOF_EPI_BIAS_GATE_RES_F32 || base == VCOF_EPI_GATE_ACCUM_BF16;
  const void* b = base == VCOF_EPI_GATE_ACCUM_BF16 ? nullptr : bias;
  int rc = 0;
  for (int i = 0; i < 3; ++i) rc |= vcof_gemm_bf16(a, K, w, K, b, needs_gate ? gate : nullptr, out, N, M, N, K, epi, nullptr);
  if (rc || cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", vcof_last_error()); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < reps; ++i) vcof_gemm_bf16(a, K, w, K, b, needs_gate ? gate : nullptr, out, N, M, N, K, epi, nullptr);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  char line[512];
  snprintf(line, sizeof line, "{\"kernel\": \"gemm\", \"M\": %d, \"N\": %d, \"K\": %d, \"epilogue\": %d, \"reps\": %d, "
           "\"ms\": %.4f, \"tflops\": %.1f, \"VCOF_GEMM_2CTA\": \"%s\", \"VCOF_GEMM_PRODUCERS\": \"%s\"}", M, N, K, epi, reps, ms,
           2.0 * M * N * K / (ms * 1e-3) / 1e12, env_or("VCOF_GEMM_2CTA"), env_or("VCOF_GEMM_PRODUCERS"));
  emit(line);
  return 0;
}

static int bench_attn(int Lq, int Lk, int heads, int reps) {
  const size_t C = (size_t)heads * 128;
  bf16* q = dev_noise((size_t)Lq * C, 1.0f);
  bf16* k = dev_noise((size_t)Lk * C, 1.0f);
  bf16* v = dev_noise((size_t)Lk * C, 1.0f);
  bf16* out = nullptr;
  cudaMalloc(&out, (size_t)Lq * C * 2);
  const float scale = 1.0f / sqrtf(128.0f);
  int rc = 0;
  for (int i = 0; i < 2; ++i) rc |= vcof_attn_fwd(q, C, k, C, v, C, out, C, Lq, Lk, Lk, heads, 128, scale, 0, nullptr);
  if (rc || cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", vcof_last_error()); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < reps; ++i) vcof_attn_fwd(q, C, k, C, v, C, out, C, Lq, Lk, Lk, heads, 128, scale, 0, nullptr);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  char line[512];
  snprintf(line, sizeof line, "{\"kernel\": \"attn\", \"Lq\": %d, \"Lk\": %d, \"heads\": %d, \"reps\": %d, \"ms\": %.4f, "
           "\"tflops\": %.1f, \"VCOF_ATTN_EMU\": \"%s\", \"VCOF_ATTN_PSPLIT\": \"%s\", \"VCOF_ATTN_SPLIT_S\": \"%s\"}", Lq, Lk, heads,
           reps, ms, 4.0 * Lq * (double)Lk * C / (ms * 1e-3) / 1e12, env_or("VCOF_ATTN_EMU"), env_or("VCOF_ATTN_PSPLIT"),
           env_or("VCOF_ATTN_SPLIT_S"));
  emit(line);
  return 0;
}

int main(int argc, char** argv) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { puts("{\"error\": \"no CUDA device\"}"); return 3; }
  if (argc >= 5 && strcmp(argv[1], "gemm") == 0)
    return bench_gemm(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 0, argc > 6 ? atoi(argv[6]) : 20);
  if (argc >= 5 && strcmp(argv[1], "attn") == 0)
    return bench_attn(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argc > 5 ? atoi(argv[5]) : 5);
  fprintf(stderr, "usage: kbench gemm M N K [epilogue] [reps] | kbench attn Lq Lk heads [reps]\n");
  return 64;
}
