// selftest_t5.cu — stand-alone parity check of the text-encoder entry points of libvcof (no Python, no torch):
// random bf16 inputs, the C-ABI call on the GPU, a double-precision CPU restatement of the same contract
// (include/vcof.h) in this file, relative Frobenius errors printed as one JSON object per case.
// Test infrastructure: built by tests/native/build.sh, run on a GPU box as `tests/native/selftest_t5 [out.jsonl]`.
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
static uint64_t g_seed = 0x243F6A8885A308D3ull;
static double urand() {  // xorshift64*, uniform (0, 1)
  g_seed ^= g_seed >> 12; g_seed ^= g_seed << 25; g_seed ^= g_seed >> 27;
  return ((g_seed * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) + 1e-17;
}
static double nrand() { return sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand()); }
static float bfr(double x) { return __bfloat162float(__float2bfloat16_rn((float)x)); }

static std::vector<bf16> rand_bf16(size_t n, double scale, double mean = 0.0) {
  std::vector<bf16> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = __float2bfloat16_rn((float)(mean + scale * nrand()));
  return v;
}
template <class T> static T* to_dev(const std::vector<T>& h) {
  T* d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(2); }
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}
template <class T> static std::vector<T> to_host(const T* d, size_t n) {
  std::vector<T> h(n);
  cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost);
  return h;
}
static FILE* g_out = nullptr;
static int g_fail = 0;
static void report(const char* name, const char* shape, const std::vector<bf16>& got, const std::vector<double>& ref,
                   double tol, int rc) {
  double num = 0, den = 0;
  bool nan = false;
  for (size_t i = 0; i < ref.size(); ++i) {
    const double g = __bfloat162float(got[i]);
    if (g != g) nan = true;
    num += (g - ref[i]) * (g - ref[i]);
    den += ref[i] * ref[i];
  }
  const double rel = sqrt(num / (den + 1e-300));
  const cudaError_t e = cudaDeviceSynchronize();
  const bool ok = rc == 0 && e == cudaSuccess && !nan && rel < tol;
  if (!ok) ++g_fail;
  char line[512];
  snprintf(line, sizeof line, "{\"case\": \"%s\", \"shape\": \"%s\", \"rc\": %d, \"cuda\": %d, \"rel_fro\": %.3e, "
           "\"tol\": %.1e, \"nan\": %s, \"ok\": %s%s%s%s}", name, shape, rc, (int)e, rel, tol, nan ? "true" : "false",
           ok ? "true" : "false", rc ? ", \"error\": \"" : "", rc ? vcof_last_error() : "", rc ? "\"" : "");
  puts(line);
  if (g_out) { fputs(line, g_out); fputc('\n', g_out); fflush(g_out); }
}

static void case_rmsnorm(int rows, int C) {
  auto x = rand_bf16((size_t)rows * C, 3.0), w = rand_bf16(C, 0.1, 1.0);
  bf16 *dx = to_dev(x), *dw = to_dev(w), *dy = nullptr;
  cudaMalloc(&dy, (size_t)rows * C * 2);
  const int rc = vcof_t5_rmsnorm(dx, C, dw, dy, C, rows, C, 1e-6f, nullptr);
  std::vector<double> ref((size_t)rows * C);
  for (int r = 0; r < rows; ++r) {
    double sq = 0;
    for (int c = 0; c < C; ++c) { const double v = __bfloat162float(x[(size_t)r * C + c]); sq += v * v; }
    const double rs = 1.0 / sqrt(sq / C + 1e-6);
    for (int c = 0; c < C; ++c)
      ref[(size_t)r * C + c] = (double)__bfloat162float(w[c]) * bfr(__bfloat162float(x[(size_t)r * C + c]) * rs);
  }
  char shape[64]; snprintf(shape, sizeof shape, "rows=%d C=%d", rows, C);
  report("t5_rmsnorm", shape, to_host(dy, (size_t)rows * C), ref, 4e-3, rc);
  cudaFree(dx); cudaFree(dw); cudaFree(dy);
}

static void case_embed(int n, int vocab, int C) {
  auto tab = rand_bf16((size_t)vocab * C, 1.0);
  std::vector<long long> ids(n);
  for (int i = 0; i < n; ++i) ids[i] = (long long)(urand() * vocab) % vocab;
  bf16 *dt = to_dev(tab), *dy = nullptr;
  long long* di = to_dev(ids);
  cudaMalloc(&dy, (size_t)n * C * 2);
  const int rc = vcof_embed_rows(di, dt, C, vocab, dy, C, n, C, nullptr);
  std::vector<double> ref((size_t)n * C);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < C; ++c) ref[(size_t)i * C + c] = __bfloat162float(tab[(size_t)ids[i] * C + c]);
  char shape[64]; snprintf(shape, sizeof shape, "n=%d vocab=%d C=%d", n, vocab, C);
  report("embed_rows", shape, to_host(dy, (size_t)n * C), ref, 1e-12, rc);
  cudaFree(dt); cudaFree(di); cudaFree(dy);
}

static void case_gemm(int M, int N, int K, int epi, int flags = 0) {
  auto a = rand_bf16((size_t)M * K, 1.0), w = rand_bf16((size_t)N * K, 1.0 / sqrt((double)K));
  auto o0 = rand_bf16((size_t)M * N, 1.0);
  bf16 *da = to_dev(a), *dw = to_dev(w), *dout = to_dev(o0);
  const int rc = vcof_gemm_bf16(da, K, dw, K, nullptr, nullptr, dout, N, M, N, K, epi | flags, nullptr);
  std::vector<double> ref((size_t)M * N);
  std::vector<float> af((size_t)M * K), wf((size_t)N * K);
  for (size_t i = 0; i < af.size(); ++i) af[i] = __bfloat162float(a[i]);
  for (size_t i = 0; i < wf.size(); ++i) wf[i] = __bfloat162float(w[i]);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      const float *ar = &af[(size_t)m * K], *wr = &wf[(size_t)n * K];
      for (int k = 0; k < K; ++k) acc += (double)ar[k] * wr[k];
      const double lin = bfr(acc), prev = __bfloat162float(o0[(size_t)m * N + n]);
      ref[(size_t)m * N + n] = epi == VCOF_EPI_MUL_BF16 ? prev * lin : epi == VCOF_EPI_ADD_BF16 ? prev + lin
                               : epi == VCOF_EPI_BIAS_GELU_BF16
                                   ? 0.5 * lin * (1.0 + tanh(0.7978845608028654 * (lin + 0.044715 * lin * lin * lin)))
                                   : acc;
    }
  char shape[64]; snprintf(shape, sizeof shape, "M=%d N=%d K=%d%s", M, N, K, flags ? " tile128" : "");
  report(epi == VCOF_EPI_MUL_BF16 ? "gemm_mul" : epi == VCOF_EPI_ADD_BF16 ? "gemm_add"
         : epi == VCOF_EPI_BIAS_GELU_BF16 ? "gemm_gelu_nobias" : "gemm_nobias",
         shape, to_host(dout, (size_t)M * N), ref, 4e-3, rc);
  cudaFree(da); cudaFree(dw); cudaFree(dout);
}

static void case_attn(int B, int L, int heads, int d, const int* lens) {
  const int C = heads * d;
  const size_t n = (size_t)B * L * C;
  auto q = rand_bf16(n, 0.7), k = rand_bf16(n, 1.0), v = rand_bf16(n, 1.0);
  std::vector<float> bias((size_t)heads * (2 * L - 1));
  for (auto& b : bias) b = (float)nrand();
  std::vector<int> mask((size_t)B * L, 1);
  if (lens) for (int b = 0; b < B; ++b) for (int j = lens[b]; j < L; ++j) mask[(size_t)b * L + j] = 0;
  bf16 *dq = to_dev(q), *dk = to_dev(k), *dv = to_dev(v), *dout = nullptr;
  float* db = to_dev(bias);
  int* dm = to_dev(mask);
  cudaMalloc(&dout, n * 2);
  const int rc = vcof_t5_attn(dq, C, dk, C, dv, C, dout, C, db, 2 * L - 1, lens ? dm : nullptr, B, L, heads, d, nullptr);
  std::vector<double> ref(n), s(L);
  const double kMin = -3.3895313892515355e38;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < heads; ++h)
      for (int i = 0; i < L; ++i) {
        const bf16* qr = &q[((size_t)b * L + i) * C + h * d];
        double mx = -INFINITY;
        for (int j = 0; j < L; ++j) {
          const bf16* kr = &k[((size_t)b * L + j) * C + h * d];
          double acc = 0;
          for (int c = 0; c < d; ++c) acc += (double)__bfloat162float(qr[c]) * __bfloat162float(kr[c]);
          s[j] = (float)(acc + (mask[(size_t)b * L + j] ? bias[(size_t)h * (2 * L - 1) + (j - i) + L - 1] : kMin));
          mx = fmax(mx, s[j]);
        }
        double sum = 0;
        for (int j = 0; j < L; ++j) { s[j] = exp(s[j] - mx); sum += s[j]; }
        for (int c = 0; c < d; ++c) {
          double acc = 0;
          for (int j = 0; j < L; ++j)
            acc += (double)bfr(s[j] / sum) * __bfloat162float(v[((size_t)b * L + j) * C + h * d + c]);
          ref[((size_t)b * L + i) * C + h * d + c] = acc;
        }
      }
  char shape[96]; snprintf(shape, sizeof shape, "B=%d L=%d heads=%d d=%d masked=%d", B, L, heads, d, lens ? 1 : 0);
  report("t5_attn", shape, to_host(dout, n), ref, 6e-3, rc);
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout); cudaFree(db); cudaFree(dm);
}

// ---- --bench: the kernel sequence of one umT5-XXL encode (24 layers, dim 4096, 64 heads x 64, ffn 10240) on `n`
// tokens, exactly what videocof_b200/text_encoder.py launches, timed with CUDA events on the launching stream.
// Weights are random and come in 4 distinct layer sets (1.5 GB, cycled) so no weight is served from L2 twice.
struct LayerW { bf16 *q, *k, *v, *o, *gate, *fc1, *fc2, *n1, *n2; };
static bf16* dev_random(size_t n, const bf16* seed_dev, size_t seed_n) {
  bf16* d = nullptr;
  if (cudaMalloc(&d, n * 2) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(2); }
  for (size_t off = 0; off < n; off += seed_n)
    cudaMemcpy(d + off, seed_dev, (n - off < seed_n ? n - off : seed_n) * 2, cudaMemcpyDeviceToDevice);
  return d;
}
static int bench(const char* out_path, int n, int narrow) {
  const int D = 4096, F = 10240, H = 64, LAYERS = 24, SETS = 4, L = n;
  const size_t seed_n = (size_t)8 << 20;
  auto seed = rand_bf16(seed_n, 1.0 / 64.0);
  bf16* dseed = to_dev(seed);
  LayerW w[SETS];
  auto ones = std::vector<bf16>(D, __float2bfloat16_rn(1.0f));
  for (auto& l : w) {
    l.q = dev_random((size_t)D * D, dseed, seed_n); l.k = dev_random((size_t)D * D, dseed + 4099, seed_n - 4099);
    l.v = dev_random((size_t)D * D, dseed + 131, seed_n - 131); l.o = dev_random((size_t)D * D, dseed + 977, seed_n - 977);
    l.gate = dev_random((size_t)F * D, dseed + 17, seed_n - 17); l.fc1 = dev_random((size_t)F * D, dseed + 3001, seed_n - 3001);
    l.fc2 = dev_random((size_t)D * F, dseed + 555, seed_n - 555);
    l.n1 = to_dev(ones); l.n2 = to_dev(ones);
  }
  auto tab = rand_bf16((size_t)1024 * D, 1.0);
  bf16* dtab = to_dev(tab);
  std::vector<long long> ids(n);
  for (auto& i : ids) i = (long long)(urand() * 1024) % 1024;
  long long* dids = to_dev(ids);
  std::vector<float> bias((size_t)H * (2 * L - 1));
  for (auto& b : bias) b = (float)nrand();
  float* dbias = to_dev(bias);
  bf16 *x, *h, *q, *k, *v, *a, *g, *y;
  cudaMalloc(&x, (size_t)n * D * 2); cudaMalloc(&h, (size_t)n * D * 2); cudaMalloc(&q, (size_t)n * D * 2);
  cudaMalloc(&k, (size_t)n * D * 2); cudaMalloc(&v, (size_t)n * D * 2); cudaMalloc(&a, (size_t)n * D * 2);
  cudaMalloc(&g, (size_t)n * F * 2); cudaMalloc(&y, (size_t)n * D * 2);
  const int fl = narrow ? VCOF_GEMM_TILE128 : 0;
  enum { K_NORM, K_QKV, K_ATTN, K_O, K_GATE, K_FC1, K_FC2, K_EMBED, K_N };
  const char* names[K_N] = {"t5_rmsnorm", "gemm_qkv", "t5_attn", "gemm_o_add", "gemm_gate_gelu", "gemm_fc1_mul",
                            "gemm_fc2_add", "embed_rows"};
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_kind;
  int rc = 0, launches = 0;
  auto pass = [&](bool per_kernel) {
    auto mark = [&](int kind) {
      if (!per_kernel) return;
      cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, 0); ev.push_back(e); ev_kind.push_back(kind);
    };
    launches = 0;
    mark(-1);
    rc |= vcof_embed_rows(dids, dtab, D, 1024, x, D, n, D, nullptr); ++launches; mark(K_EMBED);
    for (int l = 0; l < LAYERS; ++l) {
      const LayerW& W = w[l % SETS];
      rc |= vcof_t5_rmsnorm(x, D, W.n1, h, D, n, D, 1e-6f, nullptr); mark(K_NORM);
      rc |= vcof_gemm_bf16(h, D, W.q, D, nullptr, nullptr, q, D, n, D, D, VCOF_EPI_BIAS_BF16 | fl, nullptr);
      rc |= vcof_gemm_bf16(h, D, W.k, D, nullptr, nullptr, k, D, n, D, D, VCOF_EPI_BIAS_BF16 | fl, nullptr);
      rc |= vcof_gemm_bf16(h, D, W.v, D, nullptr, nullptr, v, D, n, D, D, VCOF_EPI_BIAS_BF16 | fl, nullptr); mark(K_QKV);
      rc |= vcof_t5_attn(q, D, k, D, v, D, a, D, dbias, 2 * L - 1, nullptr, 1, L, H, D / H, nullptr); mark(K_ATTN);
      rc |= vcof_gemm_bf16(a, D, W.o, D, nullptr, nullptr, x, D, n, D, D, VCOF_EPI_ADD_BF16 | fl, nullptr); mark(K_O);
      rc |= vcof_t5_rmsnorm(x, D, W.n2, h, D, n, D, 1e-6f, nullptr); mark(K_NORM);
      rc |= vcof_gemm_bf16(h, D, W.gate, D, nullptr, nullptr, g, F, n, F, D, VCOF_EPI_BIAS_GELU_BF16 | fl, nullptr); mark(K_GATE);
      rc |= vcof_gemm_bf16(h, D, W.fc1, D, nullptr, nullptr, g, F, n, F, D, VCOF_EPI_MUL_BF16 | fl, nullptr); mark(K_FC1);
      rc |= vcof_gemm_bf16(g, F, W.fc2, F, nullptr, nullptr, x, D, n, D, F, VCOF_EPI_ADD_BF16 | fl, nullptr); mark(K_FC2);
      launches += 10;
    }
    rc |= vcof_t5_rmsnorm(x, D, w[0].n1, y, D, n, D, 1e-6f, nullptr); ++launches; mark(K_NORM);
  };
  for (int i = 0; i < 3; ++i) pass(false);
  if (rc || cudaDeviceSynchronize() != cudaSuccess) {
    printf("{\"bench\": \"failed\", \"rc\": %d, \"error\": \"%s\"}\n", rc, vcof_last_error());
    return 1;
  }
  const int REPS = 10;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < REPS; ++i) pass(false);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= REPS;
  pass(true);
  cudaDeviceSynchronize();
  double per[K_N] = {0};
  for (size_t i = 1; i < ev.size(); ++i) {
    float t = 0;
    cudaEventElapsedTime(&t, ev[i - 1], ev[i]);
    per[ev_kind[i]] += t;
  }
  std::vector<bf16> yh = to_host(y, (size_t)n * D);
  bool nan = false;
  for (auto& e : yh) { const float f = __bfloat162float(e); if (f != f) nan = true; }
  const double gemm_flops = (double)LAYERS * 2.0 * n * (4.0 * D * D + 3.0 * D * F);
  const double attn_flops = (double)LAYERS * 4.0 * n * (double)n * D;
  const double weight_bytes = (double)LAYERS * 2.0 * (4.0 * D * D + 3.0 * D * F);
  char line[1024];
  int o = snprintf(line, sizeof line, "{\"bench\": \"umt5_xxl_encode\", \"tokens\": %d, \"layers\": %d, \"tile128\": %d, "
                   "\"ms_per_encode\": %.4f, \"launches\": %d, \"gemm_tflop\": %.4f, \"attn_tflop\": %.4f, "
                   "\"achieved_tflops\": %.1f, \"weight_gb\": %.3f, \"weight_stream_gbs\": %.1f, \"nan\": %s, "
                   "\"per_kernel_ms\": {", n, LAYERS, narrow, ms, launches, gemm_flops / 1e12, attn_flops / 1e12,
                   (gemm_flops + attn_flops) / (ms * 1e-3) / 1e12, weight_bytes / 1e9, weight_bytes / (ms * 1e-3) / 1e9,
                   nan ? "true" : "false");
  for (int i = 0; i < K_N; ++i)
    o += snprintf(line + o, sizeof line - o, "%s\"%s\": %.4f", i ? ", " : "", names[i], per[i]);
  snprintf(line + o, sizeof line - o, "}}");
  puts(line);
  if (out_path) { FILE* f = fopen(out_path, "a"); if (f) { fputs(line, f); fputc('\n', f); fclose(f); } }
  return nan ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "--bench") == 0) {
    const char* out = argc > 2 ? argv[2] : nullptr;
    int bad = 0;
    bad |= bench(out, 512, 1);
    bad |= bench(out, 512, 0);
    bad |= bench(out, 128, 1);
    return bad;
  }
  const bool attn_only = argc > 1 && strcmp(argv[1], "--attn") == 0;   // attention cases + one encoder bench
  if (attn_only) { --argc; ++argv; }
  if (argc > 1) g_out = fopen(argv[1], "w");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { puts("{\"error\": \"no CUDA device\"}"); return 3; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("{\"device\": \"%s\", \"sm\": %d%d, \"abi\": %d}\n", prop.name, prop.major, prop.minor, vcof_abi_version());
  if (!attn_only) {
    case_embed(777, 1000, 256);
    case_rmsnorm(1, 64); case_rmsnorm(77, 4096); case_rmsnorm(1024, 256);
  }
  const int l2[2] = {96, 1}, l3[3] = {160, 13, 100}, l1[1] = {150}, l512[2] = {512, 77}, lall[2] = {0, 40};
  case_attn(2, 96, 4, 64, l2);
  case_attn(1, 1, 2, 64, nullptr);
  case_attn(3, 160, 4, 16, l3);
  case_attn(1, 200, 3, 32, l1);
  case_attn(1, 300, 2, 128, nullptr);
  case_attn(2, 512, 4, 64, l512);
  case_attn(1, 512, 8, 64, nullptr);
  case_attn(2, 77, 3, 64, lall);                 // one sample with every key masked (uniform attention)
  case_attn(1, 129, 2, 64, nullptr);             // ragged: one row in the last 16-row tile / 64-key block
  if (attn_only) {
    printf("{\"failed\": %d}\n", g_fail);
    if (g_out) { fprintf(g_out, "{\"failed\": %d}\n", g_fail); fclose(g_out); }
    const int bad = bench(argc > 1 ? argv[1] : nullptr, 512, 1);
    return (g_fail || bad) ? 1 : 0;
  }
  const int shapes[4][3] = {{192, 256, 256}, {200, 104, 64}, {333, 1024, 512}, {512, 4096, 512}};
  for (auto& sh : shapes) {
    case_gemm(sh[0], sh[1], sh[2], VCOF_EPI_MUL_BF16);
    case_gemm(sh[0], sh[1], sh[2], VCOF_EPI_ADD_BF16);
  }
  case_gemm(192, 512, 256, VCOF_EPI_BIAS_GELU_BF16);
  case_gemm(192, 256, 512, VCOF_EPI_BIAS_BF16);
  // 128-wide tiles over several tile columns (VCOF_GEMM_TILE128: what the text encoder asks for at M = 512)
  case_gemm(512, 4096, 256, VCOF_EPI_MUL_BF16, VCOF_GEMM_TILE128);
  case_gemm(512, 4096, 256, VCOF_EPI_ADD_BF16, VCOF_GEMM_TILE128);
  case_gemm(333, 1000, 512, VCOF_EPI_BIAS_GELU_BF16, VCOF_GEMM_TILE128);
  case_gemm(77, 2560, 256, VCOF_EPI_BIAS_BF16, VCOF_GEMM_TILE128);
  case_gemm(640, 304, 192, VCOF_EPI_ADD_BF16, VCOF_GEMM_TILE128);
  printf("{\"failed\": %d}\n", g_fail);
  if (g_out) { fprintf(g_out, "{\"failed\": %d}\n", g_fail); fclose(g_out); }
  return g_fail ? 1 : 0;
}
