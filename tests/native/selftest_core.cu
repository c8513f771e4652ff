// selftest_core.cu — stand-alone parity check of the two tensor-core kernels of the denoising step over the C ABI
// (no Python): vcof_gemm_bf16 (bias / GELU / gated fp32 residual / fp32 store epilogues) and vcof_attn_fwd
// (head_dim 128, ragged lengths, kv_len < Lk, V^T input) against double-precision CPU restatements in this file.
// `frames`: the byte conversions vcof_u8_to_cl / vcof_cl_to_u8 against the library's host evaluation, bit-exact.
// Meant for kernel work: `VCOF_GEMM_2CTA=1 tests/native/selftest_core gemm`, `VCOF_ATTN_EMU=2 … attn` validate a
// variant in seconds of GPU time.  The Python suite (tests/test_kernels_gpu.py) stays the reference gate.
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
static double urand() {
  g_seed ^= g_seed >> 12; g_seed ^= g_seed << 25; g_seed ^= g_seed >> 27;
  return ((g_seed * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) + 1e-17;
}
static double nrand() { return sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand()); }
static float bfr(double x) { return __bfloat162float(__float2bfloat16_rn((float)x)); }
static std::vector<bf16> rand_bf16(size_t n, double scale) {
  std::vector<bf16> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = __float2bfloat16_rn((float)(scale * nrand()));
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
static int g_fail = 0;
static void report(const char* name, const char* shape, const std::vector<double>& got, const std::vector<double>& ref,
                   double tol, int rc) {
  double num = 0, den = 0;
  bool nan = false;
  for (size_t i = 0; i < ref.size(); ++i) {
    if (got[i] != got[i]) nan = true;
    num += (got[i] - ref[i]) * (got[i] - ref[i]);
    den += ref[i] * ref[i];
  }
  const double rel = sqrt(num / (den + 1e-300));
  const cudaError_t e = cudaDeviceSynchronize();
  const bool ok = rc == 0 && e == cudaSuccess && !nan && rel < tol;
  if (!ok) ++g_fail;
  printf("{\"case\": \"%s\", \"shape\": \"%s\", \"rc\": %d, \"cuda\": %d, \"rel_fro\": %.3e, \"tol\": %.1e, \"nan\": %s, "
         "\"ok\": %s%s%s%s}\n", name, shape, rc, (int)e, rel, tol, nan ? "true" : "false", ok ? "true" : "false",
         rc ? ", \"error\": \"" : "", rc ? vcof_last_error() : "", rc ? "\"" : "");
  fflush(stdout);
}
static double gelu_tanh(double x) { return 0.5 * x * (1.0 + tanh(0.7978845608028654 * (x + 0.044715 * x * x * x))); }

static void case_gemm(int M, int N, int K, int epi) {
  auto a = rand_bf16((size_t)M * K, 1.0), w = rand_bf16((size_t)N * K, 1.0 / sqrt((double)K)), bias = rand_bf16(N, 1.0);
  std::vector<float> gate(N), x0((size_t)M * N);
  for (auto& g : gate) g = (float)nrand();
  for (auto& v : x0) v = (float)nrand();
  bf16 *da = to_dev(a), *dw = to_dev(w), *db = to_dev(bias);
  float* dg = to_dev(gate);
  const bool f32 = epi == VCOF_EPI_BIAS_GATE_RES_F32 || epi == VCOF_EPI_BIAS_F32;
  void* dout = nullptr;
  if (f32) dout = to_dev(x0); else cudaMalloc(&dout, (size_t)M * N * 2);
  const int rc = vcof_gemm_bf16(da, K, dw, K, db, epi == VCOF_EPI_BIAS_GATE_RES_F32 ? dg : nullptr, dout, N, M, N, K, epi, nullptr);
  std::vector<double> got((size_t)M * N), ref((size_t)M * N);
  if (f32) { auto h = to_host((float*)dout, (size_t)M * N); for (size_t i = 0; i < h.size(); ++i) got[i] = h[i]; }
  else { auto h = to_host((bf16*)dout, (size_t)M * N); for (size_t i = 0; i < h.size(); ++i) got[i] = __bfloat162float(h[i]); }
  std::vector<float> af((size_t)M * K), wf((size_t)N * K);
  for (size_t i = 0; i < af.size(); ++i) af[i] = __bfloat162float(a[i]);
  for (size_t i = 0; i < wf.size(); ++i) wf[i] = __bfloat162float(w[i]);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = __bfloat162float(bias[n]);
      const float *ar = &af[(size_t)m * K], *wr = &wf[(size_t)n * K];
      for (int k = 0; k < K; ++k) acc += (double)ar[k] * wr[k];
      const size_t i = (size_t)m * N + n;
      ref[i] = epi == VCOF_EPI_BIAS_BF16 ? acc : epi == VCOF_EPI_BIAS_GELU_BF16 ? gelu_tanh(bfr(acc))
               : epi == VCOF_EPI_BIAS_GATE_RES_F32 ? x0[i] + (double)gate[n] * bfr(acc) : (double)bfr(acc);
    }
  char shape[64]; snprintf(shape, sizeof shape, "M=%d N=%d K=%d epi=%d", M, N, K, epi);
  report("gemm", shape, got, ref, 6e-3, rc);
  cudaFree(da); cudaFree(dw); cudaFree(db); cudaFree(dg); cudaFree(dout);
}

// grow > 0: the keys of the b-th block of 128 are scaled by (1 + grow * b), so the running row maximum jumps by far more
// than the lazy-rescale threshold (2^8) at every key block and the rescale path of the kernel runs at each step.
static void case_attn(int Lq, int Lk, int kv, int heads, int vt, double grow = 0.0) {
  const int C = heads * 128;
  auto q = rand_bf16((size_t)Lq * C, 1.0), k = rand_bf16((size_t)Lk * C, 1.0), v = rand_bf16((size_t)Lk * C, 1.0);
  if (grow > 0.0)
    for (int j = 0; j < Lk; ++j)
      for (int c = 0; c < C; ++c)
        k[(size_t)j * C + c] = __float2bfloat16_rn(__bfloat162float(k[(size_t)j * C + c]) * (float)(1.0 + grow * (j / 128)));
  bf16 *dq = to_dev(q), *dk = to_dev(k), *dv = nullptr, *dout = nullptr;
  long long ldv = C;
  if (vt) {                                         // V^T [C, ldv] with kv contiguous
    ldv = (Lk + 7) / 8 * 8;
    std::vector<bf16> t((size_t)C * ldv, __float2bfloat16_rn(0.f));
    for (int j = 0; j < Lk; ++j) for (int c = 0; c < C; ++c) t[(size_t)c * ldv + j] = v[(size_t)j * C + c];
    dv = to_dev(t);
  } else {
    dv = to_dev(v);
  }
  cudaMalloc(&dout, (size_t)Lq * C * 2);
  const int rc = vcof_attn_fwd(dq, C, dk, C, dv, ldv, dout, C, Lq, Lk, kv, heads, 128, (float)(1.0 / sqrt(128.0)), vt, nullptr);
  auto h = to_host(dout, (size_t)Lq * C);
  std::vector<double> got((size_t)Lq * C), ref((size_t)Lq * C), s(kv);
  for (size_t i = 0; i < h.size(); ++i) got[i] = __bfloat162float(h[i]);
  for (int hd = 0; hd < heads; ++hd)
    for (int i = 0; i < Lq; ++i) {
      double mx = -INFINITY, sum = 0;
      for (int j = 0; j < kv; ++j) {
        double acc = 0;
        for (int c = 0; c < 128; ++c)
          acc += (double)__bfloat162float(q[(size_t)i * C + hd * 128 + c]) * __bfloat162float(k[(size_t)j * C + hd * 128 + c]);
        s[j] = acc / sqrt(128.0);
        mx = fmax(mx, s[j]);
      }
      for (int j = 0; j < kv; ++j) { s[j] = exp(s[j] - mx); sum += s[j]; }
      for (int c = 0; c < 128; ++c) {
        double acc = 0;
        for (int j = 0; j < kv; ++j) acc += s[j] * __bfloat162float(v[(size_t)j * C + hd * 128 + c]);
        ref[(size_t)i * C + hd * 128 + c] = acc / sum;
      }
    }
  char shape[96]; snprintf(shape, sizeof shape, "Lq=%d Lk=%d kv=%d heads=%d vt=%d grow=%g", Lq, Lk, kv, heads, vt, grow);
  report("attn", shape, got, ref, 1.5e-2, rc);
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout);
}

// Frame bytes (vcof_u8_to_cl / vcof_cl_to_u8): the device kernels against the library's HOST evaluation of the same
// per-element functions (pinned exhaustively to the executed reference by tests/test_video_io_cpu.py).  Bit-exact.
static void report_bytes(const char* name, const char* shape, size_t mismatches, int rc) {
  const cudaError_t e = cudaDeviceSynchronize();
  const bool ok = rc == 0 && e == cudaSuccess && mismatches == 0;
  if (!ok) ++g_fail;
  printf("{\"case\": \"%s\", \"shape\": \"%s\", \"rc\": %d, \"cuda\": %d, \"mismatches\": %zu, \"ok\": %s%s%s%s}\n", name,
         shape, rc, (int)e, mismatches, ok ? "true" : "false", rc ? ", \"error\": \"" : "", rc ? vcof_last_error() : "",
         rc ? "\"" : "");
  fflush(stdout);
}

static void case_frames_out(long long npos, int C, int ld, int byte_offset) {
  // every 16-bit pattern cycles through the real channels; the padding channels hold junk
  std::vector<unsigned short> x((size_t)npos * ld + 8), xr((size_t)npos * C);
  unsigned short pat = 0;
  for (long long p = 0; p < npos; ++p)
    for (int c = 0; c < ld; ++c) {
      const unsigned short v = (unsigned short)(pat += 257);
      x[(size_t)p * ld + c + byte_offset / 2] = v;
      if (c < C) xr[(size_t)p * C + c] = v;
    }
  std::vector<unsigned char> ref((size_t)npos * C);
  vcof_debug_frame_u8_host(xr.data(), ref.data(), npos * C);
  unsigned short* dx = to_dev(x);
  unsigned char* dout = nullptr;
  cudaMalloc(&dout, (size_t)npos * C + 16);
  cudaMemset(dout, 0xAB, (size_t)npos * C + 16);
  const int rc = vcof_cl_to_u8(reinterpret_cast<char*>(dx) + byte_offset, ld, dout, npos, C, nullptr);
  std::vector<unsigned char> got = to_host(dout, (size_t)npos * C + 16);
  size_t bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) bad += got[i] != ref[i];
  for (size_t i = ref.size(); i < got.size(); ++i) bad += got[i] != 0xAB;          // nothing written past the end
  char shape[96];
  snprintf(shape, sizeof shape, "npos=%lld C=%d ld=%d offset=%dB", npos, C, ld, byte_offset);
  report_bytes("cl_to_u8", shape, bad, rc);
  cudaFree(dx); cudaFree(dout);
}

static void case_frames_in(long long npos, int C, int Cp) {
  std::vector<unsigned char> f((size_t)npos * C);
  for (size_t i = 0; i < f.size(); ++i) f[i] = (unsigned char)(i * 37 + (i >> 8));
  std::vector<unsigned short> conv(f.size());
  vcof_debug_video_bf16_host(f.data(), conv.data(), (long long)f.size());
  unsigned char* df = to_dev(f);
  unsigned short* dy = nullptr;
  cudaMalloc(&dy, ((size_t)npos * Cp + 8) * 2);
  cudaMemset(dy, 0xCD, ((size_t)npos * Cp + 8) * 2);
  const int rc = vcof_u8_to_cl(df, dy, npos, C, Cp, nullptr);
  std::vector<unsigned short> got = to_host(dy, (size_t)npos * Cp + 8);
  size_t bad = 0;
  for (long long p = 0; p < npos; ++p)
    for (int c = 0; c < Cp; ++c)
      bad += got[(size_t)p * Cp + c] != (c < C ? conv[(size_t)p * C + c] : (unsigned short)0);
  for (size_t i = (size_t)npos * Cp; i < got.size(); ++i) bad += got[i] != 0xCDCD;
  char shape[96];
  snprintf(shape, sizeof shape, "npos=%lld C=%d Cp=%d", npos, C, Cp);
  report_bytes("u8_to_cl", shape, bad, rc);
  cudaFree(df); cudaFree(dy);
}

int main(int argc, char** argv) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { puts("{\"error\": \"no CUDA device\"}"); return 3; }
  const bool all = argc < 2;
  if (all || strcmp(argv[1], "gemm") == 0) {
    const int shapes[5][3] = {{128, 256, 64}, {300, 512, 256}, {640, 1536, 512}, {777, 1024, 320}, {512, 64, 1024}};
    for (auto& sh : shapes)
      for (int epi : {VCOF_EPI_BIAS_BF16, VCOF_EPI_BIAS_GELU_BF16, VCOF_EPI_BIAS_GATE_RES_F32, VCOF_EPI_BIAS_F32})
        case_gemm(sh[0], sh[1], sh[2], epi);
    case_gemm(200, 104, 64, VCOF_EPI_BIAS_BF16);
    case_gemm(200, 304, 192, VCOF_EPI_BIAS_GATE_RES_F32);
    case_gemm(2048, 2560, 512, VCOF_EPI_BIAS_BF16);      // several waves of tiles per CTA
  }
  if (all || strcmp(argv[1], "attn") == 0) {
    for (int vt = 0; vt < 2; ++vt) {
      case_attn(128, 128, 128, 1, vt);
      case_attn(300, 336, 336, 2, vt);
      case_attn(512, 640, 600, 2, vt);
      case_attn(1280, 1280, 1280, 2, vt);
    }
    case_attn(300, 333, 333, 2, 0);
    case_attn(300, 336, 300, 2, 0);                      // last key block holds 44 keys: masking reaches the first half
    case_attn(256, 1024, 1000, 2, 0, 6.0);               // row maximum jumps at every key block: rescale path
    case_attn(200, 700, 700, 1, 1, 3.0);
    case_attn(1024, 128, 128, 40, 0);                    // 160 work items > 148 SMs: persistent CTAs walk several items
  }
  if (all || strcmp(argv[1], "frames") == 0) {
    case_frames_out(65536, 3, 8, 0);          // vector path: every bf16 pattern
    case_frames_out(1027, 3, 8, 0);           // vector path + 3-position scalar tail
    case_frames_out(3, 3, 8, 0);              // fewer than four positions
    case_frames_out(513, 3, 8, 8);            // rows 8 bytes off a 16-byte boundary: generic path
    case_frames_out(70000, 3, 3, 0);          // dense RGB input
    case_frames_out(4099, 4, 32, 0);          // generic shape
    case_frames_out(9 * 720 * 1280, 3, 8, 0); // nine 720p frames
    case_frames_in(1, 3, 32);
    case_frames_in(1000, 3, 32);              // the encoder's shape (4 groups per row)
    case_frames_in(1000, 3, 8);
    case_frames_in(777, 8, 8);
    case_frames_in(333, 1, 3);                // generic path
    case_frames_in(500, 4, 12);
    case_frames_in(9 * 720 * 1280, 3, 32);    // nine 720p frames
  }
  printf("{\"failed\": %d}\n", g_fail);
  return g_fail ? 1 : 0;
}
