// t5_sm100.cu — the text-encoder kernels that are not GEMMs: token-embedding gather,
// T5LayerNorm and the 64-head x 64-wide relative-position-bias attention of umT5
// (videox_fun/models/wan_text_encoder.py:45-57, 60-112, 199-247, 281-283).
//
// The encoder runs once per prompt on <= 512 tokens: ~4.7 TFLOP of bias-free Linears (the tcgen05 GEMM with its
// bf16 multiply / add epilogues) and ~0.1 TFLOP of attention over 64-wide heads.  Two attention kernels share one
// contract: a warp-level tensor-core kernel (mma.sync, the default: 83 us per layer at 512 tokens on a B200) and the
// first CUDA-core kernel (VCOF_T5_ATTN=simple, 394 us per layer), kept as the independently written cross-check.
// The position bias depends on (key - query) only, so it arrives as a [heads, 2L-1] table instead of the reference's
// [heads, L, L] tensor.
#include <math.h>
#include <stdlib.h>

#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

// ---------------------------------------------------------------------------
// token embedding gather: out[i, :] = table[ids[i], :]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
embed_rows_kernel(const long long* __restrict__ ids, const bf16* __restrict__ table, long long ldt, long long vocab,
                  bf16* __restrict__ out, long long ldo, int C) {
  const long long row = blockIdx.x;
  const long long id = ids[row];
  const bool ok = id >= 0 && id < vocab;          // host validates; an out-of-range id still never reads wild memory
  const uint4* src = reinterpret_cast<const uint4*>(table + (ok ? id : 0) * ldt);
  uint4* dst = reinterpret_cast<uint4*>(out + row * ldo);
  for (int i = threadIdx.x; i < (C >> 3); i += blockDim.x) dst[i] = ok ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------
// T5LayerNorm: y = bf16(w * bf16(x * rsqrt(mean(x^2) + eps)))  — the fp32 factor is NOT rounded (unlike WanRMSNorm)
// ---------------------------------------------------------------------------
constexpr int kT5NormThreads = 128;
constexpr int kT5NormMaxVec = 8;  // uint4 (8 bf16) per thread -> C <= 8192

__global__ void __launch_bounds__(kT5NormThreads)
t5_rmsnorm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ weight, bf16* __restrict__ y,
                  long long ldy, int C, float eps) {
  __shared__ float scratch[kT5NormThreads / 32];
  const long long row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
  const int nvec = C >> 3;
  uint4 v[kT5NormMaxVec];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kT5NormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kT5NormThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        sq += f.x * f.x + f.y * f.y;
      }
    }
  }
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kT5NormThreads / 32; ++w) tot += scratch[w];
  const float rs = rsqrtf(tot / float(C) + eps);
#pragma unroll
  for (int i = 0; i < kT5NormMaxVec; ++i) {
    const int idx = threadIdx.x + i * kT5NormThreads;
    if (idx < nvec) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight) + idx);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        const float2 w = __bfloat1622float2(wh[e]);
        o[e] = pack_bf16x2(w.x * bf16_round(f.x * rs), w.y * bf16_round(f.y * rs));
      }
      yr[idx] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------
// attention with relative position bias, no 1/sqrt(d) scaling, key mask
// ---------------------------------------------------------------------------
constexpr int kT5AttnWarps = 16;
constexpr int kT5AttnThreads = kT5AttnWarps * 32;
constexpr int kT5MaxL = 512;
constexpr int kT5KeysPerLane = kT5MaxL / 32;
// torch.finfo(torch.bfloat16).min: what masked_fill_ writes over the bias of a masked key (wan_text_encoder.py:98)
#define VCOF_BF16_MIN (-3.3895313892515355e38f)
// which attention kernel vcof_t5_attn runs when VCOF_T5_ATTN is unset (1: mma.sync, 0: CUDA cores)
#define VCOF_T5_ATTN_DEFAULT_MMA 1

struct T5AttnSmem {
  int k_off, v_off, p_off, q_off, bias_off, mask_off, total;
};

__host__ __device__ inline T5AttnSmem t5_attn_smem(int L, int D) {
  T5AttnSmem s;
  const int Lp = (L + 31) & ~31;
  int off = 0;
  s.v_off = off;    off += L * D * 2;                       // V rows, 16-byte aligned (D % 8 == 0)
  off = (off + 15) & ~15;
  s.k_off = off;    off += L * (D + 2) * 2;                 // K rows padded by one 32-bit word: conflict-free columns
  off = (off + 15) & ~15;
  s.p_off = off;    off += kT5AttnWarps * Lp * 4;           // per-warp probabilities (fp32 holding bf16 values)
  s.q_off = off;    off += kT5AttnWarps * D * 4;            // per-warp query row
  s.bias_off = off; off += (2 * L - 1) * 4;
  off = (off + 3) & ~3;
  s.mask_off = off; off += Lp;
  s.total = (off + 15) & ~15;
  return s;
}

template <int D>
__global__ void __launch_bounds__(kT5AttnThreads, 1)
t5_attn_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, long long ldk,
               const bf16* __restrict__ v, long long ldv, bf16* __restrict__ out, long long ldo,
               const float* __restrict__ bias_rel, int bias_ld, const int* __restrict__ key_mask, int L,
               int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t t5_smem[];
  const T5AttnSmem lay = t5_attn_smem(L, D);
  bf16* Vs = reinterpret_cast<bf16*>(t5_smem + lay.v_off);
  uint32_t* Ks = reinterpret_cast<uint32_t*>(t5_smem + lay.k_off);     // row pitch D/2 + 1 words
  float* Ps = reinterpret_cast<float*>(t5_smem + lay.p_off);
  float* Qs = reinterpret_cast<float*>(t5_smem + lay.q_off);
  float* Bs = reinterpret_cast<float*>(t5_smem + lay.bias_off);
  uint8_t* Ms = t5_smem + lay.mask_off;
  constexpr int KP = D / 2 + 1;
  const int Lp = (L + 31) & ~31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tok0 = (long long)b * L;

  // stage this head's K and V
  constexpr int VPR = D / 8;  // 16-byte vectors per row
  for (int i = tid; i < L * VPR; i += kT5AttnThreads) {
    const int row = i / VPR, cv = i - row * VPR;
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(k + (tok0 + row) * ldk + h * D) + cv);
    uint32_t* kd = Ks + row * KP + cv * 4;
    kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
    reinterpret_cast<uint4*>(Vs + row * D)[cv] = __ldg(reinterpret_cast<const uint4*>(v + (tok0 + row) * ldv + h * D) + cv);
  }
  // bias_rel[h][(j - i) + L - 1]
  for (int i = tid; i < 2 * L - 1; i += kT5AttnThreads) Bs[i] = __ldg(bias_rel + (long long)h * bias_ld + i);
  for (int j = tid; j < Lp; j += kT5AttnThreads)
    Ms[j] = (j < L) ? ((key_mask == nullptr || key_mask[tok0 + j] != 0) ? 1 : 0) : 0;
  __syncthreads();

  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(L, r_begin + rows_per_cta);
  float* myP = Ps + warp * Lp;
  float* myQ = Qs + warp * D;
  for (int r = r_begin + warp; r < r_end; r += kT5AttnWarps) {
    // query row -> per-warp shared memory (fp32), read back as broadcasts
    const bf16* qr = q + (tok0 + r) * ldq + h * D;
    for (int c = lane; c < D; c += 32) myQ[c] = __bfloat162float(qr[c]);
    __syncwarp();

    float s[kT5KeysPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < kT5KeysPerLane; ++t) {
      const int j = lane + 32 * t;
      s[t] = -INFINITY;
      if (32 * t < L && j < L) {           // first test is warp-uniform: whole groups of 32 keys drop out
        const uint32_t* kr = Ks + j * KP;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < D / 2; ++c) {
          const uint32_t kw = kr[c];
          const float k0 = __uint_as_float(kw << 16), k1 = __uint_as_float(kw & 0xffff0000u);
          acc = fmaf(myQ[2 * c], k0, acc);
          acc = fmaf(myQ[2 * c + 1], k1, acc);
        }
        s[t] = acc + (Ms[j] ? Bs[j - r + L - 1] : VCOF_BF16_MIN);
        mx = fmaxf(mx, s[t]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kT5KeysPerLane; ++t) {
      s[t] = (s[t] == -INFINITY) ? 0.f : expf(s[t] - mx);
      sum += s[t];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < kT5KeysPerLane; ++t) {
      const int j = lane + 32 * t;
      if (j < Lp) myP[j] = bf16_round(s[t] * inv);   // the reference casts the probabilities to bf16 (:103)
    }
    __syncwarp();

    bf16* orow = out + (tok0 + r) * ldo + h * D;
    for (int c = lane; c < D; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < L; ++j) acc = fmaf(myP[j], __bfloat162float(Vs[j * D + c]), acc);
      orow[c] = __float2bfloat16_rn(acc);
    }
    __syncwarp();
  }
}

template <int D>
static int launch_t5_attn(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                          void* out, long long ldo, const float* bias_rel, int bias_ld, const int* key_mask, int B,
                          int L, int heads, cudaStream_t st) {
  const T5AttnSmem lay = t5_attn_smem(L, D);
  auto kern = t5_attn_kernel<D>;
  static int attr_bytes = 0;
  if (lay.total > attr_bytes) {
    VCOF_REQUIRE(lay.total <= 227 * 1024, "vcof_t5_attn: L=%d needs %d B of shared memory", L, lay.total);
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total));
    attr_bytes = lay.total;
  }
  // enough row blocks to fill the machine, but never fewer than 2 rows per warp: K/V staging is paid per CTA
  int rows_per_cta = 128;
  while (rows_per_cta > 2 * kT5AttnWarps &&
         (long long)((L + rows_per_cta - 1) / rows_per_cta) * heads * B < 2 * sm_count())
    rows_per_cta >>= 1;
  dim3 grid((L + rows_per_cta - 1) / rows_per_cta, heads, B);
  kern<<<grid, kT5AttnThreads, lay.total, st>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), ldk, reinterpret_cast<const bf16*>(v),
      ldv, reinterpret_cast<bf16*>(out), ldo, bias_rel, bias_ld, key_mask, L, rows_per_cta);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// the same attention on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32)
// ---------------------------------------------------------------------------
// 64-wide heads over <= 512 keys are too small for the tcgen05 pipeline of attn_sm100.cu (128-row tiles, 128-wide
// heads) but far too much arithmetic for CUDA cores (measured 0.39 ms per layer, 57 % of an encode,
// profiles/r1_gpurun38_bench_t5_native.jsonl).  A warp owns 16 query rows; K sits in shared memory row-major (it is
// already the "col" B operand of Q.K^T), V transposed (the "col" B operand of P.V).  Two passes over the keys instead
// of an online rescale: pass 1 finds each row's max and sum, pass 2 recomputes the scores (tensor-core time is
// free here) and feeds P = bf16(exp(s - max) / sum) — the reference's rounding point (:103) — straight from the
// accumulator registers into the P.V MMA (the C fragment of two adjacent 8-key tiles is the A fragment of one
// 16-key step).
constexpr int kT5MmaWarps = 8;
constexpr int kT5MmaThreads = kT5MmaWarps * 32;
constexpr int kT5MmaRows = kT5MmaWarps * 16;

struct T5MmaSmem {
  int k_off, v_off, bias_off, mask_off, total, Lp;
};

__host__ __device__ inline T5MmaSmem t5_mma_smem(int L, int D) {
  T5MmaSmem s;
  s.Lp = (L + 63) & ~63;
  int off = 0;
  s.k_off = off;    off += s.Lp * (D + 8) * 2;      // K[key][D + 8]: pitch = 4 words mod 32 -> conflict-free fragments
  s.v_off = off;    off += D * (s.Lp + 8) * 2;      // V^T[d][Lp + 8]: same property (Lp % 64 == 0)
  s.bias_off = off; off += (2 * L - 1) * 4;
  off = (off + 3) & ~3;
  s.mask_off = off; off += s.Lp;
  s.total = (off + 15) & ~15;
  return s;
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
__global__ void __launch_bounds__(kT5MmaThreads, 1)
t5_attn_mma_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, long long ldk,
                   const bf16* __restrict__ v, long long ldv, bf16* __restrict__ out, long long ldo,
                   const float* __restrict__ bias_rel, int bias_ld, const int* __restrict__ key_mask, int L) {
  extern __shared__ __align__(16) uint8_t t5m_smem[];
  const T5MmaSmem lay = t5_mma_smem(L, D);
  constexpr int KPB = D + 8;            // K pitch (bf16)
  const int Lp = lay.Lp, VPB = Lp + 8;  // V^T pitch (bf16)
  bf16* Ks = reinterpret_cast<bf16*>(t5m_smem + lay.k_off);
  bf16* Vt = reinterpret_cast<bf16*>(t5m_smem + lay.v_off);
  float* Bs = reinterpret_cast<float*>(t5m_smem + lay.bias_off);
  uint8_t* Ms = t5m_smem + lay.mask_off;
  const int h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tok0 = (long long)b * L;
  constexpr int VPR = D / 8;

  // K rows (zero beyond L), 16 bytes per thread
  for (int i = tid; i < Lp * VPR; i += kT5MmaThreads) {
    const int row = i / VPR, cv = i - row * VPR;
    uint4 kv = make_uint4(0, 0, 0, 0);
    if (row < L) kv = __ldg(reinterpret_cast<const uint4*>(k + (tok0 + row) * ldk + h * D) + cv);
    *reinterpret_cast<uint4*>(Ks + row * KPB + cv * 8) = kv;
  }
  // V transposed (zero beyond L): consecutive lanes take consecutive keys of one 8-channel group, so the 2-byte
  // shared-memory stores of a warp are contiguous
  for (int i = tid; i < Lp * VPR; i += kT5MmaThreads) {
    const int cv = i / Lp, key = i - cv * Lp;
    uint4 vv = make_uint4(0, 0, 0, 0);
    if (key < L) vv = __ldg(reinterpret_cast<const uint4*>(v + (tok0 + key) * ldv + h * D) + cv);
    const bf16* e = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) Vt[(cv * 8 + j) * VPB + key] = e[j];
  }
  for (int i = tid; i < 2 * L - 1; i += kT5MmaThreads) Bs[i] = __ldg(bias_rel + (long long)h * bias_ld + i);
  for (int j = tid; j < Lp; j += kT5MmaThreads)
    Ms[j] = (j < L && (key_mask == nullptr || key_mask[tok0 + j] != 0)) ? 1 : 0;
  __syncthreads();

  const int row0 = blockIdx.x * kT5MmaRows + warp * 16;
  if (row0 >= L) return;
  const int g = lane >> 2, t = lane & 3;
  const int rA = row0 + g, rB = rA + 8;
  const int iA = min(rA, L - 1), iB = min(rB, L - 1);    // bias row (rows beyond L are computed but never stored)

  // Q fragments straight from global memory: [k-step][4]
  uint32_t qf[D / 16][4];
  {
    const bf16* qa = q + (tok0 + iA) * ldq + h * D + 2 * t;
    const bf16* qb = q + (tok0 + iB) * ldq + h * D + 2 * t;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      qf[ks][0] = *reinterpret_cast<const uint32_t*>(qa + ks * 16);
      qf[ks][1] = *reinterpret_cast<const uint32_t*>(qb + ks * 16);
      qf[ks][2] = *reinterpret_cast<const uint32_t*>(qa + ks * 16 + 8);
      qf[ks][3] = *reinterpret_cast<const uint32_t*>(qb + ks * 16 + 8);
    }
  }
  const uint32_t* Kw = reinterpret_cast<const uint32_t*>(Ks);
  const uint32_t* Vw = reinterpret_cast<const uint32_t*>(Vt);
  const int nblk = Lp >> 6;

  // scores of one 64-key block for this warp's 16 rows: s[nt][0..1] row rA, s[nt][2..3] row rB, keys kb*64+nt*8+2t(+1)
  auto scores = [&](int kb, float (&s)[8][4]) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const int n = kb * 64 + nt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t b0 = Kw[(n * KPB + ks * 16 + 2 * t) >> 1];
        const uint32_t b1 = Kw[(n * KPB + ks * 16 + 8 + 2 * t) >> 1];
        mma_bf16_16816(s[nt], qf[ks], b0, b1);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = kb * 64 + nt * 8 + 2 * t + e;
        if (j < L) {
          const bool keep = Ms[j] != 0;
          s[nt][e] += keep ? Bs[j - iA + L - 1] : VCOF_BF16_MIN;
          s[nt][2 + e] += keep ? Bs[j - iB + L - 1] : VCOF_BF16_MIN;
        } else {
          s[nt][e] = -INFINITY;
          s[nt][2 + e] = -INFINITY;
        }
      }
    }
  };

  // pass 1: row max and sum of exponentials
  float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;
  for (int kb = 0; kb < nblk; ++kb) {
    float s[8][4];
    scores(kb, s);
    float xa = -INFINITY, xb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      xa = fmaxf(xa, fmaxf(s[nt][0], s[nt][1]));
      xb = fmaxf(xb, fmaxf(s[nt][2], s[nt][3]));
    }
    xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 1)); xa = fmaxf(xa, __shfl_xor_sync(0xffffffffu, xa, 2));
    xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 1)); xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, 2));
    const float nA = fmaxf(mA, xa), nB = fmaxf(mB, xb);   // finite: every block holds at least one key < L
    lA *= __expf(mA - nA);
    lB *= __expf(mB - nB);
    mA = nA; mB = nB;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      lA += __expf(s[nt][0] - mA) + __expf(s[nt][1] - mA);
      lB += __expf(s[nt][2] - mB) + __expf(s[nt][3] - mB);
    }
  }
  lA += __shfl_xor_sync(0xffffffffu, lA, 1); lA += __shfl_xor_sync(0xffffffffu, lA, 2);
  lB += __shfl_xor_sync(0xffffffffu, lB, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
  const float invA = 1.f / lA, invB = 1.f / lB;

  // pass 2: P = bf16(exp(s - max) / sum) from the accumulator registers into P.V
  float o[D / 8][4];
#pragma unroll
  for (int dn = 0; dn < D / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
  for (int kb = 0; kb < nblk; ++kb) {
    float s[8][4];
    scores(kb, s);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {   // 16 keys = two adjacent 8-key tiles
      uint32_t pf[4];
      pf[0] = pack_bf16x2(__expf(s[2 * kk][0] - mA) * invA, __expf(s[2 * kk][1] - mA) * invA);
      pf[1] = pack_bf16x2(__expf(s[2 * kk][2] - mB) * invB, __expf(s[2 * kk][3] - mB) * invB);
      pf[2] = pack_bf16x2(__expf(s[2 * kk + 1][0] - mA) * invA, __expf(s[2 * kk + 1][1] - mA) * invA);
      pf[3] = pack_bf16x2(__expf(s[2 * kk + 1][2] - mB) * invB, __expf(s[2 * kk + 1][3] - mB) * invB);
      const int kbase = kb * 64 + kk * 16;
#pragma unroll
      for (int dn = 0; dn < D / 8; ++dn) {
        const int n = dn * 8 + g;
        const uint32_t b0 = Vw[(n * VPB + kbase + 2 * t) >> 1];
        const uint32_t b1 = Vw[(n * VPB + kbase + 8 + 2 * t) >> 1];
        mma_bf16_16816(o[dn], pf, b0, b1);
      }
    }
  }
  bf16* oa = out + (tok0 + rA) * ldo + h * D + 2 * t;
  bf16* ob = out + (tok0 + rB) * ldo + h * D + 2 * t;
#pragma unroll
  for (int dn = 0; dn < D / 8; ++dn) {
    if (rA < L) *reinterpret_cast<uint32_t*>(oa + dn * 8) = pack_bf16x2(o[dn][0], o[dn][1]);
    if (rB < L) *reinterpret_cast<uint32_t*>(ob + dn * 8) = pack_bf16x2(o[dn][2], o[dn][3]);
  }
}

template <int D>
static int launch_t5_attn_mma(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                              void* out, long long ldo, const float* bias_rel, int bias_ld, const int* key_mask, int B,
                              int L, int heads, cudaStream_t st) {
  const T5MmaSmem lay = t5_mma_smem(L, D);
  auto kern = t5_attn_mma_kernel<D>;
  static int attr_bytes = 0;
  if (lay.total > attr_bytes) {
    VCOF_REQUIRE(lay.total <= 227 * 1024, "vcof_t5_attn: L=%d, head_dim=%d need %d B of shared memory", L, D, lay.total);
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total));
    attr_bytes = lay.total;
  }
  dim3 grid((L + kT5MmaRows - 1) / kT5MmaRows, heads, B);
  kern<<<grid, kT5MmaThreads, lay.total, st>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), ldk, reinterpret_cast<const bf16*>(v),
      ldv, reinterpret_cast<bf16*>(out), ldo, bias_rel, bias_ld, key_mask, L);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vcof

using namespace vcof;

extern "C" int vcof_embed_rows(const long long* ids, const void* table, long long ldt, long long vocab, void* out,
                               long long ldo, long long n, int C, void* stream) {
  VCOF_REQUIRE(n > 0 && C > 0 && vocab > 0, "vcof_embed_rows: empty problem");
  VCOF_REQUIRE(C % 8 == 0 && ldt % 8 == 0 && ldo % 8 == 0 && ldt >= C && ldo >= C,
               "vcof_embed_rows: C=%d, ldt, ldo must be multiples of 8", C);
  VCOF_REQUIRE(n <= 0x7fffffffLL, "vcof_embed_rows: too many rows");
  embed_rows_kernel<<<(unsigned)n, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ids, reinterpret_cast<const bf16*>(table), ldt, vocab, reinterpret_cast<bf16*>(out), ldo, C);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_t5_rmsnorm(const void* x, long long ldx, const void* weight, void* y, long long ldy, long long rows,
                               int C, float eps, void* stream) {
  VCOF_REQUIRE(rows > 0 && C > 0, "vcof_t5_rmsnorm: empty problem");
  VCOF_REQUIRE(C % 8 == 0 && C <= kT5NormThreads * kT5NormMaxVec * 8 && ldx % 8 == 0 && ldy % 8 == 0,
               "vcof_t5_rmsnorm: C=%d / ldx / ldy must be multiples of 8, C <= %d", C,
               kT5NormThreads * kT5NormMaxVec * 8);
  VCOF_REQUIRE(rows <= 0x7fffffffLL, "vcof_t5_rmsnorm: too many rows");
  t5_rmsnorm_kernel<<<(unsigned)rows, kT5NormThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(weight), reinterpret_cast<bf16*>(y), ldy,
      C, eps);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_t5_attn(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                            void* out, long long ldo, const float* bias_rel, int bias_ld, const int* key_mask, int B,
                            int L, int heads, int head_dim, void* stream) {
  VCOF_REQUIRE(B > 0 && L > 0 && heads > 0, "vcof_t5_attn: empty problem");
  VCOF_REQUIRE(L <= kT5MaxL, "vcof_t5_attn: L=%d exceeds %d tokens (max_sequence_length of the pipeline)", L, kT5MaxL);
  VCOF_REQUIRE(bias_rel != nullptr && bias_ld >= 2 * L - 1, "vcof_t5_attn: bias table needs 2L-1 = %d entries per head",
               2 * L - 1);
  VCOF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && B <= 65535 && heads <= 65535,
               "vcof_t5_attn: ldq/ldk/ldv must be multiples of 8 (16-byte rows)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // VCOF_T5_ATTN=mma|simple selects the tensor-core (mma.sync) or the CUDA-core kernel
  static const bool use_mma = [] {
    const char* e = getenv("VCOF_T5_ATTN");
    return e != nullptr ? (e[0] == 'm') : VCOF_T5_ATTN_DEFAULT_MMA;
  }();
  if (use_mma) {
    VCOF_REQUIRE(ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0, "vcof_t5_attn: out must be 4-byte aligned");
    switch (head_dim) {
      case 16: return launch_t5_attn_mma<16>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
      case 32: return launch_t5_attn_mma<32>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
      case 64: return launch_t5_attn_mma<64>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
      case 128: return launch_t5_attn_mma<128>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
      default: break;
    }
    set_last_error("vcof_t5_attn: head_dim %d not in {16, 32, 64, 128}", head_dim);
    return -1;
  }
  switch (head_dim) {
    case 16: return launch_t5_attn<16>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
    case 32: return launch_t5_attn<32>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
    case 64: return launch_t5_attn<64>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
    case 128: return launch_t5_attn<128>(q, ldq, k, ldk, v, ldv, out, ldo, bias_rel, bias_ld, key_mask, B, L, heads, st);
    default: break;
  }
  set_last_error("vcof_t5_attn: head_dim %d not in {16, 32, 64, 128}", head_dim);
  return -1;
}
