// gemm_sm100.cu — persistent, warp-specialised bf16 GEMM on tcgen05 for sm_100a.
//
//   D[M,N] = A[M,K] * W[N,K]^T  (+ bias[N]) with a fused epilogue
//
// A is an activation matrix (tokens x channels, row-major) and W an nn.Linear
// weight (out x in, row-major), so both operands are K-major: the canonical
// tcgen05 case.  One CTA per SM loops over 128 x BN output tiles:
//   warp 4   : TMA producer  (A tile 128x64, W tile BNx64, 128B swizzle, mbarrier ring)
//   warp 5   : tcgen05.mma issuer (one elected lane), fp32 accumulators in TMEM,
//              double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 0-3: epilogue (tcgen05.ld 32x32b -> registers -> fused op -> global)
//
// Replaces the cuBLAS calls behind every nn.Linear on the DiT path
// (reference videox_fun/models/wan_transformer3d.py:264-267, 457-459, 543) and fuses
// what the reference runs as separate ATen kernels afterwards: bias, GELU(tanh)
// (:458), the AdaLN gate and fp32 residual accumulate (:499, :504, :511).
#include <cstdlib>
#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 256;   // warps 0-3 epilogue, 4 TMA (A, or A+B), 5 MMA, 6 TMA (B) when two producer lanes, 7 idle

struct GemmArgs {
  int M, N, K;
  const bf16* bias;   // [N] or nullptr
  const float* gate;  // [N] fp32 or nullptr (EPI 2)
  void* out;          // bf16 or fp32 [M, ldo]
  long long ldo;
  int group_m;
  int producers;      // 1: one lane issues both operand boxes per stage; 2: A and B boxes come from lanes of different
                      // warps (one lane sustains ~1 barrier round trip per ~520 clk, profiles/r1_tma_probe_v2.txt)
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +align slack
  static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk,
                                            int& n_blk) {
  int per_group = group_m * num_n;
  int g = tile / per_group;
  int first_m = g * group_m;
  int gsz = min(group_m, num_m - first_m);
  int r = tile - g * per_group;
  m_blk = first_m + r % gsz;
  n_blk = r / gsz;
}

// One accumulator tile (this thread's row of 128 lanes x BN columns at t_row) -> epilogue -> global memory.
template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& p, uint32_t t_row, int row, int n_blk) {
  const bool row_ok = row < p.M;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n0 = n_blk * BN + c * 32;
      if (n0 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld32(t_row + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const bool full_chunk = (n0 + 32 <= p.N);
      if (p.bias != nullptr) {
        if (full_chunk) {
          const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 b = __ldg(bp + q);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __bfloat1622float2(b2[e]);
              v[q * 8 + e * 2] += f.x;
              v[q * 8 + e * 2 + 1] += f.y;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) v[j] += __bfloat162float(p.bias[n0 + j]);
        }
      }
      if (EPI == VCOF_EPI_BIAS_GELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(bf16_round(v[j]));
      }
      if (!row_ok) continue;
      if (EPI == VCOF_EPI_BIAS_BF16 || EPI == VCOF_EPI_BIAS_GELU_BF16) {
        bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + n0;
        if (full_chunk) {
          uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            w.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            w.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            w.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            o4[q] = w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) o[j] = __float2bfloat16_rn(v[j]);
        }
      } else if (EPI == VCOF_EPI_BIAS_GATE_RES_F32) {
        // x[row, n] += gate[n] * bf16(acc + bias)   (reference :499 / :504 / :511:
        // the Linear output is a bf16 tensor, the residual stream is fp32)
        float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + n0;
        if (full_chunk) {
          float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 x = o4[q];
            float g0 = 1.f, g1 = 1.f, g2 = 1.f, g3 = 1.f;
            if (p.gate != nullptr) {
              float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + n0) + q);
              g0 = g.x; g1 = g.y; g2 = g.z; g3 = g.w;
            }
            x.x += g0 * bf16_round(v[q * 4 + 0]);
            x.y += g1 * bf16_round(v[q * 4 + 1]);
            x.z += g2 * bf16_round(v[q * 4 + 2]);
            x.w += g3 * bf16_round(v[q * 4 + 3]);
            o4[q] = x;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) {
              float g = p.gate ? p.gate[n0 + j] : 1.f;
              o[j] += g * bf16_round(v[j]);
            }
        }
      } else if (EPI == VCOF_EPI_GATE_ACCUM_BF16) {
        // W[row, n] = bf16(float(W[row, n]) + gate[n] * acc): in-place LoRA merge (lora_utils.py:496: the weight is
        // lifted to fp32, updated and cast back)
        bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + n0;
        if (full_chunk) {
          uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 w = o4[q];
            const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&w);
            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gate + n0) + q * 2);
            const float4 gb = __ldg(reinterpret_cast<const float4*>(p.gate + n0) + q * 2 + 1);
            const float2 f0 = __bfloat1622float2(w2[0]), f1 = __bfloat1622float2(w2[1]);
            const float2 f2 = __bfloat1622float2(w2[2]), f3 = __bfloat1622float2(w2[3]);
            uint4 r;
            r.x = pack_bf16x2(f0.x + ga.x * v[q * 8 + 0], f0.y + ga.y * v[q * 8 + 1]);
            r.y = pack_bf16x2(f1.x + ga.z * v[q * 8 + 2], f1.y + ga.w * v[q * 8 + 3]);
            r.z = pack_bf16x2(f2.x + gb.x * v[q * 8 + 4], f2.y + gb.y * v[q * 8 + 5]);
            r.w = pack_bf16x2(f3.x + gb.z * v[q * 8 + 6], f3.y + gb.w * v[q * 8 + 7]);
            o4[q] = r;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) o[j] = __float2bfloat16_rn(__bfloat162float(o[j]) + p.gate[n0 + j] * v[j]);
        }
      } else if (EPI == VCOF_EPI_MUL_BF16 || EPI == VCOF_EPI_ADD_BF16) {
        // bf16 read-modify-write with the Linear output rounded to bf16 first, as the reference's eager bf16 ops do:
        // MUL: u = fc1(x) * gelu(gate(x)) with `out` holding the gate branch (wan_text_encoder.py:129);
        // ADD: x = x + o(attn) / x + fc2(u) on the bf16 residual stream (wan_text_encoder.py:156-157)
        bf16* o = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + n0;
        if (full_chunk) {
          uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 w = o4[q];
            const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&w);
            uint32_t r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(w2[e]);
              const float a0 = bf16_round(v[q * 8 + e * 2]), a1 = bf16_round(v[q * 8 + e * 2 + 1]);
              r[e] = (EPI == VCOF_EPI_MUL_BF16) ? pack_bf16x2(f.x * a0, f.y * a1) : pack_bf16x2(f.x + a0, f.y + a1);
            }
            o4[q] = make_uint4(r[0], r[1], r[2], r[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) {
              const float f = __bfloat162float(o[j]), a = bf16_round(v[j]);
              o[j] = __float2bfloat16_rn((EPI == VCOF_EPI_MUL_BF16) ? f * a : f + a);
            }
        }
      } else if (EPI == VCOF_EPI_RAW_F32) {  // out = acc + bias, unrounded (attention scores)
        float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + n0;
        if (full_chunk) {
          float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            o4[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) o[j] = v[j];
        }
      } else {  // VCOF_EPI_BIAS_F32: out = float(bf16(acc + bias))
        float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + n0;
        if (full_chunk) {
          float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            o4[q] = make_float4(bf16_round(v[q * 4]), bf16_round(v[q * 4 + 1]),
                                bf16_round(v[q * 4 + 2]), bf16_round(v[q * 4 + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + j < p.N) o[j] = bf16_round(v[j]);
        }
      }
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 GemmArgs p) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  // bars: full[S], empty[S], tfull[2], tempty[2], then tmem base slot
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * S;
  const uint32_t bar_tfull = bar_empty + 8 * S;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 5) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(bar_full + 8 * i, p.producers);
        mbar_init(bar_empty + 8 * i, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_tfull + 8 * i, 1);
        mbar_init(bar_tempty + 8 * i, 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + kBM - 1) / kBM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + kBK - 1) / kBK;

  if (warp == 4 || (warp == 6 && p.producers == 2)) {
    // ---------------- TMA producer(s) ----------------
    // Producer and MMA warps run their loops with ALL lanes (warp-uniform control flow and operands) and elect one lane
    // for the issue: inside an `if (lane == 0)` region ptxas keeps every operand in per-thread registers and wraps each
    // UTMALDG / UTCHMMA in R2UR moves plus an ELECT / BRA.U.ANY loop (profiles/r2_attn_issue_bound.md); with uniform
    // operands the descriptors and coordinates live in uniform registers and the issue is one instruction.
    {
      const bool load_a = warp == 4;
      const bool load_b = (warp == 6) || (p.producers == 1);
      const uint32_t tx = (load_a ? Cfg::kABytes : 0) + (load_b ? Cfg::kBBytes : 0);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar_full + 8 * s, tx);
            if (load_a)
              tma_load_2d(smem_u32(sA + s * Cfg::kABytes), &tmA, bar_full + 8 * s, kb * kBK, m_blk * kBM);
            if (load_b)
              tma_load_2d(smem_u32(sB + s * Cfg::kBBytes), &tmB, bar_full + 8 * s, kb * kBK, n_blk * BN);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer ----------------
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, false, false);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint64_t a_desc = make_desc_kmajor_sw128(smem_u32(sA + s * Cfg::kABytes));
          const uint64_t b_desc = make_desc_kmajor_sw128(smem_u32(sB + s * Cfg::kBBytes));
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // +32 bytes per K=16 step = +2 in the descriptor's 16-byte address field
              umma_ss(d_tmem, a_desc + k * 2, b_desc + k * 2, idesc, (kb | k) != 0);
            }
            umma_commit(bar_empty + 8 * s);
            if (kb == num_kb - 1) umma_commit(bar_tfull + 8 * acc);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    // ---------------- epilogue warps 0..3 ----------------
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tc_fence_after();
      const int row = m_blk * kBM + warp * 32 + lane;
      const uint32_t t_row = tmem_base + ((warp * 32u) << 16) + acc * BN;
      gemm_epilogue_tile<BN, EPI>(p, t_row, row, n_blk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2), 256 x 256 tiles per pair; default for the plain bias epilogue (see vcof_gemm_bf16).
// The single-CTA kernel needs 48 KB of
// operands per 512-clock k-block (96 B/clk/SM) against a measured feed of ~77 B/clk/SM; here each CTA stages its own
// 128 rows of A and only HALF of the 256 rows of B (32 KB per k-block, 64 B/clk/SM), six stages deep.
//   rank 0 (leader): its MMA lane issues tcgen05.mma.cta_group::2 (M = 256: 128 TMEM lanes in each CTA) after
//     waiting on ITS full[s], which counts the leader's arrive.expect_tx plus the TMA bytes of both CTAs;
//   both ranks: TMA producer (own A rows, own half of B), completion bytes go to the leader's full[s];
//     slot release = the leader's commit, multicast to empty[s] of both CTAs; accumulator-ready likewise (tfull);
//   both ranks: four epilogue warps drain their own 128 lanes; accumulator-free arrivals all go to the leader's
//     tempty (count 8).
// ---------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int kBN = 256;
  static constexpr int kStages = 6;
  static constexpr int kABytes = kBM * kBK * 2;            // 128 rows of A
  static constexpr int kBBytes = (kBN / 2) * kBK * 2;      // this CTA's 128 of the 256 B rows
  static constexpr int kStageBytes = kABytes + kBBytes;    // 32 KB
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = 512;                    // two 256-column accumulators
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2cta_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     GemmArgs p) {
  using Cfg = Gemm2Cfg;
  constexpr int S = Cfg::kStages;
  constexpr int BN = Cfg::kBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  const uint32_t bar_full = smem_u32(bars);            // used in the leader only
  const uint32_t bar_empty = bar_full + 8 * S;         // per CTA, arrived by the leader's multicast commit
  const uint32_t bar_tfull = bar_empty + 8 * S;        // per CTA, multicast commit
  const uint32_t bar_tempty = bar_tfull + 16;          // leader only, 8 arrivals (4 warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 5) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(bar_full + 8 * i, 1);            // the leader's arrive.expect_tx; both CTAs' TMA bytes complete on it
        mbar_init(bar_empty + 8 * i, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_tfull + 8 * i, 1);
        mbar_init(bar_tempty + 8 * i, 8);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(smem_u32(tmem_slot), Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + 2 * kBM - 1) / (2 * kBM);   // 256-row blocks
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 4) {
    // ---------------- TMA producer (both CTAs) ----------------
    {
      const uint32_t full_leader = mapa_shared(bar_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
        const int row_a = (m_blk * 2 + int(rank)) * kBM;
        const int row_b = n_blk * BN + int(rank) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (elect_one()) {
            // Only the leader arrives (with the bytes of BOTH CTAs).  The peer's boxes may complete on the leader's
            // barrier before that arrive — a transiently negative tx-count is legal, the phase cannot complete while
            // the arrival is pending — and the peer cannot run a phase ahead: it reloads slot s only after the
            // leader's multicast commit on empty[s].  (The first version made the peer arrive remotely with
            // .release.cluster: a GPU-scope fence per stage, 1460 clk per k-block = exactly half speed.)
            if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * Cfg::kStageBytes);
            tma_load_2d_pair(smem_u32(sA + s * Cfg::kABytes), &tmA, full_leader + 8 * s, kb * kBK, row_a);
            tma_load_2d_pair(smem_u32(sB + s * Cfg::kBBytes), &tmB, full_leader + 8 * s, kb * kBK, row_b);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, BN, false, false);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint64_t a_desc = make_desc_kmajor_sw128(smem_u32(sA + s * Cfg::kABytes));
          const uint64_t b_desc = make_desc_kmajor_sw128(smem_u32(sB + s * Cfg::kBBytes));
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_ss_pair(d_tmem, a_desc + k * 2, b_desc + k * 2, idesc, (kb | k) != 0);
            umma_commit_pair(bar_empty + 8 * s);
            if (kb == num_kb - 1) umma_commit_pair(bar_tfull + 8 * acc);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    // ---------------- epilogue warps 0..3 (both CTAs, own 128 accumulator rows) ----------------
    const uint32_t tempty_leader = mapa_shared(bar_tempty, 0);
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tc_fence_after();
      const int row = (m_blk * 2 + int(rank)) * kBM + warp * 32 + lane;
      const uint32_t t_row = tmem_base + ((warp * 32u) << 16) + acc * BN;
      gemm_epilogue_tile<BN, EPI>(p, t_row, row, n_blk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // neither CTA may retire (or free TMEM) while its peer can still signal it
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

template <int EPI>
static int launch_gemm2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm2cta_bf16_kernel<EPI>;
  if (!attr_set) {
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::kSmemBytes));
    attr_set = true;
  }
  const int num_tiles = ((args.M + 2 * kBM - 1) / (2 * kBM)) * ((args.N + Gemm2Cfg::kBN - 1) / Gemm2Cfg::kBN);
  // A persistent kernel must not launch more CTA pairs than can be co-resident: a pair needs both SMs of one TPC,
  // and a part with 148 of its SMs enabled need not have 74 complete TPCs.  Pairs beyond the resident set would run
  // as a second wave with a full share of the tiles each — the first version sized the grid as sm_count / 2 pairs
  // and measured exactly half the single-CTA throughput (profiles/r1_gpurun34_gemm_2cta.log).
  static int max_pairs = 0;
  if (max_pairs == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() & ~1);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Gemm2Cfg::kSmemBytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = sm_count() / 2;
    }
    const char* e = getenv("VCOF_GEMM_2CTA_PAIRS");     // override for experiments
    if (e != nullptr && atoi(e) > 0) n = atoi(e);
    max_pairs = n < sm_count() / 2 ? n : sm_count() / 2;
  }
  const int grid = 2 * (num_tiles < max_pairs ? num_tiles : max_pairs);
  kern<<<grid, kGemmThreads, Gemm2Cfg::kSmemBytes, stream>>>(tmA, tmB, args);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int dispatch_epi_2cta(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args,
                             cudaStream_t stream) {
  switch (epi) {
    case VCOF_EPI_BIAS_BF16: return launch_gemm2cta<VCOF_EPI_BIAS_BF16>(tmA, tmB, args, stream);
    case VCOF_EPI_BIAS_GELU_BF16: return launch_gemm2cta<VCOF_EPI_BIAS_GELU_BF16>(tmA, tmB, args, stream);
    case VCOF_EPI_BIAS_GATE_RES_F32: return launch_gemm2cta<VCOF_EPI_BIAS_GATE_RES_F32>(tmA, tmB, args, stream);
    default: break;
  }
  set_last_error("vcof_gemm_bf16: epilogue %d has no CTA-pair variant", epi);
  return -1;
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<BN, EPI>;
  if (!attr_set) {
    VCOF_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int num_m = (args.M + kBM - 1) / kBM;
  const int num_n = (args.N + BN - 1) / BN;
  const int tiles = num_m * num_n;
  int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, args);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                        const GemmArgs& args, cudaStream_t stream) {
  switch (epi) {
    case VCOF_EPI_BIAS_BF16: return launch_gemm<BN, VCOF_EPI_BIAS_BF16>(tmA, tmB, args, stream);
    case VCOF_EPI_BIAS_GELU_BF16:
      return launch_gemm<BN, VCOF_EPI_BIAS_GELU_BF16>(tmA, tmB, args, stream);
    case VCOF_EPI_BIAS_GATE_RES_F32:
      return launch_gemm<BN, VCOF_EPI_BIAS_GATE_RES_F32>(tmA, tmB, args, stream);
    case VCOF_EPI_BIAS_F32: return launch_gemm<BN, VCOF_EPI_BIAS_F32>(tmA, tmB, args, stream);
    case VCOF_EPI_RAW_F32: return launch_gemm<BN, VCOF_EPI_RAW_F32>(tmA, tmB, args, stream);
    case VCOF_EPI_GATE_ACCUM_BF16:
      return launch_gemm<BN, VCOF_EPI_GATE_ACCUM_BF16>(tmA, tmB, args, stream);
    case VCOF_EPI_MUL_BF16: return launch_gemm<BN, VCOF_EPI_MUL_BF16>(tmA, tmB, args, stream);
    case VCOF_EPI_ADD_BF16: return launch_gemm<BN, VCOF_EPI_ADD_BF16>(tmA, tmB, args, stream);
  }
  set_last_error("vcof_gemm_bf16: unknown epilogue %d", epi);
  return -1;
}

}  // namespace vcof

using namespace vcof;

extern "C" int vcof_gemm_bf16(const void* a, long long lda, const void* w, long long ldw,
                              const void* bias, const float* gate, void* out, long long ldo, int M,
                              int N, int K, int epilogue, void* stream) {
  VCOF_REQUIRE(M > 0 && N > 0 && K > 0, "vcof_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  const bool narrow = (epilogue & VCOF_GEMM_TILE128) != 0;
  epilogue &= ~VCOF_GEMM_TILE128;
  VCOF_REQUIRE(lda % 8 == 0 && ldw % 8 == 0,
               "vcof_gemm_bf16: lda/ldw must be multiples of 8 (16-byte TMA rows)");
  const bool f32_out = (epilogue == VCOF_EPI_BIAS_GATE_RES_F32 || epilogue == VCOF_EPI_BIAS_F32 ||
                        epilogue == VCOF_EPI_RAW_F32);
  VCOF_REQUIRE(ldo % (f32_out ? 4 : 8) == 0, "vcof_gemm_bf16: ldo=%lld breaks 16-byte row alignment",
               ldo);
  VCOF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "vcof_gemm_bf16: out not 16B aligned");
  VCOF_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "vcof_gemm_bf16: bias not 16B aligned");
  VCOF_REQUIRE(gate == nullptr || (reinterpret_cast<uintptr_t>(gate) & 15) == 0,
               "vcof_gemm_bf16: gate not 16B aligned");
  VCOF_REQUIRE(epilogue != VCOF_EPI_GATE_ACCUM_BF16 || (gate != nullptr && bias == nullptr),
               "vcof_gemm_bf16: the accumulate epilogue needs a gate vector and takes no bias");
  // VCOF_GEMM_TILE128: the caller asks for 128-wide tiles (skinny problems whose 256-wide tiling would leave SMs idle)
  const int BN = (N > 128) ? (narrow ? 128 : 256) : (N > 64 ? 128 : 64);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, a, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, w, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, kBK, BN);
  if (rc) return rc;
  GemmArgs args;
  args.M = M; args.N = N; args.K = K;
  args.bias = reinterpret_cast<const bf16*>(bias);
  args.gate = gate;
  args.out = out;
  args.ldo = ldo;
  {
    // m-blocks per raster group (tiles of a group are walked m-fastest, so a wave of CTAs shares `group_m` A row blocks
    // and ceil(waves / group_m) B column blocks).  VCOF_GEMM_GROUP_M overrides for A/B runs.
    static const int group_m = [] {
      const char* e = getenv("VCOF_GEMM_GROUP_M");
      const int v = e ? atoi(e) : 0;
      return v > 0 ? v : 16;      // measured (profiles/r2_gpurun18_gemm_group_sweep.log): 16 beats 8 by 1-7 %, 64 loses 10 %
    }();
    args.group_m = group_m;
  }
  {
    static const int producers = [] {
      const char* e = getenv("VCOF_GEMM_PRODUCERS");
      return (e != nullptr && e[0] == '1') ? 1 : 2;
    }();
    args.producers = producers;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // CTA-pair kernel (cta_group::2, 256 x 256 tiles).  Measured on B200 (profiles/r2_gpurun9_gemm_2cta_fixed.log):
  // bias epilogue 75600 x 5120 x 5120: 1540 TFLOP/s against 1453 for the single-CTA kernel (cuBLAS: 1507).  Under ncu
  // the pair kernel keeps the tensor pipe 99.6 % busy with every epilogue and needs 11 % fewer cycles than the single-CTA
  // kernel — but a B200 under its power cap trades that for clock, so in milliseconds pairs win where their operand
  // traffic is lowest: with raster groups of 16 m-blocks the bias and GELU epilogues gain 5-7 %
  // (profiles/r2_gpurun18_gemm_group_sweep.log), the gate + fp32-residual epilogue (K = 13824: operands do not stay in
  // L2) is level, at K = 5120 it gains 7 % (1396 vs 1305 TFLOP/s, gpurun call 21).  Default ("auto"): pairs for bias,
  // bias+GELU, and gate + residual up to K = 8192.  VCOF_GEMM_2CTA=0: never, =1: all three epilogues.
  static const int pair_mode = [] {
    const char* e = getenv("VCOF_GEMM_2CTA");
    return e == nullptr ? 2 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2));
  }();
  const bool pair_ok = !narrow && N >= 256 && M >= 256 &&
                       ((pair_mode == 1 && epilogue <= VCOF_EPI_BIAS_GATE_RES_F32) ||
                        (pair_mode == 2 && (epilogue == VCOF_EPI_BIAS_BF16 || epilogue == VCOF_EPI_BIAS_GELU_BF16 ||
                                            (epilogue == VCOF_EPI_BIAS_GATE_RES_F32 && K <= 8192))));
  if (pair_ok) {
    // B box is this CTA's 128-row half of the 256-row tile
    rc = make_tmap_2d_bf16(&tmB, w, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, kBK, 128);
    if (rc) return rc;
    return dispatch_epi_2cta(epilogue, tmA, tmB, args, st);
  }
  if (BN == 256) return dispatch_epi<256>(epilogue, tmA, tmB, args, st);
  if (BN == 128) return dispatch_epi<128>(epilogue, tmA, tmB, args, st);
  return dispatch_epi<64>(epilogue, tmA, tmB, args, st);
}
