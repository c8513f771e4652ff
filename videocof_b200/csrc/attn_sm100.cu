// attn_sm100.cu — FlashAttention-style non-causal attention forward on tcgen05
// for sm_100a, head_dim 128, bf16 in / bf16 out, fp32 softmax statistics.
//
// Replaces the reference's attention() dispatch (videox_fun/models/attention_utils.py:152-210,
// i.e. flash_attn_varlen_func / SDPA) for the DiT self-attention
// (wan_transformer3d.py:294-299) and text cross-attention (:325-330).
//
// One persistent CTA per SM walks (head, 256-query-row block) work items.  Each
// item keeps TWO 128-row Q tiles in flight (ping-pong) so that the tensor pipe
// works on one tile while the softmax warps work on the other:
//
//   TMEM (512 columns): S0 | S1 | O0 | O1   (128 fp32 columns each);
//   P_t (bf16, 64 columns) aliases the start of S_t and feeds the PV MMA
//   directly from TMEM (tcgen05.mma with A in tensor memory).
//
//   warps 0-3  : softmax + epilogue for tile 0 (thread i <-> TMEM lane i <-> one query row)
//   warps 4-7  : softmax + epilogue for tile 1
//   warp  8    : TMA producer (Q tiles, K/V ring, 128B-swizzled 64-column boxes)
//   warp  9    : tcgen05.mma issuer: S_t = Q_t K_j^T, then O_t += P_t V_j
//   warp  10   : TMEM allocator
//
// Online softmax uses exp2 with the scale folded in, and a lazy rescale: the
// running row max is only advanced (and O_t rescaled in TMEM) when it grew by
// more than 2^8, so in steady state the softmax warps never touch O.
#include <stdlib.h>

#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

constexpr int kHD = 128;              // head dim
constexpr int kQT = 128;              // query rows per tile
constexpr int kKT = 128;              // kv rows per tile
constexpr int kTile = kQT * kHD * 2;  // 32 KiB bf16 tile
constexpr int kHalf = kTile / 2;      // one 64-column TMA box
constexpr int kKVStages = 2;
constexpr int kAttnThreads = 384;
constexpr float kRescaleThresh = 8.0f;  // log2 units

struct AttnArgs {
  int Lq, kv_len, heads, num_q_blocks;
  float scale_log2;
  bf16* out;
  long long ldo;
  int rows_per_chunk;   // OUT_SCATTER only
};

// OUT_SCATTER: the output rows are split into chunks of rows_per_chunk consecutive query rows and chunk c is stored to
// the dense [rows_per_chunk, ldo] slab p[c] instead of `out`.  Under the push exchange of the sequence-parallel path
// (videocof_b200/dist.py) chunk c is the rows rank c owns and p[c] is its receive buffer, mapped through NVLink peer
// memory: the return leg of the head exchange becomes the epilogue's own stores, overlapping the attention's tail.
constexpr int kMaxOutChunks = 16;
struct AttnOutChunks {
  bf16* p[kMaxOutChunks];
};

constexpr int kAttnSmem = 2 * kTile + 2 * kKVStages * kTile + 256 + 1024;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Same instruction as a volatile statement: keeps its place relative to tcgen05.wait::ld (SPEC == 2 wants a batch of
// exponentials issued BEFORE the wait for the second half of the score load).
__device__ __forceinline__ float ex2_ordered(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max rel.
// error 7.5e-5 << bf16 P precision).  MUFU.EX2 runs at 16/clk/SM, i.e. exactly as long as the MMAs of
// a 128x128 tile; moving a fraction of the exponentials off the XU pipe shortens the softmax leg of the
// S -> softmax -> PV dependency chain.  Packed fp32x2 instructions (FFMA2 / FADD2) halve the issue cost.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);  // 1.5 * 2^23: x + magic rounds x to an integer
  const float2 t = __fadd2_rn(x, magic);
  const float2 nf = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(nf, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.242611125f, 0.242611125f));
  q = __ffma2_rn(q, f, make_float2(0.693260968f, 0.693260968f));
  q = __ffma2_rn(q, f, make_float2(0.999928057f, 0.999928057f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

// EMU: of every 8 element pairs, this many take the polynomial path (0 = all MUFU).
// SPLIT_S: S_t(j+1) = Q_t K_{j+1}^T is issued as two N = 64 halves; the half that lands in score columns 64..127 —
// which P_t(j) does not alias — is issued as soon as the softmax warps have S_t(j) in registers, i.e. under the
// exponentials of step j, and only the other half follows PV_t(j) on the tile's dependency chain.
// PSPLIT: number of K-chunks (2 or 4) in which P is published to the PV MMA.
// SPEC (experimental, VCOF_ATTN_SPEC=1, not yet run on hardware): the row maximum leaves the critical path.  The
// exponentials of a tile start against the STALE reference maximum as soon as the scores are in registers; the
// maximum of the shifted scores is folded into the loop of the first 64 exponentials (ALU pipe, hidden under MUFU)
// and only if some row outgrew the reference by more than 2^8 — the same condition under which the default kernel
// rescales — is the tile redone against the advanced reference.  Same arithmetic as the default path otherwise.
// SPEC == 2 (experimental, VCOF_ATTN_SPEC=2, not yet run on hardware): SPEC == 1 plus a split score load — the
// first 64 columns are waited for alone, the second 64 stay in flight (tcgen05.ld is asynchronous until
// tcgen05.wait::ld) under the first 32 of the 64 first-half exponential pairs, and their shift + maximum is folded
// into the other 32.  The check that decides whether the tile must be redone still covers all 128 columns and still
// precedes the first publication of P, so the arithmetic is that of SPEC == 1.
template <bool V_TRANS, int EMU, bool SPLIT_S, int PSPLIT, int SPEC = 0, bool OUT_SCATTER = false>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, AttnArgs p, const __grid_constant__ AttnOutChunks oc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // 2 tiles
  uint8_t* sK = smem + 2 * kTile;              // kKVStages tiles
  uint8_t* sV = sK + kKVStages * kTile;        // kKVStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kTile);
  const uint32_t b0 = smem_u32(bars);
  // barrier map (8 bytes each)
  const uint32_t q_full = b0, q_empty = b0 + 16;
  const uint32_t k_full = b0 + 32, k_empty = b0 + 48;
  const uint32_t v_full = b0 + 64, v_empty = b0 + 80;
  const uint32_t s_full = b0 + 96, p_full = b0 + 112, o_full = b0 + 128, p_full2 = b0 + 144;
  const uint32_t s_cons = b0 + 160;   // SPLIT_S: S_t(j) is in the softmax warps' registers (4 arrivals)
  const uint32_t p_part = b0 + 176;   // [PSPLIT][2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + 8 * i, 1);
      mbar_init(q_empty + 8 * i, 1);
      mbar_init(k_full + 8 * i, 1);
      mbar_init(k_empty + 8 * i, 1);
      mbar_init(v_full + 8 * i, 1);
      mbar_init(v_empty + 8 * i, 1);
      mbar_init(s_full + 8 * i, 1);
      mbar_init(p_full + 8 * i, 4);
      mbar_init(p_full2 + 8 * i, 4);
      mbar_init(s_cons + 8 * i, 4);
      for (int q = 0; q < 4; ++q) mbar_init(p_part + 8 * (q * 2 + i), 4);
      mbar_init(o_full + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_kv = (p.kv_len + kKT - 1) / kKT;
  const int num_items = p.heads * p.num_q_blocks;

  // Register re-balancing between warpgroups: the softmax threads keep a whole 128-wide
  // score row live, the producer / MMA warpgroup needs almost nothing.
  if (warp >= 8) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
   if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t item_ph = 0;
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, item_ph ^= 1) {
        const int head = item / p.num_q_blocks;
        const int q0 = (item % p.num_q_blocks) * 2 * kQT;
        const int col = head * kHD;
        auto load_q = [&](int t) {
          mbar_wait(q_empty + 8 * t, item_ph ^ 1);
          mbar_expect_tx(q_full + 8 * t, kTile);
          tma_load_2d(smem_u32(sQ + t * kTile), &tmQ, q_full + 8 * t, col, q0 + t * kQT);
          tma_load_2d(smem_u32(sQ + t * kTile + kHalf), &tmQ, q_full + 8 * t, col + 64,
                      q0 + t * kQT);
        };
        auto load_k = [&](int j) {
          mbar_wait(k_empty + 8 * ks, kph ^ 1);
          mbar_expect_tx(k_full + 8 * ks, kTile);
          tma_load_2d(smem_u32(sK + ks * kTile), &tmK, k_full + 8 * ks, col, j * kKT);
          tma_load_2d(smem_u32(sK + ks * kTile + kHalf), &tmK, k_full + 8 * ks, col + 64, j * kKT);
          if (++ks == kKVStages) { ks = 0; kph ^= 1; }
        };
        auto load_v = [&](int j) {
          mbar_wait(v_empty + 8 * vs, vph ^ 1);
          mbar_expect_tx(v_full + 8 * vs, kTile);
          if (V_TRANS) {
            // V^T [C, Lk]: box = 64 kv columns x 128 head-dim rows
            tma_load_2d(smem_u32(sV + vs * kTile), &tmV, v_full + 8 * vs, j * kKT, col);
            tma_load_2d(smem_u32(sV + vs * kTile + kHalf), &tmV, v_full + 8 * vs, j * kKT + 64,
                        col);
          } else {
            tma_load_2d(smem_u32(sV + vs * kTile), &tmV, v_full + 8 * vs, col, j * kKT);
            tma_load_2d(smem_u32(sV + vs * kTile + kHalf), &tmV, v_full + 8 * vs, col + 64,
                        j * kKT);
          }
          if (++vs == kKVStages) { vs = 0; vph ^= 1; }
        };
        load_q(0);
        load_k(0);
        load_q(1);
        load_v(0);
        for (int j = 1; j < n_kv; ++j) {
          load_k(j);
          load_v(j);
        }
      }
    }
   } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    // The WHOLE warp walks the loop (warp-uniform control flow and operands) and one elected lane issues each group
    // of tcgen05 instructions.  Written as `if (lane == 0) { loop }` the same code kept every descriptor in a
    // per-thread register inside a divergent region, and ptxas wrapped each UTCHMMA in four R2UR moves plus an
    // ELECT / BRA.U.ANY loop: ~89 clk of issue per 64-clk MMA, measured as an MMA warp that never waits on a barrier
    // (profiles/r2_attn_issue_bound.md) — the kernel was issue-bound, not softmax-bound.  With uniform operands the
    // descriptors live in uniform registers and a group of eight MMAs is eight back-to-back UTCHMMA.
    {
      constexpr uint32_t idesc_qk = make_idesc_bf16(kQT, kKT, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kQT, kHD, false, !V_TRANS);
      uint32_t item_ph = 0;
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t pph[2] = {0, 0};
      uint32_t cph[2] = {0, 0};
      const uint32_t tS[2] = {tmem_base, tmem_base + 128};
      const uint32_t tO[2] = {tmem_base + 256, tmem_base + 384};

      constexpr uint32_t idesc_qk64 = make_idesc_bf16(kQT, 64, false, false);
      // S_t = Q_t K^T, committed to s_full; `release_q` / `release_k`: also commit the Q tile / K stage back to the
      // producer.  SPLIT_S issues it as two N = 64 halves (see the main loop): half 1 = score columns 64..127 = K rows
      // 64..127 of the stage, half 0 = columns 0..63; only the half issued last commits.
      auto mma_s_half = [&](int t, int kstage, int half, bool commit, bool release_q, bool release_k) {
        const uint64_t ad = make_desc_kmajor_sw128(smem_u32(sQ + t * kTile));
        const uint64_t bd = make_desc_kmajor_sw128(smem_u32(sK + kstage * kTile)) + ((half * 64 * 128) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) {
            const uint32_t off = ((k >> 2) * kHalf + (k & 3) * 32) >> 4;
            umma_ss(tS[t] + half * 64, ad + off, bd + off, idesc_qk64, k != 0);
          }
          if (commit) umma_commit(s_full + 8 * t);
          if (release_q) umma_commit(q_empty + 8 * t);
          if (release_k) umma_commit(k_empty + 8 * kstage);
        }
        __syncwarp();
      };
      auto mma_s = [&](int t, int kstage, bool release_q, bool release_k) {
        if (SPLIT_S) {
          mma_s_half(t, kstage, 1, false, false, false);
          mma_s_half(t, kstage, 0, true, release_q, release_k);
          return;
        }
        // descriptors differ only in the 14-bit start-address field: build once, add (byte offset >> 4)
        const uint64_t ad = make_desc_kmajor_sw128(smem_u32(sQ + t * kTile));
        const uint64_t bd = make_desc_kmajor_sw128(smem_u32(sK + kstage * kTile));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) {
            const uint32_t off = ((k >> 2) * kHalf + (k & 3) * 32) >> 4;
            umma_ss(tS[t], ad + off, bd + off, idesc_qk, k != 0);
          }
          umma_commit(s_full + 8 * t);
          if (release_q) umma_commit(q_empty + 8 * t);
          if (release_k) umma_commit(k_empty + 8 * kstage);
        }
        __syncwarp();
      };
      // PV in two K-halves: the first half (kv rows 0..63) is issued as soon as the softmax warps have
      // published P[:, 0:64], overlapping the exponentials of the second half.  `done_bar` (0 = none) is committed
      // after the last part: o_full on the last kv tile; `release_v`: commit the V stage back to the producer.
      auto mma_pv = [&](int t, int vstage, bool first, uint32_t parity, uint32_t done_bar, bool release_v) {
        const uint32_t b = smem_u32(sV + vstage * kTile);
        const uint64_t bd0 = V_TRANS ? make_desc_kmajor_sw128(b) : make_desc_mnmajor_sw128(b, kHalf, 1024);
#pragma unroll
        for (int part = 0; part < PSPLIT; ++part) {
          mbar_wait(p_part + 8 * (part * 2 + t), parity);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kKT / 16 / PSPLIT; ++kk) {
              const int k = part * (kKT / 16 / PSPLIT) + kk;
              const uint32_t off = V_TRANS ? (((k >> 2) * kHalf + (k & 3) * 32) >> 4) : ((k * 16 * 128) >> 4);
              umma_ts(tO[t], tS[t] + k * 8, bd0 + off, idesc_pv, (!first || k != 0));
            }
            if (part == PSPLIT - 1) {
              if (done_bar != 0) umma_commit(done_bar);
              if (release_v) umma_commit(v_empty + 8 * vstage);
            }
          }
          __syncwarp();
        }
      };

      for (int item = blockIdx.x; item < num_items; item += gridDim.x, item_ph ^= 1) {
        // prologue: S0(0), S1(0)
        mbar_wait(q_full + 0, item_ph);
        mbar_wait(k_full + 8 * ks, kph);
        tc_fence_after();
        mma_s(0, ks, n_kv == 1, false);
        mbar_wait(q_full + 8, item_ph);
        tc_fence_after();
        mma_s(1, ks, n_kv == 1, true);
        if (++ks == kKVStages) { ks = 0; kph ^= 1; }

        for (int j = 0; j < n_kv; ++j) {
          const bool last = (j + 1 == n_kv);
          if (SPLIT_S) {
            // P_t(j) (bf16) aliases only score columns 0..63, so columns 64..127 of S_t are free as soon as the softmax
            // warps hold S_t(j) in registers (s_cons).  The second half of S_t(j+1) is issued right then — it runs on the
            // tensor pipe while the exponentials of S_t(j) are computed — and only the first half (256 clk instead of
            // 512) stays on the tile's S -> softmax -> PV -> S dependency chain, after PV_t(j).
            if (!last) {
              mbar_wait(k_full + 8 * ks, kph);
              mbar_wait(s_cons + 0, cph[0]);
              cph[0] ^= 1;
              tc_fence_after();
              mma_s_half(0, ks, 1, false, false, false);
            }
            mbar_wait(v_full + 8 * vs, vph);
            mma_pv(0, vs, j == 0, pph[0], last ? o_full + 0 : 0u, false);
            pph[0] ^= 1;
            if (!last) {
              mma_s_half(0, ks, 0, true, j + 2 == n_kv, false);
              mbar_wait(s_cons + 8, cph[1]);
              cph[1] ^= 1;
              tc_fence_after();
              mma_s_half(1, ks, 1, false, false, false);
            }
            mma_pv(1, vs, j == 0, pph[1], last ? o_full + 8 : 0u, true);
            pph[1] ^= 1;
            if (++vs == kKVStages) { vs = 0; vph ^= 1; }
            if (!last) {
              mma_s_half(1, ks, 0, true, j + 2 == n_kv, true);
              if (++ks == kKVStages) { ks = 0; kph ^= 1; }
            }
            continue;
          }
          // ---- tile 0: O0 += P0(j) V_j ; then S0(j+1)
          mbar_wait(v_full + 8 * vs, vph);
          mma_pv(0, vs, j == 0, pph[0], last ? o_full + 0 : 0u, false);
          pph[0] ^= 1;
          if (!last) {
            mbar_wait(k_full + 8 * ks, kph);
            tc_fence_after();
            mma_s(0, ks, j + 2 == n_kv, false);
          }
          // ---- tile 1: O1 += P1(j) V_j ; then S1(j+1)
          mma_pv(1, vs, j == 0, pph[1], last ? o_full + 8 : 0u, true);
          pph[1] ^= 1;
          if (++vs == kKVStages) { vs = 0; vph ^= 1; }
          if (!last) {
            mma_s(1, ks, j + 2 == n_kv, true);
            if (++ks == kKVStages) { ks = 0; kph ^= 1; }
          }
        }
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // =========================== softmax + epilogue ===========================
    const int t = warp >> 2;                         // which Q tile
    const uint32_t lane_base = ((warp & 3) * 32u) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_base;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_base;
    const int row_in_tile = (warp & 3) * 32 + lane;
    uint32_t sph = 0, oph = 0;
    const int rem = p.kv_len - (n_kv - 1) * kKT;  // valid columns in the last kv tile (1..128)

    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int head = item / p.num_q_blocks;
      const int q0 = (item % p.num_q_blocks) * 2 * kQT;
      float m_ref = -INFINITY;  // running (possibly stale) row max, raw score units
      float l = 0.f;
      if constexpr (SPEC == 2) {
        static_assert(SPEC != 2 || (EMU == 0 && !SPLIT_S && PSPLIT == 2), "SPEC builds on the default variant");
        float mb = -INFINITY;   // running (possibly stale) row max in scaled log2 units
        for (int j = 0; j < n_kv; ++j) {
          uint32_t s[128];
          mbar_wait(s_full + 8 * t, sph);
          tc_fence_after();
          sph ^= 1;
          tmem_ld32(tS + 0, s + 0);
          tmem_ld32(tS + 32, s + 32);
          tmem_ld_wait();
          tmem_ld32(tS + 64, s + 64);      // in flight under the first exponentials; s[64..127] must not be touched
          tmem_ld32(tS + 96, s + 96);      // before the next tcgen05.wait::ld
          const bool ragged = (j == n_kv - 1 && rem < kKT);
          if (ragged) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (c >= rem) s[c] = 0xff800000u;  // -inf
          }
          if (j == 0) {   // first tile: no reference yet, the maximum of all 128 columns has to come first
            tmem_ld_wait();
            if (ragged) {
#pragma unroll
              for (int c = 64; c < 128; ++c)
                if (c >= rem) s[c] = 0xff800000u;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 128; c += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(s[c]));
              mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
            }
            mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
          }
          const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
          const float2 nmb2 = make_float2(-mb, -mb);
          float2 x[64];
#pragma unroll
          for (int c = 0; c < 32; ++c)
            x[c] = __ffma2_rn(make_float2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, nmb2);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          uint32_t pk[32];
          float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c == 16) {     // 32 exponential pairs (~256 clk of MUFU per warp) after the loads were issued
              tmem_ld_wait();
              if (j > 0 && ragged) {
#pragma unroll
                for (int e = 64; e < 128; ++e)
                  if (e >= rem) s[e] = 0xff800000u;
              }
            }
            float2 pr;
            pr.x = c < 16 ? ex2_ordered(x[c].x) : ex2(x[c].x);
            pr.y = c < 16 ? ex2_ordered(x[c].y) : ex2(x[c].y);
            if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
            pk[c] = pack_bf16x2(pr.x, pr.y);
            mxa = fmaxf(fmaxf(mxa, x[c].x), x[c].y);
            if (c >= 16) {     // shift + maximum of the second half, two pairs per iteration, hidden under MUFU
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int e = 32 + 2 * (c - 16) + q;
                x[e] = __ffma2_rn(make_float2(__uint_as_float(s[2 * e]), __uint_as_float(s[2 * e + 1])), sc2, nmb2);
                mxb = fmaxf(fmaxf(mxb, x[e].x), x[e].y);
              }
            }
          }
          const float mxx = fmaxf(mxa, mxb);
          if (j > 0 && __any_sync(0xffffffffu, mxx > kRescaleThresh)) {
            // rare: as in SPEC == 1 — advance the reference by d (per row), rescale O_t and l, redo the first half
            const float d = fmaxf(mxx, 0.f);
            const float alpha = ex2(-d);
            l *= alpha;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tO + c * 32, o);
            }
            tmem_st_wait();
            mb += d;
            const float2 nd2 = make_float2(-d, -d);
#pragma unroll
            for (int c = 32; c < 64; ++c) x[c] = __fadd2_rn(x[c], nd2);
            // First half: S_t(j) is still in TMEM (P is stored only below), so its 64 columns are reloaded here and
            // shifted against the advanced reference — keeping their shifted values live through the common path for
            // the sake of this branch cost 10-14 spilled registers (profiles/r1_sass_summary.txt).
            const float2 nmbr = make_float2(-mb, -mb);
            acc0 = make_float2(0.f, 0.f);
            acc1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {        // 32 columns at a time: this branch must not set the register budget
              uint32_t r[32];
              tmem_ld32(tS + 32 * h, r);
              tmem_ld_wait();
              if (ragged) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (32 * h + c >= rem) r[c] = 0xff800000u;  // -inf
              }
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float2 xr =
                    __ffma2_rn(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), sc2, nmbr);
                float2 pr;
                pr.x = ex2(xr.x);
                pr.y = ex2(xr.y);
                if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
                pk[16 * h + c] = pack_bf16x2(pr.x, pr.y);
              }
            }
          }
          tmem_st32(tS, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_part + 8 * (0 * 2 + t));
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float2 pr;
            pr.x = ex2(x[32 + c].x);
            pr.y = ex2(x[32 + c].y);
            if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
            pk[c] = pack_bf16x2(pr.x, pr.y);
          }
          tmem_st32(tS + 32, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_part + 8 * (1 * 2 + t));
          l += (acc0.x + acc1.x) + (acc0.y + acc1.y);
        }
      } else if constexpr (SPEC == 1) {
        static_assert(SPEC != 1 || (EMU == 0 && !SPLIT_S && PSPLIT == 2), "SPEC builds on the default variant");
        float mb = -INFINITY;   // running (possibly stale) row max in scaled log2 units
        for (int j = 0; j < n_kv; ++j) {
          uint32_t s[128];
          mbar_wait(s_full + 8 * t, sph);
          tc_fence_after();
          sph ^= 1;
          tmem_ld32(tS + 0, s + 0);
          tmem_ld32(tS + 32, s + 32);
          tmem_ld32(tS + 64, s + 64);
          tmem_ld32(tS + 96, s + 96);
          tmem_ld_wait();
          if (j == n_kv - 1 && rem < kKT) {
#pragma unroll
            for (int c = 0; c < 128; ++c)
              if (c >= rem) s[c] = 0xff800000u;  // -inf
          }
          if (j == 0) {   // first tile: no reference yet, the maximum has to come first
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 128; c += 4) {
              mx0 = fmaxf(mx0, __uint_as_float(s[c]));
              mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
              mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
            }
            mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
          }
          const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
          const float2 nmb2 = make_float2(-mb, -mb);
          float2 x[64];
#pragma unroll
          for (int c = 0; c < 64; ++c)
            x[c] = __ffma2_rn(make_float2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, nmb2);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          uint32_t pk[32];
          float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float2 pr;
            pr.x = ex2(x[c].x);
            pr.y = ex2(x[c].y);
            if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
            pk[c] = pack_bf16x2(pr.x, pr.y);
            mxa = fmaxf(fmaxf(mxa, x[c].x), x[c].y);
            mxb = fmaxf(fmaxf(mxb, x[32 + c].x), x[32 + c].y);
          }
          const float mxx = fmaxf(mxa, mxb);
          if (j > 0 && __any_sync(0xffffffffu, mxx > kRescaleThresh)) {
            // rare: advance the reference by d (per row), rescale O_t and l, redo the first half.  PV_t(j-1) retired
            // before s_full fired and PV_t(j) is not issued until we arrive on p_part: O_t is ours to rescale.
            const float d = fmaxf(mxx, 0.f);
            const float alpha = ex2(-d);
            l *= alpha;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tO + c * 32, o);
            }
            tmem_st_wait();
            mb += d;
            const float2 nd2 = make_float2(-d, -d);
#pragma unroll
            for (int c = 32; c < 64; ++c) x[c] = __fadd2_rn(x[c], nd2);
            // First half: S_t(j) is still in TMEM (P is stored only below), so its 64 columns are reloaded here and
            // shifted against the advanced reference — keeping their shifted values live through the common path for
            // the sake of this branch cost 10-14 spilled registers (profiles/r1_sass_summary.txt).
            const float2 nmbr = make_float2(-mb, -mb);
            acc0 = make_float2(0.f, 0.f);
            acc1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {        // 32 columns at a time: this branch must not set the register budget
              uint32_t r[32];
              tmem_ld32(tS + 32 * h, r);
              tmem_ld_wait();
              if (j == n_kv - 1 && rem < kKT) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (32 * h + c >= rem) r[c] = 0xff800000u;  // -inf
              }
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float2 xr =
                    __ffma2_rn(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), sc2, nmbr);
                float2 pr;
                pr.x = ex2(xr.x);
                pr.y = ex2(xr.y);
                if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
                pk[16 * h + c] = pack_bf16x2(pr.x, pr.y);
              }
            }
          }
          tmem_st32(tS, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_part + 8 * (0 * 2 + t));
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float2 pr;
            pr.x = ex2(x[32 + c].x);
            pr.y = ex2(x[32 + c].y);
            if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
            pk[c] = pack_bf16x2(pr.x, pr.y);
          }
          tmem_st32(tS + 32, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_part + 8 * (1 * 2 + t));
          l += (acc0.x + acc1.x) + (acc0.y + acc1.y);
        }
      } else
      for (int j = 0; j < n_kv; ++j) {
        uint32_t s[128];
        mbar_wait(s_full + 8 * t, sph);
        tc_fence_after();
        tmem_ld32(tS + 0, s + 0);
        tmem_ld32(tS + 32, s + 32);
        sph ^= 1;
        tmem_ld32(tS + 64, s + 64);
        tmem_ld32(tS + 96, s + 96);
        tmem_ld_wait();
        if (SPLIT_S && j + 1 < n_kv) {     // the scores are in registers: columns 64..127 of S_t may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_cons + 8 * t);
        }
        if (j == n_kv - 1 && rem < kKT) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= rem) s[c] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const float m_new = fmaxf(m_ref, m_tile);
        const bool grow = (m_new - m_ref) * p.scale_log2 > kRescaleThresh;  // true at j == 0
        if (__any_sync(0xffffffffu, grow)) {
          // The branch is warp-wide (tcgen05.ld / st are warp-collective) but the DECISION is per row: a row whose
          // maximum did not outgrow its reference keeps it (alpha = 2^0 = 1 exactly), so a row's result depends only on
          // its own scores — not on which rows share its warp, i.e. not on how the query rows are partitioned into
          // tiles, ranks or launches (the sequence-parallel shards reproduce the single-GPU bits for any input).
          const float m_next = grow ? m_new : m_ref;
          if (j > 0) {
            // PV_t(j-1) retired before s_full fired (same commit group), and PV_t(j)
            // is not issued until we arrive on p_full: O_t is ours to rescale.
            const float alpha = grow ? ex2((m_ref - m_next) * p.scale_log2) : 1.0f;
            l *= alpha;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(tO + c * 32, o);
            }
            tmem_st_wait();
          }
          m_ref = m_next;
        }
        const float mb = m_ref * p.scale_log2;
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float2 nmb2 = make_float2(-mb, -mb);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < PSPLIT; ++h) {
          constexpr int kPairs = 64 / PSPLIT;          // packed bf16x2 columns per part
          uint32_t pk[kPairs];
#pragma unroll
          for (int c = 0; c < kPairs; ++c) {
            const int e0 = h * (128 / PSPLIT) + 2 * c;
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[e0]), __uint_as_float(s[e0 + 1])),
                                        sc2, nmb2);
            float2 pr;
            if ((c & 7) < EMU) {
              pr = exp2_poly2(x);
            } else {
              pr.x = ex2(x.x);
              pr.y = ex2(x.y);
            }
            if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
            pk[c] = pack_bf16x2(pr.x, pr.y);
          }
          if (PSPLIT == 2) tmem_st32(tS + h * kPairs, pk); else tmem_st16(tS + h * kPairs, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_part + 8 * (h * 2 + t));
        }
        const float l0 = acc0.x + acc1.x, l1 = acc0.y + acc1.y, l2 = 0.f, l3 = 0.f;
        l += (l0 + l1) + (l2 + l3);
      }
      // ---- epilogue: O_t / l -> bf16 -> global
      mbar_wait(o_full + 8 * t, oph);
      oph ^= 1;
      tc_fence_after();
      const float inv = 1.0f / l;
      const int row = q0 + t * kQT + row_in_tile;
      bf16* orow;
      if constexpr (OUT_SCATTER) {
        const int chunk = row < p.Lq ? row / p.rows_per_chunk : 0;      // rows >= Lq are never stored
        orow = oc.p[chunk] + (long long)(row - chunk * p.rows_per_chunk) * p.ldo + head * kHD;
      } else {
        orow = p.out + (long long)row * p.ldo + head * kHD;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + c * 32, o);
        tmem_ld_wait();
        if (row < p.Lq) {
          uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            o4[q] = w;
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vcof

using namespace vcof;

static int attn_fwd_impl(const void* q, long long ldq, const void* k, long long ldk,
                         const void* v, long long ldv, void* out, long long ldo, int Lq, int Lk,
                         int kv_len, int heads, int head_dim, float softmax_scale,
                         int v_transposed, void* stream, const AttnOutChunks* chunks, int rows_per_chunk) {
  VCOF_REQUIRE(head_dim == kHD, "vcof_attn_fwd: head_dim %d unsupported (only 128)", head_dim);
  VCOF_REQUIRE(Lq > 0 && Lk > 0 && heads > 0, "vcof_attn_fwd: empty problem");
  VCOF_REQUIRE(kv_len >= 1 && kv_len <= Lk, "vcof_attn_fwd: kv_len %d outside [1, %d]", kv_len, Lk);
  VCOF_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "vcof_attn_fwd: leading dims must be multiples of 8 elements");
  VCOF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "vcof_attn_fwd: out not 16B aligned");
  const int C = heads * kHD;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_2d_bf16(&tmQ, q, (uint64_t)C, (uint64_t)Lq, (uint64_t)ldq * 2, 64, kQT);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmK, k, (uint64_t)C, (uint64_t)kv_len, (uint64_t)ldk * 2, 64, kKT);
  if (rc) return rc;
  if (v_transposed) {
    // V^T stored [C, ldv] with kv contiguous
    rc = make_tmap_2d_bf16(&tmV, v, (uint64_t)kv_len, (uint64_t)C, (uint64_t)ldv * 2, 64, kHD);
  } else {
    rc = make_tmap_2d_bf16(&tmV, v, (uint64_t)C, (uint64_t)kv_len, (uint64_t)ldv * 2, 64, kKT);
  }
  if (rc) return rc;
  AttnArgs a;
  a.Lq = Lq;
  a.kv_len = kv_len;
  a.heads = heads;
  a.num_q_blocks = (Lq + 2 * kQT - 1) / (2 * kQT);
  a.scale_log2 = softmax_scale * 1.4426950408889634f;
  a.out = reinterpret_cast<bf16*>(out);
  a.ldo = ldo;
  a.rows_per_chunk = rows_per_chunk;
  const AttnOutChunks oc = chunks != nullptr ? *chunks : AttnOutChunks{};
  const int items = a.heads * a.num_q_blocks;
  const int grid = items < sm_count() ? items : sm_count();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Fraction (of 8) of exponentials evaluated on the FMA pipe; VCOF_ATTN_EMU=0|3 overrides.  Measured on
  // B200 (profiles/r1_gpurun6_emu_sweep_vae_bench.log, L=75600, 8 heads): 0 -> 17.14 ms, 2 -> 17.55, 3 -> 18.02,
  // 4 -> 18.46: the FMA/ALU issue slots, not MUFU, are the scarcer resource here, so the default is 0.
  static int emu = -1;
  if (emu < 0) {
    const char* e = getenv("VCOF_ATTN_EMU");
    emu = e ? atoi(e) : 0;
    if (emu != 3) emu = 0;
  }
  auto launch = [&](auto kern) -> int {
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    kern<<<grid, kAttnThreads, kAttnSmem, st>>>(tmQ, tmK, tmV, a, oc);
    return 0;
  };
  if (chunks != nullptr) {     // push exchange: default arithmetic, natural V layout, scattered output rows
    VCOF_REQUIRE(!v_transposed, "vcof_attn_fwd_scatter: natural V layout only");
    const int lrc = launch(attn_fwd_kernel<false, 0, false, 2, 0, true>);
    if (lrc) return lrc;
    VCOF_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  // VCOF_ATTN_SPLIT_S=1 / VCOF_ATTN_PSPLIT=2|4: tuning knobs (measured defaults: profiles/README.md)
  static int split_s = -1, psplit = -1;
  if (split_s < 0) {
    const char* e = getenv("VCOF_ATTN_SPLIT_S");
    split_s = e ? (atoi(e) != 0) : 0;
    const char* q = getenv("VCOF_ATTN_PSPLIT");
    psplit = (q && atoi(q) == 4) ? 4 : 2;
  }
  static const int spec = [] {
    const char* e = getenv("VCOF_ATTN_SPEC");
    return e == nullptr ? 0 : (e[0] == '1' ? 1 : (e[0] == '2' ? 2 : 0));
  }();
  int lrc = 0;
  if (spec == 1 && emu == 0 && !split_s && psplit == 2) {   // experimental: row maximum off the critical path
    lrc = v_transposed ? launch(attn_fwd_kernel<true, 0, false, 2, 1>) : launch(attn_fwd_kernel<false, 0, false, 2, 1>);
  } else if (spec == 2 && emu == 0 && !split_s && psplit == 2) {   // ... and half of the score load under the exps
    lrc = v_transposed ? launch(attn_fwd_kernel<true, 0, false, 2, 2>) : launch(attn_fwd_kernel<false, 0, false, 2, 2>);
  } else if (emu == 3 || split_s) {            // experimental variants, natural-V layout only
    VCOF_REQUIRE(!v_transposed, "vcof_attn_fwd: tuning variants support the natural V layout only");
    if (emu == 3 && split_s) lrc = launch(attn_fwd_kernel<false, 3, true, 2>);
    else if (emu == 3) lrc = launch(attn_fwd_kernel<false, 3, false, 2>);
    else lrc = launch(attn_fwd_kernel<false, 0, true, 2>);
  } else if (v_transposed) {
    lrc = psplit == 4 ? launch(attn_fwd_kernel<true, 0, false, 4>) : launch(attn_fwd_kernel<true, 0, false, 2>);
  } else {
    lrc = psplit == 4 ? launch(attn_fwd_kernel<false, 0, false, 4>) : launch(attn_fwd_kernel<false, 0, false, 2>);
  }
  if (lrc) return lrc;
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_attn_fwd(const void* q, long long ldq, const void* k, long long ldk,
                             const void* v, long long ldv, void* out, long long ldo, int Lq, int Lk,
                             int kv_len, int heads, int head_dim, float softmax_scale,
                             int v_transposed, void* stream) {
  return attn_fwd_impl(q, ldq, k, ldk, v, ldv, out, ldo, Lq, Lk, kv_len, heads, head_dim, softmax_scale, v_transposed,
                       stream, nullptr, 0);
}

extern "C" int vcof_attn_fwd_scatter(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                     long long ldv, void* const* out_chunks, int n_chunks, int rows_per_chunk,
                                     long long ldo, int Lq, int Lk, int kv_len, int heads, int head_dim,
                                     float softmax_scale, void* stream) {
  VCOF_REQUIRE(out_chunks != nullptr && n_chunks >= 1 && n_chunks <= kMaxOutChunks,
               "vcof_attn_fwd_scatter: need 1..%d output chunks, got %d", kMaxOutChunks, n_chunks);
  VCOF_REQUIRE(rows_per_chunk > 0 && (long long)n_chunks * rows_per_chunk >= Lq,
               "vcof_attn_fwd_scatter: %d chunks of %d rows do not cover Lq=%d", n_chunks, rows_per_chunk, Lq);
  AttnOutChunks oc;
  for (int i = 0; i < kMaxOutChunks; ++i) oc.p[i] = nullptr;
  for (int i = 0; i < n_chunks; ++i) {
    VCOF_REQUIRE(out_chunks[i] != nullptr && (reinterpret_cast<uintptr_t>(out_chunks[i]) & 15) == 0,
                 "vcof_attn_fwd_scatter: output chunk %d is null or not 16-byte aligned", i);
    oc.p[i] = reinterpret_cast<bf16*>(out_chunks[i]);
  }
  return attn_fwd_impl(q, ldq, k, ldk, v, ldv, out_chunks[0], ldo, Lq, Lk, kv_len, heads, head_dim, softmax_scale, 0,
                       stream, &oc, rows_per_chunk);
}
