// vae_attn_sm100.cu — fused single-head attention with head dim 384 for the VAE's AttentionBlock on tcgen05.
//
// Replaces, per latent frame, the reference's F.scaled_dot_product_attention over h*w tokens with d = C = 384
// (videox_fun/models/wan_vae.py:244-266).  Round 1 ran it as GEMM (fp32 scores to HBM: 829 MB per 720p frame) ->
// softmax_rows -> GEMM, ~2.5 GB of HBM traffic per frame for a 3.2e11-FLOP op; here scores and probabilities never leave
// the SM.
//
// One persistent CTA per SM walks (frame, 128-query-row tile) items; keys come in blocks of 64:
//   TMEM (512 columns): S (64 fp32 columns; P, bf16, aliases its first 32) | O (384 fp32 columns = three N = 128 chunks)
//   smem: Q tile 128 x 384 (96 KB, resident per item) | K block 64 x 384 (48 KB) | V block 64 x 384 (48 KB), each as
//         six 64-column TMA boxes (128-byte swizzle) cut straight out of the to_qkv output [T, N, 3*384]
//   warps 0-3: softmax + epilogue (thread i <-> TMEM lane i <-> one query row)   warp 4: TMA producer
//   warp 5: tcgen05.mma issuer: S = Q K_j^T (24 SS MMAs, K = 384), then O_c += P V_j,c for the three column chunks (P from TMEM)
//   warp 6: TMEM allocator
// With d = 384 all 512 TMEM columns serve ONE tile, so there is no ping-pong partner: the tensor pipe idles during a
// block's exponentials (64 per row, against 36 MMAs per block), and K / V are single-buffered (K_{j+1} loads under the
// softmax and PV of block j, V_{j+1} under S_{j+1}).  Same online softmax as attn_sm100.cu: exp2 with folded scale,
// per-row lazy rescale (threshold 2^8).
#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

constexpr int kVD = 384;               // head dim = channels of the block
constexpr int kVQ = 128;               // query rows per tile
constexpr int kVK = 64;                // keys per block
constexpr int kVBoxes = kVD / 64;      // 64-column boxes per row block
constexpr int kVQBox = kVQ * 128;      // bytes of one Q box (128 rows x 128 B)
constexpr int kVKBox = kVK * 128;      // bytes of one K / V box (64 rows x 128 B)
constexpr int kVThreads = 256;
constexpr int kVSmem = kVBoxes * kVQBox + 2 * kVBoxes * kVKBox + 256 + 1024;
constexpr float kVRescale = 8.0f;      // log2 units

struct VaeAttnArgs {
  int N, T, q_tiles;
  float scale_log2;
  bf16* out;
  long long ldo;      // elements between output rows
  long long frame_o;  // elements between output frames
};

__device__ __forceinline__ float vex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kVThreads, 1)
vae_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, VaeAttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kVBoxes * kVQBox;
  uint8_t* sV = sK + kVBoxes * kVKBox;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVBoxes * kVKBox);
  const uint32_t b0 = smem_u32(bars);
  const uint32_t q_full = b0, q_empty = b0 + 8, k_full = b0 + 16, k_empty = b0 + 24, v_full = b0 + 32, v_empty = b0 + 40;
  const uint32_t s_full = b0 + 48, p_full = b0 + 56, o_full = b0 + 64, o_empty = b0 + 72;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    mbar_init(k_full, 1); mbar_init(k_empty, 1);
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 4);
    mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 6) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;            // 64 columns (P: the first 32)
  const uint32_t tO = tmem_base + 64;       // 384 columns

  const int n_kv = (p.N + kVK - 1) / kVK;
  const int items = p.T * p.q_tiles;

  if (warp == 4) {
    // ---------------- TMA producer (all lanes walk the loop, one elected lane issues) ----------------
    uint32_t qph = 0, kph = 0, vph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int frame = item / p.q_tiles;
      const int q0 = (item % p.q_tiles) * kVQ;
      mbar_wait(q_empty, qph ^ 1);
      qph ^= 1;
      if (elect_one()) {
        mbar_expect_tx(q_full, kVBoxes * kVQBox);
#pragma unroll
        for (int b = 0; b < kVBoxes; ++b) tma_load_3d(smem_u32(sQ + b * kVQBox), &tmQ, q_full, b * 64, q0, frame);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(k_empty, kph ^ 1);
        kph ^= 1;
        if (elect_one()) {
          mbar_expect_tx(k_full, kVBoxes * kVKBox);
#pragma unroll
          for (int b = 0; b < kVBoxes; ++b)
            tma_load_3d(smem_u32(sK + b * kVKBox), &tmKV, k_full, kVD + b * 64, j * kVK, frame);
        }
        __syncwarp();
        mbar_wait(v_empty, vph ^ 1);
        vph ^= 1;
        if (elect_one()) {
          mbar_expect_tx(v_full, kVBoxes * kVKBox);
#pragma unroll
          for (int b = 0; b < kVBoxes; ++b)
            tma_load_3d(smem_u32(sV + b * kVKBox), &tmKV, v_full, 2 * kVD + b * 64, j * kVK, frame);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_qk = make_idesc_bf16(kVQ, kVK, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kVQ, 128, false, true);
    uint32_t qph = 0, kph = 0, vph = 0, pph = 0, oeph = 0;
    int it = 0;
    auto mma_s = [&](bool release_q) {
      const uint64_t ad = make_desc_kmajor_sw128(smem_u32(sQ));
      const uint64_t bd = make_desc_kmajor_sw128(smem_u32(sK));
      if (elect_one()) {
#pragma unroll
        for (int b = 0; b < kVBoxes; ++b) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(tS, ad + ((b * kVQBox + ks * 32) >> 4), bd + ((b * kVKBox + ks * 32) >> 4), idesc_qk, (b | ks) != 0);
        }
        umma_commit(s_full);
        umma_commit(k_empty);
        if (release_q) umma_commit(q_empty);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      mbar_wait(q_full, qph);
      qph ^= 1;
      mbar_wait(k_full, kph);
      kph ^= 1;
      tc_fence_after();
      mma_s(n_kv == 1);
      for (int j = 0; j < n_kv; ++j) {
        const bool last = (j + 1 == n_kv);
        mbar_wait(v_full, vph);
        vph ^= 1;
        mbar_wait(p_full, pph);
        pph ^= 1;
        if (j == 0 && it > 0) {       // the previous item's epilogue has read O
          mbar_wait(o_empty, oeph);
          oeph ^= 1;
        }
        tc_fence_after();
        const uint32_t vb = smem_u32(sV);
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint64_t bd0 = make_desc_mnmajor_sw128(vb + 2 * c * kVKBox, kVKBox, 1024);
#pragma unroll
            for (int ks = 0; ks < kVK / 16; ++ks)
              umma_ts(tO + c * 128, tS + ks * 8, bd0 + ((ks * 16 * 128) >> 4), idesc_pv, (j != 0 || ks != 0));
          }
          umma_commit(v_empty);
          if (last) umma_commit(o_full);
        }
        __syncwarp();
        if (!last) {
          mbar_wait(k_full, kph);
          kph ^= 1;
          tc_fence_after();
          mma_s(j + 2 == n_kv);
        }
      }
    }
  } else if (warp < 4) {
    // ---------------- softmax + epilogue ----------------
    const uint32_t lane_base = (warp * 32u) << 16;
    const uint32_t rS = tS + lane_base, rO = tO + lane_base;
    uint32_t sph = 0, oph = 0;
    const int rem = p.N - (n_kv - 1) * kVK;     // valid keys of the last block (1..64)
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int frame = item / p.q_tiles;
      const int q0 = (item % p.q_tiles) * kVQ;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        uint32_t s[64];
        mbar_wait(s_full, sph);
        sph ^= 1;
        tc_fence_after();
        tmem_ld32(rS, s);
        tmem_ld32(rS + 32, s + 32);
        tmem_ld_wait();
        if (j == n_kv - 1 && rem < kVK) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= rem) s[c] = 0xff800000u;   // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
        }
        const float m_new = fmaxf(m_ref, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
        const bool grow = (m_new - m_ref) * p.scale_log2 > kVRescale;    // true at j == 0
        if (__any_sync(0xffffffffu, grow)) {
          // warp-wide branch (tcgen05.ld / st are collective), per-row decision: see attn_sm100.cu
          const float m_next = grow ? m_new : m_ref;
          if (j > 0) {
            // PV(j-1) retired before s_full fired (S(j) was issued after it, same thread): O is ours to rescale
            const float alpha = grow ? vex2((m_ref - m_next) * p.scale_log2) : 1.0f;
            l *= alpha;
#pragma unroll 1
            for (int c = 0; c < kVD / 32; ++c) {
              uint32_t o[32];
              tmem_ld32(rO + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st32(rO + c * 32, o);
            }
            tmem_st_wait();
          }
          m_ref = m_next;
        }
        const float mb = m_ref * p.scale_log2;
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float2 nmb2 = make_float2(-mb, -mb);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, nmb2);
          float2 pr;
          pr.x = vex2(x.x);
          pr.y = vex2(x.y);
          if (c & 1) acc1 = __fadd2_rn(acc1, pr); else acc0 = __fadd2_rn(acc0, pr);
          pk[c] = pack_bf16x2(pr.x, pr.y);
        }
        tmem_st32(rS, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        l += (acc0.x + acc1.x) + (acc0.y + acc1.y);
      }
      // ---- epilogue: O / l -> bf16 -> global
      mbar_wait(o_full, oph);
      oph ^= 1;
      tc_fence_after();
      const float inv = 1.0f / l;
      const int row = q0 + warp * 32 + lane;
      bf16* orow = p.out + (long long)frame * p.frame_o + (long long)row * p.ldo;
#pragma unroll 1
      for (int c = 0; c < kVD / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(rO + c * 32, o);
        tmem_ld_wait();
        if (row < p.N) {
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
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vcof

using namespace vcof;

extern "C" int vcof_vae_attn(const void* qkv, long long ld, void* out, long long ldo, int T, int N, int C,
                             float softmax_scale, void* stream) {
  VCOF_REQUIRE(C == kVD, "vcof_vae_attn: channel count %d unsupported (the Wan VAE attention block has %d)", C, kVD);
  VCOF_REQUIRE(T > 0 && N > 0, "vcof_vae_attn: empty problem T=%d N=%d", T, N);
  VCOF_REQUIRE(ld >= 3 * C && ld % 8 == 0 && ldo >= C && ldo % 8 == 0,
               "vcof_vae_attn: row pitches ld=%lld / ldo=%lld must be multiples of 8 and hold 3C / C columns", ld, ldo);
  VCOF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "vcof_vae_attn: out not 16B aligned");
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_3d_bf16(&tmQ, qkv, (uint64_t)3 * C, (uint64_t)N, (uint64_t)T, (uint64_t)ld * 2,
                             (uint64_t)N * ld * 2, 64, kVQ, 1);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tmKV, qkv, (uint64_t)3 * C, (uint64_t)N, (uint64_t)T, (uint64_t)ld * 2,
                         (uint64_t)N * ld * 2, 64, kVK, 1);
  if (rc) return rc;
  VaeAttnArgs a;
  a.N = N;
  a.T = T;
  a.q_tiles = (N + kVQ - 1) / kVQ;
  a.scale_log2 = softmax_scale * 1.4426950408889634f;
  a.out = reinterpret_cast<bf16*>(out);
  a.ldo = ldo;
  a.frame_o = (long long)N * ldo;
  static bool attr_set = false;
  if (!attr_set) {
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(vae_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVSmem));
    attr_set = true;
  }
  const long long items = (long long)T * a.q_tiles;
  const int grid = items < sm_count() ? (int)items : sm_count();
  vae_attn_kernel<<<grid, kVThreads, kVSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmKV, a);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
