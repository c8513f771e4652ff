// conv_sm100.cu — im2col-free implicit-GEMM convolution for the Wan 3D causal VAE on tcgen05 (sm_100a).
//
// Activations are channels-last bf16 [T, H, W, C].  An output tile is an 8 x 16 spatial patch of one
// output frame (128 positions = 128 TMEM lanes) times all (<= 384) output channels.  For every filter
// tap and every 32-channel slice, ONE TMA box load [32 c, 16 w, 1, 8 h, 1 t] of the *shifted* input
// patch lands in shared memory already in the canonical K-major 64B-swizzled UMMA layout (128 rows of
// 64 B), so no im2col buffer ever exists; out-of-range coordinates (spatial zero padding, the causal
// left padding in time, ragged edges) are zero-filled by the TMA unit.  Weights are pre-packed
// [Cout, taps * Cin] (tap-major K) and streamed by a second TMA map.
//
//   warp 4: TMA producer   warp 5: tcgen05.mma issuer (fp32 accumulators in TMEM, double-buffered
//   when Cout <= 256)      warps 0-3: epilogue (bias, optional residual, bf16 channels-last store)
//
// One kernel covers every convolution of the reference's VAE (videox_fun/models/wan_vae.py):
//   CausalConv3d 3x3x3 / 1x1x1 (:21-40), the per-frame Conv2d 3x3 of the resamplers (:80-100), the
//   stride-2 down-sampling Conv2d (5-D "parity" view of the input), the nearest-2x + Conv2d up-sampler
//   (four parity sub-convolutions with folded 2x2 taps; the 4x tensor is never materialised), and the
//   (3,1,1) temporal convs of up/down-sampling (:107-163) incl. the frame interleave on store.
#include <cstdlib>
#include <type_traits>
#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

constexpr int kCvThreads = 256;               // warps 0-3 epilogue, 4 TMA (A), 5 MMA, 6 TMA (B), 7 idle
constexpr int kCvMaxStages = 16;             // ring depth is chosen at launch: smem / (A + B bytes of this conv)
constexpr int kCvRows = 128;                 // positions per tile (= TMEM lanes)
constexpr int kCvData = 200 * 1024;          // operand ring budget
constexpr int kCvVecMax = 768;                // bias / gamma staged in shared memory (n_total <= 768)
constexpr int kCvSmem = kCvData + 512 + 2 * kCvVecMax * 4 + 1024;
constexpr int kMaxTaps = 27;

constexpr uint64_t kDescSwizzle64 = 4ull << 61;
__device__ __forceinline__ uint64_t make_desc_kmajor_sw64(uint32_t saddr) {
  // rows of 64 B, 8-row groups 512 B apart
  return kDescVersion1 | kDescSwizzle64 | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 16) |
         uint64_t((saddr & 0x3FFFF) >> 4);
}

struct ConvTap {
  short c_base;   // added to the channel coordinate (parity views: pw * Cin)
  short dw, p, dh, dt;
};

struct ConvArgs {
  ConvTap taps[kMaxTaps];
  int ntaps, cin_chunks;      // K loop = ntaps x cin_chunks slices of kc channels
  int cin;                    // padded input channels (multiple of kc): weight K stride per tap
  int kc;                     // channels per slice: 64 (128-byte rows, SW128) or 32 (64-byte rows, SW64).  The TMA
                              // unit's cost is per box ROW (profiles/r1_tma_probe.txt), so rows should be 128 B
                              // whenever the layer has >= 64 input channels
  int T_out, H_out, W_out;    // logical output grid the tiles walk over
  int t_stride;               // t_in = t_out * t_stride + dt
  int n_total, n_tile;        // output channels (padded to 16) and channels per CTA tile (<= 384)
  // store addressing: frame = t*ot_mul + ot_add (+1 for the second channel half when interleave),
  // row = h*oh_mul + oh_add, col = w*ow_mul + ow_add inside a [*, Hs, Ws, ldc] tensor
  int ot_mul, ot_add, oh_mul, oh_add, ow_mul, ow_add, Hs, Ws;
  long long ldc;
  int interleave_half;        // > 0: channels >= this go to frame+1 and are stored at (n - half)
  int n_store;                // channels actually stored per position (<= n_total)
  int stages, stage_bytes;    // TMA ring geometry chosen at launch
  int producers;              // 2: activation and weight boxes are issued by lanes of different warps (one lane
                              // sustains ~1 barrier round trip per ~520 clk, profiles/r1_tma_probe_v2.txt); 1: one lane
  int tgroup;                 // consecutive taps that differ only in dt (3 for k_t = 3, else 1): ONE 5-D box with
                              // t-extent tgroup and ONE rank-3 weight box feed tgroup taps — the TMA unit's cost
                              // is per box (profiles/r1_tma_probe.txt), so boxes must be as large as possible
  const float* bias;          // [n_total] fp32 or nullptr
  const bf16* residual;       // same addressing as out, or nullptr
  bf16* out;                  // raw output, may be nullptr when only act_out is wanted
  float clamp;                // > 0: clamp output to [-clamp, clamp]
  bf16* act_out;              // optional: silu(rms_norm(out) * act_gamma), same addressing as out
  const float* act_gamma;     // fp32 [n_store]
};

// Epilogue of ONE output position (this thread's TMEM lane, n_end accumulator columns starting at t_row):
// bias (+ residual, clamp) -> bf16 store, and optionally the next layer's RMS_norm + SiLU as a second output.
// `release_bar` (0 = none): mbarrier on which the accumulator is handed back to the MMA warp.  Returns true when the
// function arrived on it itself — on the path that keeps the whole channel row in registers this happens right after
// the last tcgen05.ld, BEFORE the activation sweep and the stores, so the next tile's MMAs into this accumulator
// overlap the epilogue's long tail (the fused RMS_norm + SiLU epilogue of a 96-channel row is ~2500 clk; released
// only at its end it stalled the line kernel's MMA warp for ~5000 clk per 31 000-clk work item).
__device__ __forceinline__ bool conv_epilogue_row(const ConvArgs& p, const float* sBias, const float* sGamma,
                                                  uint32_t t_row, int t, int h, int w, int n0, int n_end,
                                                  uint32_t release_bar = 0) {
  const bool ok = (h < p.H_out) && (w < p.W_out);
  const long long pos_row = (long long)(h * p.oh_mul + p.oh_add) * p.Ws + (w * p.ow_mul + p.ow_add);
  const int frame = t * p.ot_mul + p.ot_add;
  // one 32-channel chunk: accumulator + bias (+ residual, clamp) -> v[]; returns #valid channels.
  // Full chunks (cnt == 32, the common case) run without per-element predicates.
  auto chunk = [&](int c, float* v, long long& off, int& n) -> int {
    uint32_t rr[32];
    tmem_ld32(t_row + c, rr);
    tmem_ld_wait();
    n = n0 + c;
    int fr = frame, ns = n;
    if (p.interleave_half > 0 && n >= p.interleave_half) { fr += 1; ns = n - p.interleave_half; }
    off = ((long long)fr * p.Hs * p.Ws + pos_row) * p.ldc + ns;
    const int cnt = min(32, min(p.n_total - n, (p.interleave_half > 0 ? p.interleave_half : p.n_store) - ns));
    if (!ok) return 0;
    const float4* b4 = reinterpret_cast<const float4*>(sBias + n);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = b4[q];                       // channels beyond n_total read zeros of the padding
      v[q * 4 + 0] = __uint_as_float(rr[q * 4 + 0]) + b.x;
      v[q * 4 + 1] = __uint_as_float(rr[q * 4 + 1]) + b.y;
      v[q * 4 + 2] = __uint_as_float(rr[q * 4 + 2]) + b.z;
      v[q * 4 + 3] = __uint_as_float(rr[q * 4 + 3]) + b.w;
    }
    if (p.residual != nullptr) {
      const bf16* rp = p.residual + off;
      if (cnt == 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 u = *reinterpret_cast<const uint4*>(rp + q * 8);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h2[e]);
            // the reference adds two bf16 tensors: round the conv output first
            v[q * 8 + e * 2] = bf16_round(v[q * 8 + e * 2]) + f.x;
            v[q * 8 + e * 2 + 1] = bf16_round(v[q * 8 + e * 2 + 1]) + f.y;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < cnt) v[j] = bf16_round(v[j]) + __bfloat162float(rp[j]);
      }
    }
    if (p.clamp > 0.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -p.clamp), p.clamp);
    }
    return cnt;
  };
  auto store = [&](bf16* o, const float* v, int cnt) {
    if (cnt == 32) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
        u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
        u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
        u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(o + q * 8) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) o[j] = __float2bfloat16_rn(v[j]);
    }
  };
  // activation of one chunk: silu(r * inv * gamma) with r the bf16-rounded result (what the next layer's
  // RMS_norm reads); fp32 intermediates, ONE rounding at the store
  auto activate = [&](float* v, int n, int cnt, float inv) {
    const float4* g4 = reinterpret_cast<const float4*>(sGamma + n);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 g = g4[q];
      const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = bf16_round(v[q * 4 + e]) * inv * gg[e];
        // silu(a) = a * sigmoid(a) = a * (0.5 + 0.5 * tanh(a / 2)): one MUFU op (tanh.approx, rel. error 2^-11, well
        // under the bf16 store's 2^-9) instead of the two of a / (1 + exp(-a))
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * a));
        v[q * 4 + e] = a * fmaf(0.5f, th, 0.5f);
      }
    }
    (void)cnt;
  };
  if (p.act_out == nullptr) {
#pragma unroll 1
    for (int c = 0; c < n_end; c += 32) {
      float v[32];
      long long off;
      int n;
      const int cnt = chunk(c, v, off, n);
      if (cnt > 0) store(p.out + off, v, cnt);
    }
  } else if (n_end <= 128) {
    // fused RMS_norm + SiLU of the NEXT layer (wan_vae.py:43-58, 197-201): the thread owns the whole channel
    // row (n_tile == n_total); up to 128 channels stay in registers between the two sweeps
    float v[4][32];
    long long off[4];
    int nn[4], cnt[4];
    float sq = 0.f;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      cnt[ci] = 0;
      if (ci * 32 < n_end) {
        cnt[ci] = chunk(ci * 32, v[ci], off[ci], nn[ci]);
        if (cnt[ci] > 0) {
          if (p.out != nullptr) store(p.out + off[ci], v[ci], cnt[ci]);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < cnt[ci]) { const float r = bf16_round(v[ci][j]); sq += r * r; }
        }
      }
    }
    if (release_bar != 0) {          // every accumulator column of this row is in registers now (warp-uniform branch)
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(release_bar);
    }
    const float inv = sqrtf(float(p.n_store)) / fmaxf(sqrtf(sq), 1e-12f);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      if (cnt[ci] > 0) {
        activate(v[ci], nn[ci], cnt[ci], inv);
        store(p.act_out + off[ci], v[ci], cnt[ci]);
      }
    }
    return release_bar != 0;
  } else {
    // wider rows: second sweep over TMEM instead of 384 live registers
    float sq = 0.f;
#pragma unroll 1
    for (int c = 0; c < n_end; c += 32) {
      float v[32];
      long long off;
      int n;
      const int cnt = chunk(c, v, off, n);
      if (cnt <= 0) continue;
      if (p.out != nullptr) store(p.out + off, v, cnt);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) { const float r = bf16_round(v[j]); sq += r * r; }
    }
    const float inv = sqrtf(float(p.n_store)) / fmaxf(sqrtf(sq), 1e-12f);
#pragma unroll 1
    for (int c = 0; c < n_end; c += 32) {
      float v[32];
      long long off;
      int n;
      const int cnt = chunk(c, v, off, n);
      if (cnt <= 0) continue;
      activate(v, n, cnt, inv);
      store(p.act_out + off, v, cnt);
    }
  }
  return false;
}

__global__ void __launch_bounds__(kCvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ ConvArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvData);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kCvMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kCvMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kCvMaxStages + 4);
  float* sBias = reinterpret_cast<float*>(smem + kCvData + 512);
  float* sGamma = sBias + kCvVecMax;
  const int kCvStages = p.stages;
  const int kCvStage = p.stage_bytes;
  for (int i = threadIdx.x; i < kCvVecMax; i += kCvThreads) {
    sBias[i] = (p.bias != nullptr && i < p.n_total) ? __ldg(p.bias + i) : 0.f;
    sGamma[i] = (p.act_gamma != nullptr && i < p.n_store) ? __ldg(p.act_gamma + i) : 0.f;
  }

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 5) {
    if (lane == 0) {
      for (int i = 0; i < kCvMaxStages; ++i) {
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
    tmem_alloc(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_w = (p.W_out + 15) / 16, tiles_h = (p.H_out + 7) / 8;
  const int n_tiles = (p.n_total + p.n_tile - 1) / p.n_tile;
  const int num_tiles = p.T_out * tiles_h * tiles_w * n_tiles;
  const int ngroups = p.ntaps / p.tgroup;
  const int num_k = ngroups * p.cin_chunks;
  const int nacc = (p.n_tile <= 256) ? 2 : 1;            // accumulator buffers in TMEM
  const int nsub = (p.n_tile <= 256) ? 1 : 2;            // MMAs per k-step (N <= 256 each)
  const int n_sub = p.n_tile / nsub;
  const uint32_t row_bytes = p.kc * 2;
  const uint32_t a_bytes = kCvRows * row_bytes;
  const uint32_t b_bytes = p.n_tile * row_bytes;

  auto decode = [&](int tile, int& t, int& h0, int& w0, int& n0) {
    int nt = tile % n_tiles;
    int s = tile / n_tiles;
    w0 = (s % tiles_w) * 16;
    s /= tiles_w;
    h0 = (s % tiles_h) * 8;
    t = s / tiles_h;
    n0 = nt * p.n_tile;
  };

  if (warp == 4 || (warp == 6 && p.producers == 2)) {
    // All lanes walk the loop (warp-uniform operands -> uniform registers), one elected lane issues: see gemm_sm100.cu.
    {
      const bool load_a = warp == 4;
      const bool load_b = (warp == 6) || (p.producers == 1);
      const uint32_t tx = p.tgroup * ((load_a ? a_bytes : 0) + (load_b ? b_bytes : 0));
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t b_off_bytes = p.tgroup * a_bytes;
      const uint32_t b_sub_bytes = p.tgroup * n_sub * row_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t, h0, w0, n0;
        decode(tile, t, h0, w0, n0);
        for (int g = 0; g < ngroups; ++g) {
          const ConvTap tp = p.taps[g * p.tgroup];            // first tap of the group (lowest dt)
          // everything the issue needs is computed BEFORE the wait: the instructions between a barrier wake-up and the
          // TMA issue sit on the ring's slot cycle (single-lane code, ~5 clk per dependent op)
          const int cw = w0 + tp.dw, chh = h0 + tp.dh, ct = t * p.t_stride + tp.dt;
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            const uint32_t st = smem_base + s * kCvStage;
            const uint32_t bar = bar_full + 8 * s;
            const int cch = tp.c_base + cc * p.kc;
            const int slice = (g * p.cin_chunks + cc) * p.tgroup;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bar, tx);
              // [kc c, 16 w, 1, 8 h, tgroup t] -> tgroup consecutive 128-row K-major tiles
              if (load_a) tma_load_5d(st, &tmX, bar, cch, cw, tp.p, chh, ct);
              if (load_b) {
                tma_load_3d(st + b_off_bytes, &tmW, bar, 0, n0, slice);
                if (nsub == 2) tma_load_3d(st + b_off_bytes + b_sub_bytes, &tmW, bar, 0, n0 + n_sub, slice);
              }
            }
            __syncwarp();
            if (++s == kCvStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    {
      // The issue loop is descriptor += constant only: all smem addresses are < 256 KB and 16-byte multiples, hence an
      // offset is a plain add on the descriptor's 14-bit address field (uniform-register arithmetic: the whole warp
      // walks the loop, one elected lane issues).
      const uint32_t idesc = make_idesc_bf16(128, n_sub, false, false);
      const bool wide = p.kc == 64;
      const uint64_t dflags = wide ? (kDescVersion1 | kDescSwizzle128 | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 16))
                                   : (kDescVersion1 | kDescSwizzle64 | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 16));
      const uint64_t desc0 = dflags | uint64_t((smem_u32(smem) & 0x3FFFF) >> 4);
      const uint32_t stage_step = kCvStage >> 4;
      const uint32_t a_step = a_bytes >> 4;                         // next temporal tap of the A box
      const uint32_t b_off = (p.tgroup * a_bytes) >> 4;             // B tiles follow the tgroup A tiles
      const uint32_t b_step = (n_sub * row_bytes) >> 4;             // next temporal tap of a B sub-tile
      const int tg = p.tgroup;
      auto issue_stage = [&](auto ks_tag, uint64_t a_desc, uint32_t d_tmem, bool first) {
        constexpr int kSteps = decltype(ks_tag)::value;
        uint64_t b_desc = a_desc + b_off;
        for (int dt = 0; dt < tg; ++dt) {
#pragma unroll
          for (int ks = 0; ks < kSteps; ++ks) {
            umma_ss(d_tmem, a_desc + ks * 2, b_desc + ks * 2, idesc, !(first && dt == 0 && ks == 0));
            if (nsub == 2)
              umma_ss(d_tmem + n_sub, a_desc + ks * 2, b_desc + tg * b_step + ks * 2, idesc,
                      !(first && dt == 0 && ks == 0));
          }
          a_desc += a_step;
          b_desc += b_step;
        }
      };
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = (nacc == 2) ? (it & 1) : 0;
        const uint32_t acc_ph = (nacc == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int k = 0; k < num_k; ++k) {
          const uint64_t a_desc = desc0 + uint32_t(s) * stage_step;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            if (wide) issue_stage(std::integral_constant<int, 4>{}, a_desc, d_tmem, k == 0);
            else issue_stage(std::integral_constant<int, 2>{}, a_desc, d_tmem, k == 0);
            umma_commit(bar_empty + 8 * s);
            if (k == num_k - 1) umma_commit(bar_tfull + 8 * acc);
          }
          __syncwarp();
          if (++s == kCvStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int t, h0, w0, n0;
      decode(tile, t, h0, w0, n0);
      const int acc = (nacc == 2) ? (it & 1) : 0;
      const uint32_t acc_ph = (nacc == 2) ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tc_fence_after();
      const int r = warp * 32 + lane;
      const int h = h0 + (r >> 4), w = w0 + (r & 15);
      const bool released = conv_epilogue_row(p, sBias, sGamma, tmem_base + ((warp * 32u) << 16) + acc * 256, t, h, w,
                                              n0, min(p.n_tile, p.n_total - n0), bar_tempty + 8 * acc);
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// Line-resident variant for stride-1 3x3(x3) convolutions (vae.py uses it for layers of up to 128 output channels;
// VCOF_CONV_LINES=0|1 overrides).  Parity-green on B200 against the same contract as conv_igemm_kernel
// (tests/test_vae_gpu.py, profiles/r1_gpurun36_final_validation_lines.log); first version, untuned.
//
// conv_igemm_kernel re-fetches the shifted input patch for every filter tap and the weights for every tile: 146 B of
// operands per MMA-clock for 96 -> 96 against ~40 B/clk/SM of feed.  Here a tile is ONE output row segment of 128
// pixels and a CTA works on R consecutive rows at a time (R accumulators side by side in TMEM):
//   * an input LINE (130 haloed pixels x 32 channels of one row of one frame) is fetched once per (channel slice,
//     temporal tap) and serves up to nine taps: the three horizontal taps are row-shifted descriptor views of the same
//     swizzled buffer (valid: profiles/r1_umma_shift_probe.txt), the three vertical taps feed the accumulators of the
//     rows above / below;
//   * the nine weight tiles of a (slice, temporal tap) phase stay in shared memory for all R rows (double-buffered
//     across phases).
// 96 -> 96, R = 4: 105 KB per phase for 3456 MMA-clocks = 30 B/clk/SM.
//   warp 4: line producer   warp 6: weight producer   warp 5: MMA issuer   warps 0-3: epilogue, one row at a time
// ---------------------------------------------------------------------------
constexpr int kLnPix = 130;                          // 128 outputs + one halo pixel each side
constexpr int kLnBytes = 17 * 512;                   // 130 x 64 B rounded up to the SW64 repeat (8704)
constexpr int kLnMaxRows = 4;
constexpr int kLnMaxAcc = 5;                         // accumulators in the TMEM ring (>= rows: the spare ones absorb the epilogue's burst)
constexpr int kLnMaxRing = 8;

struct LineArgs {
  ConvArgs c;          // geometry / epilogue fields (taps unused); c.n_tile = channels per N pass
  int kt, t0;          // temporal taps and the offset of the first one: t_in = t + t0 + dt
  int rows;            // R: output rows per CTA work item
  int naccs;           // accumulators in the TMEM ring: min(kLnMaxAcc, 512 / roundup32(n_tile)) >= rows
  int ring;            // line ring depth
};

__global__ void __launch_bounds__(kCvThreads, 1)
conv_lines_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ LineArgs a) {
  const ConvArgs& p = a.c;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvData);
  const uint32_t bar_wfull = smem_u32(bars);                     // [2]
  const uint32_t bar_wempty = bar_wfull + 16;                    // [2]
  const uint32_t bar_lfull = bar_wempty + 16;                    // [kLnMaxRing]
  const uint32_t bar_lempty = bar_lfull + 8 * kLnMaxRing;        // [kLnMaxRing]
  const uint32_t bar_tfull = bar_lempty + 8 * kLnMaxRing;        // [kLnMaxAcc]
  const uint32_t bar_tempty = bar_tfull + 8 * kLnMaxAcc;         // [kLnMaxAcc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * kLnMaxRing + 2 * kLnMaxAcc);
  float* sBias = reinterpret_cast<float*>(smem + kCvData + 512);
  float* sGamma = sBias + kCvVecMax;
  for (int i = threadIdx.x; i < kCvVecMax; i += kCvThreads) {
    sBias[i] = (p.bias != nullptr && i < p.n_total) ? __ldg(p.bias + i) : 0.f;
    sGamma[i] = (p.act_gamma != nullptr && i < p.n_store) ? __ldg(p.act_gamma + i) : 0.f;
  }
  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int R = a.rows;
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmX);
  if (warp == 6 && lane == 0) tma_prefetch_desc(&tmW);
  if (warp == 5) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, 1); }
      for (int i = 0; i < kLnMaxRing; ++i) { mbar_init(bar_lfull + 8 * i, 1); mbar_init(bar_lempty + 8 * i, 1); }
      for (int i = 0; i < kLnMaxAcc; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int w_bytes = 9 * p.n_tile * 64;                         // nine [n_tile x 32 ch] tiles of one phase
  const int acc_stride = (p.n_tile + 31) & ~31;                  // TMEM columns per row accumulator
  // Accumulator RING: row j of this CTA's i-th work item uses accumulator (i * R + j) % nacc.  With nacc > R the first
  // rows of the next item land in accumulators that were drained long ago, so the MMA warp no longer waits while the
  // epilogue works through the R rows that all complete within an item's last phase (measured: 27 % of the MMA warp's
  // time with R = nacc = 4, profiles/r2_conv_lines_ncu_summary.txt).
  const int nacc = a.naccs;
  const uint32_t smem_w = smem_u32(smem);                        // two weight buffers
  const uint32_t smem_l = smem_w + 2 * w_bytes;                  // then the line ring (w_bytes is a multiple of 512)
  const int wsegs = (p.W_out + 127) / 128, hgroups = (p.H_out + R - 1) / R;
  const int n_passes = (p.n_total + p.n_tile - 1) / p.n_tile;
  const int num_items = p.T_out * hgroups * wsegs * n_passes;
  const int phases = p.cin_chunks * a.kt;
  auto decode = [&](int item, int& t, int& h0, int& w0, int& n0) {
    n0 = (item % n_passes) * p.n_tile;
    int s2 = item / n_passes;
    w0 = (s2 % wsegs) * 128;
    s2 /= wsegs;
    h0 = (s2 % hgroups) * R;
    t = s2 / hgroups;
  };

  if (warp == 4) {
    {                           // ---- line producer (all lanes walk the loop, one elected lane issues) ----
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int t, h0, w0, n0;
        decode(item, t, h0, w0, n0);
        for (int cc = 0; cc < p.cin_chunks; ++cc)
          for (int dt = 0; dt < a.kt; ++dt) {
            const int ct = t + a.t0 + dt;
            for (int r = 0; r < R + 2; ++r) {
              const uint32_t dst = smem_l + s * kLnBytes;
              mbar_wait(bar_lempty + 8 * s, ph ^ 1);
              if (elect_one()) {
                mbar_expect_tx(bar_lfull + 8 * s, kLnPix * 64);
                tma_load_5d(dst, &tmX, bar_lfull + 8 * s, cc * 32, w0 - 1, 0, h0 - 1 + r, ct);
              }
              __syncwarp();
              if (++s == a.ring) { s = 0; ph ^= 1; }
            }
          }
      }
    }
  } else if (warp == 6) {
    {                           // ---- weight producer ----
      int b = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int t, h0, w0, n0;
        decode(item, t, h0, w0, n0);
        for (int phase = 0; phase < phases; ++phase) {
          mbar_wait(bar_wempty + 8 * b, ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar_wfull + 8 * b, w_bytes);
            tma_load_3d(smem_w + b * w_bytes, &tmW, bar_wfull + 8 * b, 0, n0, phase * 9);
          }
          __syncwarp();
          if (++b == 2) { b = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    {                           // ---- MMA issuer (all lanes walk the loop, one elected lane issues per line) ----
      const uint32_t idesc = make_idesc_bf16(128, p.n_tile, false, false);
      const uint64_t dflags = kDescVersion1 | kDescSwizzle64 | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 16);
      const uint64_t wdesc0 = dflags | uint64_t((smem_w & 0x3FFFF) >> 4);
      const uint64_t ldesc0 = dflags | uint64_t((smem_l & 0x3FFFF) >> 4);
      const uint32_t w_tile = (p.n_tile * 64) >> 4;              // one (dh, dw) weight tile, in 16-byte units
      int s = 0, b = 0;
      uint32_t lph = 0, wph = 0;
      int g0 = 0;               // (first row of this item) % nacc, advanced without divisions
      uint32_t use_bits = 0;    // bit k = parity of accumulator k's next use
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int accs[kLnMaxRows];
        uint32_t par = 0;       // bit j = parity of row j's use of its accumulator
#pragma unroll
        for (int j = 0; j < kLnMaxRows; ++j) {
          int k = g0 + j;
          if (k >= nacc) k -= nacc;
          accs[j] = k;
          if (j < R) {
            par |= ((use_bits >> k) & 1u) << j;
            use_bits ^= 1u << k;
          }
        }
        g0 += R;
        if (g0 >= nacc) g0 -= nacc;
        for (int phase = 0; phase < phases; ++phase) {
          const uint64_t wdesc = wdesc0 + uint32_t(b) * (uint32_t(w_bytes) >> 4);
          mbar_wait(bar_wfull + 8 * b, wph);
          // unrolled over the (at most six) lines of a phase: the row a line feeds through tap dh and that row's
          // accumulator are then compile-time register picks instead of per-line select chains in the issuing warp
#pragma unroll
          for (int r = 0; r < kLnMaxRows + 2; ++r) {
            if (r >= R + 2) break;
            const uint64_t ldesc = ldesc0 + uint32_t(s) * (kLnBytes >> 4);
            mbar_wait(bar_lfull + 8 * s, lph);
            if (phase == 0 && r < R) {
              // first MMA ever into row r's accumulator (through dh = 0): its previous user has been drained
              mbar_wait(bar_tempty + 8 * accs[r < kLnMaxRows ? r : 0], ((par >> r) & 1u) ^ 1u);
            }
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int dh = 0; dh < 3; ++dh) {
                const int j = r - dh;                              // output row this line feeds through vertical tap dh
                if (j < 0 || j >= kLnMaxRows || j >= R) continue;
                const bool first = (phase == 0) && (dh == 0);      // first MMA ever into accumulator j of this item
                const int acc = accs[j];
                const uint32_t d_tmem = tmem_base + acc * acc_stride;
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) {
                  // horizontal tap = the same line read dw pixels (64-byte rows) further in
                  const uint64_t ad = ldesc + dw * 4;
                  const uint64_t bd = wdesc + (dh * 3 + dw) * w_tile;
                  umma_ss(d_tmem, ad, bd, idesc, !(first && dw == 0));
                  umma_ss(d_tmem, ad + 2, bd + 2, idesc, 1);
                }
                if (phase == phases - 1 && dh == 2) umma_commit(bar_tfull + 8 * acc);   // row j is complete
              }
              umma_commit(bar_lempty + 8 * s);
              if (r == R + 1) umma_commit(bar_wempty + 8 * b);
            }
            __syncwarp();
            if (++s == a.ring) { s = 0; lph ^= 1; }
          }
          if (++b == 2) { b = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp < 4) {        // ---- epilogue: rows drain one by one as they complete ----
    int g0 = 0;
    uint32_t use_bits = 0;      // same ring bookkeeping as the MMA warp
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int t, h0, w0, n0;
      decode(item, t, h0, w0, n0);
      const int n_end = min(p.n_tile, p.n_total - n0);
      const int w = w0 + warp * 32 + lane;
      for (int j = 0; j < R; ++j) {
        int acc = g0 + j;
        if (acc >= nacc) acc -= nacc;
        const uint32_t parity = (use_bits >> acc) & 1u;
        use_bits ^= 1u << acc;
        mbar_wait(bar_tfull + 8 * acc, parity);
        tc_fence_after();
        bool released = false;
        if (h0 + j < p.H_out)
          released = conv_epilogue_row(p, sBias, sGamma, tmem_base + ((warp * 32u) << 16) + acc * acc_stride, t, h0 + j, w,
                                       n0, n_end, bar_tempty + 8 * acc);
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
      }
      g0 += R;
      if (g0 >= nacc) g0 -= nacc;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// channels-last helpers (HBM-bound)
// ---------------------------------------------------------------------------
// y = silu(x / max(||x||_2, 1e-12) * sqrt(C) * gamma) per position (wan_vae.py:43-58 + nn.SiLU);
// one warp per position; with silu == 0 the activation is skipped (attention block norm).
__global__ void rms_silu_cl_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                   bf16* __restrict__ y, long long npos, int C, long long ldx,
                                   long long ldy, int silu) {
  const long long pos = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pos >= npos) return;
  const __nv_bfloat162* xr = reinterpret_cast<const __nv_bfloat162*>(x + pos * ldx);
  __nv_bfloat162* yr = reinterpret_cast<__nv_bfloat162*>(y + pos * ldy);
  const int np = C >> 1;
  float2 v[6];  // C <= 384
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int idx = lane + i * 32;
    if (idx < np) {
      v[i] = __bfloat1622float2(xr[idx]);
      sq += v[i].x * v[i].x + v[i].y * v[i].y;
    }
  }
  sq = warp_sum(sq);
  // x / max(||x||, 1e-12) * sqrt(C) * gamma with fp32 intermediates and ONE rounding at the store (the
  // reference's ATen chain rounds to bf16 after every op; fewer roundings only move us closer to fp32)
  const float inv = sqrtf(float(C)) / fmaxf(sqrtf(sq), 1e-12f);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int idx = lane + i * 32;
    if (idx < np) {
      float a = v[i].x * inv * gamma[2 * idx];
      float b = v[i].y * inv * gamma[2 * idx + 1];
      if (silu) {
        a = a / (1.f + __expf(-a));
        b = b / (1.f + __expf(-b));
      }
      yr[idx] = __floats2bfloat162_rn(a, b);
    }
  }
}

// [C, T, H, W] (bf16) -> [T, H, W, Cp] channels-last, channels >= C zero-filled; optional per-channel
// affine x / div[c] + add[c] (latent de-normalisation z / (1/std) + mean, wan_vae.py:553-558).
__global__ void nchw_to_cl_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C, int Cp,
                                  long long thw, const float* __restrict__ mul,
                                  const float* __restrict__ add) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= thw * Cp) return;
  const int c = int(i % Cp);
  const long long pos = i / Cp;
  float v = 0.f;
  if (c < C) {
    v = __bfloat162float(x[(long long)c * thw + pos]);
    if (mul != nullptr) v = bf16_round(bf16_round(v / mul[c]) + add[c]);
  }
  y[i] = __float2bfloat16_rn(v);
}

// channels-last [T, H, W, ld] -> [C, T, H, W]; optional per-channel affine (x - sub[c]) * mul[c]
// (latent normalisation of mu, wan_vae.py:540-546).
__global__ void cl_to_nchw_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C,
                                  long long ld, long long thw, const float* __restrict__ sub,
                                  const float* __restrict__ mul) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= thw * C) return;
  const long long pos = i % thw;
  const int c = int(i / thw);
  float v = __bfloat162float(x[pos * ld + c]);
  if (sub != nullptr) v = bf16_round(bf16_round(v - sub[c]) * mul[c]);
  y[i] = __float2bfloat16_rn(v);
}

// row softmax of fp32 scores -> bf16 probabilities (VAE single-head attention, wan_vae.py:251-256)
__global__ void softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int n,
                                    long long lds, long long ldp, float scale) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* sr = s + row * lds;
  bf16* pr = p + row * ldp;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, sr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = ((threadIdx.x & 31) < (blockDim.x >> 5)) ? red[threadIdx.x & 31] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  mx = __shfl_sync(0xffffffffu, mx, 0);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += __expf((sr[i] - mx) * scale);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = ((threadIdx.x & 31) < (blockDim.x >> 5)) ? red[threadIdx.x & 31] : 0.f;
  sum = warp_sum(sum);
  sum = __shfl_sync(0xffffffffu, sum, 0);
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    pr[i] = __float2bfloat16_rn(__expf((sr[i] - mx) * scale) * inv);
}

}  // namespace vcof

using namespace vcof;

extern "C" int vcof_conv_igemm(const void* x, const long long* x_dims, const long long* x_strides,
                               const void* w, int k_total, const short* taps, int ntaps, int tgroup, int cin,
                               const int* geom, const float* bias, const void* residual, void* out,
                               long long ldc, float clamp, void* act_out, const float* act_gamma,
                               void* stream) {
  // x_dims[5]: (c_inner, W, P, H, T) of the (parity-)view; x_strides[4]: element strides of dims 1..4
  // geom[17]: T_out,H_out,W_out,t_stride,n_total,n_tile,ot_mul,ot_add,oh_mul,oh_add,ow_mul,ow_add,Hs,Ws
  //           then [14] interleave_half, [15] n_store, [16] kc (channels per K slice: 32 or 64)
  VCOF_REQUIRE(ntaps >= 1 && ntaps <= kMaxTaps, "vcof_conv_igemm: ntaps %d outside [1,%d]", ntaps, kMaxTaps);
  const int kc = geom[16];
  VCOF_REQUIRE(kc == 32 || kc == 64, "vcof_conv_igemm: slice width %d must be 32 or 64", kc);
  VCOF_REQUIRE(cin % kc == 0 && cin > 0, "vcof_conv_igemm: cin %d must be a positive multiple of %d", cin, kc);
  ConvArgs a;
  for (int i = 0; i < ntaps; ++i) {
    a.taps[i].c_base = taps[i * 5 + 0];
    a.taps[i].dw = taps[i * 5 + 1];
    a.taps[i].p = taps[i * 5 + 2];
    a.taps[i].dh = taps[i * 5 + 3];
    a.taps[i].dt = taps[i * 5 + 4];
  }
  a.ntaps = ntaps;
  a.cin = cin;
  a.cin_chunks = cin / kc;
  a.kc = kc;
  a.T_out = geom[0]; a.H_out = geom[1]; a.W_out = geom[2]; a.t_stride = geom[3];
  a.n_total = geom[4]; a.n_tile = geom[5];
  a.ot_mul = geom[6]; a.ot_add = geom[7]; a.oh_mul = geom[8]; a.oh_add = geom[9];
  a.ow_mul = geom[10]; a.ow_add = geom[11]; a.Hs = geom[12]; a.Ws = geom[13];
  a.interleave_half = geom[14]; a.n_store = geom[15];
  a.ldc = ldc;
  a.bias = bias;
  a.residual = reinterpret_cast<const bf16*>(residual);
  a.out = reinterpret_cast<bf16*>(out);
  a.clamp = clamp;
  a.act_out = reinterpret_cast<bf16*>(act_out);
  a.act_gamma = act_gamma;
  VCOF_REQUIRE(out != nullptr || act_out != nullptr, "vcof_conv_igemm: no output requested");
  VCOF_REQUIRE(act_out == nullptr || (act_gamma != nullptr && a.n_tile == a.n_total && a.interleave_half == 0),
               "vcof_conv_igemm: fused norm needs gamma, a single channel tile and no frame interleave");
  VCOF_REQUIRE(a.n_total <= kCvVecMax, "vcof_conv_igemm: n_total %d > %d", a.n_total, kCvVecMax);
  VCOF_REQUIRE(a.n_total % 16 == 0 && a.n_tile % 16 == 0 && a.n_tile <= 384 && a.n_tile >= 16,
               "vcof_conv_igemm: n_total %d / n_tile %d must be multiples of 16, n_tile <= 384", a.n_total,
               a.n_tile);
  VCOF_REQUIRE(a.n_tile <= 256 || (a.n_tile / 2) % 16 == 0, "vcof_conv_igemm: n_tile/2 must be a multiple of 16");
  VCOF_REQUIRE(tgroup >= 1 && tgroup <= 3 && ntaps % tgroup == 0, "vcof_conv_igemm: bad tap group %d", tgroup);
  for (int g = 0; g < ntaps / tgroup; ++g)
    for (int j = 1; j < tgroup; ++j) {
      const ConvTap &u = a.taps[g * tgroup], &v = a.taps[g * tgroup + j];
      VCOF_REQUIRE(v.c_base == u.c_base && v.dw == u.dw && v.p == u.p && v.dh == u.dh && v.dt == u.dt + j,
                   "vcof_conv_igemm: taps of group %d must differ only by consecutive dt", g);
    }
  a.tgroup = tgroup;
  {
    static const int producers = [] {
      const char* e = getenv("VCOF_CONV_PRODUCERS");
      return (e != nullptr && e[0] == '1') ? 1 : 2;
    }();
    a.producers = producers;
  }
  a.stage_bytes = tgroup * (kCvRows + a.n_tile) * kc * 2;    // multiples of 1 KB keep the swizzled tiles aligned
  a.stages = kCvData / a.stage_bytes;
  VCOF_REQUIRE(a.stages >= 2, "vcof_conv_igemm: stage of %d bytes leaves no ring", a.stage_bytes);
  if (a.stages > kCvMaxStages) a.stages = kCvMaxStages;
  VCOF_REQUIRE(k_total == ntaps * cin, "vcof_conv_igemm: weight K %d != ntaps*cin %d", k_total, ntaps * cin);
  VCOF_REQUIRE(ldc % 8 == 0, "vcof_conv_igemm: ldc must be a multiple of 8");
  CUtensorMap tmX, tmW;
  uint64_t dims[5], strides[4];
  for (int i = 0; i < 5; ++i) dims[i] = (uint64_t)x_dims[i];
  for (int i = 0; i < 4; ++i) strides[i] = (uint64_t)x_strides[i] * 2;
  const uint32_t box[5] = {(uint32_t)kc, 16, 1, 8, (uint32_t)tgroup};
  int rc = make_tmap_nd_bf16(&tmX, x, 5, dims, strides, box, kc * 2);
  if (rc) return rc;
  const int nsub = a.n_tile <= 256 ? 1 : 2;
  // weights: [slices, n_total, kc] with slice = ((tap_group * cin_chunks + chunk) * tgroup + j)
  uint64_t wd[3] = {(uint64_t)kc, (uint64_t)a.n_total, (uint64_t)(k_total / kc)};
  uint64_t ws[2] = {(uint64_t)kc * 2, (uint64_t)a.n_total * kc * 2};
  uint32_t wb[3] = {(uint32_t)kc, (uint32_t)(a.n_tile / nsub), (uint32_t)tgroup};
  rc = make_tmap_nd_bf16(&tmW, w, 3, wd, ws, wb, kc * 2);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kCvSmem));
    attr_set = true;
  }
  const long long tiles = (long long)a.T_out * ((a.H_out + 7) / 8) * ((a.W_out + 15) / 16) *
                          ((a.n_total + a.n_tile - 1) / a.n_tile);
  VCOF_REQUIRE(tiles > 0 && tiles < (1ll << 31), "vcof_conv_igemm: bad tile count %lld", tiles);
  const int grid = tiles < sm_count() ? (int)tiles : sm_count();
  conv_igemm_kernel<<<grid, kCvThreads, kCvSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tmX, tmW, a);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_conv_lines(const void* x, const long long* x_dims, const long long* x_strides, const void* w,
                               int cin, int kt, int t0, const int* geom, const float* bias, const void* residual,
                               void* out, long long ldc, float clamp, void* act_out, const float* act_gamma,
                               void* stream) {
  // geom[7]: T_out, H_out, W_out, n_total, n_tile, rows, n_store
  VCOF_REQUIRE(cin > 0 && cin % 32 == 0, "vcof_conv_lines: cin %d must be a positive multiple of 32", cin);
  VCOF_REQUIRE(kt == 1 || kt == 3, "vcof_conv_lines: kt %d must be 1 or 3", kt);
  LineArgs a;
  ConvArgs& c = a.c;
  c.ntaps = 9 * kt; c.cin_chunks = cin / 32; c.cin = cin; c.kc = 32;
  c.T_out = geom[0]; c.H_out = geom[1]; c.W_out = geom[2]; c.t_stride = 1;
  c.n_total = geom[3]; c.n_tile = geom[4];
  c.ot_mul = 1; c.ot_add = 0; c.oh_mul = 1; c.oh_add = 0; c.ow_mul = 1; c.ow_add = 0;
  c.Hs = geom[1]; c.Ws = geom[2];
  c.ldc = ldc; c.interleave_half = 0; c.n_store = geom[6];
  c.stages = 0; c.stage_bytes = 0; c.producers = 2; c.tgroup = 1;
  c.bias = bias; c.residual = reinterpret_cast<const bf16*>(residual); c.out = reinterpret_cast<bf16*>(out);
  c.clamp = clamp; c.act_out = reinterpret_cast<bf16*>(act_out); c.act_gamma = act_gamma;
  a.kt = kt; a.t0 = t0; a.rows = geom[5];
  VCOF_REQUIRE(out != nullptr || act_out != nullptr, "vcof_conv_lines: no output requested");
  VCOF_REQUIRE(act_out == nullptr || (act_gamma != nullptr && c.n_tile == c.n_total),
               "vcof_conv_lines: fused norm needs gamma and a single channel pass");
  VCOF_REQUIRE(c.n_total <= kCvVecMax && c.n_total % 16 == 0 && c.n_tile % 16 == 0 && c.n_tile >= 16 && c.n_tile <= 256,
               "vcof_conv_lines: n_total %d / n_tile %d must be multiples of 16, n_tile <= 256", c.n_total, c.n_tile);
  a.naccs = 512 / ((c.n_tile + 31) & ~31);
  if (a.naccs > kLnMaxAcc) a.naccs = kLnMaxAcc;
  VCOF_REQUIRE(a.rows >= 1 && a.rows <= kLnMaxRows && a.rows <= a.naccs,
               "vcof_conv_lines: %d rows of %d channels do not fit 512 TMEM columns", a.rows, c.n_tile);
  VCOF_REQUIRE(ldc % 8 == 0, "vcof_conv_lines: ldc must be a multiple of 8");
  const int w_bytes = 9 * c.n_tile * 64;
  a.ring = (kCvData - 2 * w_bytes) / kLnBytes;
  if (a.ring > kLnMaxRing) a.ring = kLnMaxRing;
  VCOF_REQUIRE(a.ring >= 3, "vcof_conv_lines: %d-channel passes leave no room for the line ring", c.n_tile);
  CUtensorMap tmX, tmW;
  uint64_t dims[5], strides[4];
  for (int i = 0; i < 5; ++i) dims[i] = (uint64_t)x_dims[i];
  for (int i = 0; i < 4; ++i) strides[i] = (uint64_t)x_strides[i] * 2;
  const uint32_t box[5] = {32, (uint32_t)kLnPix, 1, 1, 1};
  int rc = make_tmap_nd_bf16(&tmX, x, 5, dims, strides, box, 64);
  if (rc) return rc;
  // weights: [slices, n_total, 32] with slice = ((chunk * kt + dt) * 3 + dh) * 3 + dw
  uint64_t wd[3] = {32, (uint64_t)c.n_total, (uint64_t)(c.cin_chunks * kt * 9)};
  uint64_t ws[2] = {64, (uint64_t)c.n_total * 64};
  uint32_t wb[3] = {32, (uint32_t)c.n_tile, 9};
  rc = make_tmap_nd_bf16(&tmW, w, 3, wd, ws, wb, 64);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    VCOF_CHECK_CUDA(cudaFuncSetAttribute(conv_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmem));
    attr_set = true;
  }
  const long long items = (long long)c.T_out * ((c.H_out + a.rows - 1) / a.rows) * ((c.W_out + 127) / 128) *
                          ((c.n_total + c.n_tile - 1) / c.n_tile);
  VCOF_REQUIRE(items > 0 && items < (1ll << 31), "vcof_conv_lines: bad item count %lld", items);
  const int grid = items < sm_count() ? (int)items : sm_count();
  conv_lines_kernel<<<grid, kCvThreads, kCvSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tmX, tmW, a);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_rms_silu_cl(const void* x, long long ldx, const float* gamma, void* y, long long ldy,
                                long long npos, int C, int silu, void* stream) {
  VCOF_REQUIRE(C % 2 == 0 && C <= 384 && npos > 0, "vcof_rms_silu_cl: C=%d must be even and <= 384", C);
  const int threads = 256;
  const long long blocks = (npos * 32 + threads - 1) / threads;
  rms_silu_cl_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), gamma, reinterpret_cast<bf16*>(y), npos, C, ldx, ldy, silu);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_nchw_to_cl(const void* x, void* y, int C, int Cp, long long thw, const float* mul,
                               const float* add, void* stream) {
  VCOF_REQUIRE(C > 0 && Cp >= C && thw > 0, "vcof_nchw_to_cl: bad shape");
  const long long n = thw * Cp;
  nchw_to_cl_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), C, Cp, thw, mul, add);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_cl_to_nchw(const void* x, long long ldx, void* y, int C, long long thw, const float* sub,
                               const float* mul, void* stream) {
  VCOF_REQUIRE(C > 0 && thw > 0, "vcof_cl_to_nchw: bad shape");
  const long long n = thw * C;
  cl_to_nchw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), C, ldx, thw, sub, mul);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_softmax_rows(const float* s, long long lds, void* p, long long ldp, int rows, int n,
                                 float scale, void* stream) {
  VCOF_REQUIRE(rows > 0 && n > 0, "vcof_softmax_rows: empty problem");
  softmax_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s, reinterpret_cast<bf16*>(p), n, lds, ldp, scale);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
