// elementwise.cu — HBM-bound DiT kernels for sm_100a: LayerNorm+AdaLN modulate,
// full-row RMSNorm + 3-axis RoPE, patchify / unpatchify, fp32 timestep linears.
// Each is a single pass over its operand (one coalesced read, one coalesced write) —
// the reference runs every one of these as 3-6 separate ATen kernels
// (videox_fun/models/wan_transformer3d.py:135-243, 495-511, 870-879, 913-929, 1108-1131).
#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = (l < THREADS / 32) ? scratch[l] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

// ---------------------------------------------------------------------------
// LayerNorm (+affine) + modulate -> bf16
// ---------------------------------------------------------------------------
constexpr int kLnThreads = 256;
constexpr int kLnMaxVec = 8;  // float4 per thread -> C <= 8192

__global__ void __launch_bounds__(kLnThreads)
ln_modulate_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ ln_w,
                   const float* __restrict__ ln_b, const float* __restrict__ shift,
                   const float* __restrict__ scale, bf16* __restrict__ out, long long ldo, int C,
                   float eps) {
  __shared__ float scratch[kLnThreads / 32];
  const long long row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int nvec = C >> 2;
  float4 v[kLnMaxVec];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = threadIdx.x + i * kLnThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = block_sum<kLnThreads>(sum, scratch) / float(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = threadIdx.x + i * kLnThreads;
    if (idx < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum<kLnThreads>(sq, scratch) / float(C) + eps);
  bf16* orow = out + row * ldo;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int idx = threadIdx.x + i * kLnThreads;
    if (idx < nvec) {
      float y[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                    (v[i].w - mean) * rstd};
      if (ln_w != nullptr) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(ln_w) + idx);
        y[0] *= w.x; y[1] *= w.y; y[2] *= w.z; y[3] *= w.w;
      }
      if (ln_b != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ln_b) + idx);
        y[0] += b.x; y[1] += b.y; y[2] += b.z; y[3] += b.w;
      }
      if (scale != nullptr) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale) + idx);
        y[0] *= (1.f + s.x); y[1] *= (1.f + s.y); y[2] *= (1.f + s.z); y[3] *= (1.f + s.w);
      }
      if (shift != nullptr) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(shift) + idx);
        y[0] += s.x; y[1] += s.y; y[2] += s.z; y[3] += s.w;
      }
      uint2 pk;
      pk.x = pack_bf16x2(y[0], y[1]);
      pk.y = pack_bf16x2(y[2], y[3]);
      reinterpret_cast<uint2*>(orow)[idx] = pk;
    }
  }
}

// ---------------------------------------------------------------------------
// full-row RMSNorm (+ RoPE), in place on bf16
// ---------------------------------------------------------------------------
constexpr int kRmsThreads = 128;
constexpr int kRmsMaxVec = 8;  // uint4 (8 bf16) per thread -> C <= 8192

// Destinations of a scattered column-blocked output: block b (cols_per_block columns of every row) is the dense
// [rows, cols_per_block] slab at p[b].  Under sequence parallelism the slabs are the receive buffers of the other
// ranks, mapped through NVLink peer memory: the kernel's own stores are the all-to-all.
constexpr int kMaxScatter = 16;
struct ScatterPtrs {
  bf16* p[kMaxScatter];
};

template <bool SCATTER>
__global__ void __launch_bounds__(kRmsThreads)
rmsnorm_rope_kernel(bf16* __restrict__ x, long long ldx, const bf16* __restrict__ weight, float eps,
                    int C, int head_dim, const float2* __restrict__ table,
                    const int* __restrict__ tpos, int F, int H, int W, int n_t, int n_h,
                    int row_offset, bf16* __restrict__ y, int cols_per_block, long long block_stride,
                    const __grid_constant__ ScatterPtrs dst) {
  __shared__ float scratch[kRmsThreads / 32];
  // cos / sin of this row's rotation, one entry per complex pair of a head: every head of the row uses the same 64
  // angles, so they are fetched from the table ONCE per row (first kRmsThreads lanes, published by the barrier inside
  // block_sum) instead of once per pair per head (20 scattered 8-byte loads per thread: the kernel ran at 0.44 of the
  // HBM rate inside a step, LSU-bound, not DRAM-bound)
  __shared__ __align__(16) float2 s_cs[kRmsThreads];
  const long long row = blockIdx.x;
  uint4* xr = reinterpret_cast<uint4*>(x + row * ldx);
  const int nvec = C >> 3;
  const long long g = (long long)row_offset + row;
  const bool do_rope = (table != nullptr) && (g < (long long)F * H * W);
  if (do_rope && threadIdx.x < (head_dim >> 1)) {
    const int f = int(g / ((long long)H * W));
    const int r = int(g % ((long long)H * W));
    const int pi = threadIdx.x;
    const int pos = (pi < n_t) ? tpos[f] : ((pi < n_t + n_h) ? r / W : r % W);
    s_cs[pi] = __ldg(table + pos * (head_dim >> 1) + pi);
  }
  uint4 v[kRmsMaxVec];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kRmsMaxVec; ++i) {
    const int idx = threadIdx.x + i * kRmsThreads;
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
  const float ms = block_sum<kRmsThreads>(sq, scratch) / float(C);
  // the reference rounds the rsqrt factor to bf16 before the multiply (:229)
  const float rs = bf16_round(rsqrtf(ms + eps));

#pragma unroll
  for (int i = 0; i < kRmsMaxVec; ++i) {
    const int idx = threadIdx.x + i * kRmsThreads;
    if (idx < nvec) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight) + idx);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
      uint32_t o[4];
      // first complex pair of this vector (a multiple of 4); head dims are powers of two in practice: a mask instead
      // of an integer modulo by a run-time value (a MUFU.RCP sequence per vector: the XU pipe was 53 % busy)
      const int c0 = idx * 8;
      const int pair0 = (((head_dim & (head_dim - 1)) == 0) ? (c0 & (head_dim - 1)) : (c0 % head_dim)) >> 1;
      float2 csv[4];
      if (do_rope) {
        const float4 c01 = *reinterpret_cast<const float4*>(&s_cs[pair0]);
        const float4 c23 = *reinterpret_cast<const float4*>(&s_cs[pair0 + 2]);
        csv[0] = make_float2(c01.x, c01.y); csv[1] = make_float2(c01.z, c01.w);
        csv[2] = make_float2(c23.x, c23.y); csv[3] = make_float2(c23.z, c23.w);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        const float2 w = __bfloat1622float2(wh[e]);
        float yr = bf16_round(bf16_round(f.x * rs) * w.x);
        float yi = bf16_round(bf16_round(f.y * rs) * w.y);
        if (do_rope) {
          const float2 cs = csv[e];
          const float r0 = yr * cs.x - yi * cs.y;
          const float r1 = yr * cs.y + yi * cs.x;
          yr = r0;
          yi = r1;
        }
        o[e] = pack_bf16x2(yr, yi);
      }
      const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
      if constexpr (SCATTER) {
        const int c = idx * 8;
        const int blk = c / cols_per_block;
        *reinterpret_cast<uint4*>(dst.p[blk] + row * cols_per_block + (c - blk * cols_per_block)) = ov;
      } else if (y == nullptr) {
        xr[idx] = ov;
      } else {
        // column-blocked output [C / cols_per_block][rows][cols_per_block]: the send layout of the head exchange
        const int c = idx * 8;
        const int blk = c / cols_per_block;
        *reinterpret_cast<uint4*>(y + blk * block_stride + row * cols_per_block + (c - blk * cols_per_block)) = ov;
      }
    }
  }
}

// Copy between a row-major [rows, C] matrix (pitch ld) and its column-blocked form
// [C / cols_per_block][rows][cols_per_block] (blocks block_stride elements apart); 16 bytes per thread.
__global__ void __launch_bounds__(256)
copy_blocked_kernel(bf16* __restrict__ rowmajor, long long ld, bf16* __restrict__ blocked, long long block_stride,
                    long long rows, int C, int cols_per_block, int to_blocked) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nvec;
    const int c = int(i - row * nvec) * 8;
    const int blk = c / cols_per_block;
    uint4* a = reinterpret_cast<uint4*>(rowmajor + row * ld + c);
    uint4* b = reinterpret_cast<uint4*>(blocked + blk * block_stride + row * cols_per_block + (c - blk * cols_per_block));
    if (to_blocked) *b = *a; else *a = *b;
  }
}

// Pack of a row-major [rows, C] matrix into per-block destination slabs (see ScatterPtrs): the V leg of the
// push-style head exchange.
__global__ void __launch_bounds__(256)
copy_scatter_kernel(const bf16* __restrict__ rowmajor, long long ld, const __grid_constant__ ScatterPtrs dst, long long rows, int C,
                    int cols_per_block) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nvec;
    const int c = int(i - row * nvec) * 8;
    const int blk = c / cols_per_block;
    *reinterpret_cast<uint4*>(dst.p[blk] + row * cols_per_block + (c - blk * cols_per_block)) =
        *reinterpret_cast<const uint4*>(rowmajor + row * ld + c);
  }
}

// Row chunks of a dense [n_chunks * rows, cols] matrix to per-chunk destination slabs [rows, cols]: the return leg
// of the push-style head exchange (chunk r = the query rows that rank r owns).
__global__ void __launch_bounds__(256)
copy_rows_scatter_kernel(const bf16* __restrict__ src, long long ld, const __grid_constant__ ScatterPtrs dst, long long rows, int cols,
                         int n_chunks) {
  const int nvec = cols >> 3;
  const long long total = rows * n_chunks * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long grow = i / nvec;                 // row of src
    const int c = int(i - grow * nvec) * 8;
    const int chunk = int(grow / rows);
    const long long r = grow - (long long)chunk * rows;
    *reinterpret_cast<uint4*>(dst.p[chunk] + r * cols + c) = *reinterpret_cast<const uint4*>(src + grow * ld + c);
  }
}

// ---------------------------------------------------------------------------
// patchify / unpatchify
// ---------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ x, bf16* __restrict__ a, int Cin, int F,
                                int H, int W) {
  const int H2 = H >> 1, W2 = W >> 1;
  const long long total = (long long)F * H2 * W2 * Cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = int(i % Cin);
  const long long l = i / Cin;
  const int w2 = int(l % W2);
  const int h2 = int((l / W2) % H2);
  const int f = int(l / ((long long)W2 * H2));
  const bf16* src = x + (((long long)c * F + f) * H + 2 * h2) * W + 2 * w2;
  const uint32_t top = *reinterpret_cast<const uint32_t*>(src);
  const uint32_t bot = *reinterpret_cast<const uint32_t*>(src + W);
  *reinterpret_cast<uint2*>(a + l * (Cin * 4) + c * 4) = make_uint2(top, bot);
}

__global__ void unpatchify_kernel(const bf16* __restrict__ y, long long ldy, bf16* __restrict__ out,
                                  int Cout, int F, int H, int W) {
  const long long total = (long long)Cout * F * H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ww = int(i % W);
  const int hh = int((i / W) % H);
  const int f = int((i / ((long long)W * H)) % F);
  const int c = int(i / ((long long)W * H * F));
  const int H2 = H >> 1, W2 = W >> 1;
  const long long l = ((long long)f * H2 + (hh >> 1)) * W2 + (ww >> 1);
  const int col = ((hh & 1) * 2 + (ww & 1)) * Cout + c;
  out[i] = y[l * ldy + col];
}

// ---------------------------------------------------------------------------
// fp32 linear with bf16 weights, one warp per output element
// ---------------------------------------------------------------------------
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

__global__ void linear_f32_kernel(const float* __restrict__ x, const bf16* __restrict__ w,
                                  const bf16* __restrict__ bias, float* __restrict__ out, int B,
                                  int N, int K, int act_in, int act_out) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)B * N) return;
  const int n = int(gw % N);
  const int b = int(gw / N);
  const uint4* wr = reinterpret_cast<const uint4*>(w + (long long)n * K);
  const float* xr = x + (long long)b * K;
  float acc = 0.f;
  for (int kv = lane; kv < (K >> 3); kv += 32) {
    const uint4 wv = __ldg(wr + kv);
    const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 wf = __bfloat1622float2(wh[e]);
      float x0 = xr[kv * 8 + e * 2], x1 = xr[kv * 8 + e * 2 + 1];
      if (act_in == 1) { x0 = silu(x0); x1 = silu(x1); }
      acc = fmaf(x0, wf.x, acc);
      acc = fmaf(x1, wf.y, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (bias != nullptr) acc += __bfloat162float(bias[n]);
    if (act_out == 1) acc = silu(acc);
    out[(long long)b * N + n] = acc;
  }
}

}  // namespace vcof

using namespace vcof;

extern "C" int vcof_ln_modulate(const float* x, long long ldx, const float* ln_w, const float* ln_b,
                                const float* shift, const float* scale, void* out, long long ldo,
                                int L, int C, float eps, void* stream) {
  VCOF_REQUIRE(L > 0 && C > 0, "vcof_ln_modulate: empty problem");
  VCOF_REQUIRE(C % 4 == 0 && C <= kLnThreads * kLnMaxVec * 4,
               "vcof_ln_modulate: C=%d must be a multiple of 4 and <= %d", C,
               kLnThreads * kLnMaxVec * 4);
  VCOF_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "vcof_ln_modulate: ldx/ldo must be multiples of 4");
  ln_modulate_kernel<<<L, kLnThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, ldx, ln_w, ln_b, shift, scale, reinterpret_cast<bf16*>(out), ldo, C, eps);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int rmsnorm_rope_launch(void* x, long long ldx, const void* weight, float eps, int L, int C,
                               int head_dim, const float* rope_table, const int* tpos, int F,
                               int H, int W, int n_t, int n_h, int row_offset, void* y, int cols_per_block,
                               long long block_stride, void* stream) {
  VCOF_REQUIRE(L > 0 && C > 0, "vcof_rmsnorm_rope: empty problem");
  VCOF_REQUIRE(y == nullptr || (cols_per_block > 0 && cols_per_block % 8 == 0 && C % cols_per_block == 0 &&
                                block_stride >= (long long)L * cols_per_block && block_stride % 8 == 0),
               "vcof_rmsnorm_rope_blocked: cols_per_block %d must divide C=%d in multiples of 8 and blocks must "
               "not overlap", cols_per_block, C);
  VCOF_REQUIRE(C % 8 == 0 && C <= kRmsThreads * kRmsMaxVec * 8 && ldx % 8 == 0,
               "vcof_rmsnorm_rope: C=%d / ldx must be multiples of 8, C <= %d", C,
               kRmsThreads * kRmsMaxVec * 8);
  VCOF_REQUIRE(head_dim % 8 == 0 && C % head_dim == 0 && head_dim <= 2 * kRmsThreads, "vcof_rmsnorm_rope: bad head_dim %d",
               head_dim);
  VCOF_REQUIRE(rope_table == nullptr || (tpos != nullptr && n_t + n_h <= head_dim / 2),
               "vcof_rmsnorm_rope: rope requested without positions");
  rmsnorm_rope_kernel<false><<<L, kRmsThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<bf16*>(x), ldx, reinterpret_cast<const bf16*>(weight), eps, C, head_dim,
      reinterpret_cast<const float2*>(rope_table), tpos, F, H, W, n_t, n_h, row_offset,
      reinterpret_cast<bf16*>(y), cols_per_block, block_stride, ScatterPtrs{});
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_rmsnorm_rope(void* x, long long ldx, const void* weight, float eps, int L, int C,
                                 int head_dim, const float* rope_table, const int* tpos, int F,
                                 int H, int W, int n_t, int n_h, int row_offset, void* stream) {
  return rmsnorm_rope_launch(x, ldx, weight, eps, L, C, head_dim, rope_table, tpos, F, H, W, n_t, n_h, row_offset,
                             nullptr, 0, 0, stream);
}

extern "C" int vcof_rmsnorm_rope_blocked(const void* x, long long ldx, void* y, int cols_per_block,
                                         long long block_stride, const void* weight, float eps, int L, int C,
                                         int head_dim, const float* rope_table, const int* tpos, int F, int H,
                                         int W, int n_t, int n_h, int row_offset, void* stream) {
  VCOF_REQUIRE(y != nullptr, "vcof_rmsnorm_rope_blocked: no output");
  return rmsnorm_rope_launch(const_cast<void*>(x), ldx, weight, eps, L, C, head_dim, rope_table, tpos, F, H, W, n_t,
                             n_h, row_offset, y, cols_per_block, block_stride, stream);
}

static int fill_scatter(ScatterPtrs* dst, void* const* ptrs, int n, const char* who) {
  VCOF_REQUIRE(ptrs != nullptr && n >= 1 && n <= kMaxScatter, "%s: need 1..%d destination pointers, got %d", who,
               kMaxScatter, n);
  for (int i = 0; i < n; ++i) {
    VCOF_REQUIRE(ptrs[i] != nullptr && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0,
                 "%s: destination %d is null or not 16-byte aligned", who, i);
    dst->p[i] = reinterpret_cast<bf16*>(ptrs[i]);
  }
  for (int i = n; i < kMaxScatter; ++i) dst->p[i] = nullptr;
  return 0;
}

extern "C" int vcof_rmsnorm_rope_scatter(const void* x, long long ldx, void* const* block_ptrs, int n_blocks,
                                         const void* weight, float eps, int L, int C, int head_dim,
                                         const float* rope_table, const int* tpos, int F, int H, int W, int n_t,
                                         int n_h, int row_offset, void* stream) {
  ScatterPtrs dst;
  if (int rc = fill_scatter(&dst, block_ptrs, n_blocks, "vcof_rmsnorm_rope_scatter")) return rc;
  VCOF_REQUIRE(L > 0 && C > 0 && C % n_blocks == 0 && (C / n_blocks) % 8 == 0,
               "vcof_rmsnorm_rope_scatter: C=%d must split into %d blocks of a multiple of 8 columns", C, n_blocks);
  VCOF_REQUIRE(C % 8 == 0 && C <= kRmsThreads * kRmsMaxVec * 8 && ldx % 8 == 0,
               "vcof_rmsnorm_rope_scatter: C=%d / ldx must be multiples of 8, C <= %d", C, kRmsThreads * kRmsMaxVec * 8);
  VCOF_REQUIRE(head_dim % 8 == 0 && C % head_dim == 0 && head_dim <= 2 * kRmsThreads,
               "vcof_rmsnorm_rope_scatter: bad head_dim %d", head_dim);
  VCOF_REQUIRE(rope_table == nullptr || (tpos != nullptr && n_t + n_h <= head_dim / 2),
               "vcof_rmsnorm_rope_scatter: rope requested without positions");
  rmsnorm_rope_kernel<true><<<L, kRmsThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      const_cast<bf16*>(reinterpret_cast<const bf16*>(x)), ldx, reinterpret_cast<const bf16*>(weight), eps, C,
      head_dim, reinterpret_cast<const float2*>(rope_table), tpos, F, H, W, n_t, n_h, row_offset, nullptr,
      C / n_blocks, 0, dst);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_copy_scatter(const void* rowmajor, long long ld, void* const* block_ptrs, int n_blocks,
                                 long long rows, int C, void* stream) {
  ScatterPtrs dst;
  if (int rc = fill_scatter(&dst, block_ptrs, n_blocks, "vcof_copy_scatter")) return rc;
  VCOF_REQUIRE(rows > 0 && C > 0 && ld % 8 == 0 && ld >= C && C % n_blocks == 0 && (C / n_blocks) % 8 == 0,
               "vcof_copy_scatter: bad shape rows=%lld C=%d ld=%lld blocks=%d", rows, C, ld, n_blocks);
  const long long total = rows * (C >> 3);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  copy_scatter_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(rowmajor), ld, dst, rows, C, C / n_blocks);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_copy_rows_scatter(const void* src, long long ld, void* const* chunk_ptrs, int n_chunks,
                                      long long rows, int cols, void* stream) {
  ScatterPtrs dst;
  if (int rc = fill_scatter(&dst, chunk_ptrs, n_chunks, "vcof_copy_rows_scatter")) return rc;
  VCOF_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && ld >= cols,
               "vcof_copy_rows_scatter: bad shape rows=%lld cols=%d ld=%lld", rows, cols, ld);
  const long long total = rows * n_chunks * (cols >> 3);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  copy_rows_scatter_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(src), ld, dst, rows, cols, n_chunks);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_copy_blocked(void* rowmajor, long long ld, void* blocked, long long block_stride, long long rows,
                                 int C, int cols_per_block, int to_blocked, void* stream) {
  VCOF_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && ld % 8 == 0 && ld >= C, "vcof_copy_blocked: bad shape rows=%lld C=%d",
               rows, C);
  VCOF_REQUIRE(cols_per_block > 0 && cols_per_block % 8 == 0 && C % cols_per_block == 0 && block_stride % 8 == 0 &&
               block_stride >= rows * cols_per_block,
               "vcof_copy_blocked: cols_per_block %d must divide C=%d in multiples of 8", cols_per_block, C);
  const long long total = rows * (C >> 3);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  copy_blocked_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<bf16*>(rowmajor), ld, reinterpret_cast<bf16*>(blocked), block_stride, rows, C, cols_per_block,
      to_blocked);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_patchify(const void* x, void* a, int Cin, int F, int H, int W, void* stream) {
  VCOF_REQUIRE(Cin > 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
               "vcof_patchify: latent H=%d W=%d must be even", H, W);
  const long long total = (long long)F * (H / 2) * (W / 2) * Cin;
  const int threads = 256;
  patchify_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0,
                    reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(a), Cin, F, H, W);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_unpatchify(const void* y, long long ldy, void* out, int Cout, int F, int H,
                               int W, void* stream) {
  VCOF_REQUIRE(Cout > 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
               "vcof_unpatchify: latent H=%d W=%d must be even", H, W);
  const long long total = (long long)Cout * F * H * W;
  const int threads = 256;
  unpatchify_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0,
                      reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(y), ldy, reinterpret_cast<bf16*>(out), Cout, F, H, W);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int vcof_linear_f32(const float* x, const void* w, const void* bias, float* out, int B,
                               int N, int K, int act_in, int act_out, void* stream) {
  VCOF_REQUIRE(B > 0 && N > 0 && K > 0 && K % 8 == 0, "vcof_linear_f32: bad shape B=%d N=%d K=%d", B,
               N, K);
  const long long warps = (long long)B * N;
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  linear_f32_kernel<<<(unsigned)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const bf16*>(w), reinterpret_cast<const bf16*>(bias), out, B, N, K, act_in,
      act_out);
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
