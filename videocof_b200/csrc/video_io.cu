// video_io.cu — the byte conversions at the two ends of the pipeline, on the device:
//   frames out: decoder output (channels-last bf16 in [-1, 1]) -> uint8 [T, H, W, 3], the composition of
//               WanPipeline.decode_latents' (frames / 2 + 0.5).clamp(0, 1) in bf16 (pipeline_wan.py:425-426), the
//               .cpu().float() that follows (:427) and save_videos_grid's (x * 255).numpy().astype(np.uint8)
//               (videox_fun/utils/utils.py:59-68);
//   frames in:  uint8 [T, H, W, 3] -> channels-last bf16 in [-1, 1], the composition of load_video_frames'
//               input_video * (2.0 / 255.0) - 1.0 in fp32 (fast_infer.py:86-88) and the pipeline's cast to the VAE
//               dtype (pipeline_wan.py:397).
// A quarter of the fp32 bytes cross PCIe and no conversion pass runs on the host.  Byte work: bit-exact against the
// reference chain.  The per-element arithmetic is __host__ __device__ and also exported as host-side debug entries,
// so the CPU test suite pins it exhaustively (all 65 536 bf16 inputs, all 256 bytes) without a GPU.
#include "vcof_common.cuh"
#include "../../include/vcof.h"

namespace vcof {

__host__ __device__ inline float bf16_bits_to_float(unsigned short b) {
  unsigned int u = (unsigned int)b << 16;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

__host__ __device__ inline unsigned short float_to_bf16_bits(float f) {   // round to nearest even, as torch / cvt.rn
  unsigned int u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);   // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}

// uint8 of one decoder output value (bf16 bits): trunc(255 * clamp(bf16(bf16(x / 2) + 0.5), 0, 1))
__host__ __device__ inline unsigned char frame_u8(unsigned short xb) {
  const float t1 = bf16_bits_to_float(float_to_bf16_bits(bf16_bits_to_float(xb) * 0.5f));
  float t2 = bf16_bits_to_float(float_to_bf16_bits(t1 + 0.5f));
  if (!(t2 == t2)) return 0;                     // NaN never leaves a clamped decoder; pin it to black
  t2 = t2 < 0.f ? 0.f : (t2 > 1.f ? 1.f : t2);
  return (unsigned char)(t2 * 255.0f);           // fp32 product, truncation toward zero (numpy astype)
}

// bf16 bits of one input byte: bf16(fp32(u) * fp32(2/255) - 1), two separately rounded fp32 operations
__host__ __device__ inline unsigned short video_bf16(unsigned char u) {
#ifdef __CUDA_ARCH__
  const float v = __fsub_rn(__fmul_rn((float)u, 0.007843137718737125f), 1.0f);
#else
  volatile float m = (float)u * 0.007843137718737125f;   // volatile: no host-side contraction into an FMA
  const float v = m - 1.0f;
#endif
  return float_to_bf16_bits(v);
}

// Generic shapes: one element per thread.
__global__ void __launch_bounds__(256)
cl_to_u8_kernel(const unsigned short* __restrict__ x, long long ldx, unsigned char* __restrict__ out, long long npos,
                int C) {
  const long long total = npos * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / C;
    const int c = int(i - pos * C);
    out[i] = frame_u8(x[pos * ldx + c]);
  }
}

// The decoder's own shape (C = 3 of ldx = 8 stored channels): a thread owns four consecutive positions, reads their
// four 16-byte rows (64 contiguous bytes) and writes their twelve bytes as three aligned 32-bit words, so a warp reads
// 2 KB and writes 384 B of contiguous memory per iteration.  npos4 = npos / 4; the tail goes through the kernel above.
__global__ void __launch_bounds__(256)
cl8_to_rgb_kernel(const uint4* __restrict__ x, unsigned int* __restrict__ out, long long npos4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npos4; i += (long long)gridDim.x * blockDim.x) {
    unsigned char b[12];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint4 v = __ldg(x + i * 4 + p);                  // channels 0,1 in v.x (little endian), 2,3 in v.y
      b[p * 3 + 0] = frame_u8((unsigned short)(v.x & 0xffffu));
      b[p * 3 + 1] = frame_u8((unsigned short)(v.x >> 16));
      b[p * 3 + 2] = frame_u8((unsigned short)(v.y & 0xffffu));
    }
#pragma unroll
    for (int w = 0; w < 3; ++w)
      out[i * 3 + w] = (unsigned int)b[4 * w] | ((unsigned int)b[4 * w + 1] << 8) | ((unsigned int)b[4 * w + 2] << 16) |
                       ((unsigned int)b[4 * w + 3] << 24);
  }
}

// Generic shapes: one element per thread.
__global__ void __launch_bounds__(256)
u8_to_cl_kernel(const unsigned char* __restrict__ frames, unsigned short* __restrict__ y, long long npos, int C, int Cp) {
  const long long total = npos * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / Cp;
    const int c = int(i - pos * Cp);
    y[i] = c < C ? video_bf16(frames[pos * C + c]) : (unsigned short)0;
  }
}

// The encoder's own shape (C <= 8 real channels in rows of Cp = 8 * groups bf16): a thread writes one 16-byte group of
// one position — the first group of a row carries the converted bytes, the others are the zero padding — so a warp
// writes 512 contiguous bytes per iteration; the 3 input bytes per position are read by one thread in `groups`.
__global__ void __launch_bounds__(256)
u8_to_cl8_kernel(const unsigned char* __restrict__ frames, uint4* __restrict__ y, long long npos, int C, int groups) {
  const long long total = npos * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i / groups;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i - pos * groups == 0) {
      unsigned int h[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) h[c] = c < C ? (unsigned int)video_bf16(__ldg(frames + pos * C + c)) : 0u;
      v = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    }
    y[i] = v;
  }
}

}  // namespace vcof

using namespace vcof;

static unsigned grid_for(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = (long long)sm_count() * 16;       // 16 resident 256-thread CTAs per SM cover the grid-stride loop
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

extern "C" int vcof_cl_to_u8(const void* x, long long ldx, unsigned char* out, long long npos, int C, void* stream) {
  VCOF_REQUIRE(x && out, "vcof_cl_to_u8: null pointer");
  VCOF_REQUIRE(npos > 0 && C > 0 && ldx >= C, "vcof_cl_to_u8: bad shape npos=%lld C=%d ldx=%lld", npos, C, ldx);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  long long done = 0;
  if (C == 3 && ldx == 8 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 &&
      npos >= 4) {
    const long long npos4 = npos / 4;
    cl8_to_rgb_kernel<<<grid_for(npos4), 256, 0, st>>>(reinterpret_cast<const uint4*>(x),
                                                       reinterpret_cast<unsigned int*>(out), npos4);
    VCOF_CHECK_CUDA(cudaGetLastError());
    done = npos4 * 4;
  }
  if (done < npos) {
    cl_to_u8_kernel<<<grid_for((npos - done) * C), 256, 0, st>>>(xs + done * ldx, ldx, out + done * C, npos - done, C);
    VCOF_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int vcof_u8_to_cl(const unsigned char* frames, void* y, long long npos, int C, int Cp, void* stream) {
  VCOF_REQUIRE(frames && y, "vcof_u8_to_cl: null pointer");
  VCOF_REQUIRE(npos > 0 && C > 0 && Cp >= C, "vcof_u8_to_cl: bad shape npos=%lld C=%d Cp=%d", npos, C, Cp);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C <= 8 && Cp % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    u8_to_cl8_kernel<<<grid_for(npos * (Cp / 8)), 256, 0, st>>>(frames, reinterpret_cast<uint4*>(y), npos, C, Cp / 8);
  } else {
    u8_to_cl_kernel<<<grid_for(npos * Cp), 256, 0, st>>>(frames, reinterpret_cast<unsigned short*>(y), npos, C, Cp);
  }
  VCOF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Host-side evaluation of the same per-element functions (HOST pointers): lets the CPU test suite pin the arithmetic
// of the two kernels above exhaustively.  Not a fallback: no product code calls these.
extern "C" int vcof_debug_frame_u8_host(const unsigned short* host_bf16_bits, unsigned char* host_out, long long n) {
  for (long long i = 0; i < n; ++i) host_out[i] = frame_u8(host_bf16_bits[i]);
  return 0;
}

extern "C" int vcof_debug_video_bf16_host(const unsigned char* host_bytes, unsigned short* host_bf16_bits, long long n) {
  for (long long i = 0; i < n; ++i) host_bf16_bits[i] = video_bf16(host_bytes[i]);
  return 0;
}
