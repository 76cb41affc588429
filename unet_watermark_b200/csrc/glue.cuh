// HBM-bound glue kernels of the path: input preparation, stem max-pool, nearest-2x upsample.
// All are pure streaming kernels: 16-byte vector accesses, consecutive threads on consecutive
// addresses, grid-stride loops sized to a multiple of the SM count.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx_sm100.cuh"

namespace uwm {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---------------------------------------------------------------------------------------------
// prep: network input -> bf16 2x2 space-to-depth tensor [n, h/2, w/2, 16]
//   channel (ph*2+pw)*3 + c  = x[c, 2i+ph, 2j+pw];  channels 12..15 = 0.
// The 7x7/s2/p3 stem conv becomes a 4x4/s1 conv over this tensor (taps -2..1).
// u8 input fuses get_val_transform's Normalize: (x/255 - mean)/std, computed in fp32.
// One thread per output pixel (32 bytes out).
// ---------------------------------------------------------------------------------------------
template <bool kU8>
__global__ void __launch_bounds__(256) prep_s2d_kernel(const void* __restrict__ in,
                                                       __nv_bfloat16* __restrict__ out, int n,
                                                       int h, int w) {
  pdl_launch_dependents();
  // u8: get_val_transform's Normalize as ONE fused multiply-add per byte, v = fma(k, 1/(255 std), -mean/std).  For the
  // ImageNet constants this rounds to the same bf16 value as the reference's fp32 operation orders for every one of the
  // 256 x 3 possible inputs - both albumentations' (k - 255 mean) * (1 / (255 std)) and (k/255 - mean)/std (checked
  // exhaustively in tests/test_packing_cpu.py).  (An earlier version tabulated the values in shared memory: ncu
  // showed it bound by the LSU queue - 24 conflicting 2-byte lookups per thread - at 19 us for 46 MB.)
  const float na[3] = {0.017124755f, 0.017507004f, 0.017429195f};       // fp32(fp32(1/255) / std)
  const float nb[3] = {-2.117904f, -2.0357141f, -1.8044444f};           // fp32(-mean / std)
  auto norm = [&](uint32_t byte, int c) { return fmaf(__uint2float_rn(byte), na[c], nb[c]); };
  pdl_wait();
  const int h2 = h >> 1, w2 = w >> 1;
  if (kU8 && (w & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0) {
    // two output pixels per thread: 4 input pixels = 12 bytes = three aligned 32-bit loads per input row
    const int w4 = w >> 2;
    const long long total = (long long)n * h2 * w4;
    const uint8_t* src = static_cast<const uint8_t*>(in);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int jj = (int)(idx % w4);
      const int i = (int)((idx / w4) % h2);
      const int b = (int)(idx / ((long long)w4 * h2));
      float v[2][12];                            // [output pixel][channel (ph*2+pw)*3 + c]
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(src + (((long long)b * h + (2 * i + ph)) * w + 4 * jj) * 3);
        const uint32_t q[3] = {__ldg(r), __ldg(r + 1), __ldg(r + 2)};
#pragma unroll
        for (int e = 0; e < 12; ++e) {           // byte e = pixel e/3 (0..3), channel e%3
          const uint32_t byte = (q[e >> 2] >> (8 * (e & 3))) & 0xffu;
          const int px = e / 3, c = e - 3 * px;
          v[px >> 1][(ph * 2 + (px & 1)) * 3 + c] = norm(byte, c);
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(out + (((long long)b * h2 + i) * w2 + 2 * jj) * 16);
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        uint4 o0, o1;
        o0.x = pack2(v[o][0], v[o][1]);   o0.y = pack2(v[o][2], v[o][3]);
        o0.z = pack2(v[o][4], v[o][5]);   o0.w = pack2(v[o][6], v[o][7]);
        o1.x = pack2(v[o][8], v[o][9]);   o1.y = pack2(v[o][10], v[o][11]);
        o1.z = 0; o1.w = 0;
        dst[2 * o] = o0;
        dst[2 * o + 1] = o1;
      }
    }
    return;
  }
  const long long total = (long long)n * h2 * w2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % w2);
    const int i = (int)((idx / w2) % h2);
    const int b = (int)(idx / ((long long)w2 * h2));
    float v[16];
#pragma unroll
    for (int k = 12; k < 16; ++k) v[k] = 0.f;
    if (kU8) {
      const uint8_t* src = static_cast<const uint8_t*>(in);
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        // 2 pixels x 3 channels = 6 consecutive bytes
        const uint8_t* r = src + (((long long)b * h + (2 * i + ph)) * w + 2 * j) * 3;
#pragma unroll
        for (int pw = 0; pw < 2; ++pw)
#pragma unroll
          for (int c = 0; c < 3; ++c) v[(ph * 2 + pw) * 3 + c] = norm(r[pw * 3 + c], c);
      }
    } else {
      const float* src = static_cast<const float*>(in);
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          const float2 t = *reinterpret_cast<const float2*>(
              src + (((long long)b * 3 + c) * h + (2 * i + ph)) * w + 2 * j);
          v[(ph * 2 + 0) * 3 + c] = t.x;
          v[(ph * 2 + 1) * 3 + c] = t.y;
        }
    }
    uint4 o0, o1;
    o0.x = pack2(v[0], v[1]);   o0.y = pack2(v[2], v[3]);
    o0.z = pack2(v[4], v[5]);   o0.w = pack2(v[6], v[7]);
    o1.x = pack2(v[8], v[9]);   o1.y = pack2(v[10], v[11]);
    o1.z = pack2(v[12], v[13]); o1.w = pack2(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(out + idx * 16);
    dst[0] = o0;
    dst[1] = o1;
  }
}

__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* ra = reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162* rb = reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162* rr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) rr[i] = __hmax2(ra[i], rb[i]);
  return r;
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, padding 1), NHWC bf16.  Padding is -inf (PyTorch semantics):
// out-of-range taps are skipped.  One thread per (output pixel, 8-channel group).
// ---------------------------------------------------------------------------------------------
constexpr int kPoolRows = 8;     // default output rows per thread: input row 2i+1 is shared by outputs i and i+1, kept in registers
template <int ROWS>
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ y, int n,
                                                           int h, int w, int c, long long x_pitch,
                                                           long long y_pitch, int reverse) {
  pdl_launch_dependents();
  pdl_wait();
  const int ho = h >> 1, wo = w >> 1, cg = c >> 3;
  const int strips = (ho + ROWS - 1) / ROWS;
  const long long total = (long long)n * strips * wo * cg;
  const long long n32 = (total + 31) >> 5;
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < n32 * 32;
       it += (long long)gridDim.x * blockDim.x) {
    // reverse: the 32-item groups (one warp each) are walked back to front, lanes keep their order
    const long long idx = reverse ? ((n32 - 1 - (it >> 5)) << 5 | (it & 31)) : it;
    if (idx >= total) continue;
    // a warp = 4 adjacent output columns x 8 channel groups: every load instruction covers whole 128-byte lines, and
    // the column 2j+1 shared with the neighbouring output is an L1 hit
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int j = (int)(t % wo); t /= wo;
    const int i0 = (int)(t % strips) * ROWS;
    const int b = (int)(t / strips);
    const __nv_bfloat16* img = x + (long long)b * h * w * x_pitch + g * 8;
    auto row_max = [&](int ih) {           // max over columns 2j-1 .. 2j+1 of input row ih (ih in range)
      const __nv_bfloat16* r = img + ((long long)ih * w + 2 * j) * x_pitch;
      uint4 m = hmax8(*reinterpret_cast<const uint4*>(r), *reinterpret_cast<const uint4*>(r + x_pitch));   // 2j+1 < w: w even
      if (j > 0) m = hmax8(m, *reinterpret_cast<const uint4*>(r - x_pitch));
      return m;
    };
    uint4 prev = make_uint4(0, 0, 0, 0);
    bool have_prev = i0 > 0;
    if (have_prev) prev = row_max(2 * i0 - 1);
#pragma unroll 2
    for (int i = i0; i < min(i0 + ROWS, ho); ++i) {
      const uint4 a = row_max(2 * i);
      const uint4 bb = row_max(2 * i + 1);                // 2i+1 < h: h even
      uint4 m = hmax8(a, bb);
      if (have_prev) m = hmax8(m, prev);
      prev = bb; have_prev = true;
      *reinterpret_cast<uint4*>(y + (((long long)b * ho + i) * wo + j) * y_pitch + g * 8) = m;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Nearest 2x upsample into channels [0,c) of a pitch-y_pitch buffer:
//   y[b, i, j, 0:c] = x[b, i/2, j/2, 0:c].   One thread per (output pixel, 8-channel group).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x,
                                                         __nv_bfloat16* __restrict__ y, int n,
                                                         int h, int w, int c, long long x_pitch,
                                                         long long y_pitch) {
  pdl_launch_dependents();
  pdl_wait();
  const int ho = h * 2, wo = w * 2, cg = c >> 3;
  const long long total = (long long)n * ho * wo * cg;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int j = (int)(t % wo); t /= wo;
    const int i = (int)(t % ho);
    const int b = (int)(t / ho);
    const uint4 v = *reinterpret_cast<const uint4*>(
        x + (((long long)b * h + (i >> 1)) * w + (j >> 1)) * x_pitch + g * 8);
    *reinterpret_cast<uint4*>(y + (((long long)b * ho + i) * wo + j) * y_pitch + g * 8) = v;
  }
}

}  // namespace uwm
