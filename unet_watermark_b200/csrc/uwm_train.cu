// libuwm_b200.so, third translation unit: the memory-bound glue of the optional training step (BASELINE.json configs[4];
// reference src/train.py:68-127 runs the smp Unet under model.train(), so every Conv2dReLU is conv -> BatchNorm2d with
// BATCH statistics -> ReLU and every residual block ends in BatchNorm -> add -> ReLU).
//
//   bn_stats_kernel / bn_fwd_finalize_kernel / bn_apply_kernel        train-mode BatchNorm2d (+ residual)(+ ReLU), forward
//   bn_bwd_reduce_kernel / bn_bwd_finalize_kernel / bn_bwd_apply_kernel   its backward (ReLU mask, d residual, dx, dgamma, dbeta)
//   upsample2x_bwd_kernel                                              backward of F.interpolate(x2, nearest): 2x2 sum
//   pack_train_weights_kernel                                          fp32 filter -> bf16 operands of the forward and the dgrad conv
//   copy_channels_kernel                                               channel slices in / out of the decoder's concat buffers
//   maxpool3x3s2_bwd_kernel                                            backward of the stem's MaxPool2d(3, 2, 1), arg-max recomputed
//
// All of it is HBM-bound: NHWC bf16, 16-byte accesses (8 channels per thread), fp32 arithmetic, per-channel sums
// accumulated per thread in fp32, per block in shared memory and across blocks with fp64 atomics.  Algorithmic bytes per
// pixel and channel: forward 2 (stats read) + 2 + 2 (apply read / write) (+2 residual); backward 4 (reduce: dy, x) + 6
// (apply: dy, x, dx) (+2 y for the ReLU mask of a residual block, +2 d residual).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/uwm.h"

extern "C" void uwm_internal_set_error(const char* msg);        // uwm_api.cu (thread-local message)
extern "C" void uwm_internal_count_launches(int n);

namespace {

int tfail(int code, const char* fmt, ...) {
  char b[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof(b), fmt, ap);
  va_end(ap);
  uwm_internal_set_error(b);
  return code;
}
int tpost(const char* what) {
  uwm_internal_count_launches(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tfail(UWM_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return UWM_OK;
}

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ uint4 ld16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Thread layout shared by every kernel below: a block is rows_per_block x groups threads, groups = c / 8; thread
// (r, g) owns channels [8g, 8g+8) of rows blockIdx.x * rpb + r, + gridDim.x * rpb, ...  A warp reads 512 contiguous
// bytes when groups <= 32; the per-channel constants live in registers for the whole loop.
struct Geom {
  int groups, rpb, threads, blocks;
};
Geom geom(long long pixels, int c, int sms) {
  Geom g;
  g.groups = c / 8;
  g.rpb = std::max(1, 256 / g.groups);
  g.threads = g.rpb * g.groups;                       // <= 256 for c <= 2048
  const long long want = (pixels + (long long)g.rpb * 4 - 1) / ((long long)g.rpb * 4);   // >= 4 rows per thread
  g.blocks = (int)std::max(1LL, std::min(want, (long long)sms * 8));
  return g;
}
int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// Fold of the 2 x 8 per-thread partial sums.  Threads of a warp that own the same channel group (lanes g, g + groups, ...:
// groups divides 32) are combined with xor shuffles first, so shared memory sees one add per warp and channel instead of
// one per thread (fp32 shared atomics are compare-and-swap loops: 128 threads on one address serialise); across blocks
// the sums go to one of kSlots copies of ws (blockIdx % kSlots) - 1184 blocks adding fp64 atomics into the SAME eight
// cache lines took longer than reading the tensor.
constexpr int kSlots = UWM_BN_WS_SLOTS;

__device__ __forceinline__ void fold_sums(float* a, float* q, int g, int c, int groups, float* sm, double* ws) {
  const bool shuffled = groups <= 32 && (groups & (groups - 1)) == 0;      // then blockDim = 256: whole warps
  if (shuffled) {
    for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], off);
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], off);
      }
    }
  }
  if (!shuffled || (threadIdx.x & 31) < groups) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(&sm[g * 8 + k], a[k]);
      atomicAdd(&sm[c + g * 8 + k], q[k]);
    }
  }
  __syncthreads();
  double* slot = ws + (size_t)(blockIdx.x % kSlots) * 2 * c;
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) atomicAdd(&slot[i], (double)sm[i]);
}

// sum of the kSlots copies of entries i and c + i, which are left zeroed for the next call: one WARP per channel, lane k
// owns slot k, so the 2 x 32 loads of a channel are one L2 round trip instead of eight dependent ones (a thread per
// channel walking its 64 values took 9-12 us per launch, 92 launches per training step)
static_assert(kSlots == 32, "take_slots: one lane per slot");
__device__ __forceinline__ void take_slots(double* ws, int i, int c, double& s0, double& s1) {
  double* p = ws + (size_t)(threadIdx.x & 31) * 2 * c + i;
  double a = p[0], b = p[c];
  p[0] = 0.0;
  p[c] = 0.0;
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
  }
  s0 = a;
  s1 = b;
}

// ---- forward ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ x, long long pixels, int c, int groups,
                                                       int rpb, double* __restrict__ ws) {
  extern __shared__ float sm[];                       // [2][c]
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int r = threadIdx.x / groups, g = threadIdx.x - r * groups;
  float a[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = q[k] = 0.f;
  const long long step = (long long)gridDim.x * rpb;
#pragma unroll 4
  for (long long row = (long long)blockIdx.x * rpb + r; row < pixels; row += step) {
    float f[8];
    unpack8(ld16(x + row * c + g * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a[k] += f[k];
      q[k] = fmaf(f[k], f[k], q[k]);
    }
  }
  fold_sums(a, q, g, c, groups, sm, ws);
}

// mean / biased variance -> save_mean, save_rstd, scale = gamma * rstd, shift = beta - mean * scale; running statistics
// updated like nn.BatchNorm2d (momentum, UNBIASED variance); the fp64 sums are left zeroed for the next call.
__global__ void bn_fwd_finalize_kernel(double* __restrict__ ws, long long pixels, int c, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                       float momentum, float eps, float* __restrict__ save_mean,
                                       float* __restrict__ save_rstd, float* __restrict__ scale, float* __restrict__ shift) {
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < c; i += (gridDim.x * blockDim.x) >> 5) {   // warp-uniform
    double s, ss;
    take_slots(ws, i, c, s, ss);
    if (threadIdx.x & 31) continue;
    const double mean = s / (double)pixels;
    double var = ss / (double)pixels - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[i] * rstd;
    save_mean[i] = (float)mean;
    save_rstd[i] = rstd;
    scale[i] = sc;
    shift[i] = fmaf(-(float)mean, sc, beta[i]);
    if (rmean) {
      const double unbiased = pixels > 1 ? var * (double)pixels / (double)(pixels - 1) : var;
      rmean[i] = (1.f - momentum) * rmean[i] + momentum * (float)mean;
      rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unbiased;
    }
  }
}

template <bool RES>
__global__ void __launch_bounds__(256) bn_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                       bf16* __restrict__ y, long long pixels, int c, int groups, int rpb,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       int relu) {
  const int r = threadIdx.x / groups, g = threadIdx.x - r * groups;
  float sc[8], sh[8];
  load8f(scale + g * 8, sc);
  load8f(shift + g * 8, sh);
  const long long step = (long long)gridDim.x * rpb;
#pragma unroll 4
  for (long long row = (long long)blockIdx.x * rpb + r; row < pixels; row += step) {
    const long long off = row * c + g * 8;
    float f[8], e[8];
    unpack8(ld16(x + off), f);
    if (RES) unpack8(ld16(res + off), e);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = fmaf(f[k], sc[k], sh[k]);
      if (RES) v += e[k];
      f[k] = relu ? fmaxf(v, 0.f) : v;
    }
    *reinterpret_cast<uint4*>(y + off) = pack8(f);
  }
}

// ---- backward --------------------------------------------------------------------------------------------------
// g = dy * [y > 0] (ReLU; the mask is recomputed as scale * x + shift > 0 without a residual, read from y with one),
// sums over pixels of g and g * (x - mean).
template <bool RES>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const bf16* __restrict__ y, long long pixels, int c, int groups,
                                                            int rpb, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const float* __restrict__ mean,
                                                            int relu, double* __restrict__ ws) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int r = threadIdx.x / groups, g = threadIdx.x - r * groups;
  float sc[8], sh[8], mu[8], a[8], q[8];
  load8f(scale + g * 8, sc);
  load8f(shift + g * 8, sh);
  load8f(mean + g * 8, mu);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = q[k] = 0.f;
  const long long step = (long long)gridDim.x * rpb;
#pragma unroll 2
  for (long long row = (long long)blockIdx.x * rpb + r; row < pixels; row += step) {
    const long long off = row * c + g * 8;
    float d[8], f[8], o[8];
    unpack8(ld16(dy + off), d);
    unpack8(ld16(x + off), f);
    if (RES) unpack8(ld16(y + off), o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool on = !relu || (RES ? o[k] > 0.f : fmaf(f[k], sc[k], sh[k]) > 0.f);
      const float gk = on ? d[k] : 0.f;
      a[k] += gk;
      q[k] = fmaf(gk, f[k] - mu[k], q[k]);
    }
  }
  fold_sums(a, q, g, c, groups, sm, ws);
}

// dbeta = sum g; dgamma = rstd * sum g (x - mean); coefficients of dx = scale * (g - kb - (x - mean) * kc) with
// kb = sum g / P, kc = rstd^2 * sum g (x - mean) / P.  Leaves the fp64 sums zeroed.
__global__ void bn_bwd_finalize_kernel(double* __restrict__ ws, long long pixels, int c, const float* __restrict__ rstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ kb,
                                       float* __restrict__ kc) {
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < c; i += (gridDim.x * blockDim.x) >> 5) {   // warp-uniform
    double sg, sgx;
    take_slots(ws, i, c, sg, sgx);
    if (threadIdx.x & 31) continue;
    const double rs = (double)rstd[i];
    dbeta[i] = (float)sg;
    dgamma[i] = (float)(rs * sgx);
    kb[i] = (float)(sg / (double)pixels);
    kc[i] = (float)(rs * rs * sgx / (double)pixels);
  }
}

template <bool RES>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                           const bf16* __restrict__ y, bf16* __restrict__ dx,
                                                           bf16* __restrict__ dres, long long pixels, int c, int groups,
                                                           int rpb, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ mean,
                                                           const float* __restrict__ kb, const float* __restrict__ kc,
                                                           int relu) {
  const int r = threadIdx.x / groups, g = threadIdx.x - r * groups;
  float sc[8], sh[8], mu[8], b[8], cc[8];
  load8f(scale + g * 8, sc);
  load8f(shift + g * 8, sh);
  load8f(mean + g * 8, mu);
  load8f(kb + g * 8, b);
  load8f(kc + g * 8, cc);
  const long long step = (long long)gridDim.x * rpb;
#pragma unroll 2
  for (long long row = (long long)blockIdx.x * rpb + r; row < pixels; row += step) {
    const long long off = row * c + g * 8;
    float d[8], f[8], o[8];
    unpack8(ld16(dy + off), d);
    unpack8(ld16(x + off), f);
    if (RES) unpack8(ld16(y + off), o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool on = !relu || (RES ? o[k] > 0.f : fmaf(f[k], sc[k], sh[k]) > 0.f);
      const float gk = on ? d[k] : 0.f;
      d[k] = gk;
      f[k] = sc[k] * (gk - b[k] - (f[k] - mu[k]) * cc[k]);
    }
    *reinterpret_cast<uint4*>(dx + off) = pack8(f);
    if (RES) *reinterpret_cast<uint4*>(dres + off) = pack8(d);
  }
}

// backward of nearest 2x: dx[n,h,w,:] = sum of the 2x2 block of dy (channels [0,c) of rows with pixel pitch dy_pitch)
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const bf16* __restrict__ dy, long long dy_pitch, int n, int h,
                                                             int w, int c, bf16* __restrict__ dx, long long dx_pitch) {
  const int groups = c / 8;
  const long long items = (long long)n * h * w * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    long long px = idx / groups;
    const int xw = (int)(px % w);
    const long long t = px / w;
    const int yh = (int)(t % h);
    const long long img = t / h;
    const long long base = ((img * 2 * h + 2 * yh) * (2LL * w) + 2 * xw) * dy_pitch + g * 8;
    float s[8], f[8];
    unpack8(ld16(dy + base), s);
    unpack8(ld16(dy + base + dy_pitch), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
    unpack8(ld16(dy + base + 2LL * w * dy_pitch), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
    unpack8(ld16(dy + base + (2LL * w + 1) * dy_pitch), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
    *reinterpret_cast<uint4*>(dx + px * dx_pitch + g * 8) = pack8(s);
  }
}

// backward of MaxPool2d(3, stride 2, padding 1), no saved indices.  A thread owns a 2x2 block of input pixels (rows 2k,
// 2k+1, columns 2j, 2j+1; 8 channels): the four windows that can route a gradient into it - (k,j), (k,j+1), (k+1,j),
// (k+1,j+1) - lie inside the 5x5 neighbourhood around it, which is read once (25 16-byte loads for four pixels; from L1 /
// L2 mostly - the tensor crosses HBM once) while the arg-max of each window is tracked with torch's rule: the FIRST
// maximum in row-major scan order (strict '>' while scanning).  A pixel then sums the gradients of the windows whose
// arg-max it is, in torch's order, in fp32.  (torch's max_pool_backward_nhwc: 0.54 ms of the step for this 134 MB tensor.)
__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, int n,
                                                               int h, int w, int c, bf16* __restrict__ dx) {
  const int groups = c / 8, oh_n = (h - 1) / 2 + 1, ow_n = (w - 1) / 2 + 1, bh = (h + 1) / 2, bw = (w + 1) / 2;
  const long long items = (long long)n * bh * bw * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    long long t = idx / groups;
    const int j = (int)(t % bw);
    t /= bw;
    const int k = (int)(t % bh);
    const long long img = t / bh;
    const bf16* xi = x + img * h * w * c + g * 8;
    const bf16* dyi = dy + img * oh_n * ow_n * c + g * 8;
    float mx[4][8];
    int am[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        mx[q][e] = -INFINITY;
        am[q][e] = -1;
      }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const int yy = 2 * k - 1 + a;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const int xx = 2 * j - 1 + b;
        if (xx < 0 || xx >= w) continue;
        float f[8];
        unpack8(ld16(xi + ((long long)yy * w + xx) * c), f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {                      // window q = (k + q / 2, j + q % 2): rows a in [2 (q / 2), +3)
          const int a0 = 2 * (q >> 1), b0 = 2 * (q & 1);
          if (a < a0 || a > a0 + 2 || b < b0 || b > b0 + 2) continue;     // compile-time after unrolling
          const int id = (a - a0) * 3 + (b - b0);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (f[e] > mx[q][e]) {
              mx[q][e] = f[e];
              am[q][e] = id;
            }
        }
      }
    }
    float d[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int oh = k + (q >> 1), ow = j + (q & 1);
      if (oh < oh_n && ow < ow_n) {
        unpack8(ld16(dyi + ((long long)oh * ow_n + ow) * c), d[q]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) d[q][e] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int yy = 2 * k + u, xx = 2 * j + v;
        if (yy >= h || xx >= w) continue;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float s = am[0][e] == (1 + u) * 3 + 1 + v ? d[0][e] : 0.f;        // window (k, j): position (1 + u, 1 + v)
          if (v == 1) s += am[1][e] == (1 + u) * 3 ? d[1][e] : 0.f;          // (k, j + 1): its first column
          if (u == 1) s += am[2][e] == 1 + v ? d[2][e] : 0.f;                // (k + 1, j): its first row
          if (u == 1 && v == 1) s += am[3][e] == 0 ? d[3][e] : 0.f;          // (k + 1, j + 1): its corner
          acc[e] = s;
        }
        *reinterpret_cast<uint4*>(dx + (((long long)img * h + yy) * w + xx) * c + g * 8) = pack8(acc);
      }
  }
}

// Filters of a training-step conv, one launch per conv and step: the fp32 parameter [cout][cin][taps] becomes the bf16
// UWM_PACK_TAPS operand of the forward conv, fwd[cout][taps][cin], and of the data-gradient conv (the same kernel run on
// dy with the taps flipped and cin / cout exchanged), dgrad[cin][taps - 1 - t][cout].  A block stages 16 x 16 filters x taps
// in shared memory: contiguous fp32 reads, 32-byte bf16 runs on both write sides (torch needed a cast + permute copy, a
// flip and a transposing copy per conv: ~130 launches and 0.9 ms per step).
constexpr int kWTile = 16;
__global__ void __launch_bounds__(256) pack_train_weights_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                                                 bf16* __restrict__ fwd, bf16* __restrict__ dgrad) {
  extern __shared__ float tile[];                              // [16 cout][16 cin * taps + 1]
  const int pitch = kWTile * taps + 1;
  const int co0 = blockIdx.y * kWTile, ci0 = blockIdx.x * kWTile;
  const int n_co = min(kWTile, cout - co0), n_ci = min(kWTile, cin - ci0);
  const int tx = threadIdx.x & (kWTile - 1), ty = threadIdx.x / kWTile;      // 16 x 16 threads
  if (ty < n_co) {
    const float* src = w + ((size_t)(co0 + ty) * cin + ci0) * taps;           // n_ci * taps contiguous floats of filter row ty
    for (int rem = tx; rem < n_ci * taps; rem += kWTile) tile[ty * pitch + rem] = __ldg(src + rem);
  }
  __syncthreads();
  if (ty < n_co && tx < n_ci) {                                               // (ty, tx) = (cout, cin): runs along cin
    bf16* dst = fwd + (size_t)(co0 + ty) * taps * cin + ci0 + tx;
    for (int t = 0; t < taps; ++t) dst[(size_t)t * cin] = __float2bfloat16_rn(tile[ty * pitch + tx * taps + t]);
  }
  if (dgrad && ty < n_ci && tx < n_co) {                                      // (ty, tx) = (cin, cout): runs along cout
    bf16* dst = dgrad + (size_t)(ci0 + ty) * taps * cout + co0 + tx;
    for (int t = 0; t < taps; ++t) dst[(size_t)(taps - 1 - t) * cout] = __float2bfloat16_rn(tile[tx * pitch + ty * taps + t]);
  }
}

// channels [0, c) of rows with pixel pitch src_pitch -> rows with pixel pitch dst_pitch, 16 bytes per thread: the skip half of
// the decoder's concat (forward) and the skip's slice of the concat gradient (backward); torch's strided copy_ runs these
// through its scalar elementwise kernel (13 launches, 0.52 ms per step).
__global__ void __launch_bounds__(256) copy_channels_kernel(const bf16* __restrict__ src, long long src_pitch,
                                                            bf16* __restrict__ dst, long long dst_pitch, long long pixels, int c) {
  const int groups = c / 8;
  const long long items = pixels * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < items; idx += (long long)gridDim.x * blockDim.x) {
    const long long px = idx / groups;
    const int g = (int)(idx - px * groups);
    *reinterpret_cast<uint4*>(dst + px * dst_pitch + g * 8) = ld16(src + px * src_pitch + g * 8);
  }
}

int check_bn(const char* who, const void* a, const void* b, long long pixels, int c) {
  if (!a || !b) return tfail(UWM_EINVAL, "%s: null tensor", who);
  if (pixels < 1) return tfail(UWM_EINVAL, "%s: pixels = %lld", who, pixels);
  if (c < 8 || c % 8 != 0 || c > 2048) return tfail(UWM_EINVAL, "%s: channels %d (multiple of 8, <= 2048)", who, c);
  return UWM_OK;
}

}  // namespace

extern "C" int uwm_bn_train_forward_nhwc_bf16(const void* d_x, long long pixels, int c, const float* d_gamma,
                                              const float* d_beta, float* d_running_mean, float* d_running_var,
                                              float momentum, float eps, const void* d_residual, int relu, void* d_y,
                                              float* d_save /*[4][c]: mean, rstd, scale, shift*/, double* d_ws /*[slots][2][c], zero*/,
                                              void* stream) {
  int rc = check_bn("bn_train_forward", d_x, d_y, pixels, c);
  if (rc) return rc;
  if (!d_gamma || !d_beta || !d_save || !d_ws) return tfail(UWM_EINVAL, "bn_train_forward: null parameter / workspace");
  if ((d_running_mean == nullptr) != (d_running_var == nullptr))
    return tfail(UWM_EINVAL, "bn_train_forward: running_mean and running_var go together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Geom g = geom(pixels, c, sm_count());
  const bf16* x = static_cast<const bf16*>(d_x);
  bn_stats_kernel<<<g.blocks, g.threads, 2 * c * sizeof(float), st>>>(x, pixels, c, g.groups, g.rpb, d_ws);
  rc = tpost("bn_stats_kernel");
  if (rc) return rc;
  float* save = d_save;
  bn_fwd_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(d_ws, pixels, c, d_gamma, d_beta, d_running_mean, d_running_var,
                                                         momentum, eps, save, save + c, save + 2 * c, save + 3 * c);
  rc = tpost("bn_fwd_finalize_kernel");
  if (rc) return rc;
  if (d_residual)
    bn_apply_kernel<true><<<g.blocks, g.threads, 0, st>>>(x, static_cast<const bf16*>(d_residual), static_cast<bf16*>(d_y),
                                                          pixels, c, g.groups, g.rpb, save + 2 * c, save + 3 * c, relu);
  else
    bn_apply_kernel<false><<<g.blocks, g.threads, 0, st>>>(x, nullptr, static_cast<bf16*>(d_y), pixels, c, g.groups, g.rpb,
                                                           save + 2 * c, save + 3 * c, relu);
  return tpost("bn_apply_kernel");
}

extern "C" int uwm_bn_train_backward_nhwc_bf16(const void* d_dy, const void* d_x, const void* d_y, long long pixels, int c,
                                               const float* d_save /*[4][c] of the forward*/, int relu, int has_residual,
                                               void* d_dx, void* d_dres, float* d_dgamma, float* d_dbeta,
                                               float* d_coef /*[2][c] scratch*/, double* d_ws /*[slots][2][c], zero*/, void* stream) {
  int rc = check_bn("bn_train_backward", d_dy, d_x, pixels, c);
  if (rc) return rc;
  if (!d_save || !d_dx || !d_dgamma || !d_dbeta || !d_coef || !d_ws)
    return tfail(UWM_EINVAL, "bn_train_backward: null output / workspace");
  if (has_residual && (!d_y || !d_dres)) return tfail(UWM_EINVAL, "bn_train_backward: a residual block needs y and d_residual");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Geom g = geom(pixels, c, sm_count());
  const bf16 *dy = static_cast<const bf16*>(d_dy), *x = static_cast<const bf16*>(d_x), *y = static_cast<const bf16*>(d_y);
  const float *mean = d_save, *rstd = d_save + c, *scale = d_save + 2 * c, *shift = d_save + 3 * c;
  if (has_residual)
    bn_bwd_reduce_kernel<true><<<g.blocks, g.threads, 2 * c * sizeof(float), st>>>(dy, x, y, pixels, c, g.groups, g.rpb, scale,
                                                                                  shift, mean, relu, d_ws);
  else
    bn_bwd_reduce_kernel<false><<<g.blocks, g.threads, 2 * c * sizeof(float), st>>>(dy, x, nullptr, pixels, c, g.groups, g.rpb,
                                                                                   scale, shift, mean, relu, d_ws);
  rc = tpost("bn_bwd_reduce_kernel");
  if (rc) return rc;
  bn_bwd_finalize_kernel<<<(c + 7) / 8, 256, 0, st>>>(d_ws, pixels, c, rstd, d_dgamma, d_dbeta, d_coef, d_coef + c);
  rc = tpost("bn_bwd_finalize_kernel");
  if (rc) return rc;
  if (has_residual)
    bn_bwd_apply_kernel<true><<<g.blocks, g.threads, 0, st>>>(dy, x, y, static_cast<bf16*>(d_dx), static_cast<bf16*>(d_dres),
                                                              pixels, c, g.groups, g.rpb, scale, shift, mean, d_coef,
                                                              d_coef + c, relu);
  else
    bn_bwd_apply_kernel<false><<<g.blocks, g.threads, 0, st>>>(dy, x, nullptr, static_cast<bf16*>(d_dx), nullptr, pixels, c,
                                                               g.groups, g.rpb, scale, shift, mean, d_coef, d_coef + c, relu);
  return tpost("bn_bwd_apply_kernel");
}

extern "C" int uwm_upsample2x_backward_nhwc_bf16(const void* d_dy, int n, int h, int w, int c, int dy_pitch, void* d_dx,
                                                 int dx_pitch, void* stream) {
  if (!d_dy || !d_dx) return tfail(UWM_EINVAL, "upsample2x_backward: null tensor");
  if (n < 1 || h < 1 || w < 1 || c < 8 || c % 8 != 0 || dy_pitch % 8 != 0 || dx_pitch % 8 != 0 || dy_pitch < c || dx_pitch < c)
    return tfail(UWM_EINVAL, "upsample2x_backward: bad geometry n=%d h=%d w=%d c=%d pitches %d/%d (channels and pitches multiples of 8)",
                 n, h, w, c, dy_pitch, dx_pitch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long items = (long long)n * h * w * (c / 8);
  const int blocks = (int)std::max(1LL, std::min((items + 255) / 256, (long long)sm_count() * 8));
  upsample2x_bwd_kernel<<<blocks, 256, 0, st>>>(static_cast<const bf16*>(d_dy), dy_pitch, n, h, w, c, static_cast<bf16*>(d_dx),
                                                dx_pitch);
  return tpost("upsample2x_bwd_kernel");
}

extern "C" int uwm_maxpool3x3s2_backward_nhwc_bf16(const void* d_dy, const void* d_x, int n, int h, int w, int c, void* d_dx,
                                                   void* stream) {
  if (!d_dy || !d_x || !d_dx) return tfail(UWM_EINVAL, "maxpool3x3s2_backward: null tensor");
  if (n < 1 || h < 1 || w < 1 || c < 8 || c % 8 != 0)
    return tfail(UWM_EINVAL, "maxpool3x3s2_backward: bad geometry n=%d h=%d w=%d c=%d (channels a multiple of 8)", n, h, w, c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long items = (long long)n * h * w * (c / 8);
  const int blocks = (int)std::max(1LL, std::min((items + 255) / 256, (long long)sm_count() * 16));
  maxpool3x3s2_bwd_kernel<<<blocks, 256, 0, st>>>(static_cast<const bf16*>(d_dy), static_cast<const bf16*>(d_x), n, h, w, c,
                                                  static_cast<bf16*>(d_dx));
  return tpost("maxpool3x3s2_bwd_kernel");
}

extern "C" int uwm_pack_train_weights(const float* d_w, int cout, int cin, int kh, int kw, void* d_fwd, void* d_dgrad,
                                      void* stream) {
  if (!d_w || !d_fwd) return tfail(UWM_EINVAL, "pack_train_weights: null tensor");
  const int taps = kh * kw;
  if (cout < 1 || cin < 1 || kh < 1 || kw < 1 || taps > 49)
    return tfail(UWM_EINVAL, "pack_train_weights: bad filter shape [%d,%d,%d,%d] (at most 49 taps)", cout, cin, kh, kw);
  const size_t smem = (size_t)kWTile * (kWTile * taps + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    static bool raised[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (!raised[dev]) {
      cudaError_t e = cudaFuncSetAttribute(pack_train_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      if (e != cudaSuccess) return tfail(UWM_ECUDA, "pack_train_weights: shared memory opt-in failed: %s", cudaGetErrorString(e));
      raised[dev] = true;
    }
  }
  dim3 grid((cin + kWTile - 1) / kWTile, (cout + kWTile - 1) / kWTile);
  pack_train_weights_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(d_w, cout, cin, taps, static_cast<bf16*>(d_fwd),
                                                                                  static_cast<bf16*>(d_dgrad));
  return tpost("pack_train_weights_kernel");
}

extern "C" int uwm_copy_channels_nhwc_bf16(const void* d_src, long long pixels, int c, int src_pitch, void* d_dst, int dst_pitch,
                                           void* stream) {
  if (!d_src || !d_dst) return tfail(UWM_EINVAL, "copy_channels: null tensor");
  if (pixels < 1 || c < 8 || c % 8 != 0 || src_pitch % 8 != 0 || dst_pitch % 8 != 0 || src_pitch < c || dst_pitch < c ||
      (reinterpret_cast<uintptr_t>(d_src) & 15) || (reinterpret_cast<uintptr_t>(d_dst) & 15))
    return tfail(UWM_EINVAL, "copy_channels: bad geometry pixels=%lld c=%d pitches %d/%d (channels, pitches and base addresses in units of 8 bf16)",
                 pixels, c, src_pitch, dst_pitch);
  const long long items = pixels * (c / 8);
  const int blocks = (int)std::max(1LL, std::min((items + 255) / 256, (long long)sm_count() * 16));
  copy_channels_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(d_src), src_pitch,
                                                                             static_cast<bf16*>(d_dst), dst_pitch, pixels, c);
  return tpost("copy_channels_kernel");
}
