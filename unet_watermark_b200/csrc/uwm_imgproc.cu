// libuwm_b200.so, second translation unit: the integer / byte kernels either side of the network on the reference's
// mask path (reference src/predict.py:588-664 and :161-301).
//
//   resize_u8_kernel          cv2.resize(uint8 RGB, (S,S), INTER_LINEAR)     bit-exact (OpenCV's 11-bit fixed point)
//   upscale_threshold_kernel  cv2.resize(float32 map, (W0,H0)) > thr -> {0,255}, OpenCV's float operation order
//   pack / morph / unpack     binary erosion / dilation with the reference's elliptical / rectangular kernels on
//                             bit-packed rows (32 pixels per word), constant border                 bit-exact
//   ccl_*                     8-connected components (union-find on pixel indices), areas, bounding boxes and
//                             OpenCV's label ORDER (first 2x2 block in raster order)                bit-exact
//
// Ragged batches: every entry takes a table of uwm_image_desc (offset, width, height, pitch) so images of different
// sizes travel in one packed buffer and one launch.  All of this is HBM / L2-bound byte work: one thread per output
// pixel or per 32-pixel word, coalesced, grid.y = image.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/uwm.h"

extern "C" void uwm_internal_set_error(const char* msg);        // uwm_api.cu (thread-local message)
extern "C" void uwm_internal_count_launches(int n);

namespace {

int ifail(int code, const char* fmt, ...) {
  char b[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof(b), fmt, ap);
  va_end(ap);
  uwm_internal_set_error(b);
  return code;
}
int ipost(const char* what, int n = 1) {
  uwm_internal_count_launches(n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ifail(UWM_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return UWM_OK;
}

constexpr int kCoefBits = 11;
constexpr float kCoefScale = 2048.f;

// OpenCV resize.cpp: scale = 1 / (dsize / ssize) in double; f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s
__device__ __forceinline__ void linear_tap(int d, double scale, int* s, float* f) {
  const float ff = (float)(((double)d + 0.5) * scale - 0.5);
  const int ss = (int)floorf(ff);
  *s = ss;
  *f = ff - (float)ss;
}
__device__ __forceinline__ double inv_inv(int ssize, int dsize) {
  const double inv_scale = (double)dsize / (double)ssize;
  return 1.0 / inv_scale;
}
// saturate_cast<short>(float): round half to even, clamp
__device__ __forceinline__ int round_short(float v) {
  int r = __float2int_rn(v);
  return max(-32768, min(32767, r));
}

// ------------------------------------------------------------------------------------------------------------
// cv2.resize(uint8 HxWx3, (dw, dh), INTER_LINEAR), one thread per destination pixel (3 channels).
// swap_rb: the source is BGR as cv2.imread returns it, the destination RGB (cvtColor folded into the read).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, const uwm_image_desc* __restrict__ desc,
                                                        int dw, int dh, int swap_rb, uint8_t* __restrict__ out) {
  const int img = blockIdx.y;
  const uwm_image_desc d = desc[img];
  const int sw = d.width, sh = d.height;
  const uint8_t* s = src + d.offset;
  const long long pitch = d.pitch;
  uint8_t* o = out + (long long)img * dw * dh * 3;
  const int total = dw * dh;
  const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;
  const bool same = (sw == dw && sh == dh);
  const bool area2 = (sw == 2 * dw && sh == 2 * dh);
  const double scale_x = inv_inv(sw, dw), scale_y = inv_inv(sh, dh);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int dy = idx / dw, dx = idx - dy * dw;
    int r[3];
    if (same) {
      const uint8_t* p = s + dy * pitch + dx * 3;
      r[0] = p[c0]; r[1] = p[1]; r[2] = p[c2];
    } else if (area2) {
      // cv::resize switches INTER_LINEAR to the fast INTER_AREA path for an exact 2x reduction: (a+b+c+d+2) >> 2
      const uint8_t* p0 = s + (2 * dy) * pitch + (2 * dx) * 3;
      const uint8_t* p1 = p0 + pitch;
      const int cc[3] = {c0, 1, c2};
#pragma unroll
      for (int c = 0; c < 3; ++c) r[c] = (p0[cc[c]] + p0[3 + cc[c]] + p1[cc[c]] + p1[3 + cc[c]] + 2) >> 2;
    } else {
      int sx, sy; float fx, fy;
      linear_tap(dx, scale_x, &sx, &fx);
      linear_tap(dy, scale_y, &sy, &fy);
      if (sx < 0) { sx = 0; fx = 0.f; }
      if (sx >= sw - 1) { sx = sw - 1; fx = 0.f; }
      const int a0 = round_short((1.f - fx) * kCoefScale), a1 = round_short(fx * kCoefScale);
      const int b0 = round_short((1.f - fy) * kCoefScale), b1 = round_short(fy * kCoefScale);
      const int sx1 = min(sx + 1, sw - 1);
      const int y0 = max(0, min(sy, sh - 1)), y1 = max(0, min(sy + 1, sh - 1));
      const uint8_t* p00 = s + y0 * pitch + sx * 3;
      const uint8_t* p01 = s + y0 * pitch + sx1 * 3;
      const uint8_t* p10 = s + y1 * pitch + sx * 3;
      const uint8_t* p11 = s + y1 * pitch + sx1 * 3;
      const int cc[3] = {c0, 1, c2};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int r0 = p00[cc[c]] * a0 + p01[cc[c]] * a1;         // horizontal pass, scaled by 2^11
        const int r1 = p10[cc[c]] * a0 + p11[cc[c]] * a1;
        r[c] = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;   // VResizeLinear<uchar>, FixedPtCast<22>
        r[c] = max(0, min(255, r[c]));
      }
    }
    uint8_t* q = o + (long long)idx * 3;
    q[0] = (uint8_t)r[0]; q[1] = (uint8_t)r[1]; q[2] = (uint8_t)r[2];
  }
}

// ------------------------------------------------------------------------------------------------------------
// cv2.resize(float32 [sh, sw], (W0, H0)) > thr  ->  uint8 {0, 255}   (reference src/predict.py:620-625)
// Float taps and OpenCV's operation order: rows = S[sx]*(1-fx) + S[sx+1]*fx, then rows0*(1-fy) + rows1*fy, every
// product and sum rounded separately (no FMA contraction).  One thread per destination pixel.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upscale_threshold_kernel(const float* __restrict__ maps, int sw, int sh,
                                                                const uwm_image_desc* __restrict__ desc, float thr,
                                                                uint8_t* __restrict__ out, float* __restrict__ out_f32) {
  const int img = blockIdx.y;
  const uwm_image_desc d = desc[img];
  const int dw = d.width, dh = d.height;
  const float* s = maps + (long long)img * sw * sh;
  uint8_t* o = out ? out + d.offset : nullptr;
  float* of = out_f32 ? out_f32 + d.offset : nullptr;              // debug/parity tap: the resized float map (pitch in elements)
  const long long total = (long long)dw * dh;
  const bool same = (sw == dw && sh == dh);
  const bool area2 = (sw == 2 * dw && sh == 2 * dh);
  const double scale_x = inv_inv(sw, dw), scale_y = inv_inv(sh, dh);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int dy = (int)(idx / dw), dx = (int)(idx - (long long)dy * dw);
    float v;
    if (same) {
      v = s[(long long)dy * sw + dx];
    } else if (area2) {
      const float* p = s + (long long)(2 * dy) * sw + 2 * dx;
      v = __fmul_rn(__fadd_rn(__fadd_rn(p[0], p[1]), __fadd_rn(p[sw], p[sw + 1])), 0.25f);   // OpenCV's vector body: (a+b)+(c+d)
    } else {
      int sx, sy; float fx, fy;
      linear_tap(dx, scale_x, &sx, &fx);
      linear_tap(dy, scale_y, &sy, &fy);
      if (sx < 0) { sx = 0; fx = 0.f; }
      if (sx >= sw - 1) { sx = sw - 1; fx = 0.f; }
      const int sx1 = min(sx + 1, sw - 1);
      const int y0 = max(0, min(sy, sh - 1)), y1 = max(0, min(sy + 1, sh - 1));
      const float a0 = 1.f - fx, b0 = 1.f - fy;
      const float r0 = __fadd_rn(__fmul_rn(s[(long long)y0 * sw + sx], a0), __fmul_rn(s[(long long)y0 * sw + sx1], fx));
      const float r1 = __fadd_rn(__fmul_rn(s[(long long)y1 * sw + sx], a0), __fmul_rn(s[(long long)y1 * sw + sx1], fx));
      v = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, fy));
    }
    if (o) o[(long long)dy * d.pitch + dx] = (v > thr) ? 255 : 0;
    if (of) of[(long long)dy * d.pitch + dx] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// bit-packed binary morphology
// ------------------------------------------------------------------------------------------------------------
struct MorphSE {           // a structuring element as one contiguous run per row (every OpenCV ellipse / rect / cross row is one)
  int rows, ay;
  int lo[16], hi[16];      // run of row i as pixel offsets [lo, hi] relative to the anchor column; lo > hi: empty row
};

struct BitImg { long long word_off; int w, h, wpr; };   // wpr = words per row

// uint8 mask (> 127, i.e. cv2.threshold(mask, 127, 255, THRESH_BINARY)) -> bits; bit b of word k = pixel 32k + b
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t* __restrict__ masks, const uwm_image_desc* __restrict__ desc,
                                                        const BitImg* __restrict__ bi, uint32_t* __restrict__ bits) {
  const int img = blockIdx.y;
  const uwm_image_desc d = desc[img];
  const BitImg b = bi[img];
  const int lane = threadIdx.x & 31;
  const long long nwords = (long long)b.h * b.wpr;
  // one warp per word: lane = pixel (coalesced 32-byte read, ballot)
  for (long long wd = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; wd < nwords; wd += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int y = (int)(wd / b.wpr), k = (int)(wd - (long long)y * b.wpr);
    const int x = 32 * k + lane;
    const bool on = (x < b.w) && masks[d.offset + (long long)y * d.pitch + x] > 127;
    const uint32_t m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[b.word_off + wd] = m;
  }
}

__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint32_t* __restrict__ bits, const BitImg* __restrict__ bi,
                                                          const uwm_image_desc* __restrict__ desc, uint8_t* __restrict__ masks) {
  const int img = blockIdx.y;
  const uwm_image_desc d = desc[img];
  const BitImg b = bi[img];
  const long long total = (long long)b.h * b.w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(idx / b.w), x = (int)(idx - (long long)y * b.w);
    const uint32_t wv = bits[b.word_off + (long long)y * b.wpr + (x >> 5)];
    masks[d.offset + (long long)y * d.pitch + x] = ((wv >> (x & 31)) & 1u) ? 255 : 0;
  }
}

// One erosion / dilation pass, OpenCV semantics: dst(y,x) = max|min over SE elements (i,j) of src(y + i - ay, x + j - ax),
// out-of-image pixels never win (constant border: 0 for dilation, 1 for erosion).  Erosion runs as the complement of
// the dilation of the complement (same offsets), with the complement's out-of-image bits forced to 0.
// One thread per output word: window = words k-1, k, k+1 of every source row the SE touches (|offset| < 32).
template <bool ERODE>
__global__ void __launch_bounds__(256) morph_bits_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                         const BitImg* __restrict__ bi, const __grid_constant__ MorphSE se) {
  const int img = blockIdx.y;
  const BitImg b = bi[img];
  const long long nwords = (long long)b.h * b.wpr;
  const uint32_t* src = in + b.word_off;
  const uint32_t tail = (b.w & 31) ? ((1u << (b.w & 31)) - 1u) : 0xffffffffu;     // valid bits of a row's last word
  for (long long wd = blockIdx.x * (long long)blockDim.x + threadIdx.x; wd < nwords; wd += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(wd / b.wpr), k = (int)(wd - (long long)y * b.wpr);
    uint32_t acc = 0;
    for (int i = 0; i < se.rows; ++i) {
      const int lo = se.lo[i], hi = se.hi[i];
      const int yy = y + i - se.ay;
      if (lo > hi || yy < 0 || yy >= b.h) continue;
      const uint32_t* row = src + (long long)yy * b.wpr;
      auto ld = [&](int kk) -> uint32_t {
        if (kk < 0 || kk >= b.wpr) return 0u;
        uint32_t v = row[kk];
        if (ERODE) { v = ~v; if (kk == b.wpr - 1) v &= tail; }
        return v;
      };
      const uint32_t prev = ld(k - 1), cur = ld(k), next = ld(k + 1);
      for (int dx = lo; dx <= hi; ++dx) {
        // bit x of the result takes source bit x + dx
        if (dx >= 0) acc |= __funnelshift_r(cur, next, dx);
        else acc |= __funnelshift_l(prev, cur, -dx);
      }
    }
    if (ERODE) acc = ~acc;
    if (k == b.wpr - 1) acc &= tail;
    out[b.word_off + wd] = acc;
  }
}

__global__ void __launch_bounds__(256) or_bits_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b2,
                                                      uint32_t* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] | b2[i];
}

// ------------------------------------------------------------------------------------------------------------
// 8-connected components: union-find over pixel indices (root = smallest index of the component)
// ------------------------------------------------------------------------------------------------------------
struct CclImg { long long pix_off; long long word_off; int w, h, wpr; int pad; };

__device__ __forceinline__ int uf_find(const int* __restrict__ L, int a) {
  int p = L[a];
  while (p != a) { a = p; p = L[a]; }
  return a;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
    else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}
__device__ __forceinline__ bool bit_at(const uint32_t* __restrict__ bits, const CclImg& c, int y, int x) {
  if (x < 0 || y < 0 || x >= c.w || y >= c.h) return false;
  return (bits[c.word_off + (long long)y * c.wpr + (x >> 5)] >> (x & 31)) & 1u;
}

__global__ void __launch_bounds__(256) ccl_init_kernel(const uint32_t* __restrict__ bits, const CclImg* __restrict__ ci,
                                                       int* __restrict__ labels, int* __restrict__ area, int* __restrict__ order,
                                                       int* __restrict__ bbox) {
  const CclImg c = ci[blockIdx.y];
  const long long total = (long long)c.w * c.h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(idx / c.w), x = (int)(idx - (long long)y * c.w);
    labels[c.pix_off + idx] = bit_at(bits, c, y, x) ? (int)idx : -1;
    area[c.pix_off + idx] = 0;
    order[c.pix_off + idx] = 0x7fffffff;
    if (bbox) {
      bbox[4 * (c.pix_off + idx) + 0] = 0x7fffffff; bbox[4 * (c.pix_off + idx) + 1] = 0x7fffffff;
      bbox[4 * (c.pix_off + idx) + 2] = -1; bbox[4 * (c.pix_off + idx) + 3] = -1;
    }
  }
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const CclImg* __restrict__ ci, int* __restrict__ labels) {
  const CclImg c = ci[blockIdx.y];
  int* L = labels + c.pix_off;
  const long long total = (long long)c.w * c.h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (L[idx] < 0) continue;
    const int y = (int)(idx / c.w), x = (int)(idx - (long long)y * c.w);
    // W, NW, N, NE (the other four directions are covered from the neighbour's side)
    if (x > 0 && L[idx - 1] >= 0) uf_union(L, (int)idx, (int)idx - 1);
    if (y > 0) {
      const long long up = idx - c.w;
      if (L[up] >= 0) uf_union(L, (int)idx, (int)up);
      else {      // N is background: NW and NE are not connected through it
        if (x > 0 && L[up - 1] >= 0) uf_union(L, (int)idx, (int)up - 1);
        if (x + 1 < c.w && L[up + 1] >= 0) uf_union(L, (int)idx, (int)up + 1);
      }
    }
  }
}

// flatten + per-root statistics: area, OpenCV order key (first 2x2 block in raster order), bounding box
__global__ void __launch_bounds__(256) ccl_stats_kernel(const CclImg* __restrict__ ci, int* __restrict__ labels,
                                                        int* __restrict__ area, int* __restrict__ order, int* __restrict__ bbox) {
  const CclImg c = ci[blockIdx.y];
  int* L = labels + c.pix_off;
  const long long total = (long long)c.w * c.h;
  const int w2 = (c.w + 1) >> 1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (L[idx] < 0) continue;
    const int r = uf_find(L, (int)idx);
    L[idx] = r;                                 // roots keep L[r] == r: concurrent finds stay correct
    const int y = (int)(idx / c.w), x = (int)(idx - (long long)y * c.w);
    // warp-aggregated area count: lanes with the same root add once
    const unsigned peers = __match_any_sync(__activemask(), r);
    if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&area[c.pix_off + r], __popc(peers));
    atomicMin(&order[c.pix_off + r], (y >> 1) * w2 + (x >> 1));
    if (bbox) {
      int* bb = bbox + 4 * (c.pix_off + r);
      atomicMin(&bb[0], x); atomicMin(&bb[1], y); atomicMax(&bb[2], x); atomicMax(&bb[3], y);
    }
  }
}

// per image: the largest component, ties broken by OpenCV's label order (np.argmax takes the first maximum)
__global__ void __launch_bounds__(256) ccl_best_kernel(const CclImg* __restrict__ ci, const int* __restrict__ labels,
                                                       const int* __restrict__ area, const int* __restrict__ order,
                                                       unsigned long long* __restrict__ best) {
  const CclImg c = ci[blockIdx.y];
  const long long total = (long long)c.w * c.h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (labels[c.pix_off + idx] != (int)idx) continue;            // roots only
    const unsigned long long key = ((unsigned long long)(unsigned)area[c.pix_off + idx] << 32) |
                                   (unsigned long long)(0xffffffffu - (unsigned)order[c.pix_off + idx]);
    atomicMax(&best[blockIdx.y], key);
  }
}

// mode 0 (reference _optimize_watermark_mask :251-266): keep the largest component; if it has fewer than 500 pixels keep
// every component with more than 200 instead.  mode 1 / 2 (text / mixed, :217-228, :289-299): keep area > area_thr.
__global__ void __launch_bounds__(256) ccl_select_kernel(const CclImg* __restrict__ ci, const int* __restrict__ labels,
                                                         const int* __restrict__ area, const int* __restrict__ order,
                                                         const unsigned long long* __restrict__ best, int mode, int area_thr,
                                                         uint32_t* __restrict__ bits_out) {
  const CclImg c = ci[blockIdx.y];
  const long long nwords = (long long)c.h * c.wpr;
  const unsigned long long bk = best[blockIdx.y];
  const int best_area = (int)(bk >> 32);
  const int best_order = (int)(0xffffffffu - (unsigned)(bk & 0xffffffffu));
  for (long long wd = blockIdx.x * (long long)blockDim.x + threadIdx.x; wd < nwords; wd += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(wd / c.wpr), k = (int)(wd - (long long)y * c.wpr);
    uint32_t m = 0;
    for (int b = 0; b < 32; ++b) {
      const int x = 32 * k + b;
      if (x >= c.w) break;
      const int r = labels[c.pix_off + (long long)y * c.w + x];
      if (r < 0) continue;
      const int a = area[c.pix_off + r];
      bool keep;
      if (mode == 0) keep = (best_area >= 500) ? (a == best_area && order[c.pix_off + r] == best_order) : (a > 200);
      else keep = a > area_thr;
      if (keep) m |= 1u << b;
    }
    bits_out[c.word_off + wd] = m;
  }
}

// reference _analyze_text_features (:443-508): per component a score from aspect ratio, fill density and area;
// counts components with score > 0.5.  score_mask bit (3*ia + ib)*3 + ic tells whether the float sum of the three
// partial scores exceeds 0.5 (tabulated by the host in double precision, as Python evaluates it).
__global__ void __launch_bounds__(256) ccl_text_features_kernel(const CclImg* __restrict__ ci, const int* __restrict__ labels,
                                                                const int* __restrict__ area, const int* __restrict__ bbox,
                                                                unsigned score_mask, int* __restrict__ out /* [n][2] */) {
  const CclImg c = ci[blockIdx.y];
  const long long total = (long long)c.w * c.h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (labels[c.pix_off + idx] != (int)idx) continue;
    const long long a = area[c.pix_off + idx];
    const int* bb = bbox + 4 * (c.pix_off + idx);
    const long long wdt = bb[2] - bb[0] + 1, hgt = bb[3] - bb[1] + 1;
    const long long mx = max(wdt, hgt), mn = min(wdt, hgt), box = wdt * hgt;
    // exact rational forms of the reference's float comparisons (all operands are small integers)
    const int ia = (mx <= 5 * mn) ? 0 : (mx <= 10 * mn ? 1 : 2);                            // 1 <= ar <= 5 | 5 < ar <= 10 | else
    int ib;
    if (10 * a >= 3 * box && 10 * a <= 8 * box) ib = 0;                                     // 0.3 <= density <= 0.8
    else if ((10 * a >= 2 * box && 10 * a < 3 * box) || (10 * a > 8 * box && 10 * a <= 9 * box)) ib = 1;
    else ib = 2;
    const int ic = (a >= 50 && a <= 5000) ? 0 : (((a >= 20 && a < 50) || (a > 5000 && a <= 10000)) ? 1 : 2);
    atomicAdd(&out[2 * blockIdx.y + 1], 1);
    if ((score_mask >> ((3 * ia + ib) * 3 + ic)) & 1u) atomicAdd(&out[2 * blockIdx.y], 1);
  }
}

// per image: {foreground pixels, components, largest component area} (reference src/scripts/model_selector.py:171-197)
__global__ void __launch_bounds__(256) ccl_summary_kernel(const CclImg* __restrict__ ci, const int* __restrict__ labels,
                                                          const int* __restrict__ area, int* __restrict__ out /* [n][3] */) {
  const CclImg c = ci[blockIdx.y];
  const long long total = (long long)c.w * c.h;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (labels[c.pix_off + idx] != (int)idx) continue;            // roots only
    const int a = area[c.pix_off + idx];
    atomicAdd(&out[3 * blockIdx.y + 0], a);
    atomicAdd(&out[3 * blockIdx.y + 1], 1);
    atomicMax(&out[3 * blockIdx.y + 2], a);
  }
}

unsigned grid_x(long long items, int threads = 256) {
  const long long blocks = (items + threads - 1) / threads;
  return (unsigned)std::max(1LL, std::min(blocks, 148LL * 8));
}

MorphSE make_se(int shape, int cols, int rows) {
  // cv2.getStructuringElement(shape, (cols, rows)), default anchor (cols/2, rows/2); 0 rect, 1 cross, 2 ellipse
  MorphSE se;
  memset(&se, 0, sizeof(se));
  se.rows = rows; se.ay = rows / 2;
  const int r = rows / 2, c = cols / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  for (int i = 0; i < rows; ++i) {
    int j1 = 0, j2 = 0;
    if (shape == 0 || (rows == 1 && cols == 1)) { j1 = 0; j2 = cols; }
    else if (shape == 1) { if (i == r) { j1 = 0; j2 = cols; } else { j1 = c; j2 = c + 1; } }
    else {
      const int dy = i - r;
      if (std::abs(dy) <= r) {
        const int dx = (int)std::lrint(c * std::sqrt(((double)r * r - (double)dy * dy) * inv_r2));
        j1 = std::max(c - dx, 0); j2 = std::min(c + dx + 1, cols);
      }
    }
    se.lo[i] = j1 - c; se.hi[i] = j2 - 1 - c;
  }
  return se;
}

struct Workspace {           // carve-up of the caller's device workspace
  uint8_t* base; size_t size, used;
  template <typename T> T* take(size_t n) {
    used = (used + 255) & ~(size_t)255;
    T* p = reinterpret_cast<T*>(base + used);
    used += n * sizeof(T);
    return used <= size ? p : nullptr;
  }
};

struct Geometry {
  std::vector<BitImg> bi;
  std::vector<CclImg> ci;
  long long words = 0, pixels = 0, max_words = 0, max_pixels = 0;
};
Geometry geometry(const uwm_image_desc* h, int n) {
  Geometry g;
  for (int i = 0; i < n; ++i) {
    const int wpr = (h[i].width + 31) / 32;
    g.bi.push_back({g.words, h[i].width, h[i].height, wpr});
    g.ci.push_back({g.pixels, g.words, h[i].width, h[i].height, wpr, 0});
    const long long nw = (long long)wpr * h[i].height, np = (long long)h[i].width * h[i].height;
    g.words += nw; g.pixels += np;
    g.max_words = std::max(g.max_words, nw); g.max_pixels = std::max(g.max_pixels, np);
  }
  return g;
}

int check_descs(const uwm_image_desc* h, int n, const char* what) {
  if (!h || n < 1) return ifail(UWM_EINVAL, "%s: need at least one image descriptor", what);
  for (int i = 0; i < n; ++i) {
    if (h[i].width < 1 || h[i].height < 1 || h[i].pitch < h[i].width || h[i].offset < 0)
      return ifail(UWM_EINVAL, "%s: image %d has a bad descriptor (%dx%d pitch %d offset %lld)", what, i, h[i].width,
                   h[i].height, h[i].pitch, (long long)h[i].offset);
    if ((long long)h[i].width * h[i].height >= (1LL << 31))
      return ifail(UWM_EINVAL, "%s: image %d has 2^31 pixels or more", what, i);
  }
  return UWM_OK;
}

size_t post_workspace_bytes(const Geometry& g, int n, bool bbox) {
  size_t b = 0;
  auto add = [&](size_t bytes) { b = ((b + 255) & ~(size_t)255) + bytes; };
  add(sizeof(BitImg) * n); add(sizeof(CclImg) * n);
  add(4 * (size_t)g.words); add(4 * (size_t)g.words); add(4 * (size_t)g.words);   // three bit planes (ping, pong, branch)
  add(4 * (size_t)g.pixels); add(4 * (size_t)g.pixels); add(4 * (size_t)g.pixels); // labels, area, order
  if (bbox) add(16 * (size_t)g.pixels);
  add(8 * (size_t)n); add(8 * (size_t)n);
  return b + 256;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" int uwm_resize_bilinear_u8(const uint8_t* d_src, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                                      int dst_w, int dst_h, int swap_rb, uint8_t* d_dst, void* stream) {
  if (!d_src || !d_desc || !d_dst) return ifail(UWM_EINVAL, "resize_bilinear_u8: null pointer");
  int rc = check_descs(h_desc, n, "resize_bilinear_u8");
  if (rc) return rc;
  if (dst_w < 1 || dst_h < 1) return ifail(UWM_EINVAL, "resize_bilinear_u8: bad destination size %dx%d", dst_w, dst_h);
  for (int i = 0; i < n; ++i)
    if (h_desc[i].pitch < 3 * h_desc[i].width) return ifail(UWM_EINVAL, "resize_bilinear_u8: image %d: pitch < 3 * width", i);
  dim3 grid(grid_x((long long)dst_w * dst_h), n);
  resize_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_src, d_desc, dst_w, dst_h, swap_rb, d_dst);
  return ipost("resize_u8_kernel");
}

extern "C" int uwm_mask_upscale_threshold(const float* d_maps, int n, int src_w, int src_h, const uwm_image_desc* h_desc,
                                          const uwm_image_desc* d_desc, float threshold, uint8_t* d_masks,
                                          float* d_resized_f32, void* stream) {
  if (!d_maps || !d_desc || (!d_masks && !d_resized_f32)) return ifail(UWM_EINVAL, "mask_upscale_threshold: null pointer");
  int rc = check_descs(h_desc, n, "mask_upscale_threshold");
  if (rc) return rc;
  long long mx = 0;
  for (int i = 0; i < n; ++i) mx = std::max(mx, (long long)h_desc[i].width * h_desc[i].height);
  dim3 grid(grid_x(mx), n);
  upscale_threshold_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_maps, src_w, src_h, d_desc, threshold,
                                                                                 d_masks, d_resized_f32);
  return ipost("upscale_threshold_kernel");
}

extern "C" size_t uwm_mask_postprocess_workspace(const uwm_image_desc* h_desc, int n) {
  if (!h_desc || n < 1) return 0;
  return post_workspace_bytes(geometry(h_desc, n), n, true);
}

namespace {

struct PostCtx {
  cudaStream_t st;
  int n;
  Geometry g;
  BitImg* d_bi; CclImg* d_ci;
  uint32_t *b0, *b1, *b2;
  int *labels, *area, *order, *bbox;
  unsigned long long* best;
  int* feat;
  int launches = 0;
};

int post_setup(PostCtx& c, const uwm_image_desc* h_desc, int n, void* d_ws, size_t ws_bytes, bool bbox, void* stream) {
  c.st = static_cast<cudaStream_t>(stream);
  c.n = n;
  c.g = geometry(h_desc, n);
  if (ws_bytes < post_workspace_bytes(c.g, n, bbox))
    return ifail(UWM_EINVAL, "mask post-processing: workspace of %zu bytes, need %zu (uwm_mask_postprocess_workspace)", ws_bytes,
                 post_workspace_bytes(c.g, n, bbox));
  Workspace ws{static_cast<uint8_t*>(d_ws), ws_bytes, 0};
  c.d_bi = ws.take<BitImg>(n); c.d_ci = ws.take<CclImg>(n);
  c.b0 = ws.take<uint32_t>(c.g.words); c.b1 = ws.take<uint32_t>(c.g.words); c.b2 = ws.take<uint32_t>(c.g.words);
  c.labels = ws.take<int>(c.g.pixels); c.area = ws.take<int>(c.g.pixels); c.order = ws.take<int>(c.g.pixels);
  c.bbox = bbox ? ws.take<int>(4 * c.g.pixels) : nullptr;
  c.best = ws.take<unsigned long long>(n);
  c.feat = reinterpret_cast<int*>(ws.take<unsigned long long>(n));
  if (!c.feat) return ifail(UWM_EINVAL, "mask post-processing: workspace too small");
  // geometry tables: small pageable copies, ordered on the stream
  if (cudaMemcpyAsync(c.d_bi, c.g.bi.data(), sizeof(BitImg) * n, cudaMemcpyHostToDevice, c.st) != cudaSuccess ||
      cudaMemcpyAsync(c.d_ci, c.g.ci.data(), sizeof(CclImg) * n, cudaMemcpyHostToDevice, c.st) != cudaSuccess ||
      cudaStreamSynchronize(c.st) != cudaSuccess)     // the host vectors die with this call
    return ifail(UWM_ECUDA, "mask post-processing: geometry upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  return UWM_OK;
}

void morph(PostCtx& c, bool erode, int shape, int cols, int rows, int iterations, uint32_t*& cur, uint32_t*& other) {
  const MorphSE se = make_se(shape, cols, rows);
  dim3 grid(grid_x(c.g.max_words), c.n);
  for (int it = 0; it < iterations; ++it) {
    if (erode) morph_bits_kernel<true><<<grid, 256, 0, c.st>>>(cur, other, c.d_bi, se);
    else morph_bits_kernel<false><<<grid, 256, 0, c.st>>>(cur, other, c.d_bi, se);
    std::swap(cur, other);
    ++c.launches;
  }
}
void m_open(PostCtx& c, int shape, int cols, int rows, int it, uint32_t*& cur, uint32_t*& other) {
  morph(c, true, shape, cols, rows, it, cur, other); morph(c, false, shape, cols, rows, it, cur, other);
}
void m_close(PostCtx& c, int shape, int cols, int rows, int it, uint32_t*& cur, uint32_t*& other) {
  morph(c, false, shape, cols, rows, it, cur, other); morph(c, true, shape, cols, rows, it, cur, other);
}

void run_ccl(PostCtx& c, const uint32_t* bits, bool bbox) {
  dim3 grid(grid_x(c.g.max_pixels), c.n);
  ccl_init_kernel<<<grid, 256, 0, c.st>>>(bits, c.d_ci, c.labels, c.area, c.order, bbox ? c.bbox : nullptr);
  ccl_merge_kernel<<<grid, 256, 0, c.st>>>(c.d_ci, c.labels);
  ccl_stats_kernel<<<grid, 256, 0, c.st>>>(c.d_ci, c.labels, c.area, c.order, bbox ? c.bbox : nullptr);
  c.launches += 3;
}

}  // namespace

extern "C" int uwm_mask_postprocess(uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                                    int mode, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (!d_masks || !d_desc || !d_workspace) return ifail(UWM_EINVAL, "mask_postprocess: null pointer");
  if (mode < 0 || mode > 2) return ifail(UWM_EINVAL, "mask_postprocess: mode %d (0 watermark, 1 text, 2 mixed)", mode);
  int rc = check_descs(h_desc, n, "mask_postprocess");
  if (rc) return rc;
  PostCtx c;
  rc = post_setup(c, h_desc, n, d_workspace, workspace_bytes, false, stream);
  if (rc) return rc;
  dim3 gw(grid_x(c.g.max_words * 32), n);
  pack_bits_kernel<<<gw, 256, 0, c.st>>>(d_masks, d_desc, c.d_bi, c.b0);        // cv2.threshold(mask, 127, 255) (:175)
  ++c.launches;
  uint32_t *cur = c.b0, *other = c.b1;
  const int E = 2, R = 0;
  int area_thr = 0;
  if (mode == 0) {                      // reference _optimize_watermark_mask :232-249
    m_open(c, E, 3, 3, 1, cur, other);
    m_close(c, E, 7, 7, 3, cur, other);
    m_close(c, E, 11, 11, 2, cur, other);
    morph(c, false, E, 9, 9, 2, cur, other);
  } else if (mode == 1) {               // reference _optimize_text_mask :192-215
    m_open(c, E, 2, 2, 1, cur, other);
    m_close(c, E, 3, 3, 2, cur, other);
    // horizontal and vertical closings of the same input, OR-ed
    uint32_t* base = cur;                // keep `base`, work in (other, b2)
    uint32_t* h_cur = other; uint32_t* h_tmp = c.b2;
    {   // mask_h = close(base, rect 5x1)
      const MorphSE se = make_se(R, 5, 1);
      dim3 grid(grid_x(c.g.max_words), n);
      morph_bits_kernel<false><<<grid, 256, 0, c.st>>>(base, h_cur, c.d_bi, se);
      morph_bits_kernel<true><<<grid, 256, 0, c.st>>>(h_cur, h_tmp, c.d_bi, se);      // mask_h in h_tmp (= b2)
      const MorphSE sv = make_se(R, 1, 5);
      morph_bits_kernel<false><<<grid, 256, 0, c.st>>>(base, h_cur, c.d_bi, sv);
      morph_bits_kernel<true><<<grid, 256, 0, c.st>>>(h_cur, base, c.d_bi, sv);       // mask_v overwrites base
      or_bits_kernel<<<grid_x(c.g.words), 256, 0, c.st>>>(h_tmp, base, h_cur, c.g.words);
      c.launches += 5;
      cur = h_cur; other = base;
    }
    morph(c, false, E, 4, 4, 1, cur, other);
    area_thr = 50;
  } else {                              // reference _optimize_mixed_mask :275-287
    m_open(c, E, 2, 2, 1, cur, other);
    m_close(c, E, 5, 5, 2, cur, other);
    morph(c, false, E, 6, 6, 1, cur, other);
    area_thr = 100;
  }
  run_ccl(c, cur, false);
  dim3 gp(grid_x(c.g.max_pixels), n);
  cudaMemsetAsync(c.best, 0, 8 * (size_t)n, c.st);
  ccl_best_kernel<<<gp, 256, 0, c.st>>>(c.d_ci, c.labels, c.area, c.order, c.best);
  dim3 gs(grid_x(c.g.max_words), n);
  ccl_select_kernel<<<gs, 256, 0, c.st>>>(c.d_ci, c.labels, c.area, c.order, c.best, mode, area_thr, other);
  // the final GaussianBlur(3x3, 0.5) + threshold 127 of the watermark branch (:269-271) is the identity on a {0,255}
  // image (centre weight 0.619 > 127/255 > 0.381 = all the others together): nothing to launch
  unpack_bits_kernel<<<gp, 256, 0, c.st>>>(other, c.d_bi, d_desc, d_masks);
  c.launches += 3;
  return ipost("mask post-processing kernels", c.launches);
}

extern "C" int uwm_mask_text_features(const uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                                      unsigned score_mask, int32_t* d_out, void* d_workspace, size_t workspace_bytes,
                                      void* stream) {
  if (!d_masks || !d_desc || !d_out || !d_workspace) return ifail(UWM_EINVAL, "mask_text_features: null pointer");
  int rc = check_descs(h_desc, n, "mask_text_features");
  if (rc) return rc;
  PostCtx c;
  rc = post_setup(c, h_desc, n, d_workspace, workspace_bytes, true, stream);
  if (rc) return rc;
  dim3 gw(grid_x(c.g.max_words * 32), n);
  pack_bits_kernel<<<gw, 256, 0, c.st>>>(d_masks, d_desc, c.d_bi, c.b0);
  ++c.launches;
  run_ccl(c, c.b0, true);
  cudaMemsetAsync(d_out, 0, 8 * (size_t)n, c.st);
  dim3 gp(grid_x(c.g.max_pixels), n);
  ccl_text_features_kernel<<<gp, 256, 0, c.st>>>(c.d_ci, c.labels, c.area, c.bbox, score_mask, d_out);
  ++c.launches;
  return ipost("mask text-feature kernels", c.launches);
}

// op-level entry for the parity tests: one morphology operation on a ragged batch of uint8 masks, in place.
//   op: 0 erode, 1 dilate, 2 open, 3 close;  shape: 0 rect, 1 cross, 2 ellipse (cv2.MORPH_*)
extern "C" int uwm_mask_morphology(uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n, int op,
                                   int shape, int ksize_w, int ksize_h, int iterations, void* d_workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!d_masks || !d_desc || !d_workspace) return ifail(UWM_EINVAL, "mask_morphology: null pointer");
  if (op < 0 || op > 3 || shape < 0 || shape > 2 || ksize_w < 1 || ksize_h < 1 || ksize_w > 16 || ksize_h > 16 || iterations < 1)
    return ifail(UWM_EINVAL, "mask_morphology: bad operation / kernel (op %d shape %d %dx%d x%d)", op, shape, ksize_w, ksize_h, iterations);
  int rc = check_descs(h_desc, n, "mask_morphology");
  if (rc) return rc;
  PostCtx c;
  rc = post_setup(c, h_desc, n, d_workspace, workspace_bytes, false, stream);
  if (rc) return rc;
  dim3 gw(grid_x(c.g.max_words * 32), n);
  pack_bits_kernel<<<gw, 256, 0, c.st>>>(d_masks, d_desc, c.d_bi, c.b0);
  uint32_t *cur = c.b0, *other = c.b1;
  if (op == 0) morph(c, true, shape, ksize_w, ksize_h, iterations, cur, other);
  else if (op == 1) morph(c, false, shape, ksize_w, ksize_h, iterations, cur, other);
  else if (op == 2) m_open(c, shape, ksize_w, ksize_h, iterations, cur, other);
  else m_close(c, shape, ksize_w, ksize_h, iterations, cur, other);
  dim3 gp(grid_x(c.g.max_pixels), n);
  unpack_bits_kernel<<<gp, 256, 0, c.st>>>(cur, c.d_bi, d_desc, d_masks);
  return ipost("mask morphology kernels", c.launches + 2);
}

// op-level entry for the parity tests: labels (root pixel index, -1 background), area and order key per pixel's root
extern "C" int uwm_mask_components(const uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc, int n,
                                   int32_t* d_labels, int32_t* d_area, int32_t* d_order, int32_t* d_bbox, void* d_workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!d_masks || !d_desc || !d_labels || !d_area || !d_order || !d_workspace) return ifail(UWM_EINVAL, "mask_components: null pointer");
  int rc = check_descs(h_desc, n, "mask_components");
  if (rc) return rc;
  PostCtx c;
  rc = post_setup(c, h_desc, n, d_workspace, workspace_bytes, false, stream);
  if (rc) return rc;
  dim3 gw(grid_x(c.g.max_words * 32), n);
  pack_bits_kernel<<<gw, 256, 0, c.st>>>(d_masks, d_desc, c.d_bi, c.b0);
  c.labels = d_labels; c.area = d_area; c.order = d_order; c.bbox = d_bbox;
  run_ccl(c, c.b0, d_bbox != nullptr);
  return ipost("mask component kernels", c.launches + 1);
}

// d_out[i] = {foreground pixels, 8-connected components, area of the largest component} of mask i
// (cv2.connectedComponentsWithStats as reference src/scripts/model_selector.py:171-197 uses it)
extern "C" int uwm_mask_component_summary(const uint8_t* d_masks, const uwm_image_desc* h_desc, const uwm_image_desc* d_desc,
                                          int n, int32_t* d_out, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (!d_masks || !d_desc || !d_out || !d_workspace) return ifail(UWM_EINVAL, "mask_component_summary: null pointer");
  int rc = check_descs(h_desc, n, "mask_component_summary");
  if (rc) return rc;
  PostCtx c;
  rc = post_setup(c, h_desc, n, d_workspace, workspace_bytes, false, stream);
  if (rc) return rc;
  dim3 gw(grid_x(c.g.max_words * 32), n);
  pack_bits_kernel<<<gw, 256, 0, c.st>>>(d_masks, d_desc, c.d_bi, c.b0);
  run_ccl(c, c.b0, false);
  cudaMemsetAsync(d_out, 0, 12 * (size_t)n, c.st);
  dim3 gp(grid_x(c.g.max_pixels), n);
  ccl_summary_kernel<<<gp, 256, 0, c.st>>>(c.d_ci, c.labels, c.area, d_out);
  return ipost("mask component summary kernels", c.launches + 2);
}
