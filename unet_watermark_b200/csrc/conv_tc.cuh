// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[pixel, co] = epilogue( sum_{tap, ci} act[pixel*stride + tap, ci] * wgt[co, tap, ci] )
//
// GEMM view: M = output pixels (128 per CTA, a tw x th x tn box of the NHW space),
//            N = output channels (block_n <= 128 per CTA), K = taps * cin.
// One k-step = one filter tap x `kc` input channels (kc = 16/32/64 = one swizzle span):
//   A: TMA 4-D tiled load of the (tn,th,tw,kc) activation box shifted by the tap offset;
//      out-of-bounds coordinates are zero-filled by TMA, which IS the conv padding;
//      stride-2 convs use the tensor map's elementStrides.
//   B: TMA 2-D load of the (block_n, kc) slice of the [cout_pad][taps*cin] weight matrix.
// Both land in shared memory in the canonical K-major swizzled UMMA layout, are multiplied by
// tcgen05.mma (M=128, N=block_n, K=16 per instruction) into a TMEM fp32 accumulator, and the
// epilogue warps read TMEM with tcgen05.ld, add the folded-BN bias (+ residual), apply ReLU,
// round ONCE to bf16 and store NHWC.  Head mode writes fp32 logits / uint8 mask instead.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#pragma once
#include "ptx_sm100.cuh"

namespace uwm {

constexpr int kConvThreads = 192;
constexpr int kTileM = 128;
constexpr int kMaxTaps = 16;

struct ConvKArgs {
  // output geometry and tiling
  int n_img, h_out, w_out;
  int tw, th, tn;               // tile extents in output pixels, tw*th*tn == 128
  int tiles_w, tiles_h, tiles_n;
  int stride;                   // 1 or 2 (input coordinate = output coordinate * stride + tap)
  // K loop
  int ntaps, chunks, kc;        // k-steps = ntaps * chunks, each kc channels wide
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  // N
  int block_n, n_tiles, cout;   // cout = channels actually stored
  // pipeline
  int stages;
  uint32_t a_stage_bytes, b_stage_bytes;   // both multiples of 1024
  uint32_t tmem_cols;
  uint32_t layout_type;         // kLayoutSw128 / Sw64 / Sw32
  // epilogue
  const float* bias;            // [n_tiles*block_n]
  const __nv_bfloat16* res;     // optional residual [pixels][res_pitch]
  __nv_bfloat16* out;           // [pixels][out_pitch]
  long long res_pitch, out_pitch;
  int relu;
  // head mode (cout == 1): fp32 logits and/or uint8 mask per pixel
  int head, apply_sigmoid;
  float* logits;
  uint8_t* mask;
  float thr_logit;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__global__ void __launch_bounds__(kConvThreads)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act,
               const __grid_constant__ CUtensorMap tm_wgt,
               const __grid_constant__ ConvKArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atom (8 rows x 128 B).
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + p.stages * p.a_stage_bytes;
  const uint32_t misc_off = p.stages * (p.a_stage_bytes + p.b_stage_bytes);
  float* s_bias = reinterpret_cast<float*>(smem_gen + misc_off);                 // 256 floats
  const uint32_t bar_base = smem_base + misc_off + 1024;                          // 8-byte barriers
  uint32_t* s_tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + misc_off + 1024 + 8 * 40);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  const uint32_t acc_bar = bar_base + 8u * 32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int n_tile = blockIdx.x % p.n_tiles;
  int mt = blockIdx.x / p.n_tiles;
  const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
  const int tile_h = mt % p.tiles_h;
  const int tile_n = mt / p.tiles_h;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.tn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_wgt);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(s_tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.block_n; i += 128) s_bias[i] = p.bias[n_tile * p.block_n + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *s_tmem_slot;

  const int nk = p.ntaps * p.chunks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = 128u * p.kc * 2u + (uint32_t)p.block_n * p.kc * 2u;
      int s = 0; uint32_t ph = 0;
      int tap = 0, ch = 0;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), tx);
        tma_load_4d(a_base + s * p.a_stage_bytes, &tm_act, full_bar(s),
                    ch * p.kc, w0 * p.stride + p.tap_dw[tap], h0 * p.stride + p.tap_dh[tap], n0);
        tma_load_2d(b_base + s * p.b_stage_bytes, &tm_wgt, full_bar(s),
                    ks * p.kc, n_tile * p.block_n);
        if (++ch == p.chunks) { ch = 0; ++tap; }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTileM, p.block_n);
      const uint32_t row_bytes = p.kc * 2;
      const int kmma = p.kc / 16;
      int s = 0; uint32_t ph = 0;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t da = make_smem_desc(a_base + s * p.a_stage_bytes, row_bytes, p.layout_type);
        const uint64_t db = make_smem_desc(b_base + s * p.b_stage_bytes, row_bytes, p.layout_type);
        for (int k = 0; k < kmma; ++k) {
          // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in 16-byte units
          umma_bf16(tmem_acc, da + 2u * k, db + 2u * k, idesc, (ks | k) != 0);
        }
        umma_commit(empty_bar(s));           // frees the smem stage when these MMAs finish
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      umma_commit(acc_bar);                  // accumulator complete
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;           // accumulator row == pixel index inside the tile
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const int ow = w0 + wi, oh = h0 + hi, on = n0 + ni;
    const bool valid = (ow < p.w_out) && (oh < p.h_out) && (on < p.n_img);
    const long long pix = ((long long)on * p.h_out + oh) * p.w_out + ow;

    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);

    if (p.head) {
      uint32_t v[16];
      tmem_ld_x16(taddr, v);
      tmem_ld_wait();
      if (valid) {
        const float z = __uint_as_float(v[0]) + s_bias[0];
        if (p.logits) p.logits[pix] = p.apply_sigmoid ? 1.f / (1.f + __expf(-z)) : z;
        if (p.mask) p.mask[pix] = (z > p.thr_logit) ? 255 : 0;
      }
    } else {
      __nv_bfloat16* orow = p.out + pix * p.out_pitch + (long long)n_tile * p.block_n;
      const __nv_bfloat16* rrow =
          p.res ? p.res + pix * p.res_pitch + (long long)n_tile * p.block_n : nullptr;
      const int ncols = min(p.block_n, p.cout - n_tile * p.block_n);
      for (int c = 0; c < ncols; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(taddr + c, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c + j];
          if (rrow) {
            const uint4 r0 = *reinterpret_cast<const uint4*>(rrow + c);
            const uint4 r1 = *reinterpret_cast<const uint4*>(rrow + c + 8);
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) { f[2 * j] += bf16_lo(rr[j]); f[2 * j + 1] += bf16_hi(rr[j]); }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
          o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
          o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          *reinterpret_cast<uint4*>(orow + c) = o0;
          *reinterpret_cast<uint4*>(orow + c + 8) = o1;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, p.tmem_cols);
}

}  // namespace uwm
