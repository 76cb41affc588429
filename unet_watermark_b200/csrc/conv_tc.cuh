// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a) — persistent, warp-specialised.
//
//   out[pixel, co] = epilogue( sum_{tap, ci} act[pixel*stride + tap, ci] * wgt[co, tap, ci] )
//
// GEMM view: M = output pixels (128 per tile, a tw x th x tn box of the NHW space),
//            N = output channels (block_n <= 256 per tile), K = taps * cin.
// One k-step = one filter tap x `kc` input channels (kc = 16/32/64 = one swizzle span):
//   A: TMA 4-D tiled load of the (tn,th,tw,kc) activation box shifted by the tap offset;
//      out-of-bounds coordinates are zero-filled by TMA, which IS the conv padding;
//      stride-2 convs use the tensor map's elementStrides.
//   B: TMA 2-D load of the (block_n, kc) slice of the [cout_pad][taps*cin] weight matrix — either
//      per k-step through the ring, or (small layers) all slices once per CTA, resident in smem.
// Both land in shared memory in the canonical K-major swizzled UMMA layout and are multiplied by
// tcgen05.mma (M=128, N=block_n, K=16 per instruction) into a TMEM fp32 accumulator.
//
// Persistent schedule: grid = min(#tiles, #SMs); CTA b processes tiles b, b+G, b+2G, ...
//   warp 0  : TMA producer — runs ahead across tile boundaries through an S-stage smem ring
//   warp 1  : TMEM owner + MMA issuer — alternates between two TMEM accumulators
//   warps 2-5: epilogue — tcgen05.ld the finished accumulator (lane quarter = warp_id % 4), add the
//             folded-BN bias (+ residual), ReLU, round ONCE to bf16, store NHWC; or head mode
//             (fp32 logits / sigmoid / uint8 mask).  Overlaps the next tile's loads and MMAs.
#pragma once
#include "ptx_sm100.cuh"

// Pipeline-isolation switches and the in-kernel trace exist only in the tools build (-DUWM_BENCH_TOOLS); in the product
// library these fold to constants and the compiler drops the code behind them.
#ifdef UWM_BENCH_TOOLS
#define UWM_DBG_OF(p) ((p).dbg)
#define UWM_TRACE_OF(p) ((p).trace)
#else
#define UWM_DBG_OF(p) 0
#define UWM_TRACE_OF(p) (static_cast<long long*>(nullptr))
#endif

namespace uwm {

constexpr int kConvThreads = 192;
constexpr int kTileM = 128;
constexpr int kMaxTaps = 16;
constexpr int kMaxStages = 32;

struct ConvKArgs {
  // output geometry and tiling
  int n_img, h_out, w_out;
  int tw, th, tn;               // tile extents in output pixels, tw*th*tn == 128
  int tiles_w, tiles_h, tiles_n;
  int total_tiles;              // tiles_w*tiles_h*tiles_n*n_tiles
  int stride;                   // 1 or 2 (input coordinate = output coordinate * stride + tap)
  // K loop
  int ntaps, chunks, kc;        // k-steps = ntaps * chunks, each kc channels wide
  int8_t tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  // N
  int block_n, n_tiles, cout;   // cout = channels actually stored
  // pipeline
  int stages;
  int kpack;                    // k-steps per ring stage (one barrier round trip per kpack k-steps)
  int b_resident;               // 1: all weight slices loaded once per CTA, ring carries A only
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
  int dbg;                      // bench-only: 1 = skip TMA loads, 2 = skip MMAs (results are garbage)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

struct TileCoord { int n_tile, w0, h0, n0; };
__device__ __forceinline__ TileCoord decode_tile(const ConvKArgs& p, int tile) {
  TileCoord t;
  t.n_tile = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  t.w0 = (mt % p.tiles_w) * p.tw; mt /= p.tiles_w;
  t.h0 = (mt % p.tiles_h) * p.th;
  t.n0 = (mt / p.tiles_h) * p.tn;
  return t;
}

template <int KC>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act,
               const __grid_constant__ CUtensorMap tm_wgt,
               const __grid_constant__ ConvKArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atom (8 rows x 128 B).
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int nk = p.ntaps * p.chunks;
  const uint32_t ring_stage = (uint32_t)p.kpack * (p.a_stage_bytes + (p.b_resident ? 0u : p.b_stage_bytes));
  const uint32_t ring_base = smem_base;
  const uint32_t bres_base = ring_base + p.stages * ring_stage;                 // resident weights
  const uint32_t misc_off = p.stages * ring_stage + (p.b_resident ? nk * p.b_stage_bytes : 0u);
  const uint32_t bar_base = smem_base + misc_off;                               // 8-byte barriers
  uint32_t* s_tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + misc_off + 8 * (2 * kMaxStages + 8));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto accf_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };       // accumulator full
  auto acce_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };   // accumulator drained
  const uint32_t bres_bar = bar_base + 8u * (2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accf_bar(a), 1); mbar_init(acce_bar(a), 4); }
    mbar_init(bres_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_act);
    tma_prefetch_desc(&tm_wgt);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(s_tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_slot;
  pdl_wait();                                // everything below reads / writes activations of the previous kernel

  const int G = gridDim.x;

  // Producer and MMA warps run their loops warp-uniformly (all 32 lanes wait on the barriers, one
  // elected lane issues): addresses, coordinates and descriptors then live in uniform registers.
  // A ring stage carries `kpack` k-steps so that one barrier round trip (~100 cycles of try_wait plus
  // the loop bookkeeping of the single issuing thread) is amortised over >= ~500 cycles of MMA work.
  const int kpack = p.kpack;
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    constexpr uint32_t kABytes = 128u * KC * 2u;
    const uint32_t b_bytes = (uint32_t)p.block_n * KC * 2u;
    if (p.b_resident) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bres_bar, (uint32_t)nk * b_bytes);
        for (int ks = 0; ks < nk; ++ks)
          tma_load_2d(bres_base + ks * p.b_stage_bytes, &tm_wgt, bres_bar, ks * KC, 0);
      }
      __syncwarp();
    }
    const uint32_t tx1 = kABytes + (p.b_resident ? 0u : b_bytes);
    const bool skip_tma = (UWM_DBG_OF(p) == 1 || UWM_DBG_OF(p) == 7);
    const uint32_t b_off0 = (uint32_t)kpack * p.a_stage_bytes;
    const bool leader = elect_one();
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += G) {
      const TileCoord t = decode_tile(p, tile);
      const int wi0 = t.w0 * p.stride, hi0 = t.h0 * p.stride;
      const int ncol = t.n_tile * p.block_n;
      int tap = 0, ch = 0;
      int cw = wi0 + p.tap_dw[0], chh = hi0 + p.tap_dh[0];
      for (int ks0 = 0; ks0 < nk; ks0 += kpack) {
        const int cnt = min(kpack, nk - ks0);
        const uint32_t dst = ring_base + s * ring_stage;
        if (leader) {      // one lane waits and issues: mbarrier ops cost ~2x more when all 32 lanes execute them
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (skip_tma) mbar_arrive(full_bar(s));
          else mbar_arrive_expect_tx(full_bar(s), (uint32_t)cnt * tx1);
        }
        for (int j = 0; j < cnt; ++j) {
          if (leader && !skip_tma) {
            tma_load_4d(dst + j * p.a_stage_bytes, &tm_act, full_bar(s), ch * KC, cw, chh, t.n0);
            if (!p.b_resident)
              tma_load_2d(dst + b_off0 + j * p.b_stage_bytes, &tm_wgt, full_bar(s), (ks0 + j) * KC, ncol);
          }
          if (++ch == p.chunks) {
            ch = 0;
            if (++tap < p.ntaps) { cw = wi0 + p.tap_dw[tap]; chh = hi0 + p.tap_dh[tap]; }
          }
        }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(kTileM, p.block_n);
    // shared-memory matrix descriptor, split in halves: hi = SBO | version | swizzle (invariant),
    // lo = start address >> 4 | LBO  (advances with the stage and with K inside the swizzle span)
    constexpr uint32_t kRowBytes = KC * 2;
    const uint32_t desc_hi = ((8u * kRowBytes) >> 4) | (1u << 14) | (p.layout_type << 29);
    const uint32_t a_lo0 = ((ring_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t bres_lo0 = ((bres_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stage_units = ring_stage >> 4, a_units = p.a_stage_bytes >> 4, b_units = p.b_stage_bytes >> 4;
    const uint32_t b_off_units = (uint32_t)kpack * a_units;
    const bool skip_mma = (UWM_DBG_OF(p) == 2 || UWM_DBG_OF(p) == 7);
    const bool leader = elect_one();
    if (p.b_resident && leader) mbar_wait(bres_bar, 0);
    int s = 0; uint32_t ph = 0;
    int it = 0;
    bool ready = false;                  // leader-only state: result of peeking at the next stage
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += G, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * p.block_n);
      if (leader) {
        mbar_wait(acce_bar(acc), (use & 1u) ^ 1u);     // epilogue of the previous use has drained it
        tc_fence_after();
      }
      for (int ks0 = 0; ks0 < nk; ks0 += kpack) {
        const int cnt = min(kpack, nk - ks0);
        const uint32_t a_lo = a_lo0 + s * stage_units;
        const uint32_t b_lo = p.b_resident ? bres_lo0 + ks0 * b_units : a_lo + b_off_units;
        const int sn = (s + 1 == p.stages) ? 0 : s + 1;
        const uint32_t phn = (s + 1 == p.stages) ? ph ^ 1u : ph;
        if (leader) {
          if (!ready) mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (skip_mma) {
            mbar_arrive(empty_bar(s));
          } else {
            for (int j = 0; j < cnt; ++j) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)   // +32 bytes (2 x 16-byte units) per K=16 slice
                umma_bf16_lohi(tmem_acc, a_lo + j * a_units + 2u * k, b_lo + j * b_units + 2u * k, desc_hi, idesc,
                               (k > 0) ? 1u : (uint32_t)((ks0 + j) != 0));
            }
            umma_commit(empty_bar(s));           // frees the smem stage when these MMAs finish
          }
          // peek at the next stage (non-blocking): usually already landed, which saves the blocking wait
          ready = mbar_test_wait(full_bar(sn), phn);
        }
        s = sn; ph = phn;
      }
      if (leader) umma_commit(accf_bar(acc));   // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;           // accumulator row == pixel index inside the tile
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += G, ++it) {
      const int acc = it & 1;
      const TileCoord t = decode_tile(p, tile);
      const int ow = t.w0 + wi, oh = t.h0 + hi, on = t.n0 + ni;
      const bool valid = (ow < p.w_out) && (oh < p.h_out) && (on < p.n_img);
      const long long pix = ((long long)on * p.h_out + oh) * p.w_out + ow;
      const int col0 = t.n_tile * p.block_n;

      if (lane == 0) mbar_wait(accf_bar(acc), (uint32_t)(it >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);

      if (UWM_DBG_OF(p) == 4 || UWM_DBG_OF(p) == 7) {
        // bench-only: no epilogue work at all
      } else if (p.head) {
        uint32_t v[16];
        tmem_ld_x16(taddr, v);
        tmem_ld_wait();
        if (valid) {
          const float z = __uint_as_float(v[0]) + __ldg(p.bias);
          if (p.logits) p.logits[pix] = p.apply_sigmoid ? 1.f / (1.f + __expf(-z)) : z;
          if (p.mask) p.mask[pix] = (z > p.thr_logit) ? 255 : 0;
        }
      } else {
        __nv_bfloat16* orow = p.out + pix * p.out_pitch + col0;
        const __nv_bfloat16* rrow = p.res ? p.res + pix * p.res_pitch + col0 : nullptr;
        const float4* brow = reinterpret_cast<const float4*>(p.bias + col0);
        const int ncols = min(p.block_n, p.cout - col0);
        for (int c = 0; c < ncols; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + c, v);
          // bias is warp-uniform: 4 broadcast 16-byte loads, issued while the TMEM load is in flight
          const float4 b0 = __ldg(brow + (c >> 2)), b1 = __ldg(brow + (c >> 2) + 1);
          const float4 b2 = __ldg(brow + (c >> 2) + 2), b3 = __ldg(brow + (c >> 2) + 3);
          uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
          if (rrow && valid) {
            r0 = *reinterpret_cast<const uint4*>(rrow + c);
            r1 = *reinterpret_cast<const uint4*>(rrow + c + 8);
          }
          tmem_ld_wait();
          if (valid) {
            float f[16];
            const float bb[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w,
                                  b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bb[j];
            if (rrow) {
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
            if (UWM_DBG_OF(p) != 3) {
              *reinterpret_cast<uint4*>(orow + c) = o0;
              *reinterpret_cast<uint4*>(orow + c + 8) = o1;
            } else if (o0.x == 0x12345678u && o1.w == 0x9abcdef0u) {   // bench-only: keep the math alive
              *reinterpret_cast<uint4*>(orow + c) = o0;
            }
          }
        }
      }
      // all tcgen05.ld of this accumulator have completed (wait::ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace uwm
