// Halo-resident implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): every conv of the Unet plan.
// Stride-1 k x k convs (3x3, the stem's 4x4 over the space-to-depth input, 1x1) with the decoder's nearest-2x
// upsample and skip concat fused into the activation loader, plus the forms documented at the kernel template
// below: sub-pixel convs over an upsampled (and concatenated) input (SPX = 1), stride-2 convs over the input's
// parity planes (SPX = 2), convs on space-to-depth tensors (S2D) and CTA pairs sharing the weights (CG2).
//
//   out[n,h,w,co] = epilogue( sum_{tap,ci} in[n, h+dh(tap), w+dw(tap), ci] * wgt[co, tap, ci] )
//   in = concat_c( up2x?(src0), src1 )          (src1 optional; up2x = F.interpolate(nearest, x2))
//
// Why a second conv kernel: conv_tc.cuh re-fetches the 128-pixel activation tile from L2 once per
// filter tap (9x for a 3x3), and the L2->SM fabric (~42 B/clk/SM) - not the tensor pipe - bounds every
// layer.  Here a CTA loads the HALO of its output tile ONCE per 64-channel chunk and every tap reads a
// shifted window of that one copy.  A tile is TG sub-tiles of 8 (w) x 16 (h) pixels side by side; each
// sub-tile is one M=128 accumulator, all of them share the halo and every weight slice.
//
// A operand, two layouts (template A_TMA):
//   A_TMA (all sources at the conv's resolution): ONE TMA box per stage - KC channels x halo_w x halo_h
//     pixels, out-of-image coordinates zero-filled (= conv padding) - in the 128B/64B/32B-swizzled
//     pixel-major layout (row = pixel, KC*2 bytes).  Tap (r,c), sub-tile g: descriptor start address
//     += (r*halo_w + c + 8g) rows; SBO = halo_w rows.  A start that is not aligned to the 8-row swizzle
//     atom still reads what TMA wrote: both units derive the XOR from the absolute smem address.
//   !A_TMA (an upsampled source): loader warps gather 16-byte pieces with cp.async (zero-fill = padding;
//     source pixel (h>>1, w>>1) for the upsampled source) into a no-swizzle K-major layout
//     [kc/8 channel groups][halo pixels][16 B]: SBO = halo_w*16 B, LBO = plane stride, start address
//     += (r*halo_w + c + 8g)*16 B; completion -> mbarrier through cp.async.mbarrier.arrive.  Neither the
//     upsampled tensor nor the concat ever exists in HBM.
// B operand (weights): TMA into 128B/64B/32B-swizzled K-major slices, resident in smem for the small layers
// or streamed through their own ring.
//
// Filter shape, TG and the weight mode are template parameters: ncu and the in-kernel trace show a warp
// running ~4-10 cycles per instruction in these role loops, so the single MMA-issuing thread (9 MMAs of 32
// cycles per 128 pixels in the smallest layers) and the epilogue are instruction-count bound unless their
// loops are fully unrolled with immediate offsets; the epilogue's item loop is instantiated per output mode.
//
// Warp roles (448 threads, 1 CTA/SM, persistent over tiles):
//   warp 0      weight TMA producer          warp 1    TMEM owner + tcgen05.mma issuer
//   warps 2-5   epilogue set 0 (TMEM -> bias/residual/ReLU -> bf16 NHWC through a swizzled smem staging row and
//               a TMA tensor store, or direct stores for 16/32 channels, or head: logits / uint8 mask)
//   warps 6-9   activation loaders (one elected thread when A_TMA)
//   warps 10-13 epilogue set 1: a warp may only read TMEM lanes 32*(warp%4).., and one warp per lane quarter
//               is latency-bound (one TMEM load -> math -> store chain at a time), so every quarter gets two
//               warps that split the sub-tiles (or the 16-column chunks when TG == 1) of each tile
#pragma once
#include <type_traits>

#include "conv_tc.cuh"

namespace uwm {

constexpr int kHaloThreads = 448;
constexpr int kHaloLoaderThreads = 128;
constexpr int kHaloLoaderWarp0 = 6;
constexpr int kHaloTW = 8, kHaloTH = 16;
constexpr int kHaloMaxStages = 16;

// geometry shared by host and device
__host__ __device__ constexpr int halo_pw(int tg, int kw) { return kHaloTW * tg + kw - 1; }
__host__ __device__ constexpr int halo_npix(int tg, int kh, int kw) { return halo_pw(tg, kw) * (kHaloTH + kh - 1); }
// plane stride = 4 (mod 8) sixteen-byte units: the 8 channel groups of one pixel land in distinct banks
__host__ __device__ constexpr uint32_t halo_plane_bytes(int tg, int kh, int kw) {
  return (((uint32_t)halo_npix(tg, kh, kw) * 16u + 127u) & ~127u) + 64u;
}

// n / d for n < 2^31 with a precomputed multiplier (host: make_fastdiv)
struct FastDiv { uint32_t mul, shr, d; };
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}

struct HaloSrc {
  const __nv_bfloat16* ptr;   // NHWC, already offset to the first channel this conv reads
  long long pitch;            // elements between consecutive pixels
  int h, w;                   // spatial size of the stored tensor (half the conv's size when up == 1)
  int up;                     // 1: nearest-2x upsample on the fly
};

struct HaloKArgs {
  // geometry: stride 1, output size == (virtual) input size
  int n_img, h, w;
  int tiles_w, tiles_h, n_tiles, total_tiles;
  FastDiv div_ntiles, div_tw, div_th;
  FastDiv div_total;                        // CHAIN: work item / total_tiles = layer
  // K loop: chunks of kc channels (one A stage each), all taps per chunk
  int chunks;
  int split_chunk;                          // chunks [0,split) come from src[0], the rest from src[1]
  int dh_min, dw_min;                       // offset of tap (0,0): tap (r,c) reads pixel (h+dh_min+r, w+dw_min+c)
  uint32_t pw_magic;                        // p / pw == (p * pw_magic) >> 16 for p < npix
  int a_stages;
  HaloSrc src[2];
  // weights
  int cin_total;                            // K index of (tap, chunk) = tap*cin_total + chunk*kc
  int block_n, cout;
  int kpb, b_stages;                        // streamed weights: kpb taps per stage (divides the tap count)
  uint32_t b_slice_bytes;                   // one (tap, chunk) slice, multiple of 1024
  uint32_t tmem_cols;
  int nacc_log2;                            // log2 of the accumulator sets in TMEM (1 or 2)
  // epilogue
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* out;
  long long res_pitch, out_pitch;
  int relu;
  int head, apply_sigmoid;
  float* logits;
  uint8_t* mask;
  float thr_logit;
  int a_scale;                              // A_TMA loaders: box start = a_scale * tile origin (2: the map walks every second pixel)
  int spx_cpp, spx_slices;                  // SPX kernels: 64-channel chunks per parity plane of src[1]; weight slices per tile
  int mix;                                  // !A_TMA kernels: chunks of src[1] arrive as TMA boxes (swizzled stage layout)
  int ep_tma, ep_cols;                      // epilogue: TMA tensor stores of (64 ch x 8 x 4 px) boxes via smem staging (ep_cols 64), else 16
  int shuffle;                              // > 0: sub-pixel mode, real cout; GEMM column n = parity*shuffle + co is stored to
                                            // pixel (2h + parity/2, 2w + parity%2), channel co of the 2x larger output (pixel shuffle)
  unsigned long long a_policy[2];           // A_TMA loaders: L2 eviction policy of the loads from src[0] / src[1] (0: default);
                                            // evict_first for a source nobody reads after this launch (build_plan)
  int reverse;                              // walk the tiles from the last to the first (serpentine order across launches, see build_plan)
  long long* trace;                         // bench-only: CTA 0 writes clock64() stamps of its pipeline events (tools/gpu_trace.py)
  int dbg;                                  // bench-only bit mask: 1 skip activation loads, 2 skip MMAs, 4 skip epilogue
};

// ---- multi-layer chains (CHAIN): one persistent launch runs several same-shaped conv layers back to back ----------
// The N = 128 streamed 3x3 layers of a ResNet stage have 128-256 tiles each on 148 SMs and spend a third of a
// launch in fill, drain and the gap to the next launch.  A chain walks the work items (layer, tile) of up to
// kMaxChainLayers consecutive layers of identical geometry in one grid: item i = layer * total_tiles + tile, CTA c
// takes items c, c + G, ...  Layer l + 1 of image n may start as soon as every tile of image n of layer l has been
// stored: the epilogue warps count their completed TMA stores into dep[l][n] (release), the activation loader
// polls it (acquire) before it requests the halo of a tile of that image.  Accumulator double-buffering, the weight
// ring and the activation ring run across layer boundaries, so only the first fill and the last drain are exposed
// and the tiles of consecutive layers fill every SM.  Tensors touched inside a chain never alias each other
// (build_plan extends their lifetimes over the chain), and a residual (layer l - 2's output of the same image) is
// complete by transitivity.
constexpr int kMaxChainLayers = 16;
struct alignas(64) HaloLayerRef {
  CUtensorMap tm_wgt, tm_out, tm_res, tm_a0;
  const float* bias;
  int relu, has_res;
  int res_layer;                            // chain layer whose output is this layer's residual, -1: produced before the chain
  int pad_[11];
};
static_assert(sizeof(HaloLayerRef) % 64 == 0, "HaloLayerRef must stay a multiple of 64 bytes (tensor maps are 64-byte aligned)");
struct HaloChain {
  HaloLayerRef layer[kMaxChainLayers];
  int n_layers;
  int dep_target;                           // arrivals per (layer, image): tiles of one image x 8 epilogue warps
  int* dep;                                 // [n_layers][n_img] counters, then one CTA ticket (all zero between launches)
};
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// all bulk-group operations of this thread have COMPLETED (writes performed), not only read their source
template <int N>
__device__ __forceinline__ void bulk_wait_done() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// orders generic-proxy accesses (the counters) against async-proxy accesses (TMA loads / stores) of this thread
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __noinline__ void chain_wait_dep(const int* ctr, int target) {
  long long t0 = 0;
  uint32_t spins = 0;
  while (ld_acquire_gpu(ctr) < target) {
    __nanosleep(40);
    if ((++spins & 4095u) != 0) continue;
    if (t0 == 0) { t0 = clock64(); continue; }
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a dependency that can never complete is a protocol bug, not a wait
      printf("uwm: chain dependency watchdog (block %d, counter %d < %d)\n", (int)blockIdx.x, ld_acquire_gpu(ctr), target);
      __trap();
    }
  }
}

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` (without raising its pending count) once all cp.async issued so far by this thread landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Waits with the watchdog kept out of line: the MMA thread's instruction footprint matters (i-cache).
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// tcgen05.mma with separate descriptor halves; ACC: accumulate into D unconditionally, else only when acc != 0
template <bool ACC, bool CG2 = false>
__device__ __forceinline__ void umma_halo(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t acc) {
  if (CG2) {
    umma_bf16_lohi2_cg2(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, ACC ? 1u : acc);
  } else if (ACC) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  } else {
    umma_bf16_lohi2(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
  }
}

// trace slots (CTA 0 only): [0] kernel start, [1] prologue done, loaders 16+2*stage+{0 slot free, 1 copies issued},
// MMA 160+3*tile+{0 accumulator free, 1 first stage landed, 2 all MMAs issued}, epilogue 400+2*tile+{0 accumulator
// full, 1 tile stored}
__device__ __forceinline__ void halo_trace(const HaloKArgs& p, int slot) {
  if (UWM_TRACE_OF(p) && blockIdx.x == 0 && slot < 1000) p.trace[slot] = clock64();
}

// dbg bit 8 (with a trace buffer of >= 1024 + 4*grid slots): every CTA stamps its start and exit in nanoseconds
// and in SM cycles,
// so consecutive launches show the gap between one grid's last exit and the next grid's first tile
__device__ __forceinline__ void halo_trace_cta(const HaloKArgs& p, int which) {
  if (UWM_TRACE_OF(p) && (UWM_DBG_OF(p) & 8)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[1024 + 4 * blockIdx.x + which] = (long long)t;
    p.trace[1024 + 4 * blockIdx.x + 2 + which] = clock64();
  }
}

// S2D: parity plane k = (ph,pw) of a space-to-depth tensor meets tap (r,c) of the 3x3 block neighbourhood only for
// r in {1-ph, 2-ph} and c in {1-pw, 2-pw}: 16 of the 36 (tap, plane) pairs.  s2d_pair = position of a meeting pair in
// (tap, plane) order = index of its [N x 16] weight block in shared memory.
__host__ __device__ constexpr bool s2d_meets(int tap, int k) {
  return (tap / 3 == 1 - (k >> 1) || tap / 3 == 2 - (k >> 1)) && (tap % 3 == 1 - (k & 1) || tap % 3 == 2 - (k & 1));
}
__host__ __device__ constexpr int s2d_pair(int tap, int k) {
  // meeting pairs per tap: 1 2 1 / 2 4 2 / 1 2 1
  int j = tap == 0 ? 0 : tap == 1 ? 1 : tap == 2 ? 3 : tap == 3 ? 4 : tap == 4 ? 6 : tap == 5 ? 10 : tap == 6 ? 12 : tap == 7 ? 13 : 15;
  for (int kk = 0; kk < k; ++kk) if (s2d_meets(tap, kk)) ++j;
  return j;
}
static_assert(s2d_pair(8, 0) == 15 && s2d_pair(4, 3) == 9 && s2d_pair(5, 2) == 11, "S2D weight block order");

// max of two packed bf16 pairs (one HMNMX2)
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v, uint32_t clamp) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(clamp));
  return r;
}

struct HaloTile { int n_tile, w0, h0, img; };
template <int TG>
__device__ __forceinline__ HaloTile halo_decode(const HaloKArgs& p, int tile) {
  HaloTile t;
  int mt = fast_div(tile, p.div_ntiles);
  t.n_tile = tile - mt * p.n_tiles;
  int q = fast_div(mt, p.div_tw);
  t.w0 = (mt - q * p.tiles_w) * (kHaloTW * TG);
  t.img = fast_div(q, p.div_th);
  t.h0 = (q - t.img * p.tiles_h) * kHaloTH;
  return t;
}

// SPX (sub-pixel conv with a skip source): conv3x3(concat(nearest-2x(x), skip)) computed on x's grid with one GEMM
// column group per output parity (the epilogue's pixel shuffle scatters them).  x chunks are ordinary 3x3 chunks with
// pre-summed weights; the skip source, stored at the OUTPUT resolution, arrives as four parity planes (TMA boxes with
// element stride 2) of 64-channel chunks, and plane (ph,pw) only meets the 2x2 taps {1-ph,2-ph} x {1-pw,2-pw} of the
// 3x3 block neighbourhood: 16 (plane, tap) pairs instead of the 36 a dense 3x3 conv over the space-to-depth input
// would issue.  Weight slices stream in issue order (p.spx_slices per tile, one per stage); in SPX kernels tm_out and
// tm_res are the weight maps with half- and quarter-height boxes (the epilogue stores without TMA).
//
// SPX == 2 (stride-2 3x3 conv, pad 1) uses the same machinery with every chunk a parity plane of the one source:
// output pixel (i,j) reads input rows 2i-1, 2i, 2i+1 = (block i-1, plane 1), (block i, plane 0), (block i, plane 1), so
// with a 2x2 block halo at origin -1 plane ph meets halo rows {1} (ph = 0) or {0,1} (ph = 1): 9 (plane, tap) pairs,
// one per kernel tap, each a full-N MMA group over a dense TMA-written stage - the stride never reaches the MMA.
//
// SPX == 3 (sub-pixel conv over an upsampled input with 4*Cout > 256 GEMM columns): one N tile per output parity
// (block_n = Cout); parity (qh,qw) only meets the 2x2 taps {qh,qh+1} x {qw,qw+1} of the 3x3 source neighbourhood, so
// an N tile streams and issues 4 of the 9 taps.  The pixel-shuffle epilogue can add a residual (the partial sum of
// the skip half of a decoder conv1, computed by an ordinary conv launch) before the ReLU; with Cout >= 64 it is the
// ordinary TMA-store epilogue over an element-stride-2 view of the output (one view serves the four parities).
//
// CG2 (CTA pair, streamed weights): the two CTAs of a 2-CTA cluster work on two M tiles of the same N tile and share
// every weight slice - each loads HALF of its rows (block_n/2) and the leader issues one M = 256 tcgen05.mma.cta_group::2
// per tap / K step that reads A from both CTAs' stages (same offsets: the rings run in lockstep), B half from each, and
// writes each CTA's 128 x N accumulator into its own TMEM.  Per CTA that halves the weight bytes through the L2->SM
// fabric and through the shared-memory write port (the two things that bound layer4 and slow the MMAs of layer2/3).
// Protocol: 'full' barriers and the accumulator-free barriers live in the leader: its producers arm them with the
// bytes of both CTAs and both CTAs' TMA loads complete on them; the accumulator-free barrier counts the 16 epilogue
// warps of the pair (the peer's arrive over the cluster address).  'empty' and accumulator-full barriers are local and
// receive the leader's multicast commits, so the two rings run in lockstep.
//
// S2D (conv3x3 on a tensor stored space-to-depth): a [n, 2h, 2w, 16] activation kept as [n, h, w, 4 x 16] (channel
// group = pixel parity (ph,pw), the layout the sub-pixel conv's GEMM produces before any pixel shuffle) is one
// 64-channel chunk whose K=16 slice k IS parity plane k.  A 3x3 conv at the full resolution is then a 3x3 conv over
// blocks with 4x16 outputs in which plane (ph,pw) only meets taps {1-ph,2-ph} x {1-pw,2-pw}: the MMA loop issues those
// 16 of the 36 (tap, k) pairs (N = 64 or 16 instead of 16 per MMA, 2.25x fewer MMAs per output pixel) and rows of
// 128 bytes go in and out by TMA.  The head variant writes the 4 logits / mask bytes of a block to its 2x2 pixels.
template <int KC, int KH, int KW, int TG, bool RESIDENT, bool A_TMA, int SPX = 0, bool S2D = false, bool CG2 = false,
          bool CHAIN = false>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_wgt, const __grid_constant__ CUtensorMap tm_out,
                 const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                 const __grid_constant__ HaloKArgs p, const HaloChain* __restrict__ chain) {
  static_assert(!CHAIN || (SPX == 0 && !S2D && !CG2 && !RESIDENT && A_TMA), "chains: plain streamed TMA-fed convs only");
  constexpr int NT = KH * KW;
  constexpr bool SPXP = (SPX == 1 || SPX == 2);            // chunk / slice sequencing of the parity-plane forms
  constexpr int CPS = KC / 8;                               // 8-channel planes per stage
  constexpr int LOG2_CPS = (KC == 64) ? 3 : (KC == 32) ? 2 : 1;
  constexpr int PW = halo_pw(TG, KW);
  constexpr int NPIX = halo_npix(TG, KH, KW);
  constexpr uint32_t PLANE_BYTES = halo_plane_bytes(TG, KH, KW);
  constexpr uint32_t PLANE_UNITS = PLANE_BYTES >> 4;
  // A_TMA: the halo is one TMA box (KC channels x halo_w x halo_h pixels) in the swizzled pixel-major layout
  // (row = pixel, KC*2 bytes); the tap / sub-tile shift is a whole number of rows.  TMA and the MMA unit both
  // derive the swizzle XOR from the absolute shared-memory address, so a start address that is not aligned to
  // the 8-row swizzle atom still reads what TMA wrote.
  constexpr uint32_t ROWB = KC * 2;
  constexpr uint32_t A_SW_BYTES = ((uint32_t)NPIX * ROWB + 1023u) & ~1023u;
  constexpr uint32_t A_PL_BYTES = (CPS * PLANE_BYTES + 1023u) & ~1023u;     // !A_TMA stages may hold either layout
  constexpr uint32_t A_STAGE_BYTES = A_TMA ? A_SW_BYTES : (A_SW_BYTES > A_PL_BYTES ? A_SW_BYTES : A_PL_BYTES);
  constexpr int ITEMS = NPIX * CPS;                         // 16-byte pieces per stage
  constexpr uint32_t B_LAYOUT = (KC == 64) ? kLayoutSw128 : (KC == 32) ? kLayoutSw64 : kLayoutSw32;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // weight slices need 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int nk = NT * p.chunks;
  const uint32_t b_stage_bytes = (uint32_t)p.kpb * p.b_slice_bytes;
  // S2D: only the 16 meeting (tap, plane) blocks of [block_n x 16] weights are kept (32-byte rows, 32B swizzle)
  const uint32_t b_bytes_total = S2D ? ((16u * (uint32_t)p.block_n * 32u + 1023u) & ~1023u)
                                 : RESIDENT ? (uint32_t)nk * p.b_slice_bytes : (uint32_t)p.b_stages * b_stage_bytes;
  const uint32_t b_base = smem_base;
  const uint32_t a_base = smem_base + b_bytes_total;
  const uint32_t stg_base = (a_base + (uint32_t)p.a_stages * A_STAGE_BYTES + 1023u) & ~1023u;   // epilogue staging
  const uint32_t stg_bytes = p.ep_tma ? 8u * 32u * 128u : 0u;
  const uint32_t bar_base = (stg_base + stg_bytes + 7u) & ~7u;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (kHaloMaxStages + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * kHaloMaxStages + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (3 * kHaloMaxStages + s); };
  auto accf_bar = [&](int a) { return bar_base + 8u * (4 * kHaloMaxStages + a); };
  auto acce_bar = [&](int a) { return bar_base + 8u * (4 * kHaloMaxStages + 4 + a); };
  const uint32_t bres_bar = bar_base + 8u * (4 * kHaloMaxStages + 8);
  auto resbar = [&](int w) { return bar_base + 8u * (4 * kHaloMaxStages + 12 + w); };   // residual box landed (per epilogue warp)
  uint32_t* s_tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * (4 * kHaloMaxStages + 10));
  // folded-BN bias of every output channel, staged once: the epilogue re-reads it per 16-column chunk, and the
  // trace showed that re-read missing L1 (an L2 round trip of 500-800 cycles in front of the first FADD)
  // CHAIN work queue: items are handed out by a global atomic counter (the weight producer grabs, the other roles
  // follow through this 4-deep FIFO) - see q_get below
  auto qfull_bar = [&](int i) { return bar_base + 8u * (4 * kHaloMaxStages + 20 + i); };
  auto qempty_bar = [&](int i) { return bar_base + 8u * (4 * kHaloMaxStages + 24 + i); };
  volatile int* s_items = reinterpret_cast<volatile int*>(smem_gen + (bar_base - smem_base) + 8 * (4 * kHaloMaxStages + 28));
  float* s_bias = reinterpret_cast<float*>(smem_gen + (bar_base - smem_base) + 8 * (4 * kHaloMaxStages + 36));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();
  if (threadIdx.x == 0) { halo_trace(p, 0); halo_trace_cta(p, 0); }
  // The prologue is on every launch's critical path (the CTA cannot start before the previous kernel's CTA has left
  // the SM): the ~45 barrier inits are spread over four threads of different warps, and the bias is staged by the
  // epilogue warps after the block-wide sync (they have thousands of cycles before the first accumulator).
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bfull_bar(s), 1); mbar_init(bempty_bar(s), 1); }
    mbar_init(bres_bar, 1);
    if (CHAIN) for (int i = 0; i < 4; ++i) { mbar_init(qfull_bar(i), 1); mbar_init(qempty_bar(i), 10); }   // consumers: weight producer, MMA, 8 epilogue warps
    fence_mbar_init();
    tma_prefetch_desc(&tm_wgt);
    if (SPX == 1) { tma_prefetch_desc(&tm_out); tma_prefetch_desc(&tm_res); }
  } else if (threadIdx.x == 64) {
    for (int a = 0; a < 4; ++a) { mbar_init(accf_bar(a), 1); mbar_init(acce_bar(a), CG2 ? 16 : 8); }
    for (int w = 0; w < 8; ++w) mbar_init(resbar(w), 1);
    fence_mbar_init();
    if (p.ep_tma) { tma_prefetch_desc(&tm_out); tma_prefetch_desc(&tm_res); }
  } else if (threadIdx.x == kHaloLoaderWarp0 * 32) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(afull_bar(s), A_TMA ? 1 : kHaloLoaderThreads); mbar_init(aempty_bar(s), 1); }
    fence_mbar_init();
    if (A_TMA) { tma_prefetch_desc(&tm_a0); tma_prefetch_desc(&tm_a1); }
    else if (p.mix) tma_prefetch_desc(&tm_a1);
  }
  const uint32_t crank = CG2 ? cluster_ctarank() : 0u;      // CTA pair: rank 0 leads (issues the MMAs, owns the full barriers)
  // the peer's barriers must exist before anything signals them: cluster barrier, its wait overlapped with the TMEM alloc
  if (CG2) cluster_arrive();
  if (warp == 1) {
    if (CG2) { tmem_alloc_cg2(smem_u32(s_tmem_slot), p.tmem_cols); tmem_relinquish_cg2(); }
    else { tmem_alloc(smem_u32(s_tmem_slot), p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CG2) cluster_wait();
  const uint32_t tmem_base = *s_tmem_slot;
  const int G = gridDim.x;
  // tile sequence of this CTA: tiles blockIdx.x, +G, ... ; a CTA pair walks pair-tiles (two M tiles of one N tile)
  const int t_first = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_layers = CHAIN ? chain->n_layers : 1;
  const int t_end = CHAIN ? n_layers * p.total_tiles : (CG2 ? (p.total_tiles >> 1) : p.total_tiles);
  const int t_step = CG2 ? (G >> 1) : G;
  // CHAIN: work item tl = layer * total_tiles + tile.  Items are not assigned statically: a global counter hands them
  // out in order (the activation loader grabs, about one item ahead of the MMAs; the weight producer, the MMA thread
  // and the epilogue warps read the same sequence from a 4-deep shared-memory FIFO).  With static striding every layer boundary
  // falls differently against the rounds of 148 CTAs, a quarter of the CTAs always depend on the round just
  // before theirs and throttle everyone (measured); with the counter every item's producers are the same number
  // of grabs back, whoever runs them.
  auto layer_of = [&](int tl) { return CHAIN ? fast_div(tl, p.div_total) : 0; };
  auto q_get = [&](int k) -> int {           // k-th item of this CTA, -1 after the last (called by ONE thread per consumer)
    mbar_wait(qfull_bar(k & 3), (uint32_t)(k >> 2) & 1u);
    const int v = s_items[k & 3];
    mbar_arrive(qempty_bar(k & 3));
    return v;
  };
  auto tile_of = [&](int tl) {
    if (CHAIN) return tl - fast_div(tl, p.div_total) * p.total_tiles;
    if (!CG2) return p.reverse ? p.total_tiles - 1 - tl : tl;
    const int mu = fast_div(tl, p.div_ntiles);
    return (2 * mu + (int)crank) * p.n_tiles + (tl - mu * p.n_tiles);
  };
  // the leader's copy of a barrier, as a cluster address (the leader's own barrier for the leader)
  auto lead = [&](uint32_t bar) { return CG2 ? mapa_u32(bar, 0) : bar; };
  // accumulator sets in TMEM (2 or 4): the MMA -> epilogue -> MMA hand-back is a ~3000-cycle round trip even
  // with nothing to do, so short tiles need more than two sets in flight
  const int nacc_log2 = p.nacc_log2, nacc_mask = (1 << nacc_log2) - 1;
  if (threadIdx.x == 0) halo_trace(p, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ weight TMA producer
    const uint32_t slice_tx = (uint32_t)p.block_n * KC * 2u;
    if (S2D) {
      // the weights arrive in the [N][9 x 64] layout of every 3x3 conv; tm_a1 (no second source in this form) views
      // them in [N x 16] boxes, and only the blocks a parity plane can meet are fetched
      if (elect_one()) {
        mbar_arrive_expect_tx(bres_bar, 16u * (uint32_t)p.block_n * 32u);
        for (int tap = 0; tap < 9; ++tap)
          for (int k = 0; k < 4; ++k)
            if (s2d_meets(tap, k))
              tma_load_2d(b_base + (uint32_t)s2d_pair(tap, k) * (uint32_t)p.block_n * 32u, &tm_a1, bres_bar, tap * 64 + 16 * k, 0);
      }
      __syncwarp();
    } else if (RESIDENT) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bres_bar, (uint32_t)nk * slice_tx);
        for (int ch = 0; ch < p.chunks; ++ch)
          for (int tap = 0; tap < NT; ++tap)
            tma_load_2d(b_base + (uint32_t)(ch * NT + tap) * p.b_slice_bytes, &tm_wgt, bres_bar,
                        tap * p.cin_total + ch * KC, 0);
      }
      __syncwarp();
    } else {
      const bool leader = elect_one();
      int s = 0; uint32_t ph = 0;
      int qk = 0;
      for (int tl = t_first; CHAIN || tl < t_end; tl += t_step) {
        if (CHAIN) {
          int item = 0;
          if (leader) item = q_get(qk);
          ++qk;
          item = __shfl_sync(0xffffffffu, item, __ffs(__ballot_sync(0xffffffffu, leader)) - 1);
          if (item < 0) break;
          tl = item;
        }
        const int tile = tile_of(tl);
        const int ncol = (tile - fast_div(tile, p.div_ntiles) * p.n_tiles) * p.block_n;
        if (SPXP) {
          // slices in the MMA warp's issue order; each carries only the rows (GEMM columns) its tap can reach
          // (see spx_tap): all bn, half of them or a quarter, through the map with the matching box height
          int sl = 0, par = 0, cc = 0;       // no divisions here: this one thread feeds ~25 slices per tile
          for (int ch = 0; ch < p.chunks; ++ch) {
            const bool is_x = ch < p.split_chunk;
            const int nt = (SPX == 2) ? ((par >> 1) + 1) * ((par & 1) + 1) : (is_x ? 9 : 4);
            for (int t = 0; t < nt; ++t, ++sl) {
              if (leader) {
                int R, C;
                if (is_x) {
                  const int tap = (t == 0) ? 4 : (t <= 4 ? t - 1 : t);                 // (1,1) first, then row-major
                  R = (tap >= 6) ? 2 : (tap >= 3 ? 1 : 0); C = tap - 3 * R;
                } else {
                  R = 1 - (par >> 1) + (t >> 1); C = 1 - (par & 1) + (t & 1);
                }
                const int qn = (SPX == 2 || R == 1) ? 4 : (C == 1 ? 2 : 1);
                const int qoff = (SPX == 2 || R == 1) ? 0 : 2 * (R == 2) + (C == 1 ? 0 : (C == 2));
                const CUtensorMap* tm = (qn == 4) ? &tm_wgt : (qn == 2 ? &tm_out : &tm_res);
                mbar_wait(bempty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(bfull_bar(s), (uint32_t)(qn * (p.block_n >> 2)) * KC * 2u);
                tma_load_2d(b_base + (uint32_t)s * b_stage_bytes, tm, bfull_bar(s), sl * KC,
                            (SPX == 2 ? ncol : 0) + qoff * (p.block_n >> 2));
              }
              if (++s == p.b_stages) { s = 0; ph ^= 1u; }
            }
            if (!is_x && ++cc == p.spx_cpp) { cc = 0; ++par; }
          }
          continue;
        }
        if (SPX == 3) {      // the 2x2 taps of this N tile's output parity, one slice per stage
          const int q = ncol / p.block_n, qh = q >> 1, qw = q & 1;
          for (int ch = 0; ch < p.chunks; ++ch) {
            for (int t = 0; t < 4; ++t) {
              if (leader) {
                const int tap = (qh + (t >> 1)) * KW + qw + (t & 1);
                mbar_wait(bempty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(bfull_bar(s), slice_tx);
                tma_load_2d(b_base + (uint32_t)s * b_stage_bytes, &tm_wgt, bfull_bar(s), tap * p.cin_total + ch * KC, ncol);
              }
              if (++s == p.b_stages) { s = 0; ph ^= 1u; }
            }
          }
          continue;
        }
        const CUtensorMap* wmap = CHAIN ? &chain->layer[layer_of(tl)].tm_wgt : &tm_wgt;
        for (int ch = 0; ch < p.chunks; ++ch) {
          for (int tap0 = 0; tap0 < NT; tap0 += p.kpb) {
            if (leader) {
              mbar_wait(bempty_bar(s), ph ^ 1u);
              const uint32_t dst = b_base + (uint32_t)s * b_stage_bytes;
              if (CG2) {
                // this CTA's half of the rows of every slice; the bytes of BOTH CTAs complete on the leader's barrier,
                // which only the leader arms (a remote arrive per slice costs the peer's producer ~500 cycles)
                const uint32_t fb = lead(bfull_bar(s));
                if (crank == 0) mbar_arrive_expect_tx(bfull_bar(s), (uint32_t)p.kpb * slice_tx);
                for (int j = 0; j < p.kpb; ++j)
                  tma_load_2d_cg2(dst + (uint32_t)j * p.b_slice_bytes, &tm_wgt, fb, (tap0 + j) * p.cin_total + ch * KC,
                                  ncol + (int)crank * (p.block_n >> 1));
              } else {
                mbar_arrive_expect_tx(bfull_bar(s), (uint32_t)p.kpb * slice_tx);
                for (int j = 0; j < p.kpb; ++j)
                  tma_load_2d(dst + (uint32_t)j * p.b_slice_bytes, wmap, bfull_bar(s),
                              (tap0 + j) * p.cin_total + ch * KC, ncol);
              }
            }
            if (++s == p.b_stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one elected thread)
    const uint32_t idesc = make_idesc_bf16(CG2 ? 2 * kTileM : kTileM, p.block_n);
    auto commit = [&](uint32_t bar) { if (CG2) umma_commit_cg2(bar, (uint16_t)3); else umma_commit(bar); };
    // A descriptors, per stage layout.  swizzled pixel-major (TMA box): hi = SBO (one halo row) | version | swizzle,
    // LBO unused.  no-swizzle planes (cp.async gather): hi = SBO (halo_w x 16 B) | version, LBO = plane stride.
    constexpr uint32_t a_hi_sw = (((uint32_t)PW * ROWB) >> 4) | (1u << 14) | (B_LAYOUT << 29);
    constexpr uint32_t a_hi_pl = (uint32_t)PW | (1u << 14);
    const uint32_t a_lo0_sw = ((a_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_lo0_pl = ((a_base & 0x3FFFFu) >> 4) | (PLANE_UNITS << 16);
    constexpr uint32_t a_stage_units = A_STAGE_BYTES >> 4;
    // B: swizzled K-major slices, rows of KC*2 bytes
    constexpr uint32_t b_hi = ((8u * KC * 2u) >> 4) | (1u << 14) | (B_LAYOUT << 29);
    const uint32_t b_lo0 = ((b_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_slice_units = p.b_slice_bytes >> 4;
    const uint32_t b_stage_units = b_stage_bytes >> 4;
    const uint32_t bn = (uint32_t)p.block_n;
    const bool skip_mma = UWM_DBG_OF(p) & 2;
    const uint32_t idesc_half = make_idesc_bf16(kTileM, SPX == 1 ? p.block_n >> 1 : p.block_n);
    const uint32_t idesc_quarter = make_idesc_bf16(kTileM, SPX == 1 ? p.block_n >> 2 : p.block_n);
    if (elect_one() && (!CG2 || crank == 0)) {
      if (RESIDENT) mbar_wait(bres_bar, 0);
      int sa = 0; uint32_t pha = 0;
      int sb = 0; uint32_t phb = 0;
      int it = 0;
      for (int tl = t_first; CHAIN || tl < t_end; tl += t_step, ++it) {
      if (CHAIN) { tl = q_get(it); if (tl < 0) break; }
      const int tile = (SPX == 3) ? tile_of(tl) : 0;
        const int acc = it & nacc_mask;
        const uint32_t tmem_acc = tmem_base + (uint32_t)acc * (TG * bn);
        mbar_wait_fast(acce_bar(acc), ((uint32_t)(it >> nacc_log2) & 1u) ^ 1u);
        tc_fence_after();
        halo_trace(p, 160 + 3 * it);
        for (int ch = 0; ch < p.chunks; ++ch) {
          mbar_wait_fast(afull_bar(sa), pha);
          if (ch == 0) halo_trace(p, 161 + 3 * it);
          if (!A_TMA) fence_proxy_async_smem();        // loaders wrote through the generic proxy (cp.async)
          tc_fence_after();
          uint32_t b_lo = b_lo0 + (uint32_t)(ch * NT) * b_slice_units;   // resident: slice (ch, tap 0)
          int j = 0;                                                      // streamed: slice inside the stage
          // one K chunk: all taps x sub-tiles x K=16 slices, instantiated per stage layout (immediate offsets)
          auto issue_chunk = [&](auto swz_c, auto r0_c, auto c0_c, auto nr_c, auto nc_c) {
            constexpr bool SWZ = decltype(swz_c)::value;
            constexpr int R0 = decltype(r0_c)::value, C0 = decltype(c0_c)::value;     // tap sub-rectangle of the halo
            constexpr int NR = decltype(nr_c)::value, NC = decltype(nc_c)::value;
            constexpr uint32_t a_hi = SWZ ? a_hi_sw : a_hi_pl;
            constexpr uint32_t a_px_units = SWZ ? (ROWB >> 4) : 1u;       // 16-byte units per halo pixel step
            constexpr uint32_t a_k_units = SWZ ? 2u : 2u * PLANE_UNITS;   // 16-byte units per K=16 slice
            const uint32_t a_st = (SWZ ? a_lo0_sw : a_lo0_pl) + (uint32_t)sa * a_stage_units;
#pragma unroll
            for (int tap = 0; tap < NR * NC; ++tap) {
              const uint32_t shift = (uint32_t)((R0 + tap / NC) * PW + (C0 + tap % NC));
              if (!RESIDENT && j == 0) {
                mbar_wait_fast(bfull_bar(sb), phb);
                tc_fence_after();
                b_lo = b_lo0 + (uint32_t)sb * b_stage_units;
              }
#pragma unroll
              for (int g = 0; g < TG; ++g) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k) {
                  if (S2D) {       // parity plane k meets 4 of the 9 taps (folds at compile time); its weight block is pair j
                    if (!s2d_meets(tap, k)) continue;
                    constexpr uint32_t b_hi_s2d = ((8u * 32u) >> 4) | (1u << 14) | (kLayoutSw32 << 29);
                    const uint32_t b_blk = b_lo0 + (uint32_t)s2d_pair(tap, k) * (bn * 2u);          // bn x 32 bytes per block
                    if (tap == 0 && k == 3)
                      umma_halo<false, CG2>(tmem_acc + g * bn, a_st + (shift + 8u * g) * a_px_units + k * a_k_units, a_hi,
                                            b_blk, b_hi_s2d, idesc, (uint32_t)ch);
                    else
                      umma_halo<true, CG2>(tmem_acc + g * bn, a_st + (shift + 8u * g) * a_px_units + k * a_k_units, a_hi,
                                           b_blk, b_hi_s2d, idesc, 1u);
                    continue;
                  }
                  if (tap == 0 && k == 0)
                    umma_halo<false, CG2>(tmem_acc + g * bn, a_st + (shift + 8u * g) * a_px_units + k * a_k_units, a_hi,
                                          b_lo + 2u * k, b_hi, idesc, (uint32_t)ch);
                  else
                    umma_halo<true, CG2>(tmem_acc + g * bn, a_st + (shift + 8u * g) * a_px_units + k * a_k_units, a_hi,
                                         b_lo + 2u * k, b_hi, idesc, 1u);
                }
              }
              b_lo += b_slice_units;
              if (!RESIDENT && ++j == p.kpb) {
                j = 0;
                commit(bempty_bar(sb));
                if (++sb == p.b_stages) { sb = 0; phb ^= 1u; }
              }
            }
          };
          using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>;
          using I2 = std::integral_constant<int, 2>;
          using IH = std::integral_constant<int, KH>; using IW = std::integral_constant<int, KW>;
          if (!skip_mma) {
            if (SPXP) {
              // One tap (R,C) of the 3x3 block neighbourhood.  Row R = 0 only reaches output parity qh = 0, R = 2 only
              // qh = 1, R = 1 both (same for columns), so with GEMM columns ordered (qh, qw, co) the tap needs
              // N = bn (R == 1), bn/2 at column qh*bn/2 (R != 1, C == 1) or bn/4 at column (2qh+qw)*bn/4: fewer
              // tensor cycles (32 + N/4 per MMA) and a smaller weight slice.  The first tap of a tile is (1,1): it
              // writes all bn columns, everything after it accumulates.
              auto spx_tap = [&](auto r_c, auto c_c, uint32_t accumulate) {
                constexpr int R = decltype(r_c)::value, C = decltype(c_c)::value;
                constexpr int QN = (SPX == 2 || R == 1) ? 4 : (C == 1 ? 2 : 1);          // stride-2 mode: always all columns
                constexpr int QOFF = (SPX == 2 || R == 1) ? 0 : 2 * (R == 2) + (C == 1 ? 0 : (C == 2));
                constexpr uint32_t shift = (uint32_t)(R * PW + C);
                const uint32_t a_st = a_lo0_sw + (uint32_t)sa * a_stage_units;
                const uint32_t idesc_t = QN == 4 ? idesc : (QN == 2 ? idesc_half : idesc_quarter);
                mbar_wait_fast(bfull_bar(sb), phb);
                tc_fence_after();
                const uint32_t b_st = b_lo0 + (uint32_t)sb * b_stage_units;
#pragma unroll
                for (int g = 0; g < TG; ++g) {
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k) {
                    const uint32_t d = tmem_acc + g * bn + QOFF * (bn >> 2);
                    if (k == 0) umma_halo<false>(d, a_st + (shift + 8u * g) * (ROWB >> 4) + 2u * k, a_hi_sw, b_st + 2u * k, b_hi, idesc_t, accumulate);
                    else umma_halo<true>(d, a_st + (shift + 8u * g) * (ROWB >> 4) + 2u * k, a_hi_sw, b_st + 2u * k, b_hi, idesc_t, 1u);
                  }
                }
                commit(bempty_bar(sb));
                if (++sb == p.b_stages) { sb = 0; phb ^= 1u; }
              };
              if (SPX == 2) {
                // stride-2 conv: plane (ph,pw) meets halo rows {1} / {0,1} (ph = 0 / 1), columns likewise
                const int par = ch / p.spx_cpp;
                if (par == 0) { spx_tap(I1{}, I1{}, (uint32_t)ch); }
                else if (par == 1) { spx_tap(I1{}, I0{}, 1u); spx_tap(I1{}, I1{}, 1u); }
                else if (par == 2) { spx_tap(I0{}, I1{}, 1u); spx_tap(I1{}, I1{}, 1u); }
                else { spx_tap(I0{}, I0{}, 1u); spx_tap(I0{}, I1{}, 1u); spx_tap(I1{}, I0{}, 1u); spx_tap(I1{}, I1{}, 1u); }
              } else if (ch < p.split_chunk) {
                spx_tap(I1{}, I1{}, (uint32_t)ch);
                spx_tap(I0{}, I0{}, 1u); spx_tap(I0{}, I1{}, 1u); spx_tap(I0{}, I2{}, 1u);
                spx_tap(I1{}, I0{}, 1u); spx_tap(I1{}, I2{}, 1u);
                spx_tap(I2{}, I0{}, 1u); spx_tap(I2{}, I1{}, 1u); spx_tap(I2{}, I2{}, 1u);
              } else {
                // parity plane (ph,pw) of the skip source: taps {1-ph, 2-ph} x {1-pw, 2-pw}
                const int par = (ch - p.split_chunk) / p.spx_cpp;
                if (par == 0) { spx_tap(I1{}, I1{}, 1u); spx_tap(I1{}, I2{}, 1u); spx_tap(I2{}, I1{}, 1u); spx_tap(I2{}, I2{}, 1u); }
                else if (par == 1) { spx_tap(I1{}, I0{}, 1u); spx_tap(I1{}, I1{}, 1u); spx_tap(I2{}, I0{}, 1u); spx_tap(I2{}, I1{}, 1u); }
                else if (par == 2) { spx_tap(I0{}, I1{}, 1u); spx_tap(I0{}, I2{}, 1u); spx_tap(I1{}, I1{}, 1u); spx_tap(I1{}, I2{}, 1u); }
                else { spx_tap(I0{}, I0{}, 1u); spx_tap(I0{}, I1{}, 1u); spx_tap(I1{}, I0{}, 1u); spx_tap(I1{}, I1{}, 1u); }
              }
            }
            else if (SPX == 3) {
              const int q = tile - fast_div(tile, p.div_ntiles) * p.n_tiles;      // N tile = output parity (qh,qw)
              if (q == 0) issue_chunk(std::true_type{}, I0{}, I0{}, I2{}, I2{});
              else if (q == 1) issue_chunk(std::true_type{}, I0{}, I1{}, I2{}, I2{});
              else if (q == 2) issue_chunk(std::true_type{}, I1{}, I0{}, I2{}, I2{});
              else issue_chunk(std::true_type{}, I1{}, I1{}, I2{}, I2{});
            }
            else if (A_TMA) issue_chunk(std::true_type{}, I0{}, I0{}, IH{}, IW{});
            else if (p.mix && ch >= p.split_chunk) issue_chunk(std::true_type{}, I0{}, I0{}, IH{}, IW{});  // skip source: TMA box
            else issue_chunk(std::false_type{}, I0{}, I0{}, IH{}, IW{});                                   // upsampled source: gather
          } else if (!RESIDENT) {
            const int par2 = (SPX == 2) ? ch / p.spx_cpp : 0;
            const int nt_ch = (SPX == 2) ? ((par2 >> 1) + 1) * ((par2 & 1) + 1) : ((SPX == 3 || (SPXP && ch >= p.split_chunk)) ? 4 : NT);
            for (int tap0 = 0; tap0 < nt_ch; tap0 += p.kpb) {
              mbar_wait_fast(bfull_bar(sb), phb);
              commit(bempty_bar(sb));
              if (++sb == p.b_stages) { sb = 0; phb ^= 1u; }
            }
          }
          commit(aempty_bar(sa));
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
        }
        commit(accf_bar(acc));
        halo_trace(p, 162 + 3 * it);
      }
    }
  } else if (warp < kHaloLoaderWarp0 || warp >= kHaloLoaderWarp0 + 4) {
    // ------------------------------------------------------------ epilogue (2 x 4 warps)
    // Per 16-column chunk the bias is loaded once and the TG sub-tiles are walked with the TMEM load of
    // sub-tile g+1 in flight while sub-tile g is biased / ReLU'd / packed / stored.
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;           // accumulator row == pixel of the 8x16 sub-tile
    const int wi = row & (kHaloTW - 1);
    const int hi = row >> 3;
    // ReLU on the packed bf16 pairs (8 HMNMX2 per 16 columns instead of 16 FMNMX): max(round(x), 0) == round(max(x, 0))
    uint32_t lo_clamp2 = p.relu ? 0u : 0xff80ff80u;              // +0 | +0, or -inf | -inf = no clamp
    const uint32_t bn = (uint32_t)p.block_n;
    const int eset = (warp >= kHaloLoaderWarp0) ? 1 : 0;       // which of the two warps of this lane quarter
    constexpr int GN = (TG >= 2) ? TG / 2 : 1;                 // sub-tiles per warp: g = GSTEP*gi + g_first
    constexpr int GSTEP = (TG >= 2) ? 2 : 1;
    const int g_first = (TG >= 2) ? eset : 0;
    const int c_first = (TG >= 2) ? 0 : 16 * eset, c_step = (TG >= 2) ? 16 : 32;
    const bool tma_out = p.ep_tma != 0;
    const int ewarp = (warp & 3) + 4 * eset;                                  // 0..7
    // stage the bias (all 8 epilogue warps, then a named barrier among them); a chain stages every layer's
    const int cout_all = p.n_tiles * p.block_n;
    if (CHAIN) {
      for (int l = 0; l < n_layers; ++l) {
        const float* bl = chain->layer[l].bias;
        for (int i = ewarp * 32 + lane; i < cout_all; i += 256) s_bias[l * cout_all + i] = __ldg(bl + i);
      }
    } else {
      for (int i = ewarp * 32 + lane; i < cout_all; i += 256) s_bias[i] = __ldg(p.bias + i);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t stg = stg_base + (uint32_t)ewarp * (32u * 128u);           // one 32-row x 128-byte buffer per warp
    // 128B swizzle of the 16-byte chunks of this lane's row: chunk index ^ (row & 7)
    const uint32_t my_row_off = (uint32_t)lane * 128u;
    const uint32_t sw_xor = (uint32_t)lane & 7u;
    uint32_t res_cnt = 0;                      // residual boxes consumed by this warp (barrier parity)
    pdl_wait();                              // residual reads / output writes depend on the previous kernel
    int it = 0;
    for (int tl = t_first; CHAIN || tl < t_end; tl += t_step, ++it) {
      if (CHAIN) {
        int item = 0;
        if (lane == 0) item = q_get(it);
        tl = __shfl_sync(0xffffffffu, item, 0);
        if (tl < 0) break;
      }
      const int tile = tile_of(tl);
      const int layer = layer_of(tl);
      // per-layer epilogue parameters of a chain: output / residual maps, bias row, ReLU, residual flag
      const CUtensorMap* omap = CHAIN ? &chain->layer[layer].tm_out : &tm_out;
      const CUtensorMap* rmap = CHAIN ? &chain->layer[layer].tm_res : &tm_res;
      const bool has_res = CHAIN ? (chain->layer[layer].has_res != 0) : (p.res != nullptr);
      if (CHAIN) lo_clamp2 = chain->layer[layer].relu ? 0u : 0xff80ff80u;
      const int acc = it & nacc_mask;
      const HaloTile t = halo_decode<TG>(p, tile);
      const int oh = t.h0 + hi;
      const int ow0 = t.w0 + wi;
      const int col0 = t.n_tile * p.block_n;
      // TMA coordinates of an output / residual box.  SPX == 3: the N tile is output parity (qh,qw) of a 2x larger tensor
      // whose map walks every second pixel, so a box starts at pixel (2h + qh, 2w + qw), channel = column inside the tile
      const int ec0 = (SPX == 3) ? 0 : col0;
      auto ex = [&](int wpx) { return (SPX == 3) ? 2 * wpx + (t.n_tile & 1) : wpx; };
      auto ey = [&](int hpx) { return (SPX == 3) ? 2 * hpx + (t.n_tile >> 1) : hpx; };
      const bool row_ok = oh < p.h;
      const long long pix0 = ((long long)t.img * p.h + oh) * p.w + ow0;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * (TG * bn);

      // Residual convs with staged TMA stores: the residual box of the warp's first (sub-tile, 64 channels) unit is
      // TMA-loaded into the warp's staging buffer NOW, long before the accumulator is complete; the epilogue adds it
      // in place.  (Lane-per-pixel global loads exposed an L2 round trip per 16-column item in the drain.)
      const bool res_tma = tma_out && has_res && !(UWM_DBG_OF(p) & 4);
      if (CHAIN && res_tma && lane == 0) {
        // the residual is an earlier chain layer's output of the same image: it must be complete before the first
        // residual box of this tile is requested (the loader's own wait does not order THIS thread's TMA reads)
        const int rl = chain->layer[layer].res_layer;
        if (rl >= 0) { chain_wait_dep(chain->dep + rl * p.n_img + t.img, chain->dep_target); fence_proxy_async_all(); }
      }
      if (res_tma && lane == 0 && c_first * 4 < min(p.block_n, p.cout - col0)) {
        bulk_wait_read<0>();                 // the previous tile's last store has read the buffer out
        mbar_arrive_expect_tx(resbar(ewarp), 32u * 128u);
        tma_load_4d(stg, rmap, resbar(ewarp), ec0 + c_first * 4, ex(t.w0 + g_first * kHaloTW), ey(t.h0 + q * 4), t.img);
      }

      if (lane == 0) mbar_wait(accf_bar(acc), (uint32_t)(it >> nacc_log2) & 1u);
      __syncwarp();
      tc_fence_after();
      if (warp == 2 && lane == 0) halo_trace(p, 400 + 2 * it);

      if (UWM_DBG_OF(p) & 4) {
        // bench-only: no epilogue work
      } else if (p.head) {
        if (S2D) {
          // 4 GEMM columns = the 2x2 output pixels of this block: logits / mask rows 2*oh + qh, columns 2*ow + qw
          if (TG >= 2 || eset == 0) {
            uint32_t z[GN][4];
#pragma unroll
            for (int gi = 0; gi < GN; ++gi) tmem_ld_x4(taddr0 + (uint32_t)(GSTEP * gi + g_first) * bn, z[gi]);
            tmem_ld_wait();
            const float b = __ldg(p.bias);
            const bool pair_ok = ((reinterpret_cast<uintptr_t>(p.mask) & 1) == 0);
#pragma unroll
            for (int gi = 0; gi < GN; ++gi) {
              const int g = GSTEP * gi + g_first;
              if (row_ok && ow0 + g * kHaloTW < p.w) {
#pragma unroll
                for (int qh = 0; qh < 2; ++qh) {
                  const long long o = ((long long)(t.img * 2 * p.h + 2 * oh + qh) * (2 * p.w)) + 2 * (ow0 + g * kHaloTW);
                  const float z0 = __uint_as_float(z[gi][2 * qh]) + b, z1 = __uint_as_float(z[gi][2 * qh + 1]) + b;
                  if (p.logits) {
                    p.logits[o] = p.apply_sigmoid ? 1.f / (1.f + __expf(-z0)) : z0;
                    p.logits[o + 1] = p.apply_sigmoid ? 1.f / (1.f + __expf(-z1)) : z1;
                  }
                  if (p.mask) {
                    const uint8_t m0 = (z0 > p.thr_logit) ? 255 : 0, m1 = (z1 > p.thr_logit) ? 255 : 0;
                    if (pair_ok) *reinterpret_cast<uchar2*>(p.mask + o) = make_uchar2(m0, m1);
                    else { p.mask[o] = m0; p.mask[o + 1] = m1; }
                  }
                }
              }
            }
          }
        } else if (TG >= 2 || eset == 0) {
          uint32_t z[GN];
#pragma unroll
          for (int gi = 0; gi < GN; ++gi) tmem_ld_x1(taddr0 + (uint32_t)(GSTEP * gi + g_first) * bn, z[gi]);
          tmem_ld_wait();
          const float b = __ldg(p.bias);
#pragma unroll
          for (int gi = 0; gi < GN; ++gi) {
            const int g = GSTEP * gi + g_first;
            if (row_ok && ow0 + g * kHaloTW < p.w) {
              const float zz = __uint_as_float(z[gi]) + b;
              if (p.logits) p.logits[pix0 + g * kHaloTW] = p.apply_sigmoid ? 1.f / (1.f + __expf(-zz)) : zz;
              if (p.mask) p.mask[pix0 + g * kHaloTW] = (zz > p.thr_logit) ? 255 : 0;
            }
          }
        }
      } else {
        // Items of this warp: 16-column chunk j of column group cg0 (64 columns with TMA stores, else 16) of
        // sub-tile g.  The TMEM load of the next item and its residual are in flight while the current one is
        // biased / ReLU'd / packed.  The item body is instantiated per output mode (generic lambda with
        // compile-time flags): a warp runs ~4-10 cycles per instruction here, so the per-item branches of a
        // run-time-configured loop cost as much as the arithmetic.
        const int ncols = min(p.block_n, p.cout - col0);
        const uint32_t bsm = smem_u32(s_bias + (CHAIN ? layer * cout_all : 0) + col0);   // folded-BN bias, staged in smem
        auto run_items = [&](auto tma_c, auto res_c, auto shuf_c) {
          constexpr bool TMA = decltype(tma_c)::value, RES = decltype(res_c)::value, SHUF = decltype(shuf_c)::value;
          constexpr int LOG2_JN = TMA ? 2 : 0, JN = 1 << LOG2_JN;           // chunks per column group
          constexpr int NIT = GN << LOG2_JN;                                // items per column group for this warp
          __nv_bfloat16* obase = p.out + pix0 * p.out_pitch + col0;
          const __nv_bfloat16* rbase = RES ? p.res + pix0 * p.res_pitch + col0 : nullptr;
          const long long ostep = (long long)kHaloTW * p.out_pitch, rstep = (long long)kHaloTW * p.res_pitch;
          for (int cg0 = c_first * JN; cg0 < ncols; cg0 += c_step * JN) {  // first column of the group
            uint32_t v[2][16];
            uint4 r0[2], r1[2];
            auto item_load = [&](int n, int b) {
              const int g = GSTEP * (n >> LOG2_JN) + g_first, c = cg0 + ((n & (JN - 1)) << 4);
              tmem_ld_x16(taddr0 + (uint32_t)g * bn + c, v[b]);
              if (RES && !TMA) {
                r0[b] = make_uint4(0, 0, 0, 0); r1[b] = make_uint4(0, 0, 0, 0);
                if (row_ok && (ow0 + g * kHaloTW < p.w)) {
                  const __nv_bfloat16* rsrc = rbase + g * rstep + c;
                  if (SHUF) {            // the residual lives at the shuffled output position
                    const int par = (col0 + c) / p.shuffle, co = col0 + c - par * p.shuffle;
                    const long long hp = ((long long)(t.img * 2 * p.h + 2 * oh + (par >> 1)) * (2 * p.w) +
                                          2 * (ow0 + g * kHaloTW) + (par & 1));
                    rsrc = p.res + hp * p.res_pitch + co;
                  }
                  r0[b] = *reinterpret_cast<const uint4*>(rsrc);
                  r1[b] = *reinterpret_cast<const uint4*>(rsrc + 8);
                }
              }
            };
            item_load(0, 0);
#pragma unroll
            for (int n = 0; n < NIT; ++n) {
              const int u = n & 1;
              const int j = n & (JN - 1);
              const int g = GSTEP * (n >> LOG2_JN) + g_first, c = cg0 + (j << 4);
              const bool valid = row_ok && (ow0 + g * kHaloTW < p.w);
              const uint4 b0 = ld_shared_v4(bsm + c * 4), b1 = ld_shared_v4(bsm + c * 4 + 16);
              const uint4 b2 = ld_shared_v4(bsm + c * 4 + 32), b3 = ld_shared_v4(bsm + c * 4 + 48);
              const float bb[16] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z), __uint_as_float(b0.w),
                                    __uint_as_float(b1.x), __uint_as_float(b1.y), __uint_as_float(b1.z), __uint_as_float(b1.w),
                                    __uint_as_float(b2.x), __uint_as_float(b2.y), __uint_as_float(b2.z), __uint_as_float(b2.w),
                                    __uint_as_float(b3.x), __uint_as_float(b3.y), __uint_as_float(b3.z), __uint_as_float(b3.w)};
              tmem_ld_wait();
              if (warp == 2 && lane == 0 && it < 20 && cg0 == c_first * JN) halo_trace(p, 600 + 16 * it + 2 * n);       // item n: accumulator chunk in registers
              if (n + 1 < NIT) item_load(n + 1, (n + 1) & 1);
              if (valid || TMA) {
                float f[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[u][k]) + bb[k];
                if (RES) {
                  uint4 x0, x1;
                  if (TMA) {               // residual box of this unit sits in the staging buffer (own row, same swizzle)
                    if (j == 0) {
                      if (lane == 0) mbar_wait(resbar(ewarp), res_cnt & 1u);
                      __syncwarp();
                      ++res_cnt;
                    }
                    x0 = ld_shared_v4(stg + my_row_off + (((uint32_t)(2 * j) ^ sw_xor) << 4));
                    x1 = ld_shared_v4(stg + my_row_off + (((uint32_t)(2 * j + 1) ^ sw_xor) << 4));
                  } else {
                    x0 = r0[u]; x1 = r1[u];
                  }
                  const uint32_t rr[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                  for (int k = 0; k < 8; ++k) { f[2 * k] += bf16_lo(rr[k]); f[2 * k + 1] += bf16_hi(rr[k]); }
                }
                uint4 o0, o1;
                o0.x = relu_bf16x2(pack_bf16x2(f[0], f[1]), lo_clamp2);   o0.y = relu_bf16x2(pack_bf16x2(f[2], f[3]), lo_clamp2);
                o0.z = relu_bf16x2(pack_bf16x2(f[4], f[5]), lo_clamp2);   o0.w = relu_bf16x2(pack_bf16x2(f[6], f[7]), lo_clamp2);
                o1.x = relu_bf16x2(pack_bf16x2(f[8], f[9]), lo_clamp2);   o1.y = relu_bf16x2(pack_bf16x2(f[10], f[11]), lo_clamp2);
                o1.z = relu_bf16x2(pack_bf16x2(f[12], f[13]), lo_clamp2); o1.w = relu_bf16x2(pack_bf16x2(f[14], f[15]), lo_clamp2);
                if (TMA) {
                  if (j == 0 && !RES) {  // the previous store of this warp must have read the buffer out
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                  }
                  st_shared_v4(stg + my_row_off + (((uint32_t)(2 * j) ^ sw_xor) << 4), o0);
                  st_shared_v4(stg + my_row_off + (((uint32_t)(2 * j + 1) ^ sw_xor) << 4), o1);
                  if (j == JN - 1) {
                    // unit complete: make the generic-proxy smem writes visible to the async proxy, one lane stores
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                      tma_store_4d(omap, stg, ec0 + cg0, ex(t.w0 + g * kHaloTW), ey(t.h0 + q * 4), t.img);
                      bulk_commit();
                      if (RES) {           // next unit of this tile: its residual box, once the store has read the buffer
                        const int gi_n = (n >> LOG2_JN) + 1;
                        const int g_n = (gi_n < GN) ? GSTEP * gi_n + g_first : g_first;
                        const int cg_n = (gi_n < GN) ? cg0 : cg0 + c_step * JN;
                        if (cg_n < ncols) {
                          bulk_wait_read<0>();
                          mbar_arrive_expect_tx(resbar(ewarp), 32u * 128u);
                          tma_load_4d(stg, rmap, resbar(ewarp), ec0 + cg_n, ex(t.w0 + g_n * kHaloTW), ey(t.h0 + q * 4), t.img);
                        }
                      }
                    }
                  }
                } else if (SHUF) {
                  const int par = (col0 + c) / p.shuffle, co = col0 + c - par * p.shuffle;   // chunk -> (parity, channel)
                  const long long hp = ((long long)(t.img * 2 * p.h + 2 * oh + (par >> 1)) * (2 * p.w) +
                                        2 * (ow0 + g * kHaloTW) + (par & 1));
                  __nv_bfloat16* dst = p.out + hp * p.out_pitch + co;
                  *reinterpret_cast<uint4*>(dst) = o0;
                  *reinterpret_cast<uint4*>(dst + 8) = o1;
                } else {
                  *reinterpret_cast<uint4*>(obase + g * ostep + c) = o0;
                  *reinterpret_cast<uint4*>(obase + g * ostep + c + 8) = o1;
                }
              }
              if (warp == 2 && lane == 0 && it < 20 && cg0 == c_first * JN) halo_trace(p, 601 + 16 * it + 2 * n);       // item n: stored / staged
            }
          }
        };
        using T_ = std::true_type; using F_ = std::false_type;
        if (tma_out) { if (has_res) run_items(T_{}, T_{}, F_{}); else run_items(T_{}, F_{}, F_{}); }
        else if (CHAIN) { }                                       // chains always store through TMA
        else if (p.shuffle) { if (p.res) run_items(F_{}, T_{}, T_{}); else run_items(F_{}, F_{}, T_{}); }
        else if (p.res) run_items(F_{}, T_{}, F_{});
        else run_items(F_{}, F_{}, F_{});
      }

      // every tcgen05.ld of this accumulator set has completed (tcgen05.wait::ld above): hand the set back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CG2 && crank != 0) mbar_arrive_cluster(lead(acce_bar(acc))); else mbar_arrive(acce_bar(acc)); }
      if (warp == 2 && lane == 0) halo_trace(p, 401 + 2 * it);
      if (CHAIN && lane == 0) {
        // (after the accumulator set went back to the MMA warp) this warp's stores of the tile are complete -> count
        // them into the image's counter of this layer.  The stores went through the async proxy: wait for their
        // completion (not just for the smem read), then a proxy fence, then the release.
        bulk_wait_done<0>();
        if (warp == 2) halo_trace(p, 560 + it);                 // trace: this warp's stores of the tile are complete
        fence_proxy_async_all();
        red_release_gpu_add(chain->dep + layer * p.n_img + t.img, 1);
      }
    }
    if (tma_out && lane == 0) bulk_wait_read<0>();     // smem must outlive the last store's read
  } else {
    // ------------------------------------------------------------ activation loaders (4 warps)
    const int lt = threadIdx.x - kHaloLoaderWarp0 * 32;
    if (lt == 0) halo_trace(p, 2);
    pdl_wait();                              // activations are written by the previous kernel
    if (lt == 0) halo_trace(p, 3);
    int s = 0; uint32_t ph = 0;
    int nstage = 0;
    if (A_TMA) {
      // one thread, one TMA box per stage; out-of-image coordinates are zero-filled (= conv padding)
      if (lt == 0) {
        int lk = 0;
        const unsigned long long pol0 = p.a_policy[0] ? p.a_policy[0] : kL2EvictNormal;
        const unsigned long long pol1 = p.a_policy[1] ? p.a_policy[1] : kL2EvictNormal;
        for (int tl = t_first; CHAIN || tl < t_end; tl += t_step) {
        if (CHAIN) {
          // grab the next work item (right after the previous item's loads were issued, i.e. about one item ahead of
          // the MMAs) and publish it to the weight producer, the MMA thread and the epilogue warps
          mbar_wait(qempty_bar(lk & 3), ((uint32_t)(lk >> 2) & 1u) ^ 1u);
          tl = atomicAdd(chain->dep + n_layers * p.n_img + 1, 1);
          if (tl >= t_end) tl = -1;
          s_items[lk & 3] = tl;
          mbar_arrive(qfull_bar(lk & 3));
          ++lk;
          if (tl < 0) break;
        }
        const int tile = tile_of(tl);
          const HaloTile t = halo_decode<TG>(p, tile);
          const int hbase = t.h0 + p.dh_min, wbase = t.w0 + p.dw_min;
          const int layer = layer_of(tl);
          const CUtensorMap* amap = CHAIN ? &chain->layer[layer].tm_a0 : &tm_a0;
          if (CHAIN && layer > 0) {
            // every tile of this image of the previous layer has been stored (acquire), then order the TMA reads after it
            halo_trace(p, 500 + 2 * (nstage / p.chunks));            // trace: dependency wait of this item (begin / end)
            chain_wait_dep(chain->dep + (layer - 1) * p.n_img + t.img, chain->dep_target);
            fence_proxy_async_all();
            halo_trace(p, 501 + 2 * (nstage / p.chunks));
          }
          for (int ch = 0; ch < p.chunks; ++ch) {
            mbar_wait(aempty_bar(s), ph ^ 1u);
            halo_trace(p, 16 + 2 * nstage);
            if (CG2) {       // both CTAs' boxes complete on the leader's barrier (CG2 kernels have one source)
              const uint32_t fa = lead(afull_bar(s));
              if (!(UWM_DBG_OF(p) & 1)) {
                if (crank == 0) mbar_arrive_expect_tx(afull_bar(s), 2u * (uint32_t)NPIX * ROWB);   // both CTAs' boxes
                tma_load_4d_cg2(a_base + (uint32_t)s * A_STAGE_BYTES, &tm_a0, fa, ch * KC, wbase * p.a_scale,
                                hbase * p.a_scale, t.img);
              } else if (crank == 0) {
                mbar_arrive(afull_bar(s));
              }
            } else if (!(UWM_DBG_OF(p) & 1)) {
              mbar_arrive_expect_tx(afull_bar(s), (uint32_t)NPIX * ROWB);
              if (ch < p.split_chunk)
                tma_load_4d_hint(a_base + (uint32_t)s * A_STAGE_BYTES, amap, afull_bar(s), ch * KC, wbase * p.a_scale,
                                 hbase * p.a_scale, t.img, pol0);
              else if (SPXP) {
                // parity plane (ph,pw) of the full-resolution skip source: halo block b is pixel 2b + parity
                const int e = ch - p.split_chunk, par = e / p.spx_cpp, cc = e - par * p.spx_cpp;
                tma_load_4d_hint(a_base + (uint32_t)s * A_STAGE_BYTES, &tm_a1, afull_bar(s), cc * KC, 2 * wbase + (par & 1),
                                 2 * hbase + (par >> 1), t.img, pol1);
              } else
                tma_load_4d_hint(a_base + (uint32_t)s * A_STAGE_BYTES, &tm_a1, afull_bar(s), (ch - p.split_chunk) * KC, wbase,
                                 hbase, t.img, pol1);
            } else {
              mbar_arrive(afull_bar(s));
            }
            halo_trace(p, 17 + 2 * nstage);
            ++nstage;
            if (++s == p.a_stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else
    for (int tl = t_first; tl < t_end; tl += t_step) {
        const int tile = tile_of(tl);
      const HaloTile t = halo_decode<TG>(p, tile);
      const int hbase = t.h0 + p.dh_min, wbase = t.w0 + p.dw_min;
      for (int ch = 0; ch < p.chunks; ++ch) {
        const HaloSrc& S = p.src[ch < p.split_chunk ? 0 : 1];
        const int cbase = (ch < p.split_chunk ? ch : ch - p.split_chunk) * KC;
        const __nv_bfloat16* img = S.ptr + (long long)t.img * S.h * S.w * S.pitch + cbase;
        const int up = S.up, sw_ = S.w;
        const int pitch = (int)S.pitch;
        if (lane == 0) mbar_wait(aempty_bar(s), ph ^ 1u);
        __syncwarp();
        if (lt == 0) halo_trace(p, 16 + 2 * nstage);
        const uint32_t dst0 = a_base + (uint32_t)s * A_STAGE_BYTES;
        if (p.mix && ch >= p.split_chunk && !(UWM_DBG_OF(p) & 1)) {
          // the skip source is stored at the conv's resolution: one TMA box (swizzled layout) instead of a gather;
          // the barrier expects one arrival per loader thread, so thread 0's expect_tx arrival is one of them
          if (lt == 0) {
            mbar_arrive_expect_tx(afull_bar(s), (uint32_t)NPIX * ROWB);
            tma_load_4d(dst0, &tm_a1, afull_bar(s), (ch - p.split_chunk) * KC, wbase, hbase, t.img);
          } else {
            mbar_arrive(afull_bar(s));
          }
          if (lt == 0) halo_trace(p, 17 + 2 * nstage);
          ++nstage;
          if (++s == p.a_stages) { s = 0; ph ^= 1u; }
          continue;
        }
        if (!(UWM_DBG_OF(p) & 1)) {
#pragma unroll 4
          for (int i = lt; i < ITEMS; i += kHaloLoaderThreads) {
            const int c = i & (CPS - 1);
            const int px = i >> LOG2_CPS;
            const int hh = (int)(((uint32_t)px * p.pw_magic) >> 16);
            const int ww = px - hh * PW;
            const int ih = hbase + hh, iw = wbase + ww;
            const bool valid = ((unsigned)ih < (unsigned)p.h) && ((unsigned)iw < (unsigned)p.w);
            const int off = valid ? ((ih >> up) * sw_ + (iw >> up)) * pitch + c * 8 : 0;   // < 2^31: checked on the host
            cp_async_16_zfill(dst0 + (uint32_t)c * PLANE_BYTES + (uint32_t)px * 16u, img + off, valid ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive_noinc(afull_bar(s));
        if (lt == 0) halo_trace(p, 17 + 2 * nstage);
        ++nstage;
        if (++s == p.a_stages) { s = 0; ph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CHAIN && threadIdx.x == 0) {
    // the last CTA to finish (nobody polls any more) zeroes the counters for the next launch of this chain
    int* ticket = chain->dep + n_layers * p.n_img;
    __threadfence();
    if (atomicAdd(ticket, 1) == G - 1) {
      for (int i = 0; i < n_layers * p.n_img; ++i) chain->dep[i] = 0;
      ticket[0] = 0;
      ticket[1] = 0;                           // the work-item counter
      __threadfence();
    }
  }
  if (CG2) cluster_sync_all();       // the peer may still signal this CTA's barriers / the leader still reads its smem
  if (warp == 1) { if (CG2) tmem_dealloc_cg2(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols); }
  if (threadIdx.x == 0) halo_trace_cta(p, 1);
}

}  // namespace uwm
