// libuwm_b200.so — C ABI (include/uwm.h) over the sm_100a kernels in conv_tc.cuh / glue.cuh.
//
// Host side of the hot path: tensor-map construction, tile selection, the static plan of
// smp.Unet(resnet34|resnet50), the activation arena and CUDA-graph replay.
// libcuda is NOT linked: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint so
// the library still loads (and exports its symbols) on a machine without a GPU driver.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/uwm.h"
#include "conv_tc.cuh"
#include "conv_halo.cuh"
#include "glue.cuh"
#ifdef UWM_BENCH_TOOLS
#include "microbench.cuh"   // sizing micro-benchmarks: only in the tools build (python -m unet_watermark_b200.build --tools)
#endif

using namespace uwm;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(UWM_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                  __FILE__, __LINE__);                                                      \
  } while (0)

static bool debug_sync() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UWM_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static int post_launch(const char* what, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(UWM_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  if (debug_sync()) {
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(UWM_ECUDA, "%s faulted: %s", what, cudaGetErrorString(e));
  }
  return UWM_OK;
}

extern "C" const char* uwm_last_error(void) { return g_err; }
// shared with the library's other translation units (uwm_imgproc.cu); not part of the public header
extern "C" void uwm_internal_set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }
extern "C" void uwm_internal_count_launches(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
extern "C" int uwm_abi_version(void) { return 1; }
extern "C" uint64_t uwm_kernel_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------
// driver entry point (no libcuda link)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Launch with programmatic stream serialization (PDL): the kernel may start while its predecessor in the
// stream drains; every kernel here calls griddepcontrol.wait before touching dependent memory.
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UWM_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// Programmatic dependent launch lets a kernel start (prologue, weight and bias loads) before the previous kernel in the
// stream has finished; only activation reads / output writes sit behind griddepcontrol.wait.  That is sound inside a
// model plan, whose weights were uploaded long before.  The single-operator entry points keep plain stream order: a
// caller may have produced the weights or the bias on this stream one kernel earlier.
static thread_local int g_plan_depth = 0;
struct PlanScope { PlanScope() { ++g_plan_depth; } ~PlanScope() { --g_plan_depth; } };
// same, as 2-CTA clusters (CTA pairs of the cta_group::2 kernels)
template <typename... KArgs, typename... Args>
static void launch_pdl_pairs(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args);
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool op_pdl = []{ const char* e = getenv("UWM_OP_PDL"); return e && e[0] == '1'; }();   // bench tools: time single ops as the plan runs them
  cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && (g_plan_depth > 0 || op_pdl)) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
static void launch_pdl_pairs(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool op_pdl = []{ const char* e = getenv("UWM_OP_PDL"); return e && e[0] == '1'; }();
  cfg.attrs = attr; cfg.numAttrs = (pdl_enabled() && (g_plan_depth > 0 || op_pdl)) ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// per-device caches (a process may drive several GPUs through separate engines)
constexpr int kMaxDevices = 64;
// cached CUDA graphs per (model, batch): one per distinct (input, output, threshold) argument set; a caller that rotates
// 12 input buffers through 8 sub-batches (Engine.forward sub_batch) needs 96
constexpr size_t kMaxGraphsPerPlan = 512;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
static int num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

// ------------------------------------------------------------------------------------------
// conv launch construction
// ------------------------------------------------------------------------------------------
struct ConvSpec {
  const void* x = nullptr;        // activation base (already offset to the channel slice)
  int n = 0, h = 0, w = 0, cin = 0; // h, w: conv input size (the UPSAMPLED size when up1)
  long long x_pitch = 0;
  int up1 = 0;                    // x is stored at (h/2, w/2): nearest-2x upsample fused into the loader
  const void* x2 = nullptr;       // optional second source, concatenated after x along channels
  int cin2 = 0;
  long long x2_pitch = 0;
  const void* wgt = nullptr;      // [cout_pad][ntaps*cin]
  const float* bias = nullptr;
  int cout = 0, cout_pad = 0;
  int ntaps = 0;
  int8_t dh[kMaxTaps] = {0}, dw[kMaxTaps] = {0};
  int stride = 1;
  int h_out = 0, w_out = 0;
  const void* res = nullptr;
  long long res_pitch = 0;
  int relu = 0;
  void* out = nullptr;
  long long out_pitch = 0;
  int shuffle = 0;                // sub-pixel conv: real cout; `cout`/`cout_pad` are then 4x that (one group per output parity)
  int in_stride2 = 0, h_in = 0, w_in = 0;   // 1x1 stride-2 conv as a stride-1 conv on every second pixel of a [h_in, w_in] source
  int s2planes = 0;               // stride-2 3x3 conv with UWM_PACK_S2_PLANES weights: parity-plane halo kernel (build_halo_s2)
  int s2d = 0;                    // x (and out, or the head's 4 logits per block) are space-to-depth: [n,h,w,4*16], see conv_halo.cuh S2D
  int head = 0, apply_sigmoid = 0;
  float* logits = nullptr;
  uint8_t* mask = nullptr;
  float thr_logit = 0.f;
};

struct ConvLaunch {
  int halo = 0;                   // 0: conv_tc_kernel (args), 1: conv_halo_kernel (hargs)
  int kh = 0, kw = 0, kc = 0, tg = 0, resident = 0, a_tma = 0, spx = 0, s2d = 0, cg2 = 0;   // halo: template instantiation
  CUtensorMap tm_act, tm_wgt, tm_out, tm_res, tm_a0, tm_a1;
  ConvKArgs args;
  HaloKArgs hargs;
  unsigned grid = 0;
  size_t smem = 0;
};

static void choose_tile(int W, int H, int N, int stride, int* tw, int* th, int* tn) {
  double best = -1.0;
  for (int a = 128; a >= 1; a >>= 1) {
    if (a * stride > 256) continue;
    for (int b = 128 / a; b >= 1; b >>= 1) {
      if (b * stride > 256) continue;
      const int c = 128 / (a * b);
      if (c > 256) continue;
      const double cover = (double)((W + a - 1) / a * a) * ((H + b - 1) / b * b) * ((N + c - 1) / c * c);
      const double eff = (double)W * H * N / cover;
      // prefer full tiles, then wide rows (coalesced epilogue), then tall tiles over batch tiles
      const double score = eff + 1e-4 * std::log2((double)a) + 1e-6 * std::log2((double)b);
      if (score > best) { best = score; *tw = a; *th = b; *tn = c; }
    }
  }
}


static FastDiv make_fastdiv(int d) {
  FastDiv f; f.d = (uint32_t)d; f.mul = 0; f.shr = 0;
  if (d > 1) {
    int l = 0; while ((1LL << l) < d) ++l;          // ceil(log2 d)
    const int p = 31 + l;
    f.mul = (uint32_t)(((1ULL << p) + (uint64_t)d - 1) / (uint64_t)d);
    f.shr = (uint32_t)(p - 32);
  }
  return f;
}

// Pipeline-isolation switches (UWM_DBG) and the in-kernel trace corrupt results by design: they exist only in the tools
// build (-DUWM_BENCH_TOOLS); the product library ignores the variable and never carries a trace pointer.
#ifdef UWM_BENCH_TOOLS
static long long* g_halo_trace = nullptr;   // uwm_debug_set_trace
static int dbg_bits() { const char* e = getenv("UWM_DBG"); return e ? atoi(e) : 0; }
#else
static long long* const g_halo_trace = nullptr;
static int dbg_bits() { return 0; }
#endif

static bool halo_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UWM_HALO"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Sub-pixel form of the decoder's conv1 WITH a skip (conv_halo.cuh, SPX): conv3x3(concat(up2x(x), skip)) on x's grid.
//   s.h, s.w : x's (low) resolution; skip and out are [n, 2h, 2w, .]; s.shuffle = real cout, s.cout_pad = 4*cout
//   weights  : UWM_PACK_UPCAT_SUBPIXEL, [4*cout][9*cin/64 + 16*cin2/64 slices][64] in issue order
// Tensor-pipe cycles per 128 source pixels (= 512 outputs), cin = cin2 = 64, cout = 32: (9 + 16) x 4 MMAs of N = 128
// = 6400, against 4 x 72 MMAs of N = 32 = 11 800 for the gathered conv at the output resolution.
static int build_halo_spx(const ConvSpec& s, ConvLaunch* L) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const int kc = 64, bn = s.cout_pad;
  if (s.cin % kc || s.cin2 % kc || !s.cin || !s.cin2)
    return fail(UWM_EINVAL, "sub-pixel upcat conv: cin=%d+%d: each must be a non-zero multiple of 64", s.cin, s.cin2);
  if (s.shuffle % 16 || bn != 4 * s.shuffle || bn > 256)
    return fail(UWM_EINVAL, "sub-pixel upcat conv: cout=%d must be a multiple of 16, <= 64", s.shuffle);
  if ((s.x_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.x) & 15) || (s.x2_pitch * 2) % 16 ||
      (reinterpret_cast<uintptr_t>(s.x2) & 15) || (s.out_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.out) & 15))
    return fail(UWM_EINVAL, "sub-pixel upcat conv: activation/output base and pitch must be 16-byte aligned");
  if (s.res || s.head) return fail(UWM_EINVAL, "sub-pixel upcat conv: no residual / head epilogue");
  L->halo = 1;
  HaloKArgs& a = L->hargs;
  memset(&a, 0, sizeof(a));
  const int tg = (bn <= 128 && s.w >= 2 * kHaloTW) ? 2 : 1;          // two accumulator sets of tg*bn columns
  L->kh = 3; L->kw = 3; L->kc = kc; L->tg = tg; L->resident = 0; L->a_tma = 1; L->spx = 1;
  a.n_img = s.n; a.h = s.h; a.w = s.w;
  a.chunks = s.cin / kc + 4 * (s.cin2 / kc);
  a.split_chunk = s.cin / kc;
  a.spx_cpp = s.cin2 / kc;
  a.spx_slices = 9 * (s.cin / kc) + 16 * (s.cin2 / kc);
  a.dh_min = -1; a.dw_min = -1;
  a.a_scale = 1;
  a.cin_total = s.cin + s.cin2;
  a.tiles_w = (s.w + kHaloTW * tg - 1) / (kHaloTW * tg);
  a.tiles_h = (s.h + kHaloTH - 1) / kHaloTH;
  a.block_n = bn; a.n_tiles = 1; a.cout = bn;
  a.total_tiles = a.tiles_w * a.tiles_h * s.n;
  a.div_ntiles = make_fastdiv(1);
  a.div_tw = make_fastdiv(a.tiles_w);
  a.div_th = make_fastdiv(a.tiles_h);
  a.pw_magic = 65536u / (uint32_t)halo_pw(tg, 3) + 1u;
  const unsigned grid = (unsigned)std::min(a.total_tiles, num_sms());
  a.b_slice_bytes = (uint32_t)bn * kc * 2u;                           // >= 8 KB, 1024-aligned
  a.kpb = 1;
  const size_t a_stage_bytes = (((size_t)halo_npix(tg, 3, 3) * kc * 2) + 1023) & ~(size_t)1023;
  const size_t kBudget = 206u * 1024u;
  // A skip-plane stage lasts 4 taps = ~1750 tensor cycles, less than the 2500-3400 cycles its halo box takes from HBM
  // (the full-resolution skip tensor never sits in L2): with two stages every plane chunk waited ~800 cycles for its
  // box (profiles/r02_trace_dec3_spx.txt), so three stages when the weight ring keeps >= 4 slices (measured, 3 runs
  // each: config 2 -0.4 %, config 3 -1.1 %; four stages starve the weight ring, +2.7 %).
  a.a_stages = ((kBudget - 3 * a_stage_bytes) / a.b_slice_bytes >= 4) ? 3 : 2;
  { const char* e = getenv("UWM_SPX_ASTAGES"); if (e && atoi(e) >= 2 && atoi(e) <= 4) a.a_stages = atoi(e); }
  a.b_stages = (int)std::min<size_t>(12, (kBudget - a.a_stages * a_stage_bytes) / a.b_slice_bytes);
  if (a.b_stages < 2) return fail(UWM_ESTATE, "sub-pixel upcat conv: rings do not fit shared memory");
  a.nacc_log2 = 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * tg * bn)) cols <<= 1;
  a.tmem_cols = cols;
  a.bias = s.bias;
  a.out = static_cast<__nv_bfloat16*>(s.out);
  a.out_pitch = s.out_pitch;
  a.relu = s.relu;
  a.shuffle = s.shuffle;
  a.ep_tma = 0; a.ep_cols = 16;
  a.dbg = dbg_bits();
  a.trace = g_halo_trace;
  a.src[0].ptr = static_cast<const __nv_bfloat16*>(s.x); a.src[0].pitch = s.x_pitch; a.src[0].h = s.h; a.src[0].w = s.w;
  a.src[1].ptr = static_cast<const __nv_bfloat16*>(s.x2); a.src[1].pitch = s.x2_pitch; a.src[1].h = 2 * s.h; a.src[1].w = 2 * s.w;
  { const char* e = getenv("UWM_VERBOSE");
    if (e && e[0] == '1')
      fprintf(stderr, "halo conv (sub-pixel upcat) %dx%dx%d cin=%d+%d cout=4x%d: bn=%d tg=%d tiles=%d grid=%u a_stages=%d b_stages=%d slices=%d\n",
              s.n, s.h, s.w, s.cin, s.cin2, s.shuffle, bn, tg, a.total_tiles, grid, a.a_stages, a.b_stages, a.spx_slices); }
  L->grid = grid;
  L->smem = 1024 + (size_t)a.b_stages * a.b_slice_bytes + (size_t)a.a_stages * a_stage_bytes + 1024 + (size_t)bn * 4;

  const cuuint64_t ktot = (cuuint64_t)a.spx_slices * kc;
  cuuint64_t dims[2] = {ktot, (cuuint64_t)bn};
  cuuint64_t strides[1] = {ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(&L->tm_wgt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(s.wgt), dims, strides, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(sub-pixel wgt) -> %d", (int)r);
  // taps that reach one output row parity / one output parity only load that half / quarter of the rows:
  // the kernel takes those two maps in the (otherwise unused) output and residual map slots
  for (int q = 2; q <= 4; q <<= 1) {
    cuuint32_t qbox[2] = {(cuuint32_t)kc, (cuuint32_t)(bn / q)};
    r = enc(q == 2 ? &L->tm_out : &L->tm_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(s.wgt), dims, strides,
            qbox, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(sub-pixel wgt, 1/%d rows) -> %d", q, (int)r);
  }
  const cuuint32_t pw = (cuuint32_t)halo_pw(tg, 3), ph = (cuuint32_t)(kHaloTH + 2);
  {   // x at its own resolution: the ordinary halo box
    cuuint64_t adims[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.w, (cuuint64_t)s.h, (cuuint64_t)s.n};
    cuuint64_t astr[3] = {(cuuint64_t)s.x_pitch * 2, (cuuint64_t)s.w * s.x_pitch * 2, (cuuint64_t)s.h * s.w * s.x_pitch * 2};
    cuuint32_t abox[4] = {(cuuint32_t)kc, pw, ph, 1};
    cuuint32_t aest[4] = {1, 1, 1, 1};
    r = enc(&L->tm_a0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.x), adims, astr, abox, aest,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(sub-pixel x) -> %d", (int)r);
  }
  {   // skip at twice the resolution, every second pixel of every second row: one parity plane per box
    cuuint64_t adims[4] = {(cuuint64_t)s.cin2, (cuuint64_t)2 * s.w, (cuuint64_t)2 * s.h, (cuuint64_t)s.n};
    cuuint64_t astr[3] = {(cuuint64_t)s.x2_pitch * 2, (cuuint64_t)2 * s.w * s.x2_pitch * 2,
                          (cuuint64_t)4 * s.h * s.w * s.x2_pitch * 2};
    cuuint32_t abox[4] = {(cuuint32_t)kc, 2 * pw, 2 * ph, 1};
    cuuint32_t aest[4] = {1, 2, 2, 1};
    r = enc(&L->tm_a1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.x2), adims, astr, abox, aest,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(sub-pixel skip, element stride 2) -> %d (box %u,%u,%u)", (int)r, abox[0],
                  abox[1], abox[2]);
  }
  return UWM_OK;
}

// Stride-2 3x3 conv (pad 1) on the halo kernel (conv_halo.cuh, SPX == 2): the input's four parity planes arrive as
// dense TMA boxes (element stride 2), plane (ph,pw) meets the 1, 2, 2 or 4 halo taps that kernel taps (kr,kc) with
// kr odd/even = ph, kc likewise map to.  s.h, s.w: INPUT size (even); output [n, h/2, w/2, cout].
//   weights: UWM_PACK_S2_PLANES, [cout][9 * cin] as 64-channel slices in issue order
static int build_halo_s2(const ConvSpec& s, ConvLaunch* L) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const int kc = 64;
  if (s.cin % kc || s.cout_pad % 64 || s.cout != s.cout_pad || (s.h & 1) || (s.w & 1) || s.x2 || s.up1 || s.res || s.head)
    return fail(UWM_EINVAL, "stride-2 plane conv: needs cin%%64 == 0, cout%%64 == 0, even input size, one source, no residual");
  if ((s.x_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.x) & 15) || (s.out_pitch * 2) % 16 ||
      (reinterpret_cast<uintptr_t>(s.out) & 15))
    return fail(UWM_EINVAL, "stride-2 plane conv: activation/output base and pitch must be 16-byte aligned");
  const int ho = s.h / 2, wo = s.w / 2;
  const int bn = (s.cout_pad % 128 == 0) ? 128 : 64;
  const int n_tiles = s.cout_pad / bn;
  const int sms = num_sms();
  auto m_tiles_of = [&](int tg) { return ((wo + kHaloTW * tg - 1) / (kHaloTW * tg)) * ((ho + kHaloTH - 1) / kHaloTH) * s.n; };
  const int tg = (wo >= 2 * kHaloTW && m_tiles_of(2) * n_tiles * 5 >= sms * 4) ? 2 : 1;   // two sub-tiles halve the weight traffic
  L->halo = 1;
  HaloKArgs& a = L->hargs;
  memset(&a, 0, sizeof(a));
  L->kh = 2; L->kw = 2; L->kc = kc; L->tg = tg; L->resident = 0; L->a_tma = 1; L->spx = 2;
  a.n_img = s.n; a.h = ho; a.w = wo;
  a.spx_cpp = s.cin / kc;
  a.chunks = 4 * a.spx_cpp;
  a.split_chunk = 0;
  a.spx_slices = 9 * a.spx_cpp;
  a.dh_min = -1; a.dw_min = -1;
  a.a_scale = 1;
  a.cin_total = s.cin;
  a.tiles_w = (wo + kHaloTW * tg - 1) / (kHaloTW * tg);
  a.tiles_h = (ho + kHaloTH - 1) / kHaloTH;
  a.block_n = bn; a.n_tiles = n_tiles; a.cout = s.cout;
  a.total_tiles = a.tiles_w * a.tiles_h * s.n * n_tiles;
  a.div_ntiles = make_fastdiv(n_tiles);
  a.div_tw = make_fastdiv(a.tiles_w);
  a.div_th = make_fastdiv(a.tiles_h);
  a.pw_magic = 65536u / (uint32_t)halo_pw(tg, 2) + 1u;
  const unsigned grid = (unsigned)std::min(a.total_tiles, sms);
  a.b_slice_bytes = (uint32_t)bn * kc * 2u;
  a.kpb = 1;
  const size_t stg_bytes = (size_t)8 * 32 * 128 + 1024;
  const size_t a_stage_bytes = (((size_t)halo_npix(tg, 2, 2) * kc * 2) + 1023) & ~(size_t)1023;
  const size_t kBudget = 206u * 1024u - stg_bytes;
  // a plane stage lasts 1-4 taps x 4 K-steps; weight slices turn over every ~512 cycles and must cover the L2 latency
  a.a_stages = (tg == 2) ? 2 : 3;
  a.b_stages = (int)std::min<size_t>(12, (kBudget - a.a_stages * a_stage_bytes) / a.b_slice_bytes);
  if (a.b_stages < 2) return fail(UWM_ESTATE, "stride-2 plane conv: rings do not fit shared memory");
  a.nacc_log2 = (4 * tg * bn <= 512) ? 2 : 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)((1 << a.nacc_log2) * tg * bn)) cols <<= 1;
  a.tmem_cols = cols;
  a.bias = s.bias;
  a.out = static_cast<__nv_bfloat16*>(s.out);
  a.out_pitch = s.out_pitch;
  a.relu = s.relu;
  a.ep_tma = 1; a.ep_cols = 64;
  a.dbg = dbg_bits();
  a.trace = g_halo_trace;
  a.src[0].ptr = static_cast<const __nv_bfloat16*>(s.x); a.src[0].pitch = s.x_pitch; a.src[0].h = s.h; a.src[0].w = s.w;
  a.src[1] = a.src[0];
  { const char* e = getenv("UWM_VERBOSE");
    if (e && e[0] == '1')
      fprintf(stderr, "halo conv (stride-2 planes) %dx%dx%d cin=%d cout=%d: bn=%d tg=%d tiles=%d grid=%u a_stages=%d b_stages=%d slices=%d\n",
              s.n, s.h, s.w, s.cin, s.cout, bn, tg, a.total_tiles, grid, a.a_stages, a.b_stages, a.spx_slices); }
  L->grid = grid;
  L->smem = 1024 + (size_t)a.b_stages * a.b_slice_bytes + (size_t)a.a_stages * a_stage_bytes + stg_bytes + 1024 +
            (size_t)s.cout_pad * 4;

  const cuuint64_t ktot = (cuuint64_t)9 * s.cin;
  cuuint64_t dims[2] = {ktot, (cuuint64_t)s.cout_pad};
  cuuint64_t strides[1] = {ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(&L->tm_wgt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(s.wgt), dims, strides, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(stride-2 wgt) -> %d", (int)r);
  {
    cuuint64_t odims[4] = {(cuuint64_t)s.cout, (cuuint64_t)wo, (cuuint64_t)ho, (cuuint64_t)s.n};
    cuuint64_t ostr[3] = {(cuuint64_t)s.out_pitch * 2, (cuuint64_t)wo * s.out_pitch * 2, (cuuint64_t)ho * wo * s.out_pitch * 2};
    cuuint32_t obox[4] = {64, (cuuint32_t)kHaloTW, 4, 1};
    cuuint32_t oest[4] = {1, 1, 1, 1};
    r = enc(&L->tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, s.out, odims, ostr, obox, oest, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(stride-2 out) -> %d", (int)r);
    L->tm_res = L->tm_out;
  }
  {   // every second pixel of every second row of the input: one parity plane per box
    cuuint64_t adims[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.w, (cuuint64_t)s.h, (cuuint64_t)s.n};
    cuuint64_t astr[3] = {(cuuint64_t)s.x_pitch * 2, (cuuint64_t)s.w * s.x_pitch * 2, (cuuint64_t)s.h * s.w * s.x_pitch * 2};
    cuuint32_t abox[4] = {(cuuint32_t)kc, 2 * (cuuint32_t)halo_pw(tg, 2), 2 * (cuuint32_t)(kHaloTH + 1), 1};
    cuuint32_t aest[4] = {1, 2, 2, 1};
    r = enc(&L->tm_a1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.x), adims, astr, abox, aest,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(stride-2 planes) -> %d", (int)r);
    L->tm_a0 = L->tm_a1;
  }
  return UWM_OK;
}

// Halo-resident kernel (conv_halo.cuh): stride-1 convs, optional fused upsample + concat.
static int build_halo(const ConvSpec& s, ConvLaunch* L) {
  if (s.shuffle && s.x2) return build_halo_spx(s, L);
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if (s.stride != 1) return fail(UWM_EINVAL, "halo conv: stride %d unsupported", s.stride);
  if (s.cin % 16 || s.cin2 % 16) return fail(UWM_EINVAL, "halo conv: cin=%d+%d: each must be a multiple of 16", s.cin, s.cin2);
  if (s.cout_pad % 16) return fail(UWM_EINVAL, "halo conv: cout_pad=%d must be a multiple of 16", s.cout_pad);
  if (s.cout_pad > 2048) return fail(UWM_EINVAL, "halo conv: cout=%d > 2048 (bias staging)", s.cout_pad);
  if (s.ntaps < 1 || s.ntaps > kMaxTaps) return fail(UWM_EINVAL, "halo conv: %d taps unsupported", s.ntaps);
  if (s.h_out != s.h || s.w_out != s.w) return fail(UWM_EINVAL, "halo conv: needs 'same' padding (%dx%d -> %dx%d)", s.h, s.w, s.h_out, s.w_out);
  if (s.up1 && ((s.h | s.w) & 1)) return fail(UWM_EINVAL, "halo conv: upsampled size %dx%d must be even", s.h, s.w);
  if ((s.x_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.x) & 15) ||
      (s.x2 && ((s.x2_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.x2) & 15))))
    return fail(UWM_EINVAL, "halo conv: activation base/pitch must be 16-byte aligned");
  if (!s.head && ((s.out_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.out) & 15)))
    return fail(UWM_EINVAL, "halo conv: output base/pitch must be 16-byte aligned");
  if (s.res && ((s.res_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.res) & 15)))
    return fail(UWM_EINVAL, "halo conv: residual base/pitch must be 16-byte aligned");

  if (s.s2d && (s.cin != 64 || s.x2 || s.up1 || s.ntaps != 9 || s.res || s.shuffle || (s.head ? s.cout_pad != 16 : s.cout_pad != 64)))
    return fail(UWM_EINVAL, "halo conv: space-to-depth form needs 4x16 input channels, one source, 3x3, 4x16 outputs (or the head)");
  L->halo = 1;
  L->s2d = s.s2d ? 1 : 0;
  HaloKArgs& a = L->hargs;
  memset(&a, 0, sizeof(a));
  a.n_img = s.n; a.h = s.h; a.w = s.w;
  const int cin_total = s.cin + (s.x2 ? s.cin2 : 0);
  int kc = 64;
  while (kc > 16 && (s.cin % kc || (s.x2 && s.cin2 % kc))) kc >>= 1;
  a.chunks = cin_total / kc;
  a.split_chunk = s.cin / kc;
  const int cps = kc / 8;
  // taps must be the full kh x kw rectangle in row-major order (taps_rect / the stem's 4x4)
  int dh_min = 127, dh_max = -128, dw_min = 127, dw_max = -128;
  for (int t = 0; t < s.ntaps; ++t) {
    dh_min = std::min(dh_min, (int)s.dh[t]); dh_max = std::max(dh_max, (int)s.dh[t]);
    dw_min = std::min(dw_min, (int)s.dw[t]); dw_max = std::max(dw_max, (int)s.dw[t]);
  }
  const int kh = dh_max - dh_min + 1, kw = dw_max - dw_min + 1;
  if (kh * kw != s.ntaps) return fail(UWM_EINVAL, "halo conv: taps are not a full %dx%d rectangle", kh, kw);
  for (int t = 0; t < s.ntaps; ++t)
    if (s.dh[t] != dh_min + t / kw || s.dw[t] != dw_min + t % kw)
      return fail(UWM_EINVAL, "halo conv: taps must be in row-major order");
  if (!((kh == 3 && kw == 3) || (kh == 1 && kw == 1) || (kh == 4 && kw == 4 && kc == 16)))
    return fail(UWM_EINVAL, "halo conv: %dx%d filters with %d-channel chunks are not instantiated", kh, kw, kc);
  a.dh_min = dh_min; a.dw_min = dw_min;
  a.a_scale = s.in_stride2 ? 2 : 1;
  a.src[0].ptr = static_cast<const __nv_bfloat16*>(s.x);
  a.src[0].pitch = s.x_pitch; a.src[0].up = s.up1 ? 1 : 0;
  a.src[0].h = s.up1 ? s.h / 2 : s.h; a.src[0].w = s.up1 ? s.w / 2 : s.w;
  a.src[1].ptr = static_cast<const __nv_bfloat16*>(s.x2 ? s.x2 : s.x);
  a.src[1].pitch = s.x2 ? s.x2_pitch : s.x_pitch; a.src[1].up = 0; a.src[1].h = s.h; a.src[1].w = s.w;
  for (int i = 0; i < 2; ++i)
    if ((long long)a.src[i].h * a.src[i].w * a.src[i].pitch >= (1LL << 31))
      return fail(UWM_EINVAL, "halo conv: one image of a source exceeds 2^31 elements");
  a.cin_total = cin_total;
  const int nk = s.ntaps * a.chunks;

  // ---- tile selection: N tile (bn) x sub-tiles per tile (tg), by a small cost model (cycles) ----------
  //   issue  : the one MMA-issuing thread: ~60 instr per stage + ~12 per tap + ~3 per MMA, ~6 clk each (ncu)
  //   tensor : tg * nk * (kc/16) MMAs of 32 + bn/4 clk (bn <= 128; 128 clk at bn = 256), measured
  //   L2     : halo(tg)*cin*2 + (weights streamed ? bn*K*2 : 0) bytes per tile; ~64 B/clk per SM, 6300 B/clk chip
  // epilogue staging for TMA tensor stores (8 warps x 32 rows x 128 B): outputs with a multiple of 64 channels
  static const bool ep_tma_enabled = []{ const char* e = getenv("UWM_EP_TMA"); return !(e && e[0] == '0'); }();
  const bool par_tiles_early = s.shuffle && !s.x2 && s.cout_pad > 256;      // sub-pixel conv, one N tile per parity
  const bool ep_tma = ep_tma_enabled && !s.head && (!s.shuffle || par_tiles_early) && (s.cout_pad % 64 == 0);
  const size_t stg_bytes = ep_tma ? (size_t)8 * 32 * 128 + 1024 : 0;
  const size_t kBudget = 206u * 1024u - stg_bytes; // rings + resident weights (barriers/alignment slack on top)
  const int sms = num_sms();
  // activation halo by TMA (swizzled pixel-major stage) unless a source is upsampled on the fly
  static const bool a_tma_enabled = []{ const char* e = getenv("UWM_A_TMA"); return !(e && e[0] == '0'); }();
  const bool a_tma = a_tma_enabled && !s.up1;
  auto a_stage_of = [&](int tg) -> size_t {
    const size_t sw = (((size_t)halo_npix(tg, kh, kw) * kc * 2) + 1023) & ~(size_t)1023;
    if (a_tma) return sw;
    const size_t pl = ((size_t)cps * halo_plane_bytes(tg, kh, kw) + 1023) & ~(size_t)1023;
    return std::max(sw, pl);          // a gathered stage (planes) or a TMA box of the skip source (swizzled rows)
  };
  // sub-pixel conv with more than 256 GEMM columns: one N tile per output parity, 4 of the 9 taps each (SPX == 3)
  const bool par_tiles = s.shuffle && !s.x2 && s.cout_pad > 256;
  if (par_tiles && (kc != 64 || (s.shuffle != 128 && s.shuffle != 256) || s.cout_pad != 4 * s.shuffle || !a_tma))
    return fail(UWM_EINVAL, "sub-pixel conv: cout=%d needs cout in {128, 256} and cin a multiple of 64", s.shuffle);
  if (s.shuffle && s.res && !par_tiles) return fail(UWM_EINVAL, "sub-pixel conv: residual only with cout in {128, 256}");
  struct Cand { int bn, tg; bool resident; double cost; } best = {0, 0, false, 1e30};
  std::vector<int> bns;
  if (par_tiles) bns.push_back(s.shuffle);
  else if (s.cout_pad <= 64) bns.push_back(s.cout_pad);
  else for (int c : {256, 128, 64}) if (s.cout_pad % c == 0) bns.push_back(c);
  if (bns.empty()) { int b = std::min(s.cout_pad, 256); while (s.cout_pad % b) b -= 16; bns.push_back(b); }
  const int w8 = (s.w + kHaloTW - 1) / kHaloTW;
  const int force_tg = []{ const char* e = getenv("UWM_HALO_TG"); return e ? atoi(e) : 0; }();
  const int force_bn = []{ const char* e = getenv("UWM_HALO_BN"); return e ? atoi(e) : 0; }();
  for (int bn : bns) {
    if (force_bn && bn != force_bn && s.cout_pad % force_bn == 0 && s.cout_pad > 64) continue;
    for (int tg = 1; tg <= 8; tg <<= 1) {
      if (2 * tg * bn > 512) break;                               // two TMEM accumulator sets
      if (tg > 1 && tg > w8) break;                               // tile wider than the image
      if (force_tg && tg != force_tg && force_tg <= w8 && 2 * force_tg * bn <= 512) continue;
      const size_t a_stage = a_stage_of(tg);
      const size_t b_slice = ((size_t)bn * kc * 2 + 1023) & ~(size_t)1023;
      const int n_tiles = s.cout_pad / bn;
      // S2D kernels keep the 16 [bn x 16] weight blocks a parity plane can meet instead of 9 x 4 (conv_halo.cuh)
      const size_t res_bytes = s.s2d ? (((size_t)16 * bn * 32 + 1023) & ~(size_t)1023) : (size_t)nk * b_slice;
      const bool resident = (n_tiles == 1) && !par_tiles && (res_bytes + 2 * a_stage <= kBudget);
      if (!resident && (tg > 2 || 2 * b_slice + 2 * a_stage > kBudget)) break;   // streamed kernels: TG 1, 2
      if (s.s2d && (!resident || tg > 4 || !a_tma)) break;         // S2D: resident weights, TG 1, 2, 4, TMA-fed
      if (s.s2d && s.head && tg > 2) break;                        // head: 79 KB stages at TG 4 leave a 2-deep ring (measured slower)
      if (kh == 4 && !resident) break;
      if (kh == 1 && (tg > 4 || !a_tma)) break;                    // 1x1: TG 1, 2, 4, TMA-fed only
      const long long tiles = (long long)((s.w + kHaloTW * tg - 1) / (kHaloTW * tg)) * ((s.h + kHaloTH - 1) / kHaloTH) * s.n * n_tiles;
      const long long waves = (tiles + sms - 1) / sms;
      const double mmas = (double)tg * nk * (kc / 16) * (s.s2d ? 16.0 / 36.0 : 1.0);   // S2D skips the (tap, plane) pairs that cannot meet
      const double tensor = mmas * (bn >= 256 ? 128.0 : 32.0 + bn / 4.0);   // measured: operand fetch (128+N)x32 B at 128 B/clk
      const double issue = 6.0 * (60.0 * a.chunks + 12.0 * nk + 3.0 * mmas);
      const double a_bytes = (double)halo_npix(tg, kh, kw) * cin_total * 2;
      const double b_bytes = resident ? 0.0 : (double)bn * nk * kc * 2;
      const int a_stages_fit = (int)((kBudget - (resident ? res_bytes : 2 * b_slice)) / a_stage);
      const double shallow = a_stages_fit < 3 ? 1.08 : 1.0;        // two stages hide the load latency a little less well (measured)
      // + ~450 cycles per tile of barrier hand-offs that nothing overlaps (trace: 2220-cycle cadence on 1764 cycles of MMAs)
      const double t_sm = waves * (std::max(std::max(tensor, issue), (a_bytes + b_bytes) / 64.0) * shallow + 450.0);
      const double t_l2 = tiles * (a_bytes + b_bytes) / 6300.0;
      const double cost = std::max(t_sm, t_l2);
      if (cost < best.cost * 0.97) best = {bn, tg, resident, cost};
    }
  }
  if (!best.bn) return fail(UWM_EINVAL, "halo conv: no tile fits shared memory (cin=%d cout=%d)", cin_total, s.cout_pad);
  const int bn = best.bn, tg = best.tg;
  L->kh = kh; L->kw = kw; L->kc = kc; L->tg = tg; L->resident = best.resident ? 1 : 0; L->a_tma = a_tma ? 1 : 0;
  if (par_tiles) L->spx = 3;
  a.tiles_w = (s.w + kHaloTW * tg - 1) / (kHaloTW * tg);
  a.tiles_h = (s.h + kHaloTH - 1) / kHaloTH;
  const int m_tiles = a.tiles_w * a.tiles_h * s.n;
  const int pw = halo_pw(tg, kw), npix = halo_npix(tg, kh, kw);
  a.pw_magic = 65536u / (uint32_t)pw + 1u;
  for (int p = 0; p < npix; ++p)
    if ((int)(((uint32_t)p * a.pw_magic) >> 16) != p / pw) return fail(UWM_ESTATE, "halo conv: division magic failed for pw=%d", pw);
  const size_t a_stage_bytes = a_stage_of(tg);
  a.block_n = bn;
  a.n_tiles = s.cout_pad / bn;
  a.cout = s.cout;
  a.total_tiles = m_tiles * a.n_tiles;
  a.div_ntiles = make_fastdiv(a.n_tiles);
  a.div_tw = make_fastdiv(a.tiles_w);
  a.div_th = make_fastdiv(a.tiles_h);
  a.div_total = make_fastdiv(a.total_tiles);
  // CTA pairs (cta_group::2) for the streamed 3x3 kernels fed by TMA from one source: two M tiles of one N tile share
  // every weight slice, each CTA loading half of its rows
  // Measured (r34 512x512 B=16): it pays where a tile's streamed bytes outrun the fabric's ~43 B/clk per SM - layer4
  // (1.18 MB of weights per 18.4 k tensor cycles: 20.3 -> 18.3 us) - and costs ~7 % on layer2/3 (32 B/clk; the pair's
  // TMEM allocation and commit round trips add ~2.5 k cycles per launch and the MMA rate does not improve).
  // UWM_CG2=0: never, 2: wherever eligible.
  static const int cg2_mode = []{ const char* e = getenv("UWM_CG2"); return e ? atoi(e) : 1; }();
  const double cg2_bytes_per_clk = ((double)halo_npix(tg, kh, kw) * cin_total * 2 + (double)bn * nk * kc * 2) /
                                   ((double)tg * nk * (kc / 16) * (bn >= 256 ? 128.0 : 32.0 + bn / 4.0));
  const bool cg2 = cg2_mode > 0 && (cg2_mode == 2 || cg2_bytes_per_clk > 50.0) && !best.resident && a_tma && !s.x2 && !par_tiles &&
                   !s.s2d && !s.head && !s.shuffle && !s.in_stride2 && kc == 64 && kh == 3 && kw == 3 &&
                   (m_tiles % 2 == 0) && bn % 32 == 0 && a.total_tiles >= 2;
  L->cg2 = cg2 ? 1 : 0;
  unsigned grid = (unsigned)std::min(a.total_tiles, sms);
  if (cg2) grid &= ~1u;
  a.b_slice_bytes = (((uint32_t)bn >> (cg2 ? 1 : 0)) * kc * 2u + 1023u) & ~1023u;    // a pair holds half of the rows per CTA
  const size_t resident_bytes = s.s2d ? (((size_t)16 * bn * 32 + 1023) & ~(size_t)1023) : (size_t)nk * a.b_slice_bytes;
  if (best.resident) {
    a.kpb = 1; a.b_stages = 1;
    a.a_stages = (int)std::min<size_t>(kHaloMaxStages, (kBudget - resident_bytes) / a_stage_bytes);
  } else {
    // taps per weight stage: enough MMA work per barrier round trip (>= ~512 cycles) if it still fits 3 deep
    const int mma_cycles = tg * (kc / 16) * std::max(32, bn / 2);
    // an activation stage lasts all taps of a chunk (thousands of cycles): two are enough; the short-lived
    // weight stages get the rest of shared memory so the ring covers the L2 latency under load
    const size_t a_min = 2 * a_stage_bytes;
    int kpb = 1;
    for (int d = 1; d <= s.ntaps; ++d) {
      if (s.ntaps % d) continue;
      if (3 * (size_t)d * a.b_slice_bytes + a_min > kBudget) break;
      kpb = d;
      if (d * mma_cycles >= 512) break;
    }
    if (par_tiles) kpb = 1;                       // the N tile's four taps are not contiguous in K
    a.kpb = kpb;
    const size_t b_stage = (size_t)kpb * a.b_slice_bytes;
    const size_t a_keep = std::min(a_min, kBudget - 2 * b_stage);
    a.b_stages = (int)std::max<size_t>(2, std::min<size_t>(12, (kBudget - a_keep) / b_stage));
    a.a_stages = (int)std::max<size_t>(2, std::min<size_t>(kHaloMaxStages, (kBudget - a.b_stages * b_stage) / a_stage_bytes));
  }
  { const char* e = getenv("UWM_HALO_ASTAGES"); if (e && atoi(e) >= 2) a.a_stages = std::min(a.a_stages, atoi(e)); }
  a.nacc_log2 = (4 * tg * bn <= 512) ? 2 : 1;
  { const char* e = getenv("UWM_NACC"); if (e && atoi(e) == 2) a.nacc_log2 = 1; }
  uint32_t cols = 32;
  while (cols < (uint32_t)((1 << a.nacc_log2) * tg * bn)) cols <<= 1;
  a.tmem_cols = cols;
  a.bias = s.bias;
  a.res = static_cast<const __nv_bfloat16*>(s.res);
  a.out = static_cast<__nv_bfloat16*>(s.out);
  a.res_pitch = s.res_pitch; a.out_pitch = s.out_pitch;
  a.relu = s.relu;
  a.head = s.head; a.apply_sigmoid = s.apply_sigmoid;
  a.logits = s.logits; a.mask = s.mask; a.thr_logit = s.thr_logit;
  a.dbg = dbg_bits();
  a.trace = g_halo_trace;
  { static const bool mix_on = []{ const char* e = getenv("UWM_MIX"); return !(e && e[0] == '0'); }();
    a.mix = (!a_tma && s.x2 && mix_on) ? 1 : 0; }
#ifdef UWM_BENCH_TOOLS
  { const char* e = getenv("UWM_TRACE_KH"); if (e && atoi(e) != kh) a.trace = nullptr; }   // trace one filter shape
#endif
  a.shuffle = (par_tiles && ep_tma) ? 0 : s.shuffle;      // the strided TMA view does the pixel shuffle
  a.ep_tma = ep_tma ? 1 : 0; a.ep_cols = ep_tma ? 64 : 16;
  { const char* e = getenv("UWM_VERBOSE");
    if (e && e[0] == '1')
      fprintf(stderr, "halo conv %dx%dx%d cin=%d(+%d%s) cout=%d %dx%d: bn=%d tg=%d kc=%d chunks=%d tiles=%d grid=%u a_stages=%d (%zu B) %s kpb=%d b_stages=%d\n",
              s.n, s.h, s.w, s.cin, s.cin2, s.up1 ? ",up" : "", s.cout, kh, kw, bn, tg, kc, a.chunks, a.total_tiles, grid,
              a.a_stages, a_stage_bytes, best.resident ? "B resident" : (cg2 ? "B streamed, CTA pairs" : "B streamed"), a.kpb, a.b_stages); }
  L->grid = grid;
  const size_t b_total = best.resident ? resident_bytes : (size_t)a.b_stages * a.kpb * a.b_slice_bytes;
  L->smem = 1024 + b_total + (size_t)a.a_stages * a_stage_bytes + stg_bytes + 1024 + (size_t)s.cout_pad * 4;

  const CUtensorMapSwizzle sw = (kc == 64) ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (kc == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                           : CU_TENSOR_MAP_SWIZZLE_32B;
  const cuuint64_t ktot = (cuuint64_t)s.ntaps * cin_total;
  cuuint64_t dims[2] = {ktot, (cuuint64_t)s.cout_pad};
  cuuint64_t strides[1] = {ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(cg2 ? bn / 2 : bn)};
  cuuint32_t est[2] = {1, 1};
  CUresult r = enc(&L->tm_wgt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(s.wgt), dims, strides, box,
                   est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(wgt) -> %d (k=%llu cout=%d)", (int)r, (unsigned long long)ktot, s.cout_pad);
  if (ep_tma) {
    // output view [n][h][w][cout] (pixel pitch out_pitch): boxes of 64 channels x 8 x 4 pixels, 128B-swizzled in smem
    // par_tiles: the output is [n][2h][2w][shuffle]; a box covers every second pixel of every second row (one parity)
    const int osc = par_tiles ? 2 : 1;
    const cuuint64_t ow_ = (cuuint64_t)s.w * osc, oh_ = (cuuint64_t)s.h * osc;
    cuuint64_t odims[4] = {(cuuint64_t)(par_tiles ? s.shuffle : s.cout), ow_, oh_, (cuuint64_t)s.n};
    cuuint64_t ostr[3] = {(cuuint64_t)s.out_pitch * 2, ow_ * s.out_pitch * 2, oh_ * ow_ * s.out_pitch * 2};
    cuuint32_t obox[4] = {64, (cuuint32_t)(kHaloTW * osc), (cuuint32_t)(4 * osc), 1};
    cuuint32_t oest[4] = {1, (cuuint32_t)osc, (cuuint32_t)osc, 1};
    r = enc(&L->tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, s.out, odims, ostr, obox, oest, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(out) -> %d (c=%d w=%d h=%d n=%d pitch=%lld)", (int)r, s.cout, s.w, s.h,
                  s.n, s.out_pitch);
    L->tm_res = L->tm_out;
    if (s.res) {              // residual view with the same boxes (added in place in the staging buffer)
      cuuint64_t rstr[3] = {(cuuint64_t)s.res_pitch * 2, ow_ * s.res_pitch * 2, oh_ * ow_ * s.res_pitch * 2};
      r = enc(&L->tm_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.res), odims, rstr, obox, oest,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(res) -> %d", (int)r);
    }
  } else {
    L->tm_out = L->tm_wgt;    // unused by the kernel
    L->tm_res = L->tm_wgt;
  }
  L->tm_a0 = L->tm_wgt; L->tm_a1 = L->tm_wgt;
  if (s.s2d) {
    // the same weights in [bn x 16] boxes (32-byte rows, 32B swizzle): one box per (tap, parity plane) pair that can meet
    cuuint32_t box16[2] = {16, (cuuint32_t)bn};
    r = enc(&L->tm_a1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(s.wgt), dims, strides, box16, est,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(s2d wgt) -> %d", (int)r);
  }
  if (a_tma || s.x2) {                  // !a_tma with a skip source: only the skip source's box map is used (mix)
    for (int i = a_tma ? 0 : 1; i < 2; ++i) {
      if (i == 1 && !s.x2) { if (!s.s2d) L->tm_a1 = L->tm_a0; break; }
      const void* base = i ? s.x2 : s.x;
      const long long pitch = i ? s.x2_pitch : s.x_pitch;
      const int csrc = i ? s.cin2 : s.cin;
      const int sc = s.in_stride2 ? 2 : 1;          // strided view: the box spans 2x the pixels, every second one lands
      const int hs = s.in_stride2 ? s.h_in : s.h, ws = s.in_stride2 ? s.w_in : s.w;
      cuuint64_t adims[4] = {(cuuint64_t)csrc, (cuuint64_t)ws, (cuuint64_t)hs, (cuuint64_t)s.n};
      cuuint64_t astr[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)ws * pitch * 2, (cuuint64_t)hs * ws * pitch * 2};
      cuuint32_t abox[4] = {(cuuint32_t)kc, (cuuint32_t)(sc * halo_pw(tg, kw)), (cuuint32_t)(sc * (kHaloTH + kh - 1)), 1};
      cuuint32_t aest[4] = {1, (cuuint32_t)sc, (cuuint32_t)sc, 1};
      r = enc(i ? &L->tm_a1 : &L->tm_a0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), adims, astr, abox,
              aest, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS)
        return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(act %d) -> %d (c=%d w=%d h=%d n=%d pitch=%lld box=%u,%u,%u)", i, (int)r,
                    csrc, s.w, s.h, s.n, pitch, abox[0], abox[1], abox[2]);
    }
  }
  return UWM_OK;
}

static int build_conv_stream(const ConvSpec& s, ConvLaunch* L);

// Kernel selection: k x k stride-1 'same' convs (and anything with a fused upsample/concat) go to the
// halo-resident kernel; stride-2 and 1x1 convs (plain GEMM, nothing to share between taps) to conv_tc.
static int build_conv(const ConvSpec& s, ConvLaunch* L) {
  if (s.s2planes) return build_halo_s2(s, L);
  const bool fused = s.up1 || s.x2;
  const bool same = (s.stride == 1 && s.h_out == s.h && s.w_out == s.w);
  if (fused || (same && (s.ntaps == 9 || s.ntaps == 16) && halo_enabled())) return build_halo(s, L);
  // 1x1 stride-1: TMA-fed halo kernel (no halo, but TMA stores and the two epilogue sets) when cin is a multiple of 64
  static const bool pw_halo = []{ const char* e = getenv("UWM_PW_HALO"); return !(e && e[0] == '0'); }();
  if (same && s.ntaps == 1 && s.cin % 64 == 0 && halo_enabled() && pw_halo) return build_halo(s, L);
  // 1x1 stride-2 (ResNet downsample): the same kernel on a TMA view of every second pixel
  static const bool ds_halo = []{ const char* e = getenv("UWM_DS_HALO"); return !(e && e[0] == '0'); }();
  if (s.stride == 2 && s.ntaps == 1 && s.dh[0] == 0 && s.dw[0] == 0 && s.cin % 64 == 0 && s.cout_pad % 64 == 0 &&
      !(s.h & 1) && !(s.w & 1) && halo_enabled() && pw_halo && ds_halo) {
    ConvSpec v = s;
    v.in_stride2 = 1; v.h_in = s.h; v.w_in = s.w;
    v.stride = 1; v.h = s.h_out; v.w = s.w_out;
    return build_halo(v, L);
  }
  return build_conv_stream(s, L);
}

static int build_conv_stream(const ConvSpec& s, ConvLaunch* L) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(UWM_ECUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if (s.cin % 16) return fail(UWM_EINVAL, "conv: cin=%d must be a multiple of 16", s.cin);
  if (s.cout_pad % 16) return fail(UWM_EINVAL, "conv: cout_pad=%d must be a multiple of 16", s.cout_pad);
  if (s.ntaps < 1 || s.ntaps > kMaxTaps) return fail(UWM_EINVAL, "conv: %d taps unsupported", s.ntaps);
  if (s.stride != 1 && s.stride != 2) return fail(UWM_EINVAL, "conv: stride %d unsupported", s.stride);
  if ((s.x_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.x) & 15))
    return fail(UWM_EINVAL, "conv: activation base/pitch must be 16-byte aligned");
  if (!s.head && ((s.out_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.out) & 15)))
    return fail(UWM_EINVAL, "conv: output base/pitch must be 16-byte aligned");
  if (s.res && ((s.res_pitch * 2) % 16 || (reinterpret_cast<uintptr_t>(s.res) & 15)))
    return fail(UWM_EINVAL, "conv: residual base/pitch must be 16-byte aligned");

  L->halo = 0;
  ConvKArgs& a = L->args;
  memset(&a, 0, sizeof(a));
  a.n_img = s.n; a.h_out = s.h_out; a.w_out = s.w_out;
  choose_tile(s.w_out, s.h_out, s.n, s.stride, &a.tw, &a.th, &a.tn);
  a.tiles_w = (s.w_out + a.tw - 1) / a.tw;
  a.tiles_h = (s.h_out + a.th - 1) / a.th;
  a.tiles_n = (s.n + a.tn - 1) / a.tn;
  a.stride = s.stride;
  a.kc = (s.cin % 64 == 0) ? 64 : (s.cin % 32 == 0) ? 32 : 16;
  a.chunks = s.cin / a.kc;
  a.ntaps = s.ntaps;
  for (int t = 0; t < s.ntaps; ++t) { a.tap_dh[t] = s.dh[t]; a.tap_dw[t] = s.dw[t]; }
  const int m_tiles = a.tiles_w * a.tiles_h * a.tiles_n;
  int bn;
  if (s.cout_pad % 256 == 0 && (long long)m_tiles * (s.cout_pad / 256) * 100 >= 85LL * num_sms()) {
    bn = 256;                         // N=256 halves the A re-reads and doubles the MMA work per k-step
  } else {
    bn = std::min(s.cout_pad, 128);
    while (s.cout_pad % bn) bn -= 16;
  }
  a.block_n = bn;
  a.n_tiles = s.cout_pad / bn;
  a.cout = s.cout;
  a.a_stage_bytes = 128u * a.kc * 2u;
  a.b_stage_bytes = ((uint32_t)bn * a.kc * 2u + 1023u) & ~1023u;
  const int nk = s.ntaps * a.chunks;
  a.total_tiles = m_tiles * a.n_tiles;
  const unsigned grid = (unsigned)std::min(a.total_tiles, num_sms());
  // small layers: keep the whole weight matrix resident in smem (loaded once per persistent CTA)
  const size_t resident_bytes = (size_t)nk * a.b_stage_bytes;
  a.b_resident = (a.n_tiles == 1 && resident_bytes <= 80u * 1024u && a.total_tiles >= 2 * (int)grid) ? 1 : 0;
  const size_t kSmemBudget = 208u * 1024u;
  const size_t fixed = 1024 /*align slack*/ + 1024 /*barriers*/ + (a.b_resident ? resident_bytes : 0);
  const uint32_t kstep_bytes = a.a_stage_bytes + (a.b_resident ? 0u : a.b_stage_bytes);
  // k-steps per ring stage: enough MMA work (>= ~768 cycles at the measured 41-cycle small-N floor) to
  // amortise one barrier round trip, while keeping at least 3 stages in flight
  const int mma_cycles = (a.kc / 16) * std::max(41, bn / 2);
  const int kp_target = (768 + mma_cycles - 1) / mma_cycles;
  const int kp_max = std::max(1, (int)((kSmemBudget - fixed) / 3 / kstep_bytes));
  int kp = std::max(1, std::min(std::min(kp_target, kp_max), nk));
  { const char* e = getenv("UWM_KPACK"); if (e && atoi(e) > 0) kp = std::min(atoi(e), std::min(kp_max, nk)); }
  const int groups = (nk + kp - 1) / kp;
  kp = (nk + groups - 1) / groups;
  a.kpack = kp;
  const uint32_t ring_stage = (uint32_t)kp * kstep_bytes;
  int stages = (int)((kSmemBudget - fixed) / ring_stage);
  stages = std::max(2, std::min(stages, kMaxStages));
  a.stages = stages;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)bn) cols <<= 1;     // two accumulators (double buffering)
  a.tmem_cols = cols;
  a.layout_type = (a.kc == 64) ? kLayoutSw128 : (a.kc == 32) ? kLayoutSw64 : kLayoutSw32;
  a.bias = s.bias;
  a.res = static_cast<const __nv_bfloat16*>(s.res);
  a.out = static_cast<__nv_bfloat16*>(s.out);
  a.res_pitch = s.res_pitch; a.out_pitch = s.out_pitch;
  a.relu = s.relu;
  a.head = s.head; a.apply_sigmoid = s.apply_sigmoid;
  a.logits = s.logits; a.mask = s.mask; a.thr_logit = s.thr_logit;

  a.dbg = dbg_bits();
  L->grid = grid;
  L->smem = fixed + (size_t)stages * ring_stage;

  const CUtensorMapSwizzle sw = (a.kc == 64) ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (a.kc == 32) ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_32B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)s.cin, (cuuint64_t)s.w, (cuuint64_t)s.h, (cuuint64_t)s.n};
    cuuint64_t strides[3] = {(cuuint64_t)s.x_pitch * 2, (cuuint64_t)s.w * s.x_pitch * 2,
                             (cuuint64_t)s.h * s.w * s.x_pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)a.kc, (cuuint32_t)(a.tw * s.stride),
                         (cuuint32_t)(a.th * s.stride), (cuuint32_t)a.tn};
    cuuint32_t est[4] = {1, (cuuint32_t)s.stride, (cuuint32_t)s.stride, 1};
    CUresult r = enc(&L->tm_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.x), dims,
                     strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(act) -> %d (c=%d w=%d h=%d n=%d pitch=%lld box=%u,%u,%u,%u)",
                  (int)r, s.cin, s.w, s.h, s.n, s.x_pitch, box[0], box[1], box[2], box[3]);
  }
  {
    const cuuint64_t ktot = (cuuint64_t)s.ntaps * s.cin;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)s.cout_pad};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)a.kc, (cuuint32_t)bn};
    cuuint32_t est[2] = {1, 1};
    CUresult r = enc(&L->tm_wgt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(s.wgt), dims,
                     strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(UWM_ECUDA, "cuTensorMapEncodeTiled(wgt) -> %d (k=%llu cout=%d)", (int)r,
                  (unsigned long long)ktot, s.cout_pad);
  }
  return UWM_OK;
}

// Instantiation table of conv_halo_kernel<KC, KH, KW, TG, RESIDENT>.  L == nullptr: raise the dynamic
// shared-memory limit of every instantiation (once); otherwise launch the one matching L.
static const HaloChain* const kNoChain = nullptr;

// Multi-layer chain kernels (conv_halo.cuh CHAIN): streamed 3x3, 64-channel chunks, TMA-fed.  L == nullptr: raise the
// shared-memory limit (a chain stages every layer's bias on top of the single-layer budget).
static int chain_dispatch(const ConvLaunch* L, const HaloChain* d_chain, cudaStream_t st) {
#define UWM_HALO_CHAIN(TG)                                                                                        \
  if (!L) {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<64, 3, 3, TG, false, true, 0, false, false, true>,             \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));                      \
  } else if (L->tg == TG) {                                                                                       \
    launch_pdl(conv_halo_kernel<64, 3, 3, TG, false, true, 0, false, false, true>, L->grid, kHaloThreads, L->smem, \
               st, L->tm_wgt, L->tm_out, L->tm_res, L->tm_a0, L->tm_a1, L->hargs, d_chain);                       \
    return UWM_OK;                                                                                                \
  }
  UWM_HALO_CHAIN(1) UWM_HALO_CHAIN(2)
#undef UWM_HALO_CHAIN
  if (!L) return UWM_OK;
  return fail(UWM_ESTATE, "chain conv: no kernel instantiated for tg=%d", L->tg);
}

static int halo_dispatch(const ConvLaunch* L, cudaStream_t st) {
#define UWM_HALO_CASE1(KC, KH, KW, TG, RES, AT)                                                                 \
  if (!L) {                                                                                                     \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<KC, KH, KW, TG, RES, AT>,                                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                    \
  } else if (!L->spx && !L->s2d && !L->cg2 && L->kc == KC && L->kh == KH && L->kw == KW && L->tg == TG && (L->resident != 0) == RES && \
             (L->a_tma != 0) == AT) {                                                                           \
    launch_pdl(conv_halo_kernel<KC, KH, KW, TG, RES, AT>, L->grid, kHaloThreads, L->smem, st, L->tm_wgt, L->tm_out, \
               L->tm_res, L->tm_a0, L->tm_a1, L->hargs, kNoChain);                                                         \
    return UWM_OK;                                                                                              \
  }
#define UWM_HALO_CASE(KC, KH, KW, TG, RES) UWM_HALO_CASE1(KC, KH, KW, TG, RES, false) UWM_HALO_CASE1(KC, KH, KW, TG, RES, true)
  UWM_HALO_CASE(16, 3, 3, 1, true) UWM_HALO_CASE(16, 3, 3, 2, true) UWM_HALO_CASE(16, 3, 3, 4, true) UWM_HALO_CASE(16, 3, 3, 8, true)
  UWM_HALO_CASE(32, 3, 3, 1, true) UWM_HALO_CASE(32, 3, 3, 2, true) UWM_HALO_CASE(32, 3, 3, 4, true) UWM_HALO_CASE(32, 3, 3, 8, true)
  UWM_HALO_CASE(64, 3, 3, 1, true) UWM_HALO_CASE(64, 3, 3, 2, true) UWM_HALO_CASE(64, 3, 3, 4, true)
  UWM_HALO_CASE(16, 3, 3, 1, false) UWM_HALO_CASE(16, 3, 3, 2, false)
  UWM_HALO_CASE(32, 3, 3, 1, false) UWM_HALO_CASE(32, 3, 3, 2, false)
  UWM_HALO_CASE(64, 3, 3, 1, false) UWM_HALO_CASE(64, 3, 3, 2, false)
  UWM_HALO_CASE(16, 4, 4, 1, true) UWM_HALO_CASE(16, 4, 4, 2, true) UWM_HALO_CASE(16, 4, 4, 4, true) UWM_HALO_CASE(16, 4, 4, 8, true)
  // 1x1 stride-1 convs (resnet50 bottlenecks): plain GEMMs, TMA-fed, 64-channel chunks
  UWM_HALO_CASE1(64, 1, 1, 1, true, true) UWM_HALO_CASE1(64, 1, 1, 2, true, true) UWM_HALO_CASE1(64, 1, 1, 4, true, true)
  UWM_HALO_CASE1(64, 1, 1, 1, false, true) UWM_HALO_CASE1(64, 1, 1, 2, false, true)
#undef UWM_HALO_CASE
#undef UWM_HALO_CASE1
  // CTA pairs (CG2): streamed 3x3, 64-channel chunks, TMA-fed, launched as 2-CTA clusters
#define UWM_HALO_CG2(TG)                                                                                          \
  if (!L) {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<64, 3, 3, TG, false, true, 0, false, true>,                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                      \
  } else if (L->cg2 && L->tg == TG) {                                                                             \
    launch_pdl_pairs(conv_halo_kernel<64, 3, 3, TG, false, true, 0, false, true>, L->grid, kHaloThreads, L->smem, \
                     st, L->tm_wgt, L->tm_out, L->tm_res, L->tm_a0, L->tm_a1, L->hargs, kNoChain);                \
    return UWM_OK;                                                                                                \
  }
  UWM_HALO_CG2(1) UWM_HALO_CG2(2)
#undef UWM_HALO_CG2
  if (L && L->cg2) return fail(UWM_ESTATE, "halo conv: no CTA-pair kernel for tg=%d", L->tg);
  // space-to-depth 3x3 convs (S2D): one 64-channel chunk = 4 parity planes x 16, resident weights, TMA-fed
#define UWM_HALO_S2D(TG)                                                                                          \
  if (!L) {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<64, 3, 3, TG, true, true, 0, true>,                            \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                      \
  } else if (L->s2d && L->tg == TG && L->kc == 64 && L->resident && L->a_tma) {                                   \
    launch_pdl(conv_halo_kernel<64, 3, 3, TG, true, true, 0, true>, L->grid, kHaloThreads, L->smem, st,           \
               L->tm_wgt, L->tm_out, L->tm_res, L->tm_a0, L->tm_a1, L->hargs, kNoChain);                          \
    return UWM_OK;                                                                                                \
  }
  UWM_HALO_S2D(1) UWM_HALO_S2D(2) UWM_HALO_S2D(4)
#undef UWM_HALO_S2D
  if (L && L->s2d) return fail(UWM_ESTATE, "halo conv: no space-to-depth kernel for tg=%d resident=%d", L->tg, L->resident);
  // sub-pixel upcat conv (SPX): 64-channel chunks, streamed weights, TMA-fed
#define UWM_HALO_SPX(TG)                                                                                          \
  if (!L) {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<64, 3, 3, TG, false, true, 1>,                                 \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                      \
  } else if (L->spx == 1 && L->tg == TG) {                                                                             \
    launch_pdl(conv_halo_kernel<64, 3, 3, TG, false, true, 1>, L->grid, kHaloThreads, L->smem, st, L->tm_wgt,     \
               L->tm_out, L->tm_res, L->tm_a0, L->tm_a1, L->hargs, kNoChain);                                     \
    return UWM_OK;                                                                                                \
  }
  UWM_HALO_SPX(1) UWM_HALO_SPX(2)
#undef UWM_HALO_SPX
  // sub-pixel convs with one N tile per output parity (SPX == 3): streamed weights, TMA-fed
#define UWM_HALO_PAR(TG)                                                                                          \
  if (!L) {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<64, 3, 3, TG, false, true, 3>,                                 \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                      \
  } else if (L->spx == 3 && L->tg == TG && !L->resident) {                                                        \
    launch_pdl(conv_halo_kernel<64, 3, 3, TG, false, true, 3>, L->grid, kHaloThreads, L->smem, st, L->tm_wgt,     \
               L->tm_out, L->tm_res, L->tm_a0, L->tm_a1, L->hargs, kNoChain);                                     \
    return UWM_OK;                                                                                                \
  }
  UWM_HALO_PAR(1) UWM_HALO_PAR(2)
#undef UWM_HALO_PAR
  // stride-2 3x3 convs over parity planes (SPX == 2): 2x2 block halo
#define UWM_HALO_S2(TG)                                                                                           \
  if (!L) {                                                                                                       \
    CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<64, 2, 2, TG, false, true, 2>,                                 \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));                      \
  } else if (L->spx == 2 && L->tg == TG) {                                                                        \
    launch_pdl(conv_halo_kernel<64, 2, 2, TG, false, true, 2>, L->grid, kHaloThreads, L->smem, st, L->tm_wgt,     \
               L->tm_out, L->tm_res, L->tm_a0, L->tm_a1, L->hargs, kNoChain);                                     \
    return UWM_OK;                                                                                                \
  }
  UWM_HALO_S2(1) UWM_HALO_S2(2)
#undef UWM_HALO_S2
  if (!L) return UWM_OK;
  return fail(UWM_ESTATE, "halo conv: no kernel instantiated for kc=%d %dx%d tg=%d resident=%d a_tma=%d", L->kc, L->kh,
              L->kw, L->tg, L->resident, L->a_tma);
}

static int set_conv_attrs() {
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: tracked per device, not per process
  static bool done_dev[kMaxDevices] = {false};
  bool& done = done_dev[current_device()];
  if (done) return UWM_OK;
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  { int rc = halo_dispatch(nullptr, nullptr); if (rc) return rc; }
  { int rc = chain_dispatch(nullptr, nullptr, nullptr); if (rc) return rc; }
  done = true;
  return UWM_OK;
}

static int launch_conv(const ConvLaunch& L, cudaStream_t st) {
  int rc = set_conv_attrs();
  if (rc) return rc;
  if (L.halo) {
    rc = halo_dispatch(&L, st);
    if (rc) return rc;
    return post_launch("conv_halo_kernel", st);
  }
  switch (L.args.kc) {
    case 64: launch_pdl(conv_tc_kernel<64>, L.grid, kConvThreads, L.smem, st, L.tm_act, L.tm_wgt, L.args); break;
    case 32: launch_pdl(conv_tc_kernel<32>, L.grid, kConvThreads, L.smem, st, L.tm_act, L.tm_wgt, L.args); break;
    default: launch_pdl(conv_tc_kernel<16>, L.grid, kConvThreads, L.smem, st, L.tm_act, L.tm_wgt, L.args); break;
  }
  return post_launch("conv_tc_kernel", st);
}

static void taps_rect(ConvSpec* s, int kh, int kw, int pad) {
  s->ntaps = kh * kw;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) { s->dh[i * kw + j] = (int8_t)(i - pad); s->dw[i * kw + j] = (int8_t)(j - pad); }
}

static unsigned stream_grid(long long work_items, int threads) {
  const long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;   // 16 x 256-thread CTAs per SM = full occupancy
  return (unsigned)std::max(1LL, std::min(blocks, cap));
}

// ------------------------------------------------------------------------------------------
// single-operator entry points
// ------------------------------------------------------------------------------------------
extern "C" int uwm_conv2d_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                    const void* d_wgt, const float* d_bias, int cout, int kh, int kw,
                                    int stride, int pad, const void* d_res, int res_pitch, int relu,
                                    void* d_y, int y_pitch, void* stream) {
  if (!d_x || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d: null pointer");
  if (cout % 16) return fail(UWM_EINVAL, "conv2d: cout=%d must be a multiple of 16", cout);
  if (kh * kw > kMaxTaps) return fail(UWM_EINVAL, "conv2d: %dx%d kernel unsupported", kh, kw);
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = cin; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = cout; s.cout_pad = cout;
  taps_rect(&s, kh, kw, pad);
  s.stride = stride;
  s.h_out = (h + 2 * pad - kh) / stride + 1;
  s.w_out = (w + 2 * pad - kw) / stride + 1;
  s.res = d_res; s.res_pitch = res_pitch; s.relu = relu;
  s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_conv(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_conv2d_upcat_nhwc_bf16(const void* d_x, int n, int h, int w, int c_x, int x_pitch, int upsample,
                                          const void* d_skip, int c_skip, int skip_pitch, const void* d_wgt,
                                          const float* d_bias, int cout, int kh, int kw, int pad, int relu,
                                          void* d_y, int y_pitch, void* stream) {
  if (!d_x || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d_upcat: null pointer");
  if (cout % 16) return fail(UWM_EINVAL, "conv2d_upcat: cout=%d must be a multiple of 16", cout);
  if (kh * kw > kMaxTaps) return fail(UWM_EINVAL, "conv2d_upcat: %dx%d kernel unsupported", kh, kw);
  if (2 * pad != kh - 1 || 2 * pad != kw - 1) return fail(UWM_EINVAL, "conv2d_upcat: needs 'same' padding");
  ConvSpec s;
  s.x = d_x; s.n = n; s.cin = c_x; s.x_pitch = x_pitch; s.up1 = upsample ? 1 : 0;
  s.h = upsample ? 2 * h : h; s.w = upsample ? 2 * w : w;
  if (d_skip) { s.x2 = d_skip; s.cin2 = c_skip; s.x2_pitch = skip_pitch; }
  s.wgt = d_wgt; s.bias = d_bias; s.cout = cout; s.cout_pad = cout;
  taps_rect(&s, kh, kw, pad);
  s.stride = 1; s.h_out = s.h; s.w_out = s.w;
  s.relu = relu; s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_halo(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_conv2d_up2x_shuffle_res_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                                     const void* d_wgt, const float* d_bias, int cout,
                                                     const void* d_res, int res_pitch, int relu, void* d_y, int y_pitch,
                                                     void* stream) {
  if (!d_x || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d_up2x_shuffle: null pointer");
  if (cout % 16 || (4 * cout > 256 && cout != 128 && cout != 256))
    return fail(UWM_EINVAL, "conv2d_up2x_shuffle: cout=%d must be a multiple of 16 up to 64, or 128 / 256", cout);
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = cin; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = 4 * cout; s.cout_pad = 4 * cout; s.shuffle = cout;
  taps_rect(&s, 3, 3, 1);
  s.stride = 1; s.h_out = h; s.w_out = w;
  s.res = d_res; s.res_pitch = res_pitch;
  s.relu = relu; s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_halo(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_conv2d_up2x_shuffle_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                                 const void* d_wgt, const float* d_bias, int cout, int relu, void* d_y,
                                                 int y_pitch, void* stream) {
  if (!d_x || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d_up2x_shuffle: null pointer");
  if (cout % 16 || 4 * cout > 256) return fail(UWM_EINVAL, "conv2d_up2x_shuffle: cout=%d must be a multiple of 16, <= 64", cout);
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = cin; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = 4 * cout; s.cout_pad = 4 * cout; s.shuffle = cout;
  taps_rect(&s, 3, 3, 1);
  s.stride = 1; s.h_out = h; s.w_out = w;
  s.relu = relu; s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_halo(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_conv2d_upcat_subpixel_nhwc_bf16(const void* d_x, int n, int h, int w, int c_x, int x_pitch,
                                                   const void* d_skip, int c_skip, int skip_pitch, const void* d_wgt,
                                                   const float* d_bias, int cout, int relu, void* d_y, int y_pitch,
                                                   void* stream) {
  if (!d_x || !d_skip || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d_upcat_subpixel: null pointer");
  if (cout % 16 || 4 * cout > 256) return fail(UWM_EINVAL, "conv2d_upcat_subpixel: cout=%d must be a multiple of 16, <= 64", cout);
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = c_x; s.x_pitch = x_pitch;
  s.x2 = d_skip; s.cin2 = c_skip; s.x2_pitch = skip_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = 4 * cout; s.cout_pad = 4 * cout; s.shuffle = cout;
  taps_rect(&s, 3, 3, 1);
  s.stride = 1; s.h_out = h; s.w_out = w;
  s.relu = relu; s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_halo_spx(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_head_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                  const void* d_wgt, const float* d_bias, float* d_logits,
                                  int apply_sigmoid, uint8_t* d_mask, float thr_logit, void* stream) {
  if (!d_x || !d_wgt || !d_bias) return fail(UWM_EINVAL, "head: null pointer");
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = cin; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = 1; s.cout_pad = 16;
  taps_rect(&s, 3, 3, 1);
  s.stride = 1; s.h_out = h; s.w_out = w;
  s.head = 1; s.apply_sigmoid = apply_sigmoid; s.logits = d_logits; s.mask = d_mask; s.thr_logit = thr_logit;
  ConvLaunch L;
  int rc = build_conv(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_conv2d_s2_planes_nhwc_bf16(const void* d_x, int n, int h, int w, int cin, int x_pitch,
                                              const void* d_wgt, const float* d_bias, int cout, int relu, void* d_y,
                                              int y_pitch, void* stream) {
  if (!d_x || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d_s2_planes: null pointer");
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = cin; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = cout; s.cout_pad = cout; s.s2planes = 1;
  taps_rect(&s, 3, 3, 1);
  s.stride = 2; s.h_out = h / 2; s.w_out = w / 2;
  s.relu = relu; s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_halo_s2(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_conv2d_s2d_nhwc_bf16(const void* d_x, int n, int h, int w, int x_pitch, const void* d_wgt,
                                        const float* d_bias, int relu, void* d_y, int y_pitch, void* stream) {
  if (!d_x || !d_wgt || !d_bias || !d_y) return fail(UWM_EINVAL, "conv2d_s2d: null pointer");
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = 64; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = 64; s.cout_pad = 64; s.s2d = 1;
  taps_rect(&s, 3, 3, 1);
  s.stride = 1; s.h_out = h; s.w_out = w;
  s.relu = relu; s.out = d_y; s.out_pitch = y_pitch;
  ConvLaunch L;
  int rc = build_halo(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

extern "C" int uwm_head_s2d_nhwc_bf16(const void* d_x, int n, int h, int w, int x_pitch, const void* d_wgt,
                                      const float* d_bias, float* d_logits, int apply_sigmoid, uint8_t* d_mask,
                                      float thr_logit, void* stream) {
  if (!d_x || !d_wgt || !d_bias) return fail(UWM_EINVAL, "head_s2d: null pointer");
  ConvSpec s;
  s.x = d_x; s.n = n; s.h = h; s.w = w; s.cin = 64; s.x_pitch = x_pitch;
  s.wgt = d_wgt; s.bias = d_bias; s.cout = 1; s.cout_pad = 16; s.s2d = 1;
  taps_rect(&s, 3, 3, 1);
  s.stride = 1; s.h_out = h; s.w_out = w;
  s.head = 1; s.apply_sigmoid = apply_sigmoid; s.logits = d_logits; s.mask = d_mask; s.thr_logit = thr_logit;
  ConvLaunch L;
  int rc = build_halo(s, &L);
  if (rc) return rc;
  return launch_conv(L, static_cast<cudaStream_t>(stream));
}

static int launch_maxpool(const void* x, int n, int h, int w, int c, long long xp, void* y, long long yp, int reverse,
                          cudaStream_t st) {
  if (c % 8 || h % 2 || w % 2) return fail(UWM_EINVAL, "maxpool: c%%8, h%%2, w%%2 must be 0");
  // output rows per thread (one thread walks a column strip and keeps the shared input row in registers): fewer rows =
  // more threads in flight for this latency-bound stream (ncu r02: 33 % occupancy, long-scoreboard stalls at 8 rows)
  static const int rows = []{ const char* e = getenv("UWM_POOL_ROWS"); int v = e ? atoi(e) : 8; return (v == 2 || v == 4 || v == 8) ? v : 8; }();   // measured r02: 2 / 4 / 8 within noise
  const long long items = (long long)n * ((h / 2 + rows - 1) / rows) * (w / 2) * (c / 8);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  if (rows == 2) launch_pdl(maxpool3x3s2_kernel<2>, stream_grid(items, 256), 256, 0, st, xb, yb, n, h, w, c, xp, yp, reverse);
  else if (rows == 4) launch_pdl(maxpool3x3s2_kernel<4>, stream_grid(items, 256), 256, 0, st, xb, yb, n, h, w, c, xp, yp, reverse);
  else launch_pdl(maxpool3x3s2_kernel<8>, stream_grid(items, 256), 256, 0, st, xb, yb, n, h, w, c, xp, yp, reverse);
  return post_launch("maxpool3x3s2_kernel", st);
}
static int launch_upsample(const void* x, int n, int h, int w, int c, long long xp, void* y, long long yp,
                           cudaStream_t st) {
  if (c % 8) return fail(UWM_EINVAL, "upsample: c%%8 must be 0");
  const long long items = (long long)n * (h * 2) * (w * 2) * (c / 8);
  launch_pdl(upsample2x_kernel, stream_grid(items, 256), 256, 0, st, static_cast<const __nv_bfloat16*>(x),
             static_cast<__nv_bfloat16*>(y), n, h, w, c, xp, yp);
  return post_launch("upsample2x_kernel", st);
}
static int launch_prep(const void* in, int fmt, int n, int h, int w, void* y, cudaStream_t st) {
  if (h % 2 || w % 2) return fail(UWM_EINVAL, "prep: h, w must be even");
  const long long items = (long long)n * (h / 2) * (w / 2);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(y);
  if (fmt == UWM_IN_U8_NHWC)
    launch_pdl(prep_s2d_kernel<true>, stream_grid((w % 4 == 0) ? items / 2 : items, 256), 256, 0, st, in, out, n, h, w);
  else if (fmt == UWM_IN_F32_NCHW)
    launch_pdl(prep_s2d_kernel<false>, stream_grid(items, 256), 256, 0, st, in, out, n, h, w);
  else
    return fail(UWM_EINVAL, "prep: unknown input format %d", fmt);
  return post_launch("prep_s2d_kernel", st);
}

extern "C" int uwm_maxpool3x3s2_nhwc_bf16(const void* d_x, int n, int h, int w, int c, int x_pitch,
                                          void* d_y, int y_pitch, void* stream) {
  if (!d_x || !d_y) return fail(UWM_EINVAL, "maxpool: null pointer");
  return launch_maxpool(d_x, n, h, w, c, x_pitch, d_y, y_pitch, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int uwm_upsample2x_nhwc_bf16(const void* d_x, int n, int h, int w, int c, int x_pitch,
                                        void* d_y, int y_pitch, void* stream) {
  if (!d_x || !d_y) return fail(UWM_EINVAL, "upsample: null pointer");
  return launch_upsample(d_x, n, h, w, c, x_pitch, d_y, y_pitch, static_cast<cudaStream_t>(stream));
}
extern "C" int uwm_prep_input(const void* d_in, int in_fmt, int n, int h, int w, void* d_y, void* stream) {
  if (!d_in || !d_y) return fail(UWM_EINVAL, "prep: null pointer");
  return launch_prep(d_in, in_fmt, n, h, w, d_y, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// model plan
// ------------------------------------------------------------------------------------------
namespace {

struct Buf {
  size_t bytes_per_img = 0;
  int first = 1 << 30, last = -1;
  size_t off_per_img = 0;
};
struct TRef {            // a [B,h,w,c] bf16 view: channels [c_off, c_off+c) of a pitch-wide buffer
  int buf = -1, c_off = 0, c = 0, pitch = 0, h = 0, w = 0;
  bool s2d = false;      // stores a [B,2h,2w,c/4] tensor space-to-depth: channel = (ph*2+pw)*(c/4) + ch
};
enum OpType { OP_PREP, OP_CONV, OP_HEAD, OP_POOL };
struct Op {
  OpType type;
  std::string name;
  TRef in, in2, out, res;     // in2: second conv source (skip), concatenated after `in` along channels
  bool has_res = false, has_in2 = false;
  bool up = false;            // `in` is nearest-2x upsampled on the fly (decoder conv1)
  bool side = false;          // independent of its neighbour in op order (downsample conv): forked graph branch
  bool join = false;          // first consumer of a side-branch result
  int layer = -1;
  int chain_group = 0;        // > 0: consecutive convs of identical shape that may run as one multi-layer launch (conv_halo.cuh CHAIN)
  double flops_per_img = 0, bytes_per_img = 0;
};
struct Layer {
  uwm_layer_desc d;
  bool stem = false;
  bool shuffle = false;    // decoder conv1 without a skip: 3x3 conv on the upsampled input as a sub-pixel conv
  bool s2planes = false;   // stride-2 3x3 conv on the parity-plane halo kernel (UWM_PACK_S2_PLANES)
  bool s2d = false;        // 3x3 conv over a space-to-depth tensor (conv_halo.cuh S2D); d.cin/d.cout stay the reference's
  void* d_w = nullptr;
  float* d_b = nullptr;
  bool set = false;
};
struct Launch {          // one kernel of an instantiated plan
  OpType type;
  ConvLaunch conv;       // OP_CONV / OP_HEAD (a chain: its first layer, grid / smem / div_total of the whole chain)
  const void* src = nullptr; void* dst = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  long long src_pitch = 0, dst_pitch = 0;
  bool side = false, join = false;
  int reverse = 0;                         // OP_POOL: walk the tensor back to front (convs: conv.hargs.reverse)
  int chain_group = 0;
  int chain_len = 0;                       // > 1: multi-layer chain launch
  HaloChain* d_chain = nullptr;            // device copy of the chain table (owned by the plan)
  std::string name;
  double flops_per_img = 0, bytes_per_img = 0;
};
struct GraphKey {        // caller-owned arguments baked into a captured forward
  const void* in; const void* logits; const void* mask; int in_fmt, apply_sigmoid; uint32_t thr_bits;
  bool operator<(const GraphKey& o) const {
    return std::tie(in, logits, mask, in_fmt, apply_sigmoid, thr_bits) <
           std::tie(o.in, o.logits, o.mask, o.in_fmt, o.apply_sigmoid, o.thr_bits);
  }
};
struct Plan {            // launches of one forward at a fixed batch size (+ the captured graphs)
  std::vector<Launch> launches;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::vector<void*> dev_allocs;           // chain tables and dependency counters
};

}  // namespace

struct uwm_model {
  int encoder = 34, H = 0, W = 0, max_batch = 0;
  int dec[5] = {0};
  bool keep_all = false;
  std::vector<Buf> bufs;
  std::vector<Op> ops;
  std::vector<Layer> layers;
  std::map<std::string, TRef> named;
  size_t arena_bytes = 0;
  uint8_t* arena = nullptr;
  double flops_per_img = 0;
  std::map<int, Plan> plans;
  int last_plan_launches = 0;              // kernels per forward of the most recently used plan
  cudaStream_t cap_stream = nullptr, side_stream = nullptr;   // capture streams (main + forked branch)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  int new_buf(size_t bytes_per_img) {
    Buf b; b.bytes_per_img = (bytes_per_img + 1023) & ~(size_t)1023;
    bufs.push_back(b);
    return (int)bufs.size() - 1;
  }
  TRef dense(int h, int w, int c) {
    TRef t; t.buf = new_buf((size_t)h * w * c * 2); t.c_off = 0; t.c = c; t.pitch = c; t.h = h; t.w = w;
    return t;
  }
  void* ptr(const TRef& t) const {
    return arena + bufs[t.buf].off_per_img * (size_t)max_batch + (size_t)t.c_off * 2;
  }
  void touch(const TRef& t, int op_idx) {
    Buf& b = bufs[t.buf];
    b.first = std::min(b.first, op_idx);
    b.last = std::max(b.last, op_idx);
  }
};

static int add_layer(uwm_model* m, const std::string& conv_key, const std::string& bn_key, int cin,
                     int cout, int k, int stride, int pad, int relu, int has_res, bool stem = false,
                     bool shuffle = false, int spx_cskip = 0, bool s2d = false) {
  Layer L;
  memset(&L.d, 0, sizeof(L.d));
  snprintf(L.d.conv_key, sizeof(L.d.conv_key), "%s", conv_key.c_str());
  snprintf(L.d.bn_key, sizeof(L.d.bn_key), "%s", bn_key.c_str());
  L.d.cin = cin; L.d.cout = cout; L.d.cout_pad = (cout + 15) / 16 * 16;
  L.d.kh = k; L.d.kw = k; L.d.stride = stride; L.d.pad = pad;
  L.d.relu = relu; L.d.has_residual = has_res;
  L.stem = stem;
  if (stem) {
    L.d.pack = UWM_PACK_STEM_S2D;
    L.d.w_elems = (int64_t)L.d.cout_pad * 16 * 16;
  } else if (s2d) {
    // 3x3 conv on a [.,2h,2w,16] tensor kept as [.,h,w,4x16]: 4*cout GEMM columns (16 for the 1-channel head), K = 9 x 64
    L.s2d = true;
    L.d.pack = UWM_PACK_S2D_CONV;
    L.d.cout_pad = (cout == 1) ? 16 : 4 * cout;
    L.d.w_elems = (int64_t)L.d.cout_pad * 9 * 4 * cin;
  } else if (shuffle && spx_cskip > 0) {
    // decoder conv1 WITH a skip, sub-pixel form: see build_halo_spx
    L.shuffle = true;
    L.d.pack = UWM_PACK_UPCAT_SUBPIXEL;
    L.d.cin_skip = spx_cskip;
    L.d.cout_pad = 4 * cout;
    L.d.w_elems = (int64_t)L.d.cout_pad * 64 * (9 * ((cin - spx_cskip) / 64) + 16 * (spx_cskip / 64));
  } else if (shuffle) {
    // conv3x3 over a nearest-2x upsampled input == 3x3 conv on the SOURCE grid producing 4*cout channels (one group
    // per output parity, weights pre-summed over the taps that hit the same source pixel) + pixel shuffle
    L.shuffle = true;
    L.d.pack = UWM_PACK_UP2X_SHUFFLE;
    L.d.cout_pad = 4 * cout;                       // GEMM N; cout must be a multiple of 16 (checked at create)
    L.d.w_elems = (int64_t)L.d.cout_pad * 9 * cin;
  } else {
    // stride-2 3x3 convs whose channels split into 64-wide chunks run on the parity-plane halo kernel
    static const bool s2_on = []{ const char* e = getenv("UWM_S2PLANES"); return !(e && e[0] == '0'); }();
    L.s2planes = s2_on && halo_enabled() && k == 3 && stride == 2 && pad == 1 && cin % 64 == 0 && cout % 64 == 0 && !has_res;
    L.d.pack = L.s2planes ? UWM_PACK_S2_PLANES : UWM_PACK_TAPS;
    L.d.w_elems = (int64_t)L.d.cout_pad * k * k * cin;
  }
  L.d.b_elems = L.d.cout_pad;
  m->layers.push_back(L);
  return (int)m->layers.size() - 1;
}

// conv op: output spatial size is derived from the layer's stride/pad
static void add_conv(uwm_model* m, int layer, const TRef& in, const TRef& out, const TRef* res, bool head = false,
                     bool up = false, const TRef* in2 = nullptr, int out_h = 0, int out_w = 0) {
  Op op;
  op.type = head ? OP_HEAD : OP_CONV;
  const uwm_layer_desc& d = m->layers[layer].d;
  op.name = d.conv_key;
  op.in = in; op.out = out; op.layer = layer; op.up = up;
  if (res) { op.res = *res; op.has_res = true; }
  if (in2) { op.in2 = *in2; op.has_in2 = true; }
  if (head) { op.out.h = out_h; op.out.w = out_w; }
  const double px = (double)op.out.h * op.out.w * (op.out.s2d ? 4.0 : 1.0);   // pixels of the reference's output
  op.flops_per_img = 2.0 * d.cin * d.cout * d.kh * d.kw * px;
  // algorithmic bytes: every source read once at its stored resolution, output written once
  double in_bytes = m->layers[layer].stem ? 2.0 * in.h * in.w * 16 : 2.0 * (double)in.h * in.w * in.c;
  if (in2) in_bytes += 2.0 * (double)in2->h * in2->w * in2->c;
  op.bytes_per_img = in_bytes + (head ? 1.0 * px : 2.0 * px * d.cout) + (res ? 2.0 * px * d.cout : 0.0);
  m->layers[layer].d.flops_per_image = op.flops_per_img;
  m->flops_per_img += op.flops_per_img;
  m->ops.push_back(op);
}

static std::string fmt(const char* f, ...) {
  char b[160];
  va_list ap; va_start(ap, f); vsnprintf(b, sizeof(b), f, ap); va_end(ap);
  return b;
}

static bool subpixel_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UWM_SUBPIXEL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static int build_plan(uwm_model* m) {
  const int H = m->H, W = m->W;
  const bool r50 = (m->encoder == 50);
  const int enc_ch[5] = {64, r50 ? 256 : 64, r50 ? 512 : 128, r50 ? 1024 : 256, r50 ? 2048 : 512};
  const int nblk[4] = {3, 4, 6, 3};
  const int planes[4] = {64, 128, 256, 512};
  const int* dec = m->dec;
  for (int i = 0; i < 5; ++i)
    if (dec[i] <= 0 || dec[i] % 16)
      return fail(UWM_EINVAL, "decoder_channels[%d]=%d: must be a positive multiple of 16", i, dec[i]);

  // decoder block i consumes concat( up2x(x_i) [cx ch], skip_i [cs ch] ) at 1/2^(4-i) scale.  Neither the
  // upsampled tensor nor the concat is materialised: conv1's loader reads x_i and skip_i directly.
  const int cx[5] = {enc_ch[4], dec[0], dec[1], dec[2], dec[3]};
  const int cs[5] = {enc_ch[3], enc_ch[2], enc_ch[1], enc_ch[0], 0};
  TRef skip[5];   // encoder feature feeding decoder block i (filled in as the encoder is laid out)

  // decoder conv1 forms (see the decoder loop below), decided up front: the skip half of a two-launch conv1 is emitted
  // right behind the encoder stage that produces its input, so it can join that stage's chain
  static const bool spx_on = []{ const char* e = getenv("UWM_SPX"); return !(e && e[0] == '0'); }();
  static const bool split_on = []{ const char* e = getenv("UWM_SPLIT_UPCAT"); return !(e && e[0] == '0'); }();
  // Multi-layer chain launches are OPT-IN (UWM_CHAIN=1).  Measured r02 (B200, B = 16, 512x512): bit-identical results; the
  // layer2 chain (8 convs, 256 tiles each, work items handed out by an atomic counter) 131 us as one launch against
  // 8 x ~17 us under programmatic dependent launch; the layer3 chain (12 convs, 128 tiles each on 148 SMs: every item's
  // producers were grabbed less than one item-time earlier) 233 us against 12 x ~17.5 us.  A tile's epilogue + store +
  // the consumer's halo load (~8 k cycles) eat the 0.73 item-times of slack a 256-tile layer leaves on 148 SMs, so the
  // dependencies arrive just in time and the step gains under 1 % (1.032 vs 1.040 ms with the layer2 chain, 1.059 ms with
  // both): kept as an experiment with its trace tooling (tools/gpu_chain_ab.py, tools/gpu_trace_chain.py).
  static const bool chain_on = []{ const char* e = getenv("UWM_CHAIN"); return e && e[0] == '1'; }();
  bool spx_f[5], split_f[5];
  TRef part_t[5];
  for (int i = 0; i < 5; ++i) {
    spx_f[i] = cs[i] > 0 && spx_on && subpixel_enabled() && 4 * dec[i] <= 256 && dec[i] % 16 == 0 && cx[i] % 64 == 0 && cs[i] % 64 == 0;
    split_f[i] = cs[i] > 0 && !spx_f[i] && split_on && subpixel_enabled() && (dec[i] == 128 || dec[i] == 256) &&
                 cx[i] % 64 == 0 && cs[i] % 64 == 0;
  }
  // Chain groups (conv_halo.cuh CHAIN): maximal runs of consecutive stride-1 3x3 convs with cin == cout >= 128 (streamed
  // weights) at one resolution.  Whether a run really launches as one kernel is decided per batch size in
  // instantiate(); here the runs are marked and their tensors' lifetimes extended over the run (no aliasing inside).
  int next_group = 0, cur_group = 0, grp_c = 0, grp_h = 0;
  auto chain_mark = [&](int cin, int cout, int k, int stride, int h) {
    const bool ok = chain_on && halo_enabled() && k == 3 && stride == 1 && cin == cout && cin >= 128 && cin % 64 == 0;
    if (ok && cur_group && grp_c == cin && grp_h == h) { m->ops.back().chain_group = cur_group; return; }
    cur_group = ok ? ++next_group : 0; grp_c = cin; grp_h = h;
    m->ops.back().chain_group = cur_group;
  };
  auto chain_break = [&]() { cur_group = 0; };

  // ---- input prep + stem ----
  TRef xs = m->dense(H / 2, W / 2, 16);
  { Op op; op.type = OP_PREP; op.name = "prep"; op.out = xs;
    op.bytes_per_img = 12.0 * H * W /*fp32 in (u8: 3)*/ + 2.0 * (H / 2) * (W / 2) * 16; m->ops.push_back(op); }
  TRef f1 = m->dense(H / 2, W / 2, 64);   // stem output, 64 ch @ H/2
  skip[3] = f1;
  int l = add_layer(m, "encoder.conv1", "encoder.bn1", 3, 64, 7, 2, 3, 1, 0, /*stem=*/true);
  add_conv(m, l, xs, f1, nullptr);
  m->named["encoder.stem"] = f1;
  TRef x = m->dense(H / 4, W / 4, 64);
  { Op op; op.type = OP_POOL; op.name = "encoder.maxpool"; op.in = f1; op.out = x;
    op.bytes_per_img = 2.0 * 64 * ((double)(H / 2) * (W / 2) + (double)(H / 4) * (W / 4)); m->ops.push_back(op); }
  m->named["encoder.maxpool"] = x;

  // ---- residual stages ----
  int cur_c = 64, cur_h = H / 4, cur_w = W / 4;
  for (int li = 0; li < 4; ++li) {
    const int out_c = r50 ? planes[li] * 4 : planes[li];
    for (int b = 0; b < nblk[li]; ++b) {
      const int stride = (b == 0 && li > 0) ? 2 : 1;
      const int oh = cur_h / stride, ow = cur_w / stride;
      const std::string pre = fmt("encoder.layer%d.%d", li + 1, b);
      const bool last = (b == nblk[li] - 1);
      TRef out = m->dense(oh, ow, out_c);
      if (last && li < 3) skip[2 - li] = out;             // layer1 -> block 2, layer2 -> block 1, layer3 -> block 0
      TRef identity = x;
      const bool need_ds = (b == 0) && (stride != 1 || cur_c != out_c);
      if (need_ds) {
        TRef ds = m->dense(oh, ow, out_c);
        int ld = add_layer(m, pre + ".downsample.0", pre + ".downsample.1", cur_c, out_c, 1, stride, 0, 0, 0);
        add_conv(m, ld, x, ds, nullptr);
        m->ops.back().side = true;      // reads only x: runs beside the block's first conv in the captured graph
        identity = ds;
      }
      if (!r50) {
        TRef t1 = m->dense(oh, ow, planes[li]);
        int l1 = add_layer(m, pre + ".conv1", pre + ".bn1", cur_c, planes[li], 3, stride, 1, 1, 0);
        add_conv(m, l1, x, t1, nullptr);
        chain_mark(cur_c, planes[li], 3, stride, oh);
        int l2 = add_layer(m, pre + ".conv2", pre + ".bn2", planes[li], planes[li], 3, 1, 1, 1, 1);
        add_conv(m, l2, t1, out, &identity);
        m->ops.back().join = need_ds;
        chain_mark(planes[li], planes[li], 3, 1, oh);
      } else {  // torchvision Bottleneck v1.5: stride on the 3x3
        TRef t1 = m->dense(cur_h, cur_w, planes[li]);
        int l1 = add_layer(m, pre + ".conv1", pre + ".bn1", cur_c, planes[li], 1, 1, 0, 1, 0);
        add_conv(m, l1, x, t1, nullptr);
        TRef t2 = m->dense(oh, ow, planes[li]);
        int l2 = add_layer(m, pre + ".conv2", pre + ".bn2", planes[li], planes[li], 3, stride, 1, 1, 0);
        add_conv(m, l2, t1, t2, nullptr);
        int l3 = add_layer(m, pre + ".conv3", pre + ".bn3", planes[li], out_c, 1, 1, 0, 1, 1);
        add_conv(m, l3, t2, out, &identity);
        m->ops.back().join = need_ds;
      }
      x = out; cur_c = out_c; cur_h = oh; cur_w = ow;
    }
    m->named[fmt("encoder.layer%d", li + 1)] = x;
    // skip half of the two-launch decoder conv1 that consumes this stage (stage li feeds decoder block 2 - li)
    const int bi = 2 - li;
    if (li < 3 && bi >= 0 && split_f[bi]) {
      const std::string dpre = fmt("decoder.blocks.%d", bi);
      part_t[bi] = m->dense(cur_h, cur_w, dec[bi]);
      int la = add_layer(m, dpre + ".conv1.0", dpre + ".conv1.1", cs[bi], dec[bi], 3, 1, 1, /*relu=*/0, 0);
      m->layers[la].d.pack = UWM_PACK_TAPS_SKIP_PART; m->layers[la].d.cin_skip = cs[bi];
      add_conv(m, la, x, part_t[bi], nullptr);
      if (!r50) chain_mark(cs[bi], dec[bi], 3, 1, cur_h);
    }
    chain_break();
  }

  // ---- decoder ----
  bool s2d_tail = false;
  for (int i = 0; i < 5; ++i) {
    const int bh = H >> (4 - i), bw = W >> (4 - i);
    const std::string pre = fmt("decoder.blocks.%d", i);
    // Block 4 of the default decoder (16 channels at the input resolution, no skip): both tensors of the block stay
    // space-to-depth [bh/2, bw/2, 4x16] - conv1's sub-pixel GEMM output as it is (no pixel shuffle), conv2 and the
    // head as S2D convs - so every row is 128 bytes (TMA in and out) and the MMAs have N = 64 instead of 16.
    static const bool s2d_on = []{ const char* e = getenv("UWM_S2D"); return !(e && e[0] == '0'); }();
    const bool s2d = (i == 4) && s2d_on && subpixel_enabled() && cs[i] == 0 && dec[i] == 16;
    s2d_tail = s2d;
    TRef t1 = s2d ? m->dense(bh / 2, bw / 2, 4 * dec[i]) : m->dense(bh, bw, dec[i]);
    t1.s2d = s2d;
    // sub-pixel forms (N = 4*cout <= 256 on the source grid): without a skip, and with a skip when both sources
    // split into 64-channel chunks (build_halo_spx)
    const bool spx = spx_f[i];
    const bool subpixel = ((cs[i] == 0) && subpixel_enabled() && 4 * dec[i] <= 256) || spx;
    // Blocks whose sub-pixel form would need more than 256 GEMM columns (default decoder: blocks 0 and 1) run conv1 as
    // two launches: the skip half as an ordinary 3x3 conv into a partial-sum tensor (bias folded here, no ReLU), then
    // the upsampled half as a sub-pixel conv with one N tile per output parity (4 of 9 taps each) whose pixel-shuffle
    // epilogue adds that partial sum and applies the ReLU.  The partial sum is stored in bf16 (one extra rounding).
    const bool split = split_f[i];
    if (split) {
      if (part_t[i].buf < 0) {                   // skip source is not a residual-stage output (the stem feature): emit it here
        part_t[i] = m->dense(bh, bw, dec[i]);
        int la = add_layer(m, pre + ".conv1.0", pre + ".conv1.1", cs[i], dec[i], 3, 1, 1, /*relu=*/0, 0);
        m->layers[la].d.pack = UWM_PACK_TAPS_SKIP_PART; m->layers[la].d.cin_skip = cs[i];
        add_conv(m, la, skip[i], part_t[i], nullptr);
      }
      TRef part = part_t[i];                     // (otherwise the skip half ran right behind its encoder stage, see above)
      int lb = add_layer(m, pre + ".conv1.0", pre + ".conv1.1", cx[i], dec[i], 3, 1, 1, /*relu=*/1, /*has_res=*/1, false,
                         /*shuffle=*/true);
      m->layers[lb].d.pack = UWM_PACK_UP2X_SHUFFLE_X_PART; m->layers[lb].d.cin_skip = cs[i];
      add_conv(m, lb, x, t1, &part, false, /*up=*/true, nullptr);
    } else {
      int l1 = add_layer(m, pre + ".conv1.0", pre + ".conv1.1", cx[i] + cs[i], dec[i], 3, 1, 1, 1, 0, false, subpixel,
                         spx ? cs[i] : 0);
      add_conv(m, l1, x, t1, nullptr, false, /*up=*/true, cs[i] ? &skip[i] : nullptr);
    }
    TRef t2 = s2d ? m->dense(bh / 2, bw / 2, 4 * dec[i]) : m->dense(bh, bw, dec[i]);
    t2.s2d = s2d;
    int l2 = add_layer(m, pre + ".conv2.0", pre + ".conv2.1", dec[i], dec[i], 3, 1, 1, 1, 0, false, false, 0, s2d);
    add_conv(m, l2, t1, t2, nullptr);
    x = t2;
    m->named[pre] = x;
  }
  // ---- head ----
  int lh = add_layer(m, "segmentation_head.0", "", dec[4], 1, 3, 1, 1, 0, 0, false, false, 0, s2d_tail);
  TRef none;
  add_conv(m, lh, x, none, nullptr, /*head=*/true, false, nullptr, H, W);

  // ---- liveness + arena offsets ----
  for (int i = 0; i < (int)m->ops.size(); ++i) {
    Op& op = m->ops[i];
    if (op.in.buf >= 0) m->touch(op.in, i);
    if (op.out.buf >= 0) m->touch(op.out, i);
    if (op.has_res) m->touch(op.res, i);
    if (op.has_in2) m->touch(op.in2, i);
  }
  // tensors touched inside a chain group live for the whole group: layer l + 1 of one image runs while layer l of
  // another is still in flight, so nothing inside the group may share memory
  {
    std::map<int, std::pair<int, int>> span;      // group -> [first op, last op]
    for (int i = 0; i < (int)m->ops.size(); ++i) {
      const int g = m->ops[i].chain_group;
      if (!g) continue;
      auto it = span.find(g);
      if (it == span.end()) span[g] = {i, i}; else it->second.second = i;
    }
    for (int i = 0; i < (int)m->ops.size(); ++i) {
      const Op& op = m->ops[i];
      if (!op.chain_group) continue;
      const auto sp = span[op.chain_group];
      auto widen = [&](const TRef& t) {
        if (t.buf < 0) return;
        Buf& b = m->bufs[t.buf];
        b.first = std::min(b.first, sp.first); b.last = std::max(b.last, sp.second);
      };
      widen(op.in); widen(op.out);
      if (op.has_res) widen(op.res);
    }
  }
  std::vector<int> order;
  for (int i = 0; i < (int)m->bufs.size(); ++i) if (m->bufs[i].last >= 0) order.push_back(i);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return m->bufs[a].first < m->bufs[b].first; });
  std::vector<int> placed;
  size_t top = 0;
  for (int id : order) {
    Buf& b = m->bufs[id];
    size_t off = 0;
    if (m->keep_all) { off = top; }
    else {
      // first-fit among buffers whose lifetime overlaps
      std::vector<std::pair<size_t, size_t>> busy;
      for (int pid : placed) {
        const Buf& pb = m->bufs[pid];
        if (pb.last >= b.first && pb.first <= b.last) busy.push_back({pb.off_per_img, pb.off_per_img + pb.bytes_per_img});
      }
      std::sort(busy.begin(), busy.end());
      for (auto& iv : busy) {
        if (off + b.bytes_per_img <= iv.first) break;
        off = std::max(off, iv.second);
      }
    }
    b.off_per_img = off;
    top = std::max(top, off + b.bytes_per_img);
    placed.push_back(id);
  }
  m->arena_bytes = top * (size_t)m->max_batch;
  return UWM_OK;
}

extern "C" int uwm_model_create(int encoder, const int* decoder_channels, int h, int w, int max_batch,
                                uwm_model** out) {
  if (!out || !decoder_channels) return fail(UWM_EINVAL, "model_create: null pointer");
  if (encoder != 34 && encoder != 50) return fail(UWM_EINVAL, "model_create: encoder resnet%d unsupported (34|50)", encoder);
  if (h <= 0 || w <= 0 || h % 32 || w % 32)
    return fail(UWM_EINVAL, "Wrong input shape height=%d, width=%d. Expected image height and width divisible by 32.", h, w);
  if (max_batch < 1) return fail(UWM_EINVAL, "model_create: max_batch must be >= 1");
  uwm_model* m = new uwm_model();
  m->encoder = encoder; m->H = h; m->W = w; m->max_batch = max_batch;
  for (int i = 0; i < 5; ++i) m->dec[i] = decoder_channels[i];
  const char* ka = getenv("UWM_KEEP_ALL");
  m->keep_all = ka && ka[0] == '1';
  int rc = build_plan(m);
  if (rc) { delete m; return rc; }
  cudaError_t e = cudaMalloc(&m->arena, m->arena_bytes);
  if (e != cudaSuccess) {
    rc = fail(UWM_ENOMEM, "cudaMalloc(%zu bytes of activation workspace) failed: %s", m->arena_bytes, cudaGetErrorString(e));
    delete m; return rc;
  }
  e = cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->side_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) { rc = fail(UWM_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)); cudaFree(m->arena); delete m; return rc; }
  *out = m;
  return UWM_OK;
}

extern "C" int uwm_model_destroy(uwm_model* m) {
  if (!m) return UWM_OK;
  for (auto& kv : m->plans) {
    for (auto& g : kv.second.graphs) cudaGraphExecDestroy(g.second);
    for (void* q : kv.second.dev_allocs) cudaFree(q);
  }
  for (auto& L : m->layers) { if (L.d_w) cudaFree(L.d_w); if (L.d_b) cudaFree(L.d_b); }
  if (m->arena) cudaFree(m->arena);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  if (m->side_stream) cudaStreamDestroy(m->side_stream);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  delete m;
  return UWM_OK;
}
extern "C" int uwm_model_num_layers(const uwm_model* m) { return m ? (int)m->layers.size() : 0; }
extern "C" int uwm_model_layer_desc(const uwm_model* m, int i, uwm_layer_desc* d) {
  if (!m || !d || i < 0 || i >= (int)m->layers.size()) return fail(UWM_EINVAL, "layer_desc: bad index %d", i);
  *d = m->layers[i].d;
  return UWM_OK;
}
extern "C" int uwm_model_set_layer(uwm_model* m, int i, const void* wgt, int64_t w_elems, const float* bias,
                                   int64_t b_elems) {
  if (!m || !wgt || !bias || i < 0 || i >= (int)m->layers.size()) return fail(UWM_EINVAL, "set_layer: bad argument");
  Layer& L = m->layers[i];
  if (w_elems != L.d.w_elems || b_elems != L.d.b_elems)
    return fail(UWM_EINVAL, "set_layer(%s): expected %lld weight / %lld bias elements, got %lld / %lld", L.d.conv_key,
                (long long)L.d.w_elems, (long long)L.d.b_elems, (long long)w_elems, (long long)b_elems);
  // a forward may still be in flight on a non-blocking stream: the blocking copies below are only ordered against
  // the legacy stream, so drain the device first (weight uploads are rare)
  if (i == 0) CUDA_TRY(cudaDeviceSynchronize());
  if (!L.d_w) CUDA_TRY(cudaMalloc(&L.d_w, (size_t)w_elems * 2));
  if (!L.d_b) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&L.d_b), (size_t)b_elems * 4));
  CUDA_TRY(cudaMemcpy(L.d_w, wgt, (size_t)w_elems * 2, cudaMemcpyDefault));
  CUDA_TRY(cudaMemcpy(L.d_b, bias, (size_t)b_elems * 4, cudaMemcpyDefault));
  L.set = true;
  return UWM_OK;
}
extern "C" size_t uwm_model_workspace_bytes(const uwm_model* m) { return m ? m->arena_bytes : 0; }
// kernels per forward: of the plan used last (chains merge several convs into one launch), else one per planned op
extern "C" int uwm_model_num_kernels(const uwm_model* m) {
  if (!m) return 0;
  return m->last_plan_launches > 0 ? m->last_plan_launches : (int)m->ops.size();
}
extern "C" double uwm_model_flops_per_image(const uwm_model* m) { return m ? m->flops_per_img : 0.0; }

// Instantiate the launches of one forward for `batch` images.
static int merge_chains(std::vector<Launch>* ls, std::vector<void*>* dev_allocs);

static int instantiate(uwm_model* m, const void* d_in, int in_fmt, int batch, float* d_logits, int apply_sigmoid,
                       uint8_t* d_mask, float thr_logit, std::vector<Launch>* out, std::vector<void*>* dev_allocs) {
  out->clear();
  out->reserve(m->ops.size());
  // Serpentine tile order (UWM_SERP, default on): a launch walks its tiles in the opposite direction of the launch that
  // wrote its input, so it starts on the most recently written - still L2-resident - end of a tensor that is larger
  // than L2 (stem -> max-pool, decoder block 4 -> head: 134 MB each at 16 x 512 x 512), instead of on the end that the
  // producer's later writes have already evicted.
  static const bool serp = []{ const char* e = getenv("UWM_SERP"); return !(e && e[0] == '0'); }();
  std::map<const void*, int> dir_of;         // tensor base -> direction its producer walked (1 = back to front)
  // evict_first loads of dead sources (UWM_L2_HINT = workspace limit in MB, default 1024, 0 = off): a conv source that
  // nobody touches after this launch is loaded with the evict_first policy, so its lines leave L2 before the lines that
  // are still waiting to be read.  That pays while a good part of the producer -> consumer hand-overs still hit L2
  // (config 2: 0.6 GB of workspace, -0.6 ... -0.8 % over three runs) and costs when the forward's workspace is many
  // times L2 and the hint only evicts the halo rows a neighbouring tile is about to re-read (config 3, 9.6 GB: +0.7 ...
  // +1.0 %; config 4: +0.3 ... +0.6 %) - hence the limit on this batch's workspace.
  static const double hint_mb = []{ const char* e = getenv("UWM_L2_HINT"); return e ? atof(e) : 1024.0; }();
  const bool hint_on = hint_mb > 0 && !m->keep_all && (double)(m->arena_bytes / (size_t)m->max_batch) * batch <= hint_mb * 1e6;
  int op_idx = -1;
  for (const Op& op : m->ops) {
    ++op_idx;
    Launch L;
    L.type = op.type;
    L.side = op.side; L.join = op.join;
    L.chain_group = op.chain_group;
    L.name = op.name; L.flops_per_img = op.flops_per_img; L.bytes_per_img = op.bytes_per_img;
    switch (op.type) {
      case OP_PREP:
        L.src = d_in; L.dst = m->ptr(op.out); L.n = batch; L.h = m->H; L.w = m->W; L.c = in_fmt;
        break;
      case OP_POOL:
        L.src = m->ptr(op.in); L.dst = m->ptr(op.out); L.n = batch; L.h = op.in.h; L.w = op.in.w; L.c = op.in.c;
        L.src_pitch = op.in.pitch; L.dst_pitch = op.out.pitch;
        break;
      case OP_CONV:
      case OP_HEAD: {
        const Layer& ly = m->layers[op.layer];
        if (!ly.set) return fail(UWM_ESTATE, "weights of layer %s not set", ly.d.conv_key);
        ConvSpec s;
        s.x = m->ptr(op.in); s.n = batch; s.h = op.in.h; s.w = op.in.w; s.x_pitch = op.in.pitch;
        s.wgt = ly.d_w; s.bias = ly.d_b; s.cout = ly.d.cout; s.cout_pad = ly.d.cout_pad;
        if (ly.stem) {
          s.cin = 16; s.stride = 1;
          s.ntaps = 16;
          for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { s.dh[r * 4 + c] = (int8_t)(r - 2); s.dw[r * 4 + c] = (int8_t)(c - 2); }
          s.h_out = op.in.h; s.w_out = op.in.w;
        } else {
          s.cin = op.in.c; s.stride = ly.d.stride;
          if (ly.s2planes) s.s2planes = 1;     // input sizes are even (H, W divisible by 32)
          if (ly.s2d) { s.s2d = 1; if (op.type != OP_HEAD) s.cout = ly.d.cout_pad; }
          if (ly.shuffle) {   // runs on the source grid, N = 4*cout; a space-to-depth output keeps the GEMM layout
            s.cout = ly.d.cout_pad;
            if (!op.out.s2d) s.shuffle = ly.d.cout;
          }
          else if (op.up) { s.up1 = 1; s.h = op.in.h * 2; s.w = op.in.w * 2; }
          if (op.has_in2) { s.x2 = m->ptr(op.in2); s.cin2 = op.in2.c; s.x2_pitch = op.in2.pitch; }
          if (s.cin + s.cin2 != (ly.s2d ? 4 : 1) * ly.d.cin)
            return fail(UWM_ESTATE, "plan bug: %s reads %d+%d channels, layer has %d", ly.d.conv_key, s.cin, s.cin2, ly.d.cin);
          taps_rect(&s, ly.d.kh, ly.d.kw, ly.d.pad);
          s.h_out = (s.h + 2 * ly.d.pad - ly.d.kh) / ly.d.stride + 1;
          s.w_out = (s.w + 2 * ly.d.pad - ly.d.kw) / ly.d.stride + 1;
        }
        s.relu = ly.d.relu;
        if (op.type == OP_HEAD) {
          s.head = 1; s.apply_sigmoid = apply_sigmoid; s.logits = d_logits; s.mask = d_mask; s.thr_logit = thr_logit;
        } else {
          const int up_out = (ly.shuffle && !op.out.s2d) ? 2 : 1;   // the pixel-shuffling sub-pixel conv writes a 2x larger tensor
          if (s.h_out * up_out != op.out.h || s.w_out * up_out != op.out.w)
            return fail(UWM_ESTATE, "plan bug: %s output %dx%d != planned %dx%d", ly.d.conv_key, s.h_out * up_out,
                        s.w_out * up_out, op.out.h, op.out.w);
          s.out = m->ptr(op.out); s.out_pitch = op.out.pitch;
          if (op.has_res) { s.res = m->ptr(op.res); s.res_pitch = op.res.pitch; }
        }
        int rc = build_conv(s, &L.conv);
        if (rc) return rc;
        break;
      }
    }
    if (hint_on && (op.type == OP_CONV || op.type == OP_HEAD) && L.conv.halo) {
      auto dead = [&](const TRef& t) { return t.buf >= 0 && m->bufs[t.buf].last == op_idx; };
      if (dead(op.in)) L.conv.hargs.a_policy[0] = kL2EvictFirst;
      if (op.has_in2 && dead(op.in2)) L.conv.hargs.a_policy[1] = kL2EvictFirst;
    }
    if (serp && op.type != OP_PREP) {
      const auto it = dir_of.find(m->ptr(op.in));
      int dir = (it == dir_of.end() ? 0 : it->second) ^ 1;
      if (op.type == OP_POOL) L.reverse = dir;
      else if (L.conv.halo && !L.conv.cg2 && op.chain_group == 0) L.conv.hargs.reverse = dir;
      else dir = 0;
      if (op.type != OP_HEAD) dir_of[m->ptr(op.out)] = dir;
    }
    out->push_back(L);
  }
  return merge_chains(out, dev_allocs);
}

// Consecutive launches of one chain group that came out as the same kernel instantiation with the same geometry
// become ONE multi-layer launch (conv_halo.cuh CHAIN).
static int merge_chains(std::vector<Launch>* ls, std::vector<void*>* dev_allocs) {
  // layers with fewer than ~1.5 tiles per SM are not chained by default: their items depend on the items grabbed
  // less than one item-time earlier, so every item would wait for its producers (measured: layer3, 128 tiles per layer)
  static const int min_tiles = []{ const char* e = getenv("UWM_CHAIN_MIN_TILES"); return e ? atoi(e) : 3 * num_sms() / 2; }();
  static const int max_len = []{ const char* e = getenv("UWM_CHAIN_MAX"); int v = e ? atoi(e) : kMaxChainLayers; return std::max(1, std::min(v, kMaxChainLayers)); }();
  static const int min_len = []{ const char* e = getenv("UWM_CHAIN_MIN"); return e ? atoi(e) : 2; }();
  auto chainable = [](const Launch& L) {
    const ConvLaunch& c = L.conv;
    return L.type == OP_CONV && L.chain_group > 0 && c.hargs.total_tiles >= min_tiles && c.halo && !c.spx && !c.s2d && !c.cg2 && !c.resident && c.a_tma &&
           c.kc == 64 && c.kh == 3 && c.kw == 3 && (c.tg == 1 || c.tg == 2) && c.hargs.ep_tma && !c.hargs.shuffle &&
           c.hargs.split_chunk == c.hargs.chunks && c.hargs.a_scale == 1;
  };
  auto same_shape = [](const ConvLaunch& a, const ConvLaunch& b) {
    const HaloKArgs &x = a.hargs, &y = b.hargs;
    return a.tg == b.tg && x.n_img == y.n_img && x.h == y.h && x.w == y.w && x.chunks == y.chunks && x.block_n == y.block_n &&
           x.n_tiles == y.n_tiles && x.cout == y.cout && x.kpb == y.kpb && x.a_stages == y.a_stages && x.b_stages == y.b_stages &&
           x.total_tiles == y.total_tiles && x.tmem_cols == y.tmem_cols && x.nacc_log2 == y.nacc_log2 && x.cin_total == y.cin_total &&
           x.b_slice_bytes == y.b_slice_bytes && a.smem == b.smem;
  };
  std::vector<Launch> out;
  size_t i = 0;
  while (i < ls->size()) {
    size_t j = i + 1;
    if (chainable((*ls)[i]))
      while (j < ls->size() && j - i < (size_t)max_len && (*ls)[j].chain_group == (*ls)[i].chain_group &&
             chainable((*ls)[j]) && same_shape((*ls)[i].conv, (*ls)[j].conv)) ++j;
    const int n = (int)(j - i);
    const size_t smem = (*ls)[i].conv.smem + (size_t)(n - 1) * (*ls)[i].conv.hargs.cout * 4;   // every layer's bias is staged
    if (n < min_len || !chainable((*ls)[i]) || smem > 225u * 1024u) {
      for (size_t k = i; k < j; ++k) out.push_back((*ls)[k]);
      i = j;
      continue;
    }
    HaloChain hc;
    memset(&hc, 0, sizeof(hc));
    Launch M = (*ls)[i];
    M.name = "chain[" + std::to_string(n) + "] " + (*ls)[i].name + " .. " + (*ls)[j - 1].name;
    M.flops_per_img = 0; M.bytes_per_img = 0;
    for (int l = 0; l < n; ++l) {
      const Launch& Lk = (*ls)[i + l];
      HaloLayerRef& r = hc.layer[l];
      r.tm_wgt = Lk.conv.tm_wgt; r.tm_out = Lk.conv.tm_out; r.tm_res = Lk.conv.tm_res; r.tm_a0 = Lk.conv.tm_a0;
      r.bias = Lk.conv.hargs.bias; r.relu = Lk.conv.hargs.relu; r.has_res = Lk.conv.hargs.res != nullptr;
      r.res_layer = -1;                      // which chain layer wrote this layer's residual (the epilogue waits for it)
      for (int k = 0; k < l; ++k)
        if (r.has_res && (*ls)[i + k].conv.hargs.out == Lk.conv.hargs.res) r.res_layer = k;
      M.flops_per_img += Lk.flops_per_img; M.bytes_per_img += Lk.bytes_per_img;
    }
    const HaloKArgs& a = M.conv.hargs;
    hc.n_layers = n;
    hc.dep_target = a.tiles_w * a.tiles_h * a.n_tiles * 8;          // tiles of one image x 8 epilogue warps
    int* d_dep = nullptr;
    const size_t dep_bytes = ((size_t)n * a.n_img + 2) * sizeof(int);     // counters, CTA ticket, work-item counter
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_dep), dep_bytes));
    dev_allocs->push_back(d_dep);
    CUDA_TRY(cudaMemset(d_dep, 0, dep_bytes));
    hc.dep = d_dep;
    HaloChain* d_chain = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_chain), sizeof(HaloChain)));
    dev_allocs->push_back(d_chain);
    CUDA_TRY(cudaMemcpy(d_chain, &hc, sizeof(HaloChain), cudaMemcpyHostToDevice));
    M.chain_len = n; M.d_chain = d_chain;
    M.conv.smem = smem;
    M.conv.grid = (unsigned)std::min((long long)n * a.total_tiles, (long long)num_sms());
    { const char* e = getenv("UWM_CHAIN_GRID"); if (e && atoi(e) > 0) M.conv.grid = (unsigned)std::min<long long>(atoi(e), M.conv.grid); }
    { const char* e = getenv("UWM_VERBOSE");
      if (e && e[0] == '1') fprintf(stderr, "chain of %d layers: %s (tiles/layer %d, grid %u, dep target %d, smem %zu)\n", n,
                                    M.name.c_str(), a.total_tiles, M.conv.grid, hc.dep_target, smem); }
    out.push_back(M);
    i = j;
  }
#ifdef UWM_BENCH_TOOLS
  // trace one chain launch only (tools/gpu_trace_chain.py): every other launch of the plan loses its trace pointer
  if (const char* e = getenv("UWM_TRACE_LAUNCH")) {       // trace the k-th launch of the plan only
    int want = atoi(e), k = 0;
    for (Launch& L : out) { if (k++ != want) L.conv.hargs.trace = nullptr; else fprintf(stderr, "tracing launch %d: %s\n", want, L.name.c_str()); }
  } else
  if (const char* e = getenv("UWM_TRACE_CHAIN")) {
    int want = atoi(e), k = 0;
    for (Launch& L : out) {
      const bool keep = L.chain_len > 1 && k++ == want;
      if (!keep) { L.conv.hargs.trace = nullptr; }
    }
  }
#endif
  ls->swap(out);
  return UWM_OK;
}

static int run_launch(const Launch& L, cudaStream_t st) {
  PlanScope in_plan;
  switch (L.type) {
    case OP_PREP: return launch_prep(L.src, L.c, L.n, L.h, L.w, L.dst, st);
    case OP_POOL: return launch_maxpool(L.src, L.n, L.h, L.w, L.c, L.src_pitch, L.dst, L.dst_pitch, L.reverse, st);
    case OP_CONV:
    case OP_HEAD:
      if (L.chain_len > 1) {
        int rc = set_conv_attrs();
        if (rc) return rc;
        rc = chain_dispatch(&L.conv, L.d_chain, st);
        if (rc) return rc;
        return post_launch("conv_halo_kernel (chain)", st);
      }
      return launch_conv(L.conv, st);
  }
  return UWM_OK;
}

static int check_forward_args(uwm_model* m, const void* d_in, int in_fmt, int batch, float* d_logits, uint8_t* d_mask) {
  if (!m || !d_in) return fail(UWM_EINVAL, "forward: null model/input");
  if (batch < 1 || batch > m->max_batch) return fail(UWM_ESTATE, "forward: batch %d outside [1, max_batch=%d]", batch, m->max_batch);
  if (in_fmt != UWM_IN_F32_NCHW && in_fmt != UWM_IN_U8_NHWC) return fail(UWM_EINVAL, "forward: unknown input format %d", in_fmt);
  if (!d_logits && !d_mask) return fail(UWM_EINVAL, "forward: both outputs are NULL");
  return UWM_OK;
}

extern "C" int uwm_model_forward(uwm_model* m, const void* d_in, int in_fmt, int batch, float* d_logits,
                                 int apply_sigmoid, uint8_t* d_mask, float thr_logit, int use_graph, void* stream) {
  int rc = check_forward_args(m, d_in, in_fmt, batch, d_logits, d_mask);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto pit = m->plans.find(batch);
  if (pit == m->plans.end()) {
    // tensor maps and tile shapes depend on the batch only; the caller-owned pointers of the first (prep)
    // and last (head) kernels are patched per call
    Plan pl;
    rc = instantiate(m, d_in, in_fmt, batch, d_logits, apply_sigmoid, d_mask, thr_logit, &pl.launches, &pl.dev_allocs);
    if (rc) { for (void* q : pl.dev_allocs) cudaFree(q); return rc; }
    pit = m->plans.emplace(batch, std::move(pl)).first;
  }
  Plan& pl = pit->second;
  m->last_plan_launches = (int)pl.launches.size();
  Launch& first = pl.launches.front();
  Launch& last = pl.launches.back();
  first.src = d_in; first.c = in_fmt;
  last.conv.args.logits = d_logits; last.conv.args.mask = d_mask;
  last.conv.args.thr_logit = thr_logit; last.conv.args.apply_sigmoid = apply_sigmoid;
  last.conv.hargs.logits = d_logits; last.conv.hargs.mask = d_mask;
  last.conv.hargs.thr_logit = thr_logit; last.conv.hargs.apply_sigmoid = apply_sigmoid;
  const size_t n = pl.launches.size();
  if (!use_graph || debug_sync()) {
    for (const Launch& L : pl.launches) { rc = run_launch(L, st); if (rc) return rc; }
    return UWM_OK;
  }
  // One CUDA graph per distinct argument set (input / output pointers, threshold, activation): the whole
  // forward, prep to head, is a single graph launch whose kernel-to-kernel edges are programmatic (PDL).
  uint32_t thr_bits; memcpy(&thr_bits, &thr_logit, 4);
  const GraphKey key{d_in, d_logits, d_mask, in_fmt, apply_sigmoid, thr_bits};
  auto git = pl.graphs.find(key);
  if (git == pl.graphs.end()) {
    if (pl.graphs.size() >= kMaxGraphsPerPlan) {          // callers that rotate through unbounded pointer sets: start over
      for (auto& kv : pl.graphs) cudaGraphExecDestroy(kv.second);
      pl.graphs.clear();
    }
    rc = set_conv_attrs();
    if (rc) return rc;
    cudaGraph_t g = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeRelaxed));
    // The downsample conv of a residual stage reads only the block input, so it can be captured on a forked stream
    // beside the block's first conv and joined before the conv that adds it as the residual (UWM_FORK=1).  That paid
    // while the downsample took 12-22 us on conv_tc_kernel; at ~10 us on the halo kernel the fork/join edges (which
    // also break the programmatic-launch chain) cost more than the overlap gives, so the default is one stream.
    static const bool fork_side = []{ const char* e = getenv("UWM_FORK"); return e && e[0] == '1'; }();
    bool pending_join = false;
    for (size_t i = 0; i < n; ++i) {
      const Launch& L = pl.launches[i];
      if (L.side && fork_side) {
        cudaEventRecord(m->ev_fork, m->cap_stream);
        cudaStreamWaitEvent(m->side_stream, m->ev_fork, 0);
        rc = run_launch(L, m->side_stream);
        cudaEventRecord(m->ev_join, m->side_stream);
        pending_join = true;
      } else {
        if (L.join && pending_join) { cudaStreamWaitEvent(m->cap_stream, m->ev_join, 0); pending_join = false; }
        rc = run_launch(L, m->cap_stream);
      }
      if (rc) { cudaStreamEndCapture(m->cap_stream, &g); if (g) cudaGraphDestroy(g); return rc; }
    }
    if (pending_join) cudaStreamWaitEvent(m->cap_stream, m->ev_join, 0);
    CUDA_TRY(cudaStreamEndCapture(m->cap_stream, &g));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(UWM_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    g_launches.fetch_sub(n, std::memory_order_relaxed);   // capture is not execution
    git = pl.graphs.emplace(key, exec).first;
  }
  CUDA_TRY(cudaGraphLaunch(git->second, st));
  g_launches.fetch_add(n, std::memory_order_relaxed);
  return UWM_OK;
}

__global__ void gather_pitched_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      long long pixels, int c, long long pitch) {
  const int cg = c >> 3;
  const long long total = pixels * cg;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long px = idx / cg; const int g = (int)(idx % cg);
    *reinterpret_cast<uint4*>(dst + px * c + g * 8) = *reinterpret_cast<const uint4*>(src + px * pitch + g * 8);
  }
}

// a space-to-depth plan tensor [px(h,w)][4 x c4] read back as dense [2h][2w][c4]
__global__ void gather_d2s_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n, int h,
                                  int w, int c4, long long pitch) {
  const int cg = c4 / 8;
  const long long total = (long long)n * 2 * h * 2 * w * cg;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long px = idx / cg;
    const int x = (int)(px % (2 * w)); px /= 2 * w;
    const int y = (int)(px % (2 * h)); const int img = (int)(px / (2 * h));
    const long long sp = ((long long)img * h + (y >> 1)) * w + (x >> 1);
    *reinterpret_cast<uint4*>(dst + (idx / cg) * c4 + g * 8) =
        *reinterpret_cast<const uint4*>(src + sp * pitch + ((y & 1) * 2 + (x & 1)) * c4 + g * 8);
  }
}

extern "C" int uwm_model_read_tensor(uwm_model* m, const char* name, int batch, void* d_dst, int64_t dst_bytes,
                                     int* h, int* w, int* c, void* stream) {
  if (!m || !name) return fail(UWM_EINVAL, "read_tensor: null argument");
  auto it = m->named.find(name);
  if (it == m->named.end()) return fail(UWM_EINVAL, "read_tensor: unknown tensor '%s'", name);
  const TRef& t = it->second;
  const int up = t.s2d ? 2 : 1;
  if (h) *h = t.h * up; if (w) *w = t.w * up; if (c) *c = t.c / (up * up);
  if (!d_dst) return UWM_OK;
  const long long px = (long long)batch * t.h * t.w;
  if (dst_bytes < px * t.c * 2) return fail(UWM_EINVAL, "read_tensor: destination too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (t.s2d) {
    gather_d2s_kernel<<<stream_grid(px * (t.c / 8), 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(m->ptr(t)), static_cast<__nv_bfloat16*>(d_dst), batch, t.h, t.w, t.c / 4, t.pitch);
    return post_launch("gather_d2s_kernel", st);
  }
  gather_pitched_kernel<<<stream_grid(px * (t.c / 8), 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(m->ptr(t)), static_cast<__nv_bfloat16*>(d_dst), px, t.c, t.pitch);
  return post_launch("gather_pitched_kernel", st);
}

extern "C" int uwm_model_profile(uwm_model* m, const void* d_in, int in_fmt, int batch, float* d_logits,
                                 uint8_t* d_mask, float thr_logit, char* names, float* ms, double* flops,
                                 double* bytes, int n_max, void* stream) {
  int rc = check_forward_args(m, d_in, in_fmt, batch, d_logits, d_mask);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<Launch> ls;
  std::vector<void*> allocs;
  struct FreeAll { std::vector<void*>* v; ~FreeAll() { for (void* q : *v) cudaFree(q); } } free_all{&allocs};
  rc = instantiate(m, d_in, in_fmt, batch, d_logits, 0, d_mask, thr_logit, &ls, &allocs);
  if (rc) return rc;
  const int n = (int)ls.size();
  if (n > n_max) return fail(UWM_EINVAL, "profile: need room for %d kernels", n);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
  CUDA_TRY(cudaEventRecord(ev[0], st));
  for (int i = 0; i < n; ++i) {
    rc = run_launch(ls[i], st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev[i + 1], st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i) {
    CUDA_TRY(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    snprintf(names + (size_t)i * 64, 64, "%s", ls[i].name.c_str());
    if (flops) flops[i] = ls[i].flops_per_img * batch;
    if (bytes) bytes[i] = ls[i].bytes_per_img * batch;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return n;
}

#ifdef UWM_BENCH_TOOLS
// ------------------------------------------------------------------------------------------
// micro-benchmarks (sizing experiments; not part of the product path)
// ------------------------------------------------------------------------------------------
extern "C" int uwm_debug_prim_cost(int which, int iters, long long* d_out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  prim_cost_kernel<<<1, 32, 0, st>>>(which, iters, d_out);
  return post_launch("prim_cost_kernel", st);
}

extern "C" int uwm_debug_set_trace(long long* d_trace) { g_halo_trace = d_trace; return UWM_OK; }

extern "C" int uwm_debug_mma_rate(int n, int iters, int distinct_stages, int mode, int blocks, long long* d_cycles, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = 1024 + (size_t)distinct_stages * (16384 + 32768);
  CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  mma_rate_kernel<<<blocks, 128, smem, st>>>(n, iters, distinct_stages, mode, d_cycles);
  return post_launch("mma_rate_kernel", st);
}

extern "C" int uwm_debug_handshake(int iters, int variant, int blocks, long long* d_cycles, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  handshake_kernel<<<blocks, 64, 0, st>>>(iters, variant, d_cycles);
  return post_launch("handshake_kernel", st);
}
#endif  // UWM_BENCH_TOOLS
