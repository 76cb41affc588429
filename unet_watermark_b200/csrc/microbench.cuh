// Micro-benchmarks used to size the conv kernel (tools/gpu_microbench.py); not on the product path.
#pragma once
#include "ptx_sm100.cuh"

namespace uwm {

// Raw tcgen05.mma issue/execute rate: one warp issues `iters` back-to-back M=128 x N x K=16 bf16 MMAs on
// (uninitialised) SW128 K-major smem tiles, commits once and waits.  cycles[blockIdx] = clock delta.
// mode bit0: tcgen05.commit to a scratch mbarrier after every 4 MMAs; bit1: also wait on an (already
// complete) mbarrier + tcgen05.fence before every 4 MMAs; bit2: ping-pong handshake with a second warp
// (it waits for the commit and re-arms a "full" barrier), ring depth = distinct_stages.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, int distinct_stages, int mode, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint64_t scratch[64];   // [0..31] "empty" (commit targets), [32..63] "full"
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 64; ++i) mbar_init(smem_u32(&scratch[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, n);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);
    const uint32_t lo0 = ((base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t a_units = 16384u >> 4, stage_units = (16384u + 32768u) >> 4;
    long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t a_lo = lo0 + s * stage_units, b_lo = a_lo + a_units;
      if (mode & 4) { mbar_wait(smem_u32(&scratch[32 + s]), ph); tc_fence_after(); }
      else if (mode & 2) { mbar_wait(smem_u32(&scratch[32 + s]), 1); tc_fence_after(); }   // fresh barrier: parity 1 passes
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(tmem, a_lo + 2u * k, b_lo + 2u * k, hi, idesc, 1u);
        if (mode & 5) umma_commit(smem_u32(&scratch[s]));
      }
      __syncwarp();
      if (++s == distinct_stages) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  if (warp == 2 && (mode & 4)) {      // "producer": recycle stages as the MMAs that used them complete
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&scratch[s]), ph ^ 1u);
      if (elect_one()) mbar_arrive(smem_u32(&scratch[32 + s]));
      __syncwarp();
      if (++s == distinct_stages) { s = 0; ph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace uwm

namespace uwm {
// mbarrier ping-pong between two warps: warp 0 arrives on b1 and waits on b2; warp 1 waits on b1 and
// arrives on b2.  variant bit0: spin on test_wait instead of try_wait; bit1: only lane 0 of each warp
// participates; bit2: ring of 4 independent barrier pairs (pipelined, depth 4) instead of 1.
__global__ void __launch_bounds__(64, 1) handshake_kernel(int iters, int variant, long long* cycles) {
  __shared__ uint64_t b1[4], b2[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&b1[i]), 1); mbar_init(smem_u32(&b2[i]), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool spin = variant & 1, one_lane = variant & 2;
  const int depth = (variant & 4) ? 4 : 1;
  if (one_lane && lane != 0) return;
  auto wait = [&](uint32_t bar, uint32_t ph) {
    if (spin) { while (!mbar_test_wait(bar, ph)) {} } else { mbar_wait(bar, ph); }
  };
  long long t0 = clock64();
  int s = 0; uint32_t ph = 0;
  if (warp == 0) {          // "producer": needs slot s free (b2), then fills it (b1)
    for (int i = 0; i < iters; ++i) {
      wait(smem_u32(&b2[s]), ph ^ 1u);
      if (one_lane || elect_one()) mbar_arrive(smem_u32(&b1[s]));
      if (!one_lane) __syncwarp();
      if (++s == depth) { s = 0; ph ^= 1u; }
    }
  } else {                  // "consumer": waits slot s full (b1), releases it (b2)
    for (int i = 0; i < iters; ++i) {
      wait(smem_u32(&b1[s]), ph);
      if (one_lane || elect_one()) mbar_arrive(smem_u32(&b2[s]));
      if (!one_lane) __syncwarp();
      if (++s == depth) { s = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
}
}  // namespace uwm

namespace uwm {
// Cost of single synchronisation primitives: one warp executes `iters` repetitions, cycles per repetition.
// which: 0 empty loop, 1 tcgen05.fence::before_thread_sync, 2 tcgen05.fence::after_thread_sync, 3 fence.proxy.async,
// 4 mbarrier.arrive (lane 0) on a count-1 barrier, 5 __syncwarp, 6 try_wait on a completed phase (lane 0),
// 7 lane-0 try_wait + __syncwarp, 8 tcgen05.commit to a barrier (lane 0), 9 clock64 + global store (lane 0)
__global__ void __launch_bounds__(32, 1) prim_cost_kernel(int which, int iters, long long* out) {
  __shared__ uint64_t bar[2];
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); fence_mbar_init(); }
  __syncwarp();
  const uint32_t b0 = smem_u32(&bar[0]);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    switch (which) {
      case 1: tc_fence_before(); break;
      case 2: tc_fence_after(); break;
      case 3: asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); break;
      case 4: if (lane == 0) mbar_arrive(b0); break;
      case 5: __syncwarp(); break;
      case 6: if (lane == 0) mbar_try_wait(b0, 1); break;
      case 7: if (lane == 0) mbar_try_wait(b0, 1); __syncwarp(); break;
      case 8: if (lane == 0) umma_commit(b0); break;
      case 9: if (lane == 0) out[8 + (i & 7)] = clock64(); break;
      default: asm volatile("" ::: "memory"); break;
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[0] = t1 - t0;
}
}  // namespace uwm
