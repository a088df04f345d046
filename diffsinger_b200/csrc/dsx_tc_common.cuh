// Constants and small device helpers shared by the tcgen05 kernels (dsx_tc.cu: per-layer / round-1 stack kernel, head,
// conditioner projection; dsx_stack.cu: the register-resident stack kernel).
#pragma once
#include "dsx_internal.h"
#include "dsx_ptx.cuh"

namespace dsx {

constexpr int kC = 256;            // residual / conditioner channels supported by this path
constexpr int kRowsPerLayer = 80 * 256;   // wpack rows (of 64 fp16) per layer: 64 W1 tiles + 16 W2 tiles

constexpr int kG = 2;                      // cta_group of the layer kernel (cluster of two CTAs)
constexpr int kUnitBytes = kTile * 128;    // ring unit: 128 rows x 64 fp16 (one A k-block tile, or one CTA's half of a W tile)
constexpr int kEpiWarps = 8;               // epilogue warps (two per TMEM lane quadrant, split by columns)
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kStageRowBytes = 48;         // epilogue-2 transpose staging: 8 fp32 + 16 B pad per row
constexpr int kStagingBytes = kEpiWarps * 32 * kStageRowBytes;

__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
// sigmoid(g) * tanh(f) = (1 - E2) / ((1 + E1)(1 + E2)), E1 = e^-g, E2 = e^-2f: three MUFU operations instead of four
// (the gate epilogue is MUFU-bound); absolute error ~2e-7.  f is clamped at -15 (tanh = -1 to 2e-13) so E2 stays finite.
__device__ __forceinline__ float gate_acc(float g, float f) {
  const float e1 = ex2_approx(-1.4426950408889634f * g);
  const float e2 = ex2_approx(-2.8853900817779268f * fmaxf(f, -15.f));
  return (1.f - e2) * rcp_approx((1.f + e1) * (1.f + e2));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

// Order in which GEMM1 (one tile per tap, P = 3) consumes its 12 k-blocks (k-block = tap*4 + channel block): the
// centre tap (this tile's own y) first, the halo taps (they need the neighbour tiles' y) last.
__device__ __forceinline__ int kb_order(int ko) { return ko < 4 ? 4 + ko : (ko < 8 ? ko - 4 : ko); }
constexpr size_t kCpChunk = 256 * kTile;   // floats of CP per (layer, tile, chunk)

// publish / wait on a tile's counter in global memory (gpu scope)
__device__ __forceinline__ void flag_publish(unsigned int* f) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
}
__device__ __forceinline__ bool flag_wait(const unsigned int* f, unsigned int target, const Watchdog& wd, int code) {
  uint32_t spins = 0;
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (static_cast<int>(v - target) >= 0) return true;
    if (((++spins) & 0xff) == 0) {
      if (*(volatile int*)wd.status != 0) return false;
      if (globaltimer_ns() > wd.deadline_ns) {
        atomicCAS(wd.status, 0, code);
        return false;
      }
    }
  }
}

// Wait until this tile's counter and its neighbours' (lo / hi may be null) have all reached `target`: the three polls
// travel to L2 together (relaxed loads), one gpu-scope fence turns the successful observation into an acquire.
__device__ __forceinline__ bool flag_wait3(const unsigned int* f, const unsigned int* lo, const unsigned int* hi,
                                           unsigned int target, const Watchdog& wd, int code) {
  uint32_t spins = 0;
  while (true) {
    unsigned int v0, v1 = target, v2 = target;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(f) : "memory");
    if (lo) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(lo) : "memory");
    if (hi) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(hi) : "memory");
    if (static_cast<int>(v0 - target) >= 0 && static_cast<int>(v1 - target) >= 0 && static_cast<int>(v2 - target) >= 0) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      return true;
    }
    if (((++spins) & 0xff) == 0) {
      if (*(volatile int*)wd.status != 0) return false;
      if (globaltimer_ns() > wd.deadline_ns) {
        atomicCAS(wd.status, 0, code);
        return false;
      }
    }
  }
}

}  // namespace dsx
