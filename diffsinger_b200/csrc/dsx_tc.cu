// tcgen05 / TMEM / TMA path of the dsx sampler (sm_100a): one fused kernel per residual layer
// (usr/diff/net.py:66-78):
//
//   GEMM1  D1[128 frames x 512] = [y(t-d) | y(t) | y(t+d) | cond(t)] (K = 1024) . W1^T      y = x + d_l
//   epi1   z = sigmoid(D1[:, gate] + b) * tanh(D1[:, filter] + b)  -> fp16 (hi, lo) in shared memory
//   GEMM2  D2[128 x 512] = z (K = 256) . W2^T
//   epi2   x <- (x + D2[:, :256] + b) / sqrt2 ;  y_next = fp16 split of (x + d_{l+1}) ;  skip += D2[:, 256:] + b
//
// Layout: activations are frames-major ([B][Tp][256], Tp = T rounded up to 128) so a 128-frame tile
// of 64 channels is one TMA box that lands in shared memory as the canonical K-major SWIZZLE_128B
// UMMA operand (rows of 128 B, 8-row atoms 1024 B apart).  The dilated taps are three boxes of the
// same tensor at frame offsets -d, 0, +d; TMA zero-fills rows outside [0, T), which is exactly the
// conv's zero padding applied after the FiLM add.  Weights are pre-packed into 256x64 fp16 tiles
// (32 KB) in the order the K loop consumes them; the gate/filter rows of a 256-wide N chunk are
// interleaved as [128 gate | 128 filter] so TMEM columns j and j+128 belong to the same channel.
//
// Precision: P = 1 uses fp16 operands (fp32 accumulate); P = 3 accumulates A_hi*W_hi + A_hi*W_lo +
// A_lo*W_hi into the same TMEM tile (hi/lo fp16 split, ~2^-22 relative) -- the K loop is simply 3x
// longer.
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread; the leader CTA only when
// cta_group::2), warp 2 = TMEM allocator, warps 4-7 = epilogue (thread = frame row = TMEM lane).
// cta_group::2 (G = 2): a cluster of two CTAs, each with its own 128 frames (UMMA M = 256); every CTA
// loads half of each weight tile, halving weight traffic from L2 and shared-memory operand reads.
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include <cmath>
#include <vector>

#include "dsx_internal.h"
#include "dsx_ptx.cuh"

namespace dsx {

constexpr int kC = 256;            // residual / conditioner channels supported by this path
constexpr int kRowsPerLayer = 80 * 256;   // wpack rows (of 64 fp16) per layer: 64 W1 tiles + 16 W2 tiles

template <int G, int P>
struct TcCfg {
  static constexpr int A_BYTES = kTile * 128;                 // 128 frames x 64 fp16
  static constexpr int W_ROWS = 256 / G;                      // rows of a weight tile held by one CTA
  static constexpr int W_BYTES = W_ROWS * 128;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int Z_PLANES = (P == 1) ? 1 : 2;
  static constexpr int Z_BYTES = Z_PLANES * 4 * A_BYTES;      // z [planes][4 k-blocks][128 x 64]
  static constexpr int STAGES = (G == 2) ? (P == 1 ? 4 : 3) : (P == 1 ? 3 : 2);
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + Z_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct TcLayerParams {
  CUtensorMap tm_w;        // packed weights, 2D [rows][64]
  CUtensorMap tm_y[2];     // this layer's conv input, planes hi/lo, 3D [B][T][256]
  CUtensorMap tm_cond[2];  // conditioner, planes hi/lo
  float* X;                // [B][Tp][256] residual stream (in/out)
  float* SKIP;             // [B][Tp][256]
  __half* Yout;            // next layer's conv input, plane 0; plane 1 at + plane_elems
  size_t plane_elems;
  const float* b1p;        // [2][256] this layer (packed order)
  const float* b2;         // [512]    this layer
  const float* dnext;      // FiLM vector of the next layer (row base), or nullptr on the last layer
  int d_row_stride;        // floats between utterances' rows in dnext
  int T, Tp, tiles_per_utt, tiles, B;
  int dil;
  int w_row0;              // first wpack row of this layer
  int skip_init;           // 1: skip = value, 0: skip += value
  int* status;
  unsigned long long budget_ns;
  long long* trace;        // debug: [2 CTAs][3 roles][256] clock64 stamps, or nullptr
};

__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float sigmoid_acc(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  // 2*sigmoid(2x) - 1, absolute error ~2e-7
  return fmaf(2.f, rcp_approx(1.f + ex2_approx(-2.8853900817779268f * x)), -1.f);
}

#define DSX_TRACE(role, slot)                                                              \
  do {                                                                                     \
    if (p.trace && blockIdx.x < 2 && (slot) < 256)                                          \
      p.trace[(blockIdx.x * 3 + (role)) * 256 + (slot)] = clock64();                       \
  } while (0)

template <int G, int P>
__global__ void __launch_bounds__(256, 1) k_tc_layer(const __grid_constant__ TcLayerParams p) {
  using Cfg = TcCfg<G, P>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* zbuf = base + S * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(zbuf + Cfg::Z_BYTES);
  uint64_t* full = bars;            // [S]
  uint64_t* empty = bars + S;       // [S]
  uint64_t* tfull = bars + 2 * S;   // [2]
  uint64_t* tempty = tfull + 2;     // [2]
  uint64_t* zfull = tempty + 2;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zfull + 1);
  auto stageA = [&](int s) { return base + s * Cfg::STAGE_BYTES; };
  auto stageW = [&](int s) { return base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (G == 2) ? cluster_ctarank() : 0;
  const int tile = blockIdx.x;      // grid is padded to a multiple of G; tiles >= p.tiles are dummies
  const bool tile_valid = tile < p.tiles;
  const int b = tile_valid ? tile / p.tiles_per_utt : p.B;          // b == B -> every TMA row is out of bounds
  const int t0 = tile_valid ? (tile % p.tiles_per_utt) * kTile : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_y[0]);
    tma_prefetch_desc(&p.tm_cond[0]);
    if (P == 3) {
      tma_prefetch_desc(&p.tm_y[1]);
      tma_prefetch_desc(&p.tm_cond[1]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * G);
    }
    mbar_init(zfull, 4 * G);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<G>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (G == 2) {
    cluster_arrive();
    cluster_wait();
  }
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  Watchdog wd{p.status, globaltimer_ns() + p.budget_ns};
  if (threadIdx.x == 0) DSX_TRACE(0, 254);

  if (warp == 0 && lane == 0) {
    // ================================ TMA producer ================================
    uint32_t it = 0;
    bool ok = true;
    for (int h = 0; h < 2 && ok; ++h)
      for (int pp = 0; pp < P && ok; ++pp)
        for (int kb = 0; kb < 16 && ok; ++kb, ++it) {
          const int s = it % S;
          ok = mbar_wait(&empty[s], ((it / S) & 1) ^ 1, wd, 101);
          if (!ok) break;
          DSX_TRACE(0, it);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], G * Cfg::STAGE_BYTES);
          const int aplane = (pp == 2) ? 1 : 0, wplane = (pp == 1) ? 1 : 0;
          if (kb < 12)
            tma_load_3d<G>(&p.tm_y[aplane], &full[s], stageA(s), (kb & 3) * 64, t0 + ((kb >> 2) - 1) * p.dil, b);
          else
            tma_load_3d<G>(&p.tm_cond[aplane], &full[s], stageA(s), (kb - 12) * 64, t0, b);
          const int wrow = p.w_row0 + ((wplane * 2 + h) * 16 + kb) * 256 + rank * Cfg::W_ROWS;
          tma_load_2d<G>(&p.tm_w, &full[s], stageW(s), 0, wrow);
        }
    for (int q = 0; q < 2 && ok; ++q)
      for (int pp = 0; pp < P && ok; ++pp)
        for (int kb = 0; kb < 4 && ok; ++kb, ++it) {
          const int s = it % S;
          ok = mbar_wait(&empty[s], ((it / S) & 1) ^ 1, wd, 102);
          if (!ok) break;
          DSX_TRACE(0, it);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], G * Cfg::W_BYTES);
          const int wplane = (pp == 1) ? 1 : 0;
          const int wrow = p.w_row0 + (64 + (wplane * 2 + q) * 4 + kb) * 256 + rank * Cfg::W_ROWS;
          tma_load_2d<G>(&p.tm_w, &full[s], stageW(s), 0, wrow);
        }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = umma_idesc_f16(128 * G, 256);
    uint32_t it = 0, tuse[2] = {0, 0};
    bool ok = true;
    for (int h = 0; h < 2 && ok; ++h) {
      const int buf = h;
      ok = mbar_wait(&tempty[buf], (tuse[buf] & 1) ^ 1, wd, 201);
      if (!ok) break;
      tuse[buf]++;
      tc_fence_after();
      const uint32_t d = tmem_base + buf * 256;
      uint32_t acc = 0;
      for (int pp = 0; pp < P && ok; ++pp)
        for (int kb = 0; kb < 16 && ok; ++kb, ++it) {
          const int s = it % S;
          ok = mbar_wait(&full[s], (it / S) & 1, wd, 202);
          if (!ok) break;
          DSX_TRACE(1, it);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(stageA(s)));
          const uint64_t bd = umma_desc_sw128(smem_u32(stageW(s)));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_f16<G>(d, ad + 2 * k4, bd + 2 * k4, idesc, acc);
            acc = 1;
          }
          umma_commit<G>(&empty[s]);
        }
      if (ok) umma_commit<G>(&tfull[buf]);
    }
    DSX_TRACE(1, 200);
    if (ok) ok = mbar_wait(zfull, 0, wd, 203);
    DSX_TRACE(1, 201);
    tc_fence_after();
    for (int q = 0; q < 2 && ok; ++q) {
      const int buf = q;
      ok = mbar_wait(&tempty[buf], (tuse[buf] & 1) ^ 1, wd, 204);
      if (!ok) break;
      DSX_TRACE(1, 202 + q);
      tuse[buf]++;
      tc_fence_after();
      const uint32_t d = tmem_base + buf * 256;
      uint32_t acc = 0;
      for (int pp = 0; pp < P && ok; ++pp)
        for (int kb = 0; kb < 4 && ok; ++kb, ++it) {
          const int s = it % S;
          ok = mbar_wait(&full[s], (it / S) & 1, wd, 205);
          if (!ok) break;
          DSX_TRACE(1, it);
          tc_fence_after();
          const int zplane = (pp == 2) ? 1 : 0;
          const uint64_t ad = umma_desc_sw128(smem_u32(zbuf + (zplane * 4 + kb) * Cfg::A_BYTES));
          const uint64_t bd = umma_desc_sw128(smem_u32(stageW(s)));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            umma_f16<G>(d, ad + 2 * k4, bd + 2 * k4, idesc, acc);
            acc = 1;
          }
          umma_commit<G>(&empty[s]);
        }
      if (ok) umma_commit<G>(&tfull[buf]);
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                 // frame row in the tile == TMEM lane
    const uint32_t tlane = static_cast<uint32_t>(quad * 32) << 16;
    const int t = t0 + r;
    const bool row_valid = tile_valid && t < p.T;
    const size_t rowoff = (static_cast<size_t>(b) * p.Tp + t) * kC;
    uint32_t tf[2] = {0, 0};
    bool ok = true;
    // ---- epi1: gate ----
    for (int h = 0; h < 2 && ok; ++h) {
      if (warp == 4 && lane == 0) DSX_TRACE(2, h * 4 + 0);
      ok = mbar_wait(&tfull[h], tf[h] & 1, wd, 301);
      if (!ok) break;
      if (warp == 4 && lane == 0) DSX_TRACE(2, h * 4 + 1);
      tf[h]++;
      tc_fence_after();
      const float* bg = p.b1p + h * 256;
#pragma unroll 1
      for (int j = 0; j < 128; j += 32) {
        uint32_t g[32], f[32];
        tmem_ld_32x32(tmem_base + tlane + h * 256 + j, g);
        tmem_ld_32x32(tmem_base + tlane + h * 256 + 128 + j, f);
        tmem_ld_wait();
        const int ch0 = h * 128 + j;                 // first z channel of this group
        const int kb = ch0 >> 6;
        uint8_t* zrow = zbuf + kb * Cfg::A_BYTES + r * 128;
        const int chunk0 = (ch0 & 63) >> 3;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float z2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = c8 * 8 + e * 2 + u;
              const float vg = __uint_as_float(g[i]) + __ldg(bg + j + i);
              const float vf = __uint_as_float(f[i]) + __ldg(bg + 128 + j + i);
              z2[u] = (P == 1) ? sigmoid_fast(vg) * tanh_approx(vf) : sigmoid_acc(vg) * tanh_acc(vf);
            }
            __half h0 = __float2half_rn(z2[0]), h1 = __float2half_rn(z2[1]);
            hi[e] = pack_h2(h0, h1);
            if (P == 3)
              lo[e] = pack_h2(__float2half_rn(z2[0] - __half2float(h0)), __float2half_rn(z2[1] - __half2float(h1)));
          }
          const int off = ((chunk0 + c8) ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(zrow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (P == 3)
            *reinterpret_cast<uint4*>(zrow + 4 * Cfg::A_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (warp == 4 && lane == 0) DSX_TRACE(2, h * 4 + 2);
      if (lane == 0) {
        if (G == 2) {
          mbar_arrive_cluster(&tempty[h], 0);
          if (h == 1) mbar_arrive_cluster(zfull, 0);
        } else {
          mbar_arrive(&tempty[h]);
          if (h == 1) mbar_arrive(zfull);
        }
      }
    }
    // ---- epi2, residual half: x <- (x + o + b)/sqrt2 ; y_next ----
    if (warp == 4 && lane == 0) DSX_TRACE(2, 8);
    if (ok) ok = mbar_wait(&tfull[0], tf[0] & 1, wd, 302);
    if (warp == 4 && lane == 0) DSX_TRACE(2, 9);
    if (ok) {
      tf[0]++;
      tc_fence_after();
      const float* dn = p.dnext ? p.dnext + static_cast<size_t>(tile_valid ? b : 0) * p.d_row_stride : nullptr;
#pragma unroll 1
      for (int j = 0; j < 256; j += 32) {
        uint32_t o[32];
        tmem_ld_32x32(tmem_base + tlane + j, o);
        tmem_ld_wait();
        if (row_valid) {
          float4* xp = reinterpret_cast<float4*>(p.X + rowoff + j);
          float xn[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            float4 xv = xp[v];
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + j) + v);
            xv.x = (xv.x + (__uint_as_float(o[4 * v + 0]) + bb.x)) * 0.70710678118654752440f;
            xv.y = (xv.y + (__uint_as_float(o[4 * v + 1]) + bb.y)) * 0.70710678118654752440f;
            xv.z = (xv.z + (__uint_as_float(o[4 * v + 2]) + bb.z)) * 0.70710678118654752440f;
            xv.w = (xv.w + (__uint_as_float(o[4 * v + 3]) + bb.w)) * 0.70710678118654752440f;
            xp[v] = xv;
            xn[4 * v + 0] = xv.x; xn[4 * v + 1] = xv.y; xn[4 * v + 2] = xv.z; xn[4 * v + 3] = xv.w;
          }
          if (dn) {
            __half* y0 = p.Yout + rowoff + j;
            __half* y1 = y0 + p.plane_elems;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = c8 * 8 + e * 2;
                const float ya = xn[i] + __ldg(dn + j + i), yb = xn[i + 1] + __ldg(dn + j + i + 1);
                __half ha = __float2half_rn(ya), hb = __float2half_rn(yb);
                hi[e] = pack_h2(ha, hb);
                lo[e] = pack_h2(__float2half_rn(ya - __half2float(ha)), __float2half_rn(yb - __half2float(hb)));
              }
              reinterpret_cast<uint4*>(y0)[c8] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (P == 3) reinterpret_cast<uint4*>(y1)[c8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (G == 2) mbar_arrive_cluster(&tempty[0], 0); else mbar_arrive(&tempty[0]);
      }
    }
    // ---- epi2, skip half ----
    if (warp == 4 && lane == 0) DSX_TRACE(2, 10);
    if (ok) ok = mbar_wait(&tfull[1], tf[1] & 1, wd, 303);
    if (warp == 4 && lane == 0) DSX_TRACE(2, 11);
    if (ok) {
      tf[1]++;
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 256; j += 32) {
        uint32_t o[32];
        tmem_ld_32x32(tmem_base + tlane + 256 + j, o);
        tmem_ld_wait();
        if (row_valid) {
          float4* sp = reinterpret_cast<float4*>(p.SKIP + rowoff + j);
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + 256 + j) + v);
            float4 sv = p.skip_init ? make_float4(0.f, 0.f, 0.f, 0.f) : sp[v];
            sv.x += __uint_as_float(o[4 * v + 0]) + bb.x;
            sv.y += __uint_as_float(o[4 * v + 1]) + bb.y;
            sv.z += __uint_as_float(o[4 * v + 2]) + bb.z;
            sv.w += __uint_as_float(o[4 * v + 3]) + bb.w;
            sp[v] = sv;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (G == 2) mbar_arrive_cluster(&tempty[1], 0); else mbar_arrive(&tempty[1]);
      }
    }
  }

  if (warp == 4 && lane == 0) DSX_TRACE(2, 12);
  if (threadIdx.x == 0) DSX_TRACE(0, 255);
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (G == 2) {
    cluster_arrive();
    cluster_wait();
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<G>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// weight packing for the tcgen05 path
// tile order per layer (tiles of 256 rows x 64 k, fp16):
//   W1: idx = (plane*2 + chunk)*16 + kb      plane 0 = hi, 1 = lo; chunk h: rows n<128 -> gate channel
//       128h+n, n>=128 -> filter channel 128h+n-128 (= conv output row C + 128h + n - 128); k = kb*64+kk
//       over [tap0 | tap1 | tap2 | cond]
//   W2: idx = 64 + (plane*2 + half)*4 + kb   rows n -> output row half*256 + n
// ------------------------------------------------------------------------------------------
__global__ void k_pack_wtc(const float* __restrict__ w1f, const float* __restrict__ w2f,
                           const float* __restrict__ b1f, __half* __restrict__ wpack, float* __restrict__ b1p) {
  const int l = blockIdx.y, tileidx = blockIdx.x, n = threadIdx.x;
  const float* src;
  int plane;
  if (tileidx < 64) {
    plane = tileidx / 32;
    const int h = (tileidx / 16) & 1, kb = tileidx & 15;
    const int j = (n < 128) ? (128 * h + n) : (kC + 128 * h + (n - 128));
    src = w1f + (static_cast<size_t>(l) * 2 * kC + j) * (4 * kC) + kb * 64;
    if (plane == 0 && kb == 0) b1p[(static_cast<size_t>(l) * 2 + h) * 256 + n] = b1f[static_cast<size_t>(l) * 2 * kC + j];
  } else {
    const int u = tileidx - 64;
    plane = u / 8;
    const int q = (u / 4) & 1, kb = u & 3;
    src = w2f + (static_cast<size_t>(l) * 2 * kC + q * 256 + n) * kC + kb * 64;
  }
  __half* dst = wpack + ((static_cast<size_t>(l) * 80 + tileidx) * 256 + n) * 64;
  for (int kk = 0; kk < 64; ++kk) {
    const float v = src[kk];
    const __half hi = __float2half_rn(v);
    dst[kk] = plane == 0 ? hi : __float2half_rn(v - __half2float(hi));
  }
}

bool tc_supported(const dsx_handle* h) { return h->m.C == kC && h->m.H == kC; }

int tc_pack_model(dsx_handle* h, cudaStream_t s) {
  DSX_CHECK(tc_supported(h), DSX_E_INVALID, "tcgen05 path needs residual_channels == hidden_size == 256 (got %d, %d)",
            h->m.C, h->m.H);
  __half* wpack;
  float* b1p;
  const size_t rows = static_cast<size_t>(h->m.L) * kRowsPerLayer;
  DSX_TRY(dev_alloc(h, reinterpret_cast<void**>(&wpack), rows * 64 * sizeof(__half), true));
  DSX_TRY(dev_alloc(h, reinterpret_cast<void**>(&b1p), static_cast<size_t>(h->m.L) * 512 * sizeof(float), true));
  dim3 grid(80, h->m.L);
  k_pack_wtc<<<grid, 256, 0, s>>>(h->m.w1f, h->m.w2f, h->m.b1f, wpack, b1p);
  h->launches++;
  DSX_CUDA(cudaGetLastError());
  h->m.wpack = wpack;
  h->m.b1p = b1p;
  return DSX_OK;
}

// ------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

static int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint32_t box_rows) {
  PFN_tmapEncodeTiled enc = get_encode();
  DSX_CHECK(enc, DSX_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "cuTensorMapEncodeTiled(2D) failed: %d", static_cast<int>(r));
  return DSX_OK;
}
// [B][T (stride Tp)][ch] fp16, box = 64 channels x 128 frames x 1
static int make_map_act(CUtensorMap* m, const void* base, int ch, int T, int Tp, int B) {
  PFN_tmapEncodeTiled enc = get_encode();
  DSX_CHECK(enc, DSX_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(ch), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ch) * 2, static_cast<cuuint64_t>(Tp) * ch * 2};
  cuuint32_t box[3] = {64, kTile, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "cuTensorMapEncodeTiled(3D) failed: %d", static_cast<int>(r));
  return DSX_OK;
}

int tc_prepare_maps(dsx_handle* h, const Geom& g) {
  const size_t plane = g.frames_padded() * kC;
  if (h->tm_geom.B == g.B && h->tm_geom.T == g.T && h->tm_base_y == h->ws.Y && h->tm_base_cond == h->ws.CONDH &&
      h->tm_group == h->tc_group)
    return DSX_OK;
  DSX_TRY(make_map_2d(&h->tm_w, h->m.wpack, static_cast<uint64_t>(h->m.L) * kRowsPerLayer, 256 / h->tc_group));
  for (int buf = 0; buf < 2; ++buf)
    for (int pl = 0; pl < 2; ++pl)
      DSX_TRY(make_map_act(&h->tm_y[buf][pl], h->ws.Y + (static_cast<size_t>(buf) * 2 + pl) * plane, kC, g.T, g.Tp, g.B));
  for (int pl = 0; pl < 2; ++pl)
    DSX_TRY(make_map_act(&h->tm_cond[pl], h->ws.CONDH + static_cast<size_t>(pl) * plane, kC, g.T, g.Tp, g.B));
  h->tm_geom = g;
  h->tm_base_y = h->ws.Y;
  h->tm_base_cond = h->ws.CONDH;
  h->tm_group = h->tc_group;
  return DSX_OK;
}

template <int G, int P>
static int launch_tc_layer_t(dsx_handle* h, const TcLayerParams& prm, int tiles, cudaStream_t s) {
  using Cfg = TcCfg<G, P>;
  static bool attr_done = false;
  if (!attr_done) {
    DSX_CUDA(cudaFuncSetAttribute(k_tc_layer<G, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>((tiles + G - 1) / G * G));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DSX_CUDA(cudaLaunchKernelEx(&cfg, k_tc_layer<G, P>, prm));
  h->launches++;
  return DSX_OK;
}

int launch_tc_layer(dsx_handle* h, int layer, const Geom& g, int row0, int row_per_b, cudaStream_t s) {
  const ModelDev& m = h->m;
  TcLayerParams prm;
  memset(&prm, 0, sizeof(prm));
  const int cur = layer & 1;
  prm.tm_w = h->tm_w;
  prm.tm_y[0] = h->tm_y[cur][0];
  prm.tm_y[1] = h->tm_y[cur][1];
  prm.tm_cond[0] = h->tm_cond[0];
  prm.tm_cond[1] = h->tm_cond[1];
  prm.X = h->ws.X;
  prm.SKIP = h->ws.SKIP;
  prm.plane_elems = g.frames_padded() * kC;
  prm.Yout = h->ws.Y + static_cast<size_t>((cur ^ 1) * 2) * prm.plane_elems;
  prm.b1p = m.b1p + static_cast<size_t>(layer) * 512;
  prm.b2 = m.b2f + static_cast<size_t>(layer) * 512;
  prm.dnext = (layer + 1 < m.L) ? h->ws.DTAB + (static_cast<size_t>(row0) * m.L + layer + 1) * kC : nullptr;
  prm.d_row_stride = row_per_b * m.L * kC;
  prm.T = g.T; prm.Tp = g.Tp; prm.tiles_per_utt = g.tiles_per_utt; prm.tiles = g.tiles; prm.B = g.B;
  prm.dil = 1 << (layer % m.cycle);
  prm.w_row0 = layer * kRowsPerLayer;
  prm.skip_init = (layer == 0);
  prm.status = h->status_dev;
  prm.budget_ns = 2000000000ull;
  prm.trace = h->trace_dev;
  const int P = (h->precision == DSX_PREC_FP16) ? 1 : 3;
  if (h->tc_group == 2) {
    return P == 1 ? launch_tc_layer_t<2, 1>(h, prm, g.tiles, s) : launch_tc_layer_t<2, 3>(h, prm, g.tiles, s);
  }
  return P == 1 ? launch_tc_layer_t<1, 1>(h, prm, g.tiles, s) : launch_tc_layer_t<1, 3>(h, prm, g.tiles, s);
}

}  // namespace dsx
