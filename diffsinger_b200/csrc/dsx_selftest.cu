// Hardware self-tests of the encodings the tcgen05 path depends on: the SWIZZLE_128B K-major shared
// memory descriptor, the kind::f16 instruction descriptor, TMA boxes with negative / out-of-range
// coordinates (zero fill), cta_group::1 and cta_group::2 MMA + multicast commit, and the 32x32b TMEM
// load mapping.  One tiny GEMM per variant: D[128*G x 256] = A[128*G x 128] . W[256 x 128]^T with A
// fetched from frame offset t0 = -3 (so the first rows must come back as zeros), checked exactly
// against a host computation (all values are small dyadic rationals, exact in fp16 / fp32).
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "dsx_internal.h"
#include "dsx_ptx.cuh"

namespace dsx {

struct SelfParams {
  CUtensorMap tm_a;   // 3D [1][T][128] fp16, box 64 x 128 x 1
  CUtensorMap tm_w;   // 2D [512 rows][64] fp16, box 64 x (256/G)
  float* out;         // [128*G][256]
  int t0;
  int* status;
};

template <int G>
__global__ void __launch_bounds__(256, 1) k_selftest(const __grid_constant__ SelfParams p) {
  constexpr int A_BYTES = 128 * 128, W_BYTES = (256 / G) * 128, STAGE = A_BYTES + W_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + 2 * STAGE);
  uint64_t* tfull = full + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (G == 2) ? cluster_ctarank() : 0;
  if (warp == 1 && lane == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<G>(slot, 256);
  tc_fence_before();
  __syncthreads();
  if (G == 2) { cluster_arrive(); cluster_wait(); }
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot);
  Watchdog wd{p.status, globaltimer_ns() + 500000000ull};
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < 2; ++kb) {
      if (rank == 0) mbar_arrive_expect_tx(&full[kb], G * STAGE);
      tma_load_3d<G>(&p.tm_a, &full[kb], base + kb * STAGE, kb * 64, p.t0 + 128 * static_cast<int>(rank), 0);
      tma_load_2d<G>(&p.tm_w, &full[kb], base + kb * STAGE + A_BYTES, 0, kb * 256 + rank * (256 / G));
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128 * G, 256);
    uint32_t acc = 0;
    bool ok = true;
    for (int kb = 0; kb < 2 && ok; ++kb) {
      ok = mbar_wait(&full[kb], 0, wd, 901);
      if (!ok) break;
      tc_fence_after();
      const uint64_t ad = umma_desc_sw128(smem_u32(base + kb * STAGE));
      const uint64_t bd = umma_desc_sw128(smem_u32(base + kb * STAGE + A_BYTES));
      for (int k4 = 0; k4 < 4; ++k4) {
        umma_f16<G>(tmem_base, ad + 2 * k4, bd + 2 * k4, idesc, acc);
        acc = 1;
      }
    }
    if (ok) umma_commit<G>(tfull);
  } else if (warp >= 4) {
    const int quad = warp & 3, r = quad * 32 + lane;
    if (mbar_wait(tfull, 0, wd, 902)) {
      tc_fence_after();
      for (int j = 0; j < 256; j += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + j, v);
        tmem_ld_wait();
        float* o = p.out + (static_cast<size_t>(rank) * 128 + r) * 256 + j;
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (G == 2) { cluster_arrive(); cluster_wait(); }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<G>(tmem_base, 256);
  }
}

// ---- experiment (informational): can a K-major SWIZZLE_128B A descriptor start `shift` rows into a tile?
// (would let the three dilated taps share one shared-memory copy of the activations)  One CTA, cta_group::1:
// A box = 160 rows x 64 channels, D = A[shift : shift+128] . W^T with K = 64.
struct ShiftParams {
  CUtensorMap tm_a;   // 3D [1][T][128] fp16, box 64 x 160 x 1
  CUtensorMap tm_w;   // 2D [512][64], box 64 x 256
  float* out;         // [128][256]
  int shift, use_base_offset;
  int* status;
};

__global__ void __launch_bounds__(256, 1) k_shift_test(const __grid_constant__ ShiftParams p) {
  constexpr int A_BYTES = 160 * 128, W_BYTES = 256 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wbuf = base + 20480 + 1024;   // keep W 1024-aligned: 21504 = 21 * 1024
  uint64_t* full = reinterpret_cast<uint64_t*>(wbuf + W_BYTES);
  uint64_t* tfull = full + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    mbar_init(full, 1);
    mbar_init(tfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot);
  Watchdog wd{p.status, globaltimer_ns() + 500000000ull};
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(full, A_BYTES + W_BYTES);
    tma_load_3d<1>(&p.tm_a, full, base, 0, 0, 0);
    tma_load_2d<1>(&p.tm_w, full, wbuf, 0, 0);
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 256);
    if (mbar_wait(full, 0, wd, 911)) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(base) + p.shift * 128;
      uint64_t ad = umma_desc_sw128(a_addr);
      if (p.use_base_offset) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      const uint64_t bd = umma_desc_sw128(smem_u32(wbuf));
      uint32_t acc = 0;
      for (int k4 = 0; k4 < 4; ++k4) {
        umma_f16<1>(tmem_base, ad + 2 * k4, bd + 2 * k4, idesc, acc);
        acc = 1;
      }
      umma_commit<1>(tfull);
    }
  } else if (warp >= 4) {
    const int quad = warp & 3, r = quad * 32 + lane;
    if (mbar_wait(tfull, 0, wd, 912)) {
      tc_fence_after();
      for (int j = 0; j < 256; j += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + j, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) p.out[static_cast<size_t>(r) * 256 + j + i] = __uint_as_float(v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 256);
  }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

template <int G>
static int run_selftest(std::string& report) {
  const int T = 300, CH = 128, t0 = -3, M = 128 * G;
  std::vector<__half> ha(static_cast<size_t>(T) * CH), hw(512 * 64);
  auto aval = [](int t, int c) { return static_cast<float>((t * 131 + c * 71) % 61 - 30) / 64.f; };
  auto wval = [](int n, int k) { return static_cast<float>((n * 37 + k * 11) % 53 - 26) / 128.f; };
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < CH; ++c) ha[static_cast<size_t>(t) * CH + c] = __float2half(aval(t, c));
  for (int kb = 0; kb < 2; ++kb)
    for (int n = 0; n < 256; ++n)
      for (int kk = 0; kk < 64; ++kk) hw[(static_cast<size_t>(kb) * 256 + n) * 64 + kk] = __float2half(wval(n, kb * 64 + kk));
  __half *da = nullptr, *dw = nullptr;
  float* dout = nullptr;
  int* dstatus = nullptr;
  DSX_CUDA(cudaMalloc(&da, ha.size() * 2));
  DSX_CUDA(cudaMalloc(&dw, hw.size() * 2));
  DSX_CUDA(cudaMalloc(&dout, static_cast<size_t>(M) * 256 * 4));
  DSX_CUDA(cudaMalloc(&dstatus, 4));
  DSX_CUDA(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  DSX_CUDA(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  DSX_CUDA(cudaMemset(dout, 0xff, static_cast<size_t>(M) * 256 * 4));
  DSX_CUDA(cudaMemset(dstatus, 0, 4));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  DSX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  DSX_CHECK(fp && q == cudaDriverEntryPointSuccess, DSX_E_CUDA, "no cuTensorMapEncodeTiled");
  PFN_tmapEncodeTiled enc = reinterpret_cast<PFN_tmapEncodeTiled>(fp);
  SelfParams prm;
  memset(&prm, 0, sizeof(prm));
  {
    cuuint64_t dims[3] = {CH, T, 1};
    cuuint64_t strides[2] = {CH * 2, static_cast<cuuint64_t>(T) * CH * 2};
    cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&prm.tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, da, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "selftest: encode A map failed %d", static_cast<int>(r));
    cuuint64_t d2[2] = {64, 512};
    cuuint64_t s2[1] = {128};
    cuuint32_t b2[2] = {64, 256 / G}, e2[2] = {1, 1};
    r = enc(&prm.tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "selftest: encode W map failed %d", static_cast<int>(r));
  }
  prm.out = dout;
  prm.t0 = t0;
  prm.status = dstatus;
  const int smem = 1024 + 2 * (128 * 128 + (256 / G) * 128) + 64;
  DSX_CUDA(cudaFuncSetAttribute(k_selftest<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DSX_CUDA(cudaLaunchKernelEx(&cfg, k_selftest<G>, prm));
  DSX_CUDA(cudaDeviceSynchronize());
  std::vector<float> out(static_cast<size_t>(M) * 256);
  int status = 0;
  DSX_CUDA(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  DSX_CUDA(cudaMemcpy(&status, dstatus, 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(dw); cudaFree(dout); cudaFree(dstatus);
  double maxerr = 0;
  int bad = 0, first_bad_m = -1, first_bad_n = -1;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < 256; ++n) {
      double ref = 0;
      int t = t0 + m;
      if (t >= 0 && t < T)
        for (int k = 0; k < CH; ++k) ref += static_cast<double>(aval(t, k)) * wval(n, k);
      double e = fabs(ref - out[static_cast<size_t>(m) * 256 + n]);
      if (!(e <= 1e-4)) {
        if (!bad) { first_bad_m = m; first_bad_n = n; }
        bad++;
      }
      if (e > maxerr || e != e) maxerr = e;
    }
  char line[256];
  snprintf(line, sizeof(line), "umma_cta_group%d: status=%d bad=%d/%d maxerr=%.3g first_bad=(%d,%d)\n", G, status, bad,
           M * 256, maxerr, first_bad_m, first_bad_n);
  report += line;
  return (status == 0 && bad == 0) ? DSX_OK : DSX_E_KERNEL;
}

// Row-shifted operand descriptors (what the dilated taps use: base_offset field 0): every shift must reproduce the reference.
// `also_base_offset_variant` additionally reports the alternative encoding (informational).
static int run_shift_experiment(std::string& report, bool also_base_offset_variant = false) {
  const int T = 300, CH = 128;
  std::vector<__half> ha(static_cast<size_t>(T) * CH), hw(512 * 64);
  auto aval = [](int t, int c) { return static_cast<float>((t * 131 + c * 71) % 61 - 30) / 64.f; };
  auto wval = [](int n, int k) { return static_cast<float>((n * 37 + k * 11) % 53 - 26) / 128.f; };
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < CH; ++c) ha[static_cast<size_t>(t) * CH + c] = __float2half(aval(t, c));
  for (int n = 0; n < 256; ++n)
    for (int kk = 0; kk < 64; ++kk) hw[static_cast<size_t>(n) * 64 + kk] = __float2half(wval(n, kk));
  __half *da = nullptr, *dw = nullptr;
  float* dout = nullptr;
  int* dstatus = nullptr;
  DSX_CUDA(cudaMalloc(&da, ha.size() * 2));
  DSX_CUDA(cudaMalloc(&dw, hw.size() * 2));
  DSX_CUDA(cudaMalloc(&dout, 128 * 256 * 4));
  DSX_CUDA(cudaMalloc(&dstatus, 4));
  DSX_CUDA(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  DSX_CUDA(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  DSX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  PFN_tmapEncodeTiled enc = reinterpret_cast<PFN_tmapEncodeTiled>(fp);
  ShiftParams prm;
  memset(&prm, 0, sizeof(prm));
  {
    cuuint64_t dims[3] = {CH, T, 1};
    cuuint64_t strides[2] = {CH * 2, static_cast<cuuint64_t>(T) * CH * 2};
    cuuint32_t box[3] = {64, 160, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&prm.tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, da, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "shift experiment: encode A map failed %d", static_cast<int>(r));
    cuuint64_t d2[2] = {64, 512};
    cuuint64_t s2[1] = {128};
    cuuint32_t b2[2] = {64, 256}, e2[2] = {1, 1};
    r = enc(&prm.tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "shift experiment: encode W map failed %d", static_cast<int>(r));
  }
  prm.out = dout;
  prm.status = dstatus;
  const int smem = 1024 + 21504 + 256 * 128 + 64;
  DSX_CUDA(cudaFuncSetAttribute(k_shift_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int shifts[] = {0, 1, 2, 3, 4, 8, 16};
  int failures = 0;
  for (int ubo = 0; ubo < (also_base_offset_variant ? 2 : 1); ++ubo)
    for (int si = 0; si < 7; ++si) {
      prm.shift = shifts[si];
      prm.use_base_offset = ubo;
      DSX_CUDA(cudaMemset(dout, 0xff, 128 * 256 * 4));
      DSX_CUDA(cudaMemset(dstatus, 0, 4));
      k_shift_test<<<1, 256, smem>>>(prm);
      DSX_CUDA(cudaDeviceSynchronize());
      std::vector<float> out(128 * 256);
      DSX_CUDA(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 256; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += static_cast<double>(aval(m + shifts[si], k)) * wval(n, k);
          if (!(fabs(ref - out[static_cast<size_t>(m) * 256 + n]) <= 1e-4)) bad++;
        }
      char line[160];
      snprintf(line, sizeof(line), "shifted_A_desc shift=%d base_offset_field=%d: bad=%d/32768\n", shifts[si], ubo, bad);
      report += line;
      if (ubo == 0 && bad) failures++;
    }
  cudaFree(da); cudaFree(dw); cudaFree(dout); cudaFree(dstatus);
  return failures ? DSX_E_KERNEL : DSX_OK;
}


#ifdef DSX_EXPERIMENTS   // informational micro-benchmark, not part of the product library (build with -DDSX_EXPERIMENTS)
// ---- experiment 3: how fast can ONE SM pull L2-resident operand tiles into shared memory? --------------------
// `nthr` threads (one per warp) each stream loads through their own ring of `nslot` slots (wait for the slot's previous
// load, re-issue).  kind 0: 2D tensor map (64 x rows box, SWIZZLE_128B) -- what the layer kernel does; kind 1: 1D bulk
// copy (cp.async.bulk, no tensor map); kind 2: 3D tensor map box 64 x 128 x 2 (two 16 KB tiles per instruction).
struct IngestParams {
  CUtensorMap tm[4];       // box rows 16, 32, 64, 128
  CUtensorMap tm3;         // 3D [64][128][tiles], box 64 x 128 x 2
  const uint8_t* src;
  long long* cycles;
  int kind, rows, nslot, nthr, nloads, src_tiles;
};
__global__ void __launch_bounds__(128, 1) k_ingest(const __grid_constant__ IngestParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NSLOT = 12;                                   // x 16 KB, shared by the issuing threads
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + NSLOT * 16384);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < p.nthr) {
    const uint32_t bytes = (p.kind == 2) ? 32768u : static_cast<uint32_t>(p.rows) * 128u;
    const int per = p.nslot / p.nthr;                         // slots of this thread: [w*per, (w+1)*per)
    const int stride = (p.kind == 2) ? 2 : 1;                 // 16 KB cells per slot
    const CUtensorMap* tm = &p.tm[p.rows == 16 ? 0 : (p.rows == 32 ? 1 : (p.rows == 64 ? 2 : 3))];
    const long long t0 = clock64();
    const int n = p.nloads / p.nthr;
    for (int i = 0; i < n; ++i) {
      const int sl = i % per, use = i / per, s = (w * per + sl) * stride;
      if (use > 0)
        for (uint32_t sp = 0; sp < (1u << 22) && !mbar_try_wait(&bars[s], (use - 1) & 1); ++sp) {}
      mbar_arrive_expect_tx(&bars[s], bytes);
      const int tile = (i * 7 + blockIdx.x * 3 + w * 11) % (p.src_tiles - 1);
      uint8_t* dst = base + s * 16384;
      if (p.kind == 0) tma_load_2d<1>(tm, &bars[s], dst, 0, tile * 128);
      else if (p.kind == 2) tma_load_3d<1>(&p.tm3, &bars[s], dst, 0, 0, tile);
      else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(p.src + static_cast<size_t>(tile) * 16384), "r"(bytes), "r"(smem_u32(&bars[s]))
                     : "memory");
    }
    for (int i = max(n - per, 0); i < n; ++i) {
      const int sl = i % per, use = i / per, s = (w * per + sl) * stride;
      for (uint32_t sp = 0; sp < (1u << 22) && !mbar_try_wait(&bars[s], use & 1); ++sp) {}
    }
    p.cycles[blockIdx.x * 4 + w] = clock64() - t0;
  }
}

static int run_ingest_experiment(std::string& report) {
  const int src_tiles = 256;                                  // 4 MB source: L2 resident
  uint8_t* dsrc = nullptr;
  long long* dcyc = nullptr;
  DSX_CUDA(cudaMalloc(&dsrc, static_cast<size_t>(src_tiles) * 16384));
  DSX_CUDA(cudaMemset(dsrc, 0, static_cast<size_t>(src_tiles) * 16384));
  DSX_CUDA(cudaMalloc(&dcyc, 1024 * sizeof(long long)));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  DSX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  DSX_CHECK(fp && q == cudaDriverEntryPointSuccess, DSX_E_CUDA, "no cuTensorMapEncodeTiled");
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fp);
  IngestParams prm;
  memset(&prm, 0, sizeof(prm));
  cuuint64_t d2[2] = {64, static_cast<cuuint64_t>(src_tiles) * 128};
  cuuint64_t s2[1] = {128};
  cuuint32_t e3[3] = {1, 1, 1};
  const int rows_opt[4] = {16, 32, 64, 128};
  for (int i = 0; i < 4; ++i) {
    cuuint32_t bx[2] = {64, static_cast<cuuint32_t>(rows_opt[i])};
    CUresult r = enc(&prm.tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dsrc, d2, s2, bx, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "ingest: map %d failed %d", i, static_cast<int>(r));
  }
  {
    cuuint64_t d3[3] = {64, 128, static_cast<cuuint64_t>(src_tiles)};
    cuuint64_t s3[2] = {128, 16384};
    cuuint32_t b3[3] = {64, 128, 2};
    CUresult r = enc(&prm.tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dsrc, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DSX_CHECK(r == CUDA_SUCCESS, DSX_E_CUDA, "ingest: 3D map failed %d", static_cast<int>(r));
  }
  prm.src = dsrc;
  prm.cycles = dcyc;
  prm.src_tiles = src_tiles;
  prm.nloads = 768;
  const int smem = 1024 + 12 * 16384 + 128;
  DSX_CUDA(cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Case { int kind, rows, nslot, nthr, grid; };
  const Case cases[] = {
      {0, 128, 12, 1, 1}, {0, 128, 12, 1, 128}, {0, 128, 6, 1, 1}, {0, 128, 3, 1, 1}, {0, 128, 1, 1, 1},
      {0, 64, 12, 1, 1},  {0, 64, 6, 1, 1},     {0, 64, 3, 1, 1},  {0, 64, 1, 1, 1},
      {0, 32, 12, 1, 1},  {0, 32, 1, 1, 1},     {0, 16, 12, 1, 1}, {0, 16, 1, 1, 1},
      {0, 128, 12, 2, 1}, {0, 128, 12, 4, 1},   {0, 128, 12, 4, 128}, {0, 32, 12, 4, 1},
      {1, 128, 12, 1, 1}, {1, 128, 12, 4, 1},   {1, 128, 1, 1, 1},
      {2, 128, 6, 1, 1},  {2, 128, 6, 2, 1},    {2, 128, 6, 2, 128}, {2, 128, 1, 1, 1},
  };
  for (const Case& c : cases) {
    prm.kind = c.kind; prm.rows = c.rows; prm.nslot = c.nslot; prm.nthr = c.nthr;
    long long cyc[1024];
    for (int rep = 0; rep < 2; ++rep) {                       // first repetition warms L2
      k_ingest<<<c.grid, 128, smem>>>(prm);
      DSX_CUDA(cudaDeviceSynchronize());
    }
    DSX_CUDA(cudaMemcpy(cyc, dcyc, c.grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < c.grid; ++i)
      for (int w = 0; w < c.nthr; ++w) mx = std::max(mx, cyc[i * 4 + w]);
    const double per = (c.kind == 2) ? 32768.0 : c.rows * 128.0;
    const int n = prm.nloads / c.nthr * c.nthr;
    char line[240];
    snprintf(line, sizeof(line), "ingest kind %d (%s) %5.0f B/load, %2d slots, %d issuing threads, grid %3d: %6.0f cycles/load, %.1f B/cycle/SM\n",
             c.kind, c.kind == 0 ? "tensor2d" : (c.kind == 1 ? "bulk1d" : "tensor3d x2"), per, c.nslot, c.nthr, c.grid,
             static_cast<double>(mx) / (n / c.nthr) , per * n / static_cast<double>(mx));
    report += line;
  }
  cudaFree(dsrc);
  cudaFree(dcyc);
  return DSX_OK;
}
#endif  // DSX_EXPERIMENTS
}  // namespace dsx

extern "C" int dsx_selftest(int device, int which, char* report, int report_bytes) {
  using namespace dsx;
  DSX_CUDA(cudaSetDevice(device));
  std::string rep;
  int rc = DSX_OK;
  std::string failed;
  if (which < 0 || which == 0) {
    int r = run_selftest<1>(rep);
    if (r != DSX_OK) { rc = DSX_E_KERNEL; failed += " umma_cta_group1"; }
  }
  if (which < 0 || which == 1) {
    int r = run_selftest<2>(rep);
    if (r != DSX_OK) { rc = DSX_E_KERNEL; failed += " umma_cta_group2"; }
  }
  if (which < 0 || which == 2) {   // row-shifted SWIZZLE_128B operand descriptors: the dilated taps of both layer kernels rely on them
    int r = run_shift_experiment(rep, which == 2);
    if (r != DSX_OK) { rc = DSX_E_KERNEL; failed += " shifted_descriptors"; }
  }
  if (which == 3) {   // informational experiment: per-SM TMA ingest rate (only in -DDSX_EXPERIMENTS builds)
#ifdef DSX_EXPERIMENTS
    int r = run_ingest_experiment(rep);
    if (r != DSX_OK) rc = r;
#else
    rep += "ingest micro-benchmark not compiled in (build with -DDSX_EXPERIMENTS)\n";
#endif
  }
  if (report && report_bytes > 0) {
    strncpy(report, rep.c_str(), static_cast<size_t>(report_bytes) - 1);
    report[report_bytes - 1] = 0;
  }
  if (rc != DSX_OK) set_error("selftest failed:%s | %s", failed.c_str(), rep.c_str());
  return rc;
}
