// RCAB tail in ONE kernel: conv2 (3x3, 64->64, tcgen05) + channel attention + RCAB skip.
//
// Replaces, for one RCAB of the reference: body[2] = default_conv (common.py:6-9), body[3] = CALayer
// (architectures.py:41-44) and `res += x` (:83).  The channel attention needs the mean of conv2's output over
// the WHOLE image before any pixel can be rescaled -- a grid-wide dependency.  Instead of writing the conv output
// `u` to HBM, launching a second kernel and reading it back, every CTA keeps its accumulators in TENSOR MEMORY
// (8 tiles x 64 fp32 columns = all 512 TMEM columns) across a grid barrier:
//   phase 1  per tile: tcgen05.ld -> +bias -> per-channel sums over the tile -> pool partials (global)
//            [training only: u is also TMA-stored, backward needs it]
//   barrier  all CTAs (one persistent CTA per SM, grid <= #SMs) arrive on a monotonic 64-bit counter
//   phase 2  per tile: y = sigmoid(W2 relu(W1 mean + b1) + b2) for the tile's image (recomputed per CTA from
//            the partials: 64->Cr->64, trivial), tcgen05.ld the SAME accumulator again, out = x + u*y with x
//            TMA-prefetched into shared memory, fp32 residual stream + bf16 operand copy TMA-stored.
// So `u` never leaves the SM in inference, and one launch (plus one barrier) replaces two launches.
// Tiles are assigned in contiguous runs per CTA so a CTA's tiles belong to one image (two at a boundary).
#pragma once
#include "conv3x3_tc.cuh"

namespace rb {

struct CaFusedArgs {
  const float *w1, *b1, *w2, *b2;          // FC weights [Cr][64], [Cr], [64][Cr], [64]
  float *save_mean, *save_hid, *save_y;    // training: CA vectors for backward ([N][64], [N][Cr], [N][64]) or null
  unsigned long long* grid_bar;            // monotonic arrival counter of THIS op (zeroed once at plan build)
  int cr, hw, partials_per_img, tiles_per_cta, store_u;
};

constexpr int kCaMaxTiles = 8;             // 8 x 64 fp32 columns = 512 TMEM columns
constexpr int kCaStgBytes = 2 * kStgF32Bytes;   // phase 1: two fp32 slots; phase 2 slot A reuses the first 48 KB
constexpr int kCaSlotBytes = kStgF32Bytes + kStgBf16Bytes;

__host__ inline size_t conv_ca_smem_bytes(int stages) {
  return 1024 + kCaStgBytes + size_t(9) * conv_b_block_bytes(64) + size_t(stages) * kAStageBytes;
}

#ifdef RB_CONV_CA_KERNEL_IMPL   // the kernel is compiled in api.cu only; other units use the structs above
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kConvCaThreads = 192;
__global__ void __launch_bounds__(kConvCaThreads, 1)
conv3x3_ca_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args, const CaFusedArgs ca) {
  constexpr int BN = 64;
  constexpr int kBBlock = BN * 128;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[kCaMaxTiles];
  __shared__ __align__(8) uint64_t b_bar;
  __shared__ __align__(8) uint64_t in_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[BN], red_s[128], mean_s[64], hid_s[16], y_s[64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  grid_dep_launch_dependents();
#define CA_STAMP(slot) do { if (args.dbg) args.dbg[(148 + blockIdx.x) * 16 + (slot)] = clock64(); } while (0)
  if (threadIdx.x == 0) CA_STAMP(0);
  const int stages = args.stages;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_res = smem + kCaStgBytes;
  uint8_t* stage0 = b_res + 9 * kBBlock;

  const int mt_begin = blockIdx.x * ca.tiles_per_cta;
  const int mt_end = min(mt_begin + ca.tiles_per_cta, args.m_tiles);
  const int tiles_per_img = args.tiles_x * args.tiles_y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < kCaMaxTiles; ++i) mbar_init(&tmem_full_bar[i], 1);
    mbar_init(&b_bar, 1);
    mbar_init(&in_bar[0], 1);
    mbar_init(&in_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&maps.a[0]); tma_prefetch_desc(&maps.w); tma_prefetch_desc(&maps.rf); }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(&b_bar, uint32_t(9) * kBBlock);
      tma_load_3d(b_res, &maps.w, &b_bar, 0, 0, 0);
    }
    __syncwarp();
    grid_dep_wait();
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt) {
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int y0 = (rem / args.tiles_x) * kTileH, x0 = (rem % args.tiles_x) * kTileW;
      for (int kx = 0; kx < 3; ++kx) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], kAStageBytes);
          tma_load_4d(stage0 + stage * kAStageBytes, &maps.a[0], &full_bar[stage], 0, x0 + kx - 1, y0 - 1, n);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer: tile `it` -> TMEM slot `it`
    mbar_wait(&b_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int mt = mt_begin; mt < mt_end; ++mt, ++it) {
      const uint32_t d_tmem = tmem_base + uint32_t(it * BN);
      for (int kx = 0; kx < 3; ++kx) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(stage0 + stage * kAStageBytes);
          const uint32_t b_addr = smem_u32(b_res + (kx * 3) * kBBlock);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t adesc = make_smem_desc(a_addr + ky * (kTileW * 128), 16, 1024, kLayoutSw128);
            const uint64_t bdesc = make_smem_desc(b_addr + ky * kBBlock, 16, 1024, kLayoutSw128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[it]);
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue (128 threads)
    const int q = warp & 3, row = q * 32 + lane, et = (warp - 2) * 32 + lane;
    const int ly = row >> 4, lx = row & 15;
    const uint32_t swz = uint32_t(row & 7);
    if (et < BN) bias_s[et] = args.bias ? args.bias[et] : 0.f;
    named_bar_sync(1, 128);
    grid_dep_wait();
    const int my_tiles = mt_end - mt_begin;

    // ------------------------------------------------------------ phase 1: pool partials (+ u store in training)
    for (int it = 0; it < my_tiles; ++it) {
      const int mt = mt_begin + it;
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int y0 = (rem / args.tiles_x) * kTileH, x0 = (rem % args.tiles_x) * kTileW;
      const bool valid = (y0 + ly < args.H) && (x0 + lx < args.W);
      uint8_t* stg = smem + (it & 1) * kStgF32Bytes;
      uint8_t* mine = stg + row * 128;
      mbar_wait(&tmem_full_bar[it], 0);
      tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(it * BN);
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 o;
          o.x = valid ? __uint_as_float(v[h * 32 + c * 4 + 0]) + bias_s[h * 32 + c * 4 + 0] : 0.f;
          o.y = valid ? __uint_as_float(v[h * 32 + c * 4 + 1]) + bias_s[h * 32 + c * 4 + 1] : 0.f;
          o.z = valid ? __uint_as_float(v[h * 32 + c * 4 + 2]) + bias_s[h * 32 + c * 4 + 2] : 0.f;
          o.w = valid ? __uint_as_float(v[h * 32 + c * 4 + 3]) + bias_s[h * 32 + c * 4 + 3] : 0.f;
          *reinterpret_cast<float4*>(mine + h * kABytes + ((uint32_t(c) ^ swz) << 4)) = o;
        }
      if (ca.store_u) {
        fence_proxy_async_smem();
        if (et == 0) tma_store_wait_read0();   // the other slot's store (one tile ago) has been read out
      }
      named_bar_sync(2, 128);
      if (ca.store_u && et == 0) {
        tma_store_4d(&maps.mb, stg, 0, x0, y0, n);               // maps.mb carries the fp32 map of `u` here
        tma_store_4d(&maps.mb, stg + kABytes, 32, x0, y0, n);
        tma_store_commit();
      }
      {
        const int c = et & 63, half = et >> 6;
        const uint8_t* base = stg + (c >> 5) * kABytes + (c & 3) * 4;
        const uint32_t ch = uint32_t((c & 31) >> 2);
        float s = 0.f;
#pragma unroll 8
        for (int r = half * 64; r < half * 64 + 64; ++r)
          s += *reinterpret_cast<const float*>(base + r * 128 + ((ch ^ uint32_t(r & 7)) << 4));
        args.pool_partial[(size_t(mt) * 2 + half) * BN + c] = s;
      }
    }
    // ------------------------------------------------------------ grid barrier: every tile's partials are out
    if (et == 0) CA_STAMP(1);
    __threadfence();
    if (et == 0) tma_store_wait_read0();       // phase 2 reuses the staging area as an input slot
    named_bar_sync(2, 128);
    // FC parameters -> the idle tail of the staging area ([48 KB, 64 KB)), fetched while the barrier is pending
    const int cr = ca.cr;
    float* w1_s = reinterpret_cast<float*>(smem + kCaSlotBytes);   // [cr][64]
    float* w2_s = w1_s + 16 * 64;                                  // [64][cr]
    float* b1_s = w2_s + 64 * 16;                                  // [cr]
    float* b2_s = b1_s + 16;                                       // [64]
    for (int i = et; i < cr * 64; i += 128) { w1_s[i] = ca.w1[i]; w2_s[i] = ca.w2[i]; }
    if (et < cr) b1_s[et] = ca.b1[et];
    if (et < 64) b2_s[et] = ca.b2[et];
    if (et == 0) {
      const unsigned long long prev = atomicAdd(ca.grid_bar, 1ull);
      const unsigned long long target = (prev / gridDim.x + 1ull) * gridDim.x;
      const long long t0 = clock64();
      while (ld_acquire_u64(ca.grid_bar) < target) {
        __nanosleep(64);
        if (clock64() - t0 > RB_WATCHDOG_CYCLES) { printf("rumpy_b200: CA grid barrier watchdog (block %d)\n", (int)blockIdx.x); __trap(); }
      }
      __threadfence();
      CA_STAMP(2);
    }
    named_bar_sync(2, 128);

    // ------------------------------------------------------------ phase 2: y, then out = x + u*y from TMEM
    uint8_t* slot_base[2] = {smem, stage0};    // slot B = the A-stage area, idle once every MMA has retired
    auto issue_inputs = [&](int it) {
      const int mt = mt_begin + it;
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int y0 = (rem / args.tiles_x) * kTileH, x0 = (rem % args.tiles_x) * kTileW;
      uint8_t* sf = slot_base[it & 1];
      mbar_expect_tx(&in_bar[it & 1], kStgF32Bytes);
      tma_load_4d(sf, &maps.rf, &in_bar[it & 1], 0, x0, y0, n);
      tma_load_4d(sf + kABytes, &maps.rf, &in_bar[it & 1], 32, x0, y0, n);
    };
    if (et == 0 && my_tiles > 0) issue_inputs(0);
    int cached_n = -1;
    for (int it = 0; it < my_tiles; ++it) {
      const int mt = mt_begin + it;
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int y0 = (rem / args.tiles_x) * kTileH, x0 = (rem % args.tiles_x) * kTileW;
      if (n != cached_n) {
        cached_n = n;
        {  // mean over the image from the per-tile partials: 2 thread groups x 64 channels, 4 loads in flight
          const int c = et & 63, g = et >> 6;
          const float* pp = args.pool_partial + size_t(n) * ca.partials_per_img * BN + c;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f;
          int i = g;
          for (; i + 14 < ca.partials_per_img; i += 16) {     // 8 independent L2 loads in flight per thread
            s0 += pp[size_t(i) * BN]; s1 += pp[size_t(i + 2) * BN]; s2 += pp[size_t(i + 4) * BN];
            s3 += pp[size_t(i + 6) * BN]; s4 += pp[size_t(i + 8) * BN]; s5 += pp[size_t(i + 10) * BN];
            s6 += pp[size_t(i + 12) * BN]; s7 += pp[size_t(i + 14) * BN];
          }
          for (; i < ca.partials_per_img; i += 2) s0 += pp[size_t(i) * BN];
          s0 += s4; s1 += s5; s2 += s6; s3 += s7;
          red_s[et] = (s0 + s1) + (s2 + s3);
        }
        named_bar_sync(2, 128);
        if (et < 64) mean_s[et] = (red_s[et] + red_s[64 + et]) / float(ca.hw);
        named_bar_sync(2, 128);
        for (int j = warp - 2; j < cr; j += 4) {   // hidden unit j: warp-shuffle dot over the 64 channels
          float s = w1_s[j * BN + lane] * mean_s[lane] + w1_s[j * BN + 32 + lane] * mean_s[32 + lane];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) hid_s[j] = fmaxf(s + b1_s[j], 0.f);
        }
        named_bar_sync(2, 128);
        if (et < 64) {
          float s = b2_s[et];
          for (int j = 0; j < cr; ++j) s = fmaf(w2_s[et * cr + j], hid_s[j], s);
          y_s[et] = 1.f / (1.f + __expf(-s));
        }
        named_bar_sync(2, 128);
        if (et == 0 && it == 0) CA_STAMP(3);
        if (ca.save_y != nullptr && rem == 0) {    // the CTA owning image n's first tile records the CA vectors
          if (et < 64) { ca.save_y[n * BN + et] = y_s[et]; ca.save_mean[n * BN + et] = mean_s[et]; }
          if (et < cr) ca.save_hid[n * cr + et] = hid_s[et];
        }
      }
      uint8_t* sf = slot_base[it & 1];
      uint8_t* sb = sf + kStgF32Bytes;
      uint8_t* my_f32 = sf + row * 128;
      uint8_t* my_bf16 = sb + row * 128;
      const bool valid = (y0 + ly < args.H) && (x0 + lx < args.W);
      uint32_t v[64];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(it * BN);
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      mbar_wait(&in_bar[it & 1], uint32_t(it >> 1) & 1u);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 r = *reinterpret_cast<const float4*>(my_f32 + h * kABytes + ((uint32_t(c) ^ swz) << 4));
          const int cb = h * 32 + c * 4;
          f[c * 4 + 0] = fmaf(__uint_as_float(v[cb + 0]) + bias_s[cb + 0], y_s[cb + 0], r.x);
          f[c * 4 + 1] = fmaf(__uint_as_float(v[cb + 1]) + bias_s[cb + 1], y_s[cb + 1], r.y);
          f[c * 4 + 2] = fmaf(__uint_as_float(v[cb + 2]) + bias_s[cb + 2], y_s[cb + 2], r.z);
          f[c * 4 + 3] = fmaf(__uint_as_float(v[cb + 3]) + bias_s[cb + 3], y_s[cb + 3], r.w);
        }
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(my_f32 + h * kABytes + ((uint32_t(c) ^ swz) << 4)) =
              make_float4(f[c * 4], f[c * 4 + 1], f[c * 4 + 2], f[c * 4 + 3]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(my_bf16 + ((uint32_t(h * 4 + c) ^ swz) << 4)) =
              make_uint4(pack_bf16x2(f[c * 8], f[c * 8 + 1]), pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]),
                         pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]));
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (et == 0) {
        CA_STAMP(4 + it);
        tma_store_4d(&maps.of, sf, 0, x0, y0, n);
        tma_store_4d(&maps.of, sf + kABytes, 32, x0, y0, n);
        tma_store_4d(&maps.ob[0], sb, 0, x0, y0, n);
        tma_store_commit();
        if (it + 1 < my_tiles) {
          tma_store_wait_read1();      // the other slot's stores (tile it-1) have been read out
          issue_inputs(it + 1);
        }
      }
    }
    if (et == 0) { tma_store_wait_all0(); CA_STAMP(12); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

#endif  // RB_CONV_CA_KERNEL_IMPL

}  // namespace rb
