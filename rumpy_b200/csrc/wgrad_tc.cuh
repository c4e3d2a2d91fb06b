// Weight gradient of the 3x3 convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces autograd's convolution_backward (weight part) behind `loss.backward()` in the reference
// (/root/reference/rumpy/shared_framework/models/base_architecture.py:432) for every nn.Conv2d created by
// common.default_conv (common.py:6-9):
//     dW[co][ci][ky][kx] = sum_{n,y,x} g[n,y,x,co] * X[n, y+ky-1, x+kx-1, ci]
//
// GEMM view (per 64x64 channel block): D[(tap,co), ci] += A[(tap,co), q] * B[ci, q], K = pixels q.
// Both operands are "MN-major" for UMMA (channels contiguous, pixels strided) -- exactly the NHWC layout,
// so the TMA boxes used by the forward kernel are the operands, untransposed:
//   B  = X tile   box {64 ch, 16 px, 8 rows}   (K rows = the tile's 128 pixels)
//   A  = g window box {64 ch, 16 px, 10 rows} per kx: the SAME three halo boxes as the forward conv, shifted by
//        (1-kx) in x; the ky shift is a 2 KB (16-pixel-row) offset inside the box, so all nine taps come from
//        three boxes.  Two taps are stacked in one M=128 instruction through the descriptor's leading-dimension
//        offset (LBO = address distance between the two taps' windows).
// Out-of-image pixels are zero-filled by TMA on both operands, which is precisely the conv's zero padding.
// One CTA = one job = (layer, 64x64 channel block, K split): it streams its pixel tiles through a 2-stage TMA
// pipeline, accumulates five [128 x 64] fp32 accumulators in TMEM (320 columns) and writes a [9][64][64] fp32
// partial; wgrad_reduce_kernel sums the K splits in a fixed order (deterministic) into the OIHW gradient.
#pragma once
#include "conv3x3_tc.cuh"

namespace rb {

struct alignas(64) WgradJob {
  CUtensorMap g;     // gradient operand, bf16 NHWC (box 64 x 16 x 10)
  CUtensorMap x;     // activation operand, bf16 NHWC (box 64 x 16 x 8)
  float* out;        // partial [9][64][64] fp32, tap index t = kx*3 + (2-ky)
  int gc0, xc0;      // channel offsets (multiples of 64) into the g / x tensors
  int tile_begin, tile_end;
  int tiles_x, tiles_y;
  int pad_[8];
};
static_assert(sizeof(WgradJob) % 64 == 0, "WgradJob must keep CUtensorMap alignment in arrays");

constexpr int kWgStageBytes = kABytes + 3 * kAStageBytes;  // X tile + three g halo boxes = 76 KB
constexpr int kWgStages = 2;
constexpr int kWgSmemBytes = 1024 + kWgStages * kWgStageBytes + 4096;  // + pad for the dummy 10th tap window
constexpr int kWgThreads = 192;

#ifdef RB_WGRAD_KERNELS_IMPL   // kernels are compiled in wgrad.cu only; other units use the job structs
__device__ __forceinline__ uint32_t wg_tap_offset(int t) {  // byte offset of tap t's window inside the g boxes
  return uint32_t(t / 3) * kAStageBytes + uint32_t(t % 3) * (kTileW * 128);
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgradJob* __restrict__ jobs) {
  constexpr uint32_t kIdesc = make_idesc_bf16(128, 64, /*a MN-major*/ 1, /*b MN-major*/ 1);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;

  const WgradJob* job = jobs + blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&job->g);
    tma_prefetch_desc(&job->x);
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int tile_begin = job->tile_begin, tile_end = job->tile_end;
  const int tiles_x = job->tiles_x, tiles_per_img = job->tiles_x * job->tiles_y;

  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const int gc0 = job->gc0, xc0 = job->xc0;
    for (int mt = tile_begin; mt < tile_end; ++mt) {
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int y0 = (rem / tiles_x) * kTileH;
      const int x0 = (rem % tiles_x) * kTileW;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* s = smem + stage * kWgStageBytes;
        mbar_expect_tx(&full_bar[stage], kWgStageBytes);
        tma_load_4d(s, &job->x, &full_bar[stage], xc0, x0, y0, n);
#pragma unroll
        for (int b = 0; b < 3; ++b)  // box b holds g shifted by (1 - kx) = (1 - b) in x, rows y0-1 .. y0+8
          tma_load_4d(s + kABytes + b * kAStageBytes, &job->g, &full_bar[stage], gc0, x0 + 1 - b, y0 - 1, n);
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = tile_begin; mt < tile_end; ++mt) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xb = smem_u32(smem + stage * kWgStageBytes);
        const uint32_t gb = xb + kABytes;
#pragma unroll 1
        for (int j = 0; j < kTileH; ++j) {  // K step = one 16-pixel tile row = 2 KB = two 8-row swizzle atoms
          const uint64_t bdesc = make_smem_desc(xb + j * 2048, 1024, 1024, kLayoutSw128);
#pragma unroll
          for (int p = 0; p < 5; ++p) {     // taps (2p, 2p+1) stacked in M = 128; tap 9 is a dummy window
            const uint32_t o0 = wg_tap_offset(2 * p);
            const uint32_t o1 = (p < 4) ? wg_tap_offset(2 * p + 1) : o0 + 2048;
            const uint64_t adesc = make_smem_desc(gb + o0 + j * 2048, o1 - o0, 1024, kLayoutSw128);
            umma_bf16(tmem_base + uint32_t(p * 64), adesc, bdesc, kIdesc, (mt != tile_begin) || (j != 0));
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  } else {
    // ===================================================================== epilogue: TMEM -> global partial
    const int q = warp & 3;
    const int row = q * 32 + lane;          // TMEM lane = (tap within the pair, co)
    const int half = row >> 6, co = row & 63;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    float* out = job->out;
#pragma unroll 1
    for (int p = 0; p < 5; ++p) {
      uint32_t v[64];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(p * 64);
      tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      const int t = 2 * p + half;
      if (t < 9) {
        float4* dst = reinterpret_cast<float4*>(out + (size_t(t) * 64 + co) * 64);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          dst[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                               __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

#endif  // RB_WGRAD_KERNELS_IMPL

// Sums the K-split partials of one 64x64 channel block (fixed order) and scatters into the OIHW gradient:
//   dW[o(co0+co)][ci0+ci][ky][kx] (=|+=) alpha * sum_s partial[s][t][co][ci],  t = kx*3 + (2-ky)
// o() undoes the pixel-shuffle row permutation of the packed weights (row = q*cpp + c  <->  o = c*r^2 + q).
struct WgradReduceJob {
  const float* partial;  // [splits][9][64][64]
  float* dw;             // OIHW fp32 gradient tensor base
  int splits, cout, cin, co0, ci0, r, accumulate;
  float alpha;
};

#ifdef RB_WGRAD_KERNELS_IMPL
__global__ void wgrad_reduce_kernel(const WgradReduceJob* __restrict__ jobs) {
  const WgradReduceJob jb = jobs[blockIdx.y];
  const int rr = jb.r * jb.r, cpp = jb.cout / rr;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 9 * 64 * 64; idx += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < jb.splits; ++k) s += jb.partial[size_t(k) * 9 * 64 * 64 + idx];
    const int ci = idx & 63, co = (idx >> 6) & 63, t = idx >> 12;
    const int kx = t / 3, ky = 2 - (t % 3);
    const int row = jb.co0 + co;
    if (row >= jb.cout) continue;   // thin tail conv: the operand is zero-padded to 64 output channels
    const int o = (jb.r > 1) ? (row % cpp) * rr + row / cpp : row;
    float* d = jb.dw + ((size_t(o) * jb.cin + jb.ci0 + ci) * 3 + ky) * 3 + kx;
    const float v = s * jb.alpha;
    *d = jb.accumulate ? *d + v : v;
  }
}
#endif  // RB_WGRAD_KERNELS_IMPL

}  // namespace rb
