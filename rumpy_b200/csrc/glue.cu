// Eval glue and training-patch pipeline on the device: the host-side numpy / PIL work that
// surrounds the trunk in the reference, restated as bandwidth-bound kernels so that a step moves uint8 images and a
// few scalars across PCIe instead of fp32 batches.
//
//   rumpy_psnr_y       clip -> Y of jpg-style YCbCr -> per-image PSNR   (reference sr_tools/metrics.py:33-44,
//                      image_tools/image_manipulation/image_functions.py:72-88, base_interface.py:208-222)
//   rumpy_quantize_u8  clip(x*255, 0, 255).astype(uint8) (truncation), NCHW fp32 -> NHWC uint8 (what gets saved:
//                      sr_tools/visualization.py:31-61)
//   rumpy_bicubic_upsample  the "LR" (bicubic) baseline row of the evaluation: ToPILImage -> Image.resize(BICUBIC)
//                      -> ToTensor, bit-exact with Pillow's 8-bit resampler (evaluation/standard_eval.py:240-275)
//   rumpy_patch_batch  random crop + hflip / vflip / transpose + ToTensor for a whole batch of LR/HR pairs from
//                      uint8 images resident in HBM (image_functions.py:287-362, sr_tools/data_handler.py:570-645)
//
// All four are integer / byte work bound by HBM: one pass, coalesced along the fastest output dimension, grids sized
// from the element count; reductions run in a fixed order (deterministic).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/rumpy_b200.h"
#include "host_util.cuh"

namespace rb {

constexpr int kPsnrMaxBlocks = 2048;   // partial sums per image (workspace: N * kPsnrMaxBlocks doubles)

__device__ __forceinline__ float luma_clipped(float r, float g, float b) {
  r = fminf(fmaxf(r, 0.f), 1.f);
  g = fminf(fmaxf(g, 0.f), 1.f);
  b = fminf(fmaxf(b, 0.f), 1.f);
  return 0.299f * r + 0.587f * g + 0.114f * b;
}

// grid (blocks, N), block 256: partial[n][blk] = sum over a slice of (Y(sr) - Y(hr))^2 in double.
// VEC: HW % 4 == 0 -> 16-byte loads (six independent float4 loads in flight per thread).
template <bool VEC>
__global__ void psnr_y_partial_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                      double* __restrict__ partial, int HW) {
  const int n = blockIdx.y;
  const float* s = sr + size_t(n) * 3 * HW;
  const float* h = hr + size_t(n) * 3 * HW;
  double acc = 0.0;
  if (VEC) {
    const int Q = HW / 4;
    const float4* s4 = reinterpret_cast<const float4*>(s);
    const float4* h4 = reinterpret_cast<const float4*>(h);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Q; i += gridDim.x * blockDim.x) {
      const float4 sr_ = s4[i], sg = s4[Q + i], sb = s4[2 * Q + i];
      const float4 hr_ = h4[i], hg = h4[Q + i], hb = h4[2 * Q + i];
      const float d0 = luma_clipped(sr_.x, sg.x, sb.x) - luma_clipped(hr_.x, hg.x, hb.x);
      const float d1 = luma_clipped(sr_.y, sg.y, sb.y) - luma_clipped(hr_.y, hg.y, hb.y);
      const float d2 = luma_clipped(sr_.z, sg.z, sb.z) - luma_clipped(hr_.z, hg.z, hb.z);
      const float d3 = luma_clipped(sr_.w, sg.w, sb.w) - luma_clipped(hr_.w, hg.w, hb.w);
      acc += (double(d0) * double(d0) + double(d1) * double(d1)) + (double(d2) * double(d2) + double(d3) * double(d3));
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      const float d = luma_clipped(s[i], s[HW + i], s[2 * HW + i]) - luma_clipped(h[i], h[HW + i], h[2 * HW + i]);
      acc += double(d) * double(d);
    }
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[size_t(n) * gridDim.x + blockIdx.x] = red[0];
}

// one CTA per image: fixed-order tree sum of the partials -> PSNR (100 for identical images, metrics.py:41-42)
__global__ void psnr_y_finalize_kernel(const double* __restrict__ partial, float* __restrict__ psnr, int blocks,
                                       double inv_count, float max_value) {
  const int n = blockIdx.x;
  __shared__ double red[256];
  double s = 0.0;
  for (int b = threadIdx.x; b < blocks; b += 256) s += partial[size_t(n) * blocks + b];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mse = red[0] * inv_count;
    psnr[n] = mse == 0.0 ? 100.f : float(20.0 * log10(double(max_value) / sqrt(mse)));
  }
}

// out[n][y][x][c] = uint8(trunc(clip(in[n][c][y][x] * 255, 0, 255)))
__device__ __forceinline__ uint32_t quant1(float x) {
  const float v = __fmul_rn(x, 255.f);                         // numpy: float32 multiply, then clip, then truncate
  return __float2uint_rz(fminf(fmaxf(v, 0.f), 255.f));
}
// generic: one thread per pixel (C <= 4)
__global__ void quantize_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int C, int HW,
                                   long long total_px) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total_px;
       p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW;
    const int i = int(p - n * HW);
    const float* s = src + n * C * HW + i;
    uint8_t* d = dst + p * C;
    for (int c = 0; c < C; ++c) d[c] = uint8_t(quant1(s[size_t(c) * HW]));
  }
}
// C == 3, HW % 4 == 0: four pixels per thread, three 16-byte loads -> three 4-byte stores (12 packed bytes)
__global__ void quantize_u8_rgb4_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int HW,
                                        long long total_quads) {
  const int Q = HW / 4;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total_quads;
       p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / Q;
    const int i = int(p - n * Q);
    const float4* s = reinterpret_cast<const float4*>(src + n * 3 * HW);
    const float4 r = s[i], g = s[Q + i], b = s[2 * Q + i];
    const uint32_t b0 = quant1(r.x), b1 = quant1(g.x), b2 = quant1(b.x), b3 = quant1(r.y), b4 = quant1(g.y),
                   b5 = quant1(b.y), b6 = quant1(r.z), b7 = quant1(g.z), b8 = quant1(b.z), b9 = quant1(r.w),
                   b10 = quant1(g.w), b11 = quant1(b.w);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + p * 12);
    d[0] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    d[1] = b4 | (b5 << 8) | (b6 << 16) | (b7 << 24);
    d[2] = b8 | (b9 << 8) | (b10 << 16) | (b11 << 24);
  }
}

// geom[n] = {image index, y, x, flags (1: hflip, 2: vflip, 4: transpose), lr_h, lr_w}
// grid (ceil(side*side / 256), N, 2): z = 0 LR patch, z = 1 HR patch
__global__ void patch_batch_kernel(const uint8_t* const* __restrict__ lr_imgs, const uint8_t* const* __restrict__ hr_imgs,
                                   const int* __restrict__ geom, float* __restrict__ lr_out, float* __restrict__ hr_out,
                                   int crop, int scale) {
  const int n = blockIdx.y;
  const bool is_hr = blockIdx.z == 1;
  const int s = is_hr ? scale : 1;
  const int side = crop * s;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= side * side) return;
  const int* g = geom + n * 6;
  const int img = g[0], y0 = g[1] * s, x0 = g[2] * s, flags = g[3], w = g[5] * s;
  const int oy = i / side, ox = i - oy * side;
  int py = (flags & 4) ? ox : oy, px = (flags & 4) ? oy : ox;   // final = T(V(H(crop)))
  if (flags & 2) py = side - 1 - py;
  if (flags & 1) px = side - 1 - px;
  const uint8_t* src = (is_hr ? hr_imgs[img] : lr_imgs[img]) + (size_t(y0 + py) * w + (x0 + px)) * 3;
  float* dst = (is_hr ? hr_out : lr_out) + size_t(n) * 3 * side * side + i;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[size_t(c) * side * side] = __fdiv_rn(float(src[c]), 255.f);   // ToTensor()
}

// ---- bicubic baseline (standard_eval.py:240-275: ToPILImage -> Image.resize(BICUBIC) -> ToTensor) -----------------
// Pillow's 8-bit resampler (src/libImaging/Resample.c, restated in oracle/pil_resample.py): horizontal pass, uint8
// intermediate, vertical pass; per output index five taps at most, double-precision Keys weights (a = -0.5) clipped
// to the image, re-normalised and rounded to 22-bit fixed point.
//   bicubic_coeff_kernel  one thread per output column / group of four output rows: the tap tables ({k[5], first
//                         tap, taps} per column; dense 8 x 4 tap matrix per row group) evaluated with explicitly
//                         rounded double arithmetic (no FMA contraction: the same bits as the C code)
//   bicubic_up_kernel     one CTA = one 64 x 128 output tile of one plane: its table entries and the LR footprint
//                         (<= 38 x 70 pixels for scale >= 2, quantised like to_pil_image) go to shared memory, both
//                         passes run there (intermediate kept as one int per pixel so the vertical pass reads four
//                         columns with one 16-byte load and no byte extraction; four consecutive output rows
//                         share their <= 8 source rows through a dense tap matrix), v/255 is a correctly rounded
//                         multiply + Newton step, and the tile is written once with 16-byte stores.
// Algorithmic bytes: 4 B read per LR element + 4 B written per output element.  Versions (1080p x4 frame): taps
// recomputed per tile, runtime divisions in the index math, __fdiv_rn per output: 0.32 ms, issue-bound; tap table +
// per-row vertical pass + quotient table in shared memory: 0.187 ms, bound by shared-memory wavefronts (95 %); grouped
// vertical pass + division-free quotient: 0.149 ms, issue-bound (84 %); row-group table, 32-row tiles: 0.147 ms with
// 28 % of the instructions in per-CTA set-up, hence 64-row tiles.
constexpr int kBicTH = 64, kBicTW = 128;
constexpr int kBicRows = kBicTH / 2 + 10, kBicCols = kBicTW / 2 + 8;   // footprint + 4 (zero-weight taps past the end)
constexpr int kBicPrec = 22;

struct __align__(16) BicTap {
  int k[5];
  int lo, n, pad;
};

__device__ __forceinline__ double pil_bicubic(double x) {
  x = fabs(x);
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output index xx of an upscaling pass (filterscale = 1)
__device__ __forceinline__ BicTap pil_coeffs(int xx, int in_size, int out_size) {
  const double scale = __ddiv_rn(double(in_size), double(out_size));
  const double center = __dmul_rn(__dadd_rn(double(xx), 0.5), scale);
  int lo = __double2int_rz(__dadd_rn(__dsub_rn(center, 2.0), 0.5));
  int hi = __double2int_rz(__dadd_rn(__dadd_rn(center, 2.0), 0.5));
  lo = lo < 0 ? 0 : lo;
  hi = hi > in_size ? in_size : hi;
  const int n = hi - lo;
  double w[5], ww = 0.0;
#pragma unroll
  for (int x = 0; x < 5; ++x) {
    w[x] = x < n ? pil_bicubic(__dadd_rn(__dsub_rn(double(x + lo), center), 0.5)) : 0.0;
    if (x < n) ww = __dadd_rn(ww, w[x]);
  }
  BicTap e;
#pragma unroll
  for (int x = 0; x < 5; ++x) {
    const double v = ww != 0.0 ? __ddiv_rn(w[x], ww) : w[x];
    const double f = __dmul_rn(v, double(1 << kBicPrec));
    e.k[x] = x < n ? __double2int_rz(v < 0.0 ? __dadd_rn(-0.5, f) : __dadd_rn(0.5, f)) : 0;
  }
  e.lo = lo;
  e.n = n;
  e.pad = 0;
  return e;
}

// four consecutive output rows share <= 8 source rows: their taps as a dense [source row - base][row in group] matrix
struct __align__(16) BicGroup {
  int k[8][4];
  int base, span, pad0, pad1;
};
__host__ __device__ inline size_t bic_groups_offset(int OW) { return size_t(OW) * sizeof(BicTap); }

// workspace = OW column taps (BicTap), then ceil(OH / 4) row groups (BicGroup); one thread per entry
__global__ void bicubic_coeff_kernel(void* __restrict__ workspace, int H, int W, int scale) {
  const int OW = W * scale, OH = H * scale, G = (OH + 3) / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < OW) {
    static_cast<BicTap*>(workspace)[i] = pil_coeffs(i, W, OW);
  } else if (i < OW + G) {
    const int g = i - OW;
    BicGroup* out = reinterpret_cast<BicGroup*>(static_cast<char*>(workspace) + bic_groups_offset(OW)) + g;
    for (int u = 0; u < 8; ++u)
      for (int j = 0; j < 4; ++j) out->k[u][j] = 0;
    int base = 0, last_lo = 0;
    for (int j = 0; j < 4 && 4 * g + j < OH; ++j) {
      const BicTap e = pil_coeffs(4 * g + j, H, OH);
      if (j == 0) base = e.lo;
      last_lo = e.lo;
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (e.lo - base + t < 8) out->k[e.lo - base + t][j] = e.k[t];      // always true for scale >= 2
    }
    out->base = base;
    out->span = min(8, min(last_lo + 5, H) - base);                         // rows past H only carry zero taps
    out->pad0 = out->pad1 = 0;
  }
}

__device__ __forceinline__ int bic_clip8(int acc) { return __vimin_s32_relu(acc >> kBicPrec, 255); }

// v / 255.f, correctly rounded, for v in 0..255 without a division: 1/255 as a two-float constant, q = fma(v, hi,
// v * lo).  Equal to __fdiv_rn for all 256 values (checked exhaustively in exact arithmetic, and the 1080p test
// compares every byte value with ToTensor's true division).
__device__ __forceinline__ float bic_div255(int v) {
  const float f = float(v);
  return __fmaf_rn(f, 0x1.010102p-8f, __fmul_rn(f, -0x1.fdfdfep-33f));
}

// grid (ceil(OW / 128), ceil(OH / 64), N * C), block 256
__global__ void __launch_bounds__(256, 6) bicubic_up_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                         const void* __restrict__ workspace, int H, int W, int scale,
                                                         int vec_store) {
  __shared__ BicTap s_cx[kBicTW];
  __shared__ BicGroup s_cg[kBicTH / 4];
  __shared__ __align__(16) uint8_t s_lr[kBicRows][kBicCols];
  __shared__ __align__(16) int s_tmp[kBicRows][kBicTW];
  const int OH = H * scale, OW = W * scale;
  const int ox0 = blockIdx.x * kBicTW, oy0 = blockIdx.y * kBicTH;
  const int tw = min(kBicTW, OW - ox0), th = min(kBicTH, OH - oy0);
  const int groups = (th + 3) >> 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < kBicTW) {
    if (tid < tw) s_cx[tid] = static_cast<const BicTap*>(workspace)[ox0 + tid];
  } else {
    const int4* g = reinterpret_cast<const int4*>(static_cast<const char*>(workspace) + bic_groups_offset(OW) +
                                                  size_t(blockIdx.y) * (kBicTH / 4) * sizeof(BicGroup));
    for (int i = tid - kBicTW; i < groups * int(sizeof(BicGroup) / 16); i += 256 - kBicTW)
      reinterpret_cast<int4*>(s_cg)[i] = g[i];
  }
  __syncthreads();
  // bounds are non-decreasing in the output index: the tile's footprint is [first tap of its first, end of its last]
  const int x_lo = s_cx[0].lo, cols = s_cx[tw - 1].lo + s_cx[tw - 1].n - x_lo;
  const int y_lo = s_cg[0].base, rows = s_cg[groups - 1].base + s_cg[groups - 1].span - y_lo;
  const float* sp = src + size_t(blockIdx.z) * H * W + size_t(y_lo) * W + x_lo;
  {                                                                          // LR footprint, quantised like
    const int c = tid & 63;                                                  // to_pil_image: pic.mul(255).byte()
    if (c < cols)
      for (int r = tid >> 6; r < rows; r += 4) s_lr[r][c] = uint8_t(quant1(sp[r * W + c]));
    if (cols > 64) {
      const int c2 = 64 + (tid & 7);
      if (c2 < cols)
        for (int r = tid >> 3; r < rows; r += 32) s_lr[r][c2] = uint8_t(quant1(sp[r * W + c2]));
    }
  }
  __syncthreads();
  {                                                                          // horizontal pass: one column per thread
    const int x = tid & (kBicTW - 1);
    if (x < tw) {
      const BicTap e = s_cx[x];
      const int lo = e.lo - x_lo;
      for (int r = tid >> 7; r < rows; r += 2) {
        const uint8_t* p = &s_lr[r][lo];               // taps past e.n carry k = 0 (the row buffer is padded by 4)
        const int acc = (1 << (kBicPrec - 1)) + int(p[0]) * e.k[0] + int(p[1]) * e.k[1] + int(p[2]) * e.k[2] +
                        int(p[3]) * e.k[3] + int(p[4]) * e.k[4];
        s_tmp[r][x] = bic_clip8(acc);
      }
    }
  }
  __syncthreads();
  // vertical pass: warp = four consecutive output rows, lane = four columns; the source rows those output rows share
  // are read once (one 16-byte load each) and hit the dense tap matrix of the group
  const int x4 = lane * 4;
  if (x4 >= tw) return;
  for (int g = warp; g < groups; g += 8) {
    const int base = s_cg[g].base - y_lo, span = s_cg[g].span;
    int acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 1 << (kBicPrec - 1);
    const int4* prow = reinterpret_cast<const int4*>(&s_tmp[base][x4]);
    const int4* krow = reinterpret_cast<const int4*>(&s_cg[g].k[0][0]);
#pragma unroll 2
    for (int u = 0; u < span; ++u) {
      const int4 p = prow[u * (kBicTW / 4)];
      const int4 k = krow[u];
      const int kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] += p.x * kk[j];
        acc[j][1] += p.y * kk[j];
        acc[j][2] += p.z * kk[j];
        acc[j][3] += p.w * kk[j];
      }
    }
    float* d = dst + size_t(blockIdx.z) * OH * OW + size_t(oy0 + 4 * g) * OW + ox0 + x4;
    const bool vec = vec_store && x4 + 3 < tw;
    const int nrow = min(4, th - 4 * g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < nrow) {
        const float4 o = make_float4(bic_div255(bic_clip8(acc[j][0])), bic_div255(bic_clip8(acc[j][1])),
                                     bic_div255(bic_clip8(acc[j][2])), bic_div255(bic_clip8(acc[j][3])));
        if (vec) {
          *reinterpret_cast<float4*>(d) = o;
        } else {
          d[0] = o.x;
          if (x4 + 1 < tw) d[1] = o.y;
          if (x4 + 2 < tw) d[2] = o.z;
          if (x4 + 3 < tw) d[3] = o.w;
        }
        d += OW;
      }
    }
  }
}


// ---- Lanczos baseline (standard_eval.py:252-253: `--lanczos_upsample`, Image.resize(..., LANCZOS)) ------------------
// Same two-pass 8-bit resampler as the bicubic baseline with Pillow's Lanczos-3 filter (support 3: up to seven taps per
// output index).  The tap tables come from the host (libm sin, see rumpy_lanczos_upsample); the kernel is a plain
// tiled FIR: one CTA = one 16 x 64 output tile of one plane, its LR footprint quantised like to_pil_image in shared
// memory, horizontal pass into a uint8 intermediate (Pillow's intermediate image is uint8), vertical pass, v / 255.
// An evaluation-time baseline, launched once per image: correctness against Pillow is the bar, not the roofline.
constexpr int kLanTaps = 7;
struct __align__(16) LanTap {
  int k[kLanTaps];
  int lo;
  int n, pad0, pad1, pad2;
};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc, Lanczos-3, upscaling (filterscale = 1); host, libm
inline LanTap lanczos_coeffs(int xx, int in_size, int out_size) {
  auto sinc = [](double v) { if (v == 0.0) return 1.0; v = v * M_PI; return sin(v) / v; };
  auto filt = [&](double v) { return (-3.0 <= v && v < 3.0) ? sinc(v) * sinc(v / 3) : 0.0; };
  const double scale = double(in_size) / double(out_size);
  const double center = (double(xx) + 0.5) * scale;
  int lo = int(center - 3.0 + 0.5), hi = int(center + 3.0 + 0.5);
  lo = lo < 0 ? 0 : lo;
  hi = hi > in_size ? in_size : hi;
  LanTap e{};
  e.lo = lo;
  e.n = hi - lo;
  double w[kLanTaps], ww = 0.0;
  for (int x = 0; x < e.n; ++x) { w[x] = filt(double(x + lo) - center + 0.5); ww += w[x]; }
  for (int x = 0; x < e.n; ++x) {
    const double v = ww != 0.0 ? w[x] / ww : w[x];
    e.k[x] = int(v < 0.0 ? -0.5 + v * double(1 << kBicPrec) : 0.5 + v * double(1 << kBicPrec));
  }
  return e;
}

constexpr int kLanTH = 16, kLanTW = 64;
constexpr int kLanRows = kLanTH / 2 + kLanTaps + 1, kLanCols = kLanTW / 2 + kLanTaps + 1;   // footprint for scale >= 2

// grid (ceil(OW / 64), ceil(OH / 16), N * C), block 256
__global__ void __launch_bounds__(256) lanczos_up_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                         const LanTap* __restrict__ taps, int H, int W, int scale) {
  __shared__ LanTap s_cx[kLanTW], s_cy[kLanTH];
  __shared__ uint8_t s_lr[kLanRows][kLanCols + 3];
  __shared__ uint8_t s_tmp[kLanRows][kLanTW];
  const int OH = H * scale, OW = W * scale;
  const int ox0 = blockIdx.x * kLanTW, oy0 = blockIdx.y * kLanTH;
  const int tw = min(kLanTW, OW - ox0), th = min(kLanTH, OH - oy0);
  const int tid = threadIdx.x;
  if (tid < tw) s_cx[tid] = taps[ox0 + tid];
  else if (tid >= kLanTW && tid - kLanTW < th) s_cy[tid - kLanTW] = taps[OW + oy0 + tid - kLanTW];
  __syncthreads();
  const int x_lo = s_cx[0].lo, cols = s_cx[tw - 1].lo + s_cx[tw - 1].n - x_lo;
  const int y_lo = s_cy[0].lo, rows = s_cy[th - 1].lo + s_cy[th - 1].n - y_lo;
  const float* sp = src + size_t(blockIdx.z) * H * W + size_t(y_lo) * W + x_lo;
  for (int i = tid; i < rows * cols; i += 256) {
    const int r = i / cols, c = i - r * cols;
    s_lr[r][c] = uint8_t(quant1(sp[r * W + c]));                  // to_pil_image: pic.mul(255).byte()
  }
  __syncthreads();
  for (int i = tid; i < rows * tw; i += 256) {                    // horizontal pass -> uint8 intermediate
    const int r = i / tw, x = i - r * tw;
    const LanTap& e = s_cx[x];
    int acc = 1 << (kBicPrec - 1);
    for (int t = 0; t < e.n; ++t) acc += int(s_lr[r][e.lo - x_lo + t]) * e.k[t];
    s_tmp[r][x] = uint8_t(bic_clip8(acc));
  }
  __syncthreads();
  float* dp = dst + size_t(blockIdx.z) * OH * OW;
  for (int i = tid; i < th * tw; i += 256) {                      // vertical pass, ToTensor's v / 255
    const int y = i / tw, x = i - y * tw;
    const LanTap& e = s_cy[y];
    int acc = 1 << (kBicPrec - 1);
    for (int t = 0; t < e.n; ++t) acc += int(s_tmp[e.lo - y_lo + t][x]) * e.k[t];
    dp[size_t(oy0 + y) * OW + ox0 + x] = bic_div255(bic_clip8(acc));
  }
}

}  // namespace rb

using namespace rb;

extern "C" {

long long rumpy_psnr_y_workspace(int N) { return N > 0 ? (long long)N * kPsnrMaxBlocks * (long long)sizeof(double) : -1LL; }

int rumpy_psnr_y(const float* sr, const float* hr, float* psnr, void* workspace, int N, int H, int W, float max_value,
                 void* stream) {
  if (!sr || !hr || !psnr || !workspace) return set_error(RUMPY_ERR_ARG, "psnr_y: null pointer");
  if (N < 1 || H < 1 || W < 1 || N > 65535) return set_error(RUMPY_ERR_ARG, "psnr_y: N=%d H=%d W=%d", N, H, W);
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  const int HW = H * W;
  const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(sr) | reinterpret_cast<uintptr_t>(hr)) % 16 == 0;
  // enough CTAs to fill the GPU (8 x 256 threads per SM) without handing a CTA less than ~2k pixels
  int blocks = (sms * 8 + N - 1) / N;
  const int by_work = (HW + 2047) / 2048;
  if (blocks > by_work) blocks = by_work;
  if (blocks > kPsnrMaxBlocks) blocks = kPsnrMaxBlocks;
  if (blocks < 1) blocks = 1;
  if (vec) psnr_y_partial_kernel<true><<<dim3(blocks, N), 256, 0, s>>>(sr, hr, part, HW);
  else psnr_y_partial_kernel<false><<<dim3(blocks, N), 256, 0, s>>>(sr, hr, part, HW);
  if (int e = check_launch("psnr_y_partial")) return e;
  psnr_y_finalize_kernel<<<N, 256, 0, s>>>(part, psnr, blocks, 1.0 / (double(H) * W), max_value);
  return check_launch("psnr_y_finalize");
}

int rumpy_quantize_u8(const float* src_nchw, unsigned char* dst_nhwc, int N, int C, int H, int W, void* stream) {
  if (!src_nchw || !dst_nhwc) return set_error(RUMPY_ERR_ARG, "quantize_u8: null pointer");
  if (N < 1 || C < 1 || C > 4 || H < 1 || W < 1) return set_error(RUMPY_ERR_ARG, "quantize_u8: N=%d C=%d H=%d W=%d", N, C, H, W);
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  const int HW = H * W;
  const long long px = (long long)N * HW;
  const bool rgb4 = C == 3 && HW % 4 == 0 && reinterpret_cast<uintptr_t>(src_nchw) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(dst_nhwc) % 4 == 0;
  const long long items = rgb4 ? px / 4 : px;
  long long blocks = (items + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  if (rgb4)
    quantize_u8_rgb4_kernel<<<int(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(src_nchw, dst_nhwc, HW, items);
  else
    quantize_u8_kernel<<<int(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(src_nchw, dst_nhwc, C, HW, px);
  return check_launch("quantize_u8");
}

int rumpy_patch_batch(const unsigned char* const* lr_imgs, const unsigned char* const* hr_imgs, const int* geom,
                      float* lr_out, float* hr_out, int N, int crop, int scale, void* stream) {
  if (!lr_imgs || !hr_imgs || !geom || !lr_out || !hr_out) return set_error(RUMPY_ERR_ARG, "patch_batch: null pointer");
  if (N < 1 || crop < 1 || scale < 1 || N > 65535) return set_error(RUMPY_ERR_ARG, "patch_batch: N=%d crop=%d scale=%d", N, crop, scale);
  if (int e = device_info(nullptr)) return e;
  const int side = crop * scale;
  patch_batch_kernel<<<dim3((side * side + 255) / 256, N, 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      lr_imgs, hr_imgs, geom, lr_out, hr_out, crop, scale);
  return check_launch("patch_batch");
}

long long rumpy_lanczos_workspace(int H, int W, int scale) {
  if (H < 1 || W < 1 || scale < 2 || scale > 8) return -1LL;
  // column taps + row taps (LanTap each) + the horizontally resampled uint8 intermediate of ONE plane batch
  return (long long)(W * scale + H * scale) * (long long)sizeof(LanTap);
}

int rumpy_lanczos_upsample(const float* lr_nchw, float* out_nchw, void* workspace, int N, int C, int H, int W, int scale,
                           void* stream) {
  if (!lr_nchw || !out_nchw || !workspace) return set_error(RUMPY_ERR_ARG, "lanczos_upsample: null pointer");
  if (N < 1 || C < 1 || H < 1 || W < 1 || scale < 2 || scale > 8 || (long long)N * C > 65535)
    return set_error(RUMPY_ERR_ARG, "lanczos_upsample: N=%d C=%d H=%d W=%d scale=%d", N, C, H, W, scale);
  if (int e = device_info(nullptr)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int OW = W * scale, OH = H * scale;
  // tap tables on the host: Pillow evaluates sin() of libm in double; the same libm here gives the same bits
  // (CUDA's double sin is 2-ulp accurate, not correctly rounded).  The vector lives until the copy has been staged:
  // cudaMemcpyAsync from pageable memory returns after the driver has copied the source.
  static thread_local std::vector<LanTap> taps;
  taps.resize(size_t(OW) + OH);
  for (int x = 0; x < OW; ++x) taps[x] = lanczos_coeffs(x, W, OW);
  for (int y = 0; y < OH; ++y) taps[size_t(OW) + y] = lanczos_coeffs(y, H, OH);
  if (cudaMemcpyAsync(workspace, taps.data(), taps.size() * sizeof(LanTap), cudaMemcpyHostToDevice, s) != cudaSuccess)
    return set_error(RUMPY_ERR_CUDA, "lanczos_upsample: table upload: %s", cudaGetErrorString(cudaGetLastError()));
  const dim3 grid((OW + 63) / 64, (OH + 15) / 16, N * C);
  lanczos_up_kernel<<<grid, 256, 0, s>>>(lr_nchw, out_nchw, static_cast<const LanTap*>(workspace), H, W, scale);
  return check_launch("lanczos_upsample");
}

long long rumpy_bicubic_workspace(int H, int W, int scale) {
  if (H < 1 || W < 1 || scale < 2 || scale > 8) return -1LL;
  return (long long)W * scale * (long long)sizeof(BicTap) + ((long long)H * scale + 3) / 4 * (long long)sizeof(BicGroup);
}

int rumpy_bicubic_upsample(const float* lr_nchw, float* out_nchw, void* workspace, int N, int C, int H, int W, int scale,
                           void* stream) {
  if (!lr_nchw || !out_nchw || !workspace) return set_error(RUMPY_ERR_ARG, "bicubic_upsample: null pointer");
  if (N < 1 || C < 1 || H < 1 || W < 1 || scale < 2 || scale > 8 || (long long)N * C > 65535 ||
      (long long)H * scale > (long long)kBicTH * 65535 || (long long)H * scale * W * scale > 0x7fffffffLL ||
      reinterpret_cast<uintptr_t>(workspace) % 16 != 0)
    return set_error(RUMPY_ERR_ARG, "bicubic_upsample: N=%d C=%d H=%d W=%d scale=%d (scale 2..8, N*C <= 65535, "
                     "16-byte aligned workspace)", N, C, H, W, scale);
  if (int e = device_info(nullptr)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int OH = H * scale, OW = W * scale;
  bicubic_coeff_kernel<<<(OW + (OH + 3) / 4 + 127) / 128, 128, 0, s>>>(workspace, H, W, scale);
  if (int e = check_launch("bicubic_coeff")) return e;
  const int vec = OW % 4 == 0 && reinterpret_cast<uintptr_t>(out_nchw) % 16 == 0;
  bicubic_up_kernel<<<dim3((OW + kBicTW - 1) / kBicTW, (OH + kBicTH - 1) / kBicTH, N * C), 256, 0, s>>>(
      lr_nchw, out_nchw, workspace, H, W, scale, vec);
  return check_launch("bicubic_upsample");
}

}  // extern "C"
