// Eval glue and training-patch pipeline on the device: the host-side numpy / PIL work that
// surrounds the trunk in the reference, restated as bandwidth-bound kernels so that a step moves uint8 images and a
// few scalars across PCIe instead of fp32 batches.
//
//   rumpy_psnr_y       clip -> Y of jpg-style YCbCr -> per-image PSNR   (reference sr_tools/metrics.py:33-44,
//                      image_tools/image_manipulation/image_functions.py:72-88, base_interface.py:208-222)
//   rumpy_quantize_u8  clip(x*255, 0, 255).astype(uint8) (truncation), NCHW fp32 -> NHWC uint8 (what gets saved:
//                      sr_tools/visualization.py:31-61)
//   rumpy_patch_batch  random crop + hflip / vflip / transpose + ToTensor for a whole batch of LR/HR pairs from
//                      uint8 images resident in HBM (image_functions.py:287-362, sr_tools/data_handler.py:570-645)
//
// All three are integer / byte work bound by HBM: one pass, coalesced along the fastest output dimension, grids sized
// from the element count; reductions run in a fixed order (deterministic).
#include <cuda_runtime.h>

#include <cstdint>

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

}  // extern "C"
