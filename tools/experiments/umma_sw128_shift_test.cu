// Experiment (GPU box): can tcgen05.mma read a 128B-swizzled, pixel-major A operand (the layout TMA writes for an
// NHWC bf16 box) at a start address that is NOT aligned to the 1024-byte swizzle atom?  If so, ONE halo box
// {64 ch, 16 px, 18 rows} could serve all nine taps of a 16-row x 8-pixel tile (36 KB per tile instead of the
// three {64, 16, 10} boxes = 60 KB the per-layer conv kernel fetches now): tap (ky,kx) = start address moved by
// (ky*16 + kx) * 128 B, SBO = 2048 (next image row = next 8-pixel group).
// -DPITCH=10: the same with the minimal {64, 10, 18} box (SBO = 1280, not a multiple of the atom).
// Variant 0: descriptor base-offset field 0; variant 1: base offset = (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/umma_sw128_shift_test tools/experiments/umma_sw128_shift_test.cu
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <cmath>
#include <cstdlib>
#include <vector>
using namespace rb;

#ifndef PITCH
#define PITCH 16
#endif
constexpr int kPitch = PITCH, kRows = 18, kPix = kPitch * kRows;   // PITCH 10: the minimal halo box

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __nv_bfloat16* act, const __nv_bfloat16* w, float* out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;                 // [9][64 rows][128 B], 128B swizzle
  uint8_t* a_s = smem + 9 * 8192;      // [kPix][128 B], 128B swizzle on the absolute address (as TMA writes it)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kPix * 8; i += 128) {
    const int p = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + p * 128 + ((c ^ (p & 7)) << 4)) = *reinterpret_cast<const uint4*>(act + p * 64 + c * 8);
  }
  for (int i = tid; i < 9 * 64 * 8; i += 128) {
    const int tap = i / 512, r = (i >> 3) & 63, c = i & 7;
    *reinterpret_cast<uint4*>(w_s + tap * 8192 + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(w + (tap * 64 + r) * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<64>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      for (int k = 0; k < 4; ++k) {
        const uint32_t start = smem_u32(a_s + (ky * kPitch + kx) * 128);
        uint64_t adesc = make_smem_desc(start, 16, kPitch * 128, kLayoutSw128) + uint64_t(2 * k);
        if (variant == 1) adesc |= uint64_t((start >> 7) & 7) << 49;
        const uint64_t bdesc = make_smem_desc(smem_u32(w_s + tap * 8192), 16, 1024, kLayoutSw128) + uint64_t(2 * k);
        umma_bf16(tmem, adesc, bdesc, kIdesc, (tap | k) != 0);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + h * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[tid * 64 + h * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

// Rate: `iters` x 36 MMAs (nine taps x four K steps) from one issuing thread, nothing else running.
//   layout 0: aligned reference -- every tap reads the box at offset 0 with SBO = 1024 (dense atoms)
//   layout 1: the shifted taps of the conv (start (ky*pitch+kx)*128, SBO = pitch*128)
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int layout) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + 9 * 8192;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (9 * 8192 + kPix * 128) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0, 0x3c003c00u, 0);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<64>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap % 3, kx = tap / 3;
        const uint32_t start = smem_u32(a_s + (layout ? (ky * kPitch + kx) * 128 : 0));
        const uint64_t adesc = make_smem_desc(start, 16, layout ? kPitch * 128 : 1024, kLayoutSw128);
        const uint64_t bdesc = make_smem_desc(smem_u32(w_s + tap * 8192), 16, 1024, kLayoutSw128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, (it | tap | k) != 0);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

int main() {
  std::vector<__nv_bfloat16> act(kPix * 64), w(9 * 64 * 64);
  std::vector<float> actf(kPix * 64), wf(9 * 64 * 64);
  srand(1);
  for (size_t i = 0; i < act.size(); ++i) { act[i] = __float2bfloat16(float(rand() % 17 - 8) / 8.f); actf[i] = __bfloat162float(act[i]); }
  for (size_t i = 0; i < w.size(); ++i) { w[i] = __float2bfloat16(float(rand() % 13 - 6) / 16.f); wf[i] = __bfloat162float(w[i]); }
  __nv_bfloat16 *dact, *dw; float* dout;
  cudaMalloc(&dact, act.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dact, act.data(), act.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + 9 * 8192 + kPix * 128;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ref(128 * 64, 0.f), got(128 * 64);
  for (int m = 0; m < 128; ++m) {
    const int ly = m >> 3, lx = m & 7;
    for (int co = 0; co < 64; ++co) {
      float s = 0.f;
      for (int tap = 0; tap < 9; ++tap) {
        const int p = (ly + tap / 3) * kPitch + lx + tap % 3;
        for (int ci = 0; ci < 64; ++ci) s += actf[p * 64 + ci] * wf[(tap * 64 + co) * 64 + ci];
      }
      ref[m * 64 + co] = s;
    }
  }
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, 128 * 64 * 4);
    shift_kernel<<<1, 128, smem>>>(dact, dw, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    int bad_rows = 0;
    for (int m = 0; m < 128; ++m) {
      double re = 0;
      for (int c = 0; c < 64; ++c) { re = fmax(re, fabs(got[m * 64 + c] - ref[m * 64 + c])); maxref = fmax(maxref, fabs(ref[m * 64 + c])); }
      maxerr = fmax(maxerr, re);
      bad_rows += re > 1e-3;
    }
    printf("variant %d (base offset %s): max |err| = %.4g (ref absmax %.4g), %d of 128 rows wrong -> %s\n", variant,
           variant == 0 ? "0" : "(start>>7)&7", maxerr, maxref, bad_rows, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
  }
  long long* dt; cudaMalloc(&dt, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int layout = 0; layout < 2; ++layout)
    for (int iters : {1, 4, 64}) {
      rate_kernel<<<1, 128, smem>>>(dt, iters, layout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, dt, 16, cudaMemcpyDeviceToHost);
      printf("layout %d (%s), %2d x 36 MMAs: issue %.1f cycles/MMA, issue+drain %.1f cycles/MMA\n", layout,
             layout ? "shifted taps, SBO = pitch*128" : "aligned, SBO = 1024", iters, double(h[0]) / (iters * 36), double(h[1]) / (iters * 36));
    }
  return 0;
}
