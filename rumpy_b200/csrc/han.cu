// HAN (holistic attention network) attention modules around the RCAN trunk (reference
// SISR/models/advanced/HAN_blocks.py:8-76, architectures.py:331-394).  The residual groups run in the trunk kernels;
// what is left is bandwidth-bound fp32 work over the L = n_groups + 1 stacked feature maps (NHWC fp32, 64 channels):
//
//   LAM  (layer attention, HAN_blocks.py:8-41):  E[i][j] = <x_i, x_j> over all C*H*W elements of an image,
//        A = softmax_j(max_j E[i][j] - E[i][j]),  out_i = gamma * sum_j A[i][j] x_j + x_i, written as ONE bf16 NHWC
//        tensor with L*64 channels (the operand of last_conv)
//   CSAM (channel-spatial attention, HAN_blocks.py:44-76): 3x3x3 conv over the (C, H, W) volume -> sigmoid -> gamma,
//        x * (.) + x, written next to bf16(out2) as the 128-channel operand of `last`
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/rumpy_b200.h"
#include "host_util.cuh"

namespace rb {

constexpr int kLamPairs = kLamLayers * (kLamLayers + 1) / 2;   // 66
constexpr int kLamChunks = 74;                                  // CTAs per image (16 images -> 8 CTAs per SM)

struct LamPtrs { const float* x[kLamLayers]; };

// grid (kLamChunks, N), block 256: partial[n][chunk][pair] = sum over a slice of x_i * x_j (i <= j)
__global__ void __launch_bounds__(256) lam_energy_kernel(LamPtrs p, float* __restrict__ partial, int elems4) {
  const int n = blockIdx.y;
  float acc[kLamPairs];
#pragma unroll
  for (int k = 0; k < kLamPairs; ++k) acc[k] = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < elems4; e += gridDim.x * blockDim.x) {
    float4 v[kLamLayers];
#pragma unroll
    for (int i = 0; i < kLamLayers; ++i) v[i] = reinterpret_cast<const float4*>(p.x[i])[size_t(n) * elems4 + e];
    int k = 0;
#pragma unroll
    for (int i = 0; i < kLamLayers; ++i)
#pragma unroll
      for (int j = i; j < kLamLayers; ++j, ++k)
        acc[k] += (v[i].x * v[j].x + v[i].y * v[j].y) + (v[i].z * v[j].z + v[i].w * v[j].w);
  }
  __shared__ float red[8][kLamPairs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kLamPairs; ++k) {
    float s = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < kLamPairs) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[(size_t(n) * gridDim.x + blockIdx.x) * kLamPairs + threadIdx.x] = s;
  }
}

// grid N, block 128: fixed-order sums of the partials (double) -> attention[n][i][j]
__global__ void lam_attention_kernel(const float* __restrict__ partial, float* __restrict__ att, int chunks) {
  const int n = blockIdx.x, t = threadIdx.x;
  __shared__ double e_s[kLamLayers][kLamLayers];
  if (t < kLamLayers * kLamLayers) {
    const int i = t / kLamLayers, j = t % kLamLayers;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const int pair = a * kLamLayers - a * (a - 1) / 2 + (b - a);
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += double(partial[(size_t(n) * chunks + c) * kLamPairs + pair]);
    e_s[i][j] = s;
  }
  __syncthreads();
  if (t < kLamLayers) {
    double mx = e_s[t][0];
    for (int j = 1; j < kLamLayers; ++j) mx = fmax(mx, e_s[t][j]);
    double en[kLamLayers], top = -1e300, sum = 0.0;
    for (int j = 0; j < kLamLayers; ++j) { en[j] = mx - e_s[t][j]; top = fmax(top, en[j]); }
    for (int j = 0; j < kLamLayers; ++j) { en[j] = exp(en[j] - top); sum += en[j]; }
    for (int j = 0; j < kLamLayers; ++j) att[(size_t(n) * kLamLayers + t) * kLamLayers + j] = float(en[j] / sum);
  }
}

// one thread per (image, pixel, 4 channels): out[n][pix][i*64 + c] = gamma * sum_j A[i][j] x_j + x_i  (bf16)
__global__ void __launch_bounds__(256) lam_apply_kernel(LamPtrs p, const float* __restrict__ att,
                                                        const float* __restrict__ gamma, __nv_bfloat16* __restrict__ out,
                                                        int HW) {
  const int n = blockIdx.y;
  __shared__ float a_s[kLamLayers * kLamLayers];
  if (threadIdx.x < kLamLayers * kLamLayers) a_s[threadIdx.x] = att[size_t(n) * kLamLayers * kLamLayers + threadIdx.x];
  __syncthreads();
  const float g = __ldg(gamma);
  const int total = HW * 16;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pix = e >> 4, c4 = e & 15;
    float4 v[kLamLayers];
#pragma unroll
    for (int i = 0; i < kLamLayers; ++i)
      v[i] = reinterpret_cast<const float4*>(p.x[i])[(size_t(n) * HW + pix) * 16 + c4];
    __nv_bfloat16* o = out + (size_t(n) * HW + pix) * (kLamLayers * 64) + c4 * 4;
#pragma unroll
    for (int i = 0; i < kLamLayers; ++i) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < kLamLayers; ++j) {
        const float a = a_s[i * kLamLayers + j];
        s.x = fmaf(a, v[j].x, s.x); s.y = fmaf(a, v[j].y, s.y); s.z = fmaf(a, v[j].z, s.z); s.w = fmaf(a, v[j].w, s.w);
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf(g, s.x, v[i].x), fmaf(g, s.y, v[i].y));
      const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf(g, s.z, v[i].z), fmaf(g, s.w, v[i].w));
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(o + i * 64) = pk;
    }
  }
}

// one thread per (image, pixel, channel): cat[..][0:64] = bf16(x * gamma*sigmoid(conv3d(x)) + x), cat[..][64:128] = bf16(out2)
__global__ void csam_cat_kernel(const float* __restrict__ x, const float* __restrict__ out2, const float* __restrict__ w,
                                const float* __restrict__ b, const float* __restrict__ gamma,
                                __nv_bfloat16* __restrict__ cat, int N, int H, int W) {
  __shared__ float w_s[27];
  if (threadIdx.x < 27) w_s[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const size_t total = size_t(N) * H * W * 64;
  const float bias = __ldg(b), g = __ldg(gamma);
  for (size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c = int(e & 63);
    const size_t pix = e >> 6;
    const int xx = int(pix % W), yy = int((pix / W) % H);
    float acc = bias;
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) {
      const int cc = c + dc;
      if (cc < 0 || cc > 63) continue;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y2 = yy + dy;
        if (y2 < 0 || y2 >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x2 = xx + dx;
          if (x2 < 0 || x2 >= W) continue;
          acc = fmaf(w_s[(dc + 1) * 9 + (dy + 1) * 3 + (dx + 1)],
                     x[(ptrdiff_t(e) + ptrdiff_t(dy * W + dx) * 64 + dc)], acc);
        }
      }
    }
    const float v = x[e];
    const float r = g * (1.f / (1.f + __expf(-acc)));
    cat[pix * 128 + c] = __float2bfloat16_rn(fmaf(v, r, v));
    cat[pix * 128 + 64 + c] = __float2bfloat16_rn(out2[e]);
  }
}

int lam_workspace_floats(int N) { return N * (kLamChunks * kLamPairs + kLamLayers * kLamLayers); }

int lam_launch(const float* const* stack, float* scratch, const float* gamma, void* out_bf16, int N, int HW,
               cudaStream_t s) {
  LamPtrs p;
  for (int i = 0; i < kLamLayers; ++i) p.x[i] = stack[i];
  float* partial = scratch;
  float* att = scratch + size_t(N) * kLamChunks * kLamPairs;
  lam_energy_kernel<<<dim3(kLamChunks, N), 256, 0, s>>>(p, partial, HW * 16);
  if (int e = check_launch("lam_energy")) return e;
  lam_attention_kernel<<<N, 128, 0, s>>>(partial, att, kLamChunks);
  if (int e = check_launch("lam_attention")) return e;
  lam_apply_kernel<<<dim3(kLamChunks, N), 256, 0, s>>>(p, att, gamma, static_cast<__nv_bfloat16*>(out_bf16), HW);
  return check_launch("lam_apply");
}

int csam_cat_launch(const float* x, const float* out2, const float* w, const float* b, const float* gamma, void* cat_bf16,
                    int N, int H, int W, cudaStream_t s) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  const size_t total = size_t(N) * H * W * 64;
  size_t blocks = (total + 255) / 256;
  if (blocks > size_t(sms) * 16) blocks = size_t(sms) * 16;
  csam_cat_kernel<<<int(blocks), 256, 0, s>>>(x, out2, w, b, gamma, static_cast<__nv_bfloat16*>(cat_bf16), N, H, W);
  return check_launch("csam_cat");
}

}  // namespace rb
