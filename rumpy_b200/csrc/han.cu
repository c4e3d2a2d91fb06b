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

// =====================================================================================================================
// Backward of the two attention modules (training).  Same bandwidth-bound style: a handful of passes over fp32 maps.
//   CSAM:  out = x (1 + g s),  s = sigmoid(c),  c = b + sum_t w[t] x[.+off(t)]
//          dc = d x g s (1 - s);  dx = d (1 + g s) + sum_t w[t] dc[.-off(t)];  dg = sum d x s;  db = sum dc;
//          dw[t] = sum dc x[.+off(t)]
//   LAM:   out_i = g sum_j A_ij x_j + x_i,  A = softmax_j(max_k E_ik - E_ij),  E = X X^T
//          M_ij = <G_i, x_j>;  dg = sum_ij A_ij M_ij;  dz_ij = A_ij (g M_ij - sum_k A_ik g M_ik);  dE = -dz;
//          dX_j = G_j + g sum_i A_ij G_i + sum_i (dE_ij + dE_ji) x_i
// =====================================================================================================================
constexpr int kCsamPart = 32;   // per-block partial: [0..26] dw, [27] db, [28] dgamma

// pass 1, one thread per (pixel, channel): dc, s (kept for pass 2), bf16 copy of the out2 half of dcat, partial sums
__global__ void csam_bwd1_kernel(const float* __restrict__ x, const float* __restrict__ dcat, const float* __restrict__ w,
                                 const float* __restrict__ b, const float* __restrict__ gamma, float* __restrict__ dc_out,
                                 float* __restrict__ sig_out, __nv_bfloat16* __restrict__ d2_b, float* __restrict__ part,
                                 int N, int H, int W) {
  __shared__ float w_s[27];
  __shared__ float red[8][kCsamPart];
  if (threadIdx.x < 27) w_s[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const size_t total = size_t(N) * H * W * 64;
  const float bias = __ldg(b), g = __ldg(gamma);
  float acc[29];
#pragma unroll
  for (int i = 0; i < 29; ++i) acc[i] = 0.f;
  for (size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c = int(e & 63);
    const size_t pix = e >> 6;
    const int xx = int(pix % W), yy = int((pix / W) % H);
    float nb[27];
    float cv = bias;
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc)
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int t = (dc + 1) * 9 + (dy + 1) * 3 + (dx + 1);
          const bool in = c + dc >= 0 && c + dc < 64 && yy + dy >= 0 && yy + dy < H && xx + dx >= 0 && xx + dx < W;
          nb[t] = in ? x[ptrdiff_t(e) + ptrdiff_t(dy * W + dx) * 64 + dc] : 0.f;
          cv = fmaf(w_s[t], nb[t], cv);
        }
    const float sg = 1.f / (1.f + __expf(-cv));
    const float d = dcat[pix * 128 + c], xv = nb[13];
    const float dcv = d * xv * g * sg * (1.f - sg);
    dc_out[e] = dcv;
    sig_out[e] = sg;
    d2_b[e] = __float2bfloat16_rn(dcat[pix * 128 + 64 + c]);
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = fmaf(dcv, nb[t], acc[t]);
    acc[27] += dcv;
    acc[28] = fmaf(d * xv, sg, acc[28]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 29; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 29) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    part[size_t(blockIdx.x) * kCsamPart + threadIdx.x] = v;
  }
}

// pass 2: dx = d (1 + g s) + sum_t w[t] dc[. - off(t)]  -> fp32 (the LAM apply adds its share of dX_0 on top)
__global__ void csam_bwd2_kernel(const float* __restrict__ dcat, const float* __restrict__ dc_in,
                                 const float* __restrict__ sig, const float* __restrict__ w,
                                 const float* __restrict__ gamma, float* __restrict__ dx_out, int N, int H, int W) {
  __shared__ float w_s[27];
  if (threadIdx.x < 27) w_s[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const size_t total = size_t(N) * H * W * 64;
  const float g = __ldg(gamma);
  for (size_t e = blockIdx.x * size_t(blockDim.x) + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c = int(e & 63);
    const size_t pix = e >> 6;
    const int xx = int(pix % W), yy = int((pix / W) % H);
    float v = dcat[pix * 128 + c] * fmaf(g, sig[e], 1.f);
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc)
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          // the output element e' = e - off(t) used x[e] through tap t
          const bool in = c - dc >= 0 && c - dc < 64 && yy - dy >= 0 && yy - dy < H && xx - dx >= 0 && xx - dx < W;
          if (in) v = fmaf(w_s[(dc + 1) * 9 + (dy + 1) * 3 + (dx + 1)],
                           dc_in[ptrdiff_t(e) - ptrdiff_t(dy * W + dx) * 64 - dc], v);
        }
    dx_out[e] = v;
  }
}

// M[n][i][j] = <G_i, x_j>: grid (chunks, N, L): z = i; partial[n][i][chunk][j]
__global__ void __launch_bounds__(256) lam_gx_kernel(LamPtrs p, const float* __restrict__ G, float* __restrict__ partial,
                                                     int HW) {
  const int n = blockIdx.y, i = blockIdx.z;
  float acc[kLamLayers];
#pragma unroll
  for (int j = 0; j < kLamLayers; ++j) acc[j] = 0.f;
  const int total = HW * 16;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pix = e >> 4, c4 = e & 15;
    const float4 g = reinterpret_cast<const float4*>(G)[(size_t(n) * HW + pix) * (kLamLayers * 16) + i * 16 + c4];
#pragma unroll
    for (int j = 0; j < kLamLayers; ++j) {
      const float4 v = reinterpret_cast<const float4*>(p.x[j])[(size_t(n) * HW + pix) * 16 + c4];
      acc[j] += (g.x * v.x + g.y * v.y) + (g.z * v.z + g.w * v.w);
    }
  }
  __shared__ float red[8][kLamLayers];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < kLamLayers; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLamLayers) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    partial[((size_t(n) * kLamLayers + i) * gridDim.x + blockIdx.x) * kLamLayers + threadIdx.x] = v;
  }
}

// per image: coefG[j][i] = g A_ij + [i == j], coefX[j][i] = dE_ij + dE_ji, dgamma[n] = sum_ij A_ij M_ij
__global__ void lam_coef_kernel(const float* __restrict__ partial, const float* __restrict__ att,
                                const float* __restrict__ gamma, float* __restrict__ coef, float* __restrict__ dgamma,
                                int chunks) {
  constexpr int L = kLamLayers;
  const int n = blockIdx.x, t = threadIdx.x;
  __shared__ double M_s[L][L], dE_s[L][L];
  __shared__ float A_s[L][L];
  const double g = double(__ldg(gamma));
  if (t < L * L) {
    const int i = t / L, j = t % L;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += double(partial[((size_t(n) * L + i) * chunks + c) * L + j]);
    M_s[i][j] = s;
    A_s[i][j] = att[(size_t(n) * L + i) * L + j];
  }
  __syncthreads();
  if (t < L * L) {
    const int i = t / L, j = t % L;
    double dot = 0.0;
    for (int k = 0; k < L; ++k) dot += double(A_s[i][k]) * g * M_s[i][k];
    dE_s[i][j] = -double(A_s[i][j]) * (g * M_s[i][j] - dot);
  }
  __syncthreads();
  if (t < L * L) {
    const int j = t / L, i = t % L;
    coef[(size_t(n) * 2 * L + j) * L + i] = float(g * double(A_s[i][j])) + (i == j ? 1.f : 0.f);
    coef[(size_t(n) * 2 * L + L + j) * L + i] = float(dE_s[i][j] + dE_s[j][i]);
  }
  if (t == 0) {
    double s = 0.0;
    for (int i = 0; i < L; ++i)
      for (int j = 0; j < L; ++j) s += double(A_s[i][j]) * M_s[i][j];
    dgamma[n] = float(s);
  }
}

struct LamOutPtrs { float* dx[kLamLayers]; };
// dX_j = sum_i coefG[j][i] G_i + coefX[j][i] x_i  (dX_0 is ADDED to what csam_bwd2 left there and also emitted in bf16)
__global__ void __launch_bounds__(256) lam_bwd_apply_kernel(LamPtrs p, const float* __restrict__ G,
                                                            const float* __restrict__ coef, LamOutPtrs o,
                                                            __nv_bfloat16* __restrict__ dx0_b, int HW) {
  constexpr int L = kLamLayers;
  const int n = blockIdx.y;
  __shared__ float c_s[2 * L * L];
  for (int i = threadIdx.x; i < 2 * L * L; i += blockDim.x) c_s[i] = coef[size_t(n) * 2 * L * L + i];
  __syncthreads();
  const int total = HW * 16;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int pix = e >> 4, c4 = e & 15;
    const size_t off = (size_t(n) * HW + pix) * 16 + c4;
    float4 gv[L], xv[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      gv[i] = reinterpret_cast<const float4*>(G)[(size_t(n) * HW + pix) * (L * 16) + i * 16 + c4];
      xv[i] = reinterpret_cast<const float4*>(p.x[i])[off];
    }
#pragma unroll
    for (int j = 0; j < L; ++j) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < L; ++i) {
        const float a = c_s[j * L + i], b2 = c_s[(L + j) * L + i];
        s.x = fmaf(a, gv[i].x, fmaf(b2, xv[i].x, s.x)); s.y = fmaf(a, gv[i].y, fmaf(b2, xv[i].y, s.y));
        s.z = fmaf(a, gv[i].z, fmaf(b2, xv[i].z, s.z)); s.w = fmaf(a, gv[i].w, fmaf(b2, xv[i].w, s.w));
      }
      float4* dst = reinterpret_cast<float4*>(o.dx[j]) + off;
      if (j == 0) {
        const float4 prev = *dst;
        s.x += prev.x; s.y += prev.y; s.z += prev.z; s.w += prev.w;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(s.x, s.y), hi = __floats2bfloat162_rn(s.z, s.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(dx0_b)[off] = pk;
      }
      *dst = s;
    }
  }
}

// fixed-order sums -> d(csa.conv.weight) [27], d(csa.conv.bias), d(csa.gamma), d(la.gamma)
__global__ void han_param_grad_kernel(const float* __restrict__ csam_part, int blocks, const float* __restrict__ lam_dg,
                                      int N, float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dg_csa,
                                      float* __restrict__ dg_la) {
  const int t = threadIdx.x;
  if (t < 29) {
    float s = 0.f;
    for (int k = 0; k < blocks; ++k) s += csam_part[size_t(k) * kCsamPart + t];
    if (t < 27) dw[t] = s; else if (t == 27) db[0] = s; else dg_csa[0] = s;
  } else if (t == 32) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += lam_dg[n];
    dg_la[0] = s;
  }
}

int han_bwd_scratch_floats(int N, int csam_blocks) {
  return csam_blocks * kCsamPart + N * kLamLayers * kLamChunks * kLamLayers + N * 2 * kLamLayers * kLamLayers + N;
}
int han_csam_blocks(int N, int H, int W) {
  int sms = 0;
  if (device_info(&sms)) return 148 * 8;
  const size_t total = size_t(N) * H * W * 64;
  size_t blocks = (total + 255) / 256;
  return int(blocks > size_t(sms) * 8 ? size_t(sms) * 8 : blocks);
}

int csam_bwd_launch(const float* x, const float* dcat, const float* w, const float* b, const float* gamma, float* dc,
                    float* sig, void* d2_b, float* dx0, float* scratch, int N, int H, int W, cudaStream_t s) {
  const int blocks = han_csam_blocks(N, H, W);
  csam_bwd1_kernel<<<blocks, 256, 0, s>>>(x, dcat, w, b, gamma, dc, sig, static_cast<__nv_bfloat16*>(d2_b), scratch, N, H, W);
  if (int e = check_launch("csam_bwd1")) return e;
  csam_bwd2_kernel<<<blocks, 256, 0, s>>>(dcat, dc, sig, w, gamma, dx0, N, H, W);
  return check_launch("csam_bwd2");
}

int lam_bwd_launch(const float* const* stack, const float* G, const float* att, const float* gamma, float* const* dx,
                   void* dx0_b, float* scratch, int csam_blocks, int N, int HW, cudaStream_t s) {
  LamPtrs p;
  LamOutPtrs o;
  for (int i = 0; i < kLamLayers; ++i) { p.x[i] = stack[i]; o.dx[i] = dx[i]; }
  float* partial = scratch + size_t(csam_blocks) * kCsamPart;
  float* coef = partial + size_t(N) * kLamLayers * kLamChunks * kLamLayers;
  float* dg = coef + size_t(N) * 2 * kLamLayers * kLamLayers;
  lam_gx_kernel<<<dim3(kLamChunks, N, kLamLayers), 256, 0, s>>>(p, G, partial, HW);
  if (int e = check_launch("lam_gx")) return e;
  lam_coef_kernel<<<N, 128, 0, s>>>(partial, att, gamma, coef, dg, kLamChunks);
  if (int e = check_launch("lam_coef")) return e;
  lam_bwd_apply_kernel<<<dim3(kLamChunks, N), 256, 0, s>>>(p, G, coef, o, static_cast<__nv_bfloat16*>(dx0_b), HW);
  return check_launch("lam_bwd_apply");
}

int han_param_grad_launch(const float* scratch, int csam_blocks, int N, float* dw, float* db, float* dg_csa, float* dg_la,
                          cudaStream_t s) {
  const float* lam_dg = scratch + size_t(csam_blocks) * kCsamPart + size_t(N) * kLamLayers * kLamChunks * kLamLayers +
                        size_t(N) * 2 * kLamLayers * kLamLayers;
  han_param_grad_kernel<<<1, 64, 0, s>>>(scratch, csam_blocks, lam_dg, N, dw, db, dg_csa, dg_la);
  return check_launch("han_param_grad");
}

const float* lam_att_ptr(const float* lam_scratch, int N) { return lam_scratch + size_t(N) * kLamChunks * kLamPairs; }

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
