// CUDA-core kernels around the tensor-core convolution: weight packing, the thin head conv, the fused
// channel-attention (CALayer) FC + rescale + residual pass, L1 loss and Adam.  All are memory-bound
// streaming kernels: coalesced 16-byte accesses, grid sized in multiples of the SM count by the host.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace rb {

// ------------------------------------------------------------------------------------------------
// OIHW fp32 -> packed bf16 operand for conv3x3_tc_kernel.
//   forward : P[tap][row][ci]            = W[o(row)][ci][ky][kx],            tap = kx*3+ky (kx-major: the
//             three ky taps of one kx are contiguous -- they share one A halo box in the conv kernel)
//   dgrad   : P[tap][ci ][col]           = W[o(col)][ci][2-ky][2-kx]   (rows = Cin, K = Cout)
// o(row) applies the pixel-shuffle permutation when r > 1: row = q*(Cout/r^2) + c  <->  o = c*r^2 + q,
// so that output chunk q holds the channels of sub-pixel q (nn.PixelShuffle, reference common.py:33,40).
// Rows >= valid rows are zero filled (thin tail conv pads Cout 3 -> 16).
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ p, int cout, int cin,
                                    int rows_padded, int r, int dgrad) {
  const int rows = dgrad ? cin : rows_padded;   // rows of the packed [tap][rows][k] operand
  const int kdim = dgrad ? cout : cin;
  const size_t total = size_t(9) * rows * kdim;
  const int rr = r * r;
  const int cpp = cout / rr;  // channels per sub-pixel
  for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < total;
       idx += size_t(gridDim.x) * blockDim.x) {
    const int k = int(idx % kdim);
    const int row = int((idx / kdim) % rows);
    const int tap = int(idx / (size_t(kdim) * rows));
    const int kx = tap / 3, ky = tap % 3;   // packed tap order is kx-major: tap' = kx*3 + ky
    float v = 0.f;
    if (!dgrad) {
      if (row < cout) {
        const int o = (r > 1) ? (row % cpp) * rr + row / cpp : row;
        v = w[((size_t(o) * cin + k) * 3 + ky) * 3 + kx];
      }
    } else {
      const int o = (r > 1) ? (k % cpp) * rr + k / cpp : k;
      v = w[((size_t(o) * cin + row) * 3 + (2 - ky)) * 3 + (2 - kx)];
    }
    p[idx] = __float2bfloat16_rn(v);
  }
}

// Whole-network variant: one launch packs every conv (job list in device memory).
struct PackJob {
  const float* w; __nv_bfloat16* p; const float* b; float* bp;   // bp == nullptr: no packed bias
  int cout, cin, rows_padded, r, dgrad;
};
__global__ void pack_conv3x3_batched_kernel(const PackJob* __restrict__ jobs) {
  const PackJob jb = jobs[blockIdx.y];
  const int rows = jb.dgrad ? jb.cin : jb.rows_padded;
  const int kdim = jb.dgrad ? jb.cout : jb.cin;
  const size_t plane = size_t(rows) * kdim;   // elements per tap
  const int rr = jb.r * jb.r, cpp = jb.cout / rr;
  // one thread per (row, k): its 9 taps are 36 contiguous bytes of the OIHW tensor; the 9 stores are coalesced in k
  for (size_t idx = blockIdx.x * size_t(blockDim.x) + threadIdx.x; idx < plane;
       idx += size_t(gridDim.x) * blockDim.x) {
    const int k = int(idx % kdim);
    const int row = int(idx / kdim);
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = 0.f;
    if (!jb.dgrad) {
      if (row < jb.cout) {
        const int o = (jb.r > 1) ? (row % cpp) * rr + row / cpp : row;
        const float* src = jb.w + (size_t(o) * jb.cin + k) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = src[t];            // src[ky * 3 + kx]
      }
    } else {
      const int o = (jb.r > 1) ? (k % cpp) * rr + k / cpp : k;
      const float* src = jb.w + (size_t(o) * jb.cin + row) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = src[8 - t];          // 180-degree rotation: (2 - ky) * 3 + (2 - kx)
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)                          // packed tap index = kx * 3 + ky
        jb.p[size_t(kx * 3 + ky) * plane + idx] = __float2bfloat16_rn(v[ky * 3 + kx]);
  }
  if (jb.bp != nullptr && blockIdx.x == 0) {
    for (int row = threadIdx.x; row < jb.rows_padded; row += blockDim.x) {
      float v = 0.f;
      if (row < jb.cout) v = jb.b[(jb.r > 1) ? (row % cpp) * rr + row / cpp : row];
      jb.bp[row] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Meta-attention (reference attention_manipulators/q_layer.py:5-45 `ParaCALayer`; used by QRCAB.forward
// architectures.py:198-219 and ParamResBlock.forward :484-493; QCALayer style 'modulate' :113-116): per
// (image, channel) multipliers that depend on the metadata only, so all blocks are evaluated up front in one launch:
//   q[n][c] = sigmoid(W2 act(W1 meta[n] + b1) + b2)[c]     (blocks with a q-layer; act = ReLU or identity)
//   q[n][c] *= meta[n][c or 0]                             (style 'modulate')
// grid (jobs, N), block C (= n_feats, 64..256).
// ------------------------------------------------------------------------------------------------
struct QScaleJobDev { const float *w1, *b1, *w2, *b2; float* out; };
__global__ void q_scale_kernel(const QScaleJobDev* __restrict__ jobs, const float* __restrict__ meta, int M, int hidden,
                               int modulate, int relu) {
  extern __shared__ float qs_smem[];   // [M] metadata, [hidden] hidden units
  float* meta_s = qs_smem;
  float* hid_s = qs_smem + M;
  const QScaleJobDev jb = jobs[blockIdx.x];
  const int n = blockIdx.y, c = threadIdx.x, C = blockDim.x;
  for (int m = c; m < M; m += C) meta_s[m] = meta[size_t(n) * M + m];
  __syncthreads();
  float q = 1.f;
  if (jb.w1 != nullptr) {
    for (int t = c; t < hidden; t += C) {
      float a = jb.b1[t];
      for (int m = 0; m < M; ++m) a = fmaf(jb.w1[size_t(t) * M + m], meta_s[m], a);
      hid_s[t] = relu ? fmaxf(a, 0.f) : a;
    }
    __syncthreads();
    float a = jb.b2[c];
    for (int t = 0; t < hidden; ++t) a = fmaf(jb.w2[size_t(c) * hidden + t], hid_s[t], a);
    q = 1.f / (1.f + __expf(-a));
  }
  if (modulate) q *= meta_s[M == 1 ? 0 : c];
  jb.out[size_t(n) * C + c] = q;
}

// bias in packed-row order (pixel-shuffle permutation), zero padded
__global__ void pack_bias_kernel(const float* __restrict__ b, float* __restrict__ p, int cout, int rows_padded,
                                 int r) {
  const int rr = r * r, cpp = cout / rr;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows_padded; row += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (row < cout) v = b[(r > 1) ? (row % cpp) * rr + row / cpp : row];
    p[row] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Head conv: in_feats (3) -> C, fp32 NCHW input (the reference's tensor layout), fp32 math,
// writes the fp32 residual stream and its bf16 operand copy in NHWC.  (architectures.py:153,172)
// One thread = one pixel x 8 output channels; weights transposed in smem as [ci*9+tap][C].
// ------------------------------------------------------------------------------------------------
template <int CIN_MAX = 4>
__global__ void head_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ bias, float* __restrict__ yf,
                                 __nv_bfloat16* __restrict__ yb, int N, int H, int W, int cin, int C) {
  extern __shared__ float wsm[];  // [cin*9][C] then bias[C]
  const int kk = cin * 9;
  for (int i = threadIdx.x; i < kk * C; i += blockDim.x) {
    const int o = i % C, k = i / C;          // k = ci*9 + tap
    wsm[i] = w[size_t(o) * kk + k];
  }
  float* bsm = wsm + kk * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) bsm[i] = bias[i];
  __syncthreads();
  const int groups = C / 8;
  const size_t npix = size_t(N) * H * W;
  for (size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x; t < npix * groups;
       t += size_t(gridDim.x) * blockDim.x) {
    const int g = int(t % groups);
    const size_t pix = t / groups;
    const int xw = int(pix % W), yh = int((pix / W) % H), n = int(pix / (size_t(W) * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bsm[g * 8 + j];
    for (int ci = 0; ci < cin; ++ci) {
      const float* xp = x + (size_t(n) * cin + ci) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = yh + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = xw + kx - 1;
          if (xx < 0 || xx >= W) continue;
          const float v = __ldg(xp + size_t(yy) * W + xx);
          const float* wp = wsm + (ci * 9 + ky * 3 + kx) * C + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
        }
      }
    }
    float4* of = reinterpret_cast<float4*>(yf + pix * C + g * 8);
    of[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    of[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[0], acc[1]), b1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[4], acc[5]), b3 = __floats2bfloat162_rn(acc[6], acc[7]);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
    o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(yb + pix * C + g * 8) = o;
  }
}

// Large images: thousands of per-tile pool partials per image.  Compact them to `slices` rows per image first so
// that every CTA of ca_apply_kernel reads a handful of rows instead of all of them.  grid (slices, N).
__global__ void pool_compact_kernel(const float* __restrict__ pool_partial, int partials_per_img,
                                    float* __restrict__ compact, int C) {
  extern __shared__ float red[];
  const int n = blockIdx.y, slices = gridDim.x;
  const int begin = int((long long)partials_per_img * blockIdx.x / slices);
  const int end = int((long long)partials_per_img * (blockIdx.x + 1) / slices);
  const int lanes = blockDim.x / C, c = threadIdx.x % C, lane = threadIdx.x / C;
  float s0 = 0.f, s1 = 0.f;
  if (lane < lanes) {
    const float* pp = pool_partial + size_t(n) * partials_per_img * C + c;
    int i = begin + lane;
    for (; i + lanes < end; i += 2 * lanes) { s0 += pp[size_t(i) * C]; s1 += pp[size_t(i + lanes) * C]; }
    if (i < end) s0 += pp[size_t(i) * C];
  }
  red[threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * C + threadIdx.x];
    compact[(size_t(n) * slices + blockIdx.x) * C + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Channel attention, fused:  y = sigmoid(W2 relu(W1 mean(u) + b1) + b2);  x_out = x_in + u * y
// (CALayer architectures.py:41-44 + the RCAB skip :83).  The global average pool arrives as per-tile
// partial sums written by conv2's epilogue, so `u` is read exactly once here.  Every CTA recomputes
// the tiny FC for its image (C<=256, C/r hidden) in smem -- cheaper than another launch.
// grid = (chunks, N); block = 256.  u is fp32 or bf16 NHWC.
// ------------------------------------------------------------------------------------------------
template <bool U_F32>
__global__ void ca_apply_kernel(const float* __restrict__ pool_partial, int partials_per_img,
                                const void* __restrict__ u_, const float* __restrict__ x_in,
                                const float* __restrict__ w1, const float* __restrict__ b1,
                                const float* __restrict__ w2, const float* __restrict__ b2,
                                float* __restrict__ x_out, __nv_bfloat16* __restrict__ x_out_b,
                                float* __restrict__ save_mean, float* __restrict__ save_hid,
                                float* __restrict__ save_y, int HW, int C, int Cr,
                                const float* __restrict__ q_scale /* [N][C] or nullptr (Q-RCAN) */) {
  __shared__ float mean_s[256], y_s[256], hid_s[64], red_s[256];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: pool partials / u come from the previous kernel
  {
    // per-image channel sums from the conv epilogue's per-tile partials: blockDim/C thread groups split the
    // partial list, 4 independent loads in flight each (a serial chain of L2 round trips was 11 us here);
    // fixed summation order -> deterministic
    const int groups = blockDim.x / C;
    const int g = tid / C, c = tid % C;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (g < groups) {
      const float* pp = pool_partial + size_t(n) * partials_per_img * C + c;
      int i = g;
      for (; i + 3 * groups < partials_per_img; i += 4 * groups) {
        s0 += pp[size_t(i) * C]; s1 += pp[size_t(i + groups) * C];
        s2 += pp[size_t(i + 2 * groups) * C]; s3 += pp[size_t(i + 3 * groups) * C];
      }
      for (; i < partials_per_img; i += groups) s0 += pp[size_t(i) * C];
    }
    red_s[tid] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < C) {
      float s = 0.f;
      for (int k = 0; k < groups; ++k) s += red_s[k * C + tid];
      mean_s[tid] = s / float(HW);
    }
  }
  __syncthreads();
  {  // hidden layer: one warp per hidden unit (strided), warp-shuffle reduction over C
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < Cr; j += nwarps) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s = fmaf(w1[j * C + c], mean_s[c], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) hid_s[j] = fmaxf(s + b1[j], 0.f);
    }
  }
  __syncthreads();
  if (tid < C) {
    float s = b2[tid];
    for (int j = 0; j < Cr; ++j) s = fmaf(w2[tid * Cr + j], hid_s[j], s);
    const float yr = 1.f / (1.f + __expf(-s));
    y_s[tid] = q_scale != nullptr ? yr * q_scale[size_t(n) * C + tid] : yr;
    if (blockIdx.x == 0 && save_y != nullptr) save_y[n * C + tid] = yr;   // backward wants the raw sigmoid
  }
  __syncthreads();
  if (blockIdx.x == 0 && save_y != nullptr) {
    if (tid < C) save_mean[n * C + tid] = mean_s[tid];
    if (tid < Cr) save_hid[n * Cr + tid] = hid_s[tid];
  }
  // elementwise: 4 channels per thread
  const int vec_per_pix = C / 4;
  const size_t total = size_t(HW) * vec_per_pix;
  const size_t base = size_t(n) * HW * C;
  auto load_u = [&](size_t off) -> float4 {
    if (U_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(u_) + off);
    const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(u_) + off);
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
  };
  auto emit = [&](size_t i, const float4& uu, const float4& xi) {
    const int c4 = int(i % vec_per_pix) * 4;
    const size_t off = base + i * 4;
    float4 o;
    o.x = fmaf(uu.x, y_s[c4], xi.x); o.y = fmaf(uu.y, y_s[c4 + 1], xi.y);
    o.z = fmaf(uu.z, y_s[c4 + 2], xi.z); o.w = fmaf(uu.w, y_s[c4 + 3], xi.w);
    *reinterpret_cast<float4*>(x_out + off) = o;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(x_out_b + off) = pk;
  };
  // four vectors (8 independent 16-byte loads) in flight per thread
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t i = blockIdx.x * size_t(blockDim.x) + tid;
  for (; i + 3 * stride < total; i += 4 * stride) {
    const float4 u0 = load_u(base + i * 4), u1 = load_u(base + (i + stride) * 4);
    const float4 u2 = load_u(base + (i + 2 * stride) * 4), u3 = load_u(base + (i + 3 * stride) * 4);
    const float4 x0 = *reinterpret_cast<const float4*>(x_in + base + i * 4);
    const float4 x1 = *reinterpret_cast<const float4*>(x_in + base + (i + stride) * 4);
    const float4 x2 = *reinterpret_cast<const float4*>(x_in + base + (i + 2 * stride) * 4);
    const float4 x3 = *reinterpret_cast<const float4*>(x_in + base + (i + 3 * stride) * 4);
    emit(i, u0, x0); emit(i + stride, u1, x1); emit(i + 2 * stride, u2, x2); emit(i + 3 * stride, u3, x3);
  }
  for (; i < total; i += stride)
    emit(i, load_u(base + i * 4), *reinterpret_cast<const float4*>(x_in + base + i * 4));
}

// ------------------------------------------------------------------------------------------------
// Layout plumbing at block boundaries: the reference's tensors are fp32 NCHW, the kernels' are NHWC.
// Tiled 32x32 transposes through padded smem: coalesced on both sides.   [N][C][P] <-> [N][P][C]
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ yf,
                                    __nv_bfloat16* __restrict__ yb, int C, int P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xn = x + size_t(n) * C * P;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < P) ? xn[size_t(c) * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < P && c < C) {
      const float v = tile[threadIdx.x][i];
      const size_t o = (size_t(n) * P + p) * C + c;
      if (yf) yf[o] = v;
      if (yb) yb[o] = __float2bfloat16_rn(v);
    }
  }
}

template <bool X_BF16>
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x_, float* __restrict__ y, int C, int P) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < P && c < C) {
      const size_t o = (size_t(n) * P + p) * C + c;
      v = X_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(x_)[o]) : static_cast<const float*>(x_)[o];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < P) y[(size_t(n) * C + c) * P + p] = tile[threadIdx.x][i];
  }
}

// Per-image channel sums of an fp32 NHWC tensor in the pool_partial format ([N][partials][C]; slot 0 gets
// the sum, the remaining slots zero).  Only used by the stand-alone CALayer.forward; inside RCAB the sums
// come for free from conv2's epilogue.
__global__ void pool_sum_kernel(const float* __restrict__ x, float* __restrict__ pool_partial, int partials, int HW,
                                int C) {
  extern __shared__ float red[];  // [blockDim.x]
  const int n = blockIdx.x;
  const int lanes_per_c = blockDim.x / C;            // blockDim.x is a multiple of C
  const int c = threadIdx.x % C, lane = threadIdx.x / C;
  float s = 0.f;
  if (lane < lanes_per_c)
    for (int p = lane; p < HW; p += lanes_per_c) s += x[(size_t(n) * HW + p) * C + c];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int l = 0; l < lanes_per_c; ++l) t += red[l * C + threadIdx.x];
    float* pp = pool_partial + size_t(n) * partials * C;
    pp[threadIdx.x] = t;
    for (int i = 1; i < partials; ++i) pp[size_t(i) * C + threadIdx.x] = 0.f;
  }
}

}  // namespace rb
