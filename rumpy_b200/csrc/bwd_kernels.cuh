// CUDA-core kernels of the backward pass / train step around the tensor-core dgrad + wgrad kernels:
// L1 loss + gradient, thin tail/head conv gradients, CALayer backward, bias gradients, gradient-stream
// adds, fused Adam.  Memory-bound streaming kernels; reductions write per-block partials that a second
// kernel sums in a fixed order (deterministic gradients).
// Reference semantics: SURVEY.md 8(a') table; base_architecture.py:40 (nn.L1Loss), :93-95 (Adam), :425-440.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace rb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// L1 loss (mean) + its gradient: dy = sign(out - y) / numel.   partial[b] = sum |d| of block b.
// ------------------------------------------------------------------------------------------------
__global__ void l1_loss_grad_kernel(const float* __restrict__ out, const float* __restrict__ y,
                                    float* __restrict__ dy, float* __restrict__ partial, size_t numel, float gscale) {
  __shared__ float red[32];
  float s = 0.f;
  const float inv = gscale / float(numel);
  for (size_t i = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) * 4; i < numel;
       i += size_t(gridDim.x) * blockDim.x * 4) {
    if (i + 3 < numel) {
      const float4 a = *reinterpret_cast<const float4*>(out + i), b = *reinterpret_cast<const float4*>(y + i);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      s += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
      if (dy) {
        float4 g;
        g.x = d0 > 0.f ? inv : (d0 < 0.f ? -inv : 0.f); g.y = d1 > 0.f ? inv : (d1 < 0.f ? -inv : 0.f);
        g.z = d2 > 0.f ? inv : (d2 < 0.f ? -inv : 0.f); g.w = d3 > 0.f ? inv : (d3 < 0.f ? -inv : 0.f);
        *reinterpret_cast<float4*>(dy + i) = g;
      }
    } else {
      for (size_t k = i; k < numel; ++k) {
        const float d = out[k] - y[k];
        s += fabsf(d);
        if (dy) dy[k] = d > 0.f ? inv : (d < 0.f ? -inv : 0.f);
      }
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
  }
}

__global__ void l1_loss_finalize_kernel(const float* __restrict__ partial, int n, float* __restrict__ loss,
                                        size_t numel) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += double(partial[i]);
    *loss = float(s / double(numel));
  }
}

// ------------------------------------------------------------------------------------------------
// Thin tail conv backward (C -> out_feats <= 4), upstream gradient dy in the reference's fp32 NCHW.
//   dgrad:  dX[p,ci] = sum_{co,ky,kx} dy[co, p - (ky-1,kx-1)] * W[co][ci][ky][kx]     -> bf16 NHWC
// One thread = one pixel x 8 input channels; weights transposed in smem [co*9+tap][C].
// ------------------------------------------------------------------------------------------------
__global__ void tail_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                  __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int cout) {
  // One thread = FOUR consecutive pixels of a row x 8 input channels: every weight vector fetched from shared
  // memory feeds 32 FMAs (the one-pixel version was bound by the shared-memory reads of the weights), and the
  // 3 x 6 window of dy per output channel is loaded once for the four pixels.
  extern __shared__ float wsm[];  // [cout*9][C]
  for (int i = threadIdx.x; i < cout * 9 * C; i += blockDim.x) {
    const int ci = i % C, k = i / C;       // k = co*9 + tap
    const int co = k / 9, tap = k % 9;
    wsm[i] = w[(size_t(co) * C + ci) * 9 + tap];
  }
  __syncthreads();
  const int groups = C / 8;
  const int quads = (W + 3) / 4;
  const size_t items = size_t(N) * H * quads * groups;
  for (size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x; t < items; t += size_t(gridDim.x) * blockDim.x) {
    const int g = int(t % groups);
    const size_t qi = t / groups;
    const int x0 = int(qi % quads) * 4, yh = int((qi / quads) % H), n = int(qi / (size_t(quads) * H));
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
    for (int co = 0; co < cout; ++co) {
      const float* gp = dy + (size_t(n) * cout + co) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = yh - (ky - 1);
        if (yy < 0 || yy >= H) continue;
        float win[6];                       // dy[yy][x0 - 1 .. x0 + 4]
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int xx = x0 - 1 + i;
          win[i] = (xx >= 0 && xx < W) ? __ldg(gp + size_t(yy) * W + xx) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(wsm + (co * 9 + ky * 3 + kx) * C + g * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wsm + (co * 9 + ky * 3 + kx) * C + g * 8 + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = win[p + 2 - kx];   // dy at x0 + p - (kx - 1)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(v, wv[j], acc[p][j]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (x0 + p >= W) break;
      __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[p][0], acc[p][1]), b1 = __floats2bfloat162_rn(acc[p][2], acc[p][3]);
      __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[p][4], acc[p][5]), b3 = __floats2bfloat162_rn(acc[p][6], acc[p][7]);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
      o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
      const size_t pix = (size_t(n) * H + yh) * W + x0 + p;
      *reinterpret_cast<uint4*>(dx + pix * C + g * 8) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Upstream gradient of the thin tail conv as a tensor-core operand: dy fp32 NCHW [N][M][H][W] (M <= 4) -> bf16 NHWC
// with 64 channels (channels >= M zero), so the tail conv's weight gradient rides in the batched tcgen05 wgrad kernel
// as one more 64x64 block instead of a CUDA-core pass (was 0.67 ms of the 15.8 ms RCAN train step).
// One thread = one pixel x one 16-byte chunk (8 channels); a warp writes 4 whole pixels (512 contiguous bytes).
// ------------------------------------------------------------------------------------------------
__global__ void pad_thin_grad_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ out, int N, int M,
                                     int HW) {
  const size_t total = size_t(N) * HW * 8;
  for (size_t t = blockIdx.x * size_t(blockDim.x) + threadIdx.x; t < total; t += size_t(gridDim.x) * blockDim.x) {
    const int chunk = int(t & 7);
    const size_t pix = t >> 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (chunk == 0) {
      const size_t n = pix / HW, p = pix - n * HW;
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      for (int m = 0; m < M; ++m) f[m] = dy[(n * M + m) * HW + p];
      const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
      v.x = *reinterpret_cast<const uint32_t*>(&a);
      v.y = *reinterpret_cast<const uint32_t*>(&b);
    }
    reinterpret_cast<uint4*>(out)[t] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Thin conv weight gradient, shared by the tail conv (C -> few) and the head conv (few -> C):
//   S[m][tap][c] = sum_p thin[m, p - or + off(tap)] * wide[p, c]      m < M <= 4, c < C
//   tail: thin = dy (NCHW), wide = X bf16 NHWC, dW[m][c][ky][kx] = sum_p dy[m,p] X[p+off,c]
//         -> iterate wide pixel q = p+off: dy index q - off            (sign = -1)
//   head: thin = x (NCHW input), wide = G fp32 NHWC, dW[c][m][ky][kx] = sum_p G[p,c] x[m,p+off]   (sign = +1)
// plus SB[m] = sum_p thin[m,p] (tail bias grad) or SB'[c] = sum_p wide[p,c] (head bias grad).
// Thread = (channel c, pixel lane); each block writes a partial [M*9 + 1][C] (last row: wide column sums).
// ------------------------------------------------------------------------------------------------
// MM = compile-time bound on M (3 for RGB: the accumulators and the window stay in ~130 registers, two blocks per
// SM instead of one at MM = 4).
template <bool WIDE_BF16, int MM = 4>
__global__ void thin_wgrad_kernel(const float* __restrict__ thin, const void* __restrict__ wide_,
                                  const float* __restrict__ wide2 /* optional second fp32 wide tensor, added */,
                                  float* __restrict__ partial, int N, int H, int W, int C, int M, int sign) {
  // One (lane, row) walks along x with a sliding 3-column window of the thin tensor in registers: per pixel one
  // coalesced wide load, 3*M broadcast thin loads (instead of 9*M) and 9*M FMAs; no div/mod in the loop.
  extern __shared__ float red[];  // [lanes][M*9+1][C] reduce buffer
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, lane = threadIdx.x / C;
  float acc[MM * 9 + 1];
#pragma unroll
  for (int i = 0; i < MM * 9 + 1; ++i) acc[i] = 0.f;
  const int rows_total = N * H;
  for (int rowi = blockIdx.x * lanes + lane; rowi < rows_total && lane < lanes; rowi += gridDim.x * lanes) {
    const int n = rowi / H, yh = rowi - n * H;
    const float* rp[MM][3];   // thin rows for (m, ky): row yh + sign*(ky-1), nullptr when outside the image
#pragma unroll
    for (int m = 0; m < MM; ++m)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = yh + sign * (ky - 1);
        rp[m][ky] = (m < M && yy >= 0 && yy < H) ? thin + ((size_t(n) * M + m) * H + yy) * W : nullptr;
      }
    float win[MM][3][3];      // [m][ky][column offset -1, 0, +1]
#pragma unroll
    for (int m = 0; m < MM; ++m)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        win[m][ky][0] = 0.f;                                           // column -1 is outside
        win[m][ky][1] = rp[m][ky] ? __ldg(rp[m][ky]) : 0.f;            // column 0
        win[m][ky][2] = (rp[m][ky] && W > 1) ? __ldg(rp[m][ky] + 1) : 0.f;
      }
    const size_t wbase = (size_t(n) * H + yh) * W;
    for (int xw = 0; xw < W; ++xw) {
      float v;
      if (WIDE_BF16) v = __bfloat162float(static_cast<const __nv_bfloat16*>(wide_)[(wbase + xw) * C + c]);
      else {
        v = static_cast<const float*>(wide_)[(wbase + xw) * C + c];
        if (wide2) v += wide2[(wbase + xw) * C + c];
      }
      acc[MM * 9] += v;
#pragma unroll
      for (int m = 0; m < MM; ++m)
        if (m < M) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {  // thin column xw + sign*(kx-1)  ->  window slot 1 + sign*(kx-1)
              const float tv = (sign > 0) ? win[m][ky][kx] : win[m][ky][2 - kx];
              acc[m * 9 + ky * 3 + kx] = fmaf(tv, v, acc[m * 9 + ky * 3 + kx]);
            }
        }
      // slide the window one column to the right
      const bool more = xw + 2 < W;
#pragma unroll
      for (int m = 0; m < MM; ++m)
        if (m < M) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            win[m][ky][0] = win[m][ky][1];
            win[m][ky][1] = win[m][ky][2];
            win[m][ky][2] = (more && rp[m][ky]) ? __ldg(rp[m][ky] + xw + 2) : 0.f;
          }
        }
    }
  }
  const int rows = M * 9 + 1;
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < MM * 9; ++i)
      if (i < M * 9) red[(lane * rows + i) * C + c] = acc[i];
    red[(lane * rows + M * 9) * C + c] = acc[MM * 9];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * C; i += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * rows * C + i];
    partial[size_t(blockIdx.x) * rows * C + i] = s;
  }
}

// partial [blocks][M*9+1][C] -> tail: dW[m][c][tap] (+ thin-sum bias comes from thin_sum_kernel)
//                               head: dW[c][m][tap], db[c] = column sums (row M*9)
// 32 outputs x 8 slices per CTA of 256 threads: slice s adds partials s, s+8, ... (coalesced over the 32 outputs), the
// eight slice sums are combined in a fixed order (deterministic).
__global__ void thin_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ dw,
                                         float* __restrict__ db_wide, int C, int M, int is_head) {
  const int rows = M * 9 + 1;
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < rows * C)
    for (int b = slice; b < blocks; b += 8) s += partial[size_t(b) * rows * C + i];
  red[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || i >= rows * C) return;
  s = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) +
      ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
  const int c = i % C, row = i / C;
  if (row == M * 9) { if (db_wide) db_wide[c] = s; return; }
  const int m = row / 9, tap = row % 9;
  if (is_head) dw[(size_t(c) * M + m) * 9 + tap] = s;
  else dw[(size_t(m) * C + c) * 9 + tap] = s;
}

// per-plane sums of an NCHW tensor: out[m] = sum_{n,p} t[n,m,p]   (tail bias gradient; M <= 4 planes)
// grid (slices, M): partial[m][slice]; plane_sum_finalize_kernel adds the slices in order.
__global__ void plane_sum_kernel(const float* __restrict__ t, float* __restrict__ partial, int N, int M, int P) {
  __shared__ float red[32];
  const int m = blockIdx.y, slices = gridDim.x;
  const size_t total = size_t(N) * P;
  const size_t begin = total * blockIdx.x / slices, end = total * (blockIdx.x + 1) / slices;
  float s = 0.f;
  for (size_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const size_t n = i / P, p = i - n * P;
    s += t[(n * M + m) * P + p];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[m * slices + blockIdx.x] = v;
  }
}
__global__ void plane_sum_finalize_kernel(const float* __restrict__ partial, int slices, float* __restrict__ out,
                                          int M) {
  if (threadIdx.x < M) {
    float s = 0.f;
    for (int i = 0; i < slices; ++i) s += partial[threadIdx.x * slices + i];
    out[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Bias gradients of the tensor-core convs: phase-aware column sums of a bf16 NHWC gradient operand.
// g viewed as [outer = N*H][r (i)][inner = W][vec = r*C (j,c)]; db[c*r*r + i*r + j] = alpha * sum.
// Job list -> one launch for every layer; per-(job, slice) partials, fixed-order reduce.
// ------------------------------------------------------------------------------------------------
struct ColsumJob {
  const __nv_bfloat16* g;
  float* db;
  float* partial;   // [slices][r][r*C]
  int outer, r, inner, C;
  float alpha;
};
constexpr int kColsumSlices = 128;   // per-job parallelism: the HR upsampler gradients are 134 MB / 33 MB tensors

__global__ void colsum_kernel(const ColsumJob* __restrict__ jobs) {
  extern __shared__ float red[];  // [blockDim.x]
  const ColsumJob jb = jobs[blockIdx.y];
  const int vec = jb.r * jb.C;          // <= 768; blockDim.x = 768 covers every case (3 * 256)
  const int slice = blockIdx.x;
  const int o_begin = int((long long)jb.outer * slice / kColsumSlices);
  const int o_end = int((long long)jb.outer * (slice + 1) / kColsumSlices);
  const int lanes = blockDim.x / vec;   // row lanes per element
  const int e = threadIdx.x % vec, lane = threadIdx.x / vec;
  for (int i = 0; i < jb.r; ++i) {
    float s0 = 0.f, s1 = 0.f;
    if (lane < lanes) {
      for (int o = o_begin + lane; o < o_end; o += lanes) {
        const __nv_bfloat16* row = jb.g + ((size_t(o) * jb.r + i) * jb.inner) * vec + e;
        int w = 0;
        for (; w + 7 < jb.inner; w += 8) {   // eight independent 2-byte loads in flight per thread
          const float v0 = __bfloat162float(row[size_t(w) * vec]), v1 = __bfloat162float(row[size_t(w + 1) * vec]);
          const float v2 = __bfloat162float(row[size_t(w + 2) * vec]), v3 = __bfloat162float(row[size_t(w + 3) * vec]);
          const float v4 = __bfloat162float(row[size_t(w + 4) * vec]), v5 = __bfloat162float(row[size_t(w + 5) * vec]);
          const float v6 = __bfloat162float(row[size_t(w + 6) * vec]), v7 = __bfloat162float(row[size_t(w + 7) * vec]);
          s0 += (v0 + v2) + (v4 + v6);
          s1 += (v1 + v3) + (v5 + v7);
        }
        for (; w < jb.inner; ++w) s0 += __bfloat162float(row[size_t(w) * vec]);
      }
    }
    red[threadIdx.x] = s0 + s1;
    __syncthreads();
    if (threadIdx.x < vec) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[l * vec + threadIdx.x];
      jb.partial[(size_t(slice) * jb.r + i) * vec + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

__global__ void colsum_reduce_kernel(const ColsumJob* __restrict__ jobs) {
  const ColsumJob jb = jobs[blockIdx.x];
  const int vec = jb.r * jb.C, total = jb.r * vec;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < kColsumSlices; ++k) s += jb.partial[size_t(k) * total + idx];
    const int i = idx / vec, e = idx % vec, j = e / jb.C, c = e % jb.C;
    jb.db[c * jb.r * jb.r + i * jb.r + j] = s * jb.alpha;
  }
}

// ------------------------------------------------------------------------------------------------
// CALayer backward (SURVEY 8a'):  out = u*y, y = sigmoid(W2 relu(W1 mean(u)+b1)+b2)
//   s[n,c]  = sum_hw G*u            (ca_bwd_reduce_kernel -> per-(image, chunk) partials)
//   dz2 = s*y*(1-y); dh = (W2^T dz2) * 1[hid>0]; dmean = W1^T dh
//   du = G*y + dmean/HW  -> bf16 operand of conv2's dgrad / wgrad      (ca_bwd_apply_kernel)
//   dW2 += dz2 (x) hid; db2 += dz2; dW1 += dh (x) mean; db1 += dh       (block (0,0), images in fixed order)
// ------------------------------------------------------------------------------------------------
// Scratch/argument block of the CALayer backward pair (one per plan; counters self-reset every launch).
struct CaBwdArgs {
  const float* G; const void* u;               // upstream gradient (fp32 NHWC), saved conv2 output
  const float *save_mean, *save_hid, *save_y;  // forward CA vectors [N][C] / [N][Cr]
  const float *w1, *w2;                        // FC weights [Cr][C], [C][Cr]
  float *dw1, *db1, *dw2, *db2;                // parameter gradients (overwritten)
  float* s_partial;                            // [N][chunks][C]
  float* coef;                                 // [N][C]: dmean/HW, consumed by ca_bwd_apply_kernel
  float* pg_scratch;                           // [N][2*C*Cr + C + Cr] per-image parameter-gradient terms
  int* counters;                               // [N + 1] zero-initialised
  int N, HW, C, Cr;
  const float* q_scale;                        // Q-RCAN: forward multipliers [N][C] (out = x + u*y*q), or nullptr
  float* dq;                                   // Q-RCAN: d(loss)/dq = s*y [N][C], or nullptr
};

template <bool U_F32>
__global__ void ca_bwd_reduce_kernel(const CaBwdArgs a) {
  extern __shared__ float red[];   // [lanes][C]
  __shared__ float dz2_s[256], dh_s[64];
  __shared__ int last_s;
  const int C = a.C, HW = a.HW, Cr = a.Cr;
  const int n = blockIdx.y, chunks = gridDim.x, tid = threadIdx.x;
  const int vpp = C / 4;                      // float4 vectors per pixel
  const int lanes = blockDim.x / vpp, c4 = (tid % vpp) * 4, lane = tid / vpp;
  const int p_begin = int((long long)HW * blockIdx.x / chunks), p_end = int((long long)HW * (blockIdx.x + 1) / chunks);
  const float* G = a.G;
  auto ldu = [&](size_t o) -> float4 {
    if (U_F32) return *reinterpret_cast<const float4*>(static_cast<const float*>(a.u) + o);
    const uint2 raw = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(a.u) + o);
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
  };
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  int p = p_begin + lane;
  for (; p + lanes < p_end; p += 2 * lanes) {
    const size_t o0 = (size_t(n) * HW + p) * C + c4, o1 = o0 + size_t(lanes) * C;
    const float4 g0 = *reinterpret_cast<const float4*>(G + o0), g1 = *reinterpret_cast<const float4*>(G + o1);
    const float4 u0 = ldu(o0), u1 = ldu(o1);
    a0.x = fmaf(g0.x, u0.x, a0.x); a0.y = fmaf(g0.y, u0.y, a0.y); a0.z = fmaf(g0.z, u0.z, a0.z); a0.w = fmaf(g0.w, u0.w, a0.w);
    a1.x = fmaf(g1.x, u1.x, a1.x); a1.y = fmaf(g1.y, u1.y, a1.y); a1.z = fmaf(g1.z, u1.z, a1.z); a1.w = fmaf(g1.w, u1.w, a1.w);
  }
  if (p < p_end) {
    const size_t o0 = (size_t(n) * HW + p) * C + c4;
    const float4 g0 = *reinterpret_cast<const float4*>(G + o0), u0 = ldu(o0);
    a0.x = fmaf(g0.x, u0.x, a0.x); a0.y = fmaf(g0.y, u0.y, a0.y); a0.z = fmaf(g0.z, u0.z, a0.z); a0.w = fmaf(g0.w, u0.w, a0.w);
  }
  float* r = red + lane * C + c4;
  r[0] = a0.x + a1.x; r[1] = a0.y + a1.y; r[2] = a0.z + a1.z; r[3] = a0.w + a1.w;
  __syncthreads();
  if (tid < C) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * C + tid];
    a.s_partial[(size_t(n) * chunks + blockIdx.x) * C + tid] = s;
  }
  // ---- the last chunk-block of image n finishes the tiny FC backward once per image (not once per CTA of
  // the apply kernel): s -> dz2 -> dh -> dmean/HW, plus this image's parameter-gradient terms
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = (atomicAdd(a.counters + n, 1) == chunks - 1);
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  {
    // s[c] over the chunk partials: thread groups take interleaved chunks, 4 independent loads in flight each
    // (a serial 32-deep chain of L2 round trips cost ~5 us here), combined in a fixed order
    const int groups = blockDim.x / C, g = tid / C, c = tid % C;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (g < groups) {
      const float* pp = a.s_partial + size_t(n) * chunks * C + c;
      int k = g;
      for (; k + 3 * groups < chunks; k += 4 * groups) {
        s0 += pp[size_t(k) * C]; s1 += pp[size_t(k + groups) * C];
        s2 += pp[size_t(k + 2 * groups) * C]; s3 += pp[size_t(k + 3 * groups) * C];
      }
      for (; k < chunks; k += groups) s0 += pp[size_t(k) * C];
    }
    __syncthreads();                 // `red` is being reused
    red[tid] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < C) {
      float s = 0.f;
      for (int k = 0; k < groups; ++k) s += red[k * C + tid];
      const float y = a.save_y[n * C + tid];
      const float qv = a.q_scale != nullptr ? a.q_scale[n * C + tid] : 1.f;   // dy = s*q, dq = s*y
      dz2_s[tid] = s * qv * y * (1.f - y);
      if (a.dq != nullptr) a.dq[n * C + tid] = s * y;
    }
  }
  __syncthreads();
  {
    const int warp = tid >> 5, wl = tid & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < Cr; j += nwarps) {
      float v = 0.f;
      for (int c = wl; c < C; c += 32) v = fmaf(a.w2[c * Cr + j], dz2_s[c], v);
      v = warp_sum(v);
      if (wl == 0) dh_s[j] = a.save_hid[n * Cr + j] > 0.f ? v : 0.f;
    }
  }
  __syncthreads();
  const int per = 2 * C * Cr + C + Cr;
  float* mine = a.pg_scratch + size_t(n) * per;
  if (tid < C) {
    float v = 0.f;
    for (int j = 0; j < Cr; ++j) v = fmaf(a.w1[j * C + tid], dh_s[j], v);
    a.coef[n * C + tid] = v / float(HW);
    const float dz = dz2_s[tid], mean = a.save_mean[n * C + tid];
    for (int j = 0; j < Cr; ++j) {
      mine[tid * Cr + j] = dz * a.save_hid[n * Cr + j];            // dW2[c][j]
      mine[C * Cr + j * C + tid] = dh_s[j] * mean;                 // dW1[j][c]
    }
    mine[2 * C * Cr + tid] = dz;                                   // db2[c]
  }
  if (tid < Cr) mine[2 * C * Cr + C + tid] = dh_s[tid];            // db1[j]
  a.counters[n] = 0;                                               // re-arm (all chunk blocks of n have arrived)
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = (atomicAdd(a.counters + a.N, 1) == a.N - 1);
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  for (int i = tid; i < per; i += blockDim.x) {      // parameter gradients: images summed in index order
    float s = 0.f;
    for (int img = 0; img < a.N; ++img) s += a.pg_scratch[size_t(img) * per + i];
    if (i < C * Cr) a.dw2[i] = s;
    else if (i < 2 * C * Cr) a.dw1[i - C * Cr] = s;
    else if (i < 2 * C * Cr + C) a.db2[i - 2 * C * Cr] = s;
    else a.db1[i - 2 * C * Cr - C] = s;
  }
  if (tid == 0) a.counters[a.N] = 0;
}

// du = G*y + coef  -> bf16 operand of conv2's dgrad / wgrad; pure streaming (y / coef come from the reduce
// kernel's per-image epilogue).  Also emits per-block column sums of du (conv2's bias gradient).
__global__ void ca_bwd_apply_kernel(const float* __restrict__ G, const float* __restrict__ save_y,
                                    const float* __restrict__ coef, __nv_bfloat16* __restrict__ du,
                                    float* __restrict__ du_colsum /* [N][gridDim.x][C] */, int HW, int C,
                                    const float* __restrict__ q_scale /* Q-RCAN: du = G*y*q + coef; or nullptr */) {
  const int tid = threadIdx.x, n = blockIdx.y;
  const int vec_per_pix = C / 4;
  const size_t total = size_t(HW) * vec_per_pix;
  const size_t base = size_t(n) * HW * C;
  const size_t stride = size_t(gridDim.x) * blockDim.x;     // multiple of C/4: each thread keeps its 4 channels
  const size_t i0 = blockIdx.x * size_t(blockDim.x) + tid;
  const int c4 = int(i0 % vec_per_pix) * 4;
  float4 y4 = *reinterpret_cast<const float4*>(save_y + n * C + c4);
  if (q_scale != nullptr) {
    const float4 q4 = *reinterpret_cast<const float4*>(q_scale + n * C + c4);
    y4.x *= q4.x; y4.y *= q4.y; y4.z *= q4.z; y4.w *= q4.w;
  }
  const float4 k4 = *reinterpret_cast<const float4*>(coef + n * C + c4);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  auto emit = [&](size_t i, const float4& g) {
    const float a = fmaf(g.x, y4.x, k4.x), b = fmaf(g.y, y4.y, k4.y);
    const float c = fmaf(g.z, y4.z, k4.z), d = fmaf(g.w, y4.w, k4.w);
    cs.x += a; cs.y += b; cs.z += c; cs.w += d;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(du + base + i * 4) = pk;
  };
  size_t i = i0;
  for (; i + 3 * stride < total; i += 4 * stride) {        // four independent 16-byte loads in flight
    const float4 g0 = *reinterpret_cast<const float4*>(G + base + i * 4);
    const float4 g1 = *reinterpret_cast<const float4*>(G + base + (i + stride) * 4);
    const float4 g2 = *reinterpret_cast<const float4*>(G + base + (i + 2 * stride) * 4);
    const float4 g3 = *reinterpret_cast<const float4*>(G + base + (i + 3 * stride) * 4);
    emit(i, g0); emit(i + stride, g1); emit(i + 2 * stride, g2); emit(i + 3 * stride, g3);
  }
  for (; i < total; i += stride) emit(i, *reinterpret_cast<const float4*>(G + base + i * 4));
  if (du_colsum != nullptr) {
    __shared__ float4 cs_s[256];
    cs_s[tid] = cs;
    __syncthreads();
    if (tid < vec_per_pix) {      // threads tid, tid + vpp, tid + 2 vpp ... share the channel quad
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = tid; k < int(blockDim.x); k += vec_per_pix) {
        t.x += cs_s[k].x; t.y += cs_s[k].y; t.z += cs_s[k].z; t.w += cs_s[k].w;
      }
      *reinterpret_cast<float4*>(du_colsum + (size_t(n) * gridDim.x + blockIdx.x) * C + tid * 4) = t;
    }
  }
}

// db[c] = alpha * sum_i partial[i][c]  -- bias gradients from per-tile (conv epilogue POOL) / per-block partial rows
struct PartialSumJob {
  const float* partial; float* db; int count, C; float alpha;
};
__global__ void partial_sum_kernel(const PartialSumJob* __restrict__ jobs) {
  extern __shared__ float red[];
  const PartialSumJob jb = jobs[blockIdx.x];
  const int lanes = blockDim.x / jb.C, c = threadIdx.x % jb.C, lane = threadIdx.x / jb.C;
  float s0 = 0.f, s1 = 0.f;
  if (lane < lanes) {
    int i = lane;
    for (; i + lanes < jb.count; i += 2 * lanes) {
      s0 += jb.partial[size_t(i) * jb.C + c];
      s1 += jb.partial[size_t(i + lanes) * jb.C + c];
    }
    if (i < jb.count) s0 += jb.partial[size_t(i) * jb.C + c];
  }
  red[threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.x < jb.C) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * jb.C + threadIdx.x];
    jb.db[threadIdx.x] = s * jb.alpha;
  }
}

// dst_f = a + b (fp32 NHWC), dst_b = bf16(dst_f): joins of the fp32 gradient stream at group boundaries
__global__ void add_f32_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst_f,
                                    __nv_bfloat16* __restrict__ dst_b, size_t n4) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float4 o = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    if (dst_f) reinterpret_cast<float4*>(dst_f)[i] = o;
    if (dst_b) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
      reinterpret_cast<uint2*>(dst_b)[i] = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused Adam over flat fp32 buffers (torch.optim.Adam defaults: no weight decay, no amsgrad;
// base_architecture.py:93-95).  grad_scale folds gradient clipping (clip_coef) / DDP averaging.
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float beta1, float beta2, float eps,
                            float bc1, float bc2_sqrt, const float* __restrict__ grad_scale_dev, float grad_scale) {
  const float gs = grad_scale_dev ? grad_scale * (*grad_scale_dev) : grad_scale;
  const float step = lr / bc1;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gs;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// sum of squares of a flat buffer -> partial[b]; finalize writes clip_coef = min(1, max_norm/(norm+1e-6))
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
  }
}
__global__ void clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += double(partial[i]);
    const float norm = float(sqrt(s));
    out[0] = fminf(1.f, max_norm / (norm + 1e-6f));
    out[1] = norm;
  }
}

}  // namespace rb
