// Backward of the whole RCAN / EDSR-baseline body in ONE persistent dataflow kernel (sm_100a).
//
// What autograd derives from RCAN.forward / ResidualGroup.forward / RCAB.forward / CALayer.forward
// (/root/reference/rumpy/SISR/models/advanced/architectures.py:41-44, 81-84, 121-124, 172-174; the backward
// obligations are listed in SURVEY.md 8a') for every layer between the body-tail conv and the head conv:
//   per group (last to first):   Q  = dgrad(group-tail conv)(GB)                               [kBwdFresh]
//     per RCAB (last to first):  s  = sum_hw Q*u;  du = Q*y + W1^T(relu'(.) W2^T(y(1-y) s))/HW   [kBwdCA, no conv]
//                                dt = dgrad(conv2)(du) * 1[t > 0]                                [kBwdMask]
//                                Q += dgrad(conv1)(dt)                                           [kBwdAcc]
//     P  = P + Q  (gradient w.r.t. the group input; bf16 copy = next group's GB)               [kBwdAcc, emit]
// Same machinery as trunk_pipe.cuh: tile-stationary CTAs (<= 4 tiles each), per-tile epochs in global memory,
// TMA halo-box loads, weights (the dgrad packing) streamed in kx thirds, the fp32 gradient stream Q of the owned
// tiles lives in TENSOR MEMORY for the whole kernel; the per-image reduction s is exchanged as (value, epoch)
// pairs.  du / dt are written once (bf16) -- they are also the operands of the batched wgrad kernel -- and their
// per-tile column sums (the conv bias gradients) come out of the same epilogues.
// Warp roles (352 threads): 0-3 / 4-7 epilogue groups (tile slots 0,2 / 1,3) | 8 A producer | 9 MMA | 10 weights.
#pragma once
#include "trunk_pipe.cuh"

namespace rb {

enum TrunkBwdKind : int { kBwdFresh = 3, kBwdCA = 4, kBwdMask = 5, kBwdAcc = 6 };

// One layer of the backward program (device table).
struct TrunkBwdLayer {
  int kind;
  int in_map, out_map;     // bf16 operand read by the conv (-1: no conv) / bf16 tensor written (-1: none)
  int w_idx;               // layer index inside the dgrad weight map (-1: no conv)
  int wait_epoch;          // the conv input is complete when the tile epochs reach this value (0: produced before the kernel)
  int ca_slot;             // kBwdCA: ordinal among the CA layers of this program (epoch of the s exchange)
  int pad0_, pad1_;
  const __nv_bfloat16* aux;     // kBwdCA: saved pre-attention activation u; kBwdMask: saved post-ReLU t (sign only)
  float* colsum;                // kBwdCA / kBwdMask: per-tile column sums of the emitted tensor, [T][64] (bias gradient)
  const float *w1, *w2;         // kBwdCA: FC weights [cr][64], [64][cr]
  const float *save_mean, *save_hid, *save_y;   // kBwdCA: forward CA vectors [N][64], [N][cr], [N][64]
  float* pg;                    // kBwdCA: per-image parameter-gradient terms [N][2*64*cr + 64 + cr]
  const float* q_scale;         // kBwdCA, Q-RCAN: forward meta-attention multipliers [N][64] (out = x + u*y*q), or nullptr
  float* dq;                    // kBwdCA, Q-RCAN: d(loss)/dq = s*y per (image, channel) [N][64], or nullptr
  const float* res_f32;         // kBwdAcc with emit: P (fp32 NHWC) read ...
  float* out_f32;               // ... and P + Q written (may alias res_f32); nullptr = no emit
  const float* add_f32;         // kBwdAcc with emit, HAN: the layer-attention gradient of the same tensor, added too; or nullptr
};

struct TrunkBwdArgs {
  const TrunkBwdLayer* layers;
  const CUtensorMap* in_maps;
  const CUtensorMap* out_maps;
  int* ready;                          // [T]
  unsigned long long* pool_partial;    // [2][T][64] (value, epoch) pairs
  int n_layers, N, H, W, tiles_x, tiles_y, tiles_per_img, T, K, cr;
  float inv_hw;
};

#ifdef RB_TRUNK_KERNEL_IMPL

__global__ void __launch_bounds__(kTrunkThreads, 1)
trunk_bwd_kernel(const __grid_constant__ CUtensorMap w_map, const TrunkBwdArgs args) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kTrunkAStages];
  __shared__ __align__(8) uint64_t a_empty[kTrunkAStages];
  __shared__ __align__(8) uint64_t w_full[3];
  __shared__ __align__(8) uint64_t w_empty[3];
  __shared__ __align__(8) uint64_t acc_full[kTrunkMaxK];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float y_s[2][64], coef_s[2][64];
  __shared__ float red_s[2][4][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpA = 8, kWarpMma = 9, kWarpW = 10, kWarpMma2 = 11;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + kTrunkWBytes;
  uint8_t* stg_s = a_s + kTrunkAStages * kAStageBytes;

  const int G = gridDim.x, cta = blockIdx.x;
  const int T = args.T, P = args.tiles_per_img, n_layers = args.n_layers;
  int my_k = 0;
  for (int j = 0; j < args.K; ++j) my_k += (cta + j * G < T) ? 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTrunkAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], my_k >= 2 ? 2 : 1); }
    for (int i = 0; i < kTrunkMaxK; ++i) mbar_init(&acc_full[i], 1);
    fence_mbar_init();
  }
  if (warp == kWarpW && lane == 0) tma_prefetch_desc(&w_map);
  if (warp == kWarpMma) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == kWarpA) {
    // ===================================================================== A-operand producer (conv layers only)
    // Two MMA issuers => two independent stage rings (stages {0,1} feed issuer 0, {2,3} issuer 1): a parity wait only
    // tells consecutive phases apart, so every ring must have exactly one consumer.  One tile per CTA: one ring of 4.
    const bool two = my_k >= 2;
    uint32_t cnt[2] = {0, 0};   // groups loaded into each ring
    int g = 0;                  // running tile counter: tile g belongs to issuer g & 1
    for (int L = 0; L < n_layers; ++L) {
      const TrunkBwdLayer* lay = args.layers + L;
      if (lane == 0 && L + 1 < n_layers) {
        const TrunkBwdLayer* nx = lay + 1;
        if (nx->in_map >= 0) tma_prefetch_desc(args.in_maps + nx->in_map);
        if (nx->out_map >= 0) tma_prefetch_desc(args.out_maps + nx->out_map);
      }
      if (lay->w_idx < 0) continue;
      const CUtensorMap* im = args.in_maps + lay->in_map;
      const int wait_epoch = lay->wait_epoch;
      for (int j = 0; j < my_k; ++j, ++g) {
        const int ring = two ? (g & 1) : 0;
        const int t = cta + j * G;
        const int n = t / P, rem = t - n * P;
        const int ty = rem / args.tiles_x, tx = rem - ty * args.tiles_x;
        if (wait_epoch > 0) {
          if (lane < 9) {
            const int nty = ty + lane / 3 - 1, ntx = tx + lane % 3 - 1;
            if (nty >= 0 && nty < args.tiles_y && ntx >= 0 && ntx < args.tiles_x)
              poll_ge(args.ready + n * P + nty * args.tiles_x + ntx, wait_epoch, 1, false);
          }
          __syncwarp();
        }
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t c = cnt[ring]++;
          const int stage = two ? 2 * ring + int(c & 1u) : int(c & 3u);
          const uint32_t phase = (two ? (c >> 1) : (c >> 2)) & 1u;
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&a_full[stage], kAStageBytes);
            tma_load_4d(a_s + stage * kAStageBytes, im, &a_full[stage], 0, tx * kTileW + kx - 1, ty * kTileH - 1, n);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma2) {
    // ===================================================================== MMA issuers (two warps, see trunk_pipe.cuh)
    const int me = warp == kWarpMma ? 0 : 1;
    const bool two = my_k >= 2;
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
    int cl = 0, g = 0;   // cl: ordinal of the conv layer (parity of the weight / accumulator barriers)
    uint32_t cnt = 0;    // groups consumed from this issuer's own stage ring (stages {2 me, 2 me + 1}; one ring of 4 when alone)
    if (two || me == 0) {
      for (int L = 0; L < n_layers; ++L) {
        if (args.layers[L].w_idx < 0) continue;
        int my_last = -1;
        for (int j = 0; j < my_k; ++j)
          if (!two || ((g + j) & 1) == me) my_last = j;
        bool have_w = false;
        for (int j = 0; j < my_k; ++j, ++g) {
          if (two && (g & 1) != me) continue;
          const uint32_t d_tmem = tmem_base + uint32_t(kTrunkAccCol + j * 64);
          for (int kx = 0; kx < 3; ++kx) {
            if (!have_w) mbar_wait_trap(&w_full[kx], uint32_t(cl & 1));
            const uint32_t c = cnt++;
            const int stage = two ? 2 * me + int(c & 1u) : int(c & 3u);
            mbar_wait_trap(&a_full[stage], (two ? (c >> 1) : (c >> 2)) & 1u);
            {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_addr = smem_u32(a_s + stage * kAStageBytes);
                const uint32_t b_addr = smem_u32(w_s + kx * kTrunkWThird);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                  const uint64_t adesc = make_smem_desc(a_addr + ky * (kTileW * 128), 16, 1024, kLayoutSw128);
                  const uint64_t bdesc = make_smem_desc(b_addr + ky * 8192, 16, 1024, kLayoutSw128);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0);
                }
                umma_commit(&a_empty[stage]);
                if (j == my_last) umma_commit(&w_empty[kx]);
              }
              __syncwarp();
            }
          }
          {
            have_w = true;
            if (elect_one()) umma_commit(&acc_full[j]);
            __syncwarp();
          }
        }
        ++cl;
      }
    }
  } else if (warp == kWarpW) {
    // ===================================================================== weight producer (three kx thirds)
    int cl = 0;
    for (int L = 0; L < n_layers; ++L) {
      const int w_idx = args.layers[L].w_idx;
      if (w_idx < 0) continue;
      for (int kx = 0; kx < 3; ++kx) {
        mbar_wait(&w_empty[kx], uint32_t(cl & 1) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&w_full[kx], kTrunkWThird);
          tma_load_4d(w_s + kx * kTrunkWThird, &w_map, &w_full[kx], 0, 0, kx * 3, w_idx);
        }
        __syncwarp();
      }
      ++cl;
    }
  } else if (warp < 8) {
    // ===================================================================== epilogue groups (2 x 128 threads)
    const int e = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;           // pixel of the tile == TMEM lane == thread index in the group
    const int ly = row >> 4, lx = row & 15;
    const uint32_t swz = uint32_t(row & 7);
    const uint32_t bar_id = 1u + uint32_t(e);
    uint8_t* stg = stg_s + e * kABytes;
    uint8_t* my_stg = stg + row * 128;
    unsigned long long* pool_scr = reinterpret_cast<unsigned long long*>(stg_s + 2 * kABytes + e * kTrunkPoolBytes);
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const int cr = args.cr;

    int cl = 0;
    for (int L = 0; L < n_layers; ++L) {
      const TrunkBwdLayer* lay = args.layers + L;
      const int kind = lay->kind;

      auto coords = [&](int j, int& t, int& n, int& rem, int& ty, int& tx, bool& valid, size_t& pix) {
        t = cta + j * G;
        n = t / P; rem = t - n * P;
        ty = rem / args.tiles_x; tx = rem - ty * args.tiles_x;
        const int y = ty * kTileH + ly, x = tx * kTileW + lx;
        valid = y < args.H && x < args.W;
        pix = ((size_t(n) * args.H + y) * args.W + x) * 64;
      };
      // staged bf16 tile -> global (TMA), then publish the tile's epoch (release: see DESIGN.md trunk protocol)
      auto finish_tile = [&](int t, int n, int ty, int tx) {
        tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (row == 0) {
          tma_store_4d(args.out_maps + lay->out_map, stg, 0, tx * kTileW, ty * kTileH, n);
          tma_store_commit();
          tma_store_wait_all0();
          st_release_s32(args.ready + t, L + 1);
        }
      };
      auto stage_bf16 = [&](const float (&f)[32], int h) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(my_stg + ((uint32_t(h * 4 + c) ^ swz) << 4)) =
              make_uint4(pack_bf16x2(f[c * 8], f[c * 8 + 1]), pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]),
                         pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]));
      };
      // per-tile column sums of the emitted tensor (f of both halves already reduced into red_s): one row of [T][64]
      auto colsum_row = [&](int t) {
        named_bar_sync(bar_id, 128);
        if (row < 64 && lay->colsum != nullptr)
          lay->colsum[size_t(t) * 64 + row] =
              (red_s[e][0][row] + red_s[e][1][row]) + (red_s[e][2][row] + red_s[e][3][row]);
      };

      if (kind == kBwdFresh) {
        // ------------------------------------------------------------ Q = dgrad(group tail)(GB): accumulator -> stream
        for (int j = e; j < my_k; j += 2) {
          mbar_wait(&acc_full[j], uint32_t(cl & 1));
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld32(lane_addr + uint32_t(kTrunkAccCol + j * 64 + h * 32), v);
            tmem_ld_wait();
            tmem_st32(lane_addr + uint32_t(j * 64 + h * 32), v);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      } else if (kind == kBwdAcc) {
        // ------------------------------------------------------------ Q += dgrad(conv1)(dt); group start: P += Q, emit
        const float* res = lay->res_f32;
        float* outf = lay->out_f32;
        for (int j = e; j < my_k; j += 2) {
          int t, n, rem, ty, tx; bool valid; size_t pix;
          coords(j, t, n, rem, ty, tx, valid, pix);
          mbar_wait(&acc_full[j], uint32_t(cl & 1));
          tc_fence_after();
          if (outf != nullptr) named_bar_sync(bar_id, 128);   // row 0 is past the previous store's wait: staging is free
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32], s[32];
            float f[32];
            tmem_ld32(lane_addr + uint32_t(kTrunkAccCol + j * 64 + h * 32), v);
            tmem_ld32(lane_addr + uint32_t(j * 64 + h * 32), s);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(s[i]) + __uint_as_float(v[i]);
            if (outf == nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) s[i] = __float_as_uint(f[i]);
              tmem_st32(lane_addr + uint32_t(j * 64 + h * 32), s);
            } else {
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) r = *reinterpret_cast<const float4*>(res + pix + h * 32 + c4 * 4);
                if (valid && lay->add_f32 != nullptr) {
                  const float4 a = *reinterpret_cast<const float4*>(lay->add_f32 + pix + h * 32 + c4 * 4);
                  r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
                }
                f[c4 * 4 + 0] += r.x; f[c4 * 4 + 1] += r.y; f[c4 * 4 + 2] += r.z; f[c4 * 4 + 3] += r.w;
                if (valid)
                  *reinterpret_cast<float4*>(outf + pix + h * 32 + c4 * 4) =
                      make_float4(f[c4 * 4], f[c4 * 4 + 1], f[c4 * 4 + 2], f[c4 * 4 + 3]);
              }
              stage_bf16(f, h);
            }
          }
          tmem_st_wait();
          if (outf != nullptr) finish_tile(t, n, ty, tx); else tc_fence_before();
        }
      } else if (kind == kBwdMask) {
        // ------------------------------------------------------------ dt = dgrad(conv2)(du) * 1[t > 0] (+ column sums)
        for (int j = e; j < my_k; j += 2) {
          int t, n, rem, ty, tx; bool valid; size_t pix;
          coords(j, t, n, rem, ty, tx, valid, pix);
          mbar_wait(&acc_full[j], uint32_t(cl & 1));
          tc_fence_after();
          named_bar_sync(bar_id, 128);   // staging + red_s are free (row 0 is past the previous store's wait)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            float f[32];
            tmem_ld32(lane_addr + uint32_t(kTrunkAccCol + j * 64 + h * 32), v);
            uint4 m[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
              m[c] = valid ? __ldg(reinterpret_cast<const uint4*>(lay->aux + pix + h * 32) + c) : make_uint4(0, 0, 0, 0);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t mw[4] = {m[c].x, m[c].y, m[c].z, m[c].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t lo = mw[k] & 0xFFFFu, hi = mw[k] >> 16;   // bf16 > 0 <=> sign clear and magnitude non-zero
                f[c * 8 + k * 2] = (lo != 0 && lo < 0x8000u) ? __uint_as_float(v[c * 8 + k * 2]) : 0.f;
                f[c * 8 + k * 2 + 1] = (hi != 0 && hi < 0x8000u) ? __uint_as_float(v[c * 8 + k * 2 + 1]) : 0.f;
              }
            }
            stage_bf16(f, h);
            red_s[e][q][h * 32 + lane] = lane_transpose_sum32(f, lane);
          }
          colsum_row(t);
          finish_tile(t, n, ty, tx);
        }
        } else if (kind == kBwdCA) {
        // ------------------------------------------------------------ CALayer backward, phase 1: s partials = sum Q*u
        const unsigned epoch = unsigned(lay->ca_slot + 1);
        unsigned long long* pbase = args.pool_partial + size_t(lay->ca_slot & 1) * T * 64;
        for (int j = e; j < my_k; j += 2) {
          int t, n, rem, ty, tx; bool valid; size_t pix;
          coords(j, t, n, rem, ty, tx, valid, pix);
          named_bar_sync(bar_id, 128);   // red_s is free
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t s[32];
            float f[32];
            tmem_ld32(lane_addr + uint32_t(j * 64 + h * 32), s);
            uint4 u4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
              u4[c] = valid ? __ldg(reinterpret_cast<const uint4*>(lay->aux + pix + h * 32) + c) : make_uint4(0, 0, 0, 0);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t uw[4] = {u4[c].x, u4[c].y, u4[c].z, u4[c].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                f[c * 8 + k * 2] = __uint_as_float(s[c * 8 + k * 2]) * __uint_as_float(uw[k] << 16);
                f[c * 8 + k * 2 + 1] = __uint_as_float(s[c * 8 + k * 2 + 1]) * __uint_as_float(uw[k] & 0xFFFF0000u);
              }
            }
            red_s[e][q][h * 32 + lane] = lane_transpose_sum32(f, lane);
          }
          named_bar_sync(bar_id, 128);
          if (row < 64) {
            const float sp = (red_s[e][0][row] + red_s[e][1][row]) + (red_s[e][2][row] + red_s[e][3][row]);
            st_relaxed_u64(pbase + size_t(t) * 64 + row,
                           (static_cast<unsigned long long>(epoch) << 32) | __float_as_uint(sp));
          }
        }
        // ------------------------------------------------------------ phase 2: FC backward per image, du = Q*y + coef
        for (int j = e; j < my_k; j += 2) {
          int t, n, rem, ty, tx; bool valid; size_t pix;
          coords(j, t, n, rem, ty, tx, valid, pix);
          const int c = row & 63, hsel = row >> 6;
          const unsigned long long* pp = pbase + size_t(n) * P * 64;
          // CA vectors / FC weights of this image into registers while the partials are in flight
          const float y0 = __ldg(lay->save_y + n * 64 + lane), y1 = __ldg(lay->save_y + n * 64 + 32 + lane);
          const float yc = __ldg(lay->save_y + n * 64 + c);
          const float* qs = lay->q_scale;   // Q-RCAN: out = x + u*y*q  =>  du = Q*y*q + coef, dy = s*q, dq = s*y
          const float q0 = qs ? __ldg(qs + n * 64 + lane) : 1.f, q1 = qs ? __ldg(qs + n * 64 + 32 + lane) : 1.f;
          const float qc = qs ? __ldg(qs + n * 64 + c) : 1.f;
          float w2a[4], w2b[4], w1c[4], hid[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const bool on = h < cr;
            w2a[h] = on ? __ldg(lay->w2 + lane * cr + h) : 0.f;
            w2b[h] = on ? __ldg(lay->w2 + (lane + 32) * cr + h) : 0.f;
            w1c[h] = on ? __ldg(lay->w1 + h * 64 + c) : 0.f;
            hid[h] = on ? __ldg(lay->save_hid + n * cr + h) : 0.f;
          }
          const float ssum = pool_column_sum(pp, P, epoch, pool_scr, row, bar_id);
          named_bar_sync(bar_id, 128);   // previous readers of red_s / y_s / coef_s are done; staging is free
          red_s[e][hsel][c] = ssum;
          named_bar_sync(bar_id, 128);
          // dz2 = s*y*(1-y) for channels lane, lane+32; dh_j = (W2^T dz2)_j * 1[hid_j > 0]; dmean_c = (W1^T dh)_c
          const float dz0 = (red_s[e][0][lane] + red_s[e][1][lane]) * q0 * y0 * (1.f - y0);
          const float dz1 = (red_s[e][0][lane + 32] + red_s[e][1][lane + 32]) * q1 * y1 * (1.f - y1);
          float dmean = 0.f;
          float dh[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float sdot = w2a[h] * dz0 + w2b[h] * dz1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
            dh[h] = hid[h] > 0.f ? sdot : 0.f;
            dmean = fmaf(w1c[h], dh[h], dmean);
          }
          const bool saver = rem == 0 && hsel == 0 && lay->pg != nullptr;   // one tile per image records the FC gradients
          float* mine = lay->pg + size_t(n) * (2 * 64 * cr + 64 + cr);
          const float my_dz = q == 0 ? dz0 : dz1;   // (hsel == 0: c = lane or 32 + lane)
          if (saver) {
            const float mean_c = __ldg(lay->save_mean + n * 64 + c);
#pragma unroll
            for (int h = 0; h < 4; ++h)
              if (h < cr) {
                mine[c * cr + h] = my_dz * hid[h];                 // dW2[c][h]
                mine[64 * cr + h * 64 + c] = dh[h] * mean_c;       // dW1[h][c]
                if (c == 0) mine[2 * 64 * cr + 64 + h] = dh[h];    // db1[h]
              }
            mine[2 * 64 * cr + c] = my_dz;                         // db2[c]
            if (lay->dq != nullptr) lay->dq[n * 64 + c] = (red_s[e][0][c] + red_s[e][1][c]) * yc;
          }
          for (int h = 4; h < cr; ++h) {   // generic tail (cr > 4)
            float sdot = __ldg(lay->w2 + lane * cr + h) * dz0 + __ldg(lay->w2 + (lane + 32) * cr + h) * dz1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
            const float hv = __ldg(lay->save_hid + n * cr + h);
            const float dhh = hv > 0.f ? sdot : 0.f;
            dmean = fmaf(__ldg(lay->w1 + h * 64 + c), dhh, dmean);
            if (saver) {
              mine[c * cr + h] = my_dz * hv;
              mine[64 * cr + h * 64 + c] = dhh * __ldg(lay->save_mean + n * 64 + c);
              if (c == 0) mine[2 * 64 * cr + 64 + h] = dhh;
            }
          }
          if (hsel == 0) { y_s[e][c] = yc * qc; coef_s[e][c] = dmean * args.inv_hw; }
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t s[32];
            float f[32];
            tmem_ld32(lane_addr + uint32_t(j * 64 + h * 32), s);
            float yy[32], cf[32];
            lds_bcast32(&y_s[e][h * 32], yy);
            lds_bcast32(&coef_s[e][h * 32], cf);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = valid ? fmaf(__uint_as_float(s[i]), yy[i], cf[i]) : 0.f;
            stage_bf16(f, h);
            red_s[e][q][h * 32 + lane] = lane_transpose_sum32(f, lane);
          }
          colsum_row(t);
          finish_tile(t, n, ty, tx);
        }
      }
      if (lay->w_idx >= 0) ++cl;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// d(loss)/d(CA FC parameters) = sum over images (fixed order) of the per-image terms written by trunk_bwd_kernel.
struct CaPgJob { const float* pg; float *dw1, *db1, *dw2, *db2; };
__global__ void ca_pg_finalize_kernel(const CaPgJob* __restrict__ jobs, int N, int cr) {
  const CaPgJob jb = jobs[blockIdx.x];
  const int per = 2 * 64 * cr + 64 + cr;
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    float s = 0.f;
    for (int img = 0; img < N; ++img) s += jb.pg[size_t(img) * per + i];
    if (i < 64 * cr) jb.dw2[i] = s;
    else if (i < 2 * 64 * cr) jb.dw1[i - 64 * cr] = s;
    else if (i < 2 * 64 * cr + 64) jb.db2[i - 2 * 64 * cr] = s;
    else jb.db1[i - 2 * 64 * cr - 64] = s;
  }
}

// d(loss)/d(q-layer parameters) of one block (reference q_layer.py:5-45, 2-layer ParaCALayer): q = sigmoid(a),
// a = W2 act(W1 meta + b1) + b2.  dq[n][c] = sum_hw g*u*y comes from trunk_bwd_kernel.  One CTA per block; the sums
// over the images run in a fixed order (deterministic).
struct QGradJob { const float *w1, *b1, *w2, *b2, *q, *dq; float *dw1, *db1, *dw2, *db2; };
// dq_slices == 0: jb.dq = dq [N][C] (Q-RCAN).  dq_slices > 0: jb.dq = [N][dq_slices][C] partial sums of dq * q
// (Q-EDSR: sum_hw g * (out - x) = q * dq), so da = dq * q * (1 - q) = (sum of the slices) * (1 - q).
__global__ void q_grad_kernel(const QGradJob* __restrict__ jobs, const float* __restrict__ meta, int N, int M,
                              int hidden, int C, int relu, int dq_slices) {
  extern __shared__ float qg_smem[];
  float* meta_s = qg_smem;                 // [N][M]
  float* hid_s = meta_s + N * M;           // [N][hidden]  act(W1 meta + b1)
  float* da_s = hid_s + N * hidden;        // [N][C]       dq * q * (1 - q)
  float* dh_s = da_s + N * C;              // [N][hidden]
  const QGradJob jb = jobs[blockIdx.x];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < N * M; i += nt) meta_s[i] = meta[i];
  for (int i = tid; i < N * C; i += nt) {
    const float q = jb.q[i];
    if (dq_slices == 0) {
      da_s[i] = jb.dq[i] * q * (1.f - q);
    } else {
      const int n = i / C, c = i - n * C;
      float s = 0.f;
      for (int k = 0; k < dq_slices; ++k) s += jb.dq[(size_t(n) * dq_slices + k) * C + c];
      da_s[i] = s * (1.f - q);
    }
  }
  __syncthreads();
  for (int i = tid; i < N * hidden; i += nt) {
    const int n = i / hidden, t = i - n * hidden;
    float a = jb.b1[t];
    for (int m = 0; m < M; ++m) a = fmaf(jb.w1[size_t(t) * M + m], meta_s[n * M + m], a);
    hid_s[i] = relu ? fmaxf(a, 0.f) : a;
  }
  __syncthreads();
  for (int i = tid; i < C * hidden; i += nt) {   // dW2[c][t]
    const int c = i / hidden, t = i - c * hidden;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(da_s[n * C + c], hid_s[n * hidden + t], s);
    jb.dw2[i] = s;
  }
  for (int c = tid; c < C; c += nt) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += da_s[n * C + c];
    jb.db2[c] = s;
  }
  for (int i = tid; i < N * hidden; i += nt) {
    const int n = i / hidden, t = i - n * hidden;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(jb.w2[size_t(c) * hidden + t], da_s[n * C + c], s);
    dh_s[i] = (relu && hid_s[i] <= 0.f) ? 0.f : s;
  }
  __syncthreads();
  for (int i = tid; i < hidden * M; i += nt) {   // dW1[t][m]
    const int t = i / M, m = i - t * M;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dh_s[n * hidden + t], meta_s[n * M + m], s);
    jb.dw1[i] = s;
  }
  for (int t = tid; t < hidden; t += nt) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dh_s[n * hidden + t];
    jb.db1[t] = s;
  }
}

// Q-EDSR: partial[n][slice][c] = sum over the slice's pixels of g[n,p,c] * (out[n,p,c] - x[n,p,c]); out - x = the
// block's scaled branch r * q, so the sum over slices is q * dq.  grid (slices, N), block 256 = (256 / C) pixel lanes x C.
__global__ void dq_reduce_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ out,
                                 const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, int HW, int C) {
  extern __shared__ float dq_red[];   // [lanes][C]
  const int n = blockIdx.y, lanes = blockDim.x / C;
  const int c = threadIdx.x % C, lane = threadIdx.x / C;
  const int begin = int((long long)HW * blockIdx.x / gridDim.x), end = int((long long)HW * (blockIdx.x + 1) / gridDim.x);
  float s = 0.f;
  for (int p = begin + lane; p < end; p += lanes) {
    const size_t i = (size_t(n) * HW + p) * C + c;
    s = fmaf(g[i], __bfloat162float(out[i]) - __bfloat162float(x[i]), s);
  }
  dq_red[lane * C + c] = s;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l) s += dq_red[l * C + c];
    partial[(size_t(n) * gridDim.x + blockIdx.x) * C + c] = s;
  }
}

#endif  // RB_TRUNK_KERNEL_IMPL

}  // namespace rb
