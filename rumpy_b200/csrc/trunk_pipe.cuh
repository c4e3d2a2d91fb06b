// The whole 64-channel residual trunk in ONE persistent, tile-stationary dataflow kernel (sm_100a).
//
// Replaces, for a whole RCAN body (or an EDSR-baseline body): every `default_conv` 3x3 64->64 call
// (/root/reference/rumpy/SISR/models/advanced/common.py:6-9), RCAB.forward (architectures.py:81-84) with its
// CALayer (:41-44), ResidualGroup.forward (:121-124), the body-tail conv + global skip of RCAN.forward
// (:172-174), and ResBlock.forward (common.py:71-75).
//
// Why: at the benchmark shapes (16 x 48x48 / 16 x 64x64 patches) one layer is only 2-4 output tiles per SM, so
// a kernel per layer is bounded by launch, prologue and drain latency, not by the tensor cores.  Here the layer
// loop runs INSIDE the kernel:
//   * each CTA (one per SM, all co-resident) owns up to 4 fixed 8x16-pixel tiles, taken from different images;
//   * layer L+1 of a tile starts as soon as the tile and its 8 spatial neighbours have published layer L
//     (per-tile epoch flags in global memory, release/acquire) -- no grid-wide barrier anywhere;
//   * the fp32 residual stream of the owned tiles never leaves the SM: it lives in TENSOR MEMORY (64 columns
//     per tile) next to the accumulators (64 columns per tile) -> 4 x 128 = all 512 TMEM columns;
//   * bf16 operand tiles travel through L2 only: TMA store -> flag -> neighbour's TMA halo-box load;
//   * the 72 KB of weights per layer stream through one smem buffer in three `kx` thirds that are refilled by
//     a dedicated warp as soon as the last tile's MMAs of the previous layer have consumed them;
//   * channel attention: per-tile channel sums (warp-shuffle transpose reduction, straight from the TMEM
//     accumulator) -> per-image arrival counter -> a dedicated warp computes y = sigmoid(W2 relu(W1 mean+b1)+b2)
//     -> the SAME accumulator is read again and x + u*y is written to the TMEM stream and the bf16 operand.
//
// Warp roles (384 threads):  0-3 epilogue group 0 (tile slots 0,2) | 4-7 epilogue group 1 (slots 1,3) | 8 A-operand
// TMA producer (polls the neighbour flags) | 9 MMA issuer | 10 weight producer | 11 channel-attention warp.
// The service warps carry the highest warp ids: the SM sub-partition arbiter favours them over the epilogue warps.
#pragma once
#include "conv3x3_tc.cuh"

namespace rb {

enum TrunkKind : int { kTrunkRelu = 0, kTrunkCA = 1, kTrunkRes = 2 };

struct TrunkLayer {
  int kind;
  int in_map, out_map;   // indices into TrunkArgs::in_maps (box 10 rows) / out_maps (box 8 rows)
  int ca_slot;           // kTrunkCA: ordinal among the CA layers (epoch of the per-image pool counter)
  int update_s;          // kTrunkRes: the result replaces the fp32 residual stream in TMEM
  int u_map;             // kTrunkCA (training): out map of the saved pre-attention activation u (bf16), or -1
  float alpha;           // kTrunkRes: out = alpha * (acc + bias) + residual
  int no_res;            // kTrunkRes: no residual term at all (out = alpha * (acc + bias)); HAN's body conv
  const float* bias;
  const float* res_f32;  // kTrunkRes: fp32 NHWC residual in global memory; nullptr = the TMEM stream
  float* out_f32;        // kTrunkRes: optional fp32 NHWC copy of the result (a later layer's res_f32)
  const float *w1, *b1, *w2, *b2;          // kTrunkCA: FC weights [cr][64], [cr], [64][cr], [64]
  float *save_mean, *save_hid, *save_y;    // kTrunkCA (training): CA vectors for backward, or nullptr
  const float* q_scale;  // [N][64] meta-attention multipliers or nullptr: of the CA vector (kTrunkCA, Q-RCAN) /
                         // of the scaled branch alpha * (acc + bias) (kTrunkRes, Q-EDSR)
};

struct TrunkArgs {
  const TrunkLayer* layers;
  const CUtensorMap* in_maps;
  const CUtensorMap* out_maps;
  const float* s_init;     // fp32 NHWC: initial residual stream (the head conv's output)
  int* ready;              // [T]   number of layers whose bf16 output of this tile is visible
  unsigned long long* pool_partial;   // [2][T][64] (fp32 channel sum, epoch) pairs, one 8-byte store each
  long long* dbg;          // optional timeline, [grid][dbg_layers][2][16] clock64 stamps
  int n_layers, N, H, W, tiles_x, tiles_y, tiles_per_img, T, K, w_layer0, cr, interleave, dbg_layers;
  float inv_hw;
  int sync_mode;   // debug: 1 acquire polls, 2 consumer proxy fence, 4 producer proxy fence, 8 release flag store
};

constexpr int kTrunkThreads = 384;   // 8 epilogue warps | A producer | MMA issuer 0 | weight producer | MMA issuer 1
constexpr int kTrunkAStages = 4;
constexpr int kTrunkMaxK = 4;
constexpr int kTrunkWBytes = 9 * 64 * 128;   // 72 KB: [kx*3+ky][64 rows][64 k] bf16
constexpr int kTrunkWThird = 3 * 64 * 128;   // 24 KB: the three ky taps of one kx
constexpr int kTrunkAccCol = 256;            // TMEM: stream S_j at column 64 j, accumulator j at 256 + 64 j
constexpr int kTrunkPoolRows = 32;             // rows of (value, epoch) pairs fetched per cp.async round
constexpr int kTrunkPoolBytes = kTrunkPoolRows * 64 * 8;   // 16 KB per epilogue group
constexpr size_t kTrunkSmemBytes =
    1024 + kTrunkWBytes + size_t(kTrunkAStages) * kAStageBytes + 2 * kABytes + 2 * kTrunkPoolBytes;

#ifdef RB_TRUNK_KERNEL_IMPL

static __device__ __noinline__ void trunk_watchdog_fail(const int* p, int target, int what) {
  printf("rumpy_b200: trunk watchdog: block %d thread %d waiting (%d) for %d at %p (now %d)\n", (int)blockIdx.x,
         (int)threadIdx.x, what, target, (const void*)p, *(volatile const int*)p);
  __trap();
}
__device__ __forceinline__ void poll_ge(const int* p, int target, int what, bool acq) {
  if ((acq ? ld_acquire_s32(p) : ld_relaxed_s32(p)) >= target) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while ((acq ? ld_acquire_s32(p) : ld_relaxed_s32(p)) < target) {
    __nanosleep(20);
    if ((++spins & 255u) == 0 && clock64() - t0 > RB_WATCHDOG_CYCLES) trunk_watchdog_fail(p, target, what);
  }
}

// Column sum of one image's (value, epoch) partial rows for channel c = row & 63 over the rows of parity row >> 6
// (fixed order: deterministic).  The rows are fetched with cp.async.cg -- 16-byte copies that bypass L1 and stay
// in flight together; gpu-scope relaxed / .cg loads are issued one at a time by the hardware (16 of them cost
// ~4 000 cycles here).  A stale entry (epoch mismatch) falls back to a polling reload of that entry.
// Called by all 128 threads of an epilogue group; `scratch` = the group's kTrunkPoolBytes of shared memory.
__device__ __forceinline__ float pool_column_sum(const unsigned long long* pp, int P, unsigned epoch,
                                                 unsigned long long* scratch, int row, uint32_t bar_id) {
  const int c = row & 63, hsel = row >> 6;
  float ssum = 0.f;
  const long long t0 = clock64();
  for (int r0 = 0; r0 < P; r0 += kTrunkPoolRows) {
    const int rows = P - r0 < kTrunkPoolRows ? P - r0 : kTrunkPoolRows;
    for (int r = row >> 5; r < rows; r += 4)
      cp_async_cg16(smem_u32(scratch + r * 64 + (row & 31) * 2), pp + size_t(r0 + r) * 64 + (row & 31) * 2);
    cp_async_commit();
    cp_async_wait_all();
    named_bar_sync(bar_id, 128);
    for (int r = hsel; r < rows; r += 2) {
      unsigned long long v = scratch[r * 64 + c];
      uint32_t spins = 0;
      while (unsigned(v >> 32) != epoch) {
        __nanosleep(20);
        v = ld_relaxed_u64(pp + size_t(r0 + r) * 64 + c);
        if ((++spins & 255u) == 0 && clock64() - t0 > RB_WATCHDOG_CYCLES)
          trunk_watchdog_fail(reinterpret_cast<const int*>(pp + size_t(r0 + r) * 64 + c) + 1, int(epoch), 2);
      }
      ssum += __uint_as_float(unsigned(v));
    }
    if (r0 + kTrunkPoolRows < P) named_bar_sync(bar_id, 128);   // scratch is refilled by the next round
  }
  return ssum;
}

// f[i] of lane l  ->  returns sum over the 32 lanes of element f[lane] (fixed order: deterministic)
template <int HALF>
__device__ __forceinline__ void xpose_step(float (&f)[32], int lane) {
  const bool up = (lane & HALF) != 0;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const float send = up ? f[i] : f[i + HALF];
    const float keep = up ? f[i + HALF] : f[i];
    f[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
  }
}
// 32 broadcast floats from shared memory as eight unconditional 16-byte loads (inside a `valid ? ... : 0` select the
// compiler emits 32 predicated scalar loads instead: four times the shared-memory wavefronts).
__device__ __forceinline__ void lds_bcast32(const float* p, float (&b)[32]) {
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const float4 t = reinterpret_cast<const float4*>(p)[i4];
    b[4 * i4] = t.x; b[4 * i4 + 1] = t.y; b[4 * i4 + 2] = t.z; b[4 * i4 + 3] = t.w;
  }
}
__device__ __forceinline__ float lane_transpose_sum32(float (&f)[32], int lane) {
  xpose_step<16>(f, lane);
  xpose_step<8>(f, lane);
  xpose_step<4>(f, lane);
  xpose_step<2>(f, lane);
  xpose_step<1>(f, lane);
  return f[0];
}

// 16 columns per lane: after the steps over lane bits 4..1 lane l holds column (l >> 1) summed over 16 rows; the last
// exchange adds the other 16.  Returns the sum over the warp's 32 rows of column (lane >> 1), at every lane.
__device__ __forceinline__ float lane_transpose_sum16(float (&f)[16], int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool up = (lane & (2 * half)) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < half) {
        const float send = up ? f[i] : f[i + half];
        const float keep = up ? f[i + half] : f[i];
        f[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * half);
      }
    }
  }
  return f[0] + __shfl_xor_sync(0xffffffffu, f[0], 1);
}

__global__ void __launch_bounds__(kTrunkThreads, 1)
trunk_pipe_kernel(const __grid_constant__ CUtensorMap w_map, const TrunkArgs args) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kTrunkAStages];
  __shared__ __align__(8) uint64_t a_empty[kTrunkAStages];
  __shared__ __align__(8) uint64_t w_full[3];
  __shared__ __align__(8) uint64_t w_empty[3];
  __shared__ __align__(8) uint64_t acc_full[kTrunkMaxK];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float y_s[kTrunkMaxK][64];
  __shared__ __align__(16) float bias_s[2][64], alpha_s[2][64], red_s[2][4][64];

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

#define TR_STAMP(L_, j_, slot_)                                                                       \
  do {                                                                                                \
    if (args.dbg && (L_) < args.dbg_layers && (j_) < 2)                                               \
      args.dbg[((size_t(cta) * args.dbg_layers + (L_)) * 2 + (j_)) * 16 + (slot_)] = clock64();        \
  } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTrunkAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    // w_empty: one tcgen05.commit per MMA-issuing warp (two when the CTA owns at least two tiles)
    for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], my_k >= 2 ? 2 : 1); }
    for (int i = 0; i < kTrunkMaxK; ++i) { mbar_init(&acc_full[i], 1); }
    fence_mbar_init();
  }
  if (warp == kWarpW && lane == 0) tma_prefetch_desc(&w_map);
  if (warp == kWarpMma) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == kWarpA) {
    // ===================================================================== A-operand producer
    // Two MMA issuers => two independent stage rings (stages {0,1} feed issuer 0, {2,3} issuer 1): a parity wait only
    // tells consecutive phases apart, so every ring must have exactly one consumer.  One tile per CTA: one ring of 4.
    const bool two = my_k >= 2;
    uint32_t cnt[2] = {0, 0};   // groups loaded into each ring
    int g = 0;                  // running tile counter: tile g belongs to issuer g & 1
    for (int L = 0; L < n_layers; ++L) {
      const CUtensorMap* im = args.in_maps + args.layers[L].in_map;
      if (lane == 0 && L + 1 < n_layers) {   // descriptors of the next layer (training plans use one map per layer)
        tma_prefetch_desc(args.in_maps + args.layers[L + 1].in_map);
        tma_prefetch_desc(args.out_maps + args.layers[L + 1].out_map);
        if (args.layers[L + 1].u_map >= 0) tma_prefetch_desc(args.out_maps + args.layers[L + 1].u_map);
      }
      for (int j = 0; j < my_k; ++j, ++g) {
        const int ring = two ? (g & 1) : 0;
        const int t = cta + j * G;
        const int n = t / P, rem = t - n * P;
        const int ty = rem / args.tiles_x, tx = rem - ty * args.tiles_x;
        if (L > 0) {
          // layer L reads layer L-1's output of this tile and of its 8 neighbours (same image)
          if (lane < 9) {
            const int nty = ty + lane / 3 - 1, ntx = tx + lane % 3 - 1;
            if (nty >= 0 && nty < args.tiles_y && ntx >= 0 && ntx < args.tiles_x)
              poll_ge(args.ready + n * P + nty * args.tiles_x + ntx, L, 1, (args.sync_mode & 1) != 0);
          }
          __syncwarp();
        }
        if (lane == 0) TR_STAMP(L, j, 0);
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t c = cnt[ring]++;
          const int stage = two ? 2 * ring + int(c & 1u) : int(c & 3u);
          const uint32_t phase = (two ? (c >> 1) : (c >> 2)) & 1u;
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (elect_one()) {
            if (args.sync_mode & 2) fence_proxy_async_all();
            mbar_expect_tx(&a_full[stage], kAStageBytes);
            tma_load_4d(a_s + stage * kAStageBytes, im, &a_full[stage], 0, tx * kTileW + kx - 1, ty * kTileH - 1, n);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma2) {
    // ===================================================================== MMA issuers (two warps)
    // One warp cannot issue 128x64x16 MMAs as fast as the tensor pipe retires them (48 cycles each; the descriptor
    // arithmetic costs 55-70, tools/experiments/umma_env_test.cu), so consecutive tiles -- different accumulators --
    // alternate between two issuing warps.  Each has its own A-stage ring (a parity wait only tells consecutive phases
    // apart: a ring must have one consumer), commits the stages and accumulators of its own tiles, and the weight
    // thirds are released by one commit per issuer.
    const int me = warp == kWarpMma ? 0 : 1;
    const bool two = my_k >= 2;
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
    int g = 0;
    uint32_t cnt = 0;   // groups consumed from this issuer's own stage ring (stages {2 me, 2 me + 1}; one ring of 4 when alone)
    if (two || me == 0) {
      for (int L = 0; L < n_layers; ++L) {
        int my_last = -1;
        for (int j = 0; j < my_k; ++j)
          if (!two || ((g + j) & 1) == me) my_last = j;
        bool have_w = false;
        for (int j = 0; j < my_k; ++j, ++g) {
          if (two && (g & 1) != me) continue;
          const uint32_t d_tmem = tmem_base + uint32_t(kTrunkAccCol + j * 64);
          for (int kx = 0; kx < 3; ++kx) {
            if (!have_w) mbar_wait_trap(&w_full[kx], uint32_t(L & 1));
            const uint32_t c = cnt++;
            const int stage = two ? 2 * me + int(c & 1u) : int(c & 3u);
            mbar_wait_trap(&a_full[stage], (two ? (c >> 1) : (c >> 2)) & 1u);
            {
              tc_fence_after();
              if (kx == 0 && lane == 0) TR_STAMP(L, j, 1);
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
                if (j == my_last) umma_commit(&w_empty[kx]);   // this issuer's last read of the third in layer L
              }
              __syncwarp();
            }
          }
          {
            have_w = true;
            if (elect_one()) umma_commit(&acc_full[j]);
            __syncwarp();
            if (lane == 0) TR_STAMP(L, j, 2);
          }
        }
      }
    }
  } else if (warp == kWarpW) {
    // ===================================================================== weight producer (three kx thirds)
    for (int L = 0; L < n_layers; ++L) {
      for (int kx = 0; kx < 3; ++kx) {
        mbar_wait(&w_empty[kx], uint32_t(L & 1) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&w_full[kx], kTrunkWThird);
          tma_load_4d(w_s + kx * kTrunkWThird, &w_map, &w_full[kx], 0, 0, kx * 3, args.w_layer0 + L);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ===================================================================== epilogue groups (2 x 128 threads)
    const int e = warp >> 2;
    const int q = warp & 3;                  // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;           // pixel of the tile == TMEM lane == thread index in the group
    const int ly = row >> 4, lx = row & 15;
    const uint32_t swz = uint32_t(row & 7);
    const uint32_t bar_id = 1u + uint32_t(e);
    uint8_t* stg = stg_s + e * kABytes;
    uint8_t* my_stg = stg + row * 128;
    unsigned long long* pool_scr = reinterpret_cast<unsigned long long*>(stg_s + 2 * kABytes + e * kTrunkPoolBytes);
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    float* bias_e = bias_s[e];
    float* alpha_e = alpha_s[e];   // kTrunkRes: alpha (x the Q-EDSR meta-attention multiplier of this image's channel)

    // ---- residual stream of the owned tiles: global fp32 -> TMEM
    for (int j = e; j < my_k; j += 2) {
      const int t = cta + j * G;
      const int n = t / P, rem = t - n * P;
      const int ty = rem / args.tiles_x, tx = rem - ty * args.tiles_x;
      const int y = ty * kTileH + ly, x = tx * kTileW + lx;
      const bool valid = y < args.H && x < args.W;
      const float4* src = reinterpret_cast<const float4*>(args.s_init + ((size_t(n) * args.H + y) * args.W + x) * 64);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) f = __ldg(src + h * 8 + c4);
          v[c4 * 4 + 0] = __float_as_uint(f.x); v[c4 * 4 + 1] = __float_as_uint(f.y);
          v[c4 * 4 + 2] = __float_as_uint(f.z); v[c4 * 4 + 3] = __float_as_uint(f.w);
        }
        tmem_st32(lane_addr + uint32_t(j * 64 + h * 32), v);
      }
    }
    tmem_st_wait();

    int ca_seen = 0;
    for (int L = 0; L < n_layers; ++L) {
      const TrunkLayer* lay = args.layers + L;
      const int kind = lay->kind;
      const float* bias = lay->bias;
      const CUtensorMap* om = args.out_maps + lay->out_map;

      // the staged bf16 tile -> global memory, then publish the tile's epoch once the data is visible gpu-wide.
      // Called by all 128 threads of the group after they have written their staging rows.
      auto finish_tile = [&](int j, int t, int n, int ty, int tx) {
        tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (row == 0) {
          TR_STAMP(L, j, 10);
          tma_store_4d(om, stg, 0, tx * kTileW, ty * kTileH, n);
          tma_store_commit();
          tma_store_wait_all0();   // the tile is in L2 (the gpu-scope point of coherence) when this returns
          TR_STAMP(L, j, 11);
          if (args.sync_mode & 4) fence_proxy_async_all();
          if (args.sync_mode & 8) st_release_s32(args.ready + t, L + 1);
          else st_relaxed_s32(args.ready + t, L + 1);
          TR_STAMP(L, j, 6);
        }
      };
      auto stage_bf16 = [&](const float (&f)[32], int h) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(my_stg + ((uint32_t(h * 4 + c) ^ swz) << 4)) =
              make_uint4(pack_bf16x2(f[c * 8], f[c * 8 + 1]), pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]),
                         pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]));
      };

      // ---------------------------------------------------------------- conv + bias (+ReLU | + residual)
      auto plain = [&](int j) {
        const int t = cta + j * G;
        const int n = t / P, rem = t - n * P;
        const int ty = rem / args.tiles_x, tx = rem - ty * args.tiles_x;
        const int y = ty * kTileH + ly, x = tx * kTileW + lx;
        const bool valid = y < args.H && x < args.W;
        const float bv = row < 64 ? __ldg(bias + row) : 0.f;
        const float* res = lay->res_f32;
        float* outf = lay->out_f32;
        const float alpha = lay->alpha;
        const int update_s = lay->update_s;
        const size_t pix = ((size_t(n) * args.H + y) * args.W + x) * 64;
        mbar_wait(&acc_full[j], uint32_t(L & 1));
        tc_fence_after();
        if (row == 0) TR_STAMP(L, j, 3);
        if (row < 64) {
          bias_e[row] = bv;
          alpha_e[row] = lay->q_scale != nullptr ? alpha * __ldg(lay->q_scale + n * 64 + row) : alpha;
        }
        named_bar_sync(bar_id, 128);
        if (row == 0) TR_STAMP(L, j, 8);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          float f[32];
          tmem_ld32(lane_addr + uint32_t(kTrunkAccCol + j * 64 + h * 32), v);
          if (kind == kTrunkRelu) {
            float bb[32];
            lds_bcast32(bias_e + h * 32, bb);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(__uint_as_float(v[i]) + bb[i], 0.f);
          } else {
            if (lay->no_res) {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = 0.f;
            } else if (res != nullptr) {
              tmem_ld_wait();
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) r = *reinterpret_cast<const float4*>(res + pix + h * 32 + c4 * 4);
                f[c4 * 4 + 0] = r.x; f[c4 * 4 + 1] = r.y; f[c4 * 4 + 2] = r.z; f[c4 * 4 + 3] = r.w;
              }
            } else {
              uint32_t s[32];
              tmem_ld32(lane_addr + uint32_t(j * 64 + h * 32), s);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(s[i]);
            }
            {
              float bb[32], aa[32];
              lds_bcast32(bias_e + h * 32, bb);
              lds_bcast32(alpha_e + h * 32, aa);
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = (__uint_as_float(v[i]) + bb[i]) * aa[i] + f[i];
            }
            if (update_s) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(f[i]);
              tmem_st32(lane_addr + uint32_t(j * 64 + h * 32), v);
            }
            if (outf != nullptr && valid) {
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4)
                *reinterpret_cast<float4*>(outf + pix + h * 32 + c4 * 4) =
                    make_float4(f[c4 * 4], f[c4 * 4 + 1], f[c4 * 4 + 2], f[c4 * 4 + 3]);
            }
          }
          stage_bf16(f, h);
        }
        tmem_st_wait();
        if (row == 0) TR_STAMP(L, j, 9);
        finish_tile(j, t, n, ty, tx);
      };

      // ---------------------------------------------------------------- channel attention, phase 1: pool
      auto ca_pool = [&](int j) {
        const int t = cta + j * G;
        const int n = t / P, rem = t - n * P;
        const int ty = rem / args.tiles_x, tx = rem - ty * args.tiles_x;
        const bool valid = (ty * kTileH + ly) < args.H && (tx * kTileW + lx) < args.W;
        const float bv = row < 64 ? __ldg(bias + row) : 0.f;
        const int u_map = lay->u_map;
        mbar_wait(&acc_full[j], uint32_t(L & 1));
        tc_fence_after();
        if (row == 0) TR_STAMP(L, j, 3);
        if (row < 64) bias_e[row] = bv;
        if (u_map >= 0 && row == 0) tma_store_wait_read0();   // an earlier u store has read the staging tile
        named_bar_sync(bar_id, 128);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          float f[32];
          tmem_ld32(lane_addr + uint32_t(kTrunkAccCol + j * 64 + h * 32), v);
          float bb[32];
          lds_bcast32(bias_e + h * 32, bb);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = valid ? __uint_as_float(v[i]) + bb[i] : 0.f;
          if (u_map >= 0) stage_bf16(f, h);
          red_s[e][q][h * 32 + lane] = lane_transpose_sum32(f, lane);
        }
        tc_fence_before();
        if (u_map >= 0) fence_proxy_async_smem();
        if (row == 0) TR_STAMP(L, j, 12);
        named_bar_sync(bar_id, 128);
        if (row < 64) {
          const float s = (red_s[e][0][row] + red_s[e][1][row]) + (red_s[e][2][row] + red_s[e][3][row]);
          st_relaxed_u64(args.pool_partial + (size_t(lay->ca_slot & 1) * T + t) * 64 + row,
                         (static_cast<unsigned long long>(unsigned(lay->ca_slot + 1)) << 32) | __float_as_uint(s));
        }
        if (row == 0) TR_STAMP(L, j, 4);
        if (u_map >= 0 && row == 0) {
          tma_store_4d(args.out_maps + u_map, stg, 0, tx * kTileW, ty * kTileH, n);
          tma_store_commit();
        }
        named_bar_sync(bar_id, 128);   // red_s is free for the group's next tile
      };

      // ---------------------------------------------------------------- phase 2: x + u*y from the same accumulator
      auto ca_apply = [&](int j) {
        const int t = cta + j * G;
        const int n = t / P, rem = t - n * P;
        const int ty = rem / args.tiles_x, tx = rem - ty * args.tiles_x;
        // ---- y = sigmoid(W2 relu(W1 mean + b1) + b2) for this tile's image, computed by the group itself.
        // Pool partials are (value, epoch) pairs written with single 8-byte stores: no fence, no counter.
        const int cr = args.cr;
        const unsigned epoch = unsigned(lay->ca_slot + 1);
        const unsigned long long* pp = args.pool_partial + (size_t(lay->ca_slot & 1) * T + size_t(n) * P) * 64;
        const int c = row & 63, hsel = row >> 6;
        // FC parameters of the first four hidden units into registers while the partials are still in flight
        const float *w1 = lay->w1, *b1 = lay->b1, *w2 = lay->w2;
        float w1a[4], w1b[4], w2r[4], b1r[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const bool on = h < cr;
          w1a[h] = on ? __ldg(w1 + h * 64 + lane) : 0.f;
          w1b[h] = on ? __ldg(w1 + h * 64 + 32 + lane) : 0.f;
          w2r[h] = on ? __ldg(w2 + c * cr + h) : 0.f;
          b1r[h] = on ? __ldg(b1 + h) : 0.f;
        }
        float yacc = __ldg(lay->b2 + c);
        const float ssum = pool_column_sum(pp, P, epoch, pool_scr, row, bar_id);
        if (row == 0) TR_STAMP(L, j, 13);
        if (lay->u_map >= 0 && row == 0) tma_store_wait_read0();   // the u store has read the staging tile
        red_s[e][hsel][c] = ssum;
        named_bar_sync(bar_id, 128);   // (also: row 0 is past its previous store's wait, the staging tile is free)
        const float m0 = (red_s[e][0][lane] + red_s[e][1][lane]) * args.inv_hw;
        const float m1 = (red_s[e][0][lane + 32] + red_s[e][1][lane + 32]) * args.inv_hw;
        const bool saver = lay->save_y != nullptr && rem == 0 && q == 0;   // one warp records the CA vectors
#pragma unroll
        for (int h = 0; h < 4; ++h) {   // every warp computes every hidden unit (cr is tiny): no second barrier
          float sdot = w1a[h] * m0 + w1b[h] * m1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
          const float hv = fmaxf(sdot + b1r[h], 0.f);
          yacc = fmaf(w2r[h], hv, yacc);
          if (saver && lane == 0 && h < cr) lay->save_hid[n * cr + h] = hv;
        }
        for (int h = 4; h < cr; ++h) {
          float sdot = __ldg(w1 + h * 64 + lane) * m0 + __ldg(w1 + h * 64 + 32 + lane) * m1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
          const float hv = fmaxf(sdot + __ldg(b1 + h), 0.f);
          yacc = fmaf(__ldg(w2 + c * cr + h), hv, yacc);
          if (saver && lane == 0) lay->save_hid[n * cr + h] = hv;
        }
        if (hsel == 0) {
          const float yr = 1.f / (1.f + __expf(-yacc));
          y_s[j][c] = yr * (lay->q_scale ? __ldg(lay->q_scale + n * 64 + c) : 1.f);
          if (lay->save_y != nullptr && rem == 0) lay->save_y[n * 64 + c] = yr;   // backward wants the raw sigmoid
        }
        if (saver) {
          lay->save_mean[n * 64 + lane] = m0;
          lay->save_mean[n * 64 + 32 + lane] = m1;
        }
        named_bar_sync(bar_id, 128);
        if (row == 0) TR_STAMP(L, j, 14);
        const float* yv = y_s[j];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32], s[32];
          float f[32];
          tmem_ld32(lane_addr + uint32_t(kTrunkAccCol + j * 64 + h * 32), v);
          tmem_ld32(lane_addr + uint32_t(j * 64 + h * 32), s);
          float bb[32], yy[32];
          lds_bcast32(bias_e + h * 32, bb);
          lds_bcast32(yv + h * 32, yy);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaf(__uint_as_float(v[i]) + bb[i], yy[i], __uint_as_float(s[i]));
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = __float_as_uint(f[i]);
          tmem_st32(lane_addr + uint32_t(j * 64 + h * 32), s);
          stage_bf16(f, h);
        }
        tmem_st_wait();
        if (row == 0) TR_STAMP(L, j, 9);
        finish_tile(j, t, n, ty, tx);
      };

      if (kind != kTrunkCA) {
        for (int j = e; j < my_k; j += 2) plain(j);
      } else {
        if (args.interleave) {
          // every image lives in ONE tile slot (the grid is a whole number of images): slot j's exchange only needs
          // slot-j pools, which never wait on anything => pool(j), apply(j) back to back is deadlock-free, and the
          // first tile of the next layer does not have to wait for this group's last MMA
          for (int j = e; j < my_k; j += 2) { ca_pool(j); ca_apply(j); }
        } else {
          // any tile-to-slot assignment: all pools first (no pool ever waits on another CTA), then the applies
          for (int j = e; j < my_k; j += 2) ca_pool(j);
          for (int j = e; j < my_k; j += 2) ca_apply(j);
        }
        ++ca_seen;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
#undef TR_STAMP
}

#endif  // RB_TRUNK_KERNEL_IMPL

}  // namespace rb
