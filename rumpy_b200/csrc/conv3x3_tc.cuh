// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces the reference's `common.default_conv` -> nn.Conv2d call sites on the trunk
// (/root/reference/rumpy/SISR/models/advanced/common.py:6-9 as used by architectures.py:70-78,114-119,
// 153-165,216-231 and common.py:30-41,60-66), forward and (with repacked weights) dgrad.
//
// Data layout in HBM:  activations NHWC, bf16 operand copies (64 channels = one 128-byte row = one
// SWIZZLE_128B atom row) and an fp32 copy of the residual stream; weights packed [tap][Cout][Cin] bf16.
//
// GEMM view: D[128 pixels, BN out-channels] += A[128 pixels, 64 in-channels] * B[BN, 64]^T, summed over
// K-blocks = 9 taps x (Cin/64) channel chunks.  The 128-pixel M tile is an 8x16 spatial patch of one image;
// the A operand of tap (ky,kx) is the same patch shifted by (ky-1,kx-1), fetched by ONE 4-D TMA box
// {64ch,16,8,1} whose out-of-bounds elements the TMA unit zero-fills -- the convolution's zero padding and
// ragged image edges cost nothing.  Accumulators live in TMEM (double buffered, 2 x BN columns).
//
// Warp roles (192 threads, one persistent CTA per SM):
//   warp 0      TMA producer (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one lane)
//   warps 2..5  epilogue: tcgen05.ld -> bias / ReLU / alpha / mask / fp32 residual -> swizzled smem staging
//               -> TMA store (bf16 and/or fp32), plus per-tile channel sums for the channel-attention pool
#pragma once
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace rb {

enum ConvFlags : uint32_t {
  kConvRelu = 1u,       // v = max(v, 0)
  kConvOutBf16 = 2u,    // store bf16 NHWC (optionally through r*r pixel-shuffle maps)
  kConvOutF32 = 4u,     // store fp32 NHWC
  kConvResF32 = 8u,     // v += residual (fp32 NHWC)
  kConvMask = 16u,      // v = mask > 0 ? v : 0   (mask: bf16 NHWC; ReLU backward)
  kConvPool = 32u,      // write per-tile channel sums (2 half-tile partials) for the CA global average pool
};

struct ConvMaps {
  CUtensorMap a[9];   // bf16 input maps (1 normally; r*r when the input is read through a pixel-unshuffle)
  CUtensorMap w;      // packed weights, dims {Cin, Cout, 9}
  CUtensorMap ob[9];  // bf16 output maps (1 normally; r*r for the pixel-shuffle store)
  CUtensorMap of;     // fp32 output
  CUtensorMap rf;     // fp32 residual input
  CUtensorMap mb;     // bf16 mask input
};

struct ConvArgs {
  int N, H, W;
  int tiles_x, tiles_y, m_tiles, n_tiles;
  int cin_chunks;        // Cin / 64
  int a_chunks_per_map;  // 64-channel K chunks served by each input map
  int o_chunks_per_map;  // 64-channel output chunks served by each bf16 output map
  int cout;              // total output channels (packed-row order)
  int stages;            // smem pipeline depth
  int stg_bufs;          // epilogue staging slots (1 or 2)
  float alpha;           // scale on (acc + bias) (res_scale / gradient scale)
  const float* ch_scale; // [N][cout] extra per-(image, channel) scale on the same term (meta-attention), or nullptr
  const float* bf16_scale;  // [N][cout] scale applied to the bf16 output ONLY (fp32 output unscaled): Q-EDSR backward,
                            // the next dgrad / wgrad operand is g * q while the fp32 skip gradient stays g; or nullptr
  const float* bias;     // [cout] in packed-row order, or nullptr
  float* pool_partial;   // [m_tiles][2][cout]
  float* out_nchw;       // BN == 16 variant only: fp32 NCHW [N][cout_real][H][W]
  int cout_real;         // BN == 16 variant only: number of real output channels (<= 16)
  long long* dbg;        // optional per-CTA timeline (2 x 16 clock64 slots per CTA); nullptr in production
  int dbg_mode;          // timing experiments (garbage results), 0 in production: 1 = A boxes fetched only for the first
                         // `stages` fills, 4 = no MMAs issued, 8 = timeline stamps from the pooled convs only,
                         // 16 = epilogue stages nothing (drains the accumulator only), 32 = no pool sums, 64 = no stores
  uint32_t flags;
};

constexpr int kTileH = 8, kTileW = 16, kTileM = 128;
constexpr int kABoxH = kTileH + 2;                 // A box rows: the tile plus one halo row above and below
constexpr int kABytes = kTileM * 128;              // one 128-pixel x 64-channel bf16 operand / staging tile
constexpr int kAStageBytes = kABoxH * kTileW * 128;  // 20 KB: serves the three ky taps of one (kx, chunk)
// Resident-weights geometry (64 input channels): the tile is 16 rows x 8 pixels and ONE halo box {64 ch, 10 px, 18 rows}
// serves all nine taps: tap (ky,kx) is the same descriptor with the start address moved by (ky*10 + kx) * 128 B and
// SBO = 1280 B (next image row = next 8-pixel group).  The tensor core applies the 128-byte swizzle to the absolute
// shared-memory address, so a start that is not atom-aligned reads exactly what TMA wrote
// (tools/experiments/umma_sw128_shift_test.cu: exact at pitch 16 and 10).  23 KB fetched per tile instead of 60 KB.
constexpr int kTallH = 16, kTallW = 8;
constexpr int kTallBoxH = kTallH + 2, kTallBoxW = kTallW + 2;
constexpr int kTallBoxBytes = kTallBoxH * kTallBoxW * 128;            // 23,040
constexpr int kTallStageBytes = (kTallBoxBytes + 1023) & ~1023;       // stages stay 1024-byte aligned
constexpr int kStgF32Bytes = 2 * kABytes;          // fp32 staging: two 32-channel halves
constexpr int kStgBf16Bytes = kABytes;
constexpr int kMaxStages = 8;
constexpr int kEpiGroups = 2;                      // epilogue groups of 4 warps: one 32-channel half each
constexpr int kEpiThreads = 128 * kEpiGroups;
// Warp roles: 0 TMA producer | 1 MMA issuer (even tiles) | 2 .. epilogue sets of kEpiGroups x 4 warps | store warp |
// second MMA issuer.  Resident weights: two epilogue sets (even / odd tiles) and two issuers = 20 warps; streamed
// weights: one set, one issuer = 11 warps (and twice the registers per thread).
__host__ __device__ constexpr int conv_epi_sets(bool resident) { return resident ? 2 : 1; }
__host__ __device__ constexpr int conv_warp_store(bool resident) { return 2 + 4 * kEpiGroups * conv_epi_sets(resident); }
__host__ __device__ constexpr int conv_threads(bool resident) { return (conv_warp_store(resident) + (resident ? 2 : 1)) * 32; }
constexpr size_t kConvSmemBudget = 227 * 1024 - 5120;   // static shared memory: barriers, bias, pool sums (~4.5 KB)

__host__ __device__ constexpr int conv_b_block_bytes(int bn) { return bn * 128; }

// Epilogue staging: fp32 tile (32 KB) if any fp32 input/output, bf16 tile (16 KB) if any bf16 input/output;
// double buffered (args.stg_bufs = 2) whenever shared memory allows: with TMA inputs (residual / mask) the second
// slot lets the next tile's inputs be prefetched while the current tile is processed in place.
__host__ __device__ inline int conv_stg_buf_bytes(uint32_t flags) {
  return ((flags & (kConvOutF32 | kConvResF32)) ? kStgF32Bytes : 0) +
         ((flags & (kConvOutBf16 | kConvMask)) ? kStgBf16Bytes : 0);
}


// Dynamic smem: [staging | resident B (optional) | stages x (A box [+ 3 B taps])], 1024-byte aligned.
__host__ inline size_t conv_smem_bytes(int bn, bool resident_b, int cin_chunks, int stages, uint32_t flags,
                                       int stg_bufs) {
  size_t s = 1024;
  if (bn != 16) s += size_t(stg_bufs) * conv_stg_buf_bytes(flags);
  if (resident_b) s += size_t(9) * cin_chunks * conv_b_block_bytes(bn);
  s += size_t(stages) * (resident_b ? kTallStageBytes : kAStageBytes + 3 * conv_b_block_bytes(bn));
  return s;
}

constexpr int kDbgTile = 40;   // the steady-state tile whose epilogue / MMA issue the second timeline region records
#define RB_STAMP_ON (args.dbg && (!(args.dbg_mode & 8) || (args.flags & kConvPool)))   // mode 8: pooled convs only
#define RB_STAMP(slot) do { if (RB_STAMP_ON) args.dbg[blockIdx.x * 16 + (slot)] = clock64(); } while (0)
// second region (after gridDim.x * 16 slots): the epilogue of this CTA's third tile (its TMA-issuing thread) in slots
// 0..7, cycles the MMA warp waited for an accumulator / for A stages and the producer for free stages in 12..14
#define RB_STAMP2(slot) do { if (RB_STAMP_ON && it == kDbgTile && et == 0 && tp == 0) args.dbg[(gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define RB_STAMP_NS(slot) do { if (RB_STAMP_ON) args.dbg[blockIdx.x * 16 + (slot)] = global_timer_ns(); } while (0)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int BN, bool RESIDENT_B>
__global__ void __launch_bounds__(conv_threads(RESIDENT_B), 1)
conv3x3_tc_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args) {
  static_assert(BN == 16 || BN == 64 || BN == 128 || BN == 256, "BN must be 16, 64, 128 or 256");
  constexpr int kBBlock = BN * 128;
  constexpr uint32_t kTmemCols = 2 * BN;   // 32 is the minimum TMEM allocation
  constexpr int kChunksPerTile = BN / 64;  // 0 for the thin (BN = 16) NCHW-output variant
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN);
  constexpr int kStageBytes = RESIDENT_B ? kTallStageBytes : kAStageBytes + 3 * kBBlock;
  constexpr int TH = RESIDENT_B ? kTallH : kTileH, TW = RESIDENT_B ? kTallW : kTileW;   // tile geometry

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t b_bar;
  __shared__ __align__(8) uint64_t in_bar[2];
  __shared__ __align__(8) uint64_t stg_full[2];    // epilogue -> store warp: slot staged
  __shared__ __align__(8) uint64_t stg_empty[2];   // store warp -> epilogue: slot read out
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[BN];
  constexpr int kEpiSets = conv_epi_sets(RESIDENT_B);
  constexpr int kWarpStore = conv_warp_store(RESIDENT_B);
  constexpr int kWarpMma2 = RESIDENT_B ? kWarpStore + 1 : -1;   // second MMA-issuing warp (odd tiles)
  __shared__ __align__(16) float pool_s[kEpiSets * 8][64];   // per epilogue warp channel sums of the current chunk

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  grid_dep_launch_dependents();   // PDL: let the next layer's CTAs start their prologue as SMs free up
  if (threadIdx.x == 0) { RB_STAMP(0); RB_STAMP_NS(14); }
  const int stages = args.stages;
  const int cin_chunks = args.cin_chunks;
  const uint32_t flags = args.flags;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stg_buf_bytes = (BN == 16) ? 0 : conv_stg_buf_bytes(flags);
  const int stg_bufs = args.stg_bufs;
  uint8_t* b_res = smem + stg_bufs * stg_buf_bytes;
  uint8_t* stage0 = b_res + (RESIDENT_B ? 9 * cin_chunks * kBBlock : 0);

  const int n_tile = blockIdx.x % args.n_tiles;
  const int mt_first = blockIdx.x / args.n_tiles;
  const int mt_stride = gridDim.x / args.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      // one arrival per epilogue WARP (lane 0 after __syncwarp): 32 lanes arriving on one barrier are 32 serialised
      // shared-memory operations on the data pipe the tensor core reads its operands through
      mbar_init(&tmem_empty_bar[i], BN == 16 ? 4 : kEpiThreads / 32);
      mbar_init(&stg_full[i], kEpiThreads / 32);
      mbar_init(&stg_empty[i], 1);
    }
    mbar_init(&b_bar, 1);
    mbar_init(&in_bar[0], 1);
    mbar_init(&in_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.w);
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) RB_STAMP(1);

  const int tiles_per_img = args.tiles_x * args.tiles_y;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // The whole warp walks the loop (warp-uniform control flow, uniform registers); one elected lane issues.
    if (RESIDENT_B) {
      if (elect_one()) {
        mbar_expect_tx(&b_bar, uint32_t(9 * cin_chunks) * kBBlock);
        for (int chunk = 0; chunk < cin_chunks; ++chunk)   // one box {64, BN, 9 taps} per 64-channel chunk
          tma_load_3d(b_res + chunk * 9 * kBBlock, &maps.w, &b_bar, chunk * 64, n_tile * BN, 0);
        RB_STAMP(2);
      }
      __syncwarp();
    }
    grid_dep_wait();   // weights are static; activations come from the previous kernel
    // two stage rings, one per MMA issuer (tile parity): a parity wait needs a single consumer per ring
    // (resident weights only; the streamed-weights variant keeps one issuer and one ring -- its tiles are 12+ fills
    // long and halving the ring cost more than the second issuer gained: EDSR 256-channel step 27.3 -> 32.6 ms)
    const int ring_size[2] = {RESIDENT_B ? stages - stages / 2 : stages, RESIDENT_B ? stages / 2 : 0};
    int ring_pos[2] = {0, 0};
    uint32_t ring_phase[2] = {0, 0};
    long long w_empty = 0;
    int it_p = 0;
    for (int mt = mt_first; mt < args.m_tiles; mt += mt_stride, ++it_p) {
      const int ring = RESIDENT_B ? (it_p & 1) : 0;
      const int n = mt / tiles_per_img;
      const int rem = mt - n * tiles_per_img;
      const int y0 = (rem / args.tiles_x) * TH;
      const int x0 = (rem % args.tiles_x) * TW;
      int mi = 0, cc = 0;  // input map index / 64-channel chunk inside that map
      for (int chunk = 0; chunk < cin_chunks; ++chunk) {
        for (int kx = 0; kx < (RESIDENT_B ? 1 : 3); ++kx) {
          const int stage = ring * ring_size[0] + ring_pos[ring];
          const uint32_t phase = ring_phase[ring];
          const long long t0 = args.dbg ? clock64() : 0;
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 200);
          if (args.dbg) w_empty += clock64() - t0;
          if ((args.dbg_mode & 1) && (phase || it_p > 1)) {
            if (elect_one()) mbar_arrive(&full_bar[stage]);
          } else if (elect_one()) {
            uint8_t* sa = stage0 + stage * kStageBytes;
            if constexpr (RESIDENT_B) {
              // one halo box {64 ch, 10 px, 18 rows} serves all nine taps
              mbar_expect_tx(&full_bar[stage], kTallBoxBytes);
              tma_load_4d(sa, &maps.a[mi], &full_bar[stage], cc * 64, x0 - 1, y0 - 1, n);
            } else {
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              // box {64 ch, 16 px, 10 rows}: rows y0-1 .. y0+8 at column offset kx-1 serve ky = 0, 1, 2
              tma_load_4d(sa, &maps.a[mi], &full_bar[stage], cc * 64, x0 + kx - 1, y0 - 1, n);
              tma_load_3d(sa + kAStageBytes, &maps.w, &full_bar[stage], chunk * 64, n_tile * BN, kx * 3);
            }
          }
          __syncwarp();
          if (++ring_pos[ring] == ring_size[ring]) { ring_pos[ring] = 0; ring_phase[ring] ^= 1; }
        }
        if (++cc == args.a_chunks_per_map) { cc = 0; ++mi; }
      }
    }
    if (RB_STAMP_ON && lane == 0) args.dbg[(gridDim.x + blockIdx.x) * 16 + 14] = w_empty;
  } else if (warp == 1 || warp == kWarpMma2) {
    // ===================================================================== MMA issuers (one elected lane each)
    // Two warps on different schedulers, alternating tiles: issuer `me` owns accumulator `me`.  One thread issues an
    // MMA every ~56-68 cycles here; two reach the tensor pipe's own rate.  Each issuer has its own stage ring (a
    // parity wait on a shared ring can be more than one phase away from the barrier and alias).
    const int me = warp == 1 ? 0 : 1;
    if (RESIDENT_B) mbar_wait_trap(&b_bar, 0);
    if (lane == 0 && me == 0) RB_STAMP(3);
    if (RESIDENT_B || me == 0) {   // streamed weights: one issuer, one ring
    const int ring_size = !RESIDENT_B ? stages : (me == 0 ? stages - stages / 2 : stages / 2);   // see the producer
    const int ring_base = (RESIDENT_B && me == 1) ? stages - stages / 2 : 0;
    int ring_pos = 0;
    uint32_t ring_phase = 0;
    int it = 0;
    long long w_acc = 0, w_full = 0;
    for (int mt = mt_first; mt < args.m_tiles; mt += mt_stride, ++it) {
      if (RESIDENT_B && (it & 1) != me) continue;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long t0 = args.dbg ? clock64() : 0;
      mbar_wait_trap(&tmem_empty_bar[acc], acc_phase ^ 1);
      if (args.dbg) w_acc += clock64() - t0;
      if (RB_STAMP_ON && it == kDbgTile && lane == 0) {
        args.dbg[(gridDim.x + blockIdx.x) * 16 + 10] = t0;          // tile kDbgTile: MMA warp ready for the tile
        args.dbg[(gridDim.x + blockIdx.x) * 16 + 11] = clock64();   // ... accumulator free, issue starts
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
      for (int chunk = 0; chunk < cin_chunks; ++chunk) {
        for (int kx = 0; kx < (RESIDENT_B ? 1 : 3); ++kx) {
          const int stage = ring_base + ring_pos;
          const uint32_t phase = ring_phase;
          if (++ring_pos == ring_size) { ring_pos = 0; ring_phase ^= 1; }
          const long long t1 = args.dbg ? clock64() : 0;
          mbar_wait_trap(&full_bar[stage], phase);
          if (args.dbg) w_full += clock64() - t1;
          tc_fence_after();
          if (it == 0 && chunk == 0 && kx == 0 && lane == 0) RB_STAMP(4);
          if (RESIDENT_B && elect_one()) {
            const uint32_t a_addr = smem_u32(stage0 + stage * kStageBytes);
            const uint32_t b_addr = smem_u32(b_res + chunk * 9 * kBBlock);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              if (args.dbg_mode & 4) break;
              // packed tap slot = kx*3 + ky: the box shifted by ky rows and kx pixels; 8-pixel groups one box row
              // (1280 B) apart
              const uint64_t adesc = make_smem_desc(a_addr + ((tap % 3) * kTallBoxW + tap / 3) * 128, 16,
                                                    kTallBoxW * 128, kLayoutSw128);
              const uint64_t bdesc = make_smem_desc(b_addr + tap * kBBlock, 16, 1024, kLayoutSw128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc, (chunk | tap | k) != 0);
            }
            umma_commit(&empty_bar[stage]);
          }
          if (!RESIDENT_B && elect_one()) {
            const uint32_t a_addr = smem_u32(stage0 + stage * kStageBytes);
            const uint32_t b_addr = a_addr + kAStageBytes;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              if (args.dbg_mode & 4) break;
              // tap (ky,kx): rows [16*ky, 16*ky+128) of the box -- a 2 KB (two swizzle atoms) shift
              const uint64_t adesc = make_smem_desc(a_addr + ky * (kTileW * 128), 16, 1024, kLayoutSw128);
              const uint64_t bdesc = make_smem_desc(b_addr + ky * kBBlock, 16, 1024, kLayoutSw128);
#pragma unroll
              for (int k = 0; k < 4; ++k)  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
                umma_bf16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), kIdesc,
                          (chunk | kx | ky | k) != 0);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (it == 0 && lane == 0) RB_STAMP(5);
      if (RB_STAMP_ON && it == kDbgTile && lane == 0) args.dbg[(gridDim.x + blockIdx.x) * 16 + 15] = clock64();  // issued
    }
    if (RB_STAMP_ON && lane == 0 && me == 0) {
      args.dbg[(gridDim.x + blockIdx.x) * 16 + 12] = w_acc;
      args.dbg[(gridDim.x + blockIdx.x) * 16 + 13] = w_full;
    }
    }
  } else if (warp == kWarpStore) {
    // ===================================================================== store warp (offload configuration)
    // Issues the output stores of every staged chunk, so that no epilogue thread queues behind the producer's boxes
    // in the TMA unit (measured at 1080p: ~650 cycles per tile on the epilogue's critical path), and hands the
    // staging slot back once the store has read it.  One lane: bulk groups are per thread.
    const bool has_in_s = (flags & (kConvResF32 | kConvMask)) != 0;
    if (BN != 16 && !has_in_s && args.stg_bufs == 2 && lane == 0) {
      const bool use_f32_s = (flags & kConvOutF32) != 0;
      grid_dep_wait();
      int cc = 0;
      for (int mt = mt_first; mt < args.m_tiles; mt += mt_stride) {
        const int n = mt / tiles_per_img;
        const int rem = mt - n * tiles_per_img;
        const int y0 = (rem / args.tiles_x) * TH;
        const int x0 = (rem % args.tiles_x) * TW;
        for (int j = 0; j < kChunksPerTile; ++j, ++cc) {
          const int oc = n_tile * kChunksPerTile + j;
          const int slot = cc & 1;
          uint8_t* stg_f32 = smem + slot * stg_buf_bytes;
          uint8_t* stg_bf16 = stg_f32 + (use_f32_s ? kStgF32Bytes : 0);
          mbar_wait_sleep(&stg_full[slot], uint32_t(cc >> 1) & 1u, 200);
          if (args.dbg_mode & 64) { mbar_arrive(&stg_empty[slot]); continue; }
          if (flags & kConvOutF32) {
            tma_store_4d(&maps.of, stg_f32, oc * 64, x0, y0, n);
            tma_store_4d(&maps.of, stg_f32 + kABytes, oc * 64 + 32, x0, y0, n);
          }
          if (flags & kConvOutBf16) {
            const int mi = oc / args.o_chunks_per_map;
            const int c0 = (oc % args.o_chunks_per_map) * 64;
            tma_store_4d(&maps.ob[mi], stg_bf16, c0, x0, y0, n);
          }
          tma_store_commit();
          tma_store_wait_read0();          // this chunk's store has read its slot: the epilogue may refill it (chunk
          mbar_arrive(&stg_empty[slot]);   // cc+2, two tiles from now -- releasing cc-1 here instead came one store
                                           // issue too late for chunk cc+1 and stalled the accumulator drain)
        }
      }
      tma_store_wait_all0();
    }
  } else {
    // ===================================================================== epilogue (2 sets x 2 groups x 128 threads)
    // Set tp (warps 2+8tp .. 9+8tp) takes the tiles of parity tp = accumulator tp = staging slot tp, so that two
    // tiles' epilogue chains (shared-memory round trips that queue behind the tensor core's operand reads) overlap;
    // inside a set, group g owns channels [32g, 32g+32) of every 64-channel chunk of every pixel row.  Configurations
    // with TMA inputs (residual / mask), one staging slot or several chunks per tile run on set 0 alone.
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;               // pixel row of the tile == TMEM lane
    const int tp = (warp - 2) >> 3;              // epilogue set
    const int grp = ((warp - 2) >> 2) & 1;       // channel half
    const int et = ((warp - 2) & 7) * 32 + lane; // 0..255 inside the set; et==0 issues TMA in the non-offload configurations
    const int ly = RESIDENT_B ? row >> 3 : row >> 4, lx = RESIDENT_B ? row & 7 : row & 15;
    for (int i = et + tp * kEpiThreads; i < BN; i += kEpiSets * kEpiThreads)
      bias_s[i] = args.bias ? args.bias[n_tile * BN + i] : 0.f;
    named_bar_sync(1, kEpiSets * kEpiThreads);
    grid_dep_wait();   // residual / mask inputs and every global write must follow the previous kernel
    int it = 0;
    if constexpr (BN == 16) {
      // thin tail conv (C -> out_feats <= 16): fp32 NCHW written straight from registers, no staging; group 0 only
      if (grp == 0 && tp == 0)
      for (int mt = mt_first; mt < args.m_tiles; mt += mt_stride, ++it) {
        const int n = mt / tiles_per_img;
        const int rem = mt - n * tiles_per_img;
        const int y = (rem / args.tiles_x) * TH + ly;
        const int x = (rem % args.tiles_x) * TW + lx;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (y < args.H && x < args.W) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < args.cout_real)
              args.out_nchw[((size_t(n) * args.cout_real + c) * args.H + y) * args.W + x] =
                  (__uint_as_float(v[c]) + bias_s[c]) * args.alpha;
        }
      }
    } else {
      const bool has_in = (flags & (kConvResF32 | kConvMask)) != 0;
      const bool use_f32 = (flags & (kConvOutF32 | kConvResF32)) != 0;
      const bool offload = !has_in && stg_bufs == 2;   // the store warp issues the stores (see above)
      const bool split = kEpiSets == 2 && offload && kChunksPerTile == 1;   // both sets work, one tile parity each
      const uint32_t bar_a = split ? 2u + 2u * tp : 2u, bar_b = bar_a + 1;   // the set's named barriers
      const uint32_t swz = uint32_t(row & 7);
      int cc = 0;  // running chunk counter -> staging slot
      const int my_tiles = (mt_first < args.m_tiles) ? (args.m_tiles - 1 - mt_first) / mt_stride + 1 : 0;
      const int total_chunks = my_tiles * kChunksPerTile;
      // TMA the residual / mask tiles of chunk `cidx` into its staging slot (consumed in place by the epilogue)
      auto issue_inputs = [&](int cidx) {
        const int t_mt = mt_first + (cidx / kChunksPerTile) * mt_stride;
        const int t_oc = n_tile * kChunksPerTile + (cidx % kChunksPerTile);
        const int t_n = t_mt / tiles_per_img;
        const int t_rem = t_mt - t_n * tiles_per_img;
        const int t_y0 = (t_rem / args.tiles_x) * TH, t_x0 = (t_rem % args.tiles_x) * TW;
        const int slot = (stg_bufs == 2) ? (cidx & 1) : 0;
        uint8_t* sf = smem + slot * stg_buf_bytes;
        uint8_t* sb = sf + (use_f32 ? kStgF32Bytes : 0);
        uint32_t bytes = 0;
        if (flags & kConvResF32) bytes += kStgF32Bytes;
        if (flags & kConvMask) bytes += kStgBf16Bytes;
        mbar_expect_tx(&in_bar[slot], bytes);
        if (flags & kConvResF32) {
          tma_load_4d(sf, &maps.rf, &in_bar[slot], t_oc * 64, t_x0, t_y0, t_n);
          tma_load_4d(sf + kABytes, &maps.rf, &in_bar[slot], t_oc * 64 + 32, t_x0, t_y0, t_n);
        }
        if (flags & kConvMask) tma_load_4d(sb, &maps.mb, &in_bar[slot], t_oc * 64, t_x0, t_y0, t_n);
      };
      if (has_in && stg_bufs == 2 && et == 0 && tp == 0 && total_chunks > 0) issue_inputs(0);
      if (split || tp == 0)
      for (int mt = mt_first; mt < args.m_tiles; mt += mt_stride, ++it) {
        if (split) {
          if ((it & 1) != tp) continue;
          cc = it;   // one chunk per tile: the global chunk index (staging slot = tile parity)
        }
        const int n = mt / tiles_per_img;
        const int rem = mt - n * tiles_per_img;
        const int y0 = (rem / args.tiles_x) * TH;
        const int x0 = (rem % args.tiles_x) * TW;
        const bool valid = (y0 + ly < args.H) && (x0 + lx < args.W);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        for (int j = 0; j < kChunksPerTile; ++j, ++cc) {
          const int oc = n_tile * kChunksPerTile + j;  // 64-channel output chunk index
          const int slot = (stg_bufs == 2) ? (cc & 1) : 0;
          uint8_t* stg = smem + slot * stg_buf_bytes;
          uint8_t* stg_f32 = stg;
          uint8_t* stg_bf16 = stg + (use_f32 ? kStgF32Bytes : 0);
          uint8_t* my_f32 = stg_f32 + grp * kABytes + row * 128;   // this group's 32-channel fp32 half
          uint8_t* my_bf16 = stg_bf16 + row * 128;
          RB_STAMP2(0);
          if (has_in && stg_bufs == 1 && et == 0) {
            // single slot: wait until the previous chunk's stores have read it, then fetch this chunk's inputs
            tma_store_wait_read0();
            issue_inputs(cc);
          }
          if (!has_in && stg_bufs == 1 && cc > 0) {
            // single slot without inputs (two-CTAs-per-SM configuration): the previous chunk's store must have
            // read the slot before anyone overwrites it
            if (et == 0) tma_store_wait_read0();
            named_bar_sync(1, kEpiThreads);
          }
          if (j == 0) {
            mbar_wait_sleep(&tmem_full_bar[acc], acc_phase, 100);
            tc_fence_after();
            if (et == 0 && tp == 0) RB_STAMP(it == 0 ? 6 : 9);
          }
          RB_STAMP2(1);
          uint32_t v[32];
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + j * 64 + grp * 32), v);
          tmem_ld_wait();
          RB_STAMP2(2);
          if (j == kChunksPerTile - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);  // accumulator drained -> MMA may reuse it
          }
          if (has_in) mbar_wait(&in_bar[slot], uint32_t(cc / stg_bufs) & 1u);
          // (after the accumulator was handed back: a late slot must not hold up the MMA warp)
          if (offload && cc >= 2) mbar_wait(&stg_empty[slot], uint32_t((cc >> 1) - 1) & 1u);
          if (!(args.dbg_mode & 16)) {
            float f[32];
            const float4* b4 = reinterpret_cast<const float4*>(&bias_s[j * 64 + grp * 32]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = b4[i];
              const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x = __uint_as_float(v[i * 4 + e]) + bb[e];
                if (flags & kConvRelu) x = fmaxf(x, 0.f);
                f[i * 4 + e] = x * args.alpha;
              }
            }
            if (args.ch_scale != nullptr) {
              const float* cs = args.ch_scale + size_t(n) * args.cout + oc * 64 + grp * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] *= __ldg(cs + i);
            }
            if (flags & kConvMask) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 m = *reinterpret_cast<const uint4*>(my_bf16 + (((uint32_t(grp * 4 + c)) ^ swz) << 4));
                const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                  const uint32_t lo = mw[e] & 0xFFFFu, hi = mw[e] >> 16;
                  if (!(lo != 0 && lo < 0x8000u)) f[c * 8 + e * 2] = 0.f;
                  if (!(hi != 0 && hi < 0x8000u)) f[c * 8 + e * 2 + 1] = 0.f;
                }
              }
            }
            if (flags & kConvResF32) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 r = *reinterpret_cast<const float4*>(my_f32 + ((uint32_t(c) ^ swz) << 4));
                f[c * 4 + 0] += r.x; f[c * 4 + 1] += r.y; f[c * 4 + 2] += r.z; f[c * 4 + 3] += r.w;
              }
            }
            if (!valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = 0.f;
            }
            if (flags & kConvOutF32) {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(my_f32 + ((uint32_t(c) ^ swz) << 4)) =
                    make_float4(f[c * 4], f[c * 4 + 1], f[c * 4 + 2], f[c * 4 + 3]);
            }
            if ((flags & kConvOutBf16) && args.bf16_scale != nullptr) {
              const float* bs = args.bf16_scale + size_t(n) * args.cout + oc * 64 + grp * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] *= __ldg(bs + i);
            }
            if (flags & kConvOutBf16) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(my_bf16 + ((uint32_t(grp * 4 + c) ^ swz) << 4)) =
                    make_uint4(pack_bf16x2(f[c * 8], f[c * 8 + 1]), pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]),
                               pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]));
            }
          }
          RB_STAMP2(3);
          fence_proxy_async_smem();
          if (offload) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_full[slot]);   // -> store warp
            RB_STAMP2(4);
            if (flags & kConvPool) named_bar_sync(bar_a, kEpiThreads);   // the pool below reads other threads' rows
            RB_STAMP2(5);
          } else {
            // double-buffered staging: the store issued one chunk ago must have finished READING the other
            // buffer before anyone refills it in the next chunk -- checked here, a whole chunk later
            if (!has_in && stg_bufs == 2 && et == 0) tma_store_wait_read0();
            RB_STAMP2(4);
            named_bar_sync(2, kEpiThreads);
            RB_STAMP2(5);
            if (et == 0) {
              RB_STAMP(it == 0 ? 8 : 10);
              if (flags & kConvOutF32) {
                tma_store_4d(&maps.of, stg_f32, oc * 64, x0, y0, n);
                tma_store_4d(&maps.of, stg_f32 + kABytes, oc * 64 + 32, x0, y0, n);
              }
              if (flags & kConvOutBf16) {
                const int mi = oc / args.o_chunks_per_map;
                const int c0 = (oc % args.o_chunks_per_map) * 64;
                tma_store_4d(&maps.ob[mi], stg_bf16, c0, x0, y0, n);
              }
              tma_store_commit();
              if (has_in && stg_bufs == 2 && cc + 1 < total_chunks) {
                // prefetch the next chunk's inputs into the other slot once its previous store (chunk cc-1) has
                // been read out; the store just committed (this chunk) may stay in flight
                tma_store_wait_read1();
                issue_inputs(cc + 1);
              }
            }
          }
          RB_STAMP2(6);
          if ((flags & kConvPool) && !(args.dbg_mode & 32)) {
            // Channel sums over the tile's valid pixels (invalid rows were zeroed), two 64-row halves.  Thread t sums
            // 8 channels (one 16-byte unit per row) over rows 4*(t>>3) .. +3, lanes with the same unit combine by
            // shuffle, the 8 warps (16 rows each) through shared memory.
            const int c8 = et & 7, r0 = (et >> 3) * 4;
            float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (flags & kConvOutF32) {
              // fp32 staging: two 32-channel halves; s[0..3] = channels 4*c8.., s[4..7] = channels 32 + 4*c8..
#pragma unroll
              for (int r = r0; r < r0 + 4; ++r) {
                const uint32_t off = uint32_t(r) * 128 + ((uint32_t(c8) ^ uint32_t(r & 7)) << 4);
                const float4 a = *reinterpret_cast<const float4*>(stg_f32 + off);
                const float4 b = *reinterpret_cast<const float4*>(stg_f32 + kABytes + off);
                s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
                s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
              }
            } else {
#pragma unroll
              for (int r = r0; r < r0 + 4; ++r) {
                const uint4 a = *reinterpret_cast<const uint4*>(stg_bf16 + uint32_t(r) * 128 +
                                                                ((uint32_t(c8) ^ uint32_t(r & 7)) << 4));
                const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  s[2 * e] += __uint_as_float(w[e] << 16);
                  s[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              s[e] += __shfl_xor_sync(0xffffffffu, s[e], 8);
              s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);
            }
            RB_STAMP2(8);
            const int ew = tp * 8 + (et >> 5);   // the set's warp 0..7 = rows 16*w .. 16*w+15
            if (lane < 8) {
              if (flags & kConvOutF32) {
                *reinterpret_cast<float4*>(&pool_s[ew][4 * c8]) = make_float4(s[0], s[1], s[2], s[3]);
                *reinterpret_cast<float4*>(&pool_s[ew][32 + 4 * c8]) = make_float4(s[4], s[5], s[6], s[7]);
              } else {
                *reinterpret_cast<float4*>(&pool_s[ew][8 * c8]) = make_float4(s[0], s[1], s[2], s[3]);
                *reinterpret_cast<float4*>(&pool_s[ew][8 * c8 + 4]) = make_float4(s[4], s[5], s[6], s[7]);
              }
            }
            named_bar_sync(bar_b, kEpiThreads);   // (also: every reader of the slot is done before the next input TMA)
            RB_STAMP2(9);
            if (et < 128) {
              const int c = et & 63, half = et >> 6;
              const int w0 = tp * 8 + 4 * half;
              const float t = (pool_s[w0][c] + pool_s[w0 + 1][c]) + (pool_s[w0 + 2][c] + pool_s[w0 + 3][c]);
              args.pool_partial[(size_t(mt) * 2 + half) * args.cout + oc * 64 + c] = t;
            }
          }
          RB_STAMP2(7);
        }
      }
      if (!offload && et == 0 && tp == 0) { RB_STAMP(11); tma_store_wait_all0(); RB_STAMP(12); }
      if (offload && et == 0 && tp == 0) RB_STAMP(11);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) RB_STAMP(13);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
    if (lane == 0) RB_STAMP_NS(15);
  }
}

}  // namespace rb
