// The 64-channel residual trunk, ROLE-SWAPPED: weights are the tcgen05 A operand and live in TENSOR MEMORY, the
// resident activations are the B operand (N = up to 144 pixels per MMA).  One thread-block cluster per image, each
// CTA owns a BAND of full-width image rows.
//
// Same layer program and arithmetic as trunk_cluster.cuh / trunk_pipe.cuh (reference call sites: common.py:6-9,
// architectures.py:41-44 / 81-84 / 121-124 / 172-174, common.py:71-75).  What changes is the MMA shape:
//   * The old kernels compute D[128 pixels][64 c_out] with both operands in shared memory: 6 KB of operand reads per
//     32-cycle MMA, so the 128 B/clk shared-memory port (not the tensor pipe) bounds them at 48 (67-68 measured)
//     cycles per MMA.  Here D^T[c_out][pixel] = W[c_out][c_in] . X^T[c_in][pixel]: A = the layer's weights, read
//     from TMEM (no shared-memory traffic), B = N pixels x 16 channels = N * 32 B per MMA (64 B/clk at the tensor
//     floor of N/2 cycles).  Measured (tools/experiments/r02_probe2.cu): 65 / 89 / 129 cycles at N = 128 / 176 / 256.
//   * M must be 128 for full rate while c_out is 64, so TWO TAPS ARE STACKED ALONG M: rows 0-63 = W(ky, kx=0),
//     rows 64-127 = W(ky, kx=1); both multiply the same activation window, so accumulator lane c holds tap
//     (ky,0)'s term of output pixel j in column j and lane 64+c holds tap (ky,1)'s term of pixel j in column j+1.
//     The kx=2 taps run as half-empty blocks (rows 64-127 zero): 6 MMAs per K step instead of 4.5 (75 %), ONE
//     accumulator of N columns per chunk.  (tools/experiments/r02_probe.cu checks this bookkeeping bit-exactly on the
//     device, tools/experiments/stacked_tap_conv_check.py in numpy.)
//   * Activations: bf16, channel-planar, un-swizzled: 8 planes (8 channels = 16 B per pixel cell) over the band's
//     halo-padded pixel grid, addressed LINEARLY (pitch P = W + 1: the right pad of a row is the left pad of the
//     next one; both are the conv's zero padding).  The B operand of tap offset o for the chunk of output pixels
//     [q0, q0 + nq) is the same descriptor with its start address moved by (q0 + o) cells (SBO = 128 B = 8 cells,
//     LBO = plane stride).  Pad cells are never written and stay zero.
//   * Epilogue thread = one accumulator lane = one output channel, registers = pixels.  Lanes c and 64 + c meet
//     through a small shared-memory exchange; the channel-attention pool is a thread-local sum; bf16 results are
//     written as 2-byte stores into the next layer's planes and (first / last band row) as 4-byte st.async into
//     the neighbour CTA's halo row.
//   * The fp32 residual stream is kept as (bf16 operand already in shared memory) + (fp16 remainder in registers):
//     x = hi + lo with |error| <= 2^-20 |x|.
//   * Weights: 9 TMA boxes of 8 KB per layer into 5 staging slots (SW128 K-major), then tcgen05.cp into TMEM, issued
//     by the MMA thread right after the last MMAs of the previous layer that read the block (tcgen05.mma and
//     tcgen05.cp execute in issue order: no barrier).
// TMEM: accumulators at columns [0, 144) and [144, 288) (double buffered over chunks), weights at [296, 488).
// Warps: 0-11 epilogue (lane quadrant = warp & 3, column third = warp >> 2) | 12 MMA issuer | 13 weight producer |
// 14, 15 idle (they complete the fourth warpgroup, which hands its registers to the epilogue: setmaxnreg 56 / 152).
#pragma once
#include "trunk_cluster.cuh"
#include <cuda_fp16.h>
#include <type_traits>

namespace rb {

constexpr int kBandEpiWarps = 12;
constexpr int kBandThreads = (kBandEpiWarps + 4) * 32;   // 512: three epilogue warpgroups + one service warpgroup
constexpr int kBandEpiRegs = 152, kBandSvcRegs = 56;     // setmaxnreg split of the 64 K registers (launch: 128 each)
constexpr int kBandMaxChunks = 3;
constexpr int kBandKeep = 24;            // pixels finalised per epilogue thread and chunk (static register arrays)
constexpr int kBandNQMax = 143;          // output pixels per chunk (MMA N <= 144, one spare column for the shift)
constexpr int kBandSlots = 5;            // weight staging slots
constexpr int kBandSlotBytes = 16384;    // one stacked block: 128 rows x 64 k bf16
constexpr int kBandXchgBytes = kBandEpiWarps * 8 * 32 * 4;   // 12 KB: 8 values per thread and round
constexpr uint32_t kBandAccStride = 144, kBandWCol = 296;   // columns [288, 296): guard (see the CA write-back)
constexpr int kBandMaxCluster = 8;

struct BandArgs {
  const TrunkLayer* layers;
  const float* s_init;             // fp32 NHWC: initial residual stream (head conv output)
  __nv_bfloat16* out_bf16;         // bf16 NHWC: output of the last layer
  long long* dbg;                  // optional timeline [grid][dbg_layers][16]
  int n_layers, n_ca, N, H, W, C, R, P, n_chunks, cr, dbg_layers;
  int nq[kBandMaxChunks], nn[kBandMaxChunks];   // output pixels / MMA N (multiple of 16, > nq) per chunk
  uint32_t plane_bytes;            // plane stride (= 16 mod 128: the 2-byte epilogue stores are conflict-free)
  uint32_t pmagic;                 // ceil(2^24 / P): q / P == (q * pmagic) >> 24 for q * P < 2^24
  float inv_hw;
};

__host__ __device__ inline uint32_t band_plane_bytes(int R, int P) {
  uint32_t b = uint32_t((R + 2) * P + 1) * 16u;
  while ((b & 127u) != 16u) b += 16u;
  return b;
}
// dynamic smem: [staging 5 x 16 KB | buffer 0 | buffer 1 | exchange 12 KB | pool slots 2 x C x 64 floats]
__host__ __device__ inline size_t band_smem_bytes(int R, int P, int C) {
  return 1024 + size_t(kBandSlots) * kBandSlotBytes + 2 * size_t(8) * band_plane_bytes(R, P) + kBandXchgBytes +
         size_t(2) * C * 64 * sizeof(float);
}

#ifdef RB_TRUNK_KERNEL_IMPL

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 128 lanes x 256 bit (= one K = 16 slice of a 128-row bf16 A operand) shared memory -> tensor memory
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// 24 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_ld24(uint32_t taddr, uint32_t (&v)[24]) {
  uint32_t a[16];
  tmem_ld16(taddr, a);
  tmem_ld8(taddr + 16, v + 16);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = a[i];
}
__device__ __forceinline__ void tmem_st24(uint32_t taddr, const uint32_t (&v)[24]) {
  uint32_t a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = v[i];
  tmem_st16(taddr, a);
  tmem_st8(taddr + 16, v + 16);
}
// explicit shared-space accesses (the 1024-byte alignment cast of the dynamic shared memory base makes plain pointer
// accesses generic)
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(uint16_t(v)) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr) : "memory");
  return uint32_t(v);
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds_v4(uint32_t saddr, float& a, float& b, float& c, float& d) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(saddr) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t bf16_bits(float f) {
  return uint32_t(__bfloat16_as_ushort(__float2bfloat16_rn(f)));
}
__device__ __forceinline__ float bf16_from_bits(uint32_t b) { return __uint_as_float(b << 16); }

// mbarrier waits for this kernel: the watchdog traps inline.  (ptx.cuh's waits call a __noinline__ printf helper;
// a call site inside the layer loop makes ptxas keep every live value in callee-saved registers or spill it -- with
// ~100 live registers per epilogue thread that cost 1 KB of spills and, at 20 % L1 hit rate, most of the run time.)
__device__ __forceinline__ void band_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > RB_WATCHDOG_CYCLES) asm volatile("trap;");
  }
}
__device__ __forceinline__ void band_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > RB_WATCHDOG_CYCLES) asm volatile("trap;");
  }
}

// block order of a chunk: single(ky=0) | pair(ky=0) | pair(ky=1) | pair(ky=2) | single(ky=1) | single(ky=2); the two
// singles that share staging slot 3 are five blocks apart, so the second one's TMA has time to land
__device__ __forceinline__ int band_slot(int bi) { return bi == 0 ? 3 : (bi <= 3 ? bi - 1 : (bi == 4 ? 4 : 3)); }
__device__ __forceinline__ int band_off(int bi, int P) {   // linear offset of the block's lower-half tap
  return bi == 0 ? -P + 1 : (bi == 1 ? -P - 1 : (bi == 2 ? -1 : (bi == 3 ? P - 1 : (bi == 4 ? 1 : P + 1))));
}
// k-th fill / consumption of the block's staging slot (slot 3 is used twice per layer)
__device__ __forceinline__ uint32_t band_fill(int L, int bi) {
  return (bi == 0 || bi == 5) ? uint32_t(2 * L + (bi == 5)) : uint32_t(L);
}

__global__ void __launch_bounds__(kBandThreads, 1)
trunk_band_kernel(const __grid_constant__ CUtensorMap w_map, const BandArgs args) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full[kBandSlots];
  __shared__ __align__(8) uint64_t w_empty[kBandSlots];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ __align__(8) uint64_t in_full[2];
  __shared__ __align__(8) uint64_t pool_full[2];
  __shared__ uint32_t tmem_base_s, halo_bytes_s;
  __shared__ float red_s[6][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpMma = kBandEpiWarps, kWarpW = kBandEpiWarps + 1;
  const int n_layers = args.n_layers, n_chunks = args.n_chunks;
  const int C = args.C, R = args.R, P = args.P;
  const int rank = int(cluster_ctarank());
  const int n = blockIdx.x / C;
  const int rows_me = min(R, args.H - rank * R);          // >= 1 (host guarantees)
  const uint32_t plane = args.plane_bytes, buf_bytes = 8 * plane;
  const int q_first = P + 1;                              // linear cell of pixel (row 0, col 0)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_s = smem;
  uint8_t* buf0 = smem + kBandSlots * kBandSlotBytes;
  float* xchg_s = reinterpret_cast<float*>(buf0 + 2 * buf_bytes);
  float* pool_s = xchg_s + kBandXchgBytes / 4;            // [2][C][64]

#define BD_STAMP(L_, slot_)                                                                              \
  do {                                                                                                   \
    if (args.dbg && (L_) < args.dbg_layers)                                                              \
      args.dbg[(size_t(blockIdx.x) * args.dbg_layers + (L_)) * 16 + (slot_)] = clock64();                \
  } while (0)

  // ---- zero both activation buffers, the exchange area and the staging slots (upper halves of slots 3 / 4 are the
  // permanent zero rows of the single-tap blocks)
  {
    const uint32_t total = kBandSlots * kBandSlotBytes + 2 * buf_bytes + kBandXchgBytes;
    for (uint32_t i = threadIdx.x * 16; i < total; i += kBandThreads * 16)
      *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    int halo = 0;                                         // halo pixels this CTA is owed per layer
    if (rank > 0) halo += args.W;
    if (rank < C - 1) halo += args.W;
    for (int i = 0; i < kBandSlots; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kBandEpiWarps);
      mbar_init(&in_full[i], kBandEpiWarps + 1);          // epilogue warps + the arming arrival; halos complete tx bytes
      mbar_init(&pool_full[i], 1);
    }
    halo_bytes_s = uint32_t(halo) * 128u;
    fence_mbar_init();
    for (int i = 0; i < 2; ++i) {
      mbar_expect_tx(&in_full[i], uint32_t(halo) * 128u);
      mbar_expect_tx(&pool_full[i], uint32_t(C) * 64u * 4u);
    }
  }
  if (warp == kWarpW && lane == 0) tma_prefetch_desc(&w_map);
  if (warp == kWarpMma) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();                               // the zeroed staging halves are read by tcgen05.cp
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // every CTA's buffers are zeroed and its barriers initialised before any remote access
  const uint32_t tmem_base = tmem_base_s;

  if (warp >= kBandEpiWarps) {
  setmaxnreg_dec<kBandSvcRegs>();
  if (warp == kWarpMma) {
    // ===================================================================== MMA issuer
    const uint32_t halo_bytes = halo_bytes_s;
    const uint64_t bdesc0 = make_smem_desc(0, plane, 128, 0);
    const uint64_t wdesc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    const uint32_t stage16 = (smem_u32(stage_s) & 0x3FFFF) >> 4;
    const uint32_t kstep = (2 * plane) >> 4;              // K = 16 channels = two planes
    // weights of (layer Lw, block bi): staging slot -> TMEM columns [288 + 32 bi, +32)
    auto copy_block = [&](int Lw, int bi) {
      const int slot = band_slot(bi);
      band_wait(&w_full[slot], band_fill(Lw, bi) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t d = wdesc0 + uint64_t(stage16 + uint32_t(slot) * (kBandSlotBytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tmem_base + kBandWCol + uint32_t(bi * 32 + k * 8), d + uint64_t(2 * k));
        umma_commit(&w_empty[slot]);                      // the slot may be refilled once the copies have read it
      }
      __syncwarp();
    };
    for (int bi = 0; bi < 6; ++bi) copy_block(0, bi);
    uint32_t use = 0;                                     // running chunk counter: accumulator buffer = use & 1
#pragma unroll 1
    for (int L = 0; L < n_layers; ++L) {
      // own pixels of layer L-1 and both halo rows have landed in buffer L & 1
      band_wait_cluster(&in_full[L & 1], uint32_t(L >> 1) & 1u);
      if (lane == 0 && L + 2 < n_layers) mbar_expect_tx(&in_full[L & 1], halo_bytes);   // arm layer L+2's phase
      fence_proxy_async_smem();   // generic-proxy writes of the epilogue threads -> async-proxy (tensor core) reads
      tc_fence_after();
      if (lane == 0) BD_STAMP(L, 0);
      const uint32_t abuf16 = (smem_u32(buf0 + (L & 1) * buf_bytes) & 0x3FFFF) >> 4;
      int q0 = q_first;
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch, ++use) {
        const uint32_t buf = use & 1u;
        band_wait(&acc_empty[buf], ((use >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0) BD_STAMP(L, 8 + ch);
        const uint32_t d_tmem = tmem_base + buf * kBandAccStride;
        const uint32_t idesc = make_idesc_bf16(128, uint32_t(ch == 0 ? args.nn[0] : (ch == 1 ? args.nn[1] : args.nn[2])));
        const bool refill = ch == n_chunks - 1 && L + 1 < n_layers;
        for (int bi = 0; bi < 6; ++bi) {
          if (elect_one()) {
            const uint64_t bdesc = bdesc0 + uint64_t(abuf16 + uint32_t(q0 + band_off(bi, P)));
            const uint32_t a_tmem = tmem_base + kBandWCol + uint32_t(bi * 32);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(d_tmem, a_tmem + uint32_t(8 * k), bdesc + uint64_t(k * kstep), idesc, (bi | k) != 0);
          }
          __syncwarp();
          // Next layer's weights.  tcgen05.cp executes in issue order behind the MMAs, and an MMA behind a copy waits
          // for it: every copy between two MMAs drains the tensor pipe.  So block 0 is copied right after its last
          // MMAs (one drain; staging slot 3 is then refilled with block 5 while the other blocks run) and blocks 1-5
          // after the chunk's last MMA, where the copies overlap the epilogue tail.
          if (refill && bi == 0) {
            if (lane == 0) BD_STAMP(L, 14);
            copy_block(L + 1, 0);
            if (lane == 0) BD_STAMP(L, 15);
          }
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
        __syncwarp();
        if (refill) {
#pragma unroll 1
          for (int bi = 1; bi < 6; ++bi) copy_block(L + 1, bi);
        }
        if (lane == 0) BD_STAMP(L, 5 + ch);
        q0 += ch == 0 ? args.nq[0] : (ch == 1 ? args.nq[1] : args.nq[2]);   // no dynamic index: the parameter struct
                                                                            // would be copied to local memory
      }
      if (lane == 0) BD_STAMP(L, 1);
    }
  } else if (warp == kWarpW) {
    // ===================================================================== weight producer: 9 tap boxes per layer
    for (int L = 0; L < n_layers; ++L) {
      for (int bi = 0; bi < 6; ++bi) {
        const int slot = band_slot(bi);
        band_wait(&w_empty[slot], (band_fill(L, bi) & 1u) ^ 1u);
        if (elect_one()) {
          uint8_t* dst = stage_s + slot * kBandSlotBytes;
          const bool pair = bi >= 1 && bi <= 3;
          const int ky = pair ? bi - 1 : (bi == 0 ? 0 : bi - 3);
          mbar_expect_tx(&w_full[slot], pair ? 16384u : 8192u);
          if (pair) {                                     // packed tap index = kx * 3 + ky
            tma_load_4d(dst, &w_map, &w_full[slot], 0, 0, ky, L);
            tma_load_4d(dst + 8192, &w_map, &w_full[slot], 0, 0, 3 + ky, L);
          } else {
            tma_load_4d(dst, &w_map, &w_full[slot], 0, 0, 6 + ky, L);
          }
        }
        __syncwarp();
      }
    }
  }
  } else {
    setmaxnreg_inc<kBandEpiRegs>();
    // ===================================================================== epilogue: 12 warps, thread = channel
    const int q = warp & 3, s = warp >> 2, half = q >> 1;
    const int c = (q & 1) * 32 + lane;                    // output channel of this thread
    const int et = threadIdx.x;                           // 0 .. 383
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t pair_bar = 1u + uint32_t(q & 1) + 2u * uint32_t(s);   // the two warps holding lanes c and 64 + c
    // exchange area: [warp][2 slots][32 lanes][16 B]
    const uint32_t xb_mine = smem_u32(xchg_s) + uint32_t(warp * 64 + lane) * 16;
    const uint32_t xb_partner = smem_u32(xchg_s) + uint32_t((warp ^ 2) * 64 + lane) * 16;
    const uint32_t buf0_a = smem_u32(buf0);
    const uint32_t ch_byte = uint32_t(c >> 3) * plane + uint32_t(c & 7) * 2;   // plane + position of channel c in a cell

    // ---- per chunk: this thread finalises pixels j = 48 s + 24 half + i, i < 24, of the chunk's nq.  vm = valid
    // pixels, pad = index of the pad cell among the 24 (24: none).  All per-pixel conditions are warp-uniform.
    // Chunk loops are NOT unrolled (code size); per-chunk values are picked with selects.
    const int jbase = 48 * s + 24 * half;
    uint32_t vp0 = 24u << 24, vp1 = 24u << 24, vp2 = 24u << 24;   // bits 0-23: valid mask, bits 24-28: pad index
    {
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int nq = ch == 0 ? args.nq[0] : (ch == 1 ? args.nq[1] : args.nq[2]);
        const int qb = q_first + kBandNQMax * ch + jbase;  // every chunk but the last has kBandNQMax pixels
        uint32_t vm = 0, pad = 24;
#pragma unroll 1
        for (int i = 0; i < kBandKeep; ++i) {
          const int j = jbase + i, ql = qb + i;
          const int row = int((uint32_t(ql) * args.pmagic) >> 24) - 1, col = ql - (row + 1) * P - 1;
          if (j < nq && row < rows_me) {
            if (col >= 0) vm |= 1u << i; else pad = uint32_t(i);
          }
        }
        const uint32_t vp = vm | (pad << 24);
        if (ch == 0) vp0 = vp; else if (ch == 1) vp1 = vp; else vp2 = vp;
      }
    }
#define BD_SEL(ch_, a_, b_, c_) ((ch_) == 0 ? (a_) : ((ch_) == 1 ? (b_) : (c_)))
    // fp16 remainders of the fp32 residual stream, two pixels per register.  s0 always belongs to the chunk that is
    // processed next: the three arrays are rotated once per chunk (period n_chunks) instead of being indexed by the
    // chunk number, so the chunk loops stay rolled and every register array keeps static indices.
    uint32_t s0[kBandKeep / 2], s1[kBandKeep / 2], s2[kBandKeep / 2];
#pragma unroll
    for (int k = 0; k < kBandKeep / 2; ++k) s0[k] = s1[k] = s2[k] = 0;
    auto rotate1 = [&](int k, uint32_t nw) {              // element k of the rotation; nw = the processed chunk's new value
      const uint32_t a = s1[k], b = s2[k];
      s0[k] = n_chunks == 1 ? nw : a;
      s1[k] = n_chunks == 2 ? nw : b;
      s2[k] = nw;
    };
    float u0[kBandKeep];                                  // channel attention: chunk 0's conv output between the passes

    // f[24] (fp32) -> bf16 cells of the buffer at `ob` [+ fp16 remainders nw]; pad cells are re-zeroed by the thread
    // that overwrote them (program order), pixels outside the band are never written
    // f[24] (fp32) -> bf16 cells of the buffer at `ob`.  MODE 0: leave the stream alone, 1: rotate it unchanged (the
    // layer only read it), 2: replace it (fp16 remainders).  FULL: every pixel but (at most) one pad cell is valid: the
    // pad cell is re-zeroed by the thread that overwrote it (program order); otherwise stores are predicated.
    // The conditions are template parameters so that the unrolled pixel loop has no branches.
    auto store24t = [&](auto mode_c, auto full_c, uint32_t cell0, uint32_t vm, uint32_t pad, const float (&f)[kBandKeep]) {
      constexpr int MODE = decltype(mode_c)::value;
      constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
      for (int k = 0; k < kBandKeep / 2; ++k) {
        const uint32_t pr = pack_bf16x2(f[2 * k], f[2 * k + 1]);
        if (MODE == 2) {
          const __half2 lo = __floats2half2_rn(f[2 * k] - __uint_as_float(pr << 16),
                                               f[2 * k + 1] - __uint_as_float(pr & 0xffff0000u));
          rotate1(k, *reinterpret_cast<const uint32_t*>(&lo));
        } else if (MODE == 1) {
          rotate1(k, s0[k]);
        }
        if (FULL) {
          sts_u16(cell0 + uint32_t(2 * k) * 16, pr);
          sts_u16(cell0 + uint32_t(2 * k + 1) * 16, pr >> 16);
        } else {
          if ((vm >> (2 * k)) & 1u) sts_u16(cell0 + uint32_t(2 * k) * 16, pr);
          if ((vm >> (2 * k + 1)) & 1u) sts_u16(cell0 + uint32_t(2 * k + 1) * 16, pr >> 16);
        }
      }
      if (FULL && pad < 24) sts_u16(cell0 + pad * 16, 0u);
    };
    auto store24 = [&](uint32_t ob, int qb, uint32_t vm, uint32_t pad, const float (&f)[kBandKeep], int mode) {
      const uint32_t cell0 = ob + uint32_t(qb) * 16 + ch_byte;
      const bool full = (vm | (pad < 24 ? (1u << pad) : 0u)) == 0xFFFFFFu;
      using std::integral_constant;
      if (full) {
        if (mode == 2) store24t(integral_constant<int, 2>{}, integral_constant<bool, true>{}, cell0, vm, pad, f);
        else if (mode == 1) store24t(integral_constant<int, 1>{}, integral_constant<bool, true>{}, cell0, vm, pad, f);
        else store24t(integral_constant<int, 0>{}, integral_constant<bool, true>{}, cell0, vm, pad, f);
      } else {
        if (mode == 2) store24t(integral_constant<int, 2>{}, integral_constant<bool, false>{}, cell0, vm, pad, f);
        else if (mode == 1) store24t(integral_constant<int, 1>{}, integral_constant<bool, false>{}, cell0, vm, pad, f);
        else store24t(integral_constant<int, 0>{}, integral_constant<bool, false>{}, cell0, vm, pad, f);
      }
    };
    // the stream bookkeeping alone (last layer: nothing is written to shared memory)
    auto rotate_only = [&](int mode, const float (&f)[kBandKeep]) {
#pragma unroll
      for (int k = 0; k < kBandKeep / 2; ++k) {
        if (mode == 2) {
          const uint32_t pr = pack_bf16x2(f[2 * k], f[2 * k + 1]);
          const __half2 lo = __floats2half2_rn(f[2 * k] - __uint_as_float(pr << 16),
                                               f[2 * k + 1] - __uint_as_float(pr & 0xffff0000u));
          rotate1(k, *reinterpret_cast<const uint32_t*>(&lo));
        } else if (mode == 1) {
          rotate1(k, s0[k]);
        }
      }
    };
    // end of a layer's stores: first / last band row -> the neighbours' halo rows (16-byte st.async straight from the
    // finished buffer), then this warp's arrival on the local barrier
    auto layer_done = [&](int par) {
      fence_proxy_async_smem();
      named_bar_sync(7, kBandEpiWarps * 32);              // every cell of the layer's output is written
      const uint32_t ob = buf0_a + uint32_t(par) * buf_bytes;
      const int per_dir = args.W * 8;
      const int items = ((rank > 0) + (rank < C - 1)) * per_dir;
#pragma unroll 1
      for (int idx = et; idx < items; idx += kBandEpiWarps * 32) {
        const bool up = rank > 0 && idx < per_dir;
        const int rem = up ? idx : idx - (rank > 0 ? per_dir : 0);
        const int pl = rem & 7, x = rem >> 3;
        // up: my row 0 -> the upper CTA's row R;  down: my last row -> the lower CTA's row -1
        const uint32_t src = uint32_t(up ? P + 1 + x : rows_me * P + 1 + x);
        const uint32_t dst = up ? src + uint32_t(R * P) : uint32_t(1 + x);
        const uint32_t nbr = uint32_t(up ? rank - 1 : rank + 1);
        const uint4 v = lds_u4(ob + uint32_t(pl) * plane + src * 16);
        st_async_v4(mapa_u32(ob + uint32_t(pl) * plane + dst * 16, nbr), v, mapa_u32(smem_u32(&in_full[par]), nbr));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_full[par]);
    };
    // global NHWC element index of channel c of this thread's first pixel of a chunk, and its column; consecutive
    // non-pad cells are consecutive pixels (full-width bands), so `walk` steps both
    auto first_pix = [&](int qb, int& col) -> size_t {
      const int row = int((uint32_t(qb) * args.pmagic) >> 24) - 1;   // qb / P (exact for qb * P < 2^24)
      col = qb - (row + 1) * P - 1;
      return ((size_t(n) * args.H + rank * R + row) * args.W + col) * 64 + c;
    };
    auto walk = [&](int& col, size_t& pix) { if (col == args.W - 1) col = -1; else { ++col; pix += 64; } };

    // ---- residual stream (fp32 -> bf16 operand in buffer 0 + fp16 remainder in registers)
#pragma unroll 1
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int qb = q_first + kBandNQMax * ch + jbase;
      const uint32_t vp = BD_SEL(ch, vp0, vp1, vp2), vm = vp & 0xFFFFFFu, pad = vp >> 24;
      int col;
      size_t pix = first_pix(qb, col);
      float f[kBandKeep];
#pragma unroll
      for (int i = 0; i < kBandKeep; ++i) {
        f[i] = ((vm >> i) & 1u) ? __ldg(args.s_init + pix) : 0.f;
        walk(col, pix);
      }
      store24(buf0_a, qb, vm, pad, f, 2);
    }
    layer_done(0);

    uint32_t use = 0;
    int ca_seen = 0;
#pragma unroll 1
    for (int L = 0; L < n_layers; ++L) {
      const TrunkLayer* lay = args.layers + L;
      const int kind = lay->kind;
      const float bias_c = __ldg(lay->bias + c);
      const int par_out = (L + 1) & 1;
      const bool last = L == n_layers - 1;
      const uint32_t ob = buf0_a + uint32_t(par_out) * buf_bytes;

      // conv output (+ bias) of this thread's 24 pixels of chunk ch; lanes c and 64 + c swap the halves they do not
      // finalise through shared memory, 8 values per round; the next round's columns are loaded meanwhile
      // conv output (+ bias) of this thread's 24 pixels of chunk ch, 8 per round: lanes c and 64 + c swap the halves
      // they do not finalise through shared memory; the next round's columns are loaded meanwhile
      auto rounds = [&](auto pipelined_c, int ch, uint32_t buf, auto&& consume) {
        constexpr bool PIPE = decltype(pipelined_c)::value;   // load the next round's columns during the exchange
        band_wait(&acc_full[buf], (use >> 1) & 1u);
        tc_fence_after();
        // columns: pixel j sits in column j of lanes 0-63 and in column j + 1 of lanes 64-127
        const uint32_t col_keep = lane_addr + buf * kBandAccStride + uint32_t(jbase + half);
        const uint32_t col_send = lane_addr + buf * kBandAccStride + uint32_t(jbase + (half ? -24 : 24) + half);
        const int nq = ch == 0 ? args.nq[0] : (ch == 1 ? args.nq[1] : args.nq[2]);
        uint32_t sv[PIPE ? 2 : 1][8], kv[PIPE ? 2 : 1][8];
        if (PIPE) { tmem_ld8(col_send, sv[0]); tmem_ld8(col_keep, kv[0]); }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (!PIPE) { tmem_ld8(col_send + 8 * r, sv[0]); tmem_ld8(col_keep + 8 * r, kv[0]); }
          tmem_ld_wait();
          if (PIPE && r < 2) {
            tmem_ld8(col_send + 8 * (r + 1), sv[(r + 1) & 1]);
            tmem_ld8(col_keep + 8 * (r + 1), kv[(r + 1) & 1]);
          }
          float p8[8];
          if (48 * s + 8 * r < nq) {                      // else neither warp of the pair has a pixel in this round
            sts_v4(xb_mine, sv[PIPE ? (r & 1) : 0][0], sv[PIPE ? (r & 1) : 0][1], sv[PIPE ? (r & 1) : 0][2], sv[PIPE ? (r & 1) : 0][3]);
            sts_v4(xb_mine + 512, sv[PIPE ? (r & 1) : 0][4], sv[PIPE ? (r & 1) : 0][5], sv[PIPE ? (r & 1) : 0][6], sv[PIPE ? (r & 1) : 0][7]);
            named_bar_sync(pair_bar, 64);
            float o[8];
            lds_v4(xb_partner, o[0], o[1], o[2], o[3]);
            lds_v4(xb_partner + 512, o[4], o[5], o[6], o[7]);
#pragma unroll
            for (int i = 0; i < 8; ++i) p8[i] = (__uint_as_float(kv[PIPE ? (r & 1) : 0][i]) + o[i]) + bias_c;
            named_bar_sync(pair_bar, 64);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) p8[i] = 0.f;
          }
          consume(r, p8);
        }
      };
      auto release_acc = [&](uint32_t buf) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      };

      // ---------------------------------------------------------------- channel attention, first pass: pool, y
      const bool is_ca = kind == kTrunkCA;
      const uint32_t use0 = use;
      float scale = lay->alpha;                           // multiplier of (acc + bias): alpha | CA vector y
      if (is_ca) {
        const int cpar = ca_seen & 1;
        float psum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < n_chunks; ++ch, ++use) {
          const uint32_t buf = use & 1u;
          const uint32_t vm = BD_SEL(ch, vp0, vp1, vp2) & 0xFFFFFFu;
          // write-back columns = the keep columns of the same round (already read).  For lanes 64-127 of the last
          // column third they end one column past the accumulator: column 144 of buffer 0 is column 0 of buffer 1
          // (pixel -1 of the upper lanes: unused), column 288 is the guard in front of the weights.
          const uint32_t wb = lane_addr + buf * kBandAccStride + uint32_t(jbase + half);
          rounds(std::false_type{}, ch, buf, [&](int r, const float (&p8)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i)                   // psum += valid ? p : 0 (predicated add)
              asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p add.f32 %0, %0, %1;\n\t}"
                  : "+f"(psum) : "f"(p8[i]), "r"(vm), "r"(1u << (r * 8 + i)));
            if (ch == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) u0[r * 8 + i] = p8[i];
            } else {
              // conv output back into accumulator columns only this thread touches (own lane, own column third,
              // this round's loads are done), for the second pass
              uint32_t wv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) wv[i] = __float_as_uint(p8[i]);
              tmem_st8(wb + 8 * r, wv);
            }
          });
          if (ch == 0) release_acc(buf);                  // the third chunk reuses this accumulator
        }
        tmem_st_wait();
        if (lane == 0 && warp == 0) BD_STAMP(L, 2);
        // per-CTA channel sums (fixed order) -> slot `rank` of every CTA of the cluster
        red_s[s * 2 + half][c] = psum;
        named_bar_sync(7, kBandEpiWarps * 32);
        if (et < 64) {
          const float t = ((red_s[0][et] + red_s[1][et]) + (red_s[2][et] + red_s[3][et])) + (red_s[4][et] + red_s[5][et]);
          const uint32_t slot = smem_u32(pool_s + (size_t(cpar) * C + rank) * 64 + et);
          const uint32_t pbar = smem_u32(&pool_full[cpar]);
#pragma unroll 1
          for (int d = 0; d < C; ++d) st_async_b32(mapa_u32(slot, uint32_t(d)), __float_as_uint(t), mapa_u32(pbar, uint32_t(d)));
        }
        // y = sigmoid(W2 relu(W1 mean + b1) + b2) for channel c (every warp computes the hidden units itself)
        const int cr = args.cr;
        const float *w1 = lay->w1, *b1 = lay->b1, *w2 = lay->w2;
        float yacc = __ldg(lay->b2 + c);
        band_wait_cluster(&pool_full[cpar], uint32_t(ca_seen >> 1) & 1u);   // every CTA of the image has pushed
        if (et == 0) {
          BD_STAMP(L, 4);
          if (ca_seen + 2 < args.n_ca) mbar_expect_tx(&pool_full[cpar], uint32_t(C) * 64u * 4u);   // arm CA layer +2
        }
        float m0 = 0.f, m1 = 0.f;
        const float* ps = pool_s + size_t(cpar) * C * 64;
#pragma unroll 1
        for (int d = 0; d < C; ++d) { m0 += ps[d * 64 + lane]; m1 += ps[d * 64 + 32 + lane]; }
        m0 *= args.inv_hw; m1 *= args.inv_hw;
#pragma unroll 1
        for (int h = 0; h < cr; ++h) {
          float sdot = __ldg(w1 + h * 64 + lane) * m0 + __ldg(w1 + h * 64 + 32 + lane) * m1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
          yacc = fmaf(__ldg(w2 + c * cr + h), fmaxf(sdot + __ldg(b1 + h), 0.f), yacc);
        }
        scale = 1.f / (1.f + __expf(-yacc));
        ++ca_seen;
      }
      if (lay->q_scale != nullptr) scale *= __ldg(lay->q_scale + n * 64 + c);

      // ---------------------------------------------------------------- every layer: finish the chunks
      const float* res = is_ca ? nullptr : lay->res_f32;
      float* outf = is_ca ? nullptr : lay->out_f32;
      const bool relu = kind == kTrunkRelu;
      const bool update_s = is_ca || lay->update_s != 0;
      const bool use_stream = is_ca || (!relu && res == nullptr && lay->no_res == 0);
      const bool touch = update_s || use_stream;          // the layer reads or replaces the residual stream
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int qb = q_first + kBandNQMax * ch + jbase;
        const uint32_t vp = BD_SEL(ch, vp0, vp1, vp2), vm = vp & 0xFFFFFFu, pad = vp >> 24;
        float f[kBandKeep];
        if (!is_ca) {
          const uint32_t buf = use & 1u;
          rounds(std::true_type{}, ch, buf, [&](int r, const float (&p8)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[r * 8 + i] = p8[i];
          });
          release_acc(buf);
          ++use;
          if (lane == 0 && warp == 0) BD_STAMP(L, 11 + ch);
          if (lane == 0 && warp == 0 && ch == n_chunks - 1) BD_STAMP(L, 2);
        } else if (ch == 0) {
#pragma unroll
          for (int i = 0; i < kBandKeep; ++i) f[i] = u0[i];
        } else {
          const uint32_t buf = (use0 + uint32_t(ch)) & 1u;
          uint32_t wv[kBandKeep];
          tmem_ld24(lane_addr + buf * kBandAccStride + uint32_t(jbase + half), wv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < kBandKeep; ++i) f[i] = __uint_as_float(wv[i]);
          release_acc(buf);
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < kBandKeep; ++i) f[i] = fmaxf(f[i], 0.f);
        } else if (use_stream) {
          // fp32 residual stream: bf16 part from the cell this layer overwrites + fp16 remainder
          const uint32_t cell0 = ob + uint32_t(qb) * 16 + ch_byte;
#pragma unroll
          for (int k = 0; k < kBandKeep / 2; ++k) {
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&s0[k]));
            const float b0 = __uint_as_float(lds_u16(cell0 + uint32_t(2 * k) * 16) << 16) + lo.x;
            const float b1 = __uint_as_float(lds_u16(cell0 + uint32_t(2 * k + 1) * 16) << 16) + lo.y;
            f[2 * k] = fmaf(f[2 * k], scale, b0);
            f[2 * k + 1] = fmaf(f[2 * k + 1], scale, b1);
          }
        } else if (res != nullptr) {                      // fp32 skip from global memory (group / global skip)
          int col;
          size_t pix = first_pix(qb, col);
#pragma unroll
          for (int i = 0; i < kBandKeep; ++i) {
            const float base = ((vm >> i) & 1u) ? res[pix] : 0.f;   // may have been written by this kernel: no __ldg
            f[i] = fmaf(f[i], scale, base);
            walk(col, pix);
          }
        } else {
#pragma unroll
          for (int i = 0; i < kBandKeep; ++i) f[i] *= scale;
        }
        if (outf != nullptr || last) {                    // rare layers: fp32 copy for a later skip / network output
          int col;
          size_t pix = first_pix(qb, col);
#pragma unroll
          for (int i = 0; i < kBandKeep; ++i) {
            if ((vm >> i) & 1u) {
              if (outf != nullptr) outf[pix] = f[i];
              if (last) args.out_bf16[pix] = __float2bfloat16_rn(f[i]);
            }
            walk(col, pix);
          }
        }
        if (!last) store24(ob, qb, vm, pad, f, update_s ? 2 : (touch ? 1 : 0));
        else rotate_only(update_s ? 2 : (touch ? 1 : 0), f);
      }
      if (lane == 0 && warp == 0) BD_STAMP(L, 3);
      if (!last) layer_done(par_out);
    }
#undef BD_SEL
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
#undef BD_STAMP
}

#endif  // RB_TRUNK_KERNEL_IMPL

}  // namespace rb
