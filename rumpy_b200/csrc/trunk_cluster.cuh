// The 64-channel residual trunk with ONE THREAD-BLOCK CLUSTER PER IMAGE: activations never leave shared memory.
//
// Same layer program and same arithmetic as trunk_pipe.cuh (it replaces the same reference call sites:
// common.py:6-9, architectures.py:41-44 / 81-84 / 121-124 / 172-174, common.py:71-75), different machine mapping,
// used when a whole image fits the shared memory of one cluster (e.g. BASELINE configs[1]: 48x48 -> 6 CTAs):
//   * CTA `rank` of the cluster owns a fixed rectangle of the image (th x tw tiles of 16 rows x 8 pixels) and keeps
//     its bf16 activations -- plus a one-pixel halo -- in shared memory, in a channel-planar, UNSWIZZLED layout:
//     8 planes (16 bytes = 8 channels each) of [rows+2][cols+2] pixels.  In that layout the tcgen05 A operand of
//     tap (ky,kx) is the SAME descriptor with its start address moved by (ky*pitch + kx)*16 bytes (SBO = row pitch,
//     LBO = plane stride): no TMA load, no im2col, no per-tap copy -- the conv reads the resident tensor in place;
//   * an epilogue thread owns one pixel: it writes the pixel's 8 chunks into the CTA's own output buffer and, for
//     pixels on the rectangle's edge, straight into the halo cells of the neighbouring CTAs through distributed
//     shared memory (st.shared::cluster), then arrives (release.cluster) on the neighbour's mbarrier.  A CTA starts
//     layer L+1 when its own tiles and every halo pixel it is owed have arrived: no global memory, no fences;
//   * the channel-attention pool is exchanged the same way: per-tile channel sums are pushed into every CTA of the
//     cluster, each CTA reduces them in a fixed order and computes y itself;
//   * fp32 residual stream + accumulators live in TMEM, weights stream through one 72 KB buffer in kx thirds (as
//     in trunk_pipe.cuh).
// Warp roles, NG epilogue groups (template parameter, 2 or 4): warps 0 .. 4*NG-1 epilogue (group e = warp / 4 owns
// channels [64/NG * e, 64/NG * (e+1)) of every tile) | 4*NG MMA issuer 0 | 4*NG+1 weight producer | 4*NG+2 MMA issuer 1.
// TWO MMA-ISSUING WARPS: a 128x64x16 MMA keeps the tensor pipe busy for only 48 cycles (operand reads), the pipe
// accepts hardly more than one MMA ahead, and one warp needs 55-70 cycles per MMA for the descriptor arithmetic
// (vector registers -> R2UR -> UTCHMMA; tools/experiments/umma_env_test.cu: 58-63 cycles per MMA with one issuer,
// 48.2-49.5 with two).  Consecutive tiles (different accumulators, so no ordering between them) alternate between the
// two warps; each one follows the hand-over barriers of the layers its tiles belong to.
#pragma once
#include "trunk_pipe.cuh"

namespace rb {

struct ClusterArgs {
  const TrunkLayer* layers;
  const float* s_init;             // fp32 NHWC: initial residual stream (head conv output)
  const __nv_bfloat16* x_init;     // bf16 NHWC: operand of layer 0 (head conv output)
  __nv_bfloat16* out_bf16;         // bf16 NHWC: output of the last layer
  long long* dbg;                  // optional timeline [grid][dbg_layers][16]
  int n_layers, n_ca, N, H, W, cx, cy, th, tw, cr, dbg_layers;   // n_ca = number of kTrunkCA layers
  // Two-phase layer hand-over (vertical strips only: tw == 1, cy == 1, th >= 3): tiles [0, split) of a layer -- own
  // pixels and the halo columns the neighbours owe for those rows -- complete barrier A, the remaining tiles barrier
  // B.  Tile j of the next layer reads rows of tiles j-1 .. j+1 only, so its MMAs start after A when j + 1 < split:
  // the last tile's epilogue (and, on channel-attention layers, its apply pass) runs underneath the next layer's
  // first MMAs instead of in front of them.  split == 0: one barrier per layer (any rectangle).
  int split;
  int dbg_flags;   // timing experiments only (results are garbage): 1 skip the epilogue's TMEM reads, 2 skip its local stores
  float inv_hw;
};

constexpr int kClusterThreads = 352;   // NG = 2
__host__ __device__ constexpr int cluster_threads(int ng) { return (4 * ng + 3) * 32; }
constexpr int kClusterTileH = 16, kClusterTileW = 8;

__host__ __device__ inline size_t cluster_buf_bytes(int th, int tw) {
  return size_t(8) * (kClusterTileH * th + 2) * (kClusterTileW * tw + 2) * 16;
}
// dynamic smem: [weights 72 KB | buffer 0 | buffer 1 | pool slots 2 x (C * tiles) x 64 floats]
__host__ __device__ inline size_t cluster_smem_bytes(int th, int tw, int C) {
  return 1024 + kTrunkWBytes + 2 * cluster_buf_bytes(th, tw) + size_t(2) * C * th * tw * 64 * sizeof(float);
}

#ifdef RB_TRUNK_KERNEL_IMPL

template <int NG>
__global__ void __launch_bounds__(cluster_threads(NG), 1)
trunk_cluster_kernel_t(const __grid_constant__ CUtensorMap w_map, const ClusterArgs args) {
  constexpr int KC = 64 / NG;      // channels (TMEM columns) per epilogue group
  constexpr int KCH = KC / 8;      // 16-byte chunks (8 channels) per pixel and group
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full[3];
  __shared__ __align__(8) uint64_t w_empty[3];
  __shared__ __align__(8) uint64_t acc_full[kTrunkMaxK];
  __shared__ __align__(8) uint64_t in_full[2];     // barrier A (all of the layer when split == 0)
  __shared__ __align__(8) uint64_t in_full_b[2];   // barrier B (tiles >= split)
  __shared__ __align__(8) uint64_t pool_full[2];
  __shared__ uint32_t tmem_base_s, halo_bytes_s, halo_bytes_b_s;
  __shared__ __align__(16) float y_s[NG][64];
  __shared__ __align__(16) float bias_s[NG][KC], alpha_s[NG][KC], red_s[NG][4][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpMma = 4 * NG, kWarpW = 4 * NG + 1, kWarpMma2 = 4 * NG + 2;
  const int n_layers = args.n_layers;
  const int th = args.th, tw = args.tw, n_tiles = th * tw;
  const int C = args.cx * args.cy;
  const int rank = int(cluster_ctarank());
  const int n = blockIdx.x / C;                     // image of this cluster
  const int ry = rank / args.cx, rx = rank - ry * args.cx;
  const int RH = kClusterTileH * th, RW = kClusterTileW * tw;
  const int PR = RH + 2, PP = RW + 2;
  const uint32_t plane = uint32_t(PR) * PP * 16;    // bytes per 8-channel plane
  const uint32_t buf_bytes = 8 * plane;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* buf0 = smem + kTrunkWBytes;
  float* pool_s = reinterpret_cast<float*>(buf0 + 2 * buf_bytes);   // [2][C * n_tiles][64]
  const int pool_slots = C * n_tiles;

#define CL_STAMP(L_, slot_)                                                                              \
  do {                                                                                                   \
    if (args.dbg && (L_) < args.dbg_layers)                                                              \
      args.dbg[(size_t(blockIdx.x) * args.dbg_layers + (L_)) * 16 + (slot_)] = clock64();                \
  } while (0)

  // ---- zero both activation buffers (halo cells outside the image must read as the conv's zero padding)
  for (uint32_t i = threadIdx.x * 16; i < 2 * buf_bytes; i += cluster_threads(NG) * 16)
    *reinterpret_cast<uint4*>(buf0 + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    // halo pixels this CTA is owed per layer: the halo cells that lie inside the image (each has one owner)
    int halo = 0, halo_b = 0;   // halo_b: owed by the neighbours' tiles >= split (rows of those tiles; cy == 1 there)
    for (int hy = 0; hy < PR; ++hy)
      for (int hx = 0; hx < PP; ++hx) {
        if (hy != 0 && hy != PR - 1 && hx != 0 && hx != PP - 1) continue;
        const int y = ry * RH + hy - 1, x = rx * RW + hx - 1;
        if (!(y >= 0 && y < args.H && x >= 0 && x < args.W)) continue;
        if (args.split > 0 && hy - 1 >= kClusterTileH * args.split) ++halo_b; else ++halo;
      }
    // w_empty: one tcgen05.commit per issuing warp (after its last MMAs of the layer that read the third)
    for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], n_tiles >= 2 ? 2 : 1); }
    for (int i = 0; i < kTrunkMaxK; ++i) mbar_init(&acc_full[i], 1);
    // in_full: one arrival per epilogue group + the arming arrival; the halo pixels complete transaction bytes.
    // Both parities are armed here for layers 0 / 1 (CA layers 0 / 1); later phases are re-armed by their consumer.
    halo_bytes_s = uint32_t(halo) * 128u;
    halo_bytes_b_s = uint32_t(halo_b) * 128u;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], NG + 1);
      mbar_init(&in_full_b[i], NG + 1);
      mbar_init(&pool_full[i], 1);
    }
    fence_mbar_init();
    for (int i = 0; i < 2; ++i) {
      mbar_expect_tx(&in_full[i], uint32_t(halo) * 128u);
      if (args.split > 0) mbar_expect_tx(&in_full_b[i], uint32_t(halo_b) * 128u);
      mbar_expect_tx(&pool_full[i], uint32_t(pool_slots) * 64u * 4u);
    }
  }
  if (warp == kWarpW && lane == 0) tma_prefetch_desc(&w_map);
  if (warp == kWarpMma) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();   // every CTA's buffers are zeroed and its barriers initialised before any remote access
  const uint32_t tmem_base = tmem_base_s;

  if (warp == kWarpMma || warp == kWarpMma2) {
    // ===================================================================== MMA issuers
    const int me = warp == kWarpMma ? 0 : 1;
    const bool two = n_tiles >= 2;                 // one tile per CTA: issuer 0 alone
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
    const uint32_t halo_bytes = halo_bytes_s, halo_bytes_b = halo_bytes_b_s;
    const int split = args.split;
    // descriptors: everything but the 14-bit start-address field (16-byte units) is constant for the whole kernel
    const uint64_t adesc0 = make_smem_desc(0, plane, uint32_t(PP) * 16, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(w_s), 16, 1024, kLayoutSw128);
    const uint32_t kstep = (2 * plane) >> 4;   // K = 16 channels = two planes
    if (two || me == 0) {
      int g = 0;                                   // running tile counter: tile g belongs to issuer g & 1
      for (int L = 0; L < n_layers; ++L) {
        const uint32_t abuf16 = (smem_u32(buf0 + (L & 1) * buf_bytes) & 0x3FFFF) >> 4;
        bool have_a = false, have_b = split == 0, have_w = false;
        int my_last = -1;
        for (int j = 0; j < n_tiles; ++j)
          if (!two || ((g + j) & 1) == me) my_last = j;
        for (int j = 0; j < n_tiles; ++j, ++g) {
          if (two && (g & 1) != me) continue;
          if (!have_a) {
            // own tiles [0, split) of layer L-1 and the halo pixels owed for them have landed in buffer L & 1
            mbar_wait_cluster_trap(&in_full[L & 1], uint32_t(L >> 1) & 1u);
            // the issuer of tile 0 arms layer L+2's phase (it is the only one sure to pass here before anything of
            // layer L+1 is produced)
            if (j == 0 && lane == 0 && L + 2 < n_layers) mbar_expect_tx(&in_full[L & 1], halo_bytes);
            have_a = true;
            if (j >= split - 1) {   // (also split == 0: have_b is set and the single barrier is this one)
              if (!have_b) {
                mbar_wait_cluster_trap(&in_full_b[L & 1], uint32_t(L >> 1) & 1u);
                have_b = true;
              }
            }
            fence_proxy_async_smem();   // generic-proxy writes of the epilogue threads -> async-proxy (tensor core) reads
            tc_fence_after();
            if (j == 0 && lane == 0) CL_STAMP(L, 0);
          }
          if (!have_b && j >= split - 1) {
            // tile j reads rows of tile j + 1 >= split: those complete barrier B
            mbar_wait_cluster_trap(&in_full_b[L & 1], uint32_t(L >> 1) & 1u);
            have_b = true;
            fence_proxy_async_smem();
            tc_fence_after();
          }
          if (split > 0 && j == n_tiles - 1 && lane == 0) {
            // the issuer of the last tile has passed barrier B of this layer: it arms layer L+2's phase
            if (L + 2 < n_layers) mbar_expect_tx(&in_full_b[L & 1], halo_bytes_b);
            CL_STAMP(L, 5);
          }
          const int ta = j / tw, tb = j - ta * tw;
          const uint32_t d_tmem = tmem_base + uint32_t(kTrunkAccCol + j * 64);
          const uint32_t tile16 = abuf16 + uint32_t(kClusterTileH * ta * PP + kClusterTileW * tb);
          for (int kx = 0; kx < 3; ++kx) {
            if (!have_w) { mbar_wait_trap(&w_full[kx], uint32_t(L & 1)); tc_fence_after(); }
            if (elect_one()) {
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                // tap (ky,kx) of tile (ta,tb): the resident plane, start moved by whole pixels
                const uint64_t adesc = adesc0 + uint64_t(tile16 + uint32_t(ky * PP + kx));
                const uint64_t bdesc = bdesc0 + uint64_t(((kx * 3 + ky) * 8192) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // K = 16 channels = planes 2k, 2k+1 (LBO = plane stride)
                  umma_bf16(d_tmem, adesc + uint64_t(k * kstep), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0);
              }
              if (j == my_last) umma_commit(&w_empty[kx]);   // this issuer's last read of the third in this layer
            }
            __syncwarp();
          }
          have_w = true;
          if (elect_one()) umma_commit(&acc_full[j]);
          __syncwarp();
          if (j == n_tiles - 1 && lane == 0) CL_STAMP(L, 1);
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
          tma_load_4d(w_s + kx * kTrunkWThird, &w_map, &w_full[kx], 0, 0, kx * 3, L);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue: NG groups x 128 threads
    // Every group works on EVERY tile: group e owns channels [KC*e, KC*e + KC) of each pixel (balanced for any tile
    // count, and the exposed epilogue of a layer's last tile shrinks with NG).
    const int e = warp >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;           // pixel of the tile == TMEM lane == thread index in the group
    const int ly = row >> 3, lx = row & 7;   // 16 rows x 8 pixels
    const uint32_t bar_id = 1u + uint32_t(e);
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(KC * e);
    const uint32_t plane0 = uint32_t(KCH * e) * plane;   // this group's channel planes
    float* bias_e = bias_s[e];                // KC values: channels KC*e ..
    float* alpha_e = alpha_s[e];              // kTrunkRes: alpha (x the Q-EDSR multiplier of this image's channel)
    float* y_e = y_s[e];

    // One pixel's KC bf16 channels (KCH chunks) -> this CTA's buffer, and (edge pixels) the neighbours' halo cells.
    // `par` selects the destination buffer AND the mbarrier whose transaction count the remote bytes complete.
    const int split = args.split;
    auto write_pixel = [&](int par, int qy, int qx, bool valid, const uint4 (&ch)[KCH]) {
      uint8_t* ob = buf0 + par * buf_bytes + plane0;
      const uint32_t cell = uint32_t((qy + 1) * PP + qx + 1) * 16;
      if (!(args.dbg_flags & 2)) {
#pragma unroll
        for (int c = 0; c < KCH; ++c) *reinterpret_cast<uint4*>(ob + c * plane + cell) = ch[c];
      }
      if (!valid) return;
      const int dyv = qy == 0 ? -1 : (qy == RH - 1 ? 1 : 0);
      const int dxv = qx == 0 ? -1 : (qx == RW - 1 ? 1 : 0);
      if (dyv == 0 && dxv == 0) return;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int dy = k == 1 ? 0 : dyv, dx = k == 0 ? 0 : dxv;   // (dy,0), (0,dx), (dy,dx)
        if ((k == 0 && dyv == 0) || (k == 1 && dxv == 0) || (k == 2 && (dyv == 0 || dxv == 0))) continue;
        const int nry = ry + dy, nrx = rx + dx;
        if (nry < 0 || nry >= args.cy || nrx < 0 || nrx >= args.cx) continue;
        const uint32_t drank = uint32_t(nry * args.cx + nrx);
        const uint32_t rcell = uint32_t((qy + 1 - dy * RH) * PP + (qx + 1 - dx * RW)) * 16;
        const uint32_t raddr = mapa_u32(smem_u32(ob) + rcell, drank);
        const uint32_t rbar = mapa_u32(smem_u32(split > 0 && qy >= kClusterTileH * split ? &in_full_b[par] : &in_full[par]), drank);
#pragma unroll
        for (int c = 0; c < KCH; ++c) st_async_v4(raddr + c * plane, ch[c], rbar);   // KCH x 16 B of this pixel
      }
    };
    // every thread of the group has written its half pixels of ALL tiles: one local arrival per group and layer
    // (with a split: once after tile split - 1 on barrier A, once at the end of the layer on barrier B)
    auto group_done = [&](int par, bool second = false) {
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (row == 0) mbar_arrive(second ? &in_full_b[par] : &in_full[par]);
    };

    // ---- residual stream (fp32 -> TMEM) and the layer-0 operand (bf16 -> buffer 0 + neighbours' halos)
    for (int j = 0; j < n_tiles; ++j) {
      const int ta = j / tw, tb = j - ta * tw;
      const int qy = kClusterTileH * ta + ly, qx = kClusterTileW * tb + lx;
      const int y = ry * RH + qy, x = rx * RW + qx;
      const bool valid = y < args.H && x < args.W;
      const size_t pix = ((size_t(n) * args.H + y) * args.W + x) * 64 + KC * e;
      uint32_t v[KC];
#pragma unroll
      for (int c4 = 0; c4 < KC / 4; ++c4) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) f = __ldg(reinterpret_cast<const float4*>(args.s_init + pix) + c4);
        v[c4 * 4 + 0] = __float_as_uint(f.x); v[c4 * 4 + 1] = __float_as_uint(f.y);
        v[c4 * 4 + 2] = __float_as_uint(f.z); v[c4 * 4 + 3] = __float_as_uint(f.w);
      }
      tmem_st(lane_addr + uint32_t(j * 64), v);
      uint4 ch[KCH];
#pragma unroll
      for (int c = 0; c < KCH; ++c)
        ch[c] = valid ? __ldg(reinterpret_cast<const uint4*>(args.x_init + pix) + c) : make_uint4(0, 0, 0, 0);
      write_pixel(0, qy, qx, valid, ch);
      if (split > 0 && j == split - 1) group_done(0);
    }
    group_done(0, split > 0);
    tmem_st_wait();

    int ca_seen = 0;
    for (int L = 0; L < n_layers; ++L) {
      const TrunkLayer* lay = args.layers + L;
      const int kind = lay->kind;
      const float* bias = lay->bias + KC * e;
      const int par_out = (L + 1) & 1;
      const bool last = L == n_layers - 1;
      if (row < KC) {   // the previous layer ended with a group barrier
        bias_e[row] = __ldg(bias + row);
        alpha_e[row] = (kind == kTrunkRes && lay->q_scale != nullptr)
                           ? lay->alpha * __ldg(lay->q_scale + n * 64 + KC * e + row) : lay->alpha;
      }
      if (kind == kTrunkCA && e == 0 && row < 8 + 2 * args.cr) {
        // the layer's FC parameters (w1 / w2 256*cr B each, b1, b2) towards L2 now: ca_y reads each of them once per
        // CTA, first touch, in the serial chain after the layer's last MMA
        const int cr = args.cr;
        const char* pf = row < 2 * cr ? reinterpret_cast<const char*>(lay->w1) + row * 128
                       : row < 4 * cr ? reinterpret_cast<const char*>(lay->w2) + (row - 2 * cr) * 128
                       : row < 4 * cr + 2 ? reinterpret_cast<const char*>(lay->b2) + (row - 4 * cr) * 128
                                          : reinterpret_cast<const char*>(lay->b1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      }
      named_bar_sync(bar_id, 128);
      // This group's KC bias values, ONCE per layer and thread (16-byte broadcast loads): read per tile -- and, in the
      // ReLU path, as 32 predicated scalar loads -- they were 770 of a layer's shared-memory wavefronts, on the data
      // pipe the tensor core reads its operands through (ncu source page, profiles/README.md).
      float breg[KC];
#pragma unroll
      for (int i4 = 0; i4 < KC / 4; ++i4) {
        const float4 t = reinterpret_cast<const float4*>(bias_e)[i4];
        breg[4 * i4] = t.x; breg[4 * i4 + 1] = t.y; breg[4 * i4 + 2] = t.z; breg[4 * i4 + 3] = t.w;
      }
      const bool alpha_vec = kind == kTrunkRes && lay->q_scale != nullptr;   // per-channel alpha (meta-attention) only
      const float alpha_u = lay->alpha;

      // KC fp32 results of this thread's part of a pixel -> bf16 chunks -> buffer / halos (or global, last layer)
      auto emit = [&](int qy, int qx, bool valid, size_t pix, const float (&f)[KC]) {
        uint4 ch[KCH];
#pragma unroll
        for (int c = 0; c < KCH; ++c)
          ch[c] = valid ? make_uint4(pack_bf16x2(f[c * 8], f[c * 8 + 1]), pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]),
                                     pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]), pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]))
                        : make_uint4(0, 0, 0, 0);
        if (last) {
          if (valid) {
#pragma unroll
            for (int c = 0; c < KCH; ++c) *(reinterpret_cast<uint4*>(args.out_bf16 + pix) + c) = ch[c];
          }
        } else {
          write_pixel(par_out, qy, qx, valid, ch);
        }
      };

      // ---------------------------------------------------------------- conv + bias (+ReLU | + residual)
      auto plain = [&](int j) {
        const int ta = j / tw, tb = j - ta * tw;
        const int qy = kClusterTileH * ta + ly, qx = kClusterTileW * tb + lx;
        const int y = ry * RH + qy, x = rx * RW + qx;
        const bool valid = y < args.H && x < args.W;
        const size_t pix = ((size_t(n) * args.H + y) * args.W + x) * 64 + KC * e;
        const float* res = lay->res_f32;
        float* outf = lay->out_f32;
        const int update_s = lay->update_s;
        mbar_wait(&acc_full[j], uint32_t(L & 1));
        tc_fence_after();
        if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 2);
        uint32_t v[KC];
        float f[KC];
        if (args.dbg_flags & 1) {
#pragma unroll
          for (int i = 0; i < KC; ++i) v[i] = 0;
        } else {
          tmem_ld(lane_addr + uint32_t(kTrunkAccCol + j * 64), v);
        }
        if (kind == kTrunkRelu) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < KC; ++i) f[i] = fmaxf(__uint_as_float(v[i]) + breg[i], 0.f);
        } else {
          if (lay->no_res) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < KC; ++i) f[i] = 0.f;
          } else if (res != nullptr) {
            tmem_ld_wait();
#pragma unroll
            for (int c4 = 0; c4 < KC / 4; ++c4) {
              float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
              if (valid) r = *reinterpret_cast<const float4*>(res + pix + c4 * 4);
              f[c4 * 4 + 0] = r.x; f[c4 * 4 + 1] = r.y; f[c4 * 4 + 2] = r.z; f[c4 * 4 + 3] = r.w;
            }
          } else {
            uint32_t s[KC];
            tmem_ld(lane_addr + uint32_t(j * 64), s);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < KC; ++i) f[i] = __uint_as_float(s[i]);
          }
#pragma unroll
          if (alpha_vec) {
#pragma unroll
            for (int i = 0; i < KC; ++i) f[i] = (__uint_as_float(v[i]) + breg[i]) * alpha_e[i] + f[i];
          } else {
#pragma unroll
            for (int i = 0; i < KC; ++i) f[i] = (__uint_as_float(v[i]) + breg[i]) * alpha_u + f[i];
          }
          if (update_s) {
#pragma unroll
            for (int i = 0; i < KC; ++i) v[i] = __float_as_uint(f[i]);
            tmem_st(lane_addr + uint32_t(j * 64), v);
          }
          if (outf != nullptr && valid) {
#pragma unroll
            for (int c4 = 0; c4 < KC / 4; ++c4)
              *reinterpret_cast<float4*>(outf + pix + c4 * 4) =
                  make_float4(f[c4 * 4], f[c4 * 4 + 1], f[c4 * 4 + 2], f[c4 * 4 + 3]);
          }
        }
        tc_fence_before();
        emit(qy, qx, valid, pix, f);
        if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 3);
      };

      // ---------------------------------------------------------------- channel attention, phase 1: pool
      const int cpar = ca_seen & 1;
      auto ca_pool = [&](int j) {
        const int ta = j / tw, tb = j - ta * tw;
        const int y = ry * RH + kClusterTileH * ta + ly, x = rx * RW + kClusterTileW * tb + lx;
        const bool valid = y < args.H && x < args.W;
        mbar_wait(&acc_full[j], uint32_t(L & 1));
        tc_fence_after();
        if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 2);
        if constexpr (KC == 32) {
          // Fragment-shaped reads: this thread gets 8 columns of 4 of the warp's 32 pixels (per call: rows (lane >> 2)
          // and + 8 of a 16-lane half) -- the per-column sums are local adds + 7 shuffles.  u = accumulator + bias
          // goes back to tensor memory (its own port) in the same shape: the apply pass needs no bias.
          const int tq = lane & 3, tr = lane >> 2;
          const uint32_t a0 = lane_addr + uint32_t(kTrunkAccCol + j * 64);
          uint32_t r0[16], r1[16];
          tmem_ld_frag(a0, r0);
          tmem_ld_frag(a0 + (16u << 16), r1);
          float b8[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 t = *reinterpret_cast<const float2*>(bias_e + 8 * k + 2 * tq);
            b8[2 * k] = t.x; b8[2 * k + 1] = t.y;
          }
          const int y0 = ry * RH + kClusterTileH * ta + 4 * q, x0 = rx * RW + kClusterTileW * tb + tr;
          const bool xin = x0 < args.W;
          tmem_ld_wait();
          if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 14);
          float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int hs = 0; hs < 4; ++hs) {   // pixel row 4q + hs of the tile: half hs >> 1, register pair hs & 1
            const bool ok = xin && (y0 + hs) < args.H;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                const int i = 4 * k + 2 * (hs & 1) + b;
                uint32_t& reg = (hs >> 1) ? r1[i] : r0[i];
                const float u = __uint_as_float(reg) + b8[2 * k + b];
                reg = __float_as_uint(u);
                cs[2 * k + b] += ok ? u : 0.f;
              }
            }
          }
          tmem_st_frag(a0, r0);
          tmem_st_frag(a0 + (16u << 16), r1);
          red_s[e][q][frag_col(lane)] = colsum32_frag(cs, lane);
        } else {
          uint32_t v[KC];
          float f[KC];
          tmem_ld(lane_addr + uint32_t(kTrunkAccCol + j * 64), v);
          tmem_ld_wait();
          if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 14);
#pragma unroll
          for (int i = 0; i < KC; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + breg[i]);
          tmem_st(lane_addr + uint32_t(kTrunkAccCol + j * 64), v);
#pragma unroll
          for (int i = 0; i < KC; ++i) f[i] = valid ? __uint_as_float(v[i]) : 0.f;
          const float cs = lane_transpose_sum16(f, lane);   // column (lane >> 1), at both lanes of the pair
          if ((lane & 1) == 0) red_s[e][q][lane >> 1] = cs;
        }
        tc_fence_before();
        if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 15);
        named_bar_sync(bar_id, 128);
        if (row < KC) {
          // this tile's channel sum -> slot (rank, j) of EVERY CTA of the cluster (fixed slot => fixed sum order)
          const float s = (red_s[e][0][row] + red_s[e][1][row]) + (red_s[e][2][row] + red_s[e][3][row]);
          const uint32_t slot =
              smem_u32(pool_s + (size_t(cpar) * pool_slots + rank * n_tiles + j) * 64 + KC * e + row);
          const uint32_t pbar = smem_u32(&pool_full[cpar]);
          for (int d = 0; d < C; ++d)
            st_async_b32(mapa_u32(slot, uint32_t(d)), __float_as_uint(s), mapa_u32(pbar, uint32_t(d)));
        }
        named_bar_sync(bar_id, 128);   // red_s is free for the next tile
      };

      // ---------------------------------------------------------------- y for the image (once per CA layer and group)
      auto ca_y = [&]() {
        const int cr = args.cr;
        const int c = row & 63, hsel = row >> 6;
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
        mbar_wait_cluster(&pool_full[cpar], uint32_t(ca_seen >> 1) & 1u);   // every tile of the image has pushed
        if (row == 0 && e == 0) {
          CL_STAMP(L, 4);
          if (ca_seen + 2 < args.n_ca)                                     // arm this barrier for CA layer +2
            mbar_expect_tx(&pool_full[cpar], uint32_t(pool_slots) * 64u * 4u);
        }
        // this thread's half of the slots of channel c: four independent partial sums (one dependent generic load
        // per slot was 600 cycles of the chain every CA layer waits on)
        const uint32_t ps = smem_u32(pool_s + size_t(cpar) * pool_slots * 64 + c);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        int sl = hsel;
        for (; sl + 6 < pool_slots; sl += 8) {
#pragma unroll
          for (int k = 0; k < 4; ++k) s4[k] += lds_f32(ps + uint32_t(sl + 2 * k) * 256u);
        }
        for (; sl < pool_slots; sl += 2) s4[0] += lds_f32(ps + uint32_t(sl) * 256u);
        const float ssum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        red_s[e][hsel][c] = ssum;
        if (row == 0 && e == 0) CL_STAMP(L, 10);
        named_bar_sync(bar_id, 128);
        if (row == 0 && e == 0) CL_STAMP(L, 11);
        const float m0 = (red_s[e][0][lane] + red_s[e][1][lane]) * args.inv_hw;
        const float m1 = (red_s[e][0][lane + 32] + red_s[e][1][lane + 32]) * args.inv_hw;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float sdot = w1a[h] * m0 + w1b[h] * m1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
          yacc = fmaf(w2r[h], fmaxf(sdot + b1r[h], 0.f), yacc);
        }
        for (int h = 4; h < cr; ++h) {
          float sdot = __ldg(w1 + h * 64 + lane) * m0 + __ldg(w1 + h * 64 + 32 + lane) * m1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
          yacc = fmaf(__ldg(w2 + c * cr + h), fmaxf(sdot + __ldg(b1 + h), 0.f), yacc);
        }
        if (hsel == 0) y_e[c] = (1.f / (1.f + __expf(-yacc))) * (lay->q_scale ? __ldg(lay->q_scale + n * 64 + c) : 1.f);
        if (row == 0 && e == 0) CL_STAMP(L, 12);
        named_bar_sync(bar_id, 128);
        if (row == 0 && e == 0) CL_STAMP(L, 13);
      };

      // ---------------------------------------------------------------- phase 2: x + u*y from the same accumulator
      // The tensor-memory reads of a tile (accumulator + residual stream, 2 x KC columns) do not depend on y: tile 0's
      // are issued BEFORE the wait for the pool exchange, tile j + 1's before tile j's results are written out, so the
      // ~400-cycle TMEM round trip leaves the critical path pool -> y -> apply -> hand-over.
      uint32_t av[KC], as[KC];
      float yreg[KC];
      auto ca_prefetch = [&](int j) {
        tmem_ld(lane_addr + uint32_t(kTrunkAccCol + j * 64), av);
        tmem_ld(lane_addr + uint32_t(j * 64), as);
      };
      auto ca_apply = [&](int j) {
        const int ta = j / tw, tb = j - ta * tw;
        const int qy = kClusterTileH * ta + ly, qx = kClusterTileW * tb + lx;
        const int y = ry * RH + qy, x = rx * RW + qx;
        const bool valid = y < args.H && x < args.W;
        const size_t pix = ((size_t(n) * args.H + y) * args.W + x) * 64 + KC * e;
        float f[KC];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < KC; ++i) f[i] = fmaf(__uint_as_float(av[i]), yreg[i], __uint_as_float(as[i]));
        {
          uint32_t sv[KC];
#pragma unroll
          for (int i = 0; i < KC; ++i) sv[i] = __float_as_uint(f[i]);
          tmem_st(lane_addr + uint32_t(j * 64), sv);
        }
        if (j + 1 < n_tiles) ca_prefetch(j + 1);
        tc_fence_before();
        emit(qy, qx, valid, pix, f);
        if (row == 0 && e == 0 && j == n_tiles - 1) CL_STAMP(L, 3);
      };

      // hand-over of tiles [0, split): the next layer's first MMAs may start while the remaining tiles finish
      auto early = [&](int j) { if (split > 0 && j == split - 1 && !last) group_done(par_out); };
      if (kind != kTrunkCA) {
        for (int j = 0; j < n_tiles; ++j) { plain(j); early(j); }
      } else {
        for (int j = 0; j < n_tiles; ++j) ca_pool(j);
        if (row == 0 && e == 0) CL_STAMP(L, 9);
        tmem_st_wait();     // u is back in tensor memory
        ca_prefetch(0);
        ca_y();
        {
          const float* yv = y_e + KC * e;   // this group's KC channel scales, once per layer
#pragma unroll
          for (int i4 = 0; i4 < KC / 4; ++i4) {
            const float4 t = reinterpret_cast<const float4*>(yv)[i4];
            yreg[4 * i4] = t.x; yreg[4 * i4 + 1] = t.y; yreg[4 * i4 + 2] = t.z; yreg[4 * i4 + 3] = t.w;
          }
        }
        if (row == 0 && e == 0) CL_STAMP(L, 6);
        for (int j = 0; j < n_tiles; ++j) {
          ca_apply(j);
          early(j);
          if (row == 0 && e == 0 && j < 2) CL_STAMP(L, 7 + j);
        }
        ++ca_seen;
      }
      tmem_st_wait();   // the residual-stream updates of this layer (one wait per layer, not per tile)
      if (!last) group_done(par_out, split > 0); else named_bar_sync(bar_id, 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
#undef CL_STAMP
}

#endif  // RB_TRUNK_KERNEL_IMPL

}  // namespace rb
