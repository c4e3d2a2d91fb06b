// Round-2 design probe (GPU box).  Questions it answers before the role-swapped trunk kernel is written:
//   1. how many clusters of size 6..16 are co-resident at the trunk kernel's shared-memory footprint;
//   2. tcgen05.ld throughput with 4 and 8 warps (the role-swapped epilogue reads 2x the accumulator columns);
//   3. is `D^T[c_out][pixel] = W . X^T` with TWO TAPS STACKED ALONG M (A operand in TENSOR MEMORY, written with
//      tcgen05.st or tcgen05.cp; B = resident un-swizzled channel-planar activations addressed LINEARLY over the
//      halo-padded rectangle, SBO = 128 B) numerically the 3x3 conv?  (tools/experiments/stacked_tap_conv_check.py is
//      the numpy statement of the same bookkeeping);
//   4. issue rate of the TS-mode MMAs at N = 80 / 96 / 160 / 176 against un-swizzled B, alone and with four epilogue
//      warps hammering tcgen05.ld;
//   5. tcgen05.cp 128x256b rate (72-80 KB of weights per layer go smem -> TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/r02_probe tools/experiments/r02_probe.cu
#include "ptx.cuh"
#include <cuda_bf16.h>
#include <cmath>
#include <cstdlib>
#include <vector>
using namespace rb;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// ---------------------------------------------------------------------------------------------- 1. occupancy
__global__ void __launch_bounds__(320, 1) dummy_cluster_kernel(int* p) { if (p) *p = 1; }

static void occupancy_table() {
  cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int kb : {120, 170, 200, 220}) {
    cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
    printf("occupancy, %3d KB dynamic smem, 320 threads:", kb);
    for (int cs : {1, 2, 4, 6, 8, 9, 10, 12, 16}) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(cs * 16); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = kb * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_cluster_kernel, &cfg);
      if (e != cudaSuccess) { n = -1; (void)cudaGetLastError(); }
      printf("  cs%d:%d", cs, n);
    }
    printf("\n");
  }
}

// ---------------------------------------------------------------------------------------------- 2. tcgen05.ld rate
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) ld_rate_kernel(long long* out, int iters, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_base_s + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t a[32], b[32];
    tmem_ld32(base + ((it * 64) & 255), a);
    tmem_ld32(base + ((it * 64 + 32) & 255), b);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]) + __uint_as_float(b[i]);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

template <int NW>
static void ld_rate(long long* dout, float* sink) {
  const int iters = 2000;
  ld_rate_kernel<NW><<<148, NW * 32>>>(dout, iters, sink);
  ld_rate_kernel<NW><<<148, NW * 32>>>(dout, iters, sink);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  const double bytes = double(iters) * 2 * 4096 * NW;
  printf("tcgen05.ld 32x32b.x32, %d warps: %.1f cycles per pair of loads per warp, %.1f B/clk/SM  [%s]\n", NW,
         double(mx) / iters, bytes / double(mx), cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------- 3. stacked taps
constexpr int kRH = 16, kRW = 16, kP = kRW + 2;
constexpr int kNQ = 72;                       // output pixels (linear, halo-padded) per chunk
constexpr int kN1 = 80, kN2 = 96;             // accumulator columns: delta = 2 group, delta = P group
constexpr int kChunks = 4;
constexpr int kCells = 353;                   // cells per plane (>= last window's end), stride = 16 mod 128 bytes
constexpr int kPlaneB = kCells * 16;
constexpr int kFirst = kP + 1;

// pair blocks: lo tap offset, hi tap offset (linear; 9999 = none)
__constant__ int c_lo[5] = {-kP - 1, -1, kP - 1, -kP, kP};
__constant__ int c_hi[5] = {-kP + 1, 1, kP + 1, 0, 9999};

// wstack [5 blocks][128 rows][64 k] bf16 (rows 0-63 = lo tap, 64-127 = hi tap or zeros)
// use_cp: 0 = tcgen05.st from registers, 1 = tcgen05.cp from a SW128 K-major smem image
__global__ void __launch_bounds__(160, 1)
stacked_kernel(const __nv_bfloat16* act /*[324 px][64]*/, const __nv_bfloat16* wstack, float* out /*[chunks][2][128][96]*/,
               int use_cp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, drained;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;                        // 5 x 16 KB (SW128 image for tcgen05.cp)
  uint8_t* a_s = smem + 5 * 16384;            // [8 planes][kCells][16 B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kCells * 8; i += 160) {
    const int p = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (p < (kRH + 2) * kP) v = *reinterpret_cast<const uint4*>(act + p * 64 + c * 8);
    *reinterpret_cast<uint4*>(a_s + c * kPlaneB + p * 16) = v;
  }
  for (int i = tid; i < 5 * 128 * 8; i += 160) {
    const int blk = i / 1024, r = (i >> 3) & 127, c = i & 7;
    *reinterpret_cast<uint4*>(w_s + blk * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(wstack + (blk * 128 + r) * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&drained, 128); fence_mbar_init(); }
  if (warp == 4) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t kWCol = 352;                 // weights: 5 blocks x 32 columns
  if (!use_cp && warp < 4) {
    // thread = TMEM lane = row of the stacked block; 32 registers = 64 bf16 k-values
    for (int blk = 0; blk < 5; ++blk) {
      uint32_t v[32];
      const uint4* src = reinterpret_cast<const uint4*>(wstack + (blk * 128 + tid) * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 q = src[c];
        v[c * 4] = q.x; v[c * 4 + 1] = q.y; v[c * 4 + 2] = q.z; v[c * 4 + 3] = q.w;
      }
      tmem_st32(tmem + (uint32_t(warp * 32) << 16) + kWCol + blk * 32, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && lane == 0) {
    if (use_cp) {
      for (int blk = 0; blk < 5; ++blk)
        for (int k = 0; k < 4; ++k)
          tmem_cp_128x256b(tmem + kWCol + blk * 32 + k * 8,
                           make_smem_desc(smem_u32(w_s + blk * 16384), 16, 1024, kLayoutSw128) + uint64_t(2 * k));
    }
    const uint64_t bdesc0 = make_smem_desc(0, kPlaneB, 128, 0);
    const uint32_t a16 = (smem_u32(a_s) & 0x3FFFF) >> 4;
    const uint32_t kstep = (2 * kPlaneB) >> 4;
    for (int ch = 0; ch < kChunks; ++ch) {
      if (ch > 0) { mbar_wait(&drained, (ch - 1) & 1); tc_fence_after(); }
      const int q0 = kFirst + ch * kNQ;
      const uint32_t d1 = tmem + (ch & 1) * 176, d2 = d1 + kN1;
      for (int blk = 0; blk < 5; ++blk) {
        const bool g1 = blk < 3;
        const uint32_t idesc = g1 ? make_idesc_bf16(128, kN1) : make_idesc_bf16(128, kN2);
        const uint64_t bdesc = bdesc0 + uint64_t(a16 + uint32_t(q0 + c_lo[blk]));
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(g1 ? d1 : d2, tmem + kWCol + blk * 32 + k * 8, bdesc + uint64_t(k * kstep), idesc,
                       ((g1 ? blk : blk - 3) | k) != 0);
      }
      umma_commit(&bar);
    }
  }
  if (warp < 4) {
    for (int ch = 0; ch < kChunks; ++ch) {
      mbar_wait(&bar, ch & 1);
      tc_fence_after();
      for (int g = 0; g < 2; ++g)
        for (int c0 = 0; c0 < 96; c0 += 32) {
          if (g == 0 && c0 >= kN1) continue;
          uint32_t v[32];
          const int col = (ch & 1) * 176 + (g ? kN1 : 0) + c0;
          if (g == 0 && c0 == 64) {   // 80 columns: the last 16
            uint32_t w16[16];
            tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + col, w16);
            tmem_ld_wait();
            for (int i = 0; i < 16; ++i) out[((ch * 2 + g) * 128 + tid) * 96 + c0 + i] = __uint_as_float(w16[i]);
            continue;
          }
          tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + col, v);
          tmem_ld_wait();
          for (int i = 0; i < 32; ++i) out[((ch * 2 + g) * 128 + tid) * 96 + c0 + i] = __uint_as_float(v[i]);
        }
      tc_fence_before();
      mbar_arrive(&drained);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

static void stacked_test() {
  const int npx = (kRH + 2) * kP;
  std::vector<__nv_bfloat16> act(npx * 64), ws(5 * 128 * 64);
  std::vector<float> actf(npx * 64, 0.f), wf(9 * 64 * 64);
  srand(3);
  for (int p = 0; p < npx; ++p) {
    const int r = p / kP, c = p % kP;
    const bool inside = r >= 1 && r <= kRH && c >= 1 && c <= kRW;
    for (int ch = 0; ch < 64; ++ch) {
      const float v = inside ? float(rand() % 17 - 8) / 8.f : 0.f;
      act[p * 64 + ch] = __float2bfloat16(v); actf[p * 64 + ch] = v;
    }
  }
  for (auto& v : wf) v = float(rand() % 13 - 6) / 16.f;    // [tap = ky*3+kx][co][ci]
  const int lo[5] = {-kP - 1, -1, kP - 1, -kP, kP}, hi[5] = {-kP + 1, 1, kP + 1, 0, 9999};
  auto tap_of = [&](int off) { const int ky = (off + kP + 1) / kP, kx = (off + kP + 1) % kP; return ky * 3 + kx; };
  for (int b = 0; b < 5; ++b)
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 64; ++k) {
        const int off = r < 64 ? lo[b] : hi[b];
        const float v = off == 9999 ? 0.f : wf[(tap_of(off) * 64 + (r & 63)) * 64 + k];
        ws[(b * 128 + r) * 64 + k] = __float2bfloat16(v);
      }
  __nv_bfloat16 *dact, *dws; float* dout;
  const size_t out_n = size_t(kChunks) * 2 * 128 * 96;
  cudaMalloc(&dact, act.size() * 2); cudaMalloc(&dws, ws.size() * 2); cudaMalloc(&dout, out_n * 4);
  cudaMemcpy(dact, act.data(), act.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dws, ws.data(), ws.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + 5 * 16384 + 8 * kPlaneB;
  cudaFuncSetAttribute(stacked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // direct conv reference on the interior
  std::vector<float> ref(npx * 64, 0.f);
  for (int r = 1; r <= kRH; ++r)
    for (int c = 1; c <= kRW; ++c)
      for (int co = 0; co < 64; ++co) {
        float s = 0.f;
        for (int t = 0; t < 9; ++t) {
          const int p = (r + t / 3 - 1) * kP + c + t % 3 - 1;
          for (int ci = 0; ci < 64; ++ci) s += actf[p * 64 + ci] * wf[(t * 64 + co) * 64 + ci];
        }
        ref[(r * kP + c) * 64 + co] = s;
      }
  for (int use_cp = 0; use_cp < 2; ++use_cp) {
    cudaMemset(dout, 0, out_n * 4);
    stacked_kernel<<<1, 160, smem>>>(dact, dws, dout, use_cp);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("stacked taps (%s): CUDA error %s\n", use_cp ? "tcgen05.cp" : "tcgen05.st", cudaGetErrorString(e)); return; }
    std::vector<float> got(out_n);
    cudaMemcpy(got.data(), dout, out_n * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int ch = 0; ch < kChunks; ++ch)
      for (int j = 0; j < kNQ; ++j) {
        const int q = kFirst + ch * kNQ + j;
        const int r = q / kP, c = q % kP;
        if (r < 1 || r > kRH || c < 1 || c > kRW) continue;
        for (int co = 0; co < 64; ++co) {
          auto A = [&](int g, int row, int col) { return got[((size_t(ch) * 2 + g) * 128 + row) * 96 + col]; };
          const float v = A(0, co, j) + A(0, 64 + co, j + 2) + A(1, co, j) + A(1, 64 + co, j + kP);
          maxerr = fmax(maxerr, fabs(v - ref[q * 64 + co])); maxref = fmax(maxref, fabs(ref[q * 64 + co]));
        }
      }
    printf("stacked taps, A in TMEM via %s, B planar linear (SBO=128, LBO=plane), N=%d/%d: max |err| %.4g (ref absmax %.4g) -> %s\n",
           use_cp ? "tcgen05.cp 128x256b" : "tcgen05.st", kN1, kN2, maxerr, maxref, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
  }
}

// ---------------------------------------------------------------------------------------------- 4./5. rates
// MMA issuer (warp 4) runs `iters` x 20 TS-mode MMAs (3 x 4 of N1, 2 x 4 of N2) against un-swizzled B; warps 0-3
// optionally stream tcgen05.ld over the OTHER accumulator half; optionally 20 tcgen05.cp per iteration in the MMA stream.
template <int N1, int N2>
__global__ void __launch_bounds__(160, 1) ts_rate_kernel(long long* out, int iters, int ld_load, int cp_load, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (140 * 1024) / 16; i += 160) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop_s = 0; }
  if (warp == 4) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 4) {
    if (lane == 0) {
      const uint64_t bdesc0 = make_smem_desc(smem_u32(smem + 81920), kPlaneB, 128, 0);
      const uint32_t kstep = (2 * kPlaneB) >> 4;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d1 = tmem + (it & 1) * 176, d2 = d1 + N1;
        for (int blk = 0; blk < 5; ++blk) {
          const uint32_t idesc = blk < 3 ? make_idesc_bf16(128, N1) : make_idesc_bf16(128, N2);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(blk < 3 ? d1 : d2, tmem + 352 + blk * 32 + k * 8, bdesc0 + uint64_t(blk * 3 + k * kstep), idesc, 1);
          if (cp_load)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tmem_cp_128x256b(tmem + 352 + blk * 32 + k * 8,
                               make_smem_desc(smem_u32(smem + blk * 16384), 16, 1024, kLayoutSw128) + uint64_t(2 * k));
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[blockIdx.x] = clock64() - t0;
      stop_s = 1;
    }
  } else if (ld_load) {
    float acc = 0.f;
    int it = 0;
    while (!stop_s) {
      uint32_t a[32];
      tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + ((it * 32) % 160), a);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]);
      ++it;
    }
    if (acc == 123.456f) sink[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N1, int N2>
static void ts_rate(long long* dout, float* sink, int ld_load, int cp_load) {
  const int iters = 500, smem = 1024 + 140 * 1024;
  cudaFuncSetAttribute(ts_rate_kernel<N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ts_rate_kernel<N1, N2><<<148, 160, smem>>>(dout, iters, ld_load, cp_load, sink);
  ts_rate_kernel<N1, N2><<<148, 160, smem>>>(dout, iters, ld_load, cp_load, sink);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  printf("TS MMAs N=%d x12 + N=%d x8 per chunk%s%s: %.0f cycles per chunk (floor %d)  [%s]\n", N1, N2,
         ld_load ? ", 4 warps of tcgen05.ld" : "", cp_load ? ", + 20 tcgen05.cp (80 KB)" : "", double(mx) / iters,
         12 * N1 / 2 + 8 * N2 / 2, cudaGetErrorString(e));
}

int main() {
  long long* dout; float* sink;
  cudaMalloc(&dout, 148 * 8); cudaMalloc(&sink, 4);
  occupancy_table();
  ld_rate<4>(dout, sink);
  ld_rate<8>(dout, sink);
  stacked_test();
  ts_rate<80, 96>(dout, sink, 0, 0);
  ts_rate<80, 96>(dout, sink, 1, 0);
  ts_rate<80, 96>(dout, sink, 0, 1);
  ts_rate<80, 96>(dout, sink, 1, 1);
  ts_rate<160, 176>(dout, sink, 0, 0);
  ts_rate<160, 176>(dout, sink, 1, 0);
  ts_rate<128, 128>(dout, sink, 0, 0);
  return 0;
}
