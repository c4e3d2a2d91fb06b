// Experiment (GPU box): what slows the cluster kernel's MMA sequence from 49 cycles per MMA (alone) to ~70 (in the
// kernel)?  The kernel's issue loop with its environment added piece by piece:
//   flag 1  run-time geometry (pitch, tiles per CTA, tile columns as kernel arguments, like ClusterArgs)
//   flag 2  tcgen05.commit per tile (acc_full) and per kx on the last tile (w_empty)
//   flag 4  eight epilogue warps reading the accumulators with tcgen05.ld (32 columns each) in a loop
//   flag 8  eight epilogue warps writing 16-byte cells into the OTHER activation buffer in a loop
//   flag 16 the epilogue warps read AND write tensor memory (tcgen05.ld + tcgen05.st of the residual stream)
//   flag 32 TWO issuing warps (8 and 9), alternating tiles (different accumulators: no ordering between them)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/umma_env_test tools/experiments/umma_env_test.cu
#include "ptx.cuh"
#include <vector>
using namespace rb;

constexpr int kThreads = 320, kMmaWarp = 8;

__global__ void __launch_bounds__(kThreads, 1) env_kernel(long long* out, int iters, int flags, int PP, int PR, int n_tiles, int tw) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, acc_full[4], w_empty[3];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const uint32_t plane = uint32_t(PR) * PP * 16;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + 73728;
  uint8_t* b_s = a_s + 8 * plane;
  for (int i = threadIdx.x; i < (73728 + 16 * int(plane)) / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f803f80u * (i & 1), 0x3c003c00u, i, 0x40004000u);
  if (threadIdx.x == 0) {
    mbar_init(&bar, (flags & 32) ? 2 : 1);
    for (int i = 0; i < 4; ++i) mbar_init(&acc_full[i], 1);
    for (int i = 0; i < 3; ++i) mbar_init(&w_empty[i], 1);
    fence_mbar_init();
    stop = 0;
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
  const uint32_t tmem = tmem_base_s;
  const bool two = flags & 32;
  if (warp == kMmaWarp || (two && warp == kMmaWarp + 1)) {
    const int me = warp - kMmaWarp;
    const bool rt = flags & 1;
    const int pp = rt ? PP : 10, ntl = rt ? n_tiles : 3, twv = rt ? tw : 1;
    const uint32_t pl = rt ? plane : 8000u;
    const uint64_t adesc0 = make_smem_desc(0, pl, uint32_t(pp) * 16, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(w_s), 16, 1024, kLayoutSw128);
    const uint32_t abuf16 = (smem_u32(a_s) & 0x3FFFF) >> 4;
    const uint32_t kstep = (2 * pl) >> 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int j = 0; j < ntl; ++j) {
        if (two && ((it * ntl + j) & 1) != me) continue;
        const int ta = j / twv, tb = j - ta * twv;
        const uint32_t tile16 = abuf16 + uint32_t(16 * ta * pp + 8 * tb);
        const uint32_t d_tmem = tmem + uint32_t(256 + j * 64);
        for (int kx = 0; kx < 3; ++kx) {
          if (elect_one()) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint64_t adesc = adesc0 + uint64_t(tile16 + uint32_t(ky * pp + kx));
              const uint64_t bdesc = bdesc0 + uint64_t(((kx * 3 + ky) * 8192) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, adesc + uint64_t(k * kstep), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0);
            }
            if ((flags & 2) && j == ntl - 1) umma_commit(&w_empty[kx]);
          }
          __syncwarp();
        }
        if ((flags & 2) && elect_one()) umma_commit(&acc_full[j]);
        __syncwarp();
      }
    if (lane == 0) {
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      if (me == 0) { out[blockIdx.x] = clock64() - t0; stop = 1; }
    }
  } else if (warp < 8 && (flags & (4 | 8 | 16))) {
    const int q = warp & 3, e = warp >> 2;
    const uint32_t lane_addr = tmem + (uint32_t(q * 32) << 16) + uint32_t(32 * e);
    uint32_t acc = 0;
    int j = 0;
    while (!stop) {
      if (flags & (4 | 16)) {
        uint32_t v[32];
        tmem_ld(lane_addr + uint32_t(256 + j * 64), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += v[i];
        if (flags & 16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += 1;
          tmem_st(lane_addr + uint32_t(j * 64), v);
          tmem_st_wait();
        }
      }
      if (flags & 8) {
        const int row = q * 32 + lane;
        const uint32_t cell = uint32_t(((row >> 3) + 16 * j + 1) * PP + (row & 7) + 1) * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(b_s + (4 * e + c) * plane + cell) = make_uint4(acc, j, c, row);
      }
      j = j == 2 ? 0 : j + 1;
    }
    if (acc == 0x12345678u) out[148 + blockIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 2 * 148 * 8);
  const int PP = 10, PR = 50;
  const int iters = 200, smem = 1024 + 73728 + 16 * PR * PP * 16;
  cudaFuncSetAttribute(env_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int sets[] = {0, 3 | 16 | 8, 32, 32 | 3, 32 | 3 | 16 | 8};
  for (int f : sets) {
    env_kernel<<<148, kThreads, smem>>>(dout, iters, f, PP, PR, 3, 1);
    env_kernel<<<148, kThreads, smem>>>(dout, iters, f, PP, PR, 3, 1);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long x : h) mx = x > mx ? x : mx;
    printf("flags %2d%s%s%s%s%s%s: %.1f cycles per MMA  [%s]\n", f, f & 32 ? " two-issuers" : "", f & 1 ? " runtime-geometry" : "", f & 2 ? " commits" : "",
           f & 4 ? " tmem-ld" : "", f & 8 ? " smem-stores" : "", f & 16 ? " tmem-ld+st" : "",  double(mx) / (iters * 108), cudaGetErrorString(e));
  }
  return 0;
}
