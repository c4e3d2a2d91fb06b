// Experiment (GPU box): cycles per tcgen05.mma (M=128, N=64, K=16, both operands in shared memory) for the EXACT
// descriptor sequence of the cluster kernel -- planar un-swizzled A whose start address moves with the tap, one 8 KB
// SW128 weight block per tap -- against sequences with a fixed A window and / or a fixed B block, nothing else running.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/umma_seq_test tools/experiments/umma_seq_test.cu
#include "ptx.cuh"
#include <vector>
using namespace rb;

constexpr int PP = 10, PR = 50;                 // 48x8-pixel strip + halo, as in the kernel (th = 3, tw = 1)
constexpr uint32_t kPlane = PR * PP * 16;       // 8000 B

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) seq_kernel(long long* out, int iters, int variant, int spin, int mma_warp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;                          // 72 KB of weights
  uint8_t* a_s = smem + 73728;                  // 8 planes
  for (int i = threadIdx.x; i < (73728 + 8 * int(kPlane)) / 16; i += THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f803f80u * (i & 1), 0x3c003c00u, i, 0x40004000u);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (spin && (threadIdx.x >> 5) != mma_warp) {
    // the kernel's waiters: every other warp polls an mbarrier (mbarrier.try_wait suspends up to a time limit)
    while (!mbar_try_wait(&done, 0)) {}
  }
  if (threadIdx.x == mma_warp * 32) {
    constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
    const uint64_t adesc0 = make_smem_desc(0, kPlane, uint32_t(PP) * 16, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(w_s), 16, 1024, kLayoutSw128);
    const uint32_t abuf16 = (smem_u32(a_s) & 0x3FFFF) >> 4;
    const uint32_t kstep = (2 * kPlane) >> 4;
    const bool a_moves = variant == 0 || variant == 2, b_moves = variant == 0 || variant == 1;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int j = 0; j < 3; ++j) {
        const uint32_t tile16 = abuf16 + uint32_t(16 * j * PP);
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t adesc = adesc0 + uint64_t(tile16 + (a_moves ? uint32_t(ky * PP + kx) : 0u));
            const uint64_t bdesc = bdesc0 + uint64_t(b_moves ? (((kx * 3 + ky) * 8192) >> 4) : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base_s + uint32_t(256 + j * 64), adesc + uint64_t(k * kstep), bdesc + uint64_t(2 * k), kIdesc,
                        (kx | ky | k) != 0);
          }
      }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
    mbar_arrive(&done);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  const int iters = 200, smem = 1024 + 73728 + 8 * kPlane;
  const char* names[6] = {"kernel sequence, 128 threads, issued by warp 0", "kernel sequence, 128 threads, issued by warp 1",
                          "kernel sequence, 128 threads, issued by warp 2", "kernel sequence, 128 threads, issued by warp 3", "kernel sequence, 320 threads, warp 8 issues, others parked at the barrier",
                          "kernel sequence, 320 threads, warp 8 issues, nine warps polling an mbarrier"};
  cudaFuncSetAttribute(seq_kernel<320>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(seq_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int v = 0; v < 6; ++v) {
    if (v < 4) { seq_kernel<128><<<148, 128, smem>>>(dout, iters, 0, 0, v); seq_kernel<128><<<148, 128, smem>>>(dout, iters, 0, 0, v); }
    else { seq_kernel<320><<<148, 320, smem>>>(dout, iters, 0, v - 4, 8); seq_kernel<320><<<148, 320, smem>>>(dout, iters, 0, v - 4, 8); }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long x : h) mx = x > mx ? x : mx;
    printf("%-62s %.1f cycles per MMA  [%s]\n", names[v], double(mx) / (iters * 108), cudaGetErrorString(e));
  }
  return 0;
}
