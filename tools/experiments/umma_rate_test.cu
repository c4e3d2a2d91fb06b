// Experiment (GPU box): issue rate of tcgen05.mma kind::f16, M=128, K=16, cta_group::1, operands in shared memory,
// as a function of N -- is the 64->64 conv (N=64: 4 KB of A + 2 KB of B per MMA) bound by the operand reads?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/umma_rate_test tools/experiments/umma_rate_test.cu
#include "ptx.cuh"
#include <vector>
using namespace rb;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int a_swizzled) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint32_t kIdesc = make_idesc_bf16(128, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    // A: 128 rows x 64 k (SW128, 16 KB) or the unswizzled planar layout; B: N rows x 64 k (SW128)
    const uint64_t adesc = a_swizzled ? make_smem_desc(a_addr, 16, 1024, kLayoutSw128) : make_smem_desc(a_addr, 2880, 160, 0);
    const uint64_t bdesc = make_smem_desc(b_addr, 16, 1024, kLayoutSw128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base_s, adesc + uint64_t(a_swizzled ? 2 * k : (2 * k * 2880) >> 4), bdesc + uint64_t(2 * k), kIdesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

template <int N>
void run(long long* dout, int a_sw) {
  const int iters = 2000, smem = 1024 + 64 * 1024;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N><<<148, 128, smem>>>(dout, iters, a_sw);
  rate_kernel<N><<<148, 128, smem>>>(dout, iters, a_sw);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  printf("N=%3d A %s: %.1f cycles per 128xNx16 MMA (floor %d)  [%s]\n", N, a_sw ? "SW128   " : "unswizzled", double(mx) / (iters * 4),
         N / 2, cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  for (int a_sw = 1; a_sw >= 0; --a_sw) {
    run<16>(dout, a_sw); run<32>(dout, a_sw); run<64>(dout, a_sw); run<128>(dout, a_sw); run<256>(dout, a_sw);
  }
  return 0;
}
