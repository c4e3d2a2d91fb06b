// Experiment (GPU box, NOT run yet -- written at the end of round 1 when the GPU budget was spent): issue rate of
// tcgen05.mma kind::f16 with the A operand in TENSOR MEMORY (M = 128, K = 16, cta_group::1) and B in shared memory, as a
// function of N.  Question behind it (candidate 3): with the layer's weights as a TMEM-resident A
// operand and pixels as N, is a 64-channel conv bound by the tensor pipe (N / 2 cycles per MMA) instead of the 128 B/clk
// shared-memory operand port (48 cycles at N = 64 with both operands in shared memory, umma_rate_test.cu)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/umma_ts_rate_test tools/experiments/umma_ts_rate_test.cu
// Expected output per N: cycles per 128 x N x 16 MMA next to the floor N / 2; B traffic is N * 32 B per MMA.
// The A values are whatever the TMEM columns hold (zeros after the allocation on a fresh context are NOT guaranteed):
// this measures rate only, the accumulators are never read.
#include "ptx.cuh"
#include <vector>
using namespace rb;

// D[tmem_d] (+)= A[tmem_a] * B[bdesc]: A is 128 lanes x 16 bf16 = 8 consecutive 32-bit TMEM columns starting at tmem_a
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) ts_rate_kernel(long long* out, int iters) {
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
    const uint32_t b_addr = smem_u32(smem);
    // B: N rows x 64 k (SW128, 128-byte rows: 8 rows per 1024-byte swizzle atom)
    const uint64_t bdesc = make_smem_desc(b_addr, 16, 1024, kLayoutSw128);
    // accumulator in columns [0, N), A operand (four K = 16 slices of 8 columns each) in columns [256, 288)
    const uint32_t a_tmem = tmem_base_s + 256;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base_s, a_tmem + 8 * k, bdesc + uint64_t(2 * k), kIdesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

template <int N>
void run(long long* dout) {
  const int iters = 2000, smem = 1024 + 64 * 1024;
  cudaFuncSetAttribute(ts_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ts_rate_kernel<N><<<148, 128, smem>>>(dout, iters);
  ts_rate_kernel<N><<<148, 128, smem>>>(dout, iters);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  printf("N=%3d A in TMEM, B in smem (%d B per MMA): %.1f cycles per 128xNx16 MMA (floor %d)  [%s]\n", N, N * 32,
         double(mx) / (iters * 4), N / 2, cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  run<64>(dout); run<128>(dout); run<256>(dout);
  return 0;
}
