// Round-2 probe 2: why do TS-mode MMAs against the un-swizzled linear B take ~96 cycles regardless of N (r02_probe)?
// Variants: B layout (no swizzle linear / SW128), D column, A column, N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/r02_probe2 tools/experiments/r02_probe2.cu
#include "ptx.cuh"
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

constexpr int kPlaneB = 353 * 16;

// mode 0: TS, B no-swizzle (LBO = lbo, SBO = sbo)   mode 1: TS, B SW128   mode 2: SS, A SW128 + B no-swizzle
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int mode, int d_col, int a_col,
                                                      int lbo, int sbo, int shift) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < (128 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint32_t kIdesc = make_idesc_bf16(128, N);
    const uint32_t tmem = tmem_base_s;
    const uint64_t b_nosw = make_smem_desc(smem_u32(smem + 32768 + shift * 16), lbo, sbo, 0);
    const uint64_t b_sw = make_smem_desc(smem_u32(smem + 32768), 16, 1024, kLayoutSw128);
    const uint64_t a_sw = make_smem_desc(smem_u32(smem), 16, 1024, kLayoutSw128);
    const uint32_t kstep = (2 * lbo) >> 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (mode == 0) umma_bf16_ts(tmem + d_col, tmem + a_col + 8 * k, b_nosw + uint64_t(k * kstep), kIdesc, 1);
        else if (mode == 1) umma_bf16_ts(tmem + d_col, tmem + a_col + 8 * k, b_sw + uint64_t(2 * k), kIdesc, 1);
        else umma_bf16(tmem + d_col, a_sw + uint64_t(2 * k), b_nosw + uint64_t(k * kstep), kIdesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

template <int N>
void run(long long* dout, int mode, int d_col, int a_col, int lbo, int sbo, int shift, const char* what) {
  const int iters = 1000, smem = 1024 + 128 * 1024;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N><<<148, 128, smem>>>(dout, iters, mode, d_col, a_col, lbo, sbo, shift);
  rate_kernel<N><<<148, 128, smem>>>(dout, iters, mode, d_col, a_col, lbo, sbo, shift);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long v : h) mx = v > mx ? v : mx;
  printf("N=%3d %-52s d_col %3d a_col %3d: %6.1f cycles/MMA (floor %d) [%s]\n", N, what, d_col, a_col,
         double(mx) / (iters * 4), N / 2, cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  run<128>(dout, 1, 0, 256, 0, 0, 0, "TS, B SW128");
  run<128>(dout, 1, 176, 352, 0, 0, 0, "TS, B SW128");
  run<128>(dout, 0, 0, 256, kPlaneB, 128, 0, "TS, B nosw linear LBO=5648 SBO=128");
  run<128>(dout, 0, 0, 256, kPlaneB, 128, 3, "TS, B nosw linear LBO=5648 SBO=128 shift 3 px");
  run<128>(dout, 0, 0, 256, 8192, 128, 0, "TS, B nosw linear LBO=8192 SBO=128");
  run<128>(dout, 0, 0, 256, 8192 + 16, 128, 0, "TS, B nosw linear LBO=8208 SBO=128");
  run<128>(dout, 0, 0, 256, 8192 + 64, 128, 0, "TS, B nosw linear LBO=8256 SBO=128");
  run<128>(dout, 0, 0, 256, 128, 256, 0, "TS, B nosw interleaved LBO=128 SBO=256");
  run<128>(dout, 0, 0, 256, kPlaneB, 18 * 16, 0, "TS, B nosw LBO=5648 SBO=288 (row pitch 18)");
  run<128>(dout, 0, 0, 256, kPlaneB, 10 * 16, 0, "TS, B nosw LBO=5648 SBO=160 (row pitch 10)");
  run<128>(dout, 2, 0, 256, kPlaneB, 128, 0, "SS, A SW128, B nosw linear LBO=5648 SBO=128");
  run<64>(dout, 2, 0, 256, kPlaneB, 128, 0, "SS, A SW128, B nosw linear LBO=5648 SBO=128");
  run<80>(dout, 0, 0, 256, kPlaneB, 128, 0, "TS, B nosw linear LBO=5648 SBO=128");
  run<80>(dout, 1, 0, 256, 0, 0, 0, "TS, B SW128");
  run<96>(dout, 1, 0, 256, 0, 0, 0, "TS, B SW128");
  run<160>(dout, 1, 0, 256, 0, 0, 0, "TS, B SW128");
  run<176>(dout, 1, 0, 256, 0, 0, 0, "TS, B SW128");
  run<176>(dout, 0, 0, 256, kPlaneB, 128, 0, "TS, B nosw linear LBO=5648 SBO=128");
  run<256>(dout, 0, 0, 256, kPlaneB, 128, 0, "TS, B nosw linear LBO=5648 SBO=128");
  return 0;
}
