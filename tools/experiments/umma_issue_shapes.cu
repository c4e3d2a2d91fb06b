// Experiment (GPU box): the SAME tcgen05.mma sequence as the cluster kernel (108 MMAs per "layer": 3 tiles x 9 taps x 4
// K-steps, planar A moving with the tap, one SW128 weight block per tap), issued from differently shaped code.  The
// tensor pipe needs 48 cycles per MMA (operand reads); what the issuing warp achieves depends on whether ptxas keeps
// the descriptors in UNIFORM registers (UIADD3.64 -> UTCHMMA) or in vector registers (R2UR -> UTCHMMA per operand).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/umma_issue_shapes tools/experiments/umma_issue_shapes.cu
#include "ptx.cuh"
#include <vector>
using namespace rb;

constexpr int PP = 10, PR = 50;
constexpr uint32_t kPlane = PR * PP * 16;
constexpr int kThreads = 320, kMmaWarp = 8;

// predicated issue: every lane executes the (uniform) address arithmetic, one elected lane issues
__device__ __forceinline__ void umma_pred(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t go) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(go)
      : "memory");
}

template <int SHAPE>
__global__ void __launch_bounds__(kThreads, 1) shape_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + 73728;
  for (int i = threadIdx.x; i < (73728 + 8 * int(kPlane)) / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f803f80u * (i & 1), 0x3c003c00u, i, 0x40004000u);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, 64);
  const uint64_t adesc0 = make_smem_desc(0, kPlane, uint32_t(PP) * 16, 0);
  const uint64_t bdesc0 = make_smem_desc(smem_u32(w_s), 16, 1024, kLayoutSw128);
  const uint32_t abuf16 = (smem_u32(a_s) & 0x3FFFF) >> 4;
  const uint32_t kstep = (2 * kPlane) >> 4;
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0;
  if (SHAPE == 0) {                                   // the kernel today: elected lane inside the warp's branch
    if (warp == kMmaWarp) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it)
        for (int j = 0; j < 3; ++j) {
          const uint32_t tile16 = abuf16 + uint32_t(16 * j * PP);
          for (int kx = 0; kx < 3; ++kx) {
            if (elect_one()) {
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const uint64_t adesc = adesc0 + uint64_t(tile16 + uint32_t(ky * PP + kx));
                const uint64_t bdesc = bdesc0 + uint64_t(((kx * 3 + ky) * 8192) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem + uint32_t(256 + j * 64), adesc + uint64_t(k * kstep), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0);
              }
            }
            __syncwarp();
          }
        }
    }
  } else if (SHAPE == 1) {                            // one fixed thread, compile-time index
    if (threadIdx.x == kMmaWarp * 32) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it)
        for (int j = 0; j < 3; ++j) {
          const uint32_t tile16 = abuf16 + uint32_t(16 * j * PP);
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint64_t adesc = adesc0 + uint64_t(tile16 + uint32_t(ky * PP + kx));
              const uint64_t bdesc = bdesc0 + uint64_t(((kx * 3 + ky) * 8192) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem + uint32_t(256 + j * 64), adesc + uint64_t(k * kstep), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0);
            }
        }
    }
  } else {                                            // convergent warp, predicated instruction
    if (warp == kMmaWarp) {
      const uint32_t go = lane == 0 ? 1u : 0u;
      t0 = clock64();
      for (int it = 0; it < iters; ++it)
        for (int j = 0; j < 3; ++j) {
          const uint32_t tile16 = abuf16 + uint32_t(16 * j * PP);
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint64_t adesc = adesc0 + uint64_t(tile16 + uint32_t(ky * PP + kx));
              const uint64_t bdesc = bdesc0 + uint64_t(((kx * 3 + ky) * 8192) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_pred(tmem + uint32_t(256 + j * 64), adesc + uint64_t(k * kstep), bdesc + uint64_t(2 * k), kIdesc, (kx | ky | k) != 0, go);
            }
            __syncwarp();
          }
        }
    }
  }
  if (threadIdx.x == kMmaWarp * 32) {
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}

template <int SHAPE>
void run(long long* dout, const char* name) {
  const int iters = 200, smem = 1024 + 73728 + 8 * kPlane;
  cudaFuncSetAttribute(shape_kernel<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  shape_kernel<SHAPE><<<148, kThreads, smem>>>(dout, iters);
  shape_kernel<SHAPE><<<148, kThreads, smem>>>(dout, iters);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), dout, 148 * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long x : h) mx = x > mx ? x : mx;
  printf("%-70s %.1f cycles per MMA  [%s]\n", name, double(mx) / (iters * 108), cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 148 * 8);
  run<0>(dout, "shape 0: if (warp == W) { if (elect_one()) { 12 MMAs } __syncwarp(); }");
  run<1>(dout, "shape 1: if (threadIdx.x == W * 32) { all MMAs }");
  run<2>(dout, "shape 2: if (warp == W) { convergent arithmetic, @elected tcgen05.mma }");
  return 0;
}
