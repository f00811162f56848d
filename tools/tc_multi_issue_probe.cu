// Probe: tensor-pipe throughput when tcgen05.mma chains are issued by 1, 2 or 4 warps at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_multi_issue_probe tools/tc_multi_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../wakeword_detection_b200/csrc/tc_common.cuh"
using namespace wwb::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
               "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

template <bool TS>
__global__ void multi_kernel(int n_issuers, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 60000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  long long t0 = clock64();
  if (warp < n_issuers) {
    const uint32_t idesc = make_idesc_f16(128, 32);
    const uint64_t da = make_desc(smem_u32(smem) + warp * 4096, 656 * 16, 128);
    const uint64_t db = make_desc(smem_u32(smem) + 2 * 656 * 16 + 8192, 32 * 16, 128);
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (TS) mma_f16_ts(tmem + warp * 64 + (j & 1) * 32, tmem + 480, db, idesc, true);
          else mma_f16_ss(tmem + warp * 64 + (j & 1) * 32, da, db, idesc, true);
        }
      }
      mma_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    long long t2 = clock64();
    if (lane == 0) out[warp] = t2 - t0;
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(multi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  cudaFuncSetAttribute(multi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  for (int ts = 0; ts < 2; ++ts)
    for (int ni : {1, 2, 4}) {
      long long h[2][4];
      for (int i = 0; i < 2; ++i) {
        if (ts) multi_kernel<true><<<1, 128, 61440>>>(ni, i == 0 ? 8 : 40, d);
        else multi_kernel<false><<<1, 128, 61440>>>(ni, i == 0 ? 8 : 40, d);
        if (cudaMemcpy(h[i], d, 32, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
      }
      long long mx = 0;
      for (int w = 0; w < ni; ++w) mx = (h[1][w] - h[0][w]) > mx ? (h[1][w] - h[0][w]) : mx;
      printf("A=%s N=32, %d issuing warps: %.1f clk per MMA overall (%.1f per warp-MMA)\n", ts ? "tmem" : "smem", ni,
             (double)mx / (32 * 16 * ni), (double)mx / (32 * 16));
    }
  return 0;
}
