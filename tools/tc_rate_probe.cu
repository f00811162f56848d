// Probe: true tensor-pipe rate of back-to-back tcgen05.mma (M=128, K=16, kind::f16) for small N.
// The chain is fully unrolled with loop-invariant descriptors, so the SASS is UTCHMMA after UTCHMMA
// (check with cuobjdump -sass): the number is the pipe's, not the issuing thread's.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_rate_probe tools/tc_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../wakeword_detection_b200/csrc/tc_common.cuh"
using namespace wwb::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

template <int N, int M, bool TS, int NACC = 2, int SHIFT = 0>
__global__ void rate_kernel(int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 40000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  if (tid < 32) {
    const uint32_t idesc = make_idesc_f16(M, N);
    const uint64_t da = make_desc(smem_u32(smem) + SHIFT * 16, 656 * 16, 128);   // SHIFT rows: start not 128-byte aligned
    const uint64_t db = make_desc(smem_u32(smem) + 2 * 656 * 16, N * 16, 128);
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (TS) mma_f16_ts(tmem + (j % NACC) * 64, tmem + 480, db, idesc, true);
          else mma_f16_ss(tmem + (j % NACC) * 64, da, db, idesc, true);
        }
      }
      mma_commit(&bar);
    }
    long long t1 = clock64();
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  fence_before_sync(); __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

template <int N, int M, bool TS, int NACC = 2, int SHIFT = 0>
void run(long long* d) {
  cudaFuncSetAttribute(rate_kernel<N, M, TS, NACC, SHIFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  long long h[2][2];
  for (int i = 0; i < 2; ++i) {
    rate_kernel<N, M, TS, NACC, SHIFT><<<1, 128, 61440>>>(i == 0 ? 2 : 10, d);
    if (cudaMemcpy(h[i], d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); exit(1); }
  }
  printf("M=%3d N=%3d %s n_acc=%d shift=%d rows : %.1f clk/MMA (pipe), issue %.1f clk/MMA\n", M, N, TS ? "A=tmem" : "A=smem", NACC, SHIFT,
         (double)(h[1][1] - h[0][1]) / (8 * 16), (double)(h[1][0] - h[0][0]) / (8 * 16));
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  run<8, 128, false>(d); run<16, 128, false>(d); run<32, 128, false>(d); run<48, 128, false>(d); run<64, 128, false>(d);
  run<96, 128, false>(d); run<128, 128, false>(d); run<256, 128, false>(d);
  run<16, 128, true>(d); run<32, 128, true>(d); run<48, 128, true>(d); run<64, 128, true>(d); run<96, 128, true>(d);
  run<128, 128, true>(d); run<256, 128, true>(d);
  run<32, 64, false>(d); run<64, 64, false>(d); run<128, 64, false>(d); run<256, 64, false>(d);
  run<32, 64, true>(d); run<256, 64, true>(d);
  // dependent chains: every MMA accumulates into the same TMEM columns (n_acc = 1) vs. 2 / 4 independent accumulators
  run<32, 128, true, 1>(d); run<32, 128, true, 2>(d); run<32, 128, true, 4>(d);
  run<48, 128, true, 1>(d); run<48, 128, true, 4>(d);
  run<32, 128, false, 1>(d); run<32, 128, false, 2>(d); run<32, 128, false, 4>(d);
  run<64, 128, false, 1>(d); run<64, 128, true, 1>(d);
  // A start address shifted by whole rows (16 B each): the dilated taps of the WaveNet
  run<32, 128, false, 1, 1>(d); run<32, 128, false, 1, 2>(d); run<32, 128, false, 1, 4>(d); run<32, 128, false, 1, 7>(d);
  run<32, 128, false, 1, 8>(d); run<32, 128, false, 1, 16>(d); run<64, 128, false, 1, 1>(d); run<128, 128, false, 1, 1>(d);
  return 0;
}
