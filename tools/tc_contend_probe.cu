// Probe: how much do concurrent epilogue-style instructions slow the tensor pipe?  One warp issues a chain of
// tcgen05.mma (A from smem or TMEM, N = 32) while 16 other warps run a background loop of one kind:
//   0 none, 1 LDS.128 broadcast, 2 tcgen05.ld, 3 tcgen05.st, 4 FFMA/MUFU (issue slots only), 5 STS.128, 6 LDS.128 conflict-free
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_contend_probe tools/tc_contend_probe.cu
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
__global__ void contend_kernel(int bg, int reps, long long* out, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 60000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(128, 32);
    const uint64_t da = make_desc(smem_u32(smem), 656 * 16, 128);
    const uint64_t db = make_desc(smem_u32(smem) + 2 * 656 * 16, 32 * 16, 128);
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (TS) mma_f16_ts(tmem + (j & 1) * 32, tmem + 480, db, idesc, true);
          else mma_f16_ss(tmem + (j & 1) * 32, da, db, idesc, true);
        }
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0) { out[0] = t2 - t0; stop = 1; }
  } else if (warp <= 16) {
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 64 + (warp >> 2) * 64;
    float acc = 0.f;
    const float4* bp = reinterpret_cast<const float4*>(smem + 30000);
    float4* sp = reinterpret_cast<float4*>(smem + 32768) + tid;
    long long n = 0;
    while (!stop) {
      if (bg == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(bp + (i & 3)))); acc += v.x; }
      } else if (bg == 6) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(bp + lane + (i & 3) * 32))); acc += v.x; }
      } else if (bg == 2) {
        float v[16]; tmem_ld16(tb, v); tmem_ld_wait(); acc += v[0];
      } else if (bg == 3) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(acc) + i;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tb), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else if (bg == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc = fmaf(acc, 1.0001f, 0.5f); acc = __expf(acc * 1e-6f) + acc; }
      } else if (bg == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(smem_u32(sp)), "f"(acc) : "memory");
      } else {
        __nanosleep(100);
      }
      ++n;
    }
    if (acc == 123.456f) sink[tid] = acc + (float)n;
  }
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  float* sink; cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(contend_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  cudaFuncSetAttribute(contend_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  const char* names[] = {"none", "LDS.128 broadcast", "tcgen05.ld x16", "tcgen05.st x8", "FFMA+MUFU", "STS.128", "LDS.128 conflict-free"};
  for (int ts = 0; ts < 2; ++ts)
    for (int bg = 0; bg < 7; ++bg) {
      long long h[2];
      for (int i = 0; i < 2; ++i) {
        if (ts) contend_kernel<true><<<1, 17 * 32, 61440>>>(bg, i == 0 ? 8 : 40, d, sink);
        else contend_kernel<false><<<1, 17 * 32, 61440>>>(bg, i == 0 ? 8 : 40, d, sink);
        if (cudaMemcpy(&h[i], d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
      }
      printf("A=%s N=32, background %-22s: %.1f clk/MMA\n", ts ? "tmem" : "smem", names[bg], (double)(h[1] - h[0]) / (32 * 16));
    }
  return 0;
}
