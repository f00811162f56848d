// K7 — FAR/FRR numerators on the device (reference: utils/evaluate_models.py:183-218,
// utils/plot_eval_models.py:84-129).
//
//  FRR_MAX  : per segment (clip) the maximum posterior; counts[t] += (max > thr[t])
//  FAR_EDGES: per segment a 30-tap mean ('same': out[i] = mean(p[i-15 .. i+14]), zeros
//             beyond the true ends, fp64 like np.convolve on a float64 kernel), then the
//             number of rising edges of (out > thr[t]).
// Thresholds are ascending (np.arange), so every posterior touches a contiguous range of
// threshold indices: the kernels add +1/-1 into a difference array and a final scan
// produces the int64 counts.  Comparisons are strict `>` in fp64.
#include "common.cuh"

namespace wwb {

__device__ __forceinline__ int lower_bound_d(const double* __restrict__ thr, int n, double v) {
  int lo = 0, hi = n;   // first index with thr[idx] >= v
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (thr[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void frr_max_kernel(const float* __restrict__ post, const int64_t* __restrict__ seg_off, int64_t n_seg,
                               const double* __restrict__ thr, int n_thr, long long* __restrict__ diff) {
  const int64_t seg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seg >= n_seg) return;
  const int lane = threadIdx.x & 31;
  const int64_t a = seg_off[seg], b = seg_off[seg + 1];
  if (b <= a) return;      // np.max([]) raises in the reference; the host wrapper rejects empty clips
  float m = -INFINITY;
  for (int64_t i = a + lane; i < b; i += 32) m = fmaxf(m, post[i]);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if (lane == 0) {
    int hi = lower_bound_d(thr, n_thr, (double)m);   // thresholds [0, hi) are < m
    if (hi > 0) {
      atomicAdd((unsigned long long*)&diff[0], 1ull);
      atomicAdd((unsigned long long*)&diff[hi], (unsigned long long)(-1ll));
    }
  }
}

__device__ __forceinline__ double smooth_at(const float* __restrict__ p, int64_t a, int64_t b, int64_t i,
                                            int win, double inv) {
  // mean of p[i-win/2 .. i+win/2-1] clipped to the segment [a, b)
  // even win (30): [i-15, i+14]; odd win (e.g. 5): [i-2, i+2]
  int64_t lo = i - win / 2, hi = i - win / 2 + win;
  if (lo < a) lo = a;
  if (hi > b) hi = b;
  double s = 0.0;
  for (int64_t k = lo; k < hi; ++k) s += (double)p[k] * inv;
  return s;
}

__global__ void far_edges_kernel(const float* __restrict__ post, const int64_t* __restrict__ seg_off, int64_t n_seg,
                                 const int32_t* __restrict__ halo_lo, const int32_t* __restrict__ halo_hi,
                                 const double* __restrict__ thr, int n_thr, int win, int64_t n_total,
                                 long long* __restrict__ diff) {
  extern __shared__ int sdiff[];   // [n_thr + 1]
  for (int i = threadIdx.x; i <= n_thr; i += blockDim.x) sdiff[i] = 0;
  __syncthreads();
  const double inv = 1.0 / (double)win;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total;
       i += (int64_t)gridDim.x * blockDim.x) {
    // segment of i
    int64_t lo = 0, hi = n_seg;
    while (hi - lo > 1) {
      int64_t mid = (lo + hi) >> 1;
      if (seg_off[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t a = seg_off[lo], b = seg_off[lo + 1];
    const int64_t ca = a + (halo_lo ? halo_lo[lo] : 0), cb = b - (halo_hi ? halo_hi[lo] : 0);
    if (i < ca || i >= cb) continue;
    const double cur = smooth_at(post, a, b, i, win, inv);
    int l = 0;
    if (i > a) {
      const double prev = smooth_at(post, a, b, i - 1, win, inv);
      l = lower_bound_d(thr, n_thr, prev);        // thresholds >= prev : "not prev_wake"
    }
    const int h = lower_bound_d(thr, n_thr, cur); // thresholds < cur  : "posterior > threshold"
    if (h > l) {
      atomicAdd(&sdiff[l], 1);
      atomicAdd(&sdiff[h], -1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= n_thr; i += blockDim.x)
    if (sdiff[i]) atomicAdd((unsigned long long*)&diff[i], (unsigned long long)(long long)sdiff[i]);
}

__global__ void scan_counts_kernel(const long long* __restrict__ diff, int n_thr, int64_t* __restrict__ counts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < n_thr; ++i) { run += diff[i]; counts[i] = run; }
  }
}

int launch_eval_counts(wwb_ctx* ctx, const float* post, const int64_t* seg_off, int64_t n_seg,
                       const int32_t* halo_lo, const int32_t* halo_hi, int64_t n_total,
                       const double* thr, int n_thr, int mode, int smooth, int64_t* counts, cudaStream_t st) {
  if (n_thr <= 0 || n_thr > 8192) return fail(ctx, WWB_ERR_ARG, "n_thr out of range");
  void* diff;
  int rc = workspace(ctx, 5, (size_t)(n_thr + 1) * sizeof(long long), &diff);
  if (rc) return rc;
  WWB_CUDA(ctx, cudaMemsetAsync(diff, 0, (size_t)(n_thr + 1) * sizeof(long long), st));
  if (n_seg > 0 && n_total > 0) {
    if (mode == WWB_COUNT_FRR_MAX) {
      frr_max_kernel<<<(unsigned)((n_seg + 7) / 8), 256, 0, st>>>(post, seg_off, n_seg, thr, n_thr, (long long*)diff);
      WWB_CHECK_LAUNCH(ctx);
    } else if (mode == WWB_COUNT_FAR_EDGES) {
      if (smooth < 1) return fail(ctx, WWB_ERR_ARG, "smooth must be >= 1");
      int64_t blocks = std::min<int64_t>((n_total + 255) / 256, (int64_t)ctx->sm_count * 8);
      far_edges_kernel<<<(unsigned)blocks, 256, (n_thr + 1) * sizeof(int), st>>>(
          post, seg_off, n_seg, halo_lo, halo_hi, thr, n_thr, smooth, n_total, (long long*)diff);
      WWB_CHECK_LAUNCH(ctx);
    } else {
      return fail(ctx, WWB_ERR_ARG, "bad count mode %d", mode);
    }
  }
  scan_counts_kernel<<<1, 32, 0, st>>>((long long*)diff, n_thr, counts);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

}  // namespace wwb
