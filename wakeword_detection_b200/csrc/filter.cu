// K1 — fused filter kernel: PCM -> (scale/clip) -> pre-emphasis -> frames of 512 @ hop 160
// -> Hann -> 512-point real FFT -> |.| -> 257x40 mel (sparse) -> 0.5*(ln(max(.,1e-5))+11.5129).
//
// Replaces (reference): spokestack/wakeword/tflite.py:148-191, utils/tf_lite/filter.py:38-75
// and filter.tflite.  HBM traffic is the algorithmic minimum (each PCM sample read once
// per tile + 6 % halo, each mel value written once); the limiter is fp32 ALU work of the
// FFT, so the layout is chosen for ALU efficiency:
//   * 16 threads per frame; the 512 real samples are packed into 256 complex points,
//     z[n] = x[2n] + i x[2n+1], and transformed as 16 x 16 (two in-register radix-16
//     passes, one transpose through shared memory, twiddles W256^(j*k) held in registers
//     because j is fixed per thread for the whole kernel);
//   * the split step of the real FFT pairs bin k with 256-k; the partner value lives in
//     lane (16-j) of the same 16-lane group and is fetched with one shuffle per value;
//   * magnitudes go to shared memory, the mel projection is evaluated from a segment
//     table (<= 8 non-zeros per segment, fixed summation order => deterministic results).
#include "common.cuh"

namespace wwb {

constexpr int FR_TILE = 16;                                // frames per CTA pass
constexpr int F_THREADS = FR_TILE * 16;                    // 256
constexpr int TILE_SAMPLES = (FR_TILE - 1) * kHop + kFFT;  // 2912
constexpr int XP = 17;                                     // transpose pitch (float2)
constexpr int MAGP = 273;                                  // magnitude row pitch (floats)
constexpr int MAX_SEG = 128;

struct MelParams {
  MelTables mt;
  float mel_floor, mel_log_offset, mel_scale;
};

struct FilterParams {
  const void* pcm;
  int dtype;
  int64_t n_streams, n_samples, pitch;
  int64_t n_frames;        // per stream
  int tiles_per_stream;
  int64_t n_tiles;
  float a;                 // pre-emphasis
  float* mel;              // [S, F, 40]
  const float* hann_half;  // [512] 0.5*hann
  const float2* tw256;
  const float2* tw512;
  MelParams mp;
};

struct __align__(16) FilterSmem {
  float samples[TILE_SAMPLES + 8];
  float2 xch[FR_TILE][16 * XP];
  float mag[FR_TILE][MAGP];
  float hann[kFFT];
  float2 tw512[256];
  float partial[FR_TILE][MAX_SEG];
  // mel tables (batch kernel): segment -> taps, band -> segments
  float seg_w[MAX_SEG][8];             // every segment padded to 8 taps (weight 0, bin 0): fma(0, x, acc) == acc
  unsigned short seg_bin[MAX_SEG][8];
  int band_seg0[kMel + 1];
  float band_bias[kMel];
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward radix-4 butterfly (W4 = -i), in place, natural output order
__device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x);
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x);
}

// in-register forward DFT of 16 points.  Input natural order; output X[k] is left in
// v[4*(k&3) + (k>>2)]  (see XI()).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
#pragma unroll
  for (int a = 0; a < 4; ++a) bfly4(v[a], v[a + 4], v[a + 8], v[a + 12]);
  // twiddles W16^(a*q) on v[a + 4q]
  v[1 + 4] = cmul(v[1 + 4], make_float2(C1, -S1));                                   // W^1
  v[1 + 8] = make_float2((v[1 + 8].x + v[1 + 8].y) * R, (v[1 + 8].y - v[1 + 8].x) * R);   // W^2
  v[1 + 12] = cmul(v[1 + 12], make_float2(S1, -C1));                                 // W^3
  v[2 + 4] = make_float2((v[2 + 4].x + v[2 + 4].y) * R, (v[2 + 4].y - v[2 + 4].x) * R);   // W^2
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);                                   // W^4 = -i
  v[2 + 12] = make_float2((v[2 + 12].y - v[2 + 12].x) * R, -(v[2 + 12].x + v[2 + 12].y) * R);  // W^6
  v[3 + 4] = cmul(v[3 + 4], make_float2(S1, -C1));                                   // W^3
  v[3 + 8] = make_float2((v[3 + 8].y - v[3 + 8].x) * R, -(v[3 + 8].x + v[3 + 8].y) * R);  // W^6
  v[3 + 12] = cmul(v[3 + 12], make_float2(-C1, S1));                                 // W^9
#pragma unroll
  for (int q = 0; q < 4; ++q) bfly4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ constexpr int XI(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ float pcm_to_float(int16_t s) {
  // frame.astype(np.float32) / (2**15 - 1), np.clip(-1, 1)   (wakeword/tflite.py:150-151)
  float x = __fdiv_rn((float)s, 32767.0f);
  return fminf(fmaxf(x, -1.0f), 1.0f);
}

// One frame per 16-lane group: samples (already pre-emphasised) in shared memory at `x`
// (8-byte aligned); writes 257 magnitudes to `mag`.
__device__ __forceinline__ void frame_spectrum(const float* __restrict__ x, const float* __restrict__ hann,
                                               const float2* __restrict__ tw512, const float2 (&twj)[16],
                                               float2* __restrict__ xch, float* __restrict__ mag, int j,
                                               unsigned group_mask) {
  float2 v[16];
  const float2* x2 = reinterpret_cast<const float2*>(x);
  const float2* h2 = reinterpret_cast<const float2*>(hann);
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    float2 s = x2[j + 16 * m];
    float2 h = h2[j + 16 * m];
    v[m] = make_float2(s.x * h.x, s.y * h.y);
  }
  fft16(v);
  // Y[k2] *= W256^(j*k2); transpose through shared memory
  xch[j * XP + 0] = v[XI(0)];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) xch[j * XP + k2] = cmul(v[XI(k2)], twj[k2]);
  __syncwarp(group_mask);
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = xch[n1 * XP + j];
  __syncwarp(group_mask);
  fft16(v);
  // Z[16*k1 + j] = v[XI(k1)].  Real-FFT split: X[k] = E + W512^k * O with
  // E = (Zk + conj(Zp))/2, O = -i (Zk - conj(Zp))/2, Zp = Z[(256-k) mod 256]; the 1/2 is
  // folded into the window table.
  const int lane = threadIdx.x & 31;
  const int src = (lane & 16) | ((16 - j) & 15);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    float2 zk = v[XI(k1)];
    float2 send = v[XI(15 - k1)];
    float2 zp;
    zp.x = __shfl_sync(group_mask, send.x, src);
    zp.y = __shfl_sync(group_mask, send.y, src);
    if (j == 0) zp = v[XI((16 - k1) & 15)];
    float er = zk.x + zp.x, ei = zk.y - zp.y;
    float orr = zk.y + zp.y, oi = zp.x - zk.x;
    float2 w = tw512[16 * k1 + j];
    float xr = er + w.x * orr - w.y * oi;
    float xi = ei + w.x * oi + w.y * orr;
    mag[16 * k1 + j] = sqrtf(xr * xr + xi * xi);
    if (j == 0 && k1 == 0) mag[256] = fabsf(zk.x - zk.y);   // W512^256 = -1
  }
}

template <typename T>
__device__ __forceinline__ float load_sample(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, int64_t i) {
  return pcm_to_float(__ldg(p + i));
}
template <>
__device__ __forceinline__ float load_sample<float>(const float* p, int64_t i) {
  return __ldg(p + i);
}

// mel projection + log compression for `nf` frames whose magnitudes are in sm.mag;
// writes out[frame*40 + band].
__device__ __forceinline__ void mel_phase(FilterSmem& sm, const MelParams& P, int nf, float* __restrict__ out) {
  const MelTables& mt = P.mt;
  const int n_items = mt.n_seg * FR_TILE;
  for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
    int f = it & (FR_TILE - 1), seg = it / FR_TILE;
    if (f < nf) {
      int first = __ldg(mt.seg_first + seg), cnt = __ldg(mt.seg_count + seg);
      float acc = 0.f;
      for (int t = 0; t < cnt; ++t)
        acc = fmaf(__ldg(mt.tap_w + first + t), sm.mag[f][__ldg(mt.tap_bin + first + t)], acc);
      sm.partial[f][seg] = acc;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nf * kMel; o += blockDim.x) {
    int f = o / kMel, band = o - f * kMel;
    int s0 = __ldg(mt.band_seg0 + band), s1 = __ldg(mt.band_seg0 + band + 1);
    float acc = 0.f;
    for (int s = s0; s < s1; ++s) acc += sm.partial[f][s];
    acc += __ldg(mt.bias + band);
    acc = fmaxf(acc, P.mel_floor);
    float y = logf(acc);
    y = __fsub_rn(y, P.mel_log_offset);
    out[o] = __fmul_rn(y, P.mel_scale);
  }
}

// Same arithmetic as mel_phase (per segment an fma chain from 0, segments added in order, bias, floor, log), i.e.
// bit-identical to mel_phase / mel_from_mag_kernel, but with the tables in shared memory and every segment padded
// to exactly 8 taps, so the 8 (weight, bin, magnitude) loads of an item are issued together instead of one
// dependent load chain per tap.  Contains one barrier.
__device__ __forceinline__ void mel_fast(FilterSmem& sm, const MelParams& P, int n_seg, int nf, float* __restrict__ out) {
  for (int it = threadIdx.x; it < n_seg * FR_TILE; it += blockDim.x) {
    const int f = it & (FR_TILE - 1), seg = it / FR_TILE;
    if (f < nf) {
      const float4 w0 = *reinterpret_cast<const float4*>(&sm.seg_w[seg][0]), w1 = *reinterpret_cast<const float4*>(&sm.seg_w[seg][4]);
      const uint4 bb = *reinterpret_cast<const uint4*>(&sm.seg_bin[seg][0]);
      const float* row = sm.mag[f];
      const float m0 = row[bb.x & 0xffff], m1 = row[bb.x >> 16], m2 = row[bb.y & 0xffff], m3 = row[bb.y >> 16];
      const float m4 = row[bb.z & 0xffff], m5 = row[bb.z >> 16], m6 = row[bb.w & 0xffff], m7 = row[bb.w >> 16];
      float acc = 0.f;
      acc = fmaf(w0.x, m0, acc); acc = fmaf(w0.y, m1, acc); acc = fmaf(w0.z, m2, acc); acc = fmaf(w0.w, m3, acc);
      acc = fmaf(w1.x, m4, acc); acc = fmaf(w1.y, m5, acc); acc = fmaf(w1.z, m6, acc); acc = fmaf(w1.w, m7, acc);
      sm.partial[f][seg] = acc;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nf * kMel; o += blockDim.x) {
    const int f = o / kMel, band = o - f * kMel;
    const int s0 = sm.band_seg0[band], s1 = sm.band_seg0[band + 1];
    float acc = 0.f;
    for (int s = s0; s < s1; ++s) acc += sm.partial[f][s];
    acc += sm.band_bias[band];
    acc = fmaxf(acc, P.mel_floor);
    float y = logf(acc);
    y = __fsub_rn(y, P.mel_log_offset);
    out[o] = __fmul_rn(y, P.mel_scale);
  }
}

template <typename T>
__global__ void __launch_bounds__(F_THREADS, 2) filter_kernel(const FilterParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FilterSmem& sm = *reinterpret_cast<FilterSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int j = tid & 15;          // lane within the frame group
  const int g = tid >> 4;          // frame within the tile
  const unsigned group_mask = (tid & 16) ? 0xffff0000u : 0x0000ffffu;

  for (int i = tid; i < kFFT; i += F_THREADS) sm.hann[i] = P.hann_half[i];
  for (int i = tid; i < 256; i += F_THREADS) sm.tw512[i] = P.tw512[i];
  float2 twj[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) twj[k2] = P.tw256[(j * k2) & 255];
  __syncthreads();

  {
    const MelTables& mt = P.mp.mt;
    for (int i = tid; i < mt.n_seg * 8; i += F_THREADS) {
      const int seg = i >> 3, t = i & 7;
      const bool in = t < mt.seg_count[seg];
      sm.seg_w[seg][t] = in ? mt.tap_w[mt.seg_first[seg] + t] : 0.0f;
      sm.seg_bin[seg][t] = in ? (unsigned short)mt.tap_bin[mt.seg_first[seg] + t] : (unsigned short)0;
    }
    for (int i = tid; i <= kMel; i += F_THREADS) sm.band_seg0[i] = mt.band_seg0[i];
    for (int i = tid; i < kMel; i += F_THREADS) sm.band_bias[i] = mt.bias[i];
  }
  __syncthreads();

  const T* pcm = reinterpret_cast<const T*>(P.pcm);
  // Fast path (int16, 8-byte aligned rows, no pre-emphasis = the reference's default): the next tile's PCM is
  // prefetched into registers while the current tile's spectra are computed, and a tile costs two barriers.
  const bool fast = sizeof(T) == 2 && P.a == 0.0f && (reinterpret_cast<uintptr_t>(pcm) & 7) == 0 && (P.pitch & 3) == 0;
  constexpr int NPRE = (TILE_SAMPLES / 4 + F_THREADS - 1) / F_THREADS;   // 3 uint2 per thread
  uint2 pre[NPRE];
  auto tile_geom = [&](int64_t tile, int64_t& s_out, int64_t& f0_out, int& nf_out) {
    s_out = tile / P.tiles_per_stream;
    f0_out = (tile - s_out * P.tiles_per_stream) * FR_TILE;
    nf_out = (int)min((int64_t)FR_TILE, P.n_frames - f0_out);
  };
  auto prefetch = [&](int64_t tile) {
    int64_t s, f0; int nf;
    tile_geom(tile, s, f0, nf);
    const int ns = (nf - 1) * kHop + kFFT;
    const uint2* v4 = reinterpret_cast<const uint2*>(reinterpret_cast<const int16_t*>(P.pcm) + s * P.pitch + f0 * kHop);
#pragma unroll
    for (int c = 0; c < NPRE; ++c) {
      const int i = tid + c * F_THREADS;
      pre[c] = (i < ns / 4) ? __ldg(v4 + i) : make_uint2(0u, 0u);
    }
  };
  if (fast && (int64_t)blockIdx.x < P.n_tiles) prefetch(blockIdx.x);

  for (int64_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    int64_t s, f0; int nf;
    tile_geom(tile, s, f0, nf);
    const int ns = (nf - 1) * kHop + kFFT;
    const T* row = pcm + s * P.pitch;
    const int64_t start = f0 * kHop;

    // stage samples (converted to float) in shared memory
    if (fast) {
#pragma unroll
      for (int c = 0; c < NPRE; ++c) {
        const int i = tid + c * F_THREADS;
        if (i < ns / 4) {
          const uint2 r = pre[c];
          *reinterpret_cast<float4*>(&sm.samples[4 * i]) =
              make_float4(pcm_to_float((int16_t)(r.x & 0xffff)), pcm_to_float((int16_t)(r.x >> 16)),
                          pcm_to_float((int16_t)(r.y & 0xffff)), pcm_to_float((int16_t)(r.y >> 16)));
        }
      }
    } else if (sizeof(T) == 2 && ((reinterpret_cast<uintptr_t>(row + start) & 7) == 0)) {
      const uint2* v4 = reinterpret_cast<const uint2*>(row + start);
      for (int i = tid; i < ns / 4; i += F_THREADS) {
        uint2 r = __ldg(v4 + i);
        sm.samples[4 * i + 0] = pcm_to_float((int16_t)(r.x & 0xffff));
        sm.samples[4 * i + 1] = pcm_to_float((int16_t)(r.x >> 16));
        sm.samples[4 * i + 2] = pcm_to_float((int16_t)(r.y & 0xffff));
        sm.samples[4 * i + 3] = pcm_to_float((int16_t)(r.y >> 16));
      }
    } else {
      for (int i = tid; i < ns; i += F_THREADS) sm.samples[i] = load_sample<T>(row, start + i);
    }
    if (P.a != 0.0f) {
      // y[n] = x[n] - a*x[n-1] with fp32 product and difference (numpy semantics)
      __syncthreads();
      float y[(TILE_SAMPLES + F_THREADS - 1) / F_THREADS];
      int c = 0;
      for (int i = tid; i < ns; i += F_THREADS, ++c) {
        float xp = (i > 0) ? sm.samples[i - 1] : (start > 0 ? load_sample<T>(row, start - 1) : 0.0f);
        y[c] = __fsub_rn(sm.samples[i], __fmul_rn(P.a, xp));
      }
      __syncthreads();
      c = 0;
      for (int i = tid; i < ns; i += F_THREADS, ++c) sm.samples[i] = y[c];
    }
    __syncthreads();
    if (fast && tile + gridDim.x < P.n_tiles) prefetch(tile + gridDim.x);   // latency hidden behind the FFTs

    if (g < nf)
      frame_spectrum(sm.samples + g * kHop, sm.hann, sm.tw512, twj, sm.xch[g], sm.mag[g], j, group_mask);
    __syncthreads();
    // (no barrier after this: the next store goes to `samples`, which nobody reads any more, and the next
    //  spectra overwrite `mag` / the next mel phase `partial` only after the barrier that follows that store)
    mel_fast(sm, P.mp, P.mp.mt.n_seg, nf, P.mel + (s * P.n_frames + f0) * kMel);
  }
}

int launch_filter(wwb_ctx* ctx, const void* pcm, int dtype, int64_t S, int64_t N, int64_t pitch,
                  float a, float* mel, cudaStream_t st) {
  if (dtype != WWB_PCM_I16 && dtype != WWB_PCM_F32) return fail(ctx, WWB_ERR_ARG, "bad pcm dtype %d", dtype);
  if (S < 0 || N < 0 || pitch < N) return fail(ctx, WWB_ERR_ARG, "bad filter geometry");
  int64_t F = wwb_num_frames(N);
  if (S == 0 || F == 0) return WWB_OK;
  if (ctx->mel.n_seg > MAX_SEG) return fail(ctx, WWB_ERR_ARG, "mel matrix has too many segments");
  FilterParams P;
  P.pcm = pcm; P.dtype = dtype; P.n_streams = S; P.n_samples = N; P.pitch = pitch;
  P.n_frames = F;
  P.tiles_per_stream = (int)((F + FR_TILE - 1) / FR_TILE);
  P.n_tiles = S * P.tiles_per_stream;
  P.a = a; P.mel = mel; P.hann_half = ctx->hann; P.tw256 = ctx->tw256; P.tw512 = ctx->tw512;
  P.mp.mt = ctx->mel;
  P.mp.mel_floor = ctx->mel_floor; P.mp.mel_log_offset = ctx->mel_log_offset; P.mp.mel_scale = ctx->mel_scale;
  size_t smem = sizeof(FilterSmem);
  int64_t grid = std::min<int64_t>(P.n_tiles, (int64_t)ctx->sm_count * 2);
  if (dtype == WWB_PCM_I16) {
    WWB_CUDA(ctx, cudaFuncSetAttribute(filter_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_kernel<int16_t><<<(unsigned)grid, F_THREADS, smem, st>>>(P);
  } else {
    WWB_CUDA(ctx, cudaFuncSetAttribute(filter_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    filter_kernel<float><<<(unsigned)grid, F_THREADS, smem, st>>>(P);
  }
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

// ---------------------------------------------------------------------------------------
// Streaming front end (WakewordTrigger._sample, wakeword/tflite.py:148-168) for many
// streams: one CTA per stream appends the chunk to the stream's pending samples, emits
// every frame that completes (only analysed while is_speech, :166-167), pushes the mel
// frames into the stream's ring and records one encoder window per analysed frame.
struct StreamFilterParams {
  const int16_t* pcm;
  int64_t n;
  const uint8_t* is_speech;
  const uint8_t* is_active;
  float a;
  StreamState st;
  int L;
  const float* hann_half;
  const float2* tw256;
  const float2* tw512;
  MelParams mp;
};

__global__ void __launch_bounds__(F_THREADS, 2) stream_filter_kernel(const StreamFilterParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FilterSmem& sm = *reinterpret_cast<FilterSmem*>(smem_raw);
  __shared__ float mel_out[FR_TILE * kMel];
  const int tid = threadIdx.x, j = tid & 15, g = tid >> 4;
  const unsigned group_mask = (tid & 16) ? 0xffff0000u : 0x0000ffffu;
  const int64_t s = blockIdx.x;
  const StreamState& S = P.st;

  for (int i = tid; i < kFFT; i += F_THREADS) sm.hann[i] = P.hann_half[i];
  for (int i = tid; i < 256; i += F_THREADS) sm.tw512[i] = P.tw512[i];
  float2 twj[16];
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) twj[k2] = P.tw256[(j * k2) & 255];

  if (tid == 0) S.n_new[s] = 0;
  if (P.is_active && P.is_active[s]) return;   // "if not context.is_active: self._sample(...)" (:139-140)

  const int np = S.n_pending[s];
  const int total = np + (int)P.n;
  float* pend = S.pending + s * S.pend_cap;
  // pending (already pre-emphasised) + new chunk -> shared memory, processed FR_TILE frames at a time
  const bool speech = P.is_speech ? (P.is_speech[s] != 0) : true;
  const int n_frames = total >= kFFT ? (total - kFFT) / kHop + 1 : 0;
  const float prev = S.prev_sample[s];
  const int16_t* chunk = P.pcm + s * P.n;
  __syncthreads();

  int head = S.ring_head[s];
  for (int f0 = 0; f0 < n_frames; f0 += FR_TILE) {
    const int nf = min(FR_TILE, n_frames - f0);
    const int ns = (nf - 1) * kHop + kFFT;
    const int base = f0 * kHop;
    for (int i = tid; i < ns; i += F_THREADS) {
      int p = base + i;
      float v;
      if (p < np) {
        v = pend[p];
      } else {
        int c = p - np;
        float x = pcm_to_float(chunk[c]);
        float xp = c > 0 ? pcm_to_float(chunk[c - 1]) : prev;
        v = (P.a != 0.0f) ? __fsub_rn(x, __fmul_rn(P.a, xp)) : x;
      }
      sm.samples[i] = v;
    }
    __syncthreads();
    if (speech) {
      if (g < nf)
        frame_spectrum(sm.samples + g * kHop, sm.hann, sm.tw512, twj, sm.xch[g], sm.mag[g], j, group_mask);
      __syncthreads();
      mel_phase(sm, P.mp, nf, mel_out);
      __syncthreads();
      // push into the ring: the window for new frame q ends at ring row (head + L + q) % ring
      for (int o = tid; o < nf * kMel; o += F_THREADS) {
        int q = o / kMel, band = o - q * kMel;
        int r = (head + P.L + f0 + q) % S.ring;
        S.mel_ring[(s * S.ring + r) * kMel + band] = mel_out[o];
      }
    }
    __syncthreads();
  }
  // bookkeeping by one thread: windows to evaluate, ring head, pending tail
  if (speech && tid == 0 && n_frames > 0) {
    int slot = atomicAdd(S.n_win, n_frames);
    for (int q = 0; q < n_frames; ++q) {
      S.win_stream[slot + q] = (int32_t)s;
      S.win_start[slot + q] = (head + 1 + q) % S.ring;
    }
    S.n_new[s] = n_frames;
    S.win_slot[s] = slot;
    S.ring_head[s] = (head + n_frames) % S.ring;
  }
  // new pending tail = samples [n_frames*hop, total)
  const int keep0 = n_frames * kHop;
  const int keep = total - keep0;
  __syncthreads();
  // move through registers to avoid overlapping read/write hazards
  float tmp[4];
  int c = 0;
  for (int i = tid; i < keep && c < 4; i += F_THREADS, ++c) {
    int p = keep0 + i;
    float v;
    if (p < np) v = pend[p];
    else {
      int cc = p - np;
      float x = pcm_to_float(chunk[cc]);
      float xp = cc > 0 ? pcm_to_float(chunk[cc - 1]) : prev;
      v = (P.a != 0.0f) ? __fsub_rn(x, __fmul_rn(P.a, xp)) : x;
    }
    tmp[c] = v;
  }
  __syncthreads();
  c = 0;
  for (int i = tid; i < keep && c < 4; i += F_THREADS, ++c) pend[i] = tmp[c];
  if (tid == 0) {
    S.n_pending[s] = keep;
    if (P.n > 0) S.prev_sample[s] = pcm_to_float(chunk[P.n - 1]);
  }
}


// After the encoder has produced win_post: per-stream trigger bookkeeping
// (wakeword/tflite.py:233-239) and the reset on a VAD fall (:135-146, :241-246).
struct StreamFinishParams {
  StreamState st;
  int L;
  const uint8_t* is_speech;
  const uint8_t* is_active;
  float threshold;
  float* post_out;        // [S, max_frames]
  int32_t* n_post_out;    // [S]
  uint8_t* trigger_out;   // [S]
  float* post_max_out;    // [S]
  int64_t n_streams;
};

__global__ void stream_finish_kernel(const StreamFinishParams P) {
  const StreamState& S = P.st;
  const int64_t s = blockIdx.x;
  if (s >= P.n_streams) return;
  __shared__ int do_reset;
  if (threadIdx.x == 0) {
    const int n = S.n_new[s];
    const int slot = n > 0 ? S.win_slot[s] : 0;
    float pm = S.post_max[s];
    bool trig = false;
    const bool active = P.is_active && P.is_active[s];
    for (int q = 0; q < S.max_frames; ++q) {
      float p = nanf("");
      if (q < n) {
        p = S.win_post[slot + q];
        if (p > pm) pm = p;
        if (p > P.threshold && !active) trig = true;
      }
      if (P.post_out) P.post_out[s * S.max_frames + q] = p;
    }
    if (P.n_post_out) P.n_post_out[s] = n;
    if (P.trigger_out) P.trigger_out[s] = trig ? 1 : 0;
    if (P.post_max_out) P.post_max_out[s] = pm;
    const bool speech = P.is_speech ? (P.is_speech[s] != 0) : true;
    const bool fall = S.was_speech[s] && !speech;
    S.was_speech[s] = speech ? 1 : 0;
    S.post_max[s] = fall ? 0.0f : pm;
    if (fall) {
      S.n_pending[s] = 0;
      S.ring_head[s] = 0;
    }
    do_reset = fall ? 1 : 0;
  }
  __syncthreads();
  if (do_reset) {
    float* ring = S.mel_ring + s * S.ring * kMel;
    for (int i = threadIdx.x; i < S.ring * kMel; i += blockDim.x) ring[i] = 0.0f;
  }
}

__global__ void stream_reset_kernel(StreamState S, const uint8_t* mask, int64_t n_streams) {
  const int64_t s = blockIdx.x;
  if (s >= n_streams || (mask && !mask[s])) return;
  float* ring = S.mel_ring + s * S.ring * kMel;
  for (int i = threadIdx.x; i < S.ring * kMel; i += blockDim.x) ring[i] = 0.0f;
  if (threadIdx.x == 0) {
    S.n_pending[s] = 0;
    S.ring_head[s] = 0;
    S.post_max[s] = 0.0f;
  }
}

// filter.tflite on magnitudes: one thread per (frame, band)
__global__ void mel_from_mag_kernel(const float* __restrict__ mag, int64_t B, MelParams P, float* __restrict__ mel) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= B * kMel) return;
  const int64_t f = o / kMel;
  const int band = (int)(o - f * kMel);
  const MelTables& mt = P.mt;
  const float* row = mag + f * kBins;
  float acc = 0.f;
  for (int s = mt.band_seg0[band]; s < mt.band_seg0[band + 1]; ++s) {
    float part = 0.f;
    const int first = mt.seg_first[s], cnt = mt.seg_count[s];
    for (int t = 0; t < cnt; ++t) part = fmaf(mt.tap_w[first + t], row[mt.tap_bin[first + t]], part);
    acc += part;
  }
  acc += mt.bias[band];
  acc = fmaxf(acc, P.mel_floor);
  mel[o] = __fmul_rn(__fsub_rn(logf(acc), P.mel_log_offset), P.mel_scale);
}

int launch_mel_from_mag(wwb_ctx* ctx, const float* mag, int64_t B, float* mel, cudaStream_t st) {
  if (B == 0) return WWB_OK;
  MelParams P;
  P.mt = ctx->mel;
  P.mel_floor = ctx->mel_floor; P.mel_log_offset = ctx->mel_log_offset; P.mel_scale = ctx->mel_scale;
  mel_from_mag_kernel<<<(unsigned)((B * kMel + 255) / 256), 256, 0, st>>>(mag, B, P, mel);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

int launch_stream_filter(wwb_ctx* ctx, const int16_t* pcm, int64_t S, int64_t n,
                         const uint8_t* is_speech, const uint8_t* is_active, float a,
                         cudaStream_t st) {
  StreamFilterParams P;
  P.pcm = pcm; P.n = n; P.is_speech = is_speech; P.is_active = is_active; P.a = a;
  P.st = ctx->st; P.L = ctx->L;
  P.hann_half = ctx->hann; P.tw256 = ctx->tw256; P.tw512 = ctx->tw512;
  P.mp.mt = ctx->mel;
  P.mp.mel_floor = ctx->mel_floor; P.mp.mel_log_offset = ctx->mel_log_offset; P.mp.mel_scale = ctx->mel_scale;
  WWB_CUDA(ctx, cudaMemsetAsync(ctx->st.n_win, 0, sizeof(int32_t), st));
  size_t smem = sizeof(FilterSmem);
  WWB_CUDA(ctx, cudaFuncSetAttribute(stream_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stream_filter_kernel<<<(unsigned)S, F_THREADS, smem, st>>>(P);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

int launch_stream_finish(wwb_ctx* ctx, int64_t S, const uint8_t* is_speech, const uint8_t* is_active,
                         float threshold, float* post_out, int32_t* n_post_out, uint8_t* trigger_out,
                         float* post_max_out, cudaStream_t st) {
  StreamFinishParams P;
  P.st = ctx->st; P.L = ctx->L; P.is_speech = is_speech; P.is_active = is_active;
  P.threshold = threshold; P.post_out = post_out; P.n_post_out = n_post_out;
  P.trigger_out = trigger_out; P.post_max_out = post_max_out; P.n_streams = S;
  stream_finish_kernel<<<(unsigned)S, 128, 0, st>>>(P);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

int launch_stream_reset(wwb_ctx* ctx, const uint8_t* mask, int64_t S, cudaStream_t st) {
  stream_reset_kernel<<<(unsigned)S, 128, 0, st>>>(ctx->st, mask, S);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

}  // namespace wwb
