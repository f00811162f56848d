// fp64 building blocks of the K1 filter kernel: 512-point real FFT of one frame by 16 cooperating lanes.
//
// Why fp64: the reference multiplies the f32 frame by np.hanning (f64) and runs np.fft.rfft in f64
// (spokestack/wakeword/tflite.py:174-176 == utils/tf_lite/filter.py:63-65).  The parity bound is on the
// LOG-mel value (|d| <= 1e-4 max(|ref|, 1)), i.e. 2e-4 RELATIVE on every band whose energy is above the 1e-5
// floor - including bands 100 dB below the frame's loudest bin.  An fp32 FFT has a rounding floor of ~1e-7 of the
// frame's PEAK (measured in round 1: 4.6e-4 log-mel error on a clipping sine), so it cannot meet the bound on
// high-dynamic-range frames; fp64 butterflies (B200: 64 DFMA/clk/SM) do, with 1e-12 to spare.
//
// Decomposition (same as the round-1 fp32 kernel): the 512 real samples are packed into 256 complex points
// z[n] = x[2n] + i x[2n+1]; n = n1 + 16 n2, k = 16 k1 + k2:
//   pass 1  lane j = n1 : 16-point FFT over n2 (registers)  -> Y[n1][k2] * W256^(n1 k2)
//   transpose through shared memory (16-byte elements, pitch 17: conflict-free per quarter-warp)
//   pass 2  lane j = k2 : 16-point FFT over n1              -> Z[16 k1 + j]
//   split   X[k] = E + W512^k O,  E = (Z[k] + conj Z[256-k]) / 2,  O = -i (Z[k] - conj Z[256-k]) / 2
//           X[256-k] = conj(E - W512^k O) comes from the same products, so lane j handles the pairs (k, 256-k) of its
//           bins k < 128 and fetches Z[256-k] from lane 16-j with one shuffle per value (1/2 folded into the window)
// The file is plain C++ apart from the shuffle, so tests/fft64_host.cpp can run the same code on the CPU
// (lanes emulated in lock step) against numpy.
#pragma once

#if defined(__CUDACC__)
#define F64_HD __host__ __device__ __forceinline__
#else
#define F64_HD inline
struct double2 { double x, y; };
struct float2 { float x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

namespace wwb {
namespace f64 {

constexpr int XP = 17;   // transpose pitch (double2 elements)

F64_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
F64_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
F64_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward radix-4 butterfly (W4 = -i), in place, natural output order
F64_HD void bfly4(double2& a0, double2& a1, double2& a2, double2& a3) {
  const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = make_double2(d02.x + d13.y, d02.y - d13.x);
  a3 = make_double2(d02.x - d13.y, d02.y + d13.x);
}

// second half of fft16: W16 twiddles + the four output butterflies.  X[k] is left in v[4*(k&3) + (k>>2)] (XI()).
F64_HD void fft16_tail(double2 (&v)[16]) {
  constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178, R = 0.70710678118654752440;
  v[1 + 4] = cmul(v[1 + 4], make_double2(C1, -S1));                                                   // W^1
  v[1 + 8] = make_double2((v[1 + 8].x + v[1 + 8].y) * R, (v[1 + 8].y - v[1 + 8].x) * R);              // W^2
  v[1 + 12] = cmul(v[1 + 12], make_double2(S1, -C1));                                                 // W^3
  v[2 + 4] = make_double2((v[2 + 4].x + v[2 + 4].y) * R, (v[2 + 4].y - v[2 + 4].x) * R);              // W^2
  v[2 + 8] = make_double2(v[2 + 8].y, -v[2 + 8].x);                                                   // W^4 = -i
  v[2 + 12] = make_double2((v[2 + 12].y - v[2 + 12].x) * R, -(v[2 + 12].x + v[2 + 12].y) * R);        // W^6
  v[3 + 4] = cmul(v[3 + 4], make_double2(S1, -C1));                                                   // W^3
  v[3 + 8] = make_double2((v[3 + 8].y - v[3 + 8].x) * R, -(v[3 + 8].x + v[3 + 8].y) * R);             // W^6
  v[3 + 12] = cmul(v[3 + 12], make_double2(-C1, S1));                                                 // W^9
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) bfly4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// in-register forward DFT of 16 points, natural input order
F64_HD void fft16(double2 (&v)[16]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) bfly4(v[a], v[a + 4], v[a + 8], v[a + 12]);
  fft16_tail(v);
}
F64_HD constexpr int XI(int k) { return 4 * (k & 3) + (k >> 2); }

// First radix-4 layer of pass 1 with the window multiply fused into it: a_i = x_i * h_i never exists on its own,
// s = x0 h0 + x2 h2 and d = x0 h0 - x2 h2 are one multiply and two FMAs (3 instead of 4 fp64 instructions per pair).
// load(m) returns the lane's sample pair m (n = j + 16 m) as doubles; the pairs are fetched butterfly by butterfly so
// that at most 4 of them are live next to the 16 results.
template <typename LoadFn>
F64_HD void bfly4_windowed(LoadFn load, const double2* hann2, int j, int a, double2 (&v)[16]) {
  const double2 x0 = load(a), x1 = load(a + 4), x2 = load(a + 8), x3 = load(a + 12);
  const double2 h0 = hann2[j + 16 * a], h1 = hann2[j + 16 * (a + 4)], h2 = hann2[j + 16 * (a + 8)], h3 = hann2[j + 16 * (a + 12)];
  const double p0x = x0.x * h0.x, p0y = x0.y * h0.y;
  const double p1x = x1.x * h1.x, p1y = x1.y * h1.y;
  const double2 s02 = make_double2(x2.x * h2.x + p0x, x2.y * h2.y + p0y);
  const double2 d02 = make_double2(p0x - x2.x * h2.x, p0y - x2.y * h2.y);
  const double2 s13 = make_double2(x3.x * h3.x + p1x, x3.y * h3.y + p1y);
  const double2 d13 = make_double2(p1x - x3.x * h3.x, p1y - x3.y * h3.y);
  v[a] = cadd(s02, s13);
  v[a + 8] = csub(s02, s13);
  v[a + 4] = make_double2(d02.x + d13.y, d02.y - d13.x);
  v[a + 12] = make_double2(d02.x - d13.y, d02.y + d13.x);
}

// Tables (shared memory in the kernel; both 16-lane halves of a warp read the same addresses, so a warp-wide
// 16-byte load costs two wavefronts instead of four):
//   hann2[n] = 0.5 * (np.hanning(512)[2n], np.hanning(512)[2n+1])     n = 0..255   (the window of sample pair n)
//   twj[16 j + k2] = W256^(j k2)
// pass 1 for lane j: load(m) = the lane's sample pair m -> xch[j][k2] = Y[j][k2] W256^(j k2)
template <typename LoadFn>
F64_HD void pass1(LoadFn load, const double2* hann2, const double2* twj, double2* xch, int j) {
  double2 v[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int a = 0; a < 4; ++a) bfly4_windowed(load, hann2, j, a, v);
  fft16_tail(v);
  xch[j * XP + 0] = v[XI(0)];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k2 = 1; k2 < 16; ++k2) xch[j * XP + k2] = cmul(v[XI(k2)], twj[16 * k2 + j]);   // (the table is symmetric; this index is conflict-free across lanes)
}

// pass 2 for lane j: Z[16 k1 + j] in v[XI(k1)]
F64_HD void pass2(const double2* xch, int j, double2 (&v)[16]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int n1 = 0; n1 < 16; ++n1) v[n1] = xch[n1 * XP + j];
  fft16(v);
}

F64_HD float mag_f32(double xr, double xi) {
  // fp64 up to |X|^2; the square root is taken in fp32 (the reference casts |X| to f32): ~2 ulp of f32, 2e-7 relative
  const float m2 = (float)(xr * xr + xi * xi);
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m2));
  return r;
#else
  return __builtin_sqrtf(m2);
#endif
}

// Real-FFT split of the bin pair (k, 256 - k) from Z[k] (zk), Z[256-k] (zp) and w = W512^k:
//   E = Zk + conj(Zp), O = -i (Zk - conj(Zp))   (the 1/2 is folded into the window),  t = w O
//   X[k] = E + t,   X[256-k] = conj(E - t)      (E and O of bin 256-k are the conjugates, W512^(256-k) = -conj(w))
// 16 fp64 instructions for two magnitudes.
F64_HD void split_pair(double2 zk, double2 zp, double2 w, float& mag_k, float& mag_nk) {
  const double er = zk.x + zp.x, ei = zk.y - zp.y;
  const double orr = zk.y + zp.y, oi = zp.x - zk.x;
  const double tr = w.x * orr - w.y * oi, ti = w.x * oi + w.y * orr;
  mag_k = mag_f32(er + tr, ei + ti);
  mag_nk = mag_f32(er - tr, ei - ti);
}

}  // namespace f64
}  // namespace wwb
