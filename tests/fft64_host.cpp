// CPU harness for wakeword_detection_b200/csrc/fft64.cuh: runs the SAME per-lane code the K1 kernel runs, with the 16
// lanes of a frame emulated in lock step (shared-memory transpose = an array, shuffle = a read of the partner lane's
// registers), so the index algebra of the fp64 FFT is checked against numpy without a GPU (tests/test_host.py).
//   fft64_host < frames.f32 > mags.f32      (n x 512 float32 in, n x 257 float32 out)
#include <math.h>
#include <stdio.h>
#include <vector>

#include "../wakeword_detection_b200/csrc/fft64.cuh"

using namespace wwb::f64;

int main() {
  const double PI = 3.14159265358979323846;
  std::vector<float> in;
  float buf[512];
  while (fread(buf, sizeof(float), 512, stdin) == 512) in.insert(in.end(), buf, buf + 512);
  const size_t n = in.size() / 512;
  std::vector<double2> hann2(256), twj(256), tw512(256);
  for (int n = 0; n < 256; ++n) {
    hann2[n] = make_double2(0.5 * (0.5 - 0.5 * cos(2.0 * PI * (2 * n) / 511.0)), 0.5 * (0.5 - 0.5 * cos(2.0 * PI * (2 * n + 1) / 511.0)));
    tw512[n] = make_double2(cos(2 * PI * n / 512), -sin(2 * PI * n / 512));
  }
  for (int j = 0; j < 16; ++j)
    for (int k2 = 0; k2 < 16; ++k2) {
      const int e = (j * k2) & 255;
      twj[16 * j + k2] = make_double2(cos(2 * PI * e / 256), -sin(2 * PI * e / 256));
    }
  std::vector<double2> xch(16 * XP);
  for (size_t f = 0; f < n; ++f) {
    const float* x = &in[f * 512];
    for (int j = 0; j < 16; ++j) {
      auto load = [&](int m) { return make_double2((double)x[2 * (j + 16 * m)], (double)x[2 * (j + 16 * m) + 1]); };
      pass1(load, hann2.data(), twj.data(), xch.data(), j);
    }
    double2 v[16][16];
    for (int j = 0; j < 16; ++j) pass2(xch.data(), j, v[j]);
    float mag[257];
    for (int i = 0; i < 257; ++i) mag[i] = -1.f;
    // the kernel's pairing (filter.cu frame_spectrum64): lane j, slot k1 = 0..7: bins k = 16 k1 + j and 256 - k
    for (int j = 0; j < 16; ++j) {
      const int pl = (16 - j) & 15;
      for (int k1 = 0; k1 < 8; ++k1) {
        const double2 zk = v[j][XI(k1)];
        // what lane pl sends for slot k1: its upper half, except lane 0 (self-paired: Z[16 (16 - k1)], Z[256] = Z[0])
        const double2 zp = pl == 0 ? (k1 == 0 ? v[0][XI(0)] : v[0][XI(16 - k1)]) : v[pl][XI(15 - k1)];
        const int k = 16 * k1 + j;
        split_pair(zk, zp, tw512[k], mag[k], mag[256 - k]);
      }
    }
    {  // bin 128 (lane 0, k1 = 8) pairs with itself
      float a, b;
      split_pair(v[0][XI(8)], v[0][XI(8)], tw512[128], a, b);
      mag[128] = a;
    }
    fwrite(mag, sizeof(float), 257, stdout);
  }
  return 0;
}
