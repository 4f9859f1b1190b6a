"""CPU unit test of the radix-8 Stockham FFT index/twiddle logic used by csrc/mel.cu: the header is
compiled as plain C++ and driven by a harness that emulates the kernel's gather/barrier/scatter rounds."""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = r"""
#include <cstdio>
#include <cmath>
#include <vector>
#include "fft_core.cuh"
int main() {
  std::vector<float> x(1024);
  for (int i = 0; i < 1024; ++i) if (scanf("%f", &x[i]) != 1) return 1;
  std::vector<float2> tw(1024), z(512), regs(512 * 8);
  for (int m = 0; m < 512; ++m) {
    tw[m] = make_float2((float)cos(2 * M_PI * m / 512), (float)-sin(2 * M_PI * m / 512));
    tw[512 + m] = make_float2((float)cos(2 * M_PI * m / 1024), (float)-sin(2 * M_PI * m / 1024));
  }
  for (int n = 0; n < 512; ++n) z[n] = make_float2(x[2 * n], x[2 * n + 1]);
  for (int Ns = 1; Ns <= 64; Ns *= 8) {
    std::vector<float2> out(512);
    for (int j = 0; j < 64; ++j) {
      float2 v[8];
      for (int r = 0; r < 8; ++r) v[r] = z[j + 64 * r];
      stockham8_compute(v, j, Ns, tw.data());
      for (int r = 0; r < 8; ++r) out[stockham8_dst(j, Ns, r)] = v[r];
    }
    z = out;
  }
  for (int k = 0; k <= 512; ++k) {
    float2 X = (k < 512) ? rfft_split(z[k], z[(512 - k) & 511], tw[512 + k]) : make_float2(z[0].x - z[0].y, 0.f);
    printf("%.9g %.9g\n", X.x, X.y);
  }
  return 0;
}
"""


def test_stockham_rfft_matches_numpy():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "h.cpp")
        open(src, "w").write(HARNESS)
        exe = os.path.join(d, "h")
        subprocess.run(["g++", "-O2", "-x", "c++", "-I", os.path.join(ROOT, "efficientat_b200", "csrc"), src, "-o", exe],
                       check=True)
        rng = np.random.default_rng(0)
        for trial in range(3):
            x = rng.standard_normal(1024).astype(np.float32)
            out = subprocess.run([exe], input=" ".join(f"{v:.9g}" for v in x), capture_output=True, text=True,
                                 check=True).stdout
            got = np.array([[float(a) for a in line.split()] for line in out.strip().splitlines()])
            ref = np.fft.rfft(x.astype(np.float64))
            err = np.abs(got[:, 0] + 1j * got[:, 1] - ref).max()
            assert err < 2e-4 * np.abs(ref).max(), err
