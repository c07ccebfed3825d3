"""Host-side checks of the exact / near-exact math helpers the kernels substitute for slower library calls
(mppi-generic_b200/csrc/device_utils.cuh). They are __host__ __device__, so nvcc builds a CPU program from the same
source the kernels use; no GPU needed."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "mppi-generic_b200/csrc/device_utils.cuh"
// reference form, utils/angle_utils.cuh:20-26
static float normalizeAngle_ref(float angle) {
  const float result = fmodf(angle + MPPIB_PI_F, 2.0f * MPPIB_PI_F);
  if (result <= 0.0f) return result + MPPIB_PI_F;
  return result - MPPIB_PI_F;
}
int main() {
  const float two_pi = 2.0f * MPPIB_PI_F;
  long bad = 0, n = 0;
  srand(1);
  for (long i = 0; i < 40000000L; i++) {
    unsigned u = ((unsigned)rand() << 16) ^ rand() ^ ((unsigned)rand() << 1);
    float a; memcpy(&a, &u, 4);
    if (!std::isfinite(a) || fabsf(a) > 1.0e6f) continue;
    n++;
    float x = mppib::fmod_2pi_exact(a), y = fmodf(a, two_pi);
    if (memcmp(&x, &y, 4) && !(x == 0 && y == 0)) bad++;
    float p = mppib::normalizeAngle(a), q = normalizeAngle_ref(a);
    if (memcmp(&p, &q, 4) && !(p == 0 && q == 0)) bad++;
  }
  for (int k = -2000; k <= 2000; k++)
    for (int d = -50; d <= 50; d++) {
      float a = k * two_pi; unsigned u; memcpy(&u, &a, 4); u += d; memcpy(&a, &u, 4);
      if (!std::isfinite(a)) continue;
      n++;
      float x = mppib::fmod_2pi_exact(a), y = fmodf(a, two_pi);
      if (memcmp(&x, &y, 4) && !(x == 0 && y == 0)) bad++;
    }
  printf("%ld %ld\n", n, bad);
  return 0;
}
'''


def test_fmod_2pi_and_normalize_angle_are_bit_identical_to_fmodf():
    with tempfile.TemporaryDirectory() as d:
        cu = os.path.join(d, "t.cu")
        open(cu, "w").write(SRC)
        exe = os.path.join(d, "t")
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-I", ROOT, "-Xcompiler", "-ffp-contract=off", cu, "-o", exe])
        n, bad = map(int, subprocess.check_output([exe]).split())
    assert n > 20_000_000 and bad == 0, (n, bad)
