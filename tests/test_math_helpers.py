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


SINCOS_SRC = r'''
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "mppi-generic_b200/csrc/device_utils.cuh"
static double ulp_err(float got, double ref) {
  float r = (float)ref; int e; frexpf(r == 0 ? 1e-30f : r, &e);
  return fabs((double)got - ref) / ldexp(1.0, e - 24);
}
int main() {
  double ms = 0, mc = 0; srand(3); long n = 0;
  for (long i = 0; i < 20000000; i++) {
    float x = ((float)rand() / RAND_MAX * 2 - 1) * (i % 3 == 0 ? 48039.f : (i % 3 == 1 ? 100.f : 3.2f));
    float s, c; mppib::sincos_cw(x, &s, &c);
    ms = fmax(ms, ulp_err(s, sin((double)x))); mc = fmax(mc, ulp_err(c, cos((double)x))); n++;
  }
  // beyond the Cody-Waite range the library call takes over
  float s, c; mppib::sincos_cw(1.0e6f, &s, &c);
  ms = fmax(ms, ulp_err(s, sin(1.0e6))); mc = fmax(mc, ulp_err(c, cos(1.0e6)));
  printf("%ld %.4f %.4f\n", n, ms, mc);
  return 0;
}
'''


def test_sincos_cw_is_within_1p5_ulp():
    """device_utils.cuh: sincos_cw (one range reduction for the sine and cosine of the Autorally kinematics) against
    float64 libm on 2e7 arguments up to the Cody-Waite bound."""
    with tempfile.TemporaryDirectory() as d:
        cu = os.path.join(d, "t.cu")
        open(cu, "w").write(SINCOS_SRC)
        exe = os.path.join(d, "t")
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-I", ROOT, "-Xcompiler", "-ffp-contract=off", cu, "-o", exe])
        n, ms, mc = subprocess.check_output([exe]).split()
    assert int(n) == 20_000_000 and float(ms) < 1.5 and float(mc) < 1.5, (n, ms, mc)
