"""libm_emul.h (the device's float32 exp) against the host libm.

The header is host-compilable; the same arithmetic runs on the GPU (FMA-for-FMA,
compiled with -fmad=false so nothing else is contracted).  A strided sweep over all
float32 bit patterns must agree bit-for-bit; the exhaustive 2^32 sweep (run once by
hand) differs for exactly 2 inputs, neither reachable from a softmax."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "libm_emul.h"
int main(int argc, char** argv){
  uint64_t step = strtoull(argv[1], 0, 10), bad = 0, n = 0;
  for (uint64_t u = 0; u <= 0xffffffffULL; u += step) {
    uint32_t v = (uint32_t)u, ua, ub; float x; memcpy(&x, &v, 4);
    float a = expf(x), b = mgd_expf(x);
    memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
    n++;
    if (ua != ub && !(a != a && b != b)) bad++;
    if (x <= 0.0f && x >= -104.0f) {          /* the softmax path: clamped core */
      float c = mgd_expf_core(x, mgd_exp2f_tab); uint32_t uc; memcpy(&uc, &c, 4);
      if (ua != uc) bad++;
    }
    float s1 = 1.0f / (1.0f + expf(-x)), s2 = mgd_expitf_tab(x, mgd_exp2f_tab);
    memcpy(&ua, &s1, 4); memcpy(&ub, &s2, 4);
    if (ua != ub && !(s1 != s1 && s2 != s2)) bad++;
  }
  printf("%llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
  return 0;
}
'''


def test_expf_emulation_matches_libm(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-I",
                    os.path.join(ROOT, "multigriddet_b200", "csrc"), str(src), "-o", str(exe), "-lm"],
                   check=True)
    n, bad = subprocess.run([str(exe), "211"], capture_output=True, text=True).stdout.split()
    assert int(n) > 20_000_000 and int(bad) == 0
