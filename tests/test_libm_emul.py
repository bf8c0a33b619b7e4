"""libm_emul.h (the device's float32 exp) against the host libm.

The header is host-compilable; the same arithmetic runs on the GPU (FMA-for-FMA,
compiled with -fmad=false so nothing else is contracted).  Every one of the 2^32
float32 bit patterns must agree bit-for-bit with the host's expf (glibc selects its FMA
build on every x86-64 CPU with FMA3, which the header follows), for expf itself, for the
clamped core the softmax uses and for the sigmoid built on it.  The sweep runs on all
host cores (pthreads); ``MGD_EXPF_SWEEP_STEP`` thins it out on a slow machine."""
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
#include <pthread.h>
typedef struct { uint64_t lo, hi, step, n, bad; } job_t;
static void* run(void* p);
int main(int argc, char** argv){
  uint64_t step = strtoull(argv[1], 0, 10);
  int nt = atoi(argv[2]); if (nt < 1) nt = 1; if (nt > 256) nt = 256;
  pthread_t th[256]; job_t jobs[256];
  uint64_t per = (0x100000000ULL / step + nt) / nt * step;          /* a multiple of step */
  for (int t = 0; t < nt; ++t) {
    jobs[t].lo = (uint64_t)t * per; jobs[t].hi = jobs[t].lo + per; jobs[t].step = step;
    if (jobs[t].hi > 0x100000000ULL) jobs[t].hi = 0x100000000ULL;
    pthread_create(&th[t], 0, run, &jobs[t]);
  }
  uint64_t n = 0, bad = 0;
  for (int t = 0; t < nt; ++t) { pthread_join(th[t], 0); n += jobs[t].n; bad += jobs[t].bad; }
  printf("%llu %llu\n", (unsigned long long)n, (unsigned long long)bad);
  return 0;
}
static void* run(void* p){
  job_t* j = (job_t*)p; uint64_t step = j->step, bad = 0, n = 0;
  for (uint64_t u = j->lo; u < j->hi; u += step) {
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
  j->n = n; j->bad = bad;
  return 0;
}
'''


def test_expf_emulation_matches_libm(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-pthread", "-I",
                    os.path.join(ROOT, "multigriddet_b200", "csrc"), str(src), "-o", str(exe), "-lm"],
                   check=True)
    step = int(os.environ.get("MGD_EXPF_SWEEP_STEP", "1"))
    n, bad = subprocess.run([str(exe), str(step), str(os.cpu_count() or 1)], capture_output=True,
                            text=True, timeout=900).stdout.split()
    assert int(n) == (2 ** 32 + step - 1) // step            # step 1: all 4 294 967 296 inputs
    assert int(bad) == 0
