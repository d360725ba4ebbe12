"""CPU (needs nvcc, no GPU): the parity-critical device functions do not depend on the compiler's multiply-add
contraction.  Bit-exact OHEM masks rest on `neg_class_score` (csrc/common.cuh) and bit-exact box corners on the
rotating calipers of csrc/rect.cuh reproducing the reference's / OpenCV's rounding sequence; they are written
with explicit round-to-nearest intrinsics instead of compiling the library with -fmad=false (which would slow the
loss terms down for nothing).  This test compiles a probe that instantiates exactly those functions with the
library's flags, with and without -fmad=false, and requires the two SASS listings to be identical: an edit that
introduces a contractible `a * b + c` into them shows up as a difference."""
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tensorflow_ocr_b200", "csrc")

PROBE = r'''
#include "common.cuh"
#include "rect.cuh"
namespace plh { std::atomic<long long> g_launch_count{0}; }
using namespace plh;
extern "C" __global__ void probe_score(const float2* x, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = neg_class_score(x[i].x, x[i].y);
}
extern "C" __global__ void __launch_bounds__(256) probe_rect(const int* pts, int total, int npad, int* box, float* rect) {
  extern __shared__ __align__(16) unsigned char smem[];
  RectSmem S = rect_carve(smem, npad);
  int nsort = 32;
  while (nsort < total) nsort <<= 1;
  for (int i = threadIdx.x; i < nsort; i += blockDim.x) S.keys[i] = i < total ? make_key(pts[2 * i], pts[2 * i + 1], i) : ~0ull;
  __syncthreads();
  bitonic_sort(S.keys, nsort);
  int b[8];
  float r[5];
  min_area_box_sorted(S, total, npad, b, r);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) box[i] = b[i];
    for (int i = 0; i < 5; ++i) rect[i] = r[i];
  }
}
extern "C" __global__ void __launch_bounds__(256) probe_rect_distinct(int total, int npad, int* box, float* rect) {
  extern __shared__ __align__(16) unsigned char smem[];
  RectSmem S = rect_carve(smem, npad);
  int b[8];
  float r[5];
  min_area_box_distinct(S, total, b, r);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) box[i] = b[i];
    for (int i = 0; i < 5; ++i) rect[i] = r[i];
  }
}
'''


def _sass(flags, workdir, tag):
    src = os.path.join(workdir, "probe.cu")
    obj = os.path.join(workdir, "probe_%s.cubin" % tag)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I", CSRC, "-cubin", src, "-o", obj] + flags
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    # instruction text only (drop addresses / encodings)
    return [re.sub(r"/\*[0-9a-fx ]+\*/", "", l).strip() for l in out.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_parity_critical_functions_have_no_contractible_fma():
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "probe.cu"), "w").write(PROBE)
        a = _sass([], td, "default")
        b = _sass(["-fmad=false"], td, "nofmad")
    assert len(a) > 500, "probe did not compile to what it should"
    assert a == b, "SASS differs with -fmad=false: a parity-critical function contains a contractible multiply-add"
