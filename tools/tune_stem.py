"""Build and time variants of the K6 stem kernel (warps x tiles per warp x tap-row unroll).

  python tools/tune_stem.py build      # here (no GPU): nvcc each variant into build_variants/
  python tools/tune_stem.py            # on the GPU box: time every variant on 32768 frames

Variants are separate shared objects (abi.cu + stem.cu with -D knobs) loaded through ctypes, so
the shipped library is untouched; outputs are checked bit-identical to the default variant.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build_variants")
VARIANTS = [(9, 3, 2), (9, 3, 1), (9, 3, 4), (9, 3, 8), (13, 2, 2), (13, 2, 4), (13, 2, 8),
            (25, 1, 2), (25, 1, 8), (7, 4, 2)]


def path(v):
  return os.path.join(OUT, "libstem_w%d_t%d_u%d.so" % v)


def build():
  sys.path.insert(0, ROOT)
  from derl_b200 import build as b
  os.makedirs(OUT, exist_ok=True)
  for v in VARIANTS:
    cmd = [b.find_nvcc(), *b.NVCC_FLAGS, "-I", b.INCLUDE, "-I", b.CSRC, "-Xptxas", "-v",
           "-DDERL_STEM_WARPS=%d" % v[0], "-DDERL_STEM_TILES=%d" % v[1],
           "-DDERL_STEM_UNROLL=%d" % v[2], os.path.join(b.CSRC, "abi.cu"),
           os.path.join(b.CSRC, "stem.cu"), "-o", path(v)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    regs = [l for l in proc.stderr.splitlines() if "registers" in l or "spill" in l]
    print(v, "ok" if proc.returncode == 0 else "FAILED", "|", " ".join(regs[-2:])[:160], flush=True)


def run():
  import torch
  frames = torch.randint(0, 256, (32768, 84, 84, 4), device="cuda", dtype=torch.uint8)
  weight = torch.randn(32, 4, 8, 8, device="cuda") * 0.1
  bias = torch.randn(32, device="cuda") * 0.1
  want = None
  stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
  P = ctypes.c_void_p
  for v in VARIANTS:
    if not os.path.exists(path(v)):
      continue
    lib = ctypes.CDLL(path(v))
    fn = lib.derl_b200_stem_conv_relu
    fn.argtypes = [P, P, ctypes.c_int64, P, P, P, ctypes.c_int, ctypes.c_int, P]
    out = torch.empty((32768, 10, 10, 128), device="cuda")
    times = []
    for _ in range(5):
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      rc = fn(frames.data_ptr(), None, 32768, weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
              0, 2, stream)
      e.record()
      torch.cuda.synchronize()
      assert rc == 0, rc
      times.append(s.elapsed_time(e))
    if want is None:
      want = out.clone()
    print("warps %2d tiles %d unroll %d: %.4f ms  identical=%s" %
          (*v, min(times[1:]), bool(torch.equal(out, want))), flush=True)


if __name__ == "__main__":
  build() if sys.argv[1:] == ["build"] else run()
