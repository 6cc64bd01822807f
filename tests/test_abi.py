"""The C-ABI library loads and exports every symbol include/derl_b200.h declares; argument
validation and the no-device failure mode work without a GPU.  No compute is performed."""
import ctypes
import os
import re

import pytest
import torch

from derl_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
  with open(os.path.join(REPO, "include", "derl_b200.h")) as f:
    text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
  return sorted(set(re.findall(r"\b(derl_b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  lib = ctypes.CDLL(_lib.LIB_PATH)
  names = declared_symbols()
  assert len(names) >= 15
  for name in names:
    assert hasattr(lib, name), f"{name} declared in include/derl_b200.h but not exported"
  assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"


def declared_prototypes():
  """name -> list of C parameter declarations, parsed from the header."""
  with open(os.path.join(REPO, "include", "derl_b200.h")) as f:
    text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
  protos = {}
  for ret, name, params in re.findall(r"\b([\w\s\*]+?)\b(derl_b200_\w+)\s*\(([^)]*)\)\s*;", text):
    params = " ".join(params.split())
    protos[name] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
  return protos


def test_ctypes_signatures_match_the_header_prototypes():
  """Same arity and same argument classes (pointer / 64-bit / int / double / size_t) in the
  ctypes table as in include/derl_b200.h: an ABI change must touch both (and ABI_VERSION)."""
  def classify(decl):
    if "*" in decl:
      return ctypes.c_void_p
    kind = decl.rsplit(" ", 1)[0].replace("const ", "").strip()
    return {"int64_t": ctypes.c_int64, "int": ctypes.c_int, "double": ctypes.c_double,
            "size_t": ctypes.c_size_t, "float": ctypes.c_float}[kind]
  protos = declared_prototypes()
  assert sorted(protos) == sorted(_lib.SIGNATURES)
  for name, params in protos.items():
    _, argtypes = _lib.SIGNATURES[name]
    assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters"
    for i, (decl, argtype) in enumerate(zip(params, argtypes)):
      want = classify(decl)
      same = argtype is want or (want is ctypes.c_void_p and argtype in (ctypes.c_void_p,
                                                                        ctypes.c_char_p))
      assert same, f"{name} argument {i} ({decl!r}): ctypes table says {argtype.__name__}"
  with open(os.path.join(REPO, "include", "derl_b200.h")) as f:
    assert f"#define DERL_B200_ABI_VERSION {_lib.ABI_VERSION}\n" in f.read()


def test_abi_version_and_error_string():
  lib = _lib.load()
  assert lib.derl_b200_abi_version() == _lib.ABI_VERSION
  assert isinstance(lib.derl_b200_last_error(), bytes)


def test_argument_validation_needs_no_device():
  lib = _lib.load()
  null = ctypes.c_void_p(None)
  rc = lib.derl_b200_gae(null, 0, null, null, null, 4, 4, 0.99, 0.95, null, null, null, null, 0,
                         0, null)
  assert rc == 1 and b"null pointer" in lib.derl_b200_last_error()
  rc = lib.derl_b200_gae(null, 0, null, null, null, 0, 4, 0.99, 0.95, null, null, null, null, 0,
                         0, null)
  assert rc == 1 and b"T >= 1" in lib.derl_b200_last_error()
  rc = lib.derl_b200_gather_rows(null, 1, 16, null, 0, 1, null, null)
  assert rc == 1
  assert lib.derl_b200_gae_workspace_bytes(128, 4096) >= 16 + 16 * 128
  assert lib.derl_b200_ppo_loss_workspace_bytes(1 << 17) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_device_fails_loudly_instead_of_falling_back():
  lib = _lib.load()
  assert lib.derl_b200_device_ok() == 2  # DERL_E_NO_DEVICE
  assert b"no CPU fallback" in lib.derl_b200_last_error()
  import derl_b200  # noqa: F401  registers the ops
  x = torch.zeros(4, 4)
  with pytest.raises(NotImplementedError):
    torch.ops.derl_b200.gae(x, x, x.bool(), x[0], 0.99, 0.95)
  with pytest.raises(NotImplementedError):
    torch.ops.derl_b200.gather_rows(x, torch.arange(4), 0, 4)
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    _lib.check(lib.derl_b200_device_ok(), "device_ok")


def test_plain_c_caller_links_and_runs(tmp_path):
  """tests/abi_smoke.c compiled with gcc against include/derl_b200.h + libderl_b200.so: on a
  B200 it must report bit-identical GAE, without a GPU it must fail with DERL_E_NO_DEVICE."""
  import subprocess
  exe = tmp_path / "abi_smoke"
  pkg = os.path.join(REPO, "derl_b200")
  subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-std=c99",
                  os.path.join(REPO, "tests", "abi_smoke.c"), "-I", os.path.join(REPO, "include"),
                  "-L", pkg, "-lderl_b200", f"-Wl,-rpath,{pkg}", "-o", str(exe)], check=True)
  proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
  if torch.cuda.is_available():
    assert proc.returncode == 0 and "bit-identical" in proc.stdout, proc.stdout + proc.stderr
  else:
    assert proc.returncode == 2 and "no CPU fallback" in proc.stdout, proc.stdout + proc.stderr


def test_product_never_imports_the_oracle():
  """oracle/ is test infrastructure: nothing under derl_b200/ may reference it."""
  pkg = os.path.join(REPO, "derl_b200")
  for root, _, files in os.walk(pkg):
    for name in files:
      if name.endswith((".py", ".cu", ".cuh", ".h")):
        with open(os.path.join(root, name)) as f:
          text = f.read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), name
        assert "liboracle" not in text, name
