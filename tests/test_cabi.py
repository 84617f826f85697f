"""The C-ABI library builds, loads without a GPU and exports every symbol include/diffpir_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "diffpir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpir_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from diffpir_b200 import _lib
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound but not declared in include/diffpir_b200.h"
    assert lib.dpir_version() >= 100


def test_sm100a_sass_contains_blackwell_instructions():
    """cross-compiled cubin must really carry tcgen05 / TMA code (UTCHMMA, UTMALDG, LDTM)."""
    import shutil
    import subprocess
    from diffpir_b200 import _lib
    _lib.load()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffpir_b200 import arch
    from diffpir_b200.unet import UNetModel
    with pytest.raises(RuntimeError):
        UNetModel(arch.ffhq_config())
    from diffpir_b200 import ops
    with pytest.raises(RuntimeError):
        ops.affine(torch.zeros(4), 1.0, 0.0)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: only tests/, smoke() and bench.py's CPU legs may execute it - never the package, nor tools/."""
    for sub in ("diffpir_b200", "tools"):
        pkg = os.path.join(ROOT, sub)
        for fn in os.listdir(pkg):
            if fn.endswith(".py"):
                src = open(os.path.join(pkg, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), (sub, fn)


def test_shipped_library_matches_sources():
    """The in-tree .so is the one the committed sources build (sha256 stamp written by diffpir_b200.build; file times do not
    survive the copy to the GPU box, and two ranks importing at once must not both rebuild - build() takes a file lock)."""
    from diffpir_b200 import _lib, build
    _lib.load()
    assert os.path.exists(build.LIB) and os.path.exists(build.STAMP)
    assert not build.needs_build()
    assert open(build.STAMP).read().strip() == build._source_hash()


def test_weight_comm_binding_and_argument_checks():
    """The in-library weight broadcast (dpir_comm_*, dpir_bcast_weights): NCCL binds at run time to the copy the process already
    carries (PyTorch's here), the id is NCCL's 128 bytes, and bad arguments fail loudly without touching a GPU."""
    import ctypes
    import torch  # noqa: F401  (loads PyTorch's libnccl.so.2 first: the library binds to the loaded copy)
    from diffpir_b200 import _lib
    lib = _lib.load()
    v = ctypes.c_int32()
    assert lib.dpir_nccl_version(ctypes.byref(v)) == 0 and v.value >= 21800
    a, b = ctypes.create_string_buffer(128), ctypes.create_string_buffer(128)
    assert lib.dpir_comm_unique_id(a, 128) == 0 and lib.dpir_comm_unique_id(b, 128) == 0
    assert a.raw != b.raw and any(a.raw)
    assert lib.dpir_comm_unique_id(a, 64) == -1                      # DPIR_E_ARG: buffer too small
    h = ctypes.c_void_p()
    assert lib.dpir_comm_create(a, 2, 2, ctypes.byref(h)) == -1      # rank outside the world
    assert lib.dpir_comm_create(None, 1, 0, ctypes.byref(h)) == -1
    assert lib.dpir_bcast_weights(None, None, 0, None) == -1
    assert lib.dpir_comm_adopt(None, 1, 0, ctypes.byref(h)) == -1    # no communicator to adopt
    assert lib.dpir_comm_adopt(ctypes.c_void_p(1), 2, 5, ctypes.byref(h)) == -1
    assert lib.dpir_comm_destroy(None) == 0
