"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads here (nvcc
cross-compiles sm_100a without a GPU) and exports every symbol include/rlpyt_b200.h declares;
the ctypes signature table covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rlpyt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from rlpyt_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from rlpyt_b200.csrc.build import build
        build()
    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rlpyt_b200.h but not exported"


def test_signature_table_matches_header(lib):
    from rlpyt_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_abi_version_and_error_string(lib):
    assert lib.rl_b200_abi_version() >= 1
    assert isinstance(lib.rl_b200_last_error(), bytes)


def test_bad_arguments_are_rejected_without_gpu(lib):
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    rc = lib.rl_gae_f32(None, None, None, None, None, None, 4, 4, 0.99, 0.97, 0, None)
    assert rc == -1 and b"null" in lib.rl_b200_last_error()
    assert lib.rl_adv_normalize_scratch_bytes(32768) % 8 == 0


def test_no_cpu_fallback():
    """The product path refuses CPU-only operation loudly instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from rlpyt_b200 import _lib
    from rlpyt_b200.algos import utils as U
    with pytest.raises(_lib.B200LibraryError):
        U.discount_return(np.zeros((2, 2), np.float32), np.zeros((2, 2), bool), np.zeros((1, 2), np.float32), 0.9)


def test_product_never_imports_oracle():
    """Nothing under rlpyt_b200/ may import, call or execute anything under oracle/."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "rlpyt_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src and f.endswith(".py") and "import" in src and re.search(r"importlib.*oracle", src):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
