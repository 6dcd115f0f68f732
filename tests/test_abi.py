"""The C-ABI library builds, loads and exports exactly what include/msda_b200.h declares; argument
errors come back as status codes + messages.  No compute is attempted here (CPU-only box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msda_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:msda|mask|tc|aligned_bilinear|query_init)_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from mdqe_cvpr2023_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("msda_forward", "msda_backward", "mask_logits_forward", "mask_logits_backward", "msda_forward_host",
                 "msda_backward_host", "mask_logits_forward_host", "msda_last_error", "msda_abi_version"):
        assert must in names


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mdqe_cvpr2023_b200 import _lib
    names = declared_functions()
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/msda_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototype table out of sync with the header"


def test_abi_version_and_options(lib):
    from mdqe_cvpr2023_b200 import _lib
    assert lib.msda_abi_version() == _lib.ABI_VERSION == 6
    _lib.set_option("chunk_pairs", 48)
    assert _lib.get_option("chunk_pairs") == 48
    _lib.set_option("chunk_pairs", 0)
    with pytest.raises(RuntimeError, match="unknown key"):
        _lib.set_option("no_such_knob", 1)


def test_argument_errors_are_status_codes(lib):
    from mdqe_cvpr2023_b200 import _lib
    rc = lib.msda_forward(None, 0, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, None)
    assert rc == -1 and "NULL" in _lib.last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.msda_forward(None, 9, p, p, p, p, p, 1, 1, 1, 1, 1, 1, 1, p)
    assert rc == -1 and "dtype" in _lib.last_error()
    rc = lib.msda_forward(None, 0, p, p, p, p, p, 1, 1, 0, 1, 1, 1, 1, p)
    assert rc == -1 and "bad sizes" in _lib.last_error()
    rc = lib.mask_logits_forward(None, 2, 0, p, p, 1, 1, 1, 1, p)
    assert rc == -1
    assert lib.msda_backward_workspace_bytes(_lib.MSDA_BF16, 2, 10, 8, 32) == 2 * 10 * 8 * 32 * 4
    assert lib.msda_backward_workspace_bytes(_lib.MSDA_F32, 2, 10, 8, 32) == 0


def test_no_cpu_fallback_without_a_device(lib):
    """On a box without a GPU a compute call must fail loudly (MSDA_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from mdqe_cvpr2023_b200 import _lib
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    rc = lib.msda_forward(None, 0, p, p, p, p, p, 1, 1, 1, 32, 1, 1, 1, p)
    assert rc == -3, _lib.last_error()
    t = torch.zeros(1, 1, 1, 32)
    from mdqe_cvpr2023_b200 import ops
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ops.ms_deform_attn_forward(t, torch.ones(1, 2, dtype=torch.long), torch.zeros(1, dtype=torch.long),
                                   torch.zeros(1, 1, 1, 1, 1, 2), torch.ones(1, 1, 1, 1, 1), 64)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mdqe_cvpr2023_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("# oracle-free", ""), f"{f} mentions the oracle"
