"""CPU checks of the C-ABI boundary: the library builds, loads without a GPU, and exports
every symbol include/tensoflow_b200.h declares.  No compute calls."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from tensoflow_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from tensoflow_b200 import _lib
    header = (ROOT / "include" / "tensoflow_b200.h").read_text()
    declared = set(re.findall(r"TF_API\s+[\w\s\*]+?\b(tf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())


def test_abi_version_and_error_string(lib):
    assert lib.tf_abi_version() == 1
    assert isinstance(lib.tf_last_error(), bytes)


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    from tensoflow_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))


def test_product_does_not_import_oracle():
    for p in (ROOT / "tensoflow_b200").glob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
