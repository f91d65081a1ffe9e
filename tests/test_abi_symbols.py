"""CPU: the C-ABI library builds, loads, and exports every symbol include/wast3d_b200.h declares;
the product path refuses to run without a GPU (no CPU / oracle fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "wast3d_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wast3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(built):
    lib = ctypes.CDLL(str(built["cuda"]))
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_library_exports_staged_symbols(built):
    """include/wast3d_b200_staged.h (entry points staged for the next round; not part of the drop-in ABI)."""
    text = (ROOT / "include" / "wast3d_b200_staged.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    syms = sorted(set(re.findall(r"\b(wast3d_staged_[a-z0-9_]+)\s*\(", text)))
    assert len(syms) == 3
    lib = ctypes.CDLL(str(built["cuda"]))
    for s in syms:
        assert hasattr(lib, s), s


def test_python_binding_covers_header(built):
    from wast3d_b200 import _lib
    _lib.load()
    assert _lib.MISSING == []
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.load().wast3d_abi_version() == 7
    assert _lib.load().wast3d_strerror(0) == b"ok"
    assert b"no CPU fallback" in _lib.load().wast3d_strerror(4)


def test_no_oracle_import_in_product():
    for py in (ROOT / "wast3d_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_path_fails_loudly_without_gpu(built):
    from wast3d_b200.diff_gaussian_rasterization import _C
    from wast3d_b200.simple_knn._C import distCUDA2
    from wast3d_b200 import matching
    e = torch.empty(0)
    with pytest.raises(RuntimeError):
        _C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 3), e, torch.zeros(4, 1), torch.ones(4, 3),
                               torch.ones(4, 4), 1.0, e, torch.eye(4), torch.eye(4), 1.0, 1.0, 16, 16,
                               torch.zeros(4, 1, 3), 0, torch.zeros(3), False, False, e)
    with pytest.raises(RuntimeError):
        from wast3d_b200.model_render import rasterize_model
        from wast3d_b200.diff_gaussian_rasterization import GaussianRasterizationSettings
        rs = GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                           torch.zeros(3), False, False)
        rasterize_model(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 1, 3), torch.zeros(4, 0, 3),
                        torch.zeros(4, 1), torch.zeros(4, 3), torch.ones(4, 4), rs)
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(8, 3))
    with pytest.raises(RuntimeError):
        matching.nn_match(torch.zeros(4, 3), torch.zeros(4, 3))
    from wast3d_b200 import _lib
    assert _lib.load().wast3d_device_check(0) == 4  # WAST3D_ERR_NO_DEVICE
