"""The reference-side binding documented in INTEGRATION.md (the pybind/torch-extension bodies a maintainer of the
reference would write against include/wast3d_b200.h) must at least compile: the C++ blocks are extracted from the
document and type-checked with g++ against the installed torch headers and our header (no GPU, no link)."""
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_integration_md_cpp_stub_compiles(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    from torch.utils import cpp_extension
    text = (ROOT / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```cpp\n(.*?)```", text, flags=re.S)
    assert len(blocks) >= 2, "INTEGRATION.md lost its C++ binding stubs"
    src = tmp_path / "stub.cpp"
    src.write_text("\n".join(blocks))
    inc = [f"-I{p}" for p in cpp_extension.include_paths(device_type="cuda")] + [f"-I{ROOT / 'include'}"]
    import sysconfig
    inc.append(f"-I{sysconfig.get_paths()['include']}")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-w", "-DTORCH_EXTENSION_NAME=stub", *inc, str(src)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]
