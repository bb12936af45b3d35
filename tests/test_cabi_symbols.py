"""The C-ABI library loads without a GPU and exports every symbol include/vlfm_b200.h declares."""
import os
import re

from vlfm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "vlfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlfm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vlfm_b200.h but not exported"
    assert sorted(_lib.declared_symbols()) == names
    assert lib.vlfm_version() >= 100
    assert lib.vlfm_launch_count() == 0 or lib.vlfm_launch_count() > 0
