"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/lsdgpu.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "lsdgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lsdgpu_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_loads():
    from lsd_slam_b200 import build
    path = build.build()
    assert os.path.exists(path)
    ctypes.CDLL(path)


def test_every_declared_symbol_is_exported():
    from lsd_slam_b200 import abi, build
    build.build()
    L = ctypes.CDLL(abi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/lsdgpu.h but not exported"
    bound = {s[0] for s in abi.SYMBOLS}
    assert bound == set(declared), (bound ^ set(declared))


def test_abi_version_and_struct_sizes():
    from lsd_slam_b200 import abi
    L = abi.load()
    assert L.lsdgpu_abi_version() == 2
    assert ctypes.sizeof(abi.Hyp) == 32 and abi.HYP_DTYPE.itemsize == 32     # DepthMapPixelHypothesis.h:37-61
    s = abi.default_track_settings()
    assert list(s.maxItsPerLvl) == [5, 20, 50, 100, 0]                       # settings.h:368 + SlamSystem.cpp:80-81
    assert abs(s.convergenceEps[1] - 0.999) < 1e-6 and s.huber_d == 3.0


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (the CPU restatement is the checker only)."""
    pkg = os.path.join(ROOT, "lsd_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "lsd_oracle" not in txt and "liblsd_oracle" not in txt, f


def test_docs_only_name_declared_entry_points():
    """every lsdgpu_* identifier used in INTEGRATION.md / README.md / DESIGN.md is declared in include/lsdgpu.h, and every
    reference citation of the header has the form path:line"""
    import re
    hdr = open(os.path.join(ROOT, "include", "lsdgpu.h")).read()
    declared = set(re.findall(r"\b(lsdgpu_[a-z0-9_]+)\b", hdr))
    for doc in ("INTEGRATION.md", "README.md", "DESIGN.md"):
        used = set(re.findall(r"\b(lsdgpu_[a-z0-9_]+)\b", open(os.path.join(ROOT, doc)).read()))
        assert not (used - declared), (doc, sorted(used - declared))
    # each extern "C" function of the header carries a citation (file.cpp:line or file.h:line) in the comment above it
    funcs = re.findall(r"/\*(?:(?!\*/).)*\*/\s*(?:int|void|const char\*|long long)\s+(lsdgpu_[a-z0-9_]+)\s*\(", hdr, flags=re.S)
    assert len(funcs) >= 25
