"""The C-ABI library builds, loads and exports exactly what include/valunc.h
declares; the ctypes mirror agrees with the header.  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from diffuncertainty_b200 import _lib, calibration

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "valunc.h")).read()


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = re.findall(r"^VU_API\s+[\w\s\*]+?\b(vu_[a-z0-9_]+)\s*\(", HEADER, flags=re.M)
    assert len(declared) >= 14
    assert set(declared) == set(_lib.EXPORTS), set(declared) ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.vu_abi_version() == _lib.VU_ABI_VERSION
    assert b"sm_100a" in lib.vu_build_info()


def test_header_constants_match_binding():
    defs = {m[0]: int(m[1], 0) for m in re.findall(r"^#define\s+(VU_[A-Z0-9_]+)\s+(-?(?:0x)?[0-9a-fA-F]+)u?\b", HEADER, flags=re.M)}
    for key, val in _lib.F64.items():
        assert defs[f"VU_F64_{key}"] == val
    for key, val in _lib.I64.items():
        assert defs[f"VU_I64_{key}"] == val
    assert defs["VU_N_BINS"] == _lib.N_BINS and defs["VU_N_EDGES"] == _lib.N_EDGES
    assert defs["VU_STAT_CALIB"] == _lib.STAT_CALIB and defs["VU_STAT_NCC"] == _lib.STAT_NCC
    lib = _lib.load()
    assert lib.vu_struct_size(0) == C.sizeof(_lib.FusedArgs)
    assert lib.vu_struct_size(1) == C.sizeof(_lib.MapStatsArgs)
    assert lib.vu_struct_size(2) == C.sizeof(_lib.Calib)
    assert lib.vu_struct_size(4) == C.sizeof(_lib.MemberScoresArgs)
    assert lib.vu_ged_cols(32, 4) == 2 * 32 * 4 + 4 + 32 * 32 + 32 + 2 * 16 + 3


def test_argument_errors_without_a_device():
    lib = _lib.load()
    assert lib.vu_fused_pass(None, None) == _lib.VU_ERR_BAD_ARG
    a = _lib.FusedArgs()
    a.struct_size = 4  # wrong
    assert lib.vu_fused_pass(C.byref(a), None) == _lib.VU_ERR_BAD_ARG
    assert b"struct_size" in lib.vu_last_error()
    a.struct_size = C.sizeof(_lib.FusedArgs)
    assert lib.vu_fused_pass(C.byref(a), None) == _lib.VU_ERR_BAD_ARG  # slab.data NULL
    with pytest.raises(ValueError):
        _lib.check(_lib.VU_ERR_BAD_ARG, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.VU_ERR_UNSUPPORTED, "x")
    assert lib.vu_patch_max(None, 1, 1, 8, 8, 1, 2, 2, 0, None, None, None) == _lib.VU_ERR_BAD_ARG


def test_product_has_no_cpu_fallback():
    import torch
    from diffuncertainty_b200 import uncertainty
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.ValuncError):
        uncertainty.calculate_uncertainty(torch.rand(3, 2, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffuncertainty_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


@pytest.mark.parametrize("a,b", [(3.5, -1.25), (6.0, -2.0), (-2.0, 0.5), (900.0, -300.0), (0.37, 0.0)])
def test_c_platt_inversion_agrees_with_python(a, b):
    """The C helper uses libm's expf, the Python one NumPy's exp; they may differ
    by an ulp at an edge but not more, and both must be monotone."""
    lib = _lib.load()
    c = _lib.Calib()
    assert lib.vu_platt_invert_edges_host(a, b, C.byref(c)) == 0
    py = calibration.platt_edges(a, b)
    assert c.mode == py.mode
    ce = np.array(list(c.edge_u), np.float32)
    assert np.array_equal(np.isnan(ce), np.isnan(py.edge_u))
    fin = ~np.isnan(ce)
    # both sets of thresholds sit at the crossing of their edge (u may differ by many
    # ulps where conf is flat in float32, conf itself by an ulp or two)
    edges = calibration.bin_edges()[1:20][fin]
    for thr in (ce[fin], py.edge_u[fin]):
        conf = calibration._platt_f32(thr.astype(np.float32), a, b).astype(np.float64)
        assert np.all(np.abs(conf - edges) <= 4e-7 * np.maximum(edges, 0.05) + (1e-5 if abs(a) > 100 else 0.0))


def test_header_is_plain_c():
    """include/valunc.h is the drop-in boundary: it must compile as C (no C++ / CUDA / torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = '#include "valunc.h"\nint main(void) { vu_fused_args a; vu_member_scores_args m; (void)a; (void)m; return vu_abi_version() == VU_ABI_VERSION ? 0 : 1; }\n'
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(root, "include"), "-x", "c", "-"],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_built_library_is_current_wherever_the_tree_is_copied():
    """The library is built in one place and travels with the tree (gpurun snapshot, torchrun ranks): the staleness check must not
    depend on the checkout's absolute path, or every process on the other side would rebuild -- and race -- on first use."""
    import shutil
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _lib.load()  # makes sure the in-tree library and its stamp exist
    with tempfile.TemporaryDirectory() as d:
        shutil.copytree(os.path.join(root, "diffuncertainty_b200"), os.path.join(d, "diffuncertainty_b200"),
                        ignore=shutil.ignore_patterns("__pycache__", "obj"))
        shutil.copytree(os.path.join(root, "include"), os.path.join(d, "include"))
        code = ("import sys; sys.path.insert(0, %r); from diffuncertainty_b200 import build as b; import os; "
                "print(int(b._current(os.path.join(b.LIBDIR, 'build.stamp'), b._digest())))" % d)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.strip() == "1", r.stderr[-1000:]
