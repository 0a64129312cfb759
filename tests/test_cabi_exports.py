"""CPU-only: the C-ABI library builds/loads and exports every symbol include/dca_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dca_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = declared_symbols()
    for must in ("dca_volume_gwc_concat", "dca_conv3d_direct", "dca_conv3d_tc", "dca_class_stats",
                 "dca_disp_attention", "dca_upsample_fuse", "dca_softmax_regress", "dca_convex_upsample",
                 "dca_pack_weights", "dca_version"):
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol():
    import dcanet_b200
    from dcanet_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/dca_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert _lib.load().dca_version() >= 100


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    import torch
    import dcanet_b200 as d
    with pytest.raises(d._lib.DcaError):
        d.build_gwc_volume(torch.zeros(1, 8, 2, 4), torch.zeros(1, 8, 2, 4), 2, 4)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cost-volume-aggregation-in-stereo-matching-revisited_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read(), f


def test_state_dict_layout_matches_reference():
    import dcanet_b200 as d
    ref = dict(l.strip().split(" ", 1) for l in open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.txt")))
    sd = d.GwcNet(192).state_dict()
    assert len(sd) == 726 and set(sd) == set(ref)
    for k, v in sd.items():
        assert str(tuple(v.shape)) == ref[k], k
    # DataParallel-prefixed checkpoint wrapper loads (main_dca.py:277-281)
    m = d.GwcNet(192)
    m.load_state_dict({"epoch": 1, "state_dict": {"module." + k: v for k, v in sd.items()}})


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/dca_b200.h compiles as C99 (no C++ / CUDA / torch types in any signature), and a C program that references an
    entry point links against the shared library with gcc alone."""
    import subprocess
    from dcanet_b200 import _lib
    src = tmp_path / "probe.c"
    src.write_text('#include "dca_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %d\\n", dca_version(), dca_plane_format()); '
                   'return dca_build_gwc_volume_f32(0, 0, 0, 1, 8, 4, 2, 2, 4, 0) == DCA_ERR_ARG ? 0 : 1; }\n')
    exe = tmp_path / "probe"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:libdca_b200.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    # version / plane format need no GPU; the NULL-pointer call must come back as DCA_ERR_ARG without touching the device
    assert out.returncode == 0, out
    assert out.stdout.split()[0] == str(_lib.load().dca_version())
