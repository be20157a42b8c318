"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/ugsm.h declares,
its pure-host entry points agree with the oracle's tables, and it fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ugsm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ugsm_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(capi):
    syms = _header_symbols()
    assert len(syms) >= 20
    lib = capi.lib()
    for s in syms:
        assert hasattr(lib, s), "libugsm.so does not export %s" % s
    assert sorted(capi.SYMBOLS) == syms


def test_struct_layout_matches_header(capi):
    assert C.sizeof(capi.Config) == 8 * 4
    assert C.sizeof(capi.Dims) == 4 + 6 * 14 * 4 + 3 * 4 + 8
    assert C.sizeof(capi.Timing) == 5 * 4 + 4 + 3 * 14 * 4


def test_defaults_are_reference_constants(capi):
    c = capi.default_config()
    assert (c.levels, c.fovea_level, c.iters_cutoff, c.smooth_passes, c.smooth_passes_fine) == (14, 7, 22, 5, 10)
    assert capi.lib().ugsm_version().startswith(b"ugsm-b200")


@pytest.mark.parametrize("shape", [(3264, 4928), (1632, 2464), (480, 640)])
@pytest.mark.parametrize("mode", [0, 1])
def test_query_dims_matches_oracle(capi, oracle, shape, mode):
    rows, cols = shape
    d = capi.query_dims(capi.default_config(), rows, cols, mode)
    w, h = oracle.level_dims(cols, rows)
    ocfg = oracle.default_config()
    assert list(d.width) == w and list(d.height) == h
    assert list(d.iters) == [oracle.level_iters(ocfg, i) for i in range(14)]
    assert list(d.smooth) == [oracle.level_smooth(ocfg, i) for i in range(14)]
    assert (d.fov_width, d.fov_height, d.stack_rows) == (w[6], h[6], 7 * h[6])
    if mode == 0:
        assert list(d.match_width) == w
    else:
        assert list(d.match_width) == [w[6]] * 6 + w[6:]
    if shape == (3264, 4928):
        assert d.pixel_iterations == (131429636 if mode == 0 else 21414118)


def test_query_dims_rejects_bad_input(capi):
    cfg = capi.default_config()
    with pytest.raises(capi.UgsmError):
        capi.query_dims(cfg, 2, 2, 0)  # 14 levels do not fit
    with pytest.raises(capi.UgsmError):
        capi.query_dims(cfg, 0, 10, 0)
    with pytest.raises(capi.UgsmError):
        capi.query_dims(capi.default_config(levels=15), 100, 100, 0)
    with pytest.raises(capi.UgsmError):
        capi.query_dims(capi.default_config(fovea_level=1), 480, 640, 1)


def test_no_cpu_fallback(capi):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(capi.UgsmError) as e:
        capi.Handle()
    assert e.value.code == capi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product package may not reference oracle/ (the oracle is the checker, never the path)."""
    pkg = os.path.join(ROOT, "ug_stereomatcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ugsm_oracle" not in text and "ugsmo_" not in text, os.path.join(dirpath, f)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C-ABI: include/ugsm.h must compile as C99 (no C++ or CUDA types in the signatures),
    and a C program must link against the library and reach the host-only entry points."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use.c"
    src.write_text('#include <stdio.h>\n#include "ugsm.h"\n'
                   'int main(void) { ugsm_config c; ugsm_dims d; ugsm_config_default(&c);\n'
                   '  if (ugsm_query_dims(&c, 3264, 4928, 1, &d) != UGSM_OK) return 2;\n'
                   '  printf("%d %d %d %s %d %d\\n", c.levels, d.fov_width, d.fov_height, ugsm_version(),\n'
                   '         (int)sizeof(ugsm_feeder_pair), (int)sizeof(ugsm_feeder_info_t)); return 0; }\n')
    exe = str(tmp_path / "use")
    lib_dir = os.path.join(root, "ug_stereomatcher_b200")
    p = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                        "-o", exe, str(src), "-L" + lib_dir, "-lugsm", "-Wl,-rpath," + lib_dir],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    out = subprocess.run([exe], stdout=subprocess.PIPE).stdout.decode().split()
    assert out[:3] == ["14", "615", "407"]
    from ug_stereomatcher_b200 import capi
    assert [int(out[-2]), int(out[-1])] == [C.sizeof(capi.FeederPair), C.sizeof(capi.FeederInfo)]  # ctypes mirrors
