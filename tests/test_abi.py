"""The C-ABI shared library loads, exports every symbol include/radarays_b200.h declares, mirrors the reference's
parameter schema, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from radarays_ros_b200 import RadarModel, RadarModelConfig, capi
from radarays_ros_b200.types import _CFG_FIELDS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "radarays_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libradarays_b200.so does not export %s" % n
    assert sorted(capi.SYMBOLS) == names
    assert L.rr_abi_version() == capi.ABI_VERSION == 3


def test_config_defaults_match_reference_cfg():
    """rr_config_defaults == cfg/RadarModel.cfg:11-85 == types.RadarModelConfig()."""
    L = capi.lib()
    c = RadarModelConfig()
    for n, _, _ in _CFG_FIELDS:
        setattr(c, n, 0)
    L.rr_config_defaults(C.byref(c))
    ref = RadarModelConfig()
    for n, _, d in _CFG_FIELDS:
        assert getattr(c, n) == d == getattr(ref, n), n
    m = RadarModel()
    L.rr_model_defaults(C.byref(m))
    assert (m.n_samples, m.n_reflections) == (200, 2)                  # ros_helper.h:21-28
    assert abs(m.beam_width - 8.0 * 3.141592653589793 / 180.0) < 1e-6
    assert C.sizeof(RadarModelConfig) == 256                            # same as sizeof(rr_config) in C (checked with gcc: 256)


def test_cfg_file_of_reference_has_same_parameters():
    path = "/root/reference/cfg/RadarModel.cfg"
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this machine")
    live = [ln for ln in open(path).read().splitlines() if not ln.lstrip().startswith("#")]      # 3 adds are commented out
    names = re.findall(r'gen\.add\("([a-z_0-9]+)"', "\n".join(live))
    assert names == [n for n, _, _ in _CFG_FIELDS]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = capi.lib()
    ctx = C.c_void_p()
    rc = L.rr_create(C.byref(ctx), 0)
    assert rc == -6 and not ctx.value                                  # RR_ERR_NO_DEVICE
    assert b"no CUDA device" in L.rr_last_error(None)
    from radarays_ros_b200.radar import RadarB200
    with pytest.raises(capi.RadaRaysError):
        RadarB200()


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py may load oracle/: no import, include, link or dlopen of it in the package
    (comments may mention the oracle; code may not reference it)."""
    pkg = os.path.join(ROOT, "radarays_ros_b200")
    pat = re.compile(r"from\s+oracle|import\s+oracle|liboracle|oracle/|oracle\.py|#include\s+[\"<][^\n]*oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)          # strip C comments
                code = "\n".join(ln for ln in code.splitlines() if not ln.lstrip().startswith(("#", "//", "*", '"""')) or "#include" in ln)
                m = pat.search(code)
                assert not m, "%s references the oracle: %r" % (f, m.group(0))


def test_header_is_valid_c99_and_cxx(tmp_path):
    """include/radarays_b200.h must compile as plain C (the cgo / JNI / ctypes-style binding target) and as C++."""
    import shutil
    import subprocess
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    src = tmp_path / "abi_check.c"
    src.write_text('#include <radarays_b200.h>\nint main(void) { rr_config c; rr_model m; rr_pose p; rr_radar_params g; rr_mesh s; rr_ipc_handle h;'
                   ' (void)c; (void)m; (void)p; (void)g; (void)s; (void)h; return sizeof(rr_stats) > 0 ? 0 : 1; }\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        if shutil.which(cc) is None:
            pytest.skip("no %s" % cc)
        args = [cc, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc]
        if cc == "g++":
            args += ["-x", "c++"]
        subprocess.run(args + [str(src)], check=True)
