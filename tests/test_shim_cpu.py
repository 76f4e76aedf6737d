"""CPU: the sigutils-named shim (include/sigutils/*.h, libsigutils.so).  The per-sample entry points are per-sample by
ABI (`destination[p] = su_costas_feed(&costas, origin[p])`, Tasks/CostasRecoveryTask.cpp:58-61) and run the kernels'
own step functions on the caller's thread (sdb_chain_steps.h): here the reference-shaped loops of
tests/shim/reference_tu.cpp are checked bit for bit against the oracle.  Their bulk (device) twins and the analyzer
are in tests/test_gpu_shim.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import shim_build as SB
from sigdigger_b200 import synth


def _sig(n, seed=5):
    x, _ = synth.multi_carrier(n, 1.0, [("qpsk", 0.01, 0.1, -6.0, {})], noise_db=-40.0, seed=seed)
    return np.ascontiguousarray(x, np.complex64)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_costas_task_loop(oracle, kind):
    tu = SB.reference_tu()
    n = 2 * 4096 + 777
    x = _sig(n)
    y = np.zeros(n, np.complex64)
    assert tu.tu_costas_task(x.ctypes.data, y.ctypes.data, C.c_size_t(n), C.c_float(10.0), C.c_float(2e-3), kind, 0) == 0
    ref = SB.oracle_costas(oracle, x, kind, 0.1, 2e-3)
    assert np.array_equal(y.view(np.uint32), ref.view(np.uint32))


def test_pll_agc_xlate_gardner_task_loops(oracle):
    tu = SB.reference_tu()
    n = 2 * 4096 + 100
    x = _sig(n, seed=6)
    y = np.zeros(n, np.complex64)
    assert tu.tu_pll_task(x.ctypes.data, y.ctypes.data, C.c_size_t(n), C.c_float(5e-3), 0) == 0
    assert np.array_equal(y.view(np.uint32), SB.oracle_pll(oracle, x, 5e-3).view(np.uint32))
    assert tu.tu_agc_task(x.ctypes.data, y.ctypes.data, C.c_size_t(n), C.c_float(20.0)) == 0
    assert np.array_equal(y.view(np.uint32), SB.oracle_agc(oracle, x, 20.0).view(np.uint32))
    assert tu.tu_xlate_task(x.ctypes.data, y.ctypes.data, C.c_size_t(n), C.c_float(0.0123), C.c_float(0.5), 0) == 0
    assert np.array_equal(y.view(np.uint32), SB.oracle_xlate(oracle, x, 0.0123, 0.5).view(np.uint32))
    out = np.zeros(n, np.complex64)
    got = tu.tu_gardner_task(x.ctypes.data, C.c_size_t(n), C.c_float(0.1), C.c_float(0.1), out.ctypes.data, C.c_size_t(n))
    ref = SB.oracle_gardner_frequency(oracle, x, 0.1, 0.1)
    assert got == len(ref) and np.array_equal(out[:got].view(np.uint32), ref.view(np.uint32))


def test_shim_libraries_export_every_declared_symbol():
    """libsigutils.so / libsuscan.so export every function include/sigutils/*.h and include/analyzer/*.h declare."""
    import re
    root = SB.ROOT
    for lib, dirs in (("libsigutils.so", ["sigutils"]), ("libsuscan.so", ["analyzer"])):
        L = C.CDLL(os.path.join(root, "sigdigger_b200", lib))
        names = set()
        for d in dirs:
            for dp, _, files in os.walk(os.path.join(root, "include", d)):
                for f in files:
                    txt = open(os.path.join(dp, f)).read()
                    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
                    txt = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", txt)
                    names |= set(re.findall(r"\b((?:su|suscan|sigutils)_[a-z0-9_]+)\s*\(", txt))
        names -= {"su_specttuner_on_data_fn", "suscan_source_read_fn", "suscan_analyzer_baseband_filter_func_t"}
        missing = [n for n in sorted(names) if not hasattr(L, n)]
        assert not missing, (lib, missing)


def test_config_bag_round_trip():
    L = C.CDLL(os.path.join(SB.ROOT, "sigdigger_b200", "libsuscan.so"))
    L.suscan_inspector_config_new.restype = C.c_void_p
    L.suscan_config_get_value.restype = C.c_void_p
    L.suscan_config_dup.restype = C.c_void_p
    cfg = L.suscan_inspector_config_new(b"psk", C.c_float(3.125e6))
    assert cfg
    assert L.suscan_config_set_float(C.c_void_p(cfg), b"afc.loop-bw", C.c_float(1234.5)) == 1
    assert L.suscan_config_set_integer(C.c_void_p(cfg), b"afc.costas-order", C.c_uint64(3)) == 1
    assert L.suscan_config_set_float(C.c_void_p(cfg), b"afc.costas-order", C.c_float(1.0)) == 0     # wrong type
    assert L.suscan_config_set_float(C.c_void_p(cfg), b"no.such-key", C.c_float(1.0)) == 0
    dup = L.suscan_config_dup(C.c_void_p(cfg))
    assert L.suscan_config_get_value(C.c_void_p(dup), b"clock.running")
    assert not L.suscan_config_get_value(C.c_void_p(dup), b"audio.volume")       # not a psk key
    L.suscan_config_destroy(C.c_void_p(cfg)); L.suscan_config_destroy(C.c_void_p(dup))


def test_umbrella_header_compiles_as_c_and_cxx(tmp_path):
    """<suscan.h> (Suscan/Library.cpp:24, Default/GenericInspector/InspectorUI.cpp:43) brings in the whole suscan-named
    surface, as C and as C++"""
    src = tmp_path / "umbrella.c"
    src.write_text('#include <suscan.h>\n#include <sigutils/tvproc.h>\n#include <sigutils/softtune.h>\n'
                   'int main(void) { struct sigutils_channel c = sigutils_channel_INITIALIZER; (void) c; '
                   'return suscan_sigutils_init(0) ? 0 : 1; }\n')
    inc = os.path.join(SB.ROOT, "include")
    for cmd in (["gcc", "-std=gnu11", "-Wall", "-Werror"], ["g++", "-std=c++17", "-Wall", "-x", "c++"]):
        r = subprocess.run(cmd + ["-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
