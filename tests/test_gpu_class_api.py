"""The reference-named C++ class API (csrc/host/carma_host.hpp + carma_steps.hpp: Parameter / Ensemble / StudentProposal /
AdaptiveMetro / ExchangeStep / Sampler / CARMA::StartingValue, Save, Value, GetLogDensity ...) exercised from C++ the
way the reference's own Catch tests do (cpp_tests/carma_unit_tests.cpp:783-911, 1068-1113): tests/cpp/test_class_api.cpp
is compiled against libcarma_b200.so and run on the GPU.  Also the NCCL summary gather of the C ABI with one rank."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "carma_pack_b200", "csrc", "host")


def build_cpp_test():
    exe = os.path.join(ROOT, "tests", "cpp", "test_class_api")
    srcs = [os.path.join(ROOT, "tests", "cpp", "test_class_api.cpp"), os.path.join(HOST, "carma_host.cpp"),
            os.path.join(HOST, "carma_steps.cpp")]
    deps = srcs + [os.path.join(HOST, "carma_host.hpp"), os.path.join(HOST, "carma_steps.hpp"),
                   os.path.join(ROOT, "carma_pack_b200", "libcarma_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + HOST] + srcs +
                              ["-L" + os.path.join(ROOT, "carma_pack_b200"), "-lcarma_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "carma_pack_b200"), "-o", exe])
    return exe


def test_cpp_class_api_builds_and_links_without_a_gpu():
    """CPU side: a C++ caller written against the reference's class names compiles and links."""
    assert os.path.exists(build_cpp_test())


@pytest.mark.gpu
def test_cpp_class_api_on_gpu():
    exe = build_cpp_test()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    sys.stdout.write(out.stdout[-3000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "all class-API checks passed" in out.stdout
    for name in ("CARMA/logpost_test", "CAR1/logpost_test", "hand-built PT sampler", "CholUpdateR1"):
        assert ("ok: " + name) in out.stdout


@pytest.mark.gpu
def test_gather_summaries_single_rank_nccl():
    """carma_gather_summaries through a communicator made by carma_comm_init_rank (one rank): identity gather."""
    import carma_pack_b200 as C
    lib = C._lib.lib
    ident = ctypes.create_string_buffer(128)
    C._lib.check(lib.carma_comm_unique_id(ident), "carma_comm_unique_id")
    comm = ctypes.c_void_p()
    C._lib.check(lib.carma_comm_init_rank(1, 0, ident, 0, ctypes.byref(comm)), "carma_comm_init_rank")
    local = np.arange(37, dtype=np.float64) * 0.5 - 3.0
    allv = np.empty(37)
    C._lib.check(lib.carma_gather_summaries(comm, local.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 37,
                                            allv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None), "carma_gather_summaries")
    assert np.array_equal(allv, local)
    C._lib.check(lib.carma_comm_destroy(comm), "carma_comm_destroy")
