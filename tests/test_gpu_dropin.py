"""The drop-in C++ surface on a GPU: include/quadblas/quadblas.hpp (Vector/Matrix/free functions) and the
reference-named C ABI, linked against libqblas_b200.so.

  * tests/host/dropin_test.cpp — this repo's own program (known answers of the reference's tests).
  * oracle/_ref/quadblas_test_b200 — the reference's OWN test program (/root/reference/test_quadblas.cpp,
    unmodified, 20 cases) compiled in the dev container against these headers (oracle/Makefile); it
    travels to the GPU box as a prebuilt binary because /root/reference does not exist there."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_dropin():
    exe = os.path.join(ROOT, "tests", "host", "_build", "dropin_test")
    if not os.path.exists(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-std=gnu++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "host", "dropin_test.cpp"), "-o", exe, "-L" + os.path.join(ROOT, "qblas_b200"),
                        "-lqblas_b200", "-Wl,-rpath,$ORIGIN/../../../qblas_b200"], check=True)
    return exe


def test_dropin_cpp_surface(qb):
    r = subprocess.run([_build_dropin()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all passed" in r.stdout


def test_reference_own_test_program_passes_unmodified(qb):
    exe = os.path.join(ROOT, "oracle", "_ref", "quadblas_test_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/quadblas_test_b200 not built (reference sources absent when build() ran)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    tail = r.stdout[-2500:]
    assert r.returncode == 0, tail + r.stderr[-1000:]
    assert "FAIL" not in r.stdout.upper().replace("FAILED: 0", ""), tail


@pytest.mark.parametrize("env", [{"QUADBLAS_MODE": "fast"}, {"QUADBLAS_MODE": "fast", "OMP_NUM_THREADS": "3"}, {"OMP_NUM_THREADS": "5", "QUADBLAS_KC": "256"}])
def test_reference_own_test_program_with_environment_selected_modes(qb, env):
    """The same unmodified binary with the numerical mode chosen through the environment (no source change): fast mode (tensor path for
    the large cases, window accumulator for dot / gemv), other thread counts (dot chunking), the Apple-Silicon k-panel: 20/20 each."""
    exe = os.path.join(ROOT, "oracle", "_ref", "quadblas_test_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/quadblas_test_b200 not built (reference sources absent when build() ran)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    tail = r.stdout[-2500:]
    assert r.returncode == 0, tail + r.stderr[-1000:]
    assert "FAIL" not in r.stdout.upper().replace("FAILED: 0", ""), tail


@pytest.mark.parametrize("prog,expect", [("debug_test_b200", ["a*b + 1 = 7", "15"]), ("test_sleef_simd_b200", ["14"])])
def test_reference_debug_programs_run_unmodified(qb, prog, expect):
    """/root/reference/tests/debug_test.cpp and tests/test_sleef_simd.cpp (QuadVector, QuadBLAS::dot_kernel_vectorized of
    level1.hpp:14-35, dot 1..5 = 15, the 2 x 2 gemv), compiled unmodified against the drop-in headers (oracle/Makefile)."""
    exe = os.path.join(ROOT, "oracle", "_ref", prog)
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref/{prog} not built (reference sources absent when build() ran)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1000:]
    for e in expect:
        assert e in r.stdout, (e, r.stdout[-2500:])
    assert "FAIL" not in r.stdout and "ERROR" not in r.stdout.upper().replace("ERROR = 0", ""), r.stdout[-2500:]
