"""Compiles the product's arithmetic core (qblas_b200/csrc/q128.cuh + q128_chain.cuh, host/device
dual source) with g++ and checks it bitwise against libquadmath on millions of vectors in 12
regimes (tests/host/q128_host_test.cpp).  This is how the soft-float is validated without a GPU;
the -m gpu tests re-check the nvcc build of the same source."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_q128_core_matches_libquadmath(tmp_path):
    exe = tmp_path / "q128_host_test"
    subprocess.run(["/usr/bin/g++", "-O2", "-fopenmp", "-std=gnu++17", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "q128_host_test.cpp"), "-lquadmath"], check=True)
    r = subprocess.run([str(exe), "400000", "20261017"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "failing regimes: 0" in r.stdout
    # both the inlined fast path and the generic slow path must have been exercised
    fast = [float(l.split("fast=")[1].split()[0]) for l in r.stdout.splitlines() if "fast=" in l]
    assert max(fast) > 0.7 and min(fast) < 0.05
