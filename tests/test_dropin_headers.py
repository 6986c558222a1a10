"""CPU-side checks of the drop-in boundary: the C++ headers under include/quadblas/ compile (own test
program and, when /root/reference is present, the reference's unmodified test program), keep the
reference's include paths, and the caller-side SLEEF compatibility ops are correctly rounded."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
INC = os.path.join(ROOT, "include")


def test_same_relative_include_paths_as_reference():
    mine = {os.path.relpath(os.path.join(d, f), os.path.join(INC, "quadblas")) for d, _, fs in os.walk(os.path.join(INC, "quadblas")) for f in fs}
    need = ["quadblas.hpp", "core/platform.hpp", "core/constants.hpp", "core/types.hpp", "memory/allocation.hpp", "simd/quad_vector.hpp",
            "threading/openmp_utils.hpp", "detail/blocking.hpp", "algorithms/level1.hpp", "algorithms/level2.hpp", "algorithms/level3.hpp",
            "interface/c_interface.hpp", "interface/cpp_classes.hpp"]
    assert all(n in mine for n in need)
    if os.path.isdir(os.path.join(REF, "include", "quadblas")):
        ref = {os.path.relpath(os.path.join(d, f), os.path.join(REF, "include", "quadblas")) for d, _, fs in os.walk(os.path.join(REF, "include", "quadblas")) for f in fs}
        assert ref <= mine, ref - mine


def test_dropin_test_program_compiles_and_links(tmp_path):
    exe = tmp_path / "dropin_test"
    subprocess.run(["/usr/bin/g++", "-std=gnu++17", "-O1", "-Wall", "-Werror", "-I" + INC, os.path.join(ROOT, "tests", "host", "dropin_test.cpp"), "-o", str(exe),
                    "-L" + os.path.join(ROOT, "qblas_b200"), "-lqblas_b200"], check=True)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "test_quadblas.cpp")), reason="reference sources not present")
def test_reference_test_program_compiles_unmodified(tmp_path):
    exe = tmp_path / "quadblas_test_b200"
    with open(os.path.join(REF, "test_quadblas.cpp"), "rb") as src:
        subprocess.run(["/usr/bin/g++", "-std=gnu++17", "-O0", "-x", "c++", "-", "-I" + ROOT, "-I" + INC, "-o", str(exe),
                        "-L" + os.path.join(ROOT, "qblas_b200"), "-lqblas_b200"], stdin=src, check=True, cwd=ROOT)


@pytest.mark.parametrize("prog", ["tests/debug_test.cpp", "tests/test_sleef_simd.cpp"])
@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "tests", "debug_test.cpp")), reason="reference sources not present")
def test_reference_debug_programs_compile_unmodified(tmp_path, prog):
    """QuadVector + QuadBLAS::dot_kernel_vectorized (level1.hpp:14-35) are part of the surface the reference's debug programs use."""
    exe = tmp_path / "dbg"
    with open(os.path.join(REF, prog), "rb") as src:
        subprocess.run(["/usr/bin/g++", "-std=gnu++17", "-O0", "-x", "c++", "-", "-I" + ROOT, "-I" + INC, "-I" + os.path.join(INC, "quadblas"), "-o", str(exe),
                        "-L" + os.path.join(ROOT, "qblas_b200"), "-lqblas_b200"], stdin=src, check=True, cwd=ROOT)


def test_compat_scalar_ops_are_correctly_rounded(tmp_path):
    """Sleef_*q1_u05 of include/quadblas/b200/sleefquad_compat.h (caller-side only) vs libquadmath, bitwise."""
    src = tmp_path / "c.cpp"
    src.write_text(r'''
#include <quadblas/quadblas.hpp>
#include <quadmath.h>
#include <cstdio>
#include <cstring>
#include <random>
int main() {
  std::mt19937_64 g(7); long bad = 0;
  auto rq = [&]() { unsigned __int128 b = ((unsigned __int128)g() << 64) | g(); unsigned long long hi = (unsigned long long)(b >> 64);
    hi = (hi & 0x8000ffffffffffffULL) | ((unsigned long long)(16383 - 40 + (g() % 81)) << 48); b = ((unsigned __int128)hi << 64) | (unsigned long long)b;
    __float128 q; memcpy(&q, &b, 16); return q; };
  for (int i = 0; i < 200000; ++i) {
    __float128 a = rq(), b = rq(), c = rq();
    __float128 r1 = Sleef_fmaq1_u05(a, b, c), r2 = fmaq(a, b, c); bad += memcmp(&r1, &r2, 16) != 0;
    r1 = Sleef_addq1_u05(a, b); r2 = a + b; bad += memcmp(&r1, &r2, 16) != 0;
    r1 = Sleef_mulq1_u05(a, b); r2 = a * b; bad += memcmp(&r1, &r2, 16) != 0;
    r1 = Sleef_sqrtq1_u05(Sleef_fabsq1(a)); r2 = sqrtq(fabsq(a)); bad += memcmp(&r1, &r2, 16) != 0;
    double d = (double)a; bad += Sleef_cast_to_doubleq1(a) != d; r1 = Sleef_cast_from_doubleq1(d); r2 = d; bad += memcmp(&r1, &r2, 16) != 0;
  }
  printf("bad=%ld\n", bad); return bad != 0; }
''')
    exe = tmp_path / "c"
    subprocess.run(["/usr/bin/g++", "-std=gnu++17", "-O2", "-I" + INC, str(src), "-o", str(exe), "-L" + os.path.join(ROOT, "qblas_b200"), "-lqblas_b200", "-lquadmath"], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "qblas_b200"))
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout
