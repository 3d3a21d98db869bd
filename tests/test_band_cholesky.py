"""Band Cholesky of the host minimizer (portable path, AVX2/FMA panel path, run-time dispatcher) and the AVX2 band
matrix-vector product against a dense reference (tests/cpp/band_chol_test.cpp), also with the fast path disabled (GLIO_NO_AVX2=1)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_cholesky_paths(tmp_path):
    exe = str(tmp_path / "band_chol_test")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "band_chol_test.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "glio_b200"), "-lglio_b200", "-Wl,-rpath," + os.path.join(ROOT, "glio_b200"),
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    for env in ({}, {"GLIO_NO_AVX2": "1"}):
        p = subprocess.run([exe], text=True, capture_output=True, env={**os.environ, **env})
        assert p.returncode == 0, p.stdout + p.stderr
        lines = p.stdout.splitlines()
        assert lines[-1] == "cases 17 bad 0" and all(" ok" in l for l in lines[:-1]), p.stdout
