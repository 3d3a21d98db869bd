"""Ceres' own DoglegStrategy and TrustRegionMinimizer (Powell, dogleg) known-answer tests (the reference vendors Ceres 2.0.0:
support_files/ceres-solver.tar.gz::internal/ceres/dogleg_strategy_test.cc) against the product's host minimizer
(glio_b200/csrc/solver.cpp, the a-7 row of DESIGN.md).  Fixtures and expectations are transcribed in tests/cpp/dogleg_test.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ceres_dogleg_strategy_known_answers(tmp_path):
    exe = str(tmp_path / "dogleg_test")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "dogleg_test.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "glio_b200"), "-lglio_b200", "-Wl,-rpath," + os.path.join(ROOT, "glio_b200"),
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    out = subprocess.check_output([exe], text=True).splitlines()
    names = [l.split()[0] for l in out]
    assert names[:7] == ["TrustRegionObeyedTraditional", "TrustRegionObeyedSubspace", "CorrectGaussNewtonStep", "CorrectStepLocalOptimumAlongGradient",
                         "CorrectStepGlobalOptimumAlongGradient", "ValleyTraditionalActive", "ValleyTraditionalInactive"]
    # trust_region_minimizer_test.cc: PowellsSingularFunctionUsingDogleg, the 13 column activations Ceres runs
    assert [n for n in names if n.startswith("Powell_")] == ["Powell_1110", "Powell_1011", "Powell_0111", "Powell_1100", "Powell_1010", "Powell_0110",
                                                              "Powell_1001", "Powell_0101", "Powell_0011", "Powell_1000", "Powell_0100", "Powell_0010", "Powell_0001"]
    # trust_region_minimizer_test.cc:257-280 PowellsSingularFunctionUsingLevenbergMarquardt (14 activations) and
    # levenberg_marquardt_strategy_test.cc (radius scaling, diagonal handed to the linear solver)
    assert len([n for n in names if n.startswith("PowellLM_")]) == 14
    assert "LM_AcceptRejectStepRadiusScaling" in names and "LM_CorrectDiagonalToLinearSolver" in names
    # polynomial_test.cc: the root finder of the subspace dogleg
    assert len([n for n in names if n.startswith("Poly_")]) == 11
    bad = [l for l in out if l.split()[1] != "ok"]
    assert not bad, bad
