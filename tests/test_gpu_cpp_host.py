"""GPU test: builds tests/cpp/test_planners.cpp (the C++ host mirror of the reference's plugin interface,
include/meshnav_b200/planners.hpp) with g++ and runs it against libmeshnav_b200.so and the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror(tmp_path, oracle_mod):
    exe = str(tmp_path / "test_planners")
    pkg = os.path.join(ROOT, "mesh_navigation_b200")
    orc = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_planners.cpp"),
                           f"-L{pkg}", "-lmeshnav_b200", f"-L{orc}", "-loracle", f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{orc}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp host mirror ok" in out.stdout
