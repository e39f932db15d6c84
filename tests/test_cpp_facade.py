"""The C++ host facade (include/akua_pbf.hpp) compiles with plain g++ against the C-ABI library and, on a GPU box, runs
the README scene built the way Application::prepareDamBreak builds it."""
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def _build(tmp_path, akua_lib):
    exe = tmp_path / "facade_smoke"
    libdir = REPO / "akuaengine_b200"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(REPO / "include"), str(REPO / "tests" / "cpp" / "facade_smoke.cpp"),
           "-L", str(libdir), "-l:libakua_pbf.so", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_cpp_facade_compiles_and_links(tmp_path, akua_lib):
    exe = _build(tmp_path, akua_lib)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0 and "link-ok" in res.stdout


@pytest.mark.gpu
def test_cpp_facade_runs_dam_break(tmp_path, akua_lib):
    exe = _build(tmp_path, akua_lib)
    res = subprocess.run([str(exe), "--run"], capture_output=True, text=True)
    assert res.returncode == 0 and "facade-ok" in res.stdout, res.stdout + res.stderr
