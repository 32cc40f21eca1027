"""The C++ host mirror (include/throttlecrab_b200.hpp): compiles and links on a CPU box; on a GPU box the
example replays a few of the reference's known-answer tests through it."""
import os
import subprocess

import pytest

from throttlecrab_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "cpp_known_answers")


def _build():
    _native.build()
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cpp_known_answers.cpp"),
                           "-L" + os.path.join(ROOT, "throttlecrab_b200"), "-lgcra_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "throttlecrab_b200"), "-o", EXE])


def test_cpp_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_known_answers():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp_known_answers ok" in out.stdout
