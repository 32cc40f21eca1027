"""The C++ batch-draining actor (gcra_actor_*, csrc/gcra_actor.inc): compiles and links on a CPU box; on a GPU box
the example runs the reference's actor tests (actor_tests.rs:8-70) through the C ABI and a short multi-threaded load."""
import json
import os
import subprocess

import pytest

from throttlecrab_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "actor_bench")


def _build():
    _native.build()
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "actor_bench.cpp"),
                           "-L" + os.path.join(ROOT, "throttlecrab_b200"), "-lgcra_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "throttlecrab_b200"), "-o", EXE])


def test_actor_bench_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("k1_path", ["auto"], indirect=True)
def test_actor_reference_tests_and_load(k1_path):
    _build()
    out = subprocess.run([EXE, "32", "2000", "500"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "actor ok" in out.stdout
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])["actor_bench"]
    assert line["requests"] == 64000 and line["mean_batch"] > 1.0      # callers really share batches
