"""Host logic of the product that needs no GPU: the BVH builder / child-group flattener (go-pbrt_b200/csrc/gp_bvh.h).
tests/cpp/bvh_check.cpp walks the flattened layout the way the traversal kernels read it and checks its invariants
(every primitive in exactly one leaf, leaf sizes, float32 boxes contain the float64 bounds below them, 128-byte group
alignment, determinism, the threaded path above 200 k primitives)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bvh_builder_invariants():
    src = os.path.join(ROOT, "tests", "cpp", "bvh_check.cpp")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "bvh_check")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bvh_check: ok" in out.stdout, out.stdout[-2000:]
