"""The reference's own parity harness ("Unit test/correctness_test.cpp":176-221), revived in C++ against the C ABI:
compiled with g++ on the GPU box, linked with libexahype_cuda.so, compared with the CPU oracle and with the reference's
own compiled kernel (oracle/_ref)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def test_correctness_test_cpp_against_the_drop_in(tmp_path, oracle):
    from exahype_b200 import runtime
    runtime.load()
    exe = tmp_path / "correctness_test_b200"
    subprocess.run([GXX, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
                    os.path.join(ROOT, "tests", "cpp", "correctness_test_b200.cpp"),
                    "-L", os.path.join(ROOT, "exahype_b200"), "-lexahype_cuda", "-ldl", "-o", str(exe)], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "exahype_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe), os.path.join(ROOT, "oracle", "libfv_oracle.so"),
                        os.path.join(ROOT, "oracle", "_ref", "libexahype_ref.so")], capture_output=True, text=True, env=env)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no differences! :)" in r.stdout
    assert "vs CPU oracle (all 360 values): 0 differences" in r.stdout
