"""Host code under sanitizers (the reference has none): the CPU oracle with
AddressSanitizer + UndefinedBehaviorSanitizer, the staging copy pool with ThreadSanitizer.  CUDA
code is covered on the GPU box by compute-sanitizer (profiles/r2_sanitizer.txt)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")


def _run(cmd, exe, env=None):
    subprocess.check_call(cmd)
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe], capture_output=True, text=True, env=e, timeout=300)


@pytest.mark.skipif(GCC is None, reason="no gcc")
def test_oracle_under_asan_ubsan(tmp_path):
    exe = str(tmp_path / "oracle_san")
    rc = _run([GCC, "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fopenmp",
               "-std=gnu11", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle"),
               os.path.join(ROOT, "oracle", "fosphor_oracle.c"), os.path.join(ROOT, "tests", "c", "oracle_san_main.c"),
               "-o", exe, "-lm"], exe, env={"ASAN_OPTIONS": "detect_leaks=1", "OMP_NUM_THREADS": "4"})
    assert rc.returncode == 0, rc.stdout + rc.stderr
    assert "histogram mass" in rc.stdout and "runtime error" not in rc.stderr and "ERROR: AddressSanitizer" not in rc.stderr


@pytest.mark.skipif(GXX is None, reason="no g++")
def test_copy_pool_under_tsan(tmp_path):
    exe = str(tmp_path / "pool_san")
    rc = _run([GXX, "-O1", "-g", "-fsanitize=thread", "-std=c++17",
               os.path.join(ROOT, "gr-fosphor_b200", "host", "copy_pool.cc"),
               os.path.join(ROOT, "tests", "c", "copy_pool_san_main.cc"), "-o", exe, "-lpthread"], exe)
    assert rc.returncode == 0, rc.stdout + rc.stderr
    assert "rc 0" in rc.stdout and "WARNING: ThreadSanitizer" not in rc.stderr
