"""CPU-only checks of the C ABI: the library loads, exports every symbol the
header declares, and the struct layouts match the reference's private.h."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from gr_fosphor_b200 import build
    path = build.build()
    return C.CDLL(path)


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fosphor_b200.h")).read()
    syms = sorted(set(re.findall(r"\b(fosphor_(?:cl|cu|fifo|window|host)_[a-z_]+)\s*\(", hdr)))
    assert len(syms) >= 25
    L = _lib()
    for s in syms:
        assert getattr(L, s) is not None, s
    # exactly the reference's seven boundary symbols (lib/fosphor/cl.h:22-32)
    assert [s for s in syms if s.startswith("fosphor_cl_")] == [
        "fosphor_cl_finish", "fosphor_cl_get_waterfall_position", "fosphor_cl_init",
        "fosphor_cl_load_fft_window", "fosphor_cl_process", "fosphor_cl_release",
        "fosphor_cl_set_histogram_range"]


def test_struct_fosphor_layout_matches_private_h():
    """lib/fosphor/private.h:30-55 on LP64: offsets the drop-in relies on."""
    from gr_fosphor_b200.dropin import StructFosphor
    S = StructFosphor
    assert S.cl.offset == 0 and S.gl.offset == 8 and S.flags.offset == 16
    assert S.fft_win.offset == 20 and S.fft_win.size == 4096
    assert S.img_waterfall.offset == 4120 and S.img_histogram.offset == 4128
    assert S.buf_spectrum.offset == 4136 and S.power.offset == 4144
    assert S.frequency.offset == 4160 and C.sizeof(S) == 4176
    # and the C header restating it compiles to the same numbers
    import subprocess, tempfile
    src = r'''
#include <stddef.h>
#include <stdio.h>
#include "fosphor_private_abi.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu\n", offsetof(struct fosphor, fft_win),
 offsetof(struct fosphor, img_waterfall), offsetof(struct fosphor, buf_spectrum),
 offsetof(struct fosphor, power), offsetof(struct fosphor, frequency), sizeof(struct fosphor));return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert [int(v) for v in out] == [20, 4120, 4136, 4144, 4160, 4176]


def test_reference_private_h_agrees_when_present():
    ref = "/root/reference/lib/fosphor/private.h"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present (GPU box)")
    import subprocess, tempfile
    src = r'''
#include <stddef.h>
#include <stdio.h>
#include "private.h"
int main(void){printf("%zu %zu %zu %zu\n", offsetof(struct fosphor, fft_win),
 offsetof(struct fosphor, img_waterfall), offsetof(struct fosphor, power), sizeof(struct fosphor));return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.dirname(ref), os.path.join(d, "t.c"), "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert [int(v) for v in out] == [20, 4120, 4144, 4176]


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device create() must fail (never compute)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gr_fosphor_b200.engine import Fosphor
    with pytest.raises(RuntimeError):
        Fosphor()
    from gr_fosphor_b200 import build
    from gr_fosphor_b200.dropin import FosphorCL
    with pytest.raises(RuntimeError):
        FosphorCL(build.LIB)


def test_c99_caller_builds_against_the_headers_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/dropin_min.c: a plain C99 translation unit (-pedantic) that includes the two public
    headers, links libfosphor_b200.so and drives the seven fosphor_cl_* entry points the way
    lib/fosphor/fosphor.c does.  Without a CUDA device init must report -EIO (exit code 2); with one
    the burst is processed (exit code 0)."""
    import shutil
    import subprocess
    import torch
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from gr_fosphor_b200 import build
    lib = build.build()
    exe = str(tmp_path / "dropin_min")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "dropin_min.c"), "-o", exe,
                           "-L", os.path.dirname(lib), "-lfosphor_b200", "-Wl,-rpath," + os.path.dirname(lib), "-lm"])
    rc = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert rc.returncode == 0, rc.stderr
    else:
        assert rc.returncode == 2, (rc.returncode, rc.stderr)
        assert "-EIO" in rc.stderr
    # examples/dropin_bench.c: the host-fed sink-frame loop of bench.py's e2e arm, from plain C
    exe2 = str(tmp_path / "dropin_bench")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-pedantic", "-D_POSIX_C_SOURCE=199309L",
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "dropin_bench.c"),
                           "-o", exe2, "-L", os.path.dirname(lib), "-lfosphor_b200",
                           "-Wl,-rpath," + os.path.dirname(lib), "-lm"])
    rc = subprocess.run([exe2, "3"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert rc.returncode == 0 and "Msamples_per_s" in rc.stdout, rc.stderr
    else:
        assert rc.returncode == 2, (rc.returncode, rc.stderr)


def test_reference_fosphor_c_links_against_the_dropin():
    """oracle/ref_link: the reference's UNMODIFIED lib/fosphor/fosphor.c (sole caller of the
    boundary) links against libfosphor_b200.so with --no-undefined, every fosphor_cl_* symbol
    resolved by it; without a GPU its fosphor_init() fails loudly (NULL) through the reference's own
    error path (fosphor.c:69-74: release after a failed cl init)."""
    import subprocess
    import torch
    so = os.path.join(ROOT, "oracle", "_ref", "libfosphor_facade_b200.so")
    if os.path.isdir("/root/reference/lib/fosphor"):
        from gr_fosphor_b200 import build
        build.build()
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_link"), "-s"])
    if not os.path.exists(so):
        pytest.skip("facade not built (no reference tree here)")
    und = subprocess.check_output(["nm", "-D", "--undefined-only", so]).decode()
    needed = sorted(set(re.findall(r"\b(fosphor_cl_[a-z_]+)", und)))
    assert needed == ["fosphor_cl_finish", "fosphor_cl_get_waterfall_position", "fosphor_cl_init",
                      "fosphor_cl_load_fft_window", "fosphor_cl_process", "fosphor_cl_release",
                      "fosphor_cl_set_histogram_range"]
    ldd = subprocess.check_output(["ldd", so]).decode()
    assert "libfosphor_b200.so" in ldd and "not found" not in ldd.split("libfosphor_b200.so")[1].splitlines()[0]
    if not torch.cuda.is_available():
        import facade_lib
        with pytest.raises(RuntimeError):
            facade_lib.FosphorFacade()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gr-fosphor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "fosphor_oracle" not in txt and "oracle_lib" not in txt, f
