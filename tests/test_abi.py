"""The C-ABI library loads and exports every symbol include/gat_b200.h declares; without a GPU the
product fails loudly (no CPU fallback); nothing under gat_b200/ touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gatb_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gat_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), "libgat_b200.so does not export %s" % s
    assert sorted(_lib.SYMBOLS) == syms
    assert lib.gatb_version() == 100


def test_header_is_plain_c_and_links(tmp_path):
    """include/gat_b200.h compiles as strict C99 (no C++ or torch types at the boundary) and a C program
    that uses only the header links against libgat_b200.so and runs (no device needed for gatb_version)"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "use.c"
    src.write_text('#include <stdio.h>\n#include "gat_b200.h"\n'
                   'int main(void) { gatb_ctx *c = 0; (void)c; printf("%d %d\\n", gatb_version(), GATB_NCOUNTERS); return GATB_OK; }\n')
    lib_dir = os.path.join(ROOT, "gat_b200", "lib")
    exe = tmp_path / "use"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", lib_dir, "-lgat_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out == ["100", "7"]


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gat_b200 import device, _lib
    with pytest.raises(_lib.GatB200Error) as e:
        device.Context(0)
    assert e.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(e.value)
    # the reference-style operators fail the same way instead of computing on the host
    from gat_b200 import engine as Engine
    from gat_b200.segmentlist import SegmentList
    s = SegmentList(iter=[(0, 10)], normalize=True)
    with pytest.raises(_lib.GatB200Error):
        Engine.CounterNucleotideOverlap()(s, s)
    with pytest.raises(_lib.GatB200Error):
        Engine.SamplerAnnotator().sample(s, s)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/_ref|liboracle|gat_oracle\.h", text, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_product_never_builds_or_loads_the_emulated_kernels():
    """tests/emu (the SIMT-emulated host build of the .cu sources) is test infrastructure like the oracle: the
    package must not know how to build or find it; the only trace allowed is the loader's refusal of it"""
    bad = []
    for top in ("gat_b200", "integration"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f)).read()
                    if re.search(r"build_emu|libgat_b200_emu|GATB_EMU|gatb_emu::|emu_runtime", text):
                        bad.append(os.path.join(dirpath, f))
    for f in ("bench.py", "__graft_entry__.py"):
        text = open(os.path.join(ROOT, f)).read()
        if re.search(r"build_emu|libgat_b200_emu|tests/emu|tests\.emu", text):
            bad.append(f)
    assert not bad, bad
