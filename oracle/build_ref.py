#!/usr/bin/env python
"""Build the UNMODIFIED-ALGORITHM reference (AndreasHeger/gat 1.3.6) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gat_b200/) may import oracle/_ref.
Users: tests/ (golden-vector generation), bench.py --impl reference / cpu_baseline.

The reference is Python + Cython (no CUDA, no C++).  Its sources are read from where they lie
(/root/reference, read-only); they are copied to a scratch directory under /tmp, five mechanical,
non-semantic compatibility patches are applied there (Cython 3 / numpy 2 / Python 3),
the four extensions are compiled, and only the *installed package* (gat/*.py, gat/*.so, scripts/*.py) is
written to oracle/_ref/, which is git-ignored (not gpurun-ignored: it travels to the GPU box like our
own .so files).  No reference source ever enters the git history.

The patches (nothing touches the algorithm):
  1. Cython directives language_level=2, legacy_implicit_noexcept=True   (Py2-style syntax in .pyx)
  2. numpy.int_t -> numpy.int64_t, numpy.float_t -> numpy.float64_t      (ctypedefs removed in numpy 2)
  3. numpy.int -> numpy.int64, numpy.float -> numpy.float64              (aliases removed in numpy 1.24)
  4. xrange -> range                                                      (SamplerSegments only)
  5. numpy.random.random_integers(a,b,n) -> numpy.random.randint(a,b+1,n) (Storey bootstrap only)

Run:  python oracle/build_ref.py           (needs /root/reference; a no-op success if already built
                                            and the reference is absent, e.g. on the GPU box)
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("GAT_REFERENCE", "/root/reference")


def already_built():
    return bool(glob.glob(os.path.join(DEST, "gat", "Engine*.so")))


def patch_text(text, is_pyx):
    text = re.sub(r"numpy\.int_t\b", "numpy.int64_t", text)
    text = re.sub(r"numpy\.float_t\b", "numpy.float64_t", text)
    # numpy.int / numpy.float used as dtypes (not numpy.int64, numpy.integer, numpy.floating, ...)
    text = re.sub(r"numpy\.int\b(?!_)", "numpy.int64", text)
    text = re.sub(r"numpy\.float\b(?!_)", "numpy.float64", text)
    text = re.sub(r"\bxrange\(", "range(", text)
    text = re.sub(r"numpy\.random\.random_integers\(0, m - 1, m\)",
                  "numpy.random.randint(0, m, m)", text)
    return text


def build():
    if not os.path.isdir(os.path.join(REFERENCE, "gat")):
        if already_built():
            print("oracle/_ref: reference sources absent, using prebuilt %s" % DEST)
            return 0
        print("oracle/_ref: reference sources absent and nothing prebuilt", file=sys.stderr)
        return 1

    work = tempfile.mkdtemp(prefix="gat_ref_build_")
    try:
        for sub in ("gat", "utils", "scripts"):
            shutil.copytree(os.path.join(REFERENCE, sub), os.path.join(work, sub))
        for fn in glob.glob(os.path.join(work, "gat", "*.py*")) + \
                glob.glob(os.path.join(work, "gat", "*.pxd")) + \
                glob.glob(os.path.join(work, "scripts", "*.py")):
            with open(fn) as f:
                text = f.read()
            new = patch_text(text, fn.endswith(".pyx"))
            if new != text:
                os.chmod(fn, 0o644)
                with open(fn, "w") as f:
                    f.write(new)
        # PointList.pyx is a stale duplicate that the reference's setup.py does not build
        setup_py = os.path.join(work, "setup_ref.py")
        with open(setup_py, "w") as f:
            f.write('''
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
mods = ["CoordinateList", "SegmentList", "PositionList", "Engine"]
exts = [Extension("gat." + m, ["gat/%s.pyx" % m, "utils/gat_utils.c"],
                  libraries=["z", "rt"], include_dirs=["utils", numpy.get_include()],
                  extra_compile_args=["-O2", "-w"], language="c") for m in mods]
setup(name="gat", ext_modules=cythonize(
    exts, include_path=["gat"], quiet=True,
    compiler_directives=dict(language_level=2, legacy_implicit_noexcept=True)))
''')
        env = dict(os.environ)
        r = subprocess.run([sys.executable, "setup_ref.py", "build_ext", "--inplace", "-j", "4"],
                           cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout[-6000:], file=sys.stderr)
            return r.returncode

        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        os.makedirs(os.path.join(DEST, "gat"))
        os.makedirs(os.path.join(DEST, "scripts"))
        for fn in glob.glob(os.path.join(work, "gat", "*.py")) + glob.glob(os.path.join(work, "gat", "*.so")):
            shutil.copy(fn, os.path.join(DEST, "gat"))
        for fn in glob.glob(os.path.join(work, "scripts", "*.py")):
            shutil.copy(fn, os.path.join(DEST, "scripts"))
        with open(os.path.join(DEST, "README"), "w") as f:
            f.write("Built by oracle/build_ref.py from %s (gat 1.3.6) with 5 mechanical patches.\n"
                    "Git-ignored test infrastructure; never imported by gat_b200/.\n" % REFERENCE)
        print("oracle/_ref: built reference into %s" % DEST)
        return 0
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(build())
