"""Build recipes for the parity oracle (TEST INFRASTRUCTURE ONLY).

* ``build_c_oracle()``  -> ``oracle/liblattice_oracle.so`` from ``oracle/lattice_oracle.c`` (gcc).
* ``build_ref()``       -> ``oracle/_ref/sgp_ref_lattice.so``: the reference's own CPU
  filter, compiled from the sources where they lie under ``/root/reference`` (nothing
  is copied into this repository), together with ``oracle/ref_harness.cpp`` which
  exposes its intermediates.  Only possible where ``/root/reference`` exists (the
  build container); the GPU box receives the prebuilt ``.so``.

The reference ships no build system for this path -- it JIT-compiles
``cpp/lattice.cpp`` with ``torch.utils.cpp_extension.load`` (bilateral_kernel.py:62-74)
-- so the recipe here is the same mechanism with ``-O3`` (bit-identical output to the
as-shipped flags; the code has no FMA contraction or reassociation on x86-64).

Nothing under the product package imports this module.
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SGP_REFERENCE_ROOT", "/root/reference")
REF_CPP_DIR = os.path.join(REF_ROOT, "gpytorch_lattice_kernel", "cpp")
C_ORACLE_SO = os.path.join(HERE, "liblattice_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_NAME = "sgp_ref_lattice"
REF_SO = os.path.join(REF_DIR, REF_NAME + ".so")
# The reference's hash table has a defect (permutohedral.h:104-106 computes the bucket
# h = hash % capacity BEFORE lookupOffset may double the capacity at :61-63, so the
# one lookup that triggers each doubling probes from a stale bucket: the key is then
# stored where later lookups cannot find it, which orphans / duplicates one lattice
# point per doubling, first at M = 16383).  To pin the oracle at sizes beyond that,
# the same harness is also compiled against a temporary copy of the header with that
# single statement re-ordered ("fixed" build).  The copy lives in a temp directory
# outside the repository and is deleted after the build.
REF_FIXED_NAME = "sgp_ref_lattice_fixed"
REF_FIXED_SO = os.path.join(REF_DIR, REF_FIXED_NAME + ".so")
_DEFECT_LINE = "size_t h = hash(k) % capacity;"
_FIXED_LINE = "if (filled >= (capacity / 2) - 1) grow(); size_t h = hash(k) % capacity;"


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "lattice_oracle.c")
    if force or _stale(C_ORACLE_SO, [src]):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-fPIC",
               "-Wall", "-o", C_ORACLE_SO, src, "-lm"]
        subprocess.run(cmd, check=True)
    return C_ORACLE_SO


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_CPP_DIR, "permutohedral.h"))


def build_ref(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference CPU filter + harness into oracle/_ref/. Returns the .so path or None."""
    harness = os.path.join(HERE, "ref_harness.cpp")
    if not reference_available():
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [harness, os.path.join(REF_CPP_DIR, "permutohedral.h")]
    if not force and not _stale(REF_SO, srcs) and not _stale(REF_FIXED_SO, srcs):
        return REF_SO
    os.makedirs(REF_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load

    load(name=REF_NAME, sources=[harness], extra_include_paths=[REF_CPP_DIR],
         extra_cflags=["-O3", "-w"], build_directory=REF_DIR, verbose=verbose, is_python_module=False)
    _build_fixed(harness, verbose)
    return REF_SO


def _build_fixed(harness: str, verbose: bool) -> None:
    import shutil
    import tempfile
    from torch.utils.cpp_extension import load

    hdr = open(os.path.join(REF_CPP_DIR, "permutohedral.h")).read()
    if hdr.count(_DEFECT_LINE) != 1:
        raise RuntimeError("reference header changed: cannot locate the bucket computation to re-order")
    tmp = tempfile.mkdtemp(prefix="sgp_ref_fixed_")
    try:
        with open(os.path.join(tmp, "permutohedral.h"), "w") as f:
            f.write(hdr.replace(_DEFECT_LINE, _FIXED_LINE))
        bdir = os.path.join(tmp, "build")
        os.makedirs(bdir)
        load(name=REF_FIXED_NAME, sources=[harness], extra_include_paths=[tmp],
             extra_cflags=["-O3", "-w"], build_directory=bdir, verbose=verbose, is_python_module=False)
        shutil.copy(os.path.join(bdir, REF_FIXED_NAME + ".so"), REF_FIXED_SO)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def load_ref(fixed: bool = False):
    """Import a prebuilt reference module (``filter``, ``structure``) or return None.

    ``fixed=False``: the unmodified reference.  ``fixed=True``: the build with the
    bucket computation re-ordered (see the note at the top of this file)."""
    name, so = (REF_FIXED_NAME, REF_FIXED_SO) if fixed else (REF_NAME, REF_SO)
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == "__main__":
    print(build_c_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv, verbose=True))
