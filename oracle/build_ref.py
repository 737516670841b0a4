"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference package, installed (not copied into history).

TEST / BENCH INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but
not gpurun-ignored, so the installed package travels to the GPU box, where ``/root/reference`` does not exist.  There
it serves (a) as the CPU arm of ``bench.py`` (``--impl reference`` and ``cpu_baseline.kind == "reference"``) and (b) as
the caller in the drop-in tests (the unmodified ``optimize_vp`` / ``_sieve`` / ``minimize_adam`` running on top of
``pyvbmc_b200.install()``).

    python -m oracle.build_ref          # needs /root/reference; run by __graft_entry__.build()

What it does: packs the package directory ``/root/reference/pyvbmc`` (minus ``testing/`` fixtures and bytecode
caches) VERBATIM into one archive, ``oracle/_ref/pyvbmc_ref.zip``, at build time; ``oracle.ref_loader`` unpacks it into
a temporary directory when it is needed and puts that directory on ``sys.path``.  A ``pip install --target`` was tried
first and does not work: upstream's ``pyproject.toml`` lists ``packages = ["pyvbmc", "pyvbmc.examples"]`` only, so the
wheel built offline contains ``pyvbmc/__init__.py`` but none of the subpackages (``vbmc/``, ``entropy/``,
``variational_posterior/`` ...).  The third-party imports that cannot be installed offline (gpyreg, cma, corner,
matplotlib, plotly, imageio) are stubbed by ``oracle.ref_loader``; the hot path needs none of them.  Nothing is edited.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("PYVBMC_REFERENCE_ROOT", "/root/reference")
STAMP = os.path.join(TARGET, ".built_from")


def source_id() -> str:
    """Cheap identity of the reference checkout: file count + total size + newest mtime of pyvbmc/**/*.py."""
    n = size = 0
    newest = 0.0
    for root, _dirs, files in os.walk(os.path.join(SOURCE, "pyvbmc")):
        for f in files:
            if f.endswith(".py"):
                st = os.stat(os.path.join(root, f))
                n, size, newest = n + 1, size + st.st_size, max(newest, st.st_mtime)
    return f"{SOURCE}:{n}:{size}:{int(newest)}"


ARCHIVE = os.path.join(TARGET, "pyvbmc_ref.zip")


def built() -> bool:
    return os.path.isfile(ARCHIVE)


def build(force: bool = False) -> bool:
    """Returns True if ``oracle/_ref`` holds the reference afterwards (False: no reference tree here and nothing
    prebuilt)."""
    if not os.path.isdir(os.path.join(SOURCE, "pyvbmc")):
        return built()
    sid = source_id()
    if built() and not force and os.path.isfile(STAMP) and open(STAMP).read().strip() == sid:
        return True
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET)
    import zipfile

    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        base = os.path.join(SOURCE, "pyvbmc")
        for root, dirs, files in os.walk(base):
            dirs[:] = sorted(d for d in dirs if d not in ("__pycache__", "testing"))
            for f in sorted(files):
                if f.endswith(".pyc"):
                    continue
                full = os.path.join(root, f)
                z.write(full, os.path.join("pyvbmc", os.path.relpath(full, base)))
        if os.path.isfile(os.path.join(SOURCE, "LICENSE")):
            z.write(os.path.join(SOURCE, "LICENSE"), "LICENSE")
    with open(STAMP, "w") as fh:
        fh.write(sid + "\n")
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "NOT available (no reference tree at " + SOURCE + ")")
    sys.exit(0 if ok else 1)
