"""Recipe for ``oracle/_ref/``: the reference's own pure-Python packages, where a box without ``/root/reference``
(the GPU box) can import them.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref            # in the build container; __graft_entry__.build() calls it

Copies ``utils/ model/ trajectories/ control/ cbf/ obstacles/`` (``*.py`` only, unmodified) from the reference tree
into ``oracle/_ref/``.  The directory is git-ignored (no reference source enters the history) but not gpurun-ignored,
so it travels to the GPU box like the built ``.so``.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg then
time the reference's trajectories / controllers / CBF builder / QP tracker themselves (``kind: "reference"``,
oracle/ref_pipeline.py) instead of the numpy port; the env step and ``cvxopt.solvers.qp`` stay the oracle's
(neither is in the reference tree).
"""
from __future__ import annotations

import os
import shutil

PACKAGES = ("utils", "model", "trajectories", "control", "cbf", "obstacles")
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def build(src="/root/reference", dest=DEST) -> bool:
    """-> True if ``dest`` holds the packages afterwards (copied now or already present)."""
    if not os.path.isdir(os.path.join(src, "cbf")):
        return os.path.isdir(os.path.join(dest, "cbf"))
    for pkg in PACKAGES:
        out = os.path.join(dest, pkg)
        if os.path.isdir(out):
            shutil.rmtree(out)
        shutil.copytree(os.path.join(src, pkg), out, ignore=lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))])
    with open(os.path.join(dest, "README"), "w") as f:
        f.write("Unmodified copy of the reference's pure-Python packages (oracle/build_ref.py). Not product source; git-ignored.\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref:", "ready" if build() else "reference tree not available")
