"""Recipe for ``oracle/_ref``: the reference's own model sources, UNMODIFIED, where the GPU box can import them.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; the GPU box receives the repo
snapshot.  ``python -m oracle.make_ref`` (called by ``__graft_entry__.build()`` when the mount is present) copies the
few Python files of the hot path -- ``models/*.py`` and ``pflow/models/*.py``, nothing else -- byte for byte into the
git-ignored ``oracle/_ref/`` (listed in .gitignore, not in .gpurunignore: it travels like the built ``.so``, and never
enters the history).  ``oracle/ref_import.py`` imports the reference from the mount when it is there and from this
copy otherwise, so ``bench.py --impl reference`` times the reference's own ``FlowModel.generate_samples`` on the
box's host cores, and the parity tests can check the oracle against the real modules on the box too.
"""
from __future__ import annotations

import filecmp
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PACKAGES = ("models", os.path.join("pflow", "models"))


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, "models", "flow_model.py")):
        if verbose:
            print("oracle/_ref: /root/reference is not mounted here; keeping whatever copy exists")
        return os.path.isfile(os.path.join(DST, "models", "flow_model.py"))
    n = 0
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(SRC, pkg))):
            if f.endswith(".py"):
                a, b = os.path.join(SRC, pkg, f), os.path.join(DST, pkg, f)
                if not (os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)):
                    shutil.copyfile(a, b)
                n += 1
    init = os.path.join(DST, "pflow", "__init__.py")                     # the mount's pflow/ is a namespace package; make the copy importable the same way
    if not os.path.isfile(os.path.join(SRC, "pflow", "__init__.py")) and os.path.isfile(init):
        os.remove(init)
    if verbose:
        print(f"oracle/_ref: {n} reference source files in place")
    return True


if __name__ == "__main__":
    make_ref()
