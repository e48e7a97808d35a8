"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference hot path into ``oracle/_ref/``.

    python -m oracle.stage_ref            # run by __graft_entry__.build() when /root/reference is present

The reference is pure Python: "building" it means making its two hot-path packages
(``shapleyserver/fed_client_contribution`` and ``shapleyserver/federated_learning``) importable
where ``/root/reference`` does not exist (the GPU box).  They are copied byte for byte from where
they lie into ``oracle/_ref/shapleyserver/`` -- git-ignored (never part of the history), not
gpurun-ignored (travels with the snapshot like a built ``.so``) -- together with a manifest of
sha256 sums.  ``oracle/ref_shim.py`` imports them from there when ``/root/reference`` is absent;
``bench.py --impl reference`` and the ``cpu_baseline`` leg then time the reference's own
``Game.eval_utility`` (game.py:73-114) + ``evaluation`` (federated_learning/utils.py:864-926).
Nothing in the product package reads ``oracle/_ref``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("fed_client_contribution", "federated_learning")


def stage(reference_root: str = "/root/reference", dest: str = DEST) -> dict:
    src = os.path.join(reference_root, "shapleyserver")
    if not os.path.isdir(os.path.join(src, PACKAGES[0])):
        raise RuntimeError(f"no reference tree at {reference_root}")
    out_pkg = os.path.join(dest, "shapleyserver")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(out_pkg)
    manifest = {}
    init = os.path.join(src, "__init__.py")
    if os.path.exists(init):
        shutil.copy2(init, os.path.join(out_pkg, "__init__.py"))
    for pkg in PACKAGES:
        for dirpath, dirnames, filenames in os.walk(os.path.join(src, pkg)):
            dirnames[:] = [d for d in dirnames if d != "__pycache__"]
            rel = os.path.relpath(dirpath, src)
            os.makedirs(os.path.join(out_pkg, rel), exist_ok=True)
            for fn in filenames:
                if not fn.endswith(".py"):
                    continue
                s, d = os.path.join(dirpath, fn), os.path.join(out_pkg, rel, fn)
                shutil.copy2(s, d)
                with open(s, "rb") as f:
                    manifest[os.path.join("shapleyserver", rel, fn)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = stage(*(sys.argv[1:2] or ["/root/reference"]))
    print(f"staged {len(m)} files into {DEST}")
