#!/usr/bin/env python
"""Dump outputs of the REAL reference modules for the inputs of the GPU parity tests
(see RECIPE.md).  Runs only where the reference's own `particlesfm` module (HP1) / `gcolmap`
binary (HP2) have been built; it is never imported by the product or by the test suite."""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

HP1_CASES = [(1, 32, 48, 0), (31, 32, 48, 1), (256, 64, 96, 2), (257, 64, 96, 3), (5000, 128, 256, 4), (70001, 218, 512, 5)]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def dump_hp1():
    import particlesfm                       # the REFERENCE's module (PYTHONPATH points at its build dir)
    if not hasattr(particlesfm, "optimize_location") or "particle-sfm_b200" in getattr(particlesfm, "__file__", ""):
        raise SystemExit("this is not the reference's particlesfm module")
    from particlesfm_b200 import synthetic as syn
    for n, h, w, seed in HP1_CASES:
        uv12, r1, r2, sc, f12 = syn.make_traj_inputs(n, h, w, seed=seed)
        out = particlesfm.optimize_location(uv12, r1, r2, sc, f12, n, w, h)
        path = os.path.join(GOLD, f"ref_traj_{n}_{h}_{w}_{seed}.npz")
        np.savez_compressed(path, out=np.asarray(out), inputs_sha256=digest(uv12, r1, r2, sc, f12))
        print("wrote", path)


def dump_hp2(gcolmap):
    raise SystemExit("HP2: write the test problems with sfm/colmap_utils (database + model), run\n  " + gcolmap +
                     " global_mapper ... and store the refined model; see RECIPE.md — not automated yet")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--hp1", action="store_true")
    ap.add_argument("--hp2", action="store_true")
    ap.add_argument("--gcolmap", default="gcolmap")
    a = ap.parse_args()
    if a.hp1:
        dump_hp1()
    if a.hp2:
        dump_hp2(a.gcolmap)
