#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN code, executed here.

Builds oracle/_ref/libmeso_ref.so (the reference's hot-path headers compiled unmodified from /root/reference against
the third-party stand-ins in oracle/ref_shim/; see oracle/ref_driver.cpp), runs the battery in tests/refprobe.py
through it and stores every output in tests/golden/ref_build.npz.  The GPU box has no /root/reference: there the
committed file (and the prebuilt .so, which travels with the snapshot) stand in for it.

    python tools/gen_golden_from_ref_build.py          # needs /root/reference
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refprobe  # noqa: E402


def main():
    so = refprobe.build_ref()
    if so is None or not os.path.isdir(refprobe.REF_ROOT):
        sys.exit("the reference tree is not available here; nothing generated")
    out = refprobe.probe(refprobe.RefBackend(so))
    np.savez_compressed(refprobe.GOLDEN, **out)
    n = sum(np.asarray(v).nbytes for v in out.values())
    print(f"{refprobe.GOLDEN}: {len(out)} arrays, {n} bytes raw, {os.path.getsize(refprobe.GOLDEN)} bytes on disk")
    import orc
    draw = refprobe.probe_draw(refprobe.RefBackend(so), orc)
    np.savez_compressed(refprobe.DRAW_GOLDEN, **draw)
    print(f"{refprobe.DRAW_GOLDEN}: {len(draw)} arrays, {os.path.getsize(refprobe.DRAW_GOLDEN)} bytes on disk")
    print("sphere: %d blocks, %d instances over %d chunks" % (out["sphere_counts"].sum(), out["sphere_instances"].sum(), len(out["sphere_counts"])))
    print("terrain: %d blocks, %d instances over %d chunks" % (out["terrain_counts"].sum(), out["terrain_instances"].sum(), len(out["terrain_counts"])))


if __name__ == "__main__":
    main()
