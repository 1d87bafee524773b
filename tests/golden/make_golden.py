#!/usr/bin/env python3
"""
Generate tests/golden/ref_outputs.npz by running the UNMODIFIED reference extension
(/root/reference/simplestereo/_passive.cpp compiled by `make -C oracle ref` into oracle/_ref) on
the case list in cases.py.  Also stores the md5 of every input so the tests can detect drift of the
deterministic input builders (cv2 / numpy versions).

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

The PNGs next to this file are the reference's own fixtures (examples/res/tsukuba/): the rectified
Tsukuba pair and disparityASW.png, its only known-answer image for this path.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import oracle  # noqa: E402
from tests.golden import cases  # noqa: E402


def main():
    oracle.build(ref=True)
    assert oracle.ref_available(), "oracle/_ref missing"
    out = {}
    for name, spec, kw in cases.ASW_CASES:
        l, r = cases.load_inputs(spec)
        t0 = time.time()
        out[name] = oracle.ref_asw(l, r, **kw)
        out["md5_" + name] = np.frombuffer(hashlib.md5(l.tobytes() + r.tobytes()).digest(), np.uint8)
        print(f"{name:32s} {l.shape} {time.time() - t0:6.2f}s", flush=True)
    for name, spec, kw in cases.GSW_CASES:
        l, r = cases.load_inputs(spec)
        t0 = time.time()
        out[name] = oracle.ref_gsw(l, r, **kw)
        out["md5_" + name] = np.frombuffer(hashlib.md5(l.tobytes() + r.tobytes()).digest(), np.uint8)
        print(f"{name:32s} {l.shape} {time.time() - t0:6.2f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("wrote ref_outputs.npz", os.path.getsize(os.path.join(HERE, "ref_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
