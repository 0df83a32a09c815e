"""Regenerates tests/golden/km_golden.json by running the REFERENCE's own src/km.cpp (compiled verbatim
into oracle/_ref/libkm_ref.so; needs /root/reference, i.e. this container, not the GPU box)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden", "km_golden.json")
    os.chdir(tempfile.mkdtemp())  # Km::output writes Corres.txt (src/km.cpp:148)
    cases = []
    G1 = [[-5, -2, -100], [-4, -2, -6], [-100, -1, -7]]
    CD = np.array([[11, 19, 4, 40, 10, 31], [17, 10, 16, 39, 17, 36], [20, 42, 5, 28, 11, 29],
                   [50, 21, 32, 24, 47, 32], [18, 26, 6, 7, 12, 38], [23, 36, 27, 35, 48, 30],
                   [22, 24, 7, 21, 13, 46]], float)

    def add(name, W, eps=0.01):
        W = np.asarray(W, float)
        m = oracle.km_solve(W, eps, "ref")
        cases.append(dict(name=name, eps=eps, W=W.tolist(), match=m.tolist()))

    add("G1 src/km.cpp:237-260", G1)
    add("G2 img/GH-ICPworkflow.jpg (e),(f)", oracle.km_graph(CD, 30.0))
    rng = np.random.default_rng(42)
    for n in (5, 12, 25):
        cd = np.round(rng.random((n, n - 2)) * 40, 1)
        add(f"random n={n} penalty=15", oracle.km_graph(cd, 15.0))
    json.dump(dict(generator="tests/golden/make_km_golden.py (runs the reference's src/km.cpp via oracle/_ref)",
                   cases=cases), open(out, "w"))
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
