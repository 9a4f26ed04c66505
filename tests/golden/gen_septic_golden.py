"""Writes tests/golden/septic.json: F_p^7 and septic-curve operations as the REFERENCE'S OWN C++ computes them
(crates/core/machine/include/kb31_septic_extension_t.hpp, compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`):
products, reciprocals, both Frobenius maps, the curve formula, square roots of squares, non-squares, and sums of two curve
points with different x (the twin's doubling branch differs from crates/stark/src/septic_curve.rs:62-78 and is left out).
Inputs: the vectors of the reference's own tests (septic_extension.rs tests: (i + 3, 2i + 6, 5i + 17, 6i + 91, 8i + 37, 11i +
35, 14i + 33)) and seeded random elements.  Run in the build container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402

P = 0x7F000001


def inputs():
    rng = np.random.default_rng(77)
    v = [np.array([i + 3, 2 * i + 6, 5 * i + 17, 6 * i + 91, 8 * i + 37, 11 * i + 35, 14 * i + 33], np.uint32) for i in range(24)]
    v += [rng.integers(0, P, 7).astype(np.uint32) for _ in range(24)]
    return v


def curve_points(k):
    """Points of the curve found through the oracle's lift (x from seeded messages), as 14 canonical words."""
    from ziren_b200 import tracegen as tg
    rows = o.global_trace(tg.synthetic_global_events(k, seed=5), k)
    return [rows[i, 16:30].copy() for i in range(k)]


if __name__ == "__main__":
    v = inputs()
    g = {"source": "crates/core/machine/include/kb31_septic_extension_t.hpp via oracle/_ref/libzkref_core.so", "cases": []}
    for i, a in enumerate(v):
        b = v[(i * 7 + 3) % len(v)]
        sq = o.ref_septic_op("mul", a, a)
        root = o.ref_septic_op("sqrt", sq)
        g["cases"].append({"a": a.tolist(), "b": b.tolist(),
                           "mul": o.ref_septic_op("mul", a, b).tolist(), "inv": o.ref_septic_op("inv", a).tolist(),
                           "frobenius": o.ref_septic_op("frobenius", a).tolist(),
                           "double_frobenius": o.ref_septic_op("double_frobenius", a).tolist(),
                           "curve_formula": o.ref_septic_op("curve_formula", a).tolist(),
                           "sqrt_of_square": root.tolist(), "a_is_square": o.ref_septic_op("sqrt", a) is not None})
    pts = curve_points(16)
    g["curve_add"] = [{"p": pts[i].tolist(), "q": pts[i + 1].tolist(), "sum": o.ref_septic_op("curve_add", pts[i], pts[i + 1]).tolist()}
                      for i in range(15)]
    json.dump(g, open(os.path.join(ROOT, "tests", "golden", "septic.json"), "w"))
    print(len(g["cases"]), len(g["curve_add"]), sum(c["a_is_square"] for c in g["cases"]))
