"""Extracts the known-answer vectors of the reference's own test file into a JSON fixture.

Run in the build container only (needs /root/reference):  python tests/golden/make_reference_golden.py
Source: /root/reference/test/runtests.jl:6-8 (Rodrigues / scaling / projection known answers) and
:15-27 (residuals! golden vector, asserted with ``norm(true_residuals - r) == 0``).
The numbers are parsed out of the Julia source text, not retyped.
"""
import json
import os
import re

SRC = "/root/reference/test/runtests.jl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_runtests.json")


def vec(name, text):
    m = re.search(r"^%s\s*=\s*\[([^\]]*)\]" % re.escape(name), text, re.M)
    return [float(t) for t in m.group(1).split(",")]


def main():
    text = open(SRC).read()
    rod = re.search(r"Rodrigues_rotation\(\[([^\]]*)\],\s*\[([^\]]*)\]\)\s*==\s*\[([^\]]*)\]", text)
    sc = re.search(r"scaling_factor\(\[([^\]]*)\],\s*([\d.]+),\s*([\d.]+)\)\s*==\s*([\d.]+)", text)
    pr = re.search(r"projection\(([^)]*)\)\s*==\s*\[([^\]]*)\]", text)
    fx = {
        "source": "test/runtests.jl",
        "rodrigues": {"r": [float(t) for t in rod.group(1).split(",")],
                      "x": [float(t) for t in rod.group(2).split(",")],
                      "expect": [float(t) for t in rod.group(3).split(",")]},
        "scaling_factor": {"point": [float(t) for t in sc.group(1).split()], "k1": float(sc.group(2)),
                           "k2": float(sc.group(3)), "expect": float(sc.group(4))},
        "projection": {"args_x_y_z_rx_ry_rz_tx_ty_tz_f_k1_k2": [float(t) for t in pr.group(1).split(",")],
                       "expect": [float(t) for t in pr.group(2).split()]},
        "residuals": {"pt2d": vec("pt2d", text),
                      "cam_idx": [int(v) for v in vec("cam_idx", text)],
                      "pnt_idx": [int(v) for v in vec("pnt_idx", text)],
                      "x": vec("x", text), "nobs": 5, "npnts": 1,
                      "true_residuals": vec("true_residuals", text)},
    }
    with open(OUT, "w") as f:
        json.dump(fx, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
