"""Writes tests/golden/ref_julia_inputs.bin: the inputs of tests/golden/pic_small.npz as raw little-endian values that
tools/reference_vectors.jl can read with Base Julia alone (no NPZ.jl / HDF5.jl needed).

    int64  n, nh, nt, ns, nchi          float64  L, dt, w          float64[nchi] chi       float64[n] x0, v0

    python tools/export_reference_inputs.py
"""
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "pic_small.npz"))
x0, v0, chi = g["x0"], g["v0"], g["chi"]
with open(os.path.join(ROOT, "tests", "golden", "ref_julia_inputs.bin"), "wb") as f:
    f.write(struct.pack("<5q", x0.shape[0], int(g["nh"]), int(g["nt"]), int(g["ns"]), chi.shape[0]))
    f.write(struct.pack("<3d", float(g["L"]), float(g["dt"]), float(g["w"])))
    f.write(np.asarray(chi, dtype="<f8").tobytes())
    f.write(np.asarray(x0, dtype="<f8").tobytes())
    f.write(np.asarray(v0, dtype="<f8").tobytes())
print("wrote tests/golden/ref_julia_inputs.bin")
