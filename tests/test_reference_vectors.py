"""Pins the PIC oracle (and the GPU path) to vectors produced by the REAL reference packages, when they exist.

tests/golden/ref_julia_outputs.bin is written by `julia tools/reference_vectors.jl` (VlasovMethods / PoissonSolvers /
ParticleMethods through ReducedBasisMethods' own call sites: scripts/bump_on_tail_1_training.jl:67-97) from the inputs of
tests/golden/pic_small.npz.  No Julia exists in the build image, so the file is absent there and the comparison tests
SKIP with the message "PIC parity unpinned"; the machinery itself (binary layout, convention detection, bars) is
exercised every run by a self-test that plays the reference's role with the oracle under NON-default conventions.

Conventions not pinned by /root/reference are detected, reported, and then held fixed:
sign / gauge / cyclic tap shift of the stored Phi, and whether IC.A at a save is the kick's acceleration or the
post-update! field.  Bars (north_star): rhs, phi, a <= 1e-12 relative; W, K, M traces <= 1e-10.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel
from oracle import pic_numpy as pn

REF = os.path.join(GOLDEN, "ref_julia_outputs.bin")
MAGIC = 20261020.0
SKIP = ("PIC parity unpinned: tests/golden/ref_julia_outputs.bin is absent -- run `julia tools/reference_vectors.jl` in an "
        "environment with ReducedBasisMethods.jl (see INTEGRATION.md, 'Pinning parity')")


def read_vectors(path):
    """Layout documented in tools/reference_vectors.jl."""
    a = np.fromfile(path, dtype="<f8")
    assert a[0] == MAGIC, "not a reference-vector file"
    n, nh, nt, ns, nchi = (int(q) for q in a[1:6])
    pos = [6]

    def take(m, shape=None):
        out = a[pos[0]:pos[0] + m]
        assert out.shape[0] == m, "truncated reference-vector file"
        pos[0] += m
        return out if shape is None else out.reshape(shape)
    ref = {"n": n, "nh": nh, "nt": nt, "ns": ns, "nchi": nchi,
           "rhs": take(nh), "phi": take(nh), "a": take(n), "W0": float(take(1)[0]), "runs": []}
    for _ in range(nchi):
        ref["runs"].append({"X": take(n * ns, (ns, n)), "V": take(n * ns, (ns, n)), "A": take(n * ns, (ns, n)),
                            "Phi": take(nh * ns, (ns, nh)), "W": take(ns), "K": take(ns), "M": take(ns)})
    assert pos[0] == a.shape[0], "trailing data in the reference-vector file"
    return ref


def write_vectors(path, n, nh, nt, ns, first, runs):
    """The writer the Julia script mirrors (used by the self-test)."""
    parts = [np.array([MAGIC, n, nh, nt, ns, len(runs)], dtype="<f8"), first["rhs"], first["phi"], first["a"],
             np.array([first["W0"]])]
    for o in runs:
        parts += [o[k].reshape(-1) for k in ("X", "V", "A", "Phi", "W", "K", "M")]
    np.concatenate([np.asarray(q, dtype="<f8").reshape(-1) for q in parts]).tofile(path)


def detect_phi_convention(phi_ref, phi_canon):
    """-> (Conventions-compatible dict, residual): sign, cyclic shift and additive constant that map the canonical
    coefficients (zero mean, tap 0 = own cell, a = +phi') onto the stored ones."""
    best = None
    for sign in (1.0, -1.0):
        for shift in range(-3, 4):
            cand = sign * np.roll(phi_canon, shift)
            gauge = float(np.mean(phi_ref - cand))
            err = rel(cand + gauge, phi_ref)
            if best is None or err < best[1]:
                best = ({"phi_sign": sign, "tap_shift": shift, "gauge": gauge}, err)
    return best


def compare(ref, inputs, produce_first, produce_run, who):
    """Shared by the oracle and the GPU comparison.  produce_first(chi) -> dict(rhs, phi (canonical), a, W0);
    produce_run(chi, conv) -> dict(X, V, A, Phi, W, K, M) with Phi/A stored under `conv`.  Returns the report."""
    chis = inputs["chi"]
    first = produce_first(float(chis[1]))
    report = {"who": who}
    # deposit: invariant up to the cyclic tap shift (skipped if the solver exposes no rhs field)
    if np.all(np.isfinite(ref["rhs"])):
        shifts = {k: rel(np.roll(first["rhs"], k), ref["rhs"]) for k in range(-3, 4)}
        k_rhs = min(shifts, key=shifts.get)
        # the real solver may store the deposit after mean removal and/or with the sign of the right-hand side
        alt = {}
        for k in range(-3, 4):
            r0 = np.roll(first["rhs"], k)
            alt[("demeaned", k)] = rel(r0 - r0.mean(), ref["rhs"])
            alt[("negated-demeaned", k)] = rel(-(r0 - r0.mean()), ref["rhs"])
        form, err_rhs = ("raw", k_rhs), shifts[k_rhs]
        ka = min(alt, key=alt.get)
        if alt[ka] < err_rhs:
            form, err_rhs = ka, alt[ka]
        report["rhs"] = {"form": form, "err": err_rhs}
        assert err_rhs < 1e-12, f"{who}: deposit differs from the reference ({report['rhs']})"
    pc, err_phi = detect_phi_convention(ref["phi"], first["phi"])
    report["phi"] = {"convention": pc, "err": err_phi}
    assert err_phi < 1e-12, f"{who}: potential differs from the reference under every sign/shift/gauge ({report['phi']})"
    assert rel(first["a"], ref["a"]) < 1e-12, f"{who}: gather differs from the reference"
    assert abs(first["W0"] - ref["W0"]) <= 1e-10 * abs(ref["W0"]), f"{who}: field energy differs (chi^2 W convention?)"
    # A at a save: decide on the first sample, then hold
    timing = None
    for cand in ("kick", "post_update"):
        conv = pn.Conventions(a_at_save=cand, **pc)
        o = produce_run(float(chis[0]), conv)
        if rel(o["A"], ref["runs"][0]["A"]) < 1e-9:
            timing = cand
            break
    report["a_at_save"] = timing
    assert timing is not None, f"{who}: IC.A matches neither the kick's acceleration nor the post-update field"
    conv = pn.Conventions(a_at_save=timing, **pc)
    report["conventions"] = repr(conv)
    for p, chi in enumerate(chis):
        o, g = produce_run(float(chi), conv), ref["runs"][p]
        for k in ("W", "K", "M"):
            assert rel(o[k], g[k]) < 1e-10, (who, k, p, rel(o[k], g[k]))
        assert np.max(np.abs(o["X"] - g["X"])) < 1e-9 and np.max(np.abs(o["V"] - g["V"])) < 1e-9, (who, "X/V", p)
        assert rel(o["A"], g["A"]) < 1e-9 and rel(o["Phi"], g["Phi"]) < 1e-9, (who, "A/Phi", p)
    return report


@pytest.fixture(scope="module")
def inputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "pic_small.npz"))
    return {"x0": g["x0"], "v0": g["v0"], "w": float(g["w"]), "L": float(g["L"]), "nh": int(g["nh"]), "nt": int(g["nt"]),
            "ns": int(g["ns"]), "dt": float(g["dt"]), "chi": g["chi"]}


def oracle_first(inp, chi):
    rhs = pn.deposit(inp["x0"], inp["w"], inp["L"], inp["nh"])
    phi, W0 = pn.poisson(rhs, inp["nh"], inp["L"], chi)
    return {"rhs": rhs, "phi": phi, "a": pn.gather(phi, inp["x0"], inp["L"], inp["nh"]), "W0": W0}


def oracle_run(inp, chi, conv):
    return pn.integrate(inp["x0"], inp["v0"], inp["w"], inp["L"], inp["nh"], chi, inp["dt"], inp["nt"], inp["ns"], conv)


def test_inputs_file_matches_the_golden_fixture(inputs, golden_dir):
    """tests/golden/ref_julia_inputs.bin (what the Julia script reads) holds exactly the inputs of pic_small.npz."""
    import struct

    raw = open(os.path.join(golden_dir, "ref_julia_inputs.bin"), "rb").read()
    n, nh, nt, ns, nchi = struct.unpack_from("<5q", raw, 0)
    L, dt, w = struct.unpack_from("<3d", raw, 40)
    assert (n, nh, nt, ns, nchi) == (inputs["x0"].shape[0], inputs["nh"], inputs["nt"], inputs["ns"], inputs["chi"].shape[0])
    assert (L, dt, w) == (inputs["L"], inputs["dt"], inputs["w"])
    body = np.frombuffer(raw, dtype="<f8", offset=64)
    assert np.array_equal(body[:nchi], inputs["chi"]) and np.array_equal(body[nchi:nchi + n], inputs["x0"])
    assert np.array_equal(body[nchi + n:], inputs["v0"])


def test_machinery_detects_foreign_conventions(inputs, tmp_path):
    """Self-test (runs everywhere): a file written under NON-default conventions -- -phi, shifted taps, non-zero gauge,
    post-update A -- is read back, the conventions are detected, and every bar holds.  This is what happens to the real
    file if PoissonSolvers.jl differs from SURVEY App. A.4 in any of them."""
    truth = pn.Conventions(phi_sign=-1.0, gauge=0.125, tap_shift=-2, a_at_save="post_update")
    first = oracle_first(inputs, float(inputs["chi"][1]))
    stored = dict(first, phi=truth.stored_phi(first["phi"]), rhs=np.roll(first["rhs"], -2))
    runs = [oracle_run(inputs, float(c), truth) for c in inputs["chi"]]
    f = str(tmp_path / "fake_ref.bin")
    write_vectors(f, inputs["x0"].shape[0], inputs["nh"], inputs["nt"], inputs["ns"], stored, runs)
    ref = read_vectors(f)
    rep = compare(ref, inputs, lambda chi: oracle_first(inputs, chi), lambda chi, conv: oracle_run(inputs, chi, conv), "self-test")
    assert rep["a_at_save"] == "post_update" and rep["phi"]["convention"]["phi_sign"] == -1.0
    assert rep["phi"]["convention"]["tap_shift"] == -2 and abs(rep["phi"]["convention"]["gauge"] - 0.125) < 1e-12
    assert rep["rhs"]["form"] == ("raw", -2)
    # and a genuinely different result is refused
    runs[1]["W"] = runs[1]["W"] * (1 + 1e-8)
    write_vectors(f, inputs["x0"].shape[0], inputs["nh"], inputs["nt"], inputs["ns"], stored, runs)
    with pytest.raises(AssertionError):
        compare(read_vectors(f), inputs, lambda chi: oracle_first(inputs, chi), lambda chi, conv: oracle_run(inputs, chi, conv), "self-test")


@pytest.mark.skipif(not os.path.exists(REF), reason=SKIP)
def test_oracle_against_reference_vectors(inputs):
    rep = compare(read_vectors(REF), inputs, lambda chi: oracle_first(inputs, chi), lambda chi, conv: oracle_run(inputs, chi, conv),
                  "oracle/pic_numpy.py")
    print("reference conventions:", rep)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason=SKIP)
def test_gpu_against_reference_vectors(ctx, inputs):
    """The CUDA path defaults to Conventions(); the conventions detected from the reference vectors are applied through the
    library's own switches (rbm_pic_set_conventions), so a convention difference is reported and then honoured."""
    import rbm_b200 as r

    n, nh = inputs["x0"].shape[0], inputs["nh"]
    hnd = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, inputs["L"]), ctx)

    def first(chi):
        rhs = hnd.deposit(inputs["x0"], inputs["w"])
        hnd.field_update(inputs["x0"], inputs["w"], chi)
        return {"rhs": rhs, "phi": hnd.field_coefficients(), "a": hnd.field_eval(inputs["x0"]), "W0": hnd.field_energy()}

    def run(chi, conv):
        # the library's own switches (rbm_pic_set_conventions): no host-side transform of its outputs
        ns = inputs["ns"]
        o = {k: np.zeros((1, ns, n)) for k in "XVA"}
        o.update(Phi=np.zeros((1, ns, nh)), W=np.zeros((1, ns)), K=np.zeros((1, ns)), M=np.zeros((1, ns)))
        hnd.set_conventions(conv.phi_sign, conv.tap_shift, conv.gauge, conv.a_at_save)
        hnd.set_particles(inputs["x0"], inputs["v0"], inputs["w"])
        hnd.run([chi], inputs["dt"], inputs["nt"], ns, **o)
        hnd.set_conventions()
        return {k: a[0] for k, a in o.items()}
    rep = compare(read_vectors(REF), inputs, first, run, "librbm_b200")
    print("reference conventions:", rep)
    hnd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("conv", [pn.Conventions(), pn.Conventions(-1.0, 0.25, 2, "post_update"), pn.Conventions(1.0, 0.0, -3, "post_update"),
                                  pn.Conventions(-1.0, -1.5, 0, "kick")], ids=repr)
def test_library_convention_switches_match_the_oracle(ctx, inputs, conv):
    """rbm_pic_set_conventions: under every combination of the unpinned conventions the CUDA path produces what the oracle
    produces under the same `Conventions` (Phi, A at saves; X, V, W, K, M unaffected) -- runs without the Julia file."""
    import rbm_b200 as r

    n, nh, ns, nt = inputs["x0"].shape[0], inputs["nh"], inputs["ns"], inputs["nt"]
    hnd = r.PICHandle(n, r.PoissonSolverPBSplines(3, nh, inputs["L"]), ctx)
    hnd.set_conventions(conv.phi_sign, conv.tap_shift, conv.gauge, conv.a_at_save)
    hnd.set_particles(inputs["x0"], inputs["v0"], inputs["w"])
    chis = inputs["chi"]
    o = {k: np.zeros((len(chis), ns, n)) for k in "XVA"}
    o.update(Phi=np.zeros((len(chis), ns, nh)), W=np.zeros((len(chis), ns)), K=np.zeros((len(chis), ns)), M=np.zeros((len(chis), ns)))
    hnd.run(chis, inputs["dt"], nt, ns, **o)
    hnd.field_update(inputs["x0"], inputs["w"], float(chis[1]))
    phi1 = hnd.field_coefficients()
    for p, chi in enumerate(chis):
        g = oracle_run(inputs, float(chi), conv)
        for k in ("W", "K", "M"):
            assert rel(o[k][p], g[k]) < 1e-10, (k, p)
        assert np.max(np.abs(o["X"][p] - g["X"])) < 1e-9 and rel(o["A"][p], g["A"]) < 1e-9, p
        assert np.max(np.abs(o["Phi"][p] - g["Phi"])) < 1e-9 * max(1.0, np.max(np.abs(g["Phi"]))), p
    f = oracle_first(inputs, float(chis[1]))
    assert np.max(np.abs(phi1 - conv.stored_phi(f["phi"]))) < 1e-12 * max(1.0, np.max(np.abs(conv.stored_phi(f["phi"]))))
    # device-resident outputs replay the captured time loop: the switches must invalidate the captured graph
    import torch
    Ad = torch.zeros((len(chis), ns, n), dtype=torch.float64, device="cuda")
    for _ in range(3):
        hnd.run(chis, inputs["dt"], nt, ns, A=Ad)
    hnd.set_conventions()
    A0 = torch.zeros_like(Ad)
    for _ in range(3):
        hnd.run(chis, inputs["dt"], nt, ns, A=A0)
    g0 = oracle_run(inputs, float(chis[0]), pn.Conventions())
    assert rel(A0[0].cpu().numpy(), g0["A"]) < 1e-9 and rel(Ad[0].cpu().numpy(), oracle_run(inputs, float(chis[0]), conv)["A"]) < 1e-9
    hnd.close()
