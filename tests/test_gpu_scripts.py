"""GPU: the three stage drivers of the reference workflow, end to end through files, at a small size
(scripts/bump_on_tail_1_training.jl -> 2_projections.jl -> 3_testing.jl; Python renderings under scripts/).
Stage 3 is driven from the projections file alone, as in the reference (scripts/bump_on_tail_3_testing.jl:30-42)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_three_stage_workflow(tmp_path):
    run = str(tmp_path / "BoT_small.npz")
    s1 = _run("bump_on_tail_1_training.py", "--np", "4000", "--nh", "16", "--nparam", "5", "--out", run)
    nt = int(25 // 1e-1)      # `Int(div(T, dt))` of the reference script (:26-28): 249 in floating point, in Julia as in Python
    assert f"integrated 5 samples x {nt} steps x 4000 particles" in s1 and os.path.exists(run)
    s2 = _run("bump_on_tail_2_projections.py", "--run", run)
    kp, ke = int(re.search(r"k_p = (\d+)", s2).group(1)), int(re.search(r"k_e = (\d+)", s2).group(1))
    assert 1 <= kp < 4000 and 1 <= ke < 4000
    proj = run.replace(".npz", "_projections.npz")
    z = np.load(proj)
    for key in ("kp", "ke", "Lp", "Pp", "Le", "Pe", "Pie", "parameters/kappa", "parameterspace/samples/chi",
                "initial_conditions/list", "integrator/dt", "p", "L"):          # the tree of reducedbasis.jl:68-88
        assert key in z.files, key
    vf = np.load(run.replace(".npz", "_vectorfield.npz"))
    assert vf["X"].shape == (kp, 5 * (nt + 1))
    os.remove(run)                                                               # stage 3 must not need the training set
    s3 = _run("bump_on_tail_3_testing.py", "--run", run, "--ntest", "4")
    assert f"k_p = {kp}, k_e = {ke}" in s3
    ex = float(re.search(r"rel. error X ([0-9.e+-]+)", s3).group(1))
    assert np.isfinite(ex) and ex < 0.5          # the reduced model tracks the full one (accuracy is the method's, not tested here)
