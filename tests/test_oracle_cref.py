"""The C restatement (oracle/cref.c, the CPU baseline of bench.py) agrees with the NumPy oracle."""
import numpy as np
import pytest

from oracle import cref, estimators as E, machines as M
from oracle.models import lindblad_ising_1d, random_states


@pytest.mark.parametrize("act,fock", [(M.SOFTPLUS, True), (M.LOGCOSH, False)])
def test_cref_matches_numpy_oracle(act, fock):
    N = 5
    hilb, _, _, liouv = lindblad_ising_1d(N, fock=fock)
    net = M.random_machine("ndm", N, 2, act=act, seed=4, std=0.3)
    sr, sc = random_states(hilb, 9, 1), random_states(hilb, 9, 2)
    out, O = net.logpsi_grad(sr, sc)
    cout, cO = cref.ndm_logpsi_grad(net, sr, sc)
    assert np.allclose(cout, out, rtol=1e-13, atol=1e-13) and np.allclose(cO, O, rtol=1e-13, atol=1e-13)
    loc, g = E.local_grad_super(net, liouv, sr, sc)
    cloc, cg = cref.local_grad_super(net, liouv, sr, sc, nthreads=2)
    assert np.allclose(cloc, loc, rtol=1e-12, atol=1e-12) and np.allclose(cg, g, rtol=1e-12, atol=1e-12)
    assert cref.max_threads() >= 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the contract's
    keys, on the GPU arm's metric / unit / workload, and needs no GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("MC samples/s") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "cfg4" in d["config"]["workload"]
