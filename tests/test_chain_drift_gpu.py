"""BASELINE cfg #3 (full DDPM sampling chain): drift of our chain against the fp64 oracle chain under identical injected
noise, with the fp32 PyTorch-eager chain as the yardstick (tests/diagnostics/chain_drift.py).  The chain of a
random-init denoiser amplifies differences over its last steps (c1 -> 1 as t -> 0), for ANY fp32 implementation, so
the bound is relative to the eager chain's own drift; profiles/r02b/cfg3_drift_*.json hold the 32^3 / 64^3 records."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sampling_chain_tracks_the_fp64_chain():
    r = subprocess.run([sys.executable, os.path.join(HERE, "diagnostics", "chain_drift.py"), "--resol", "16", "--steps", "80",
                        "--every", "10", "--f64", "--with-eager32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("final_rel_err", "final_eager32_rel_err", "max_rel_err")}))
    mid = [t for t in d["trace"] if t["t"] > 150]
    assert all(t["rel_err"] < 1e-5 for t in mid), mid                       # the body of the chain: 1e-6-level tracking
    assert d["final_rel_err"] < 3 * d["final_eager32_rel_err"] + 1e-4       # the last steps: no worse than eager fp32 x3
