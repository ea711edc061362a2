"""The record cost-expansion kernels (rollout.cu): the blocked kernel k_expansion_rec16b (default) must write the same records, bit for bit,
as the first one (k_expansion_rec16, TO_CEXP_V1=1) -- the kernel choice is read once per process, so the two runs are subprocesses of
profiles/scripts/cexp_ab.py (goal + control bounds at N = 101 / 33 / 16, and state + control bounds: up to three AL terms per entry)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "profiles", "scripts", "cexp_ab.py")


def _dump(path, **env):
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, SCRIPT, str(path)], check=True, env=e, timeout=600, stdout=subprocess.DEVNULL)
    return np.load(path)


def test_blocked_cost_expansion_is_bit_identical(tmp_path):
    a = _dump(tmp_path / "v1.npz", TO_CEXP_V1="1")
    b = _dump(tmp_path / "v2.npz", TO_CEXP_V1="0")
    assert sorted(a.files) == sorted(b.files) and len(a.files) >= 24
    for k in a.files:
        assert np.all(np.isfinite(a[k])), k
        assert np.array_equal(a[k], b[k]), f"{k}: max |v1 - v2| = {np.max(np.abs(a[k] - b[k])):.3e}"
