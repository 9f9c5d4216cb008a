"""Host-side logic of the hot path against the UNMODIFIED reference, on CPU.

The geometry helpers, the gaze-history weighting (the reference's O(T^2) Python loops vs one
batched matrix product) and the validity-masked sequence losses (loop over the batch vs one
expression) are pure torch arithmetic on both sides, so they can be compared directly wherever
/root/reference is mounted -- this container, not the GPU box.  oracle/check_host_logic.py does
the comparison in a subprocess (the reference's packages are called `models`, `core`, `losses`
and want `cwd = src/`; keeping them out of this process keeps the other tests' imports clean)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/src/models/common.py'


@pytest.mark.skipif(not os.path.exists(REF), reason='the reference tree is not mounted here')
def test_geometry_history_and_losses_match_the_reference_functions():
    p = subprocess.run([sys.executable, os.path.join(REPO, 'oracle', 'check_host_logic.py')],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    errs = json.loads(p.stdout.strip().splitlines()[-1])
    assert len(errs) >= 20
    # same formulas, reordered sums at most: a few ulps of fp32 on O(1) quantities
    worst = {k: v for k, v in errs.items() if not v <= 2e-6}
    assert not worst, worst
