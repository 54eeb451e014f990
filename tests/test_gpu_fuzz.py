"""Randomised configurations under `-m gpu` (reduced, fixed seeds; the long runs are tests/tools/fuzz*.py): horizons 1..19,
gait mixes, segment counts, spreads, batch sizes from 1 up — the CUDA path on both call paths against qpOASES, and the
device-side caller against its fp32 oracle bit for bit."""
import os
import sys

import pytest
import torch

from oracle import cmpc_oracle as O

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))


@pytest.mark.skipif(not O.available(), reason="oracle/_ref did not travel")
def test_randomised_configurations_against_qpoases(built_lib):
    import fuzz
    total, bad, worst = fuzz.run(rounds=30, seed=20261019, verbose=False)
    assert total > 2000 and bad == 0, (total, bad, worst)
    assert worst <= 1e-6     # measured: ~1e-8 N


def test_randomised_commands_against_the_frontend_oracle(built_lib):
    import fuzz_commands
    total, bad = fuzz_commands.run(rounds=16, seed=43, verbose=False)
    assert total > 500 and bad == 0, (total, bad)
