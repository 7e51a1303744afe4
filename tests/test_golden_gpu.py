"""The CUDA path against the committed golden fixtures directly (tests/golden/ndt_golden.json: 50-digit mpmath evaluation of the
reference's definitions, independent of oracle/ and of the kernels): per-pair residuals and Jacobians of all four functors through
EMIT, and per-pose robustified normal equations through FUSED.  Bar: 1e-8 relative (north_star: 1e-5)."""
import json
import os

import numpy as np
import pytest

from randt_slam_b200 import capi

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ndt_golden.json")
TOL = 1e-8


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_emit_matches_golden_pairs(gpu_ctx, golden):
    assert len(golden["pairs"]) >= 24
    for c in golden["pairs"]:
        cm = np.array(c["cell_m"], np.float32)[None]; cf = np.array(c["cell_f"], np.float32)[None]
        prob = gpu_ctx.problem_create(cm, cf, np.zeros(1, np.uint32), np.zeros(1, np.uint32), [0, 1])
        r, J = prob.eval_emit(np.array(c["params"], np.float64)[None], variant=c["variant"])
        prob.close()
        assert abs(r[0] - c["r"]) <= TOL * c["r"], c["variant"]
        want = np.array(c["J"])
        assert J.shape == (1, len(want))
        assert np.max(np.abs(J[0] - want)) <= TOL * np.abs(want).max(), c["variant"]


def test_fused_matches_golden_normal_equations(gpu_ctx, golden):
    assert len(golden["fused"]) >= 5
    kinds = {"barron": capi.LOSS_BARRON, "welsch": capi.LOSS_WELSCH, "none": capi.LOSS_NONE}
    for c in golden["fused"]:
        cm = np.array(c["cells_m"], np.float32); cf = np.array(c["cells_f"], np.float32)
        n_pairs = len(cm)
        idx = np.arange(n_pairs, dtype=np.uint32)
        prob = gpu_ctx.problem_create(cm, cf, idx, idx, [0, n_pairs])
        loss = capi.make_loss(kinds[c["kind"]], c["a"], c["alpha"], c["mu"], c["weight"])
        out = prob.eval_fused(np.array(c["params"], np.float64)[None], loss, variant=c["variant"])[0]
        prob.close()
        n = len(c["params"])
        H = np.array(c["H"]); g = np.array(c["g"])
        got_H = out[:16].reshape(4, 4)[:n, :n]
        assert np.max(np.abs(got_H - H)) <= TOL * np.abs(H).max(), (c["variant"], c["kind"])
        assert np.max(np.abs(out[16:16 + n] - g)) <= TOL * np.abs(g).max(), (c["variant"], c["kind"])
        assert abs(out[capi.FUSED_COST] - c["cost"]) <= TOL * abs(c["cost"])
        assert abs(out[capi.FUSED_MAXR] - c["max_r"]) <= TOL * c["max_r"]
        assert abs(out[capi.FUSED_SUMSQ] - c["sum_sq"]) <= TOL * c["sum_sq"]
        assert out[capi.FUSED_N] == n_pairs
