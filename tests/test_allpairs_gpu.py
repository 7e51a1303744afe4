"""K8 (randt_eval_allpairs): every moving cell against every fixed cell, same functor / corrector / reduction as K3 FUSED.
Checked against the CPU oracle's fused evaluation over the explicit N_m x N_f pair list (and over the window-filtered list), all four
functor variants; tolerance 1e-9 (the sums run in another order)."""
import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H
from tests.test_k3_gpu import LOSSES, loss_tuple

pytestmark = pytest.mark.gpu


def explicit_pairs(cm, cf, pose, variant, window):
    nm, nf = len(cm), len(cf)
    im = np.repeat(np.arange(nm, dtype=np.uint32), nf); jf = np.tile(np.arange(nf, dtype=np.uint32), nm)
    if window > 0:
        if variant <= 1:
            c, s = pose[0], pose[1]
            if variant == 0:
                n = np.hypot(c, s); c, s = c / n, s / n
            tx, ty = pose[2], pose[3]
        else:
            c, s, tx, ty = np.cos(pose[2]), np.sin(pose[2]), pose[0], pose[1]
        px = c * cm[:, 0].astype(np.float64) - s * cm[:, 1].astype(np.float64) + tx
        py = s * cm[:, 0].astype(np.float64) + c * cm[:, 1].astype(np.float64) + ty
        keep = (np.abs(px[im] - cf[jf, 0]) <= window) & (np.abs(py[im] - cf[jf, 1]) <= window)
        im, jf = im[keep], jf[keep]
    return im, jf


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("window", [0.0, 12.0])
def test_allpairs_matches_oracle(oracle, gpu_ctx, variant, window):
    rng = np.random.default_rng(50 + variant)
    gp = capi.grid_params(P.C3)
    sizes_m, sizes_f = [150, 0, 37], [300, 40, 129]        # ragged batch incl. an empty moving map; 300 > one slab, 150 > one CTA
    cms = [H.random_cells(rng, n, extent=20.0) for n in sizes_m]; cfs = [H.random_cells(rng, n, extent=20.0) for n in sizes_f]
    M = gpu_ctx.map_upload(np.concatenate(cms), np.concatenate([[0], np.cumsum(sizes_m)]).astype(np.uint32), gp)
    F = gpu_ctx.map_upload(np.concatenate(cfs), np.concatenate([[0], np.cumsum(sizes_f)]).astype(np.uint32), gp)
    if variant <= 1:
        poses = np.stack([synth.pose_to_se2(*rng.uniform(-1, 1, 3) * [1, 1, 0.5]) * [1.001, 1.001, 1, 1] for _ in range(3)])
    else:
        poses = rng.uniform(-1, 1, (3, 3)) * [1, 1, 2.0]
    loss = LOSSES[2]
    out = capi.unpack_fused(F.eval_allpairs(M, poses, loss, window=window, variant=variant))
    for b in range(3):
        im, jf = explicit_pairs(cms[b], cfs[b], poses[b], variant, window)
        assert int(out["n"][b]) == len(im)
        if len(im) == 0:
            assert out["cost"][b] == 0.0
            continue
        fo = oracle.fused(variant, cms[b], cfs[b], im, jf, poses[b], loss_tuple(loss), True)
        npar = 4 if variant <= 1 else 3
        assert abs(out["cost"][b] - fo["cost"]) < 1e-9 * abs(fo["cost"])
        assert H.rel_err(out["H"][b][:npar, :npar], fo["H"][:npar, :npar]) < 1e-8 and H.rel_err(out["g"][b][:npar], fo["g"][:npar]) < 1e-8
        assert abs(out["max_r"][b] - fo["max_r"]) < 1e-9 * fo["max_r"]
    # bitwise reproducible
    again = capi.unpack_fused(F.eval_allpairs(M, poses, loss, window=window, variant=variant))
    assert np.array_equal(again["H"], out["H"]) and np.array_equal(again["cost"], out["cost"])
    assert gpu_ctx.take_bad_pairs() == 0
