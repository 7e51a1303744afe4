import numpy as np, sys
sys.path.insert(0, '.')
from oracle import oracle_py as O
from randt_slam_b200 import capi, hostapi, params as P, synth
from tests import helpers as H
p = P.OXFORD
fixed = H.make_scan(p, 40, (0.0, 0.0, 0.0), 1); moving = H.make_scan(p, 40, (0.6, -0.4, 0.03), 50)
guess = synth.pose_to_se2(0.4, -0.25, 0.02)
gp = capi.grid_params(p)
f = O.voxelize(fixed, *H.vox_args(p)); m = O.voxelize(moving, *H.vox_args(p))
o = O.loop_constraint(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, m["cells"], guess, p.n_results_nn_lookup,
                      matcher_loss_scale=p.loss_function_scale, loop_scale=p.loop_closure_scale, alpha=p.loss_function_convexity,
                      divisor=p.gnc_control_parameter_divisor, max_gnc_steps=p.loop_closure_gnc_steps, on_manifold=False)
print("oracle", o)
with capi.Context(0) as ctx:
    F = ctx.voxelize(fixed, [0, len(fixed)], gp); M = ctx.voxelize(moving, [0, len(moving)], gp)
    prob = ctx.associate(F, M, guess[None], p.n_results_nn_lookup)
    im, jf = O.associate(f["cells"], f["slot"], p.size_x, p.size_y, p.resolution, p.max_neighbor_linf_distance, m["cells"], guess, p.n_results_nn_lookup)
    pm, pf, seg = prob.download()
    print("pairs equal", np.array_equal(pm, im), np.array_equal(pf, jf), len(im))
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)
    opt = capi.solver_options(use_manifold=0, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.loop_closure_gnc_steps, max_num_iterations=p.max_iteration)
    out, res = prob.register_batch(guess[None], loss, opt)
    print("capi  ", out, res)
poses, scores = hostapi.loop_constraints(gp, [fixed], [moving], guess[None], p.n_results_nn_lookup, p.loss_function_scale, p.loss_function_convexity,
                                         p.gnc_control_parameter_divisor, p.loop_closure_gnc_steps, p.loop_closure_scale)
print("host  ", poses, scores)
