"""scripts/window_chain_diff.py [n]: the window odometry (hostapi.window_replay) and the oracle's window chain over the first n scans of the replay
drive, both running free: where the two trajectories part (max |pose difference| per scan) and what the solver did there."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from randt_slam_b200 import hostapi, capi, params as P, workloads as W
from oracle import oracle_py as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
p = P.OXFORD
truth, scans = W.make_loop_drive(p, W.REPLAY_SCENE_SEED, n)
q = W.window_odometry_params(hostapi, p)
stamps = 0.2486 * np.arange(n)
poses, states, stats, tot = hostapi.window_replay(capi.grid_params(p), scans, stamps, q)
o_poses, o_states, _, _ = W.oracle_window_replay(O, p, scans, stamps, q)
d = np.max(np.abs(poses - o_poses), axis=1)
print("scan : max|pose diff| (iterations)")
for i in range(0, n, 1):
    if i < 3 or d[i] > 3 * max(d[max(i - 1, 0)], 1e-12) or i % 10 == 0:
        print("%4d : %.3e  (%d it, %d evals)" % (i, d[i], stats[i, 0], stats[i, 1]))
print("max", d.max(), "at", int(d.argmax()))
