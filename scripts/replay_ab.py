"""A/B of the full-length replay: randt_associate's fused single-map path against the general path (RANDT_NO_FUSED_ASSOC=1).
Prints the trajectory error against the ground truth of this run and saves the poses so two runs can be compared step by step."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from randt_slam_b200 import capi, params as P, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8609
truth, scans = W.make_loop_drive(P.OXFORD, W.REPLAY_SCENE_SEED, n)
with capi.Context(0) as ctx:
    poses, dt, its = W.device_replay(ctx, capi, P.OXFORD, scans)
err = np.hypot(poses[:, 2] - truth[:, 0], poses[:, 3] - truth[:, 1])
tag = "general" if os.environ.get("RANDT_NO_FUSED_ASSOC") else "fused"
np.save("gpurun_out/replay_%s.npy" % tag, poses)
print("%s: %.3f ms/scan, %.2f it/scan, max err %.3f m, final err %.3f m" % (tag, dt * 1e3 / (n - 1), its, err.max(), err[-1]))
other = "gpurun_out/replay_%s.npy" % ("fused" if tag == "general" else "general")
if os.path.exists(other):
    q = np.load(other)
    d = np.abs(q - poses).max(axis=1)
    nz = np.nonzero(d)[0]
    print("first differing step:", (int(nz[0]), float(d[nz[0]])) if len(nz) else None, "max diff", float(d.max()))
