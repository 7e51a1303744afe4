"""Where the time of one randt_scan_step goes: runs a short replay with RANDT_DEBUG_TIMING=1 in a child process and averages the
host-side laps the library prints (wall clock between the sub-calls of the composite; each ends with its own synchronisation)."""
import collections, os, re, subprocess, sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
child = ("import numpy as np\nfrom randt_slam_b200 import capi, params as P, workloads as W\n"
         "truth, scans = W.make_loop_drive(P.OXFORD, W.REPLAY_SCENE_SEED, %d)\n"
         "ctx = capi.Context(0)\nposes, dt, its = W.device_replay(ctx, capi, P.OXFORD, scans)\n"
         "print('ms/scan %%.4f  it/scan %%.2f' %% (dt * 1e3 / (len(scans) - 1), its))\n" % n)
env = dict(os.environ, RANDT_DEBUG_TIMING="1", PYTHONPATH=".")
r = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True)
print(r.stdout.strip())
laps = collections.defaultdict(list)
for line in r.stderr.splitlines():
    m = re.match(r"\[randt\] (\w+) (.+?)\s+([0-9.]+) us", line)
    if m:
        laps[(m.group(1), m.group(2).strip())].append(float(m.group(3)))
for (grp, what), v in laps.items():
    v = v[len(v) // 4:]                      # steady state: drop the first quarter (pool growth, submap still small)
    print("%-16s %-20s n=%5d  mean %8.1f us  median %8.1f us" % (grp, what, len(v), sum(v) / len(v), sorted(v)[len(v) // 2]))
