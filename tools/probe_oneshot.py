import os, sys, time
sys.path.insert(0, "/root/repo")
from scdiffcom_b200 import synthetic, perm_counts
name = sys.argv[1] if len(sys.argv) > 1 else "config2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
prob, ai, simple, mask = synthetic.make_problem(name)
for i in range(4):
    print("---- call", i, file=sys.stderr)
    t0 = time.perf_counter(); r = perm_counts(prob, iters, seed=1)
    print("python ms", round((time.perf_counter() - t0) * 1e3, 1), {k: round(v, 1) for k, v in r["stats"].items() if k.startswith("ms_")}, file=sys.stderr)
