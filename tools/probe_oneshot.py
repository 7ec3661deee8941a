import os, sys, time
sys.path.insert(0, "/root/repo")
from scdiffcom_b200 import synthetic, perm_counts
prob, ai, simple, mask = synthetic.make_problem("config2")
for i in range(5):
    print("---- call", i, file=sys.stderr)
    t0 = time.perf_counter(); r = perm_counts(prob, 10000, seed=1)
    print("python ms", round((time.perf_counter() - t0) * 1e3, 1), file=sys.stderr)
