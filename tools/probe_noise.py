"""One-shot call repeated: prints the host phase times (SCDC_TIMING=1) of the slowest and the fastest call."""
import os, sys, time, subprocess
sys.path.insert(0, "/root/repo")
os.environ["SCDC_TIMING"] = "1"
from scdiffcom_b200 import synthetic, perm_counts
prob, ai, simple, mask = synthetic.make_problem("config2")
for i in range(3):
    perm_counts(prob, 10000, seed=1)
import io, contextlib, tempfile
calls = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    # stderr of the C library: redirect fd 2 into a file per call
    with tempfile.TemporaryFile(mode="w+b") as tf:
        saved = os.dup(2); os.dup2(tf.fileno(), 2)
        t0 = time.perf_counter(); r = perm_counts(prob, 10000, seed=1); dt = (time.perf_counter() - t0) * 1e3
        os.dup2(saved, 2); os.close(saved)
        tf.seek(0); log = tf.read().decode()
    calls.append((dt, log))
calls_sorted = sorted(calls, key=lambda c: c[0])
print("call ms:", [round(c[0], 1) for c in calls])
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max unavailable", e)
print("---- fastest", round(calls_sorted[0][0], 1)); print(calls_sorted[0][1])
print("---- slowest", round(calls_sorted[-1][0], 1)); print(calls_sorted[-1][1])
