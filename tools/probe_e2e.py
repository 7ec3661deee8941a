import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from scdiffcom_b200 import synthetic, perm_counts, Engine
prob, ai, simple, mask = synthetic.make_problem("config2")
def run(tag):
    for i in range(3):
        t0 = time.perf_counter()
        r = perm_counts(prob, 10000, seed=1)
        print(tag, "python total", round((time.perf_counter() - t0) * 1e3, 1), "ms; C total", round(r["stats"]["ms_total"], 1), file=sys.stderr)
run("plain")
torch.cuda.set_device(0)
x = torch.zeros(10, device="cuda")
run("after torch cuda init")
eng = Engine(prob, device=0)
eng.run(10000, seed=1)
run("with resident engine")
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
run("with torch stream")
import subprocess
sys.path.insert(0, "/root/repo")
from bench import ClockSampler
sm = ClockSampler(0); sm.start(); time.sleep(1.0); print(sm.stop(), file=sys.stderr)
run("after sampler stop")
p = subprocess.run(["pgrep", "-a", "nvidia-smi"], capture_output=True, text=True); print("pgrep:", p.stdout, file=sys.stderr)
sm = ClockSampler(0); sm.start(); time.sleep(0.3)
run("WHILE sampler runs")
print(sm.stop(), file=sys.stderr)
