import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from scdiffcom_b200 import synthetic, perm_counts
prob, ai, simple, mask = synthetic.make_problem("config2")
for _ in range(4):
    perm_counts(prob, 10000, seed=1)
for label, env in (("no early reduce", {"SCDC_NO_EARLY_REDUCE": "1"}), ("2 chunks", {"SCDC_CHUNK_NNZ": "5000000"}),
                   ("4 chunks", {"SCDC_CHUNK_NNZ": "2600000"}), ("8 chunks", {"SCDC_CHUNK_NNZ": "1300000"}),
                   ("6 chunks", {"SCDC_CHUNK_NNZ": "1700000"}), ("3 chunks", {"SCDC_CHUNK_NNZ": "3400000"}), ("16 chunks", {"SCDC_CHUNK_NNZ": "640000"})):
    for k in ("SCDC_NO_EARLY_REDUCE", "SCDC_CHUNK_NNZ"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(8):
        t0 = time.perf_counter(); r = perm_counts(prob, 10000, seed=1); ts.append((time.perf_counter() - t0) * 1e3)
    print(label, "python ms", [round(t, 1) for t in ts], "upload", round(r["stats"]["ms_upload"], 1), "reduce", round(r["stats"]["ms_reduce"], 1), "device", round(r["stats"]["ms_device"], 1), file=sys.stderr)
