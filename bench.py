#!/usr/bin/env python
"""bench.py — permutations/second of scDiffCom's permutation-test hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Default workload = BASELINE.json configs[4], the north-star job: 250 k cells x 60 cell types, differential mode,
100 k permutations over 8 GPUs = 12 500 permutations per GPU (fits one GPU: 3.2 GB resident). A "step" is one whole
permutation job of the workload on one GPU: `iterations` replicates (label shuffles on the device, group means, scores,
exceedance counts) followed — when N > 1 — by the NCCL all-reduce of the integer counters. Weak scaling: every rank runs
its own replicate range [rank * iterations, (rank + 1) * iterations) of the same problem (replicated expression matrix),
so the job at N GPUs is N * iterations replicates (N = 8: exactly the 100 k-permutation job).

  value : replicates/s with the problem already resident in HBM (scdc_create done), counters left on the device.
  e2e   : the same job through the one-shot C-ABI call `scdc_perm_counts` with HOST buffers: problem validation,
          H2D of the matrix, device work, D2H of the counters — what `.Call` from R pays. At N > 1 rank 0 makes ONE
          in-library call with n_gpus = N (host thread per device, the matrix copied to the other devices over NVLink, NCCL all-reduce of the
          counters) while the other ranks wait at a CPU-side barrier: exactly the R caller's multi-GPU path.
  roofline : the group-sum kernels (k_scatter, + k_pass1_cond in differential mode), algorithmic bytes /
          their CUDA-event time.
  cpu_baseline : the CPU oracle port (oracle/scdc_oracle.c, OpenMP over replicates, all host cores) on a bounded sample.
          R is not installed in this image, so the literal reference (future multisession) cannot run; the port omits
          R's data.table join overhead, i.e. it is faster than the reference would be.
  other_workloads (N = 1 only): short runs of BASELINE.json configs[1..3] (config2 detection, config3 differential,
          config4 with the real LRI_human table) with the same keys, and `fp32_fast`: the headline workload in the
          SCDC_FP32_FAST precision mode (narrower than the reference's arithmetic: reported, never the headline).

Input data: synthetic (scdiffcom_b200/synthetic.py). config4 reads the reference's LRI_human table
from the decoded fixture tests/golden/LRI_human_curated.npz (input data, outside every timed region). Built problems
are cached under $SCDC_BENCH_CACHE (default /tmp/scdc_bench_cache) so that the two arms and the ranks of one box
build them once.

`--impl reference` prints the CPU arm's JSON line (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "permutations/sec"
UNIT = "permutations/s"
# replicates per GPU per step (weak scaling). config5: 100 000 / 8.
BENCH_ITERS = {"config2": 10_000, "config3": 10_000, "config4": 12_500, "config5": 12_500, "config5s": 4096}
_ARRAYS = ("colptr", "rowidx", "values", "ct_code", "cond_code", "sub_gene", "row_lri", "row_emit", "row_recv", "observed")
_SCALARS = ("n_cells", "n_genes", "n_celltypes", "n_conditions", "max_nL", "max_nR")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, n in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name):
    """(PermutationProblem, synthetic config) of a BASELINE.json workload; cached on disk per box."""
    from scdiffcom_b200 import PermutationProblem, synthetic
    cfg = dict(synthetic.CONFIGS[name])
    cache_dir = os.environ.get("SCDC_BENCH_CACHE", "/tmp/scdc_bench_cache")
    path = os.path.join(cache_dir, f"{name}_seed{cfg['seed']}_v2.npz")
    if os.path.exists(path):
        try:
            d = np.load(path)
            kw = {k: d[k] for k in _ARRAYS if k in d.files}
            kw.setdefault("cond_code", None)
            kw.update({k: int(d[k]) for k in _SCALARS})
            return PermutationProblem(**kw), cfg
        except Exception:   # truncated / foreign file: rebuild
            pass
    lri = None
    if name == "config4":   # BASELINE.json config 4 uses the real LRI_human table (decoded fixture, tests/golden/)
        lri = dict(np.load(os.path.join(ROOT, "tests", "golden", "LRI_human_curated.npz")))
    prob, ai, simple, mask = synthetic.make_problem(name, lri_table=lri)
    try:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp.npz"
        arrays = {k: np.asarray(getattr(prob, k)) for k in _ARRAYS if getattr(prob, k) is not None}
        arrays.update({k: np.int64(getattr(prob, k)) for k in _SCALARS})
        np.savez(tmp, **arrays)
        os.replace(tmp, path)
    except OSError:
        pass
    return prob, cfg


def workload_config(name, prob, iters):
    """The `config` object: identical in both arms (everything arm-specific lives in sibling keys)."""
    from scdiffcom_b200 import synthetic
    cfg = synthetic.CONFIGS[name]
    return {"workload": f"{name}: synthetic {prob.n_cells} cells x {prob.n_genes} LRI genes (of {cfg['n_genes']}), "
                        f"{prob.n_celltypes} cell types, {'differential' if prob.n_conditions == 2 else 'single-condition'} mode, "
                        f"{int(iters)} permutations per GPU",
            "nnz": prob.nnz, "tested_rows": prob.n_rows, "counters_per_row": prob.ncols, "n_lri": prob.n_lri,
            "iterations_per_gpu": int(iters), "seed": cfg["seed"]}


def cpu_port_rate(prob, target_s=12.0, threads=0):
    """CPU oracle port on a bounded sample: returns (replicates/s, cores, sample description)."""
    from oracle import oracle_c
    cores = threads or os.cpu_count() or 1
    n = max(cores, 8)
    cond, ct = oracle_c.philox_labels(prob, n, seed=1)
    t0 = time.perf_counter()
    oracle_c.perm_counts(prob, ct, cond, nthreads=cores)
    dt = time.perf_counter() - t0
    n2 = int(max(n, min(20000, (target_s / max(dt, 1e-6)) * n)))
    n2 = (n2 // cores) * cores or cores
    if n2 > n:
        cond, ct = oracle_c.philox_labels(prob, n2, seed=2)
        t0 = time.perf_counter()
        oracle_c.perm_counts(prob, ct, cond, nthreads=cores)
        dt = time.perf_counter() - t0
        n = n2
    return n / dt, cores, f"{n} replicates of the same problem, labels pre-drawn (not timed), {dt:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, cfg = build_workload(args.workload)
    iters = int(args.iterations or BENCH_ITERS.get(args.workload, cfg["iterations"]))
    from oracle import oracle_c
    cores = os.cpu_count() or 1
    # each step = a bounded sample; size it from a pilot so that steps + warmup fit in ~2 minutes
    rate, _, _ = cpu_port_rate(prob, target_s=2.0)
    per_step = max(cores, int(rate * min(20.0, 100.0 / max(1, args.steps + args.warmup))))
    per_step = min(per_step, iters)
    per_step = max(cores, (per_step // cores) * cores)
    cond, ct = oracle_c.philox_labels(prob, per_step, seed=3)
    for _ in range(args.warmup):
        oracle_c.perm_counts(prob, ct, cond, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_c.perm_counts(prob, ct, cond, nthreads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} replicates per step (bounded sample of the {iters}-replicate job), labels pre-drawn"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, prob, iters),
        "note": "CPU oracle port of the R path (R not installed in this image); OpenMP over replicates",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class Runner:
    """One process per GPU: distributed plumbing + the timed legs of one workload."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from scdiffcom_b200 import device_count
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available() or device_count() < 1:
            raise SystemExit("bench.py needs a B200: no CUDA device visible (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.cpu_group = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.cpu_group = dist.new_group(backend="gloo")   # CPU-side barrier: leaves the GPUs idle while rank 0 works
        self.stream = torch.cuda.Stream()   # a real (non-default) stream: the engine enqueues on it, the events time it
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, name, iters, steps, warmup, e2e_steps, e2e_warmup, precision="fp64", headline=True):
        torch, dist, args = self.torch, self.dist, self.args
        from scdiffcom_b200 import Engine, perm_counts, sharding
        world, rank, local, stream = self.world, self.rank, self.local, self.stream
        prob, cfg = build_workload(name)
        seed = cfg["seed"]
        ncount = prob.n_rows * prob.ncols
        counts_dev = torch.zeros(max(1, ncount), dtype=torch.int64, device="cuda")
        eng = Engine(prob, device=local)
        stats_acc = []
        kw = {} if precision == "fp64" else {"precision": precision}

        def step(collect):
            # weak scaling: the job at `world` GPUs is world * iters replicates; rank r runs [r * iters, (r + 1) * iters)
            st = sharding.run_sharded(eng, world * iters, counts_dev, seed=seed, batch_size=args.batch,
                                      stream=stream.cuda_stream, **kw)
            if collect:
                stats_acc.append(st)

        for _ in range(warmup):
            step(False)
        self.barrier()
        sampler = ClockSampler(local)
        if rank == 0 and headline:
            sampler.start()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record(stream)
        for k in range(steps):
            step(True)
            evs[k + 1].record(stream)
        self.barrier()
        ms = evs[0].elapsed_time(evs[-1])
        step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]
        clocks = sampler.stop() if (rank == 0 and headline) else None
        ms_max = self.max_over_ranks(ms)
        value = world * iters * steps / (ms_max / 1e3)
        checksum = int(counts_dev.sum().item())
        eng.close()
        del counts_dev

        # ---- e2e: one-shot C-ABI call with host buffers (H2D + D2H inside the timed region) ----
        # N > 1: rank 0 alone makes the in-library multi-GPU call; the other ranks sit at a CPU (gloo) barrier.
        e2e_call_ms, e2e_stats = [], None
        torch.cuda.synchronize()
        if self.cpu_group is not None:
            dist.barrier(group=self.cpu_group)
        if rank == 0 and args.no_e2e:
            e2e_s, e2e_stats, e2e_checksum, e2e_steps = float("inf"), {"ms_upload": 0, "ms_device": 0, "ms_merge": 0, "n_gpus": 0}, None, 0
        elif rank == 0:
            for _ in range(e2e_warmup):   # the first cudaMalloc / cudaFree cycles of new sizes are slow
                perm_counts(prob, world * iters, seed=seed, n_gpus=world, batch_size=args.batch, as_float=False, **kw)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                t_call = time.perf_counter()
                # as_float=False: the raw uint64 counters the C-ABI call returns (the .Call wrapper's view); the NumPy
                # float64 / NaN copies the Python mirror makes for its own tables are not part of the call
                r = perm_counts(prob, world * iters, seed=seed, n_gpus=world, batch_size=args.batch, as_float=False, **kw)
                e2e_call_ms.append(round((time.perf_counter() - t_call) * 1e3, 2))
            e2e_s = time.perf_counter() - t0
            e2e_stats = r["stats"]
            raw = r["counts_raw"]
            e2e_checksum = int(raw[raw != np.uint64(2**64 - 1)].sum())
        if self.cpu_group is not None:
            dist.barrier(group=self.cpu_group)
        if rank != 0:
            return None
        e2e_value = world * iters * e2e_steps / e2e_s
        peak, peak_src = load_peaks()
        ms_reduce = sum(s["ms_reduce"] for s in stats_acc)
        bytes_reduce = sum(s["reduce_bytes"] for s in stats_acc)
        # one k_scatter launch per batch; in differential mode k_pass1_cond runs next to it (persistent grid on a second stream),
        # so the unit is "one batch": the scatter's launch with pass 1 beside it, and the algorithmic bytes of both
        launches = sum(s["reduce_launches"] for s in stats_acc) // prob.n_conditions
        achieved = bytes_reduce / (ms_reduce / 1e3) / 1e9 if ms_reduce > 0 else 0.0
        wavefronts = sum(s["reduce_wavefronts"] for s in stats_acc)
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        pipe_peak = 148 * sm_clock * 1e6          # one shared-memory wavefront per clock per SM
        pipe = {"bound": "SM load-store data pipe (1 wavefront / clk / SM: shared-memory accesses and global-load write-backs)", "unit": "wavefronts/s",
                "achieved": wavefronts / (ms_reduce / 1e3) if ms_reduce > 0 else 0.0, "peak": pipe_peak,
                "frac": (wavefronts / (ms_reduce / 1e3) / pipe_peak) if ms_reduce > 0 else 0.0,
                "wavefronts_per_launch": wavefronts / max(1, launches), "sm_mhz": sm_clock,
                "note": "design wavefronts of k_scatter (entry broadcast + label load + accumulator read-modify-write) over the "
                        "CUDA-event time of the whole reduction (k_pass1_cond runs next to it in differential mode); ncu "
                        "l1tex__throughput of k_scatter alone is in profiles/NOTES_r2.md (79.5 % at K = 120, 99.2 % at K = 50)"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(name if precision == "fp64" else f"{name}_{precision}")
        roofline = {
            "bound": "hbm", "kernel": "k_scatter (k_pass1_cond runs beside it in differential mode; bytes and time of both)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "bytes_per_launch": bytes_reduce / max(1, launches), "ms_per_launch": ms_reduce / max(1, launches),
            "launches": launches, "binding_unit": pipe,
            "note": "lane = replicate batching makes the kernel load-store-pipe bound, not HBM bound (DESIGN.md section 3); an ncu "
                    "launch list serialises the two kernels, so k_pass1_cond shows its stand-alone time there (small persistent grid) "
                    "while live it overlaps k_scatter: compare ms_per_launch with k_scatter's ncu time (profiles/NOTES_r2.md)",
            "unbatched_equivalent_gbs": (12.0 * prob.nnz * prob.n_conditions) * (iters * steps) / (ms_reduce / 1e3) / 1e9 if ms_reduce > 0 else None,
        }
        shares = {k: sum(s[k] for s in stats_acc) for k in ("ms_labels", "ms_reduce", "ms_score", "ms_device")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if precision == "fp64" else "f32 sums / f64 scores (SCDC_FP32_FAST)", "data": "synthetic",
            "config": workload_config(name, prob, iters),
            "batch": stats_acc[0]["batch_size"],
            "l2": "working set per step (matrix + labels + group means, GBs) exceeds the 126 MB L2; no flush",
            "parallelism": f"replicates sharded over {world} GPU(s), expression matrix replicated, one NCCL all-reduce of int64 counters per step",
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_stats.get("h2d_bytes", 0)),
                    "d2h_bytes_per_step": int(e2e_stats.get("d2h_bytes", 0)), "steps": e2e_steps, "call_ms": e2e_call_ms,
                    "ms_upload": e2e_stats["ms_upload"], "ms_device": e2e_stats["ms_device"], "ms_merge": e2e_stats["ms_merge"],
                    "n_gpus_in_library": e2e_stats["n_gpus"], "counts_checksum": e2e_checksum,
                    "api": "scdc_perm_counts(host buffers)" + (f", ONE in-library call with n_gpus = {world} from rank 0" if world > 1 else "")},
            "gpu_launches": int(sum(s["kernel_launches"] for s in stats_acc)),
            "roofline": roofline, "clocks": clocks,
            "device_ms": shares, "step_ms": step_ms, "counts_checksum": checksum,
        }
        return line, prob


def run_ours(args):
    R = Runner(args)
    name = args.workload
    from scdiffcom_b200 import synthetic
    iters = int(args.iterations or BENCH_ITERS.get(name, synthetic.CONFIGS[name]["iterations"]))
    out = R.run(name, iters, args.steps, args.warmup, max(1, args.steps), args.e2e_warmup or max(2, min(3, args.warmup)),
                precision=args.precision)
    line = None
    if out is not None:
        line, prob = out
        cpu = None
        if R.world == 1 and not args.no_cpu:
            cpu_rate, cores, sample = cpu_port_rate(prob, target_s=args.cpu_seconds)
            cpu = {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line["cpu_baseline"] = cpu
    if R.world == 1 and not args.no_extra:
        extra = {}
        from scdiffcom_b200 import _lib
        if _lib.HAS_FP32_FAST:
            fl, _ = R.run(name, iters, 2, 1, 2, 1, precision="fp32_fast", headline=False)
            extra["fp32_fast"] = fl
        for other in ("config2", "config3", "config4"):
            if other == name:
                continue
            ol, _ = R.run(other, BENCH_ITERS[other], 2, 1, 2, 1, headline=False)
            extra[other] = ol
        line["other_workloads"] = extra
    if line is not None:
        print(json.dumps(line))
    if R.world > 1:
        R.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5")
    ap.add_argument("--iterations", type=int, default=0, help="override replicates per GPU per step")
    ap.add_argument("--batch", type=int, default=0, help="replicates per device launch (0 = auto)")
    ap.add_argument("--e2e-warmup", type=int, default=0, help="warm-up calls of the end-to-end leg (0 = 2..3)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg (e2e.value = 0)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other workloads / fp32-fast mode")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32_fast"],
                    help="profiling runs: time the workload in SCDC_FP32_FAST (the default run reports it under other_workloads)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
