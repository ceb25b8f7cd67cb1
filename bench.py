#!/usr/bin/env python
"""bench.py — TrueKNN hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle port, all host threads)

Metric (BASELINE.json): kNN queries/s, k = 10, 3-D.  A step = one full search (start-radius
estimate + every radius-doubling round) over the prebuilt LBVH, inputs resident in HBM; BVH build ms is
reported beside it (the reference's own split: "Build time" vs "True KNN time",
samples/s01-trueknn/hostCode.cpp:201-212 vs :279-347).  `e2e` is the same metric through the C ABI with
pinned HOST buffers: H2D of the points + build + search + D2H of the results inside the timed region.

Workloads:
  cfg2  (default) N = 1: 10 M uniform points (seed 42), every point is a query.
        N > 1: weak scaling of the query-sharded variant — the cloud grows to 10 M x N points, the BVH is
        replicated on every GPU, GPU g answers the g-th contiguous Morton slice (10 M queries each);
        no collective on the data path.
  cfg4  100 M uniform points at every N, queries sharded (strong scaling; BASELINE.json configs[3]).
  cfg1 / cfg3  parity-test shapes, runnable for completeness (100 K k=5; 10 M LiDAR-like k=64).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "cfg1": dict(n=100_000, k=5, cloud="uniform", seed=42, scaling="weak",
                 name="cfg1: TrueKNN k=5 on 100K uniform 3-D points (seed 42), all points are queries"),
    "cfg2": dict(n=10_000_000, k=10, cloud="uniform", seed=42, scaling="weak",
                 name="cfg2: TrueKNN k=10 on 10M uniform 3-D points (seed 42), all points are queries"),
    "cfg3": dict(n=10_000_000, k=64, cloud="lidar", seed=7, scaling="weak",
                 name="cfg3: TrueKNN k=64 on 10M clustered LiDAR-like 3-D points (seed 7)"),
    "cfg4": dict(n=100_000_000, k=10, cloud="uniform", seed=42, scaling="strong",
                 name="cfg4: query-sharded kNN k=10 on 100M uniform 3-D points, BVH replicated per GPU"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        in_timed = 0
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            # samples from the first warm-up step to the end of the timed region count as "under load"
            if self.t_begin and ts < self.t_begin - 0.05:
                continue
            if self.t_end and ts > self.t_end + 0.15:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            in_timed += 1
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (the reference has no CPU implementation of its own
# and cannot be compiled here — OptiX — so `kind` is "port").
# --------------------------------------------------------------------------------------------------
def cpu_arm(cfg: dict, n_points: int, steps: int, warmup: int, sample_queries: int, label: str):
    from oracle import oracle as O
    from owlraytracing_b200 import datasets

    O.build()
    x = datasets.uniform(n_points, cfg["seed"]) if cfg["cloud"] == "uniform" else datasets.lidar_like(n_points, cfg["seed"])
    t0 = time.perf_counter()
    tree = O.KdTree(x, leaf=8)
    build_s = time.perf_counter() - t0
    m = min(sample_queries, n_points)
    rng = np.random.default_rng(0)
    times = []
    for it in range(warmup + steps):
        ids = np.sort(rng.choice(n_points, m, replace=False)).astype(np.int32)
        q = x[ids]
        t0 = time.perf_counter()
        tree.query(q, cfg["k"], self_ids=ids)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    tree.close()
    per_step = float(np.mean(times)) if times else float("nan")
    return {
        "value": m / per_step, "unit": "queries/s", "cores": O.num_threads(), "kind": "port",
        "sample": f"{label}: {m} sampled queries per step against the full {n_points}-point exact kd-tree "
                  f"(oracle/knn_oracle.c, OpenMP, tree build {build_s:.2f} s excluded like the GPU build)",
        "build_s": build_s, "ms_per_step": per_step * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    n_points = cfg["n"] if (args.workload != "cfg2" or args.gpus == 1) else cfg["n"] * args.gpus
    if args.points:
        n_points = args.points
    res = cpu_arm(cfg, n_points, args.steps, args.warmup, args.cpu_sample, "reference arm")
    line = {
        "impl": "reference", "metric": "knn_queries_per_s", "value": res["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "n_points": n_points, "k": cfg["k"], "queries_per_step": min(args.cpu_sample, n_points)},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from owlraytracing_b200 import TrueKNN, datasets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = WORKLOADS[args.workload]
    k = cfg["k"]
    weak_grow = args.workload == "cfg2" and world > 1
    n_points = cfg["n"] * world if weak_grow else cfg["n"]
    if args.points:
        n_points = args.points
    scaling = cfg["scaling"] if args.workload != "cfg2" else "weak"

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()  # nvidia-smi needs ~1 s to start streaming; it runs through warm-up and the timed region
    t = TrueKNN(local_rank)
    stream = torch.cuda.current_stream()
    t.set_stream(stream.cuda_stream)

    # ---- synthetic input, generated on the device (uniform) or on the host (LiDAR-like) ----
    if cfg["cloud"] == "uniform":
        xd = torch.empty((n_points, 3), dtype=torch.float32, device=dev)
        t.generate_uniform(cfg["seed"], 0, n_points, out=xd)
    else:
        xd = torch.from_numpy(datasets.lidar_like(n_points, cfg["seed"])).to(dev)
    torch.cuda.synchronize()

    # ---- build (timed separately; replicated on every GPU) ----
    build_ms = []
    for _ in range(3):
        t.build(xd)
        build_ms.append(t.stats()["build_ms"])
    bstats = t.stats()
    cap = t.shard_capacity(world)
    qid = torch.empty((cap,), dtype=torch.int32, device=dev)
    idx = torch.empty((cap, k), dtype=torch.int32, device=dev)
    dst = torch.empty((cap, k), dtype=torch.float32, device=dev)

    def step():
        return t.search_shard(k, rank, world, start_radius=args.start_radius, out=(qid, idx, dst))

    if sampler:
        sampler.mark_begin()
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    kernel_ms_sum, kernel_launches, rounds_seen = 0.0, 0, []
    ev0.record(stream)
    for _ in range(args.steps):
        q_r, _, _ = step()
        s = t.stats()
        launches += s["kernel_launches"]
        kernel_ms_sum += sum(s["kernel_ms"])
        kernel_launches += len(s["kernel_ms"])
        rounds_seen.append(s["rounds"])
    ev1.record(stream)
    barrier()
    if sampler:
        sampler.mark_end()
    elapsed_ms = ev0.elapsed_time(ev1)
    my_queries = int(q_r.shape[0])
    sstats = t.stats()

    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    qtot = torch.tensor([my_queries], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(qtot, op=dist.ReduceOp.SUM)
    elapsed_ms = float(tmax.item())
    total_queries = int(qtot.item())
    ms_per_step = elapsed_ms / args.steps
    value = total_queries / (ms_per_step * 1e-3)

    clocks_hint = None
    if sampler:
        try:
            vals = [float(ln.split(",")[1]) for ts, ln in sampler.lines
                    if sampler.t_begin and sampler.t_begin - 0.05 <= ts <= (sampler.t_end or ts) + 0.15 and len(ln.split(",")) >= 9]
            clocks_hint = float(np.median(vals)) if vals else None
        except Exception:
            clocks_hint = None

    # ---- roofline of the dominant kernel (traverse_kernel): algorithmic bytes per SURVEY.md §8d ----
    t.set_option("counters", 1)
    step()
    cs = t.stats()
    t.set_option("counters", 0)
    nq = my_queries
    alg_bytes = nq * (16 + 8 * k) + 64 * cs["nodes_visited"] + 16 * cs["points_tested"]
    warp_bytes = nq * (16 + 8 * k) + 64 * cs["warp_node_visits"] + 16 * cs["warp_point_loads"]
    kern_ms_per_step = kernel_ms_sum / args.steps
    peak, peak_src = _peaks()
    achieved = alg_bytes / (kern_ms_per_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traverse_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_query"] * nq  # ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum
        except Exception:
            traffic = None
    # measured L2 read bandwidth of this pool's B200 (tools/measure_peaks.py) and the issue-slot view of the kernel
    l2_peak = None
    try:
        l2_peak = float(json.load(open(os.path.join(ROOT, "profiles", "measured_l2_hbm_read.json")))["l2_gbs"])
    except Exception:
        pass
    issue = None
    try:
        ij = json.load(open(os.path.join(ROOT, "profiles", "traverse_issue_profile.json")))
        sm_count, sm_hz = 148, (clocks_hint or 1965.0) * 1e6
        winst = ij["warp_instructions_per_query"] * nq
        peak_issue = sm_count * 4 * sm_hz                     # one warp instruction per SMSP per cycle
        issue = {"warp_instructions_per_query": ij["warp_instructions_per_query"], "source": ij["source"],
                 "achieved_warp_inst_per_s": winst / (kern_ms_per_step * 1e-3), "peak_warp_inst_per_s": peak_issue,
                 "frac": winst / (kern_ms_per_step * 1e-3) / peak_issue,
                 "note": "instruction count from the committed ncu capture of the same kernel; time and clock measured live"}
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "kernel": "tknn::trav::traverse_kernel", "kernel_ms_per_step": kern_ms_per_step,
        "kernel_share_of_step": kern_ms_per_step / ms_per_step, "peak_source": peak_src,
        "algorithmic_bytes_per_query": alg_bytes / max(nq, 1),
        "bytes_per_query_loaded_once_per_warp": warp_bytes / max(nq, 1),
        "achieved_warp_shared_gbs": warp_bytes / (kern_ms_per_step * 1e-3) / 1e9,
        "l2_peak_gbs": l2_peak, "frac_of_l2_peak": (achieved / l2_peak) if l2_peak else None, "issue_slots": issue,
        "nodes_per_query": cs["nodes_visited"] / max(nq, 1), "points_tested_per_query": cs["points_tested"] / max(nq, 1),
        "note": "SURVEY.md §8d counts every query's node/point reads; the kernel loads each once per 32-query warp and "
                "broadcasts, so it is instruction-issue bound (ncu: issue slots ~83% busy, DRAM ~3% of peak) — see DESIGN.md",
    }

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((n_points, 3), dtype=torch.float32, pin_memory=True)
        xh.copy_(xd)
        qid_h = torch.empty((cap,), dtype=torch.int32, pin_memory=True)
        idx_h = torch.empty((cap, k), dtype=torch.int32, pin_memory=True)
        dst_h = torch.empty((cap, k), dtype=torch.float32, pin_memory=True)
        xh_np, out_np = xh.numpy(), (qid_h.numpy(), idx_h.numpy(), dst_h.numpy())

        full_np = (idx_h.numpy()[:n_points], dst_h.numpy()[:n_points]) if world == 1 else None
        sh = None
        if world > 1:
            from owlraytracing_b200.sharded import ShardedTrueKNN

            sh = ShardedTrueKNN(engine=t)
            per = (n_points + world - 1) // world
            my_slice = xh[rank * per: min(n_points, (rank + 1) * per)]   # this rank's 1/N of the pinned host cloud

        def e2e_step():
            if world == 1:
                t.build(xh_np)                                 # H2D of the points + LBVH build
                if not args.e2e_shard_api:
                    # the reference-facing call: rows in file order (tknn_search), search + D2H
                    return t.search(k, args.start_radius, out=full_np)
            else:
                # every rank uploads 1/N of the cloud, one all_gather over NVLink replicates it, then build
                sh.build_from_slices(my_slice, n_points, device=dev)
            return t.search_shard(k, rank, world, start_radius=args.start_radius, out=out_np)  # search + D2H

        if args.output_chunks > 0:
            t.set_option("output_chunks", args.output_chunks)
        if args.file_order_chunks > 0:
            t.set_option("file_order_chunks", args.file_order_chunks)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        ev1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        es = t.stats()
        e_ms = max(ev0.elapsed_time(ev1), 0.0)
        tm = torch.tensor([e_ms, wall_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e_ms = float(tm[0].item()) / e2e_steps
        e2e = {"value": total_queries / (e_ms * 1e-3), "unit": "queries/s",
               "h2d_bytes_per_step": int(n_points * 12) if world == 1 else int(my_slice.shape[0] * 12),
               "d2h_bytes_per_step": int(es["d2h_bytes"]), "ms_per_step": e_ms,
               "h2d_note": "per rank: its 1/N slice of the points (pinned); one NCCL all_gather replicates the cloud" if world > 1 else "all points (pinned)", "wall_ms_per_step": float(tm[1].item()) / e2e_steps,
               "steps": e2e_steps, "includes": "H2D points (pinned) + LBVH build + search (all rounds) + D2H results (pinned)",
               "api": "tknn_build + tknn_search (rows in file order)" if (world == 1 and not args.e2e_shard_api)
               else "tknn_build + tknn_search_shard (compact Morton-order rows + query ids)"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_arm(cfg, n_points, steps=2, warmup=1, sample_queries=args.cpu_sample, label="cpu_baseline")
        cpu = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}

    clocks = sampler.stop() if sampler else None
    if rank == 0:
        line = {
            "metric": "knn_queries_per_s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": cfg["name"] if not weak_grow else
                f"cfg2 weak-scaled: query-sharded kNN k=10, {n_points} uniform points replicated per GPU, 10M queries per GPU",
                "n_points": n_points, "k": k, "queries_total": total_queries, "queries_per_gpu": my_queries,
                "sharding": "queries sharded by contiguous Morton slice, BVH replicated, no data-path collective",
                "start_radius": "auto (sampled k-th-neighbour distance)" if args.start_radius <= 0 else args.start_radius,
                "l2_policy": f"inputs larger than L2: {n_points * 16 / 1e6:.0f} MB points + {bstats['n_nodes'] * 64 / 1e6:.0f} MB nodes "
                             f"+ {my_queries * k * 8 / 1e6:.0f} MB results per step vs 126 MB L2",
                "timed_region": "search only (estimate + all rounds), CUDA events on the launching stream; build reported separately",
            },
            "build_ms": float(np.median(build_ms)), "build_phases_ms": {p: bstats[p] for p in (
                "bounds_ms", "morton_ms", "sort_ms", "leaves_ms", "hierarchy_ms", "refit_ms")},
            "rounds": int(sstats["rounds"]), "round_queries": sstats["round_queries"], "round_ms": sstats["round_ms"],
            "start_radius": sstats["start_radius"],
            "queries_per_s_incl_build": total_queries / ((ms_per_step + float(np.median(build_ms))) * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override the number of points (debug)")
    ap.add_argument("--start-radius", type=float, default=0.0, help="<= 0: auto")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="queries per CPU step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--output-chunks", type=int, default=0, help="e2e: Morton slices whose D2H overlaps the search (0 = library default)")
    ap.add_argument("--file-order-chunks", type=int, default=0, help="e2e through tknn_search: slices by original index (0 = default)")
    ap.add_argument("--e2e-shard-api", action="store_true", help="N=1 e2e through tknn_search_shard (compact Morton rows) instead of tknn_search")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
