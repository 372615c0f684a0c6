#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
bench.py -- AUCell cells/s on B200 (BASELINE.json metric), with roofline and CPU-reference figures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg3|cfg2|rank]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (fused ranking + AUC of every regulon, one kernel launch) over one
synthetic expression matrix that is resident in HBM.  Workload at N=1 (BASELINE.json: the configuration the
metric is quoted on): 1.3 M cells x 27,998 genes CSR (~1,960 nnz/cell), 500 weighted regulons incl. (-) regulons,
auc_threshold 0.05 (rank cutoff 1,399), seed 42.  Multi-GPU is weak scaling: every rank owns a shard of the same
shape (cells partition independently; no collective on the data path), `value` = all cells / max-over-ranks time.

`e2e` is the same metric through the host-buffer C-ABI entry point (aucell_csr_f32_host): pinned HOST CSR in,
pinned HOST AUC matrix out, H2D/D2H inside the timed region.  At N > 1 `e2e` is the STRONG-scaling job BASELINE.json's
config 4 describes: ONE 1.3 M-cell matrix in host memory, partitioned over the ranks by stored entries
(sharding.shard_bounds_by_nnz), every rank uploads and ranks its shard, the [cells x regulons] blocks are gathered over
NCCL to rank 0 and copied to host memory there -- all inside the timed region -- and rank 0 then recomputes the whole
matrix on its own GPU and checks the gathered result bit for bit (`sharded.identical`).  The weak-scaling host-buffer
number of round 1 stays in the line as `e2e_weak`.

`--impl reference` times the reference's CPU implementation of the path (oracle port: pandas create_rankings
call-for-call + fork/RawArray regulon-chunk workers as in aucell.py:131-181 with the ctxcore AUC restated in C;
the reference itself cannot run -- ctxcore is not installed) on a bounded cell sample per step.
"""
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
    # name: (n_cells, n_genes, kind, mean_nnz or density, n_regulons, description)
    "cfg4": dict(n_cells=1_300_000, n_genes=27_998, kind="csr", mean_nnz=1960, n_regulons=500,
                 desc="1.3M cells x 27,998 genes CSR (~1,960 nnz/cell), 500 weighted regulons incl. (-)"),
    "cfg3": dict(n_cells=100_000, n_genes=25_000, kind="csr", mean_nnz=1750, n_regulons=500,
                 desc="100k cells x 25k genes CSR (~7% nnz), 500 weighted regulons incl. (-)"),
    "cfg2": dict(n_cells=10_000, n_genes=20_000, kind="dense", density=0.10, n_regulons=300,
                 desc="10k cells x 20k genes dense f32 (~10% non-zero), 300 weighted regulons"),
}
AUC_THRESHOLD = 0.05
SEED = 42


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- workload
def make_regulon_tables(w, rs, weighted=True):
    import synth
    from pyscenic_b200.plan import derive_rank_cutoff

    reg = synth.make_regulons(rs, w["n_genes"], w["n_regulons"], 10, 400, weighted=weighted, repressors=True)
    cutoff = derive_rank_cutoff(AUC_THRESHOLD, w["n_genes"])
    tables = synth.regulon_tables_from_arrays(*reg, n_genes=w["n_genes"], rank_cutoff=cutoff, unweighted=not weighted)
    return reg, tables, cutoff


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
                break
        else:
            avail = 16 << 30
    except Exception:
        avail = 16 << 30
    for cg in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(cg).read().strip()
            if v.isdigit():
                avail = min(avail, int(v))
        except Exception:
            pass
    return avail


def cpu_reference_run(w, reg, n_sample, seed, workers):
    """Reference-equivalent CPU AUCell (oracle port) on `n_sample` synthetic cells of workload `w`.
    Returns (seconds, rank_seconds, auc DataFrame values, dense sample)."""
    import pandas as pd

    import synth
    from oracle import aucell_oracle as O

    if w["kind"] == "csr":
        ip, ix, dv = synth.synth_csr_torch(n_sample, w["n_genes"], w["mean_nnz"], seed=seed, device="cpu")
        X = synth.csr_to_dense(ip.numpy(), ix.numpy(), dv.numpy(), w["n_genes"])
    else:
        X = synth.dense_counts(np.random.RandomState(seed), n_sample, w["n_genes"], w["density"])
    return cpu_reference_on(X, reg, workers) + (X,)


def cpu_reference_on(X, reg, workers):
    import pandas as pd

    import synth
    from oracle import aucell_oracle as O

    df = pd.DataFrame(X, index=np.arange(X.shape[0]), columns=["g%d" % j for j in range(X.shape[1])])
    sigs = synth.regulons_to_signatures(*reg)
    t0 = time.perf_counter()
    rnk = O.create_rankings(df, SEED)
    t1 = time.perf_counter()
    auc = O.aucell4r(rnk, sigs, AUC_THRESHOLD, False, False, workers)
    t2 = time.perf_counter()
    return (t2 - t0, t1 - t0, auc.values)


# ------------------------------------------------------------------------------------------------- arms
def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    rs = np.random.RandomState(1234)
    reg, tables, cutoff = make_regulon_tables(w, rs, not args.unweighted)
    n_sample = args.ref_cells
    from oracle import aucell_oracle as O

    O.c_lib()
    times = []
    for it in range(args.warmup + args.steps):
        secs, rsecs, _, _ = cpu_reference_run(w, reg, n_sample, seed=100 + it, workers=cores)
        log("reference step %d: %.2fs (ranking %.2fs)" % (it, secs, rsecs))
        if it >= args.warmup:
            times.append(secs)
    total = sum(times)
    value = n_sample * len(times) / total
    sample = "%d synthetic cells per step (densified), %s" % (n_sample, w["desc"])
    line = {
        "impl": "reference", "metric": "aucell_cells_per_sec", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 keys -> u32 ranks, f64 AUC",
        "data": "synthetic",
        "config": {"workload": w["desc"], "auc_threshold": AUC_THRESHOLD, "rank_cutoff": cutoff, "seed": SEED,
                   "reference": "oracle port of pyscenic.aucell.aucell (pandas ranking, single thread, as the "
                                "reference; fork/RawArray regulon workers; ctxcore AUC restated in C)"},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import synth
    from pyscenic_b200 import _native
    from pyscenic_b200.plan import AUCellPlan

    w = dict(WORKLOADS[args.workload])
    if args.cells:
        w["n_cells"] = args.cells
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rs = np.random.RandomState(1234)                       # regulons identical on every rank
    reg, tables, cutoff = make_regulon_tables(w, rs, not args.unweighted)
    perm = np.random.RandomState(SEED).permutation(w["n_genes"])
    plan = AUCellPlan(w["n_genes"], perm, tables, device=local_rank)
    n_cells, n_genes, R = w["n_cells"], w["n_genes"], w["n_regulons"]

    t0 = time.perf_counter()
    if w["kind"] == "csr":
        indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, w["mean_nnz"], seed=1000 + rank, device=dev)
        total_nnz = int(indptr[-1].item())
        alg_bytes = 8 * total_nnz + 8 * n_cells + 8 * R * n_cells
        in_bytes = indptr.numel() * 8 + indices.numel() * 4 + data.numel() * 4
    else:
        X = torch.from_numpy(synth.dense_counts(np.random.RandomState(1000 + rank), n_cells, n_genes, w["density"])).to(dev)
        total_nnz = int((X != 0).sum().item())
        alg_bytes = 4 * n_genes * n_cells + 8 * R * n_cells
        in_bytes = X.numel() * 4
    out = torch.empty((n_cells, R), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    log("[rank %d] generated %s in %.1fs: %d cells, %.1f nnz/cell, input %.2f GB" %
        (rank, args.workload, time.perf_counter() - t0, n_cells, total_nnz / n_cells, in_bytes / 1e9))

    def step():
        if w["kind"] == "csr":
            plan.aucell_csr(indptr, indices, data, out=out)
        else:
            plan.aucell_dense(X, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _native.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    path_stats = plan.path_stats()
    fast_stats = plan.fast_stats()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_cells * args.steps / (ms_max / 1e3)

    # ---- the same launch with unit gene weights (`pyscenic aucell` without --weights; noweights=True): secondary ----
    unit_weights = None
    if world == 1 and w["kind"] == "csr" and not args.unweighted:
        _, tables_u, _ = make_regulon_tables(w, np.random.RandomState(1234), False)
        with AUCellPlan(w["n_genes"], perm, tables_u, device=local_rank) as plan_u:
            for _ in range(2):
                plan_u.aucell_csr(indptr, indices, data, out=out)
            torch.cuda.synchronize()
            eu0, eu1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_u = max(1, min(args.steps, 5))
            eu0.record()
            for _ in range(n_u):
                plan_u.aucell_csr(indptr, indices, data, out=out)
            eu1.record()
            torch.cuda.synchronize()
            ms_u = eu0.elapsed_time(eu1) / n_u
        unit_weights = {"value": n_cells / (ms_u / 1e3), "unit": "cells/s", "ms_per_step": ms_u, "steps": n_u,
                        "what": "the same matrix and regulons with unit gene weights (exact integer AUC numerators)"}
        for _ in range(1):
            step()              # `out` holds the weighted result again (the parity sample below reads it)
        torch.cuda.synchronize()

    # ---- end-to-end through the host-buffer C-ABI entry point --------------------------------------------
    e2e = None
    if not args.no_e2e:
        bytes_per_cell = (in_bytes + 8 * R * n_cells) / n_cells
        budget = 0.30 * mem_available_bytes() / world
        n_e2e = int(min(n_cells, max(1024, budget // bytes_per_cell)))
        if w["kind"] == "csr":
            e_end = int(indptr[n_e2e].item())
            h_ip = indptr[: n_e2e + 1].cpu().pin_memory()
            h_ix = indices[:e_end].cpu().pin_memory()
            h_dv = data[:e_end].cpu().pin_memory()
            h2d = h_ip.numel() * 8 + h_ix.numel() * 4 + h_dv.numel() * 4
        else:
            h_X = X[:n_e2e].cpu().pin_memory()
            h2d = h_X.numel() * 4
        h_out = torch.empty((n_e2e, R), dtype=torch.float64).pin_memory()
        d2h = h_out.numel() * 8

        def e2e_step():
            if w["kind"] == "csr":
                plan.aucell_csr_host(h_ip.numpy(), h_ix.numpy(), h_dv.numpy(), out=h_out.numpy())
            else:
                h_out.numpy()[:] = plan.aucell_dense_host(h_X.numpy())

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        l0 = _native.launch_count()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(e2e_steps):
            e2e_step()          # returns when h_out is complete (the call synchronises its own streams)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        launches += _native.launch_count() - l0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        moved_h2d, moved_d2h = plan.last_host_bytes()      # what the call really sent (uint16 indices when they fit)
        e2e = {"value": world * n_e2e * e2e_steps / float(tt.item()), "unit": "cells/s",
               "h2d_bytes_per_step": int(moved_h2d), "d2h_bytes_per_step": int(moved_d2h),
               "host_input_bytes_per_step": int(h2d), "host_output_bytes_per_step": int(d2h),
               "cells_per_step_per_gpu": n_e2e,
               "steps": e2e_steps, "timer": "host wall clock around the blocking C-ABI call, max over ranks"}
        # the device-resident and the host-buffer paths must agree bit for bit
        same = bool(torch.equal(h_out.to(dev), out[:n_e2e]))
        e2e["matches_device_path"] = same

    # ---- the call a user makes: pyscenic_b200.aucell.aucell(...) from ordinary (pageable) numpy / pandas memory --------
    e2e_api = None
    if world == 1 and w["kind"] == "csr" and not args.no_e2e and not args.no_api:
        import pandas as pd
        import scipy.sparse as sp

        from pyscenic_b200.aucell import aucell as api_aucell

        n_api = int(min(n_cells, args.api_cells))
        e_end = int(indptr[n_api].item())
        csr = sp.csr_matrix((data[:e_end].cpu().numpy(), indices[:e_end].cpu().numpy(), indptr[: n_api + 1].cpu().numpy()),
                            shape=(n_api, n_genes))
        gene_names = ["g%d" % j for j in range(n_genes)]
        sigs = synth.regulons_to_signatures(*reg)
        api_aucell(csr[:2048], sigs, auc_threshold=AUC_THRESHOLD, seed=SEED, genes=gene_names)      # library load, context
        # first call of this size: the library page-locks its staging buffers (kept for the life of the process)
        t0 = time.perf_counter()
        got_api = api_aucell(csr, sigs, auc_threshold=AUC_THRESHOLD, seed=SEED, genes=gene_names)
        dt_csr_first = time.perf_counter() - t0
        del got_api                               # (freeing 1.2 GB is not part of the next call)
        t0 = time.perf_counter()
        got_api = api_aucell(csr, sigs, auc_threshold=AUC_THRESHOLD, seed=SEED, genes=gene_names)
        dt_csr = time.perf_counter() - t0
        ref_api = out[:n_api].cpu().numpy()      # (the plans differ in the order max_auc's weights are summed: last-bit differences)
        same_api = bool(np.allclose(got_api.values, ref_api, rtol=1e-12, atol=0))
        n_df = int(min(n_api, args.api_frame_cells))
        frame = pd.DataFrame(csr[:n_df].toarray(), index=["c%d" % i for i in range(n_df)], columns=gene_names)
        t0 = time.perf_counter()
        got_df = api_aucell(frame, sigs, auc_threshold=AUC_THRESHOLD, seed=SEED)
        dt_df_first = time.perf_counter() - t0
        del got_df
        t0 = time.perf_counter()
        got_df = api_aucell(frame, sigs, auc_threshold=AUC_THRESHOLD, seed=SEED)
        dt_df = time.perf_counter() - t0
        same_df = bool(np.allclose(got_df.values, ref_api[:n_df], rtol=1e-12, atol=0))
        e2e_api = {"aucell_scipy_csr": {"value": n_api / dt_csr, "unit": "cells/s", "cells": n_api, "seconds": dt_csr,
                                        "seconds_first_call": dt_csr_first,
                                        "matches_device_path_1e-12": same_api,
                                        "what": "pyscenic_b200.aucell.aucell(scipy.sparse.csr_matrix, signatures, genes=...) from "
                                                "pageable numpy arrays to the returned DataFrame, incl. regulon preparation"},
                   "aucell_dataframe": {"value": n_df / dt_df, "unit": "cells/s", "cells": n_df, "seconds": dt_df,
                                        "seconds_first_call": dt_df_first,
                                        "matches_device_path_1e-12": same_df,
                                        "what": "aucell(float32 DataFrame [cells x genes], signatures) -> DataFrame"}}
        del csr, frame, got_api, got_df

    # ---- strong scaling (N > 1): one matrix, cell-sharded, gathered over NCCL, verified ------------------------------
    sharded = None
    if world > 1 and w["kind"] == "csr" and not args.no_e2e:
        from pyscenic_b200.sharding import aucell_csr_shard_from_host, scatter_csr_shards

        n_tot = args.sharded_cells or n_cells
        # host memory: every rank pins its shard (input / world) and rank 0 the result
        need = (in_bytes / n_cells) * n_tot + 8.0 * R * n_tot
        if need > 0.5 * mem_available_bytes():
            n_tot = int(n_tot * 0.5 * mem_available_bytes() / need)
        del out
        # ONE matrix: rank 0's (torch's CUDA cumsum is not bit-reproducible from one device to the next, so the same
        # seed does not give the same matrix on another GPU); the other ranks drop theirs and receive their shard
        s_ip = s_ix = s_dv = None
        if rank == 0:
            if n_tot == n_cells:
                s_ip, s_ix, s_dv = indptr, indices, data
            else:
                s_ip, s_ix, s_dv = synth.synth_csr_torch(n_tot, n_genes, w["mean_nnz"], seed=1000, device=dev)
        else:
            del indptr, indices, data
            torch.cuda.empty_cache()
        bounds, r_ip, r_ix, r_dv = scatter_csr_shards(s_ip, s_ix, s_dv)
        h_ip, h_ix, h_dv = r_ip.cpu().pin_memory(), r_ix.cpu().pin_memory(), r_dv.cpu().pin_memory()
        del r_ip, r_ix, r_dv
        best = None
        full = None
        for it in range(3):                                   # first pass warms NCCL and the allocator; keep the best
            tm = {}
            barrier()
            t0 = time.perf_counter()
            full = aucell_csr_shard_from_host(plan, h_ip, h_ix, h_dv, bounds, timings=tm)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt, tm["upload_and_kernel_s"], tm["gather_s"], tm["to_host_s"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if it > 0 and (best is None or float(tt[0]) < best[0]):
                best = [float(x) for x in tt]
        # the same job with the result assembled in shared host memory (every rank's pipelined host call writes its own
        # cell range over its own PCIe link; no device-side gather, no single-link copy-out)
        from pyscenic_b200.sharding import aucell_csr_shard_to_shared_host

        best_shm = None
        full_shm = None
        shm_error = None
        for it in range(3):
            tm = {}
            barrier()
            t0 = time.perf_counter()
            try:
                full_shm = aucell_csr_shard_to_shared_host(plan, h_ip, h_ix, h_dv, bounds, timings=tm)
            except RuntimeError as e:          # (every rank raises together: /dev/shm too small on this box)
                shm_error = str(e)
                break
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt, tm["setup_s"], tm["upload_kernel_download_s"], tm["barrier_s"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if it > 0 and (best_shm is None or float(tt[0]) < best_shm[0]):
                best_shm = [float(x) for x in tt]
        if rank == 0:
            want = plan.aucell_csr(s_ip, s_ix, s_dv)
            identical = bool(torch.equal(full.to(dev), want))
            if shm_error is None:
                identical_shm = bool(torch.equal(torch.from_numpy(np.ascontiguousarray(full_shm)).to(dev), want))
                sharded_shm = {"cells": int(n_tot), "seconds": best_shm[0], "cells_per_s": n_tot / best_shm[0],
                               "setup_seconds": best_shm[1], "upload_kernel_download_seconds": best_shm[2],
                               "barrier_seconds": best_shm[3], "identical": identical_shm,
                               "what": "result assembled in shared host memory (/dev/shm): every rank's pipelined host call "
                                       "writes its own cell range; the gather is a barrier"}
            else:
                sharded_shm = {"cells_per_s": 0.0, "unavailable": shm_error}
            total_in = int(s_ip.numel() * 8 + s_ix.numel() * 4 + s_dv.numel() * 4)
            sharded = {"cells": int(n_tot), "seconds": best[0], "cells_per_s": n_tot / best[0],
                       "upload_and_kernel_seconds": best[1], "gather_seconds": best[2], "to_host_seconds": best[3],
                       "identical": identical, "partition": "contiguous cell ranges balanced by stored entries",
                       "h2d_bytes": total_in, "d2h_bytes": int(n_tot) * R * 8,
                       "timer": "host wall clock around upload + kernel + NCCL gather + copy to host, max over ranks, "
                                "best of 2 after one warm-up pass"}
            sharded["shared_host_memory"] = sharded_shm
            del want
        del full_shm
        launches += 6
        out = None

    # ---- CPU baseline + parity check on a bounded sample (rank 0, N=1 only) ------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = args.ref_cells
        if w["kind"] == "csr":
            e_end = int(indptr[n_s].item())
            Xs = synth.csr_to_dense(indptr[: n_s + 1].cpu().numpy(), indices[:e_end].cpu().numpy(),
                                    data[:e_end].cpu().numpy(), n_genes)
        else:
            Xs = X[:n_s].cpu().numpy()
        cores = os.cpu_count() or 1
        secs, rsecs, auc_ref = cpu_reference_on(Xs, reg, cores)
        got = out[:n_s].cpu().numpy()
        rel = np.abs(got - auc_ref) / np.maximum(np.abs(auc_ref), 1e-300)
        rel[auc_ref == 0] = 0
        parity = {"cells": n_s, "max_rel_err": float(rel.max()),
                  "zero_pattern_equal": bool(np.array_equal(got == 0, auc_ref == 0)), "tolerance": 1e-6}
        cpu_baseline = {"value": n_s / secs, "unit": "cells/s", "cores": cores, "kind": "port",
                        "sample": "first %d cells of the benchmark matrix, densified; ranking %.1fs (1 thread, as "
                                  "the reference) + AUC %.1fs (%d fork workers)" % (n_s, rsecs, secs - rsecs, cores)}

    if rank == 0:
        peak, peak_src = peaks()
        kernel_ms = ms_max / args.steps
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                t = json.load(open(tp)).get(args.workload)
                traffic = None if t is None else float(t["dram_bytes_per_cell"]) * n_cells   # per launch, like achieved
            except Exception:
                traffic = None
        line = {
            "metric": "aucell_cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 keys -> u32 ranks, f64 AUC", "data": "synthetic",
            "config": {"workload": w["desc"], "cells_per_gpu": n_cells, "n_genes": n_genes, "n_regulons": R,
                       "nnz_per_cell": total_nnz / n_cells, "auc_threshold": AUC_THRESHOLD, "rank_cutoff": cutoff,
                       "seed": SEED, "weighted": not args.unweighted,
                       "l2": "inputs (%.1f GB per GPU) exceed the 126 MB L2; no flush needed" % (in_bytes / 1e9)},
            "clocks": clocks,
            "e2e": e2e if sharded is None else {
                "value": max(sharded["cells_per_s"], sharded["shared_host_memory"]["cells_per_s"]), "unit": "cells/s",
                "h2d_bytes_per_step": sharded["h2d_bytes"],
                "d2h_bytes_per_step": sharded["d2h_bytes"], "cells_per_step": sharded["cells"], "steps": 2,
                "scaling": "strong: ONE matrix cell-sharded over the ranks; the result reaches rank 0's host memory either "
                           "through an NCCL gather + copy-out (`sharded`) or through shared host memory that every rank "
                           "writes over its own PCIe link (`sharded.shared_host_memory`); the faster of the two is this "
                           "value, both are checked bit for bit against a single-GPU run",
                "gather": ("shared_host_memory" if sharded["shared_host_memory"]["cells_per_s"] > sharded["cells_per_s"]
                           else "nccl"),
                "timer": sharded["timer"]},
            "e2e_weak": None if sharded is None else e2e,
            "e2e_api": e2e_api,
            "unit_weights": unit_weights,
            "sharded": sharded,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": ("aucell_tie_kernel<csr> (cells it hands over: aucell_fused_kernel<csr>, same stream, inside "
                                    "the timed region)" if w["kind"] == "csr" else "aucell_fused_kernel<dense>"),
                         "algorithmic_bytes_per_launch": int(alg_bytes),
                         "bytes_per_cell": alg_bytes / n_cells},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "path_stats": path_stats,
            "tie_kernel_stats": fast_stats,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def run_rank_workload(args):
    """--workload rank: create_rankings (reference aucell.py:25-51) alone, on a dense float32 matrix of the cfg4 gene
    count as a DataFrame holds it -- the secondary kernel path (aucell_rank_dense_f32).  One GPU.  Reads 4 bytes and
    writes 4 bytes (one uint32 rank) per gene and cell: the roofline counts both."""
    import pandas as pd
    import torch

    import synth
    from oracle import aucell_oracle as O
    from pyscenic_b200 import _native
    from pyscenic_b200.plan import AUCellPlan

    n_cells, n_genes = args.cells or 40_000, 27_998
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    perm = np.random.RandomState(SEED).permutation(n_genes)
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1960, seed=1000, device=dev)
    X = torch.zeros((n_cells, n_genes), dtype=torch.float32, device=dev)
    rows = torch.repeat_interleave(torch.arange(n_cells, device=dev), (indptr[1:] - indptr[:-1]))
    X[rows, indices.long()] = data
    del rows
    ranks = torch.empty((n_cells, n_genes), dtype=torch.int32, device=dev)
    alg_bytes = 8 * n_genes * n_cells
    with AUCellPlan(n_genes, perm, None, device=0) as plan:
        for _ in range(args.warmup):
            plan.rank_dense(X, out=ranks)
        torch.cuda.synchronize()
        sampler = ClockSampler(0)
        sampler.start()
        l0 = _native.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            plan.rank_dense(X, out=ranks)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        launches = _native.launch_count() - l0
        clocks = sampler.stop()
        csr_ranks = plan.rank_csr(indptr, indices, data)
        same_as_csr = bool(torch.equal(csr_ranks, ranks))
        del csr_ranks
        # end to end: host float32 matrix -> host uint32 rankings through the product's create_rankings()
        n_e = min(n_cells, 8192)
        from pyscenic_b200.aucell import create_rankings

        frame = pd.DataFrame(X[:n_e].cpu().numpy(), columns=["g%d" % j for j in range(n_genes)])
        create_rankings(frame.iloc[:256], seed=SEED)
        t0 = time.perf_counter()
        got = create_rankings(frame, seed=SEED)
        dt = time.perf_counter() - t0
    n_s = min(args.ref_cells, n_e)
    t0 = time.perf_counter()
    want = O.create_rankings(frame.iloc[:n_s], seed=SEED)          # the reference's own pandas call chain
    cpu_s = time.perf_counter() - t0
    exact = bool(np.array_equal(got.iloc[:n_s].values, want.values) and list(got.columns) == list(want.columns))
    peak, peak_src = peaks()
    achieved = alg_bytes / (ms / 1e3) / 1e9
    print(json.dumps({
        "metric": "create_rankings_cells_per_sec", "value": n_cells / (ms / 1e3), "unit": "cells/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 keys -> u32 ranks", "data": "synthetic",
        "config": {"workload": "create_rankings on %d cells x %d genes dense f32 (~1,960 non-zero per cell), seed %d"
                               % (n_cells, n_genes, SEED),
                   "l2": "input + output (%.1f GB) exceed the 126 MB L2" % (alg_bytes / 1e9)},
        "clocks": clocks,
        "e2e": {"value": n_e / dt, "unit": "cells/s", "h2d_bytes_per_step": 4 * n_e * n_genes,
                "d2h_bytes_per_step": 4 * n_e * n_genes, "cells": n_e,
                "what": "pyscenic_b200.aucell.create_rankings(DataFrame float32) -> DataFrame uint32, pageable memory"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "kernel": "aucell_fused_kernel<dense, RANKS>",
                     "algorithmic_bytes_per_launch": int(alg_bytes), "bytes_per_cell": 8 * n_genes},
        "cpu_baseline": {"value": n_s / cpu_s, "unit": "cells/s", "cores": 1, "kind": "port",
                         "sample": "first %d cells; the reference's pandas sample + rank(method='first') chain" % n_s},
        "parity": {"cells": n_s, "bit_exact": exact, "dense_equals_csr_path": same_as_csr},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS) + ["rank"])
    ap.add_argument("--cells", type=int, default=0, help="override cells per GPU (debugging)")
    ap.add_argument("--ref-cells", type=int, default=2048, help="cells in the CPU reference sample")
    ap.add_argument("--unweighted", action="store_true", help="unit gene weights (the CLI default of the reference)")
    ap.add_argument("--sharded-cells", type=int, default=0, help="cells of the strong-scaling matrix at N > 1 (default: the workload's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the aucell(scipy_csr) / aucell(DataFrame) records")
    ap.add_argument("--api-cells", type=int, default=300_000, help="cells of the aucell(scipy_csr) record")
    ap.add_argument("--api-frame-cells", type=int, default=20_000, help="cells of the aucell(DataFrame) record")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "rank":                # secondary workload (create_rankings alone): one GPU, our arm only
        if rank == 0:
            run_rank_workload(args)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)]
        cmd += sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
