# -*- coding: utf-8 -*-
"""One process per GPU (torchrun): partition ONE synthetic CSR matrix by stored entries, run every rank's shard from
pinned host memory, gather the AUC blocks over NCCL to rank 0 and compare there, bit for bit, with the whole matrix
computed on rank 0's GPU alone.  Prints one JSON line; used by tests/test_multigpu.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sharded_check.py --cells 60000
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    import synth
    from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff
    from pyscenic_b200.sharding import aucell_csr_shard_from_host, aucell_csr_shard_to_shared_host, scatter_csr_shards

    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=60000)
    ap.add_argument("--genes", type=int, default=27998)
    ap.add_argument("--regulons", type=int, default=500)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rs = np.random.RandomState(1234)
    reg = synth.make_regulons(rs, args.genes, args.regulons, 10, 400, weighted=True, repressors=True)
    cutoff = derive_rank_cutoff(0.05, args.genes)
    tables = synth.regulon_tables_from_arrays(*reg, n_genes=args.genes, rank_cutoff=cutoff)
    perm = np.random.RandomState(42).permutation(args.genes)
    plan = AUCellPlan(args.genes, perm, tables, device=local)
    # ONE matrix, generated on rank 0 (torch's CUDA cumsum is not bit-reproducible from one device to the next, so
    # "the same seed on every rank" does not give the same matrix); every rank receives its shard and keeps it in
    # pinned host memory
    indptr = indices = data = None
    if rank == 0:
        indptr, indices, data = synth.synth_csr_torch(args.cells, args.genes, 1960, seed=77, device=dev)
    bounds, s_ip, s_ix, s_dv = scatter_csr_shards(indptr, indices, data)
    h_ip, h_ix, h_dv = s_ip.cpu().pin_memory(), s_ix.cpu().pin_memory(), s_dv.cpu().pin_memory()
    del s_ip, s_ix, s_dv
    timings = {}
    full = aucell_csr_shard_from_host(plan, h_ip, h_ix, h_dv, bounds, timings=timings)
    t_shm = {}
    try:
        full_shm = aucell_csr_shard_to_shared_host(plan, h_ip, h_ix, h_dv, bounds, timings=t_shm)  # no device-side gather
    except RuntimeError as e:                      # /dev/shm too small on this box (every rank raises together)
        full_shm = None
        t_shm = {"unavailable": str(e)}
    if rank == 0:
        want = plan.aucell_csr(indptr, indices, data)
        got = full.to(dev)
        same = bool(torch.equal(got, want))
        same_shm = True if full_shm is None else bool(torch.equal(torch.from_numpy(np.ascontiguousarray(full_shm)).to(dev), want))
        timings["identical_shared_host_memory"] = same_shm
        timings["shared_host_memory"] = t_shm
        info = {}
        if not same:
            bad = (got != want).any(dim=1).nonzero().flatten()
            info = {"bad_cells": int(bad.numel()), "first_bad": bad[:8].tolist(), "last_bad": bad[-8:].tolist(),
                    "max_abs_diff": float((got - want).abs().max().item()),
                    "again_equal": bool(torch.equal(plan.aucell_csr(indptr, indices, data), want))}
        print(json.dumps({"world": world, "cells": args.cells, "identical": same, "bounds": bounds, **timings, **info}), flush=True)
    dist.barrier()
    plan.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
