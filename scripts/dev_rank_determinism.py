"""Every rank: whole matrix on its own GPU, tie-class kernel vs general kernel, repeated; prints mismatching cells."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_cells, n_genes, R = int(sys.argv[1]), 27998, 500
rs = np.random.RandomState(1234)
reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=True, repressors=True)
cutoff = derive_rank_cutoff(0.05, n_genes)
tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff)
plan = AUCellPlan(n_genes, np.random.RandomState(42).permutation(n_genes), tables, device=local)
indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1960, seed=77, device=dev)
res = {"rank": rank, "gpu": torch.cuda.get_device_name(local)}
plan.set_fast_path(False); ref = plan.aucell_csr(indptr, indices, data); plan.set_fast_path(True)
for k in range(3):
    got = plan.aucell_csr(indptr, indices, data)
    bad = (got != ref).any(dim=1).nonzero().flatten()
    res["run%d_bad" % k] = int(bad.numel()); res["run%d_first" % k] = bad[:6].tolist()
# checksum of the input and of the reference, to compare across ranks
res["input_sum"] = float(data.double().sum().item()) + float(indices.double().sum().item())
res["ref_sum"] = float(ref.sum().item())
print(json.dumps(res), flush=True)
dist.barrier(); plan.close(); dist.destroy_process_group()
