"""Helper of tests/test_multi_gpu.py: 8 EM iterations of VIPRS on a fixed synthetic genome (two chromosomes, ragged LD
blocks); run alone or under torchrun (shard=True); rank 0 writes [ELBO, pi, sigma_epsilon, tau_beta, max|eta_diff|] per
iteration to the JSON file given as argv[1]."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
OUT = sys.argv[1]
from viprs_b200 import synth
from viprs_b200.model import VIPRS
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
sizes = [700, 512, 900, 333, 1200, 64, 800, 1024]
inp = synth.make_inputs(sizes, ld_dtype="int8", float_dtype=torch.float32, device="cpu", seed=11, n=50000)
half = sum(sizes[:4])
def chrom(r0, r1):
    ip = inp["ld_indptr"].numpy(); lb = inp["ld_left_bound"].numpy()
    return dict(ld_data=inp["ld_data"].numpy()[ip[r0]:ip[r1]], ld_indptr=(ip[r0:r1 + 1] - ip[r0]), ld_left_bound=(lb[r0:r1] - r0).astype(np.int32),
                std_beta=inp["std_beta"].numpy()[r0:r1], n_per_snp=inp["n_per_snp"].numpy()[r0:r1])
data = {1: chrom(0, half), 2: chrom(half, sum(sizes))}
m = VIPRS(data=data, float_precision="float32", shard=world > 1)
m.initialize({"pi": 0.02, "sigma_epsilon": 0.8})
hist = []
for it in range(8):
    m.e_step(); m.m_step()
    hist.append([m.elbo(), float(m.pi), float(m.sigma_epsilon), float(m.tau_beta), m.max_eta_diff()])
if rank == 0:
    json.dump(hist, open(OUT, "w"))
if world > 1:
    dist.destroy_process_group()
