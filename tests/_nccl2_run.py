"""Helper launched by tests/test_gpu_nccl2.py through torch.distributed.run (2 ranks, NCCL):
batch-sharded classifier forward; the all-gathered logits (torch.distributed AND the C-ABI
sonet_allgather) must equal the single-GPU forward of the full batch bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import build_states  # noqa: E402
from sonet_b200 import classifier, synth  # noqa: E402
from sonet_b200 import dist as sdist  # noqa: E402

B, N = 16, 2048
rank, local_rank, world = sdist.init_from_env(backend="nccl")
dev = "cuda:%d" % local_rank
keys = ("pc", "sn", "label", "node", "node_knn_I")


def build(batch):
    opt = synth.make_opt("classifier", batch_size=batch, input_pc_num=N, device=dev, gpu_id=local_rank)
    st = build_states("classifier", opt, seed=71)
    m = classifier.Model(opt)
    m.encoder.load_state_dict(st["encoder"])
    m.classifier.load_state_dict(st["head"])
    return m


inp = synth.synth_inputs(B, N, seed=71)
lo, hi = sdist.shard_bounds(B, rank, world)
m = build(hi - lo)
m.enable_cuda_graph(True)
for _ in range(2):                          # capture + replay
    m.set_input(*[inp[k][lo:hi] for k in keys])
    m.test_model()
out = sdist.all_gather_rows(m.score, B)
comm = sdist.SonetComm()
out2 = torch.empty_like(out)
comm.all_gather(m.score, out2)
torch.cuda.synchronize()
assert torch.equal(out, out2), "sonet_allgather != torch.distributed all_gather"
comm.destroy()
if rank == 0:
    full = build(B)
    full.set_input(*[inp[k] for k in keys])
    full.test_model()
    assert torch.equal(out, full.score), "gathered logits != single-GPU logits"
    print("NCCL2_OK")
dist.barrier()
dist.destroy_process_group()
