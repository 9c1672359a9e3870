"""Minimal driver for ncu captures of standalone ops at BASELINE.json sizes (B=64, N=5000, M=64):

    python tools/profile_ops.py --op query_topk|index_max|chamfer|som_train|augment [--reps 4]

so that `ncu -k regex:<kernel> -s 2 -c 1` sees only that op."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from sonet_b200 import ops, som, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", required=True)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--npts", type=int, default=5000)
ap.add_argument("--mp", type=int, default=1280)
a = ap.parse_args()
dev = "cuda:0"
B, N, M = a.batch, a.npts, 64
inp = synth.synth_inputs(B, N, seed=0)
pc, node = inp["pc"].to(dev), inp["node"].to(dev)
g = torch.Generator(device=dev).manual_seed(0)
if a.op == "query_topk":
    bs = som.BatchSOM(8, 8, 3, 0, B)
    bs.node = node
    fn = lambda: bs.query_topk(pc, 3)  # noqa: E731
elif a.op == "index_max":
    data = torch.randn(B, 384, 3 * N, device=dev, generator=g)
    index = torch.randint(0, M, (B, 3 * N), device=dev, generator=g, dtype=torch.int32)
    fn = lambda: ops.index_max(data, index, M, with_values=True)  # noqa: E731
elif a.op == "chamfer":
    pred = torch.rand(B // 2, 3, a.mp, device=dev, generator=g) * 2 - 1
    gt = pc[:B // 2].contiguous()
    fn = lambda: ops.chamfer(pred, gt)  # noqa: E731
elif a.op == "som_train":
    bs = som.BatchSOM(8, 8, 3, 0, B)
    fn = lambda: bs.optimize(pc)  # noqa: E731
elif a.op == "augment":
    sn = inp["sn"].to(dev)
    fn = lambda: ops.augment(pc, sn, node, jitter_pc=(0.01, 0.05), jitter_sn=(0.01, 0.05),  # noqa: E731
                             jitter_som=(0.04, 0.1), seed=1)
else:
    raise SystemExit("unknown op")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(a.reps):
    if i == a.reps - 1:
        ev[0].record()
    fn()
ev[1].record()
torch.cuda.synchronize()
print("done", a.op, "last rep %.4f ms" % ev[0].elapsed_time(ev[1]))
