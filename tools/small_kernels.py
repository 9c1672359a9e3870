"""Run the small SOM / node-stage kernels of the classifier step once each on the bench input
(for ncu captures):  ncu -k regex:som_group ... python tools/small_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from sonet_b200 import ops, synth  # noqa: E402

inp = synth.synth_inputs(64, 5000, seed=0)
pc, sn, node = inp["pc"].cuda(), inp["sn"].cuda(), inp["node"].cuda()
for _ in range(3):
    a = ops.som_assign(pc, node, 3, want_stats=False)
    out = ops.som_group_decenter(pc, sn, a["min_idx_i32"], 64, 3)
torch.cuda.synchronize()
print("ok", out[3].sum().item())
