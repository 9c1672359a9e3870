"""Minimal driver for ncu: build the cfg-2 classifier and run a few eval forwards (nothing else),
so that `ncu -k regex:<kernel> -s <skip> -c <n>` captures are cheap.

    python tools/profile_step.py [--steps 4] [--batch 64] [--npts 5000] [--task classifier]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from sonet_b200 import classifier, networks, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--npts", type=int, default=5000)
a = ap.parse_args()

opt = synth.make_opt("classifier", batch_size=a.batch, input_pc_num=a.npts, device="cuda:0")
cpu_opt = synth.make_opt("classifier", batch_size=a.batch, input_pc_num=a.npts)
m = classifier.Model(opt)
m.encoder.load_state_dict(synth.synth_state_dict(networks.Encoder(cpu_opt), seed=1))
m.classifier.load_state_dict(synth.synth_state_dict(networks.Classifier(cpu_opt), seed=2))
inp = synth.synth_inputs(a.batch, a.npts, seed=0)
m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
for _ in range(a.steps):
    m.test_model()
torch.cuda.synchronize()
print("done", float(m.score.sum()))
