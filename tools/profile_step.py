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

from sonet_b200 import autoencoder, classifier, networks, segmenter, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--npts", type=int, default=5000)
ap.add_argument("--task", default="classifier", choices=["classifier", "segmenter", "autoencoder"])
a = ap.parse_args()

mod = {"classifier": classifier, "segmenter": segmenter, "autoencoder": autoencoder}[a.task]
head = {"classifier": ("classifier", networks.Classifier), "segmenter": ("segmenter", networks.Segmenter),
        "autoencoder": ("decoder", networks.Decoder)}[a.task]
opt = synth.make_opt(a.task, batch_size=a.batch, input_pc_num=a.npts, device="cuda:0")
cpu_opt = synth.make_opt(a.task, batch_size=a.batch, input_pc_num=a.npts)
m = mod.Model(opt)
m.encoder.load_state_dict(synth.synth_state_dict(networks.Encoder(cpu_opt), seed=1))
getattr(m, head[0]).load_state_dict(synth.synth_state_dict(head[1](cpu_opt), seed=2))
inp = synth.synth_inputs(a.batch, a.npts, seed=0)
args = [inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"]]
if a.task == "segmenter":
    args.insert(3, torch.zeros(a.batch, a.npts, dtype=torch.int64))
m.set_input(*args)
for _ in range(a.steps):
    m.test_model()
torch.cuda.synchronize()
print("done", a.task)
