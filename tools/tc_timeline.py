"""Print the in-kernel timeline of one steady-state tile of the fused tcgen05 PointResNet.
    python tools/tc_timeline.py   (on a GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from sonet_b200 import _C, layers, synth  # noqa: E402

net = layers.PointResNet(6, [64, 128, 256, 384], 'relu', 'batch')
net.load_state_dict(synth.synth_state_dict(net, seed=1))
net = net.eval().cuda()
x = torch.randn(64, 6, 15000, device="cuda")
with torch.no_grad():
    net(x)
    blob, fpar = net._tc_params()
    out = torch.empty(64, 384, 15000, device="cuda")
    tl = torch.zeros(64, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _C.check(_C.lib().sonet_debug_pointresnet_tc_timeline(
            x.data_ptr(), 6, 64, 15000, blob.data_ptr(), fpar.data_ptr(), out.data_ptr(),
            tl.data_ptr(), None), "timeline")
    torch.cuda.synchronize()
t = tl.cpu().tolist()
mma, epi = t[:32], t[32:]
t0 = min(v for v in (mma[:13] + epi[:14]) if v > 0)
names_m = ["act0 ready", "L1 issued", "act1 ready", "L2 issued", "act2 ready"] + \
    [s for nc in range(4) for s in ("c%d start" % nc, "c%d issued" % nc)]
names_e = ["L0 start", "L0 done", "d1 ready", "epi1 done", "d2 ready", "epi2 done"] + \
    [s for nc in range(4) for s in ("c%d full" % nc, "c%d stored" % nc)]
ev = [(v - t0, "MMA  " + n) for v, n in zip(mma, names_m) if v > 0] + \
     [(v - t0, "EPI  " + n) for v, n in zip(epi, names_e) if v > 0]
prev = 0
for c, n in sorted(ev):
    print("%8d  (+%6d)  %s" % (c, c - prev, n))
    prev = c
