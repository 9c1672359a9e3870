"""Timeline of CTA 0 of the generic tcgen05 layer kernel (KNN layer 2 shape: [64,512,576] -> 512)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from sonet_b200 import _C, ops  # noqa: E402
B, C, P, Cout = 64, 512, 576, 512
x = torch.randn(B, C, P, device="cuda")
W = torch.randn(Cout, C) / 22
blob, inv = ops.pointwise_tc_pack(W)
blob = blob.cuda()
out = torch.empty(B, Cout, P, device="cuda")
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
for _ in range(3):
    _C.check(_C.lib().sonet_debug_pointwise_tc_timeline(x.data_ptr(), C, B, P, blob.data_ptr(), inv, Cout,
                                                        out.data_ptr(), tl.data_ptr(), None), "tl")
torch.cuda.synchronize()
t = tl.cpu().tolist()
t0 = t[127]
ev = [(t[126] - t0, "kernel end (thread 0)")]
for kc in range(8):
    if t[kc]: ev.append((t[kc] - t0, "TMA  issue W chunk %d" % kc))
    for j, n in enumerate(("full_w seen", "full_a seen", "issued+commit")):
        v = t[32 + 3 * kc + j]
        if v: ev.append((v - t0, "MMA  chunk %d %s" % (kc, n)))
for q in range(6, 16):     # converter marks of chunks 6..15 (end of item 0, all of item 1)
    for j, n in enumerate(("fp32 landed", "A slot free", "stored+arrive")):
        v = t[64 + 3 * (q - 6) + j]
        if v: ev.append((v - t0, "CONV chunk %d %s" % (q, n)))
for it in range(4):
    for j, n in enumerate(("d_full seen", "stored")):
        v = t[96 + 2 * it + j]
        if v: ev.append((v - t0, "EPI  item %d %s" % (it, n)))
prev = 0
for c, n in sorted(ev):
    print("%8d (+%6d) %s" % (c, c - prev, n))
    prev = c
