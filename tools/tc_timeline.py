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
    tl = torch.zeros(128, dtype=torch.int64, device="cuda")
    tl[125] = int(sys.argv[sys.argv.index("--tile") + 1]) if "--tile" in sys.argv else 3
    if "--pool" in sys.argv:
        from sonet_b200 import ops, synth as _s
        inp = _s.synth_inputs(64, 5000, seed=0)
        pc, sn, node = inp["pc"].cuda(), inp["sn"].cuda(), inp["node"].cuda()
        a = ops.som_assign(pc, node, 3)
        xs, ns, p0i = ops.som_sort_decenter(pc, sn, a["cluster_mean"], a["min_idx_i32"], a["count"], 3)
        keys = torch.empty(64, 384, 64, dtype=torch.int32, device="cuda")
        _C.check(_C.lib().sonet_pool_keys_init(keys.data_ptr(), keys.numel(), None), "init")
        p0 = torch.empty(64, 384, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(4):
            if i == 3:
                ev[0].record()
            _C.check(_C.lib().sonet_debug_pointresnet_tc_pool_timeline(
                xs.data_ptr(), 6, 64, 15000, blob.data_ptr(), fpar.data_ptr(), ns.data_ptr(),
                p0i.data_ptr(), 64, keys.data_ptr(), p0.data_ptr(), tl.data_ptr(), None), "timeline")
        ev[1].record()
        torch.cuda.synchronize()
        EVENT_MS = ev[0].elapsed_time(ev[1])
    else:
        for _ in range(2):
            _C.check(_C.lib().sonet_debug_pointresnet_tc_timeline(
                x.data_ptr(), 6, 64, 15000, blob.data_ptr(), fpar.data_ptr(), out.data_ptr(),
                tl.data_ptr(), None), "timeline")
    torch.cuda.synchronize()
t = tl.cpu().tolist()
starts = [v for v in t[64:124] if v > 0]
if starts:
    k0, k1 = t[64 + 62], t[64 + 63]
    print("CTA 0: kernel %d cycles, %d tiles; first act0-ready at +%d, last at +%d (end +%d after it)"
          % (k1 - k0, len(starts), starts[0] - k0, starts[-1] - k0, k1 - starts[-1]))
    print("tile periods:", [b - a for a, b in zip(starts, starts[1:])])
    if "EVENT_MS" in globals():
        # SM cycles of CTA 0 (clock64) against the CUDA-event duration of the same launch: the SM
        # clock the kernel actually ran at (nvidia-smi's 20 ms samples cannot see a 0.7 ms kernel)
        print("launch: %.4f ms by CUDA events, %d SM cycles -> effective SM clock %.0f MHz"
              % (EVENT_MS, k1 - k0, (k1 - k0) / EVENT_MS / 1e3))
mma, epi = t[:32], t[32:64]
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
