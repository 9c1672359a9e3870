"""Time the fused MLP+pool kernel alone on the bench input (CUDA events, L2 flushed between reps).
    [SONET_TC_CLUSTER=2] python tools/pool_kernel_time.py   (on a GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from sonet_b200 import _C, layers, ops, synth  # noqa: E402

net = layers.PointResNet(6, [64, 128, 256, 384], 'relu', 'batch')
net.load_state_dict(synth.synth_state_dict(net, seed=1))
net = net.eval().cuda()
inp = synth.synth_inputs(64, 5000, seed=0)
pc, sn, node = inp["pc"].cuda(), inp["sn"].cuda(), inp["node"].cuda()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
with torch.no_grad():
    net(torch.randn(2, 6, 2048, device="cuda"))
    blob, fpar = net._tc_params()
    a = ops.som_assign(pc, node, 3)
    xs, ns, p0i = ops.som_sort_decenter(pc, sn, a["cluster_mean"], a["min_idx_i32"], a["count"], 3)
    keys = torch.empty(64, 384, 64, dtype=torch.int32, device="cuda")
    _C.check(_C.lib().sonet_pool_keys_init(keys.data_ptr(), keys.numel(), None), "init")
    p0 = torch.empty(64, 384, device="cuda")
    ms = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _C.check(_C.lib().sonet_pointresnet_tc_pool_forward(
            xs.data_ptr(), 6, 64, 15000, blob.data_ptr(), fpar.data_ptr(), ns.data_ptr(),
            p0i.data_ptr(), 64, keys.data_ptr(), p0.data_ptr(), None), "pool")
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
ms = sorted(ms[2:])
print("cluster=%s  pool kernel ms: min %.4f median %.4f max %.4f"
      % (os.environ.get("SONET_TC_CLUSTER", "1"), ms[0], ms[len(ms) // 2], ms[-1]))
