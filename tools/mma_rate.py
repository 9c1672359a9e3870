import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)
import torch
from sonet_b200 import _C
c = torch.zeros(1, dtype=torch.int64, device="cuda")
for mode in (0, 1):
    for N in (64, 96, 128, 256):
        for sbo in (256, 512, 1024):
            for _ in range(2):
                _C.check(_C.lib().sonet_debug_tc_mma_rate(mode, N, sbo, 512, c.data_ptr(), None), "rate")
            torch.cuda.synchronize()
            print("mode=%s N=%3d sbo=%4d  %6.1f cycles/MMA (ideal %d)" % ("SS" if mode == 0 else "TS", N, sbo, c.item() / 512.0, N // 2))
