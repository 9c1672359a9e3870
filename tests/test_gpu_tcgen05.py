"""-m gpu: pins the tcgen05 encodings the fused point-MLP kernel relies on — instruction
descriptor, K-major no-swizzle shared-memory descriptors (LBO/SBO roles), and the A-operand
layout in tensor memory — with a single-tile GEMM probe against an exact bf16 reference."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _probe(A, Bm, mode, layout, swap):
    from sonet_b200 import _C
    N, K = Bm.shape
    D = torch.zeros(128, N, dtype=torch.float32, device=DEV)
    rc = _C.lib().sonet_debug_tc_probe(A.data_ptr(), Bm.data_ptr(), N, K, mode, layout, swap,
                                       D.data_ptr(), None)
    _C.check(rc, "tc_probe")
    torch.cuda.synchronize()
    return D


def _ref(A, Bm):
    a = A.to(torch.bfloat16).double().cpu()
    b = Bm.to(torch.bfloat16).double().cpu()
    return (a @ b.t()).float()


@pytest.mark.parametrize("N,K", [(64, 64), (128, 128), (64, 32), (16, 256)])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("layout", [0, 1])
def test_tcgen05_probe_matches_bf16_reference(N, K, mode, layout):
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(DEV)
    Bm = torch.randn(N, K, generator=g).to(DEV)
    want = _ref(A, Bm)
    got = _probe(A, Bm, mode, layout, swap=0).cpu()
    err = float((got - want).abs().max())
    assert err <= 1e-3 * max(1.0, float(want.abs().max())), \
        "N=%d K=%d mode=%d layout=%d: max abs err %.3e" % (N, K, mode, layout, err)


# ---- the fused tcgen05 first PointResNet (csrc/pointmlp_tc.cu) -------------------------------------
def _resnet(cin, seed):
    from sonet_b200 import layers, synth
    net = layers.PointResNet(cin, [64, 128, 256, 384], 'relu', 'batch')
    net.load_state_dict(synth.synth_state_dict(net, seed=seed))
    return net.eval()


def _ref_chain(net, x):
    """fp64 restatement of PointResNet.forward in eval mode (models/layers.py:419-432)."""
    import torch.nn.functional as F
    def layer(l, t):
        y = F.conv1d(t, l.conv.weight.double(), l.conv.bias.double())
        if l.normalization == 'batch':
            n = l.norm
            y = (y - n.running_mean.double()[None, :, None]) / torch.sqrt(
                n.running_var.double()[None, :, None] + n.eps) * n.weight.double()[None, :, None] \
                + n.bias.double()[None, :, None]
            y = F.relu(y)
        return y
    x = x.double()
    l0 = layer(net.layers[0], x)
    t = layer(net.layers[2], layer(net.layers[1], l0))
    return layer(net.layers[3], torch.cat((l0, t), 1)).float()


@pytest.mark.parametrize("B,cin,P", [(2, 6, 3072), (1, 6, 1000), (3, 3, 77), (1, 6, 128),
                                     (2, 6, 15000), (150, 6, 129)])
def test_pointresnet_tc_vs_fp64_reference(B, cin, P, monkeypatch):
    from helpers import assert_close
    net = _resnet(cin, seed=B + P)
    g = torch.Generator().manual_seed(P)
    x = torch.randn(B, cin, P, generator=g)
    want = _ref_chain(net, x)
    net = net.to(DEV)
    with torch.no_grad():
        monkeypatch.setenv("SONET_TC", "1")
        got = net(x.to(DEV))
        monkeypatch.setenv("SONET_TC", "0")
        simt = net(x.to(DEV))
    assert got.shape == (B, 384, P)
    assert_close(simt, want, "fp32 CUDA-core path vs fp64")
    assert_close(got, want, "tcgen05 bf16x3 path vs fp64")
    # the 3-product split should sit around 1e-5, well inside the 1e-4 bar
    e = float(((got.cpu() - want).abs() / want.abs().clamp(min=1)).max())
    assert e < 5e-5, e


def test_pointresnet_tc_repacks_after_weight_update(monkeypatch):
    monkeypatch.setenv("SONET_TC", "1")
    net = _resnet(6, seed=3).to(DEV)
    x = torch.randn(1, 6, 256, device=DEV)
    with torch.no_grad():
        a = net(x)
        net.layers[3].conv.bias.add_(1.0)
        b = net(x)
    assert torch.allclose(b, a + 1.0, atol=1e-5)
