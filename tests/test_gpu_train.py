"""-m gpu: train-mode kernels (csrc/train.cu, sonet_b200/train_ops.py) against
the PyTorch composition the reference trains with (conv1d -> batch_norm(training=True) -> relu,
gather; models/layers.py:22-70, 282-296, models/networks.py:185). Tolerance: |a-b| <= tol * max|b|
per tensor (gradients have no natural unit scale)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, what, tol=1e-4):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = float(b.abs().max().clamp(min=1e-12))
    err = float((a - b).abs().max()) / scale
    assert err <= tol, "%s: max|a-b|/max|b| = %.3e > %.1e" % (what, err, tol)


@pytest.mark.parametrize("B,C,P,relu", [(4, 64, 3000, True), (2, 384, 777, True), (3, 40, 1, False),
                                        (64, 128, 576, True)])
def test_bn_act_train_forward_backward_vs_torch(B, C, P, relu):
    from sonet_b200 import train_ops
    g = torch.Generator().manual_seed(B * C + P)
    x = (torch.randn(B, C, P, generator=g) * 2 + 0.5).to(DEV).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.2).to(DEV).requires_grad_(True)
    w = torch.randn(B, C, P, generator=g).to(DEV)
    y, mean, var = train_ops.BNActTrain.apply(x, gamma, beta, 1e-5, relu)
    (y * w).sum().backward()
    got = [t.grad.clone() for t in (x, gamma, beta)]
    for t in (x, gamma, beta):
        t.grad = None
    yr = F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5)
    yr = F.relu(yr) if relu else yr
    (yr * w).sum().backward()
    close(y, yr, "y", 1e-5)
    close(mean, x.detach().mean(dim=(0, 2)), "mean", 1e-5)
    close(var, x.detach().var(dim=(0, 2), unbiased=False), "var", 1e-5)
    for a, t, n in zip(got, (x, gamma, beta), ("dx", "dgamma", "dbeta")):
        close(a, t.grad, n, 2e-4)
    # bit-reproducible
    y2, _, _ = train_ops.BNActTrain.apply(x.detach(), gamma.detach(), beta.detach(), 1e-5, relu)
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("B,cin,cout,P", [(4, 64, 128, 1500), (2, 320, 384, 1024), (64, 387, 512, 576)])
def test_conv_tc_forward_dgrad_vs_torch(B, cin, cout, P):
    from sonet_b200 import train_ops
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(B, cin, P, generator=g).to(DEV).requires_grad_(True)
    W = (torch.randn(cout, cin, generator=g) * (2.0 / cin) ** 0.5).to(DEV).requires_grad_(True)
    b = (torch.randn(cout, generator=g) * 0.1).to(DEV).requires_grad_(True)
    w = torch.randn(B, cout, P, generator=g).to(DEV)
    y = train_ops.ConvTC.apply(x, W, b)
    (y * w).sum().backward()
    got = [t.grad.clone() for t in (x, W, b)]
    for t in (x, W, b):
        t.grad = None
    yr = F.conv1d(x.double(), W.double().unsqueeze(2), b.double())
    (yr * w.double()).sum().backward()
    close(y, yr, "y")
    for a, t, n in zip(got, (x, W, b), ("dx", "dW", "db")):
        close(a, t.grad, n)


@pytest.mark.parametrize("B,cin,cout,P,gscale", [(4, 64, 128, 1500, 1.0), (64, 387, 512, 576, 1e-6),
                                                 (3, 320, 384, 1000, 3e-5), (2, 40, 64, 333, 1.0)])
def test_wgrad_tc_vs_fp64(B, cin, cout, P, gscale):
    """dW = sum_{b,p} dy x^T on tcgen05 (transpose + pre-scale of dy, x packed as the weight operand,
    split-K with a fixed-order reduce) against fp64; gradients as small as 1e-6 keep full precision
    because they are pre-scaled by a power of two before the fp16 split."""
    from sonet_b200 import train_ops
    g = torch.Generator().manual_seed(B + cin + P)
    x = torch.randn(B, cin, P, generator=g).to(DEV)
    dy = (torch.randn(B, cout, P, generator=g) * gscale).to(DEV)
    got = train_ops.wgrad_tc(dy, x)
    want = torch.einsum("bop,bip->oi", dy.double(), x.double())
    close(got, want, "dW", 1e-4)
    assert torch.equal(train_ops.wgrad_tc(dy, x), got)          # deterministic


def test_dgrad_tiny_gradients_keep_precision():
    from sonet_b200 import train_ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 128, 800, generator=g).to(DEV).requires_grad_(True)
    W = (torch.randn(256, 128, generator=g) * 0.1).to(DEV).requires_grad_(True)
    y = train_ops.ConvTC.apply(x, W, None)
    w = (torch.randn(4, 256, 800, generator=g) * 1e-7).to(DEV)
    (y * w).sum().backward()
    want = torch.einsum("oi,bop->bip", W.detach().double(), w.double())
    close(x.grad, want, "dx with 1e-7 gradients", 1e-4)


def test_index_max_gather_backward_vs_torch_gather(oracle_mod):
    from sonet_b200 import train_ops
    rs = np.random.RandomState(4)
    B, C, N, K = 3, 37, 999, 64
    data = torch.from_numpy(rs.normal(size=(B, C, N)).astype(np.float32)).to(DEV).requires_grad_(True)
    index = torch.from_numpy(rs.randint(0, K, size=(B, N)).astype(np.int32))
    index[index == 5] = 6                                   # empty nodes: they all gather point 0
    index[index == 17] = 18
    index = index.to(DEV)
    w = torch.from_numpy(rs.normal(size=(B, C, K)).astype(np.float32)).to(DEV)
    out = train_ops.IndexMaxGather.apply(data, index, K)
    (out * w).sum().backward()
    got = data.grad.clone()
    data.grad = None
    gi = oracle_mod.index_max(data.detach().cpu(), index.cpu(), K).long().to(DEV)
    ref = data.gather(2, gi)                                # idx is already 0 for empty nodes
    (ref * w).sum().backward()
    assert torch.equal(out.detach(), ref.detach())
    close(got, data.grad, "d data", 1e-6)
    assert float(got[:, :, 0].abs().sum()) > 0              # the empty nodes' gradient landed on point 0


def test_classifier_training_step_kernels_vs_torch_composition():
    """One optimize() step from identical weights with the train kernels ON and OFF: same loss,
    same gradients (1e-3 of each tensor's scale: the tcgen05 GEMMs carry ~1e-5 per layer),
    same running statistics."""
    from helpers import build_states
    from sonet_b200 import classifier, synth, train_ops
    opt = synth.make_opt("classifier", batch_size=8, input_pc_num=512, device=DEV)
    opt.device = torch.device(DEV)
    st = build_states("classifier", opt, seed=81)
    inp = synth.synth_inputs(8, 512, seed=81)
    res = {}
    # the PyTorch composition must be a strict-fp32 baseline: cuDNN/cuBLAS default to TF32
    # (~1e-3), which is coarser than the kernels under test
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for flag in (True, False):
        train_ops.ENABLED = flag
        try:
            m = classifier.Model(opt)
            m.encoder.load_state_dict(st["encoder"])
            m.classifier.load_state_dict(st["head"])
            m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
            m.encoder.train()
            m.classifier.train()
            torch.manual_seed(0)                              # dropout masks
            m.forward(is_train=True, epoch=3)
            loss = m.softmax_criteria(m.score, m.label)
            m.encoder.zero_grad()
            m.classifier.zero_grad()
            loss.backward()
            res[flag] = dict(
                loss=loss.detach().clone(),
                grads={n: p.grad.detach().clone() for n, p in m.encoder.named_parameters()
                       if p.grad is not None},
                rm=m.encoder.first_pointnet.layers[1].norm.running_mean.clone(),
                rv=m.encoder.first_pointnet.layers[1].norm.running_var.clone())
        finally:
            train_ops.ENABLED = True
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    close(res[True]["loss"], res[False]["loss"], "loss", 1e-4)
    close(res[True]["rm"], res[False]["rm"], "running_mean", 1e-5)
    close(res[True]["rv"], res[False]["rv"], "running_var", 1e-5)
    assert set(res[True]["grads"]) == set(res[False]["grads"]) and len(res[True]["grads"]) > 20
    for n, gref in res[False]["grads"].items():
        got = res[True]["grads"][n]
        wname = n.rsplit(".", 1)[0] + ".weight"
        if n.endswith(".bias") and wname in res[False]["grads"]:
            # a bias whose effect is a per-channel constant removed by a LATER batch-stat BN (conv
            # biases in front of a BN; the bare last conv of the first PointResNet, whose output
            # reaches the KNN module's BN through the max-pool) has an analytically zero gradient:
            # both paths return rounding noise, negligible against the layer's weight gradient
            wscale = float(res[False]["grads"][wname].abs().max())
            if float(gref.abs().max()) < 1e-3 * wscale:
                assert float(got.abs().max()) < 1e-3 * wscale, n
                continue
        close(got, gref, "grad " + n, 2e-3)
