"""-m gpu: on-device batch augmentation + batched SOM-node kNN (SURVEY.md §8f-3,
csrc/augment.cu) against the fixture produced by the reference's own data/augmentation.py
functions, and the oracle's kNN restatement (models/layers.py:334-337)."""
import argparse

import numpy as np
import pytest
import torch

from helpers import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ulp_equal(a, b, what):
    """Bit-exact is the expectation (fp64 pipeline, one fp32 rounding). numpy's BLAS may fuse the
    3-term rotation products, which can move the fp64 value by 1 ulp and — rarely — the fp32
    rounding: tolerate <= 1 fp32 ulp on < 0.01 % of the elements."""
    a, b = a.cpu().numpy(), np.asarray(b)
    neq = a != b
    assert neq.mean() <= 1e-4, "%s: %.4f%% elements differ" % (what, 100 * neq.mean())
    if neq.any():
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, "%s: %d ulp" % (what, ulp.max())


@pytest.mark.parametrize("tag,flags", [("all", (True, True, True)), ("plain", (False, False, False))])
def test_augment_reproduces_reference_pipeline(tag, flags):
    from sonet_b200 import augmentation as A
    g = golden("augment")
    pc, sn, som = (torch.from_numpy(g[n]).to(DEV) for n in ("pc", "sn", "som"))
    B, _, N = pc.shape
    p = A.draw_params(np.random.RandomState(int(g["np_seed"])), B, N, som.shape[2], *flags,
                      host_noise=True)
    pc2, sn2, som2 = A.augment_batch(pc, sn, som, p)
    _ulp_equal(pc2, g[tag + "_pc"], "pc")
    _ulp_equal(sn2, g[tag + "_sn"], "sn")
    _ulp_equal(som2, g[tag + "_som"], "som")


def test_device_noise_statistics_and_determinism():
    from sonet_b200 import ops
    B, N, M = 4, 50000, 64
    pc = torch.zeros(B, 3, N, device=DEV)
    sn = torch.zeros(B, 3, N, device=DEV)
    som = torch.zeros(B, 3, M, device=DEV)
    kw = dict(jitter_pc=(0.01, 0.05), jitter_sn=(0.01, 0.025), jitter_som=(0.04, 0.1))
    a = ops.augment(pc, sn, som, seed=7, **kw)
    b = ops.augment(pc, sn, som, seed=7, **kw)
    c = ops.augment(pc, sn, som, seed=8, **kw)
    assert all(torch.equal(x, y) for x, y in zip(a, b))           # keyed by (seed, cloud, array, point)
    assert not torch.equal(a[0], c[0])
    j = a[0].double()
    assert abs(float(j.mean())) < 2e-4 and abs(float(j.std()) - 0.01) < 2e-4
    assert float(j.abs().max()) <= float(np.float32(0.05))
    assert abs(float((j[0] * j[1]).mean())) < 1e-6                 # clouds are independent
    assert abs(float((j * a[1].double()).mean())) < 1e-6           # arrays are independent
    jn = a[1].double()                                             # clip at 2.5 sigma is visible
    assert float(jn.abs().max()) == pytest.approx(0.025, abs=1e-8)
    frac = float((jn.abs() >= 0.025 - 1e-8).double().mean())
    assert abs(frac - 0.01242) < 2e-3                              # P(|g| > 2.5)
    # normality of the unclipped bulk: kurtosis of N(0,1) is 3
    z = j / 0.01
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05


def test_prepare_batch_and_som_knn(oracle_mod):
    from sonet_b200 import augmentation as A
    rs = np.random.RandomState(3)
    B, N, M = 5, 1000, 64
    pc = torch.from_numpy(rs.uniform(-1, 1, size=(B, 3, N)).astype(np.float32)).to(DEV)
    sn = torch.from_numpy(rs.normal(size=(B, 3, N)).astype(np.float32)).to(DEV)
    som = torch.from_numpy(rs.uniform(-1, 1, size=(B, 3, M)).astype(np.float32)).to(DEV)
    opt = argparse.Namespace(som_k=9, node_num=M, rot_horizontal=True, rot_perturbation=True,
                             translation_perturbation=True)
    pc2, sn2, som2, knn = A.prepare_batch(pc, sn, som, opt, train=True,
                                          rng=np.random.RandomState(11), seed=5)
    assert knn.dtype == torch.int64 and knn.shape == (B, M, 9)
    assert torch.equal(knn.cpu(), oracle_mod.node_knn(som2.cpu(), 9))    # exact, sorted, self first
    assert torch.equal(knn[:, :, 0].cpu(), torch.arange(M).expand(B, M))
    # augmentation moved everything, norms follow the scale: |R v| = |v| before jitter/scale
    assert not torch.equal(pc2, pc)
    ratio = (sn2.norm(dim=1) / sn.norm(dim=1)).mean(dim=1)
    assert ((ratio > 0.75) & (ratio < 1.25)).all()
    # eval mode: inputs untouched, kNN of the given nodes
    pc3, sn3, som3, knn3 = A.prepare_batch(pc, sn, som, opt, train=False)
    assert pc3 is pc and som3 is som
    assert torch.equal(knn3.cpu(), oracle_mod.node_knn(som.cpu(), 9))
    opt.som_k = 1
    k1 = A.prepare_batch(pc, sn, som, opt, train=False)[3]
    assert k1.shape == (B, M, 1) and torch.equal(k1[0, :, 0].cpu(), torch.arange(M))
