"""Seeded synthetic inputs, options and weights for benchmarks and tests.

No dataset or checkpoint is available offline, so the metric is measured on synthetic clouds of
the named shape and random weights of the named architecture. Everything is generated with
numpy's RandomState (portable across machines) so that the reference run that produced
tests/golden/ and the runs on the GPU box see bit-identical inputs and weights.
"""
import argparse
import math

import numpy as np
import torch


def make_opt(task="classifier", batch_size=8, input_pc_num=1024, device="cpu", gpu_id=0, **over):
    """The `opt` namespace the reference threads through every constructor (*/options.py)."""
    opt = argparse.Namespace(
        gpu_id=gpu_id, device=torch.device(device), batch_size=batch_size,
        input_pc_num=input_pc_num, surface_normal=True, feature_num=1024, activation='relu',
        normalization='batch', node_num=64, k=3, som_k=9, som_k_type='avg', dropout=0.7,
        classes=40, bn_momentum=0.1, bn_momentum_decay_step=None, bn_momentum_decay=0.6,
        lr=0.001, pretrain=None, pretrain_lr_ratio=1, random_pc_dropout_lower_limit=1,
        checkpoints_dir='./checkpoints', output_fc_pc_num=256, output_conv_pc_num=1024)
    if task == "segmenter":
        opt.classes, opt.som_k_type, opt.dropout = 50, 'center', 0.6
    elif task == "autoencoder":
        opt.dropout = 0.5
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def synth_inputs(B, N, M=64, som_k=9, seed=0, node_mode="sampled"):
    """pc ~ U(-1,1) [B,3,N]; sn = unit normals [B,3,N]; node [B,3,M]: 'sampled' = M distinct points
    of each cloud (stand-in for a trained SOM, no empty nodes in practice), 'uniform' = U(-1,1)
    (exercises empty nodes); node_knn_I [B,M,som_k] = exact sorted kNN among nodes, self first
    (what the loader's Faiss search returns, data/modelnet_shrec_loader.py:258-259);
    label ~ U{0..15} (valid for both classifier and the segmenter's 16 categories)."""
    rs = np.random.RandomState(1000 + seed)
    pc = rs.uniform(-1, 1, size=(B, 3, N)).astype(np.float32)
    sn = rs.normal(size=(B, 3, N)).astype(np.float32)
    sn /= np.maximum(np.sqrt((sn * sn).sum(axis=1, keepdims=True)), 1e-12)
    if node_mode == "sampled":
        node = np.stack([pc[b][:, rs.permutation(N)[:M]] if N >= M else
                         rs.uniform(-1, 1, size=(3, M)).astype(np.float32) for b in range(B)])
    elif node_mode == "uniform":
        node = rs.uniform(-1, 1, size=(B, 3, M)).astype(np.float32)
    else:
        raise ValueError(node_mode)
    node = np.ascontiguousarray(node, dtype=np.float32)
    d = ((node[:, :, :, None].astype(np.float64) - node[:, :, None, :]) ** 2).sum(axis=1)  # B,M,M
    knn = np.argsort(d, axis=2, kind="stable")[:, :, :max(som_k, 1)].astype(np.int64)
    label = rs.randint(0, 16, size=(B,)).astype(np.int64)
    return dict(pc=torch.from_numpy(pc), sn=torch.from_numpy(sn.astype(np.float32)),
                node=torch.from_numpy(node), node_knn_I=torch.from_numpy(knn),
                label=torch.from_numpy(label))


def synth_state_dict(module_or_state, seed=0):
    """Deterministic non-trivial weights for every tensor of a state_dict (same keys/shapes as the
    reference's): He-normal conv/linear weights, small biases, and randomised BatchNorm affine and
    running statistics so that eval-mode BN is not an identity."""
    state = module_or_state.state_dict() if hasattr(module_or_state, "state_dict") \
        else module_or_state
    out = {}
    for i, (name, t) in enumerate(state.items()):
        rs = np.random.RandomState((seed * 7919 + i * 104729 + 12345) % (2 ** 31 - 1))
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        owner = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros(shape, dtype=t.dtype)
            continue
        if owner == "norm" or "norm" in owner:
            if leaf == "weight":
                v = rs.uniform(0.5, 1.5, size=shape)
            elif leaf == "bias":
                v = rs.normal(0, 0.1, size=shape)
            elif leaf == "running_mean":
                v = rs.normal(0, 0.1, size=shape)
            elif leaf == "running_var":
                v = rs.uniform(0.5, 1.5, size=shape)
            else:
                v = np.zeros(shape)
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            v = rs.normal(0, math.sqrt(2.0 / max(fan_in, 1)), size=shape)
        elif leaf == "bias":
            v = rs.normal(0, 0.05, size=shape)
        else:
            v = rs.normal(0, 0.1, size=shape)
        out[name] = torch.from_numpy(np.asarray(v, dtype=np.float32)).to(t.dtype)
    return out
