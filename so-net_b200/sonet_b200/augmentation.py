"""Batched, on-device counterpart of the loader-side preprocessing that feeds the hot path
(SURVEY.md §8f-3): data/augmentation.py:52-144 as applied by
data/modelnet_shrec_loader.py:218-261 — rotation about the up axis, perturbation rotation, jitter
of points / normals / SOM nodes, random scale, random shift — and the SOM-node kNN that the loader
computes per item with a CPU Faiss index (loader :116-150, 256-259).

The reference does this per item in DataLoader workers (six numpy passes + a Faiss index build per
cloud); `prepare_batch` does it for the whole batch in two kernel launches on the device that
already holds the batch: csrc/augment.cu and sonet_node_knn (csrc/pointwise.cu).

Randomness. The per-cloud scalars (angles, scale, shift: 8 numbers per cloud) are drawn on the host
from a numpy RandomState with the loader's own expressions, in the loader's order. The per-point
jitter draws are generated inside the kernel (Philox4x32-10, keyed by seed/cloud/array/point) —
or, with `host_noise=True`, drawn from the same RandomState in the loader's order and shipped to
the device, which reproduces the reference pipeline bit for bit (what the parity tests use).
"""
import numpy as np
import torch

from . import ops

JITTER_PC = (0.01, 0.05)      # jitter_point_cloud defaults, augmentation.py:132 (points, normals)
JITTER_SOM = (0.04, 0.1)      # loader :233


def rotation_matrix_up_axis(angle):
    """augmentation.py:62-68: rotation about the y axis; data . R."""
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def rotation_matrix_perturbation(angles):
    """augmentation.py:112-122: R = Rz . (Ry . Rx) for three small angles."""
    a = angles
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    return np.dot(Rz, np.dot(Ry, Rx))


def draw_params(rng, B, N, M, rot_horizontal=False, rot_perturbation=False,
                translation_perturbation=False, host_noise=False, angle_sigma=0.06,
                angle_clip=0.18):
    """Draw one batch of augmentation parameters from `rng` (np.random.RandomState or the
    np.random module) cloud by cloud in the loader's call order (loader :224-247):
    [angle] [3 perturbation angles] [randn(N,3) points] [randn(N,3) normals] [randn(M,3) nodes]
    scale [shift]. Returns host float64 arrays (None where the option is off)."""
    rot1 = np.empty((B, 3, 3)) if rot_horizontal else None
    rot2 = np.empty((B, 3, 3)) if rot_perturbation else None
    scale = np.empty((B,))
    shift = np.empty((B, 3)) if translation_perturbation else None
    noise = [np.empty((B, N, 3)), np.empty((B, N, 3)), np.empty((B, M, 3))] if host_noise else None
    for b in range(B):
        if rot_horizontal:
            rot1[b] = rotation_matrix_up_axis(rng.uniform() * 2 * np.pi)          # :62
        if rot_perturbation:
            rot2[b] = rotation_matrix_perturbation(
                np.clip(angle_sigma * rng.randn(3), -angle_clip, angle_clip))      # :112
        if host_noise:
            noise[0][b] = rng.randn(N, 3)                                         # :141
            noise[1][b] = rng.randn(N, 3)
            noise[2][b] = rng.randn(M, 3)
        scale[b] = rng.uniform(low=0.8, high=1.2)                                 # loader :236
        if translation_perturbation:
            shift[b] = rng.uniform(-0.1, 0.1, (1, 3))[0]                          # loader :243
    return dict(rot1=rot1, rot2=rot2, scale=scale, shift=shift, noise=noise)


def _dev(a, device):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(
        device, non_blocking=True)


def augment_batch(pc, sn, som, params, seed=0):
    """Apply `params` (from draw_params) to device tensors pc, sn [B,3,N], som [B,3,M]."""
    if not pc.is_cuda:
        raise RuntimeError("sonet_b200.augmentation runs on CUDA tensors only (no CPU fallback)")
    d = pc.device
    noise = params.get("noise") or (None, None, None)
    return ops.augment(pc.contiguous(), None if sn is None else sn.contiguous(),
                       None if som is None else som.contiguous(),
                       rot1=_dev(params.get("rot1"), d), rot2=_dev(params.get("rot2"), d),
                       scale=_dev(params.get("scale"), d), shift=_dev(params.get("shift"), d),
                       jitter_pc=JITTER_PC, jitter_sn=JITTER_PC, jitter_som=JITTER_SOM,
                       noise_pc=_dev(noise[0], d), noise_sn=_dev(noise[1], d),
                       noise_som=_dev(noise[2], d), seed=seed)


def som_knn(som_node, som_k, node_num=None):
    """Batched `som_knn_I` producer: exact kNN among the SOM nodes of every cloud, ascending
    distance, self first — what KNNBuilder.self_build_search returns per item through a CPU Faiss
    IndexFlatL2 (loader :116-150, 256-259). som_node [B,3,M] -> int64 [B,M,som_k]. For
    som_k < 2 the loader emits arange(node_num) as a [M,1] column (:260-261)."""
    B, _, M = som_node.shape
    if som_k >= 2:
        return ops.node_knn(som_node.detach().contiguous(), som_k)
    return torch.arange(M, dtype=torch.int64, device=som_node.device).view(1, M, 1).expand(
        B, M, 1).contiguous()


def prepare_batch(pc, sn, som, opt, train, rng=None, seed=0, host_noise=False):
    """The loader's per-item tail (loader :218-261) for a whole batch on the device:
    augmentation (train mode only) + SOM-node kNN. Returns (pc, sn, som, som_knn_I) ready for
    Model.set_input."""
    if train:
        rng = rng if rng is not None else np.random
        B, _, N = pc.shape
        p = draw_params(rng, B, N, som.shape[2],
                        rot_horizontal=bool(getattr(opt, "rot_horizontal", False)),
                        rot_perturbation=bool(getattr(opt, "rot_perturbation", False)),
                        translation_perturbation=bool(getattr(opt, "translation_perturbation", False)),
                        host_noise=host_noise)
        pc, sn, som = augment_batch(pc, sn, som, p, seed=seed)
    return pc, sn, som, som_knn(som, opt.som_k, opt.node_num)
