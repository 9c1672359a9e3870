"""ref_shims.py — TEST INFRASTRUCTURE. Import the UNMODIFIED reference on CPU.

Source of the reference modules: /root/reference when it exists (the build container), else the
bytecode build product oracle/_ref/pyref (sourceless .pyc compiled from those sources by
oracle/build.py:build_pyref — what travels to the GPU box). Three shims, as established in
SURVEY.md §0 / Appendix B:
  1. empty stub modules for matplotlib / mpl_toolkits (imported but unused, models/networks.py:14-15);
  2. a `faiss` stub whose IndexFlatL2 is an exact brute-force search (only ChamferLoss uses it);
  3. module `index_max` = the reference's own plugin compiled in place (oracle/_ref), with
     forward_cuda redirected to the reference's forward_cpu for CPU tensors.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"


def ref_root(prefer_pyref=False):
    """Where the reference's `models` / `util` packages are importable from, or None."""
    from . import build as _b
    have_src = os.path.isdir(os.path.join(REF, "models"))
    if have_src and not prefer_pyref:
        return REF
    return _b.pyref_root() or (REF if have_src else None)


def available():
    return ref_root() is not None


class _IndexFlatL2:
    def __init__(self, d):
        self.d = d
        self.db = None

    def add(self, x):
        self.db = np.ascontiguousarray(x, dtype=np.float32)

    def search(self, q, k):
        q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
        db = torch.from_numpy(self.db)
        d = ((q[:, None, :] - db[None, :, :]) ** 2).sum(dim=2)
        D, I = torch.topk(d, k=k, dim=1, largest=False, sorted=True)
        return D.numpy(), I.numpy()


class cpu_only:
    """Context manager for constructing reference objects for the CPU arm on a machine that HAS a
    GPU: util/som.py:188 picks `cuda:%d if torch.cuda.is_available() else cpu` for the SOM nodes
    regardless of opt.device, which mixes devices when the model itself is on the CPU. Inside the
    context torch.cuda.is_available() answers False (shim 4: an environment answer, the reference
    code is untouched)."""

    def __enter__(self):
        self._orig = torch.cuda.is_available
        torch.cuda.is_available = lambda: False
        return self

    def __exit__(self, *exc):
        torch.cuda.is_available = self._orig
        return False


def install(use_ref_plugin=True, prefer_pyref=False, pool_threads=None):
    """Register the shims and put the reference on sys.path. Returns the reference's modules.
    pool_threads: when set, CPU tensors handed to index_max.forward_cuda go to the reference's
    forward_multi_thread_cpu with that many threads (its faster CPU variant, index_max.cpp:33-70)
    instead of forward_cpu — used by the timed CPU arm only."""
    root = ref_root(prefer_pyref)
    if root is None:
        raise RuntimeError("neither /root/reference nor oracle/_ref/pyref is present on this machine")
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    faiss = types.ModuleType("faiss")

    class _Res:
        def setTempMemoryFraction(self, f):
            pass

    class _Cfg:
        device = 0

    faiss.StandardGpuResources = _Res
    faiss.GpuIndexFlatConfig = _Cfg
    faiss.IndexFlatL2 = _IndexFlatL2
    faiss.index_cpu_to_gpu = lambda res, dev, idx: idx
    sys.modules["faiss"] = faiss

    from . import oracle
    plugin = oracle.ref_plugin() if use_ref_plugin else None
    shim = types.ModuleType("index_max")
    if plugin is not None:
        shim.forward_cpu = plugin.forward_cpu
        shim.forward_multi_thread_cpu = plugin.forward_multi_thread_cpu
        shim.pool_threads = pool_threads      # mutable: the timed CPU arm probes thread counts

        def _fwd(data, index, K):
            if shim.pool_threads:
                return plugin.forward_multi_thread_cpu(
                    data, index, K, max(1, min(int(shim.pool_threads), data.shape[1])))
            return plugin.forward_cpu(data, index, K)
        shim.forward_cuda = _fwd
        shim.forward_cuda_shared_mem = _fwd
        shim.is_reference_binary = True
    else:
        shim.forward_cpu = oracle.index_max
        shim.forward_cuda = oracle.index_max
        shim.forward_cuda_shared_mem = oracle.index_max
        shim.is_reference_binary = False
    sys.modules["index_max"] = shim

    if root not in sys.path:
        sys.path.insert(0, root)
    import models.autoencoder as ref_autoencoder
    import models.classifier as ref_classifier
    import models.losses as ref_losses
    import models.networks as ref_networks
    import models.segmenter as ref_segmenter
    import util.som as ref_som
    return types.SimpleNamespace(classifier=ref_classifier, segmenter=ref_segmenter,
                                 autoencoder=ref_autoencoder, networks=ref_networks,
                                 losses=ref_losses, som=ref_som, index_max=shim, root=root)
