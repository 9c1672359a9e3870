"""classifier.Model — the caller-facing wrapper of models/classifier.py:15-153 (set_input /
forward / test_model / optimize), built on the B200 networks. This is the public API bench.py
times end to end: host tensors in through set_input, scores out.

set_input is double-buffered and asynchronous: it copies into the input buffer set that the
in-flight forward is NOT reading, on a dedicated copy stream, and forward() makes the compute
stream wait for exactly that copy. The reference call order (set_input, then forward/test_model)
is unchanged; a serving loop that wants the host-to-device copy hidden simply issues
set_input(batch i+1) before it reads the scores of batch i.
"""
import torch
import torch.nn as nn

from . import networks, ops


class _InputSet:
    def __init__(self, B, N, M, som_k, dev):
        self.pc = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.sn = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.label = torch.ones(B, dtype=torch.int64, device=dev)
        self.node = torch.empty(B, 3, M, dtype=torch.float32, device=dev)
        self.node_knn_I = torch.zeros(B, M, max(som_k, 1), dtype=torch.int64, device=dev)
        self.ready = torch.cuda.Event() if dev.type == "cuda" else None      # copy finished
        self.consumed = torch.cuda.Event() if dev.type == "cuda" else None   # last forward finished


class Model():
    def __init__(self, opt):
        self.opt = opt
        dev = opt.device if isinstance(opt.device, torch.device) else torch.device(opt.device)
        self.encoder = networks.Encoder(opt).to(dev)
        self.classifier = networks.Classifier(opt).to(dev)
        self.softmax_criteria = nn.CrossEntropyLoss().to(dev)
        self._optim = None
        self._dev = dev
        B, N, M = opt.batch_size, opt.input_pc_num, opt.node_num
        self._sets = [_InputSet(B, N, M, opt.som_k, dev) for _ in range(2)]
        self._cur = 0
        self._copy_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._bind(self._sets[0])
        self.test_loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.test_accuracy = torch.zeros(1, dtype=torch.float32)
        self._use_graph = False
        self._graphs = {}
        self._graph_stream = None

    # ---- CUDA-graph replay of the eval forward ---------------------------------------------------
    def enable_cuda_graph(self, flag=True):
        """Replay test_model() as one CUDA graph per input buffer set (every C-ABI entry is
        capturable: no allocation, no synchronisation). ~16 Python/ctypes op calls per step become
        one graph launch, which is what bounds the end-to-end rate at this step time. Outputs
        (score, feature, loss, encoder attributes) then live in per-graph static buffers that the
        next replay of the same buffer set overwrites; graphs are re-captured when an input shape or
        any parameter/buffer version changes."""
        self._use_graph = bool(flag) and self._dev.type == "cuda"
        self._graphs = {}
        if self._use_graph and self._graph_stream is None:
            self._graph_stream = torch.cuda.Stream(device=self._dev)

    def invalidate(self):
        """Forget folded/packed weights and captured graphs (call after editing parameters
        through `.data`, which does not bump tensor versions — see layers.invalidate)."""
        from . import layers
        layers.invalidate(self.encoder)
        layers.invalidate(self.classifier)
        self._graphs = {}

    def _state_key(self, s):
        if not hasattr(self, "_state_tensors"):
            self._state_tensors = [t for m in (self.encoder, self.classifier)
                                   for t in list(m.parameters()) + list(m.buffers())]
        ver = 0
        for t in self._state_tensors:
            ver += t._version
        return (tuple(s.pc.shape), tuple(s.node.shape), tuple(s.node_knn_I.shape), ver,
                self.encoder.fuse_pool)

    def _eval_forward(self):
        self.feature = self.encoder(self.pc, self.sn, self.input_node, self.input_node_knn_I,
                                    False, None)
        self.score = self.classifier(self.feature, None)
        self.loss = self.softmax_criteria(self.score, self.label)

    # encoder attributes that other code reads after a forward (models/segmenter.py:90-108 style)
    _ENC_PUBLIC = ("som_node", "first_pn_out_masked_max", "knn_center_1", "knn_feature_1",
                   "final_pn_out", "feature", "_assign", "_lazy_src")
    _ENC_LAZY = ("_mask", "_centers", "_x_aug", "_first_pn_out")

    def _test_model_graph(self):
        s = self._sets[self._cur]
        cur = torch.cuda.current_stream(self._dev)
        cur.wait_event(s.ready)
        key = self._state_key(s)
        g = self._graphs.get(self._cur)
        enc = self.encoder
        with torch.no_grad():
            if g is None or g["key"] != key:
                # warm-up AND capture run on the model's capture stream, so per-stream scratch
                # (the pool keys) is created and left clean before the capture starts
                gs = self._graph_stream
                gs.wait_stream(cur)
                with torch.cuda.stream(gs):
                    for _ in range(2):           # warm every host-side cache (folded/packed weights)
                        self._eval_forward()
                torch.cuda.synchronize(self._dev)
                graph = torch.cuda.CUDAGraph()
                k0, c0 = ops.KERNEL_LAUNCHES, ops.LAUNCHES
                with torch.cuda.graph(graph, stream=gs):
                    self._eval_forward()
                g = dict(key=key, graph=graph, kernels=ops.KERNEL_LAUNCHES - k0,
                         calls=ops.LAUNCHES - c0, feature=self.feature, score=self.score,
                         loss=self.loss,
                         enc={n: getattr(enc, n) for n in self._ENC_PUBLIC if hasattr(enc, n)})
                self._graphs[self._cur] = g
            g["graph"].replay()
        ops.KERNEL_LAUNCHES += g["kernels"]      # the replay launches the captured kernels
        ops.LAUNCHES += g["calls"]
        self.feature, self.score, self.loss = g["feature"], g["score"], g["loss"]
        # the encoder's cached attributes point at THIS graph's static buffers again (two input
        # sets = two graphs = two buffer sets); lazily derived ones are recomputed on demand
        for n, v in g["enc"].items():
            setattr(enc, n, v)
        for n in self._ENC_LAZY:
            setattr(enc, n, None)
        s.consumed.record(cur)

    def _bind(self, s):
        self.input_pc, self.input_sn, self.input_label = s.pc, s.sn, s.label
        self.input_node, self.input_node_knn_I = s.node, s.node_knn_I
        self.pc, self.sn, self.label = s.pc.detach(), s.sn.detach(), s.label.detach()

    def set_input(self, input_pc, input_sn, input_label, input_node, input_node_knn_I):
        """Copy one batch (host or device tensors) into the idle device buffer set
        (models/classifier.py:64-72). Pinned host tensors are copied asynchronously on the copy
        stream; the next forward waits for this copy only."""
        self._cur ^= 1
        s = self._sets[self._cur]
        srcs = (input_pc, input_sn, input_label, input_node, input_node_knn_I)
        dsts = (s.pc, s.sn, s.label, s.node, s.node_knn_I)
        if self._copy_stream is None:
            for d, t in zip(dsts, srcs):
                d.resize_(t.size()).copy_(t)
        else:
            cur = torch.cuda.current_stream(self._dev)
            cs = self._copy_stream
            cs.wait_event(s.consumed)        # the forward that last read this set is done
            if any(t.is_cuda for t in srcs):
                cs.wait_stream(cur)          # device-side sources produced on the caller's stream
            with torch.cuda.stream(cs):
                for d, t in zip(dsts, srcs):
                    if d.size() != t.size():
                        d.resize_(t.size())
                    d.copy_(t, non_blocking=True)
                    if t.is_cuda:
                        t.record_stream(cs)
                s.ready.record(cs)
        self._bind(s)

    def forward(self, is_train=False, epoch=None):
        s = self._sets[self._cur]
        if s.ready is not None:
            torch.cuda.current_stream(self._dev).wait_event(s.ready)
        self.feature = self.encoder(self.pc, self.sn, self.input_node, self.input_node_knn_I,
                                    is_train, epoch)
        self.score = self.classifier(self.feature, epoch)
        if s.consumed is not None:
            s.consumed.record(torch.cuda.current_stream(self._dev))

    def test_model(self):
        self.encoder.eval()
        self.classifier.eval()
        if self._use_graph:
            return self._test_model_graph()
        with torch.no_grad():
            self.forward(is_train=False)
            self.loss = self.softmax_criteria(self.score, self.label)

    def optimize(self, epoch=None):
        """One training step (models/classifier.py:78-99) on the differentiable PyTorch path."""
        if self._optim is None:
            self._optim = (torch.optim.Adam(self.encoder.parameters(), lr=self.opt.lr,
                                            betas=(0.9, 0.999), weight_decay=0),
                           torch.optim.Adam(self.classifier.parameters(), lr=self.opt.lr,
                                            betas=(0.9, 0.999), weight_decay=0))
        self.encoder.train()
        self.classifier.train()
        with torch.enable_grad():
            self.forward(is_train=True, epoch=epoch)
            self.encoder.zero_grad()
            self.classifier.zero_grad()
            self.loss = self.softmax_criteria(self.score, self.label)
            self.loss.backward()
        for o in self._optim:
            o.step()
