"""classifier.Model — the caller-facing wrapper of models/classifier.py:15-153 (set_input /
forward / test_model / optimize), built on the B200 networks. This is the public API bench.py
times end to end: host tensors in through set_input, scores out.
"""
import torch
import torch.nn as nn

from . import networks


class Model():
    def __init__(self, opt):
        self.opt = opt
        dev = opt.device
        self.encoder = networks.Encoder(opt).to(dev)
        self.classifier = networks.Classifier(opt).to(dev)
        self.softmax_criteria = nn.CrossEntropyLoss().to(dev)
        self._optim = None

        B, N, M = opt.batch_size, opt.input_pc_num, opt.node_num
        self.input_pc = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.input_sn = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.input_label = torch.ones(B, dtype=torch.int64, device=dev)
        self.input_node = torch.empty(B, 3, M, dtype=torch.float32, device=dev)
        self.input_node_knn_I = torch.zeros(B, M, opt.som_k, dtype=torch.int64, device=dev)
        self.test_loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.test_accuracy = torch.zeros(1, dtype=torch.float32)

    def set_input(self, input_pc, input_sn, input_label, input_node, input_node_knn_I):
        """Copy one batch (host or device tensors) into the pre-allocated device buffers
        (models/classifier.py:64-72); pinned host tensors are copied asynchronously."""
        self.input_pc.resize_(input_pc.size()).copy_(input_pc, non_blocking=True)
        self.input_sn.resize_(input_sn.size()).copy_(input_sn, non_blocking=True)
        self.input_label.resize_(input_label.size()).copy_(input_label, non_blocking=True)
        self.input_node.resize_(input_node.size()).copy_(input_node, non_blocking=True)
        self.input_node_knn_I.resize_(input_node_knn_I.size()).copy_(input_node_knn_I,
                                                                     non_blocking=True)
        self.pc = self.input_pc.detach()
        self.sn = self.input_sn.detach()
        self.label = self.input_label.detach()

    def forward(self, is_train=False, epoch=None):
        self.feature = self.encoder(self.pc, self.sn, self.input_node, self.input_node_knn_I,
                                    is_train, epoch)
        self.score = self.classifier(self.feature, epoch)

    def test_model(self):
        self.encoder.eval()
        self.classifier.eval()
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
