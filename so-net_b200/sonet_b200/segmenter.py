"""segmenter.Model — caller-facing wrapper of models/segmenter.py:14-192 on the B200 networks.

forward() hands the Segmenter the node-level features and the point->node assignment directly
(Segmenter.forward_nodes) instead of the three per-point torch.gather copies of
models/segmenter.py:90-98.
"""
import torch

from . import networks


class Model():
    def __init__(self, opt):
        self.opt = opt
        dev = opt.device
        self.encoder = networks.Encoder(opt).to(dev)
        self.segmenter = networks.Segmenter(opt).to(dev)
        self.encoder.fuse_pool = False   # the head reads first_pn_out of every forward

        B, N, M = opt.batch_size, opt.input_pc_num, opt.node_num
        self.input_pc = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.input_sn = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.input_label = torch.ones(B, dtype=torch.int64, device=dev)
        self.input_seg = torch.ones(B, 50, dtype=torch.int64, device=dev)
        self.input_node = torch.empty(B, 3, M, dtype=torch.float32, device=dev)
        self.input_node_knn_I = torch.zeros(B, M, opt.som_k, dtype=torch.int64, device=dev)

    def set_input(self, input_pc, input_sn, input_label, input_seg, input_node, input_node_knn_I):
        self.input_pc.resize_(input_pc.size()).copy_(input_pc, non_blocking=True)
        self.input_sn.resize_(input_sn.size()).copy_(input_sn, non_blocking=True)
        self.input_label.resize_(input_label.size()).copy_(input_label, non_blocking=True)
        self.input_seg.resize_(input_seg.size()).copy_(input_seg, non_blocking=True)
        self.input_node.resize_(input_node.size()).copy_(input_node, non_blocking=True)
        self.input_node_knn_I.resize_(input_node_knn_I.size()).copy_(input_node_knn_I,
                                                                     non_blocking=True)
        self.pc = self.input_pc.detach()
        self.sn = self.input_sn.detach()
        self.seg = self.input_seg.detach()
        self.label = self.input_label.detach()

    def forward(self, is_train=False, epoch=None):
        enc = self.encoder
        self.feature = enc(self.pc, self.sn, self.input_node, self.input_node_knn_I, is_train, epoch)
        self.score_segmenter = self.segmenter.forward_nodes(
            enc.x_decentered, self.pc, enc.centers, self.sn, self.input_label, enc.first_pn_out,
            enc.first_pn_out_masked_max, enc.knn_feature_1, enc.final_pn_out, self.feature,
            enc.min_idx)

    def test_model(self):
        self.encoder.eval()
        self.segmenter.eval()
        with torch.no_grad():
            self.forward(is_train=False)
