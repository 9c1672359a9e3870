"""autoencoder.Model — caller-facing wrapper of models/autoencoder.py:13-161 on the B200
networks: encoder (kernels) + decoder (PyTorch/cuDNN) + Chamfer (kernels)."""
import torch

from . import losses, networks


class Model():
    def __init__(self, opt):
        self.opt = opt
        dev = opt.device
        self.encoder = networks.Encoder(opt).to(dev)
        self.decoder = networks.Decoder(opt).to(dev)
        self.chamfer_criteria = losses.ChamferLoss(opt).to(dev)

        B, N, M = opt.batch_size, opt.input_pc_num, opt.node_num
        self.input_pc = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.input_sn = torch.empty(B, 3, N, dtype=torch.float32, device=dev)
        self.input_label = torch.ones(B, dtype=torch.int64, device=dev)
        self.input_node = torch.empty(B, 3, M, dtype=torch.float32, device=dev)
        self.input_node_knn_I = torch.zeros(B, M, opt.som_k, dtype=torch.int64, device=dev)
        self.test_loss = torch.zeros(1, dtype=torch.float32, device=dev)

    def set_input(self, input_pc, input_sn, input_label, input_node, input_node_knn_I):
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
        self.predicted_pc = self.decoder(self.feature)

    def test_model(self):
        """models/autoencoder.py:105-126."""
        self.encoder.eval()
        self.decoder.eval()
        with torch.no_grad():
            self.forward(is_train=False)
            n_conv = self.opt.output_conv_pc_num
            if n_conv > 0:
                if n_conv == 4096:
                    self.loss_chamfer_conv5 = self.chamfer_criteria(self.decoder.conv_pc5, self.pc)
                self.loss_chamfer_conv4 = self.chamfer_criteria(self.decoder.conv_pc4, self.pc)
            self.loss_chamfer = self.chamfer_criteria(self.predicted_pc, self.pc)
            if n_conv == 1024:
                self.loss = self.loss_chamfer + self.loss_chamfer_conv4
            elif n_conv == 4096:
                self.loss = self.loss_chamfer + self.loss_chamfer_conv5 + self.loss_chamfer_conv4
            elif n_conv == 0:
                self.loss = self.loss_chamfer
