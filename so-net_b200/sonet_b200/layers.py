"""Layer classes with the reference's constructor signatures and state_dict keys
(models/layers.py:22-432), running on the sm_100a kernels of libsonet_b200 in eval/no-grad mode.

Dispatch rule (SURVEY.md §8b "Autograd"): the hand-written forward kernels run whenever the module
is in eval() mode and no INPUT tensor carries a gradient — with or without torch.no_grad(), because
the reference's own test loop (models/classifier.py:101-105, modelnet/train.py:72-76) calls
test_model() in eval mode WITHOUT no_grad and must still land on the hot path. An eval-mode forward
therefore returns tensors without grad_fn: calling backward() on them fails loudly ("does not
require grad") instead of silently dropping parameter gradients. In train() mode, when an input
requires grad (saliency / adversarial use in eval mode), or when EVAL_AUTOGRAD is set, the layers
compose differentiable PyTorch ops on the GPU exactly as the reference does — the reference's own
training path, not a CPU fallback. Non-default options ('instance' norm, elu/swish/leakyrelu)
always take the PyTorch path.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.batchnorm import _BatchNorm

from . import operations, ops, train_ops


# Set to True to make eval-mode forwards differentiable w.r.t. the parameters as well (the
# PyTorch composition); the default keeps eval() = inference kernels.
EVAL_AUTOGRAD = False


def _fast_ok(module, *tensors):
    if module.training:
        return False
    if torch.is_grad_enabled() and (EVAL_AUTOGRAD or any(t is not None and t.requires_grad
                                                         for t in tensors)):
        return False
    return all(t is not None and t.is_cuda and t.dtype == torch.float32 for t in tensors)


class Swish(nn.Module):
    """x * sigmoid(x) (models/layers.py:14-19)."""

    def forward(self, x):
        return x * torch.sigmoid(x)


class _MomentumDecayBatchNorm(_BatchNorm):
    """BatchNorm whose momentum decays with the epoch (models/layers.py:22-70, 73-120)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True,
                 momentum_decay_step=None, momentum_decay=1):
        super().__init__(num_features, eps, momentum, affine)
        self.momentum_decay_step = momentum_decay_step
        self.momentum_decay = momentum_decay
        self.momentum_original = self.momentum

    def forward(self, input, epoch=None):
        step = self.momentum_decay_step
        if epoch is not None and epoch >= 1 and step is not None and step > 0:
            self.momentum = max(self.momentum_original * (self.momentum_decay ** (epoch // step)),
                                0.01)
        return F.batch_norm(input, self.running_mean, self.running_var, self.weight, self.bias,
                            self.training, self.momentum, self.eps)


class MyBatchNorm1d(_MomentumDecayBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() not in (2, 3):
            raise ValueError('expected 2D or 3D input (got {}D input)'.format(input.dim()))


class MyBatchNorm2d(_MomentumDecayBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))


def _make_act(activation):
    if activation == 'relu':
        return nn.ReLU()
    if activation == 'elu':
        return nn.ELU(alpha=1.0)
    if activation == 'swish':
        return Swish()
    if activation == 'leakyrelu':
        return nn.LeakyReLU(0.1)
    return None


class _FoldedParams:
    """Eval-mode conv/linear + BatchNorm folded into (weights, shift), re-packed only when a
    parameter or running statistic changes (tracked through tensor version counters).

      y = gamma * (W x + bias - mean) / sqrt(var + eps) + beta  =  (s*W) x + ((bias-mean)*s + beta)
    """

    def __init__(self):
        self._key = None
        self.version = 0    # bumped on every re-fold (lets dependants cache derived packings)
        self.w = None       # [Cout, Cin] with the BN scale folded in
        self.wt = None      # [Cin, Cout] (the point-wise kernel wants K-major slabs)
        self.shift = None   # [Cout]

    def get(self, weight2d, bias, norm, transpose):
        srcs = [weight2d, bias]
        if norm is not None:
            srcs += [norm.weight, norm.bias, norm.running_mean, norm.running_var]
        key = tuple((t.data_ptr(), t._version, t.device) for t in srcs if t is not None)
        if key != self._key:
            with torch.no_grad():
                w = weight2d.detach().float()
                shift = bias.detach().float() if bias is not None else \
                    torch.zeros(w.shape[0], device=w.device)
                if norm is not None:
                    s = norm.weight.detach() / torch.sqrt(norm.running_var.detach() + norm.eps)
                    shift = (shift - norm.running_mean.detach()) * s + norm.bias.detach()
                    w = w * s[:, None]
                self.w = w.contiguous()          # [Cout, Cin]
                self.wt = None                   # [Cin, Cout], built on first use
                self.shift = shift.contiguous()
            self._key = key
            self.version += 1
        if not transpose:
            return self.w, self.shift
        if self.wt is None:
            self.wt = self.w.t().contiguous()
        return self.wt, self.shift


def invalidate(module):
    """Drop every derived weight cache (BN-folded weights, tcgen05 packings) below `module`.

    The caches are keyed on (data_ptr, tensor._version) of the parameters and BN statistics, which
    optimizer steps, load_state_dict() and in-place tensor ops all bump. Writes through `.data`
    (`p.data.mul_(..)`, `w.data.copy_(..)` — an idiom of the reference code base,
    e.g. models/networks.py:366) do NOT bump the version counter: call this after such edits."""
    for m in module.modules():
        f = getattr(m, "_folded", None)
        if isinstance(f, _FoldedParams):
            f._key = None
        for attr in ("_tc_key", "_l1_key"):
            if hasattr(m, attr):
                setattr(m, attr, None)


class MyLinear(nn.Module):
    """Linear + BN1d + act (models/layers.py:123-166)."""

    def __init__(self, in_features, out_features, activation=None, normalization=None,
                 momentum=0.1, bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        self.activation = activation
        self.normalization = normalization
        self.linear = nn.Linear(in_features, out_features, bias=True)
        if normalization == 'batch':
            self.norm = MyBatchNorm1d(out_features, momentum=momentum, affine=True,
                                      momentum_decay_step=bn_momentum_decay_step,
                                      momentum_decay=bn_momentum_decay)
        elif normalization == 'instance':
            self.norm = nn.InstanceNorm1d(out_features, momentum=momentum, affine=True)
        act = _make_act(activation)
        if act is not None:
            self.act = act
        self._folded = _FoldedParams()
        self.weight_init()

    def weight_init(self):
        nn.init.normal_(self.linear.weight, 0, math.sqrt(2. / self.linear.in_features))
        nn.init.zeros_(self.linear.bias)
        if self.normalization in ('batch', 'instance'):
            nn.init.ones_(self.norm.weight)
            nn.init.zeros_(self.norm.bias)

    def _fast(self, x):
        return (_fast_ok(self, x) and x.dim() == 2
                and self.normalization in (None, 'batch') and self.activation in (None, 'relu'))

    def forward(self, x, epoch=None):
        if self._fast(x):
            w, shift = self._folded.get(self.linear.weight, self.linear.bias,
                                        self.norm if self.normalization == 'batch' else None,
                                        transpose=False)
            return ops.linear(x.contiguous(), w, None, shift, self.activation == 'relu')
        x = self.linear(x)
        if self.normalization == 'batch':
            x = self.norm(x, epoch)
        elif self.normalization is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.act(x)
        return x


class _PointwiseConvBase(nn.Module):
    """Shared eval fast path of the 1x1 conv layers: y = act(BN(conv1x1(cat(x0, x1))))."""

    def _conv_weight2d(self):
        w = self.conv.weight
        return w.view(w.shape[0], w.shape[1])

    def _fast_eligible(self):
        ks = self.conv.kernel_size
        return (all(k == 1 for k in ks) and all(s == 1 for s in self.conv.stride)
                and all(p == 0 for p in self.conv.padding)
                and self.normalization in (None, 'batch') and self.activation in (None, 'relu'))

    def _train_kernels(self, x):
        """train() mode on CUDA fp32 with the default norm/activation: the sm_100a train kernels
        (train_ops) instead of the PyTorch composition."""
        return (train_ops.ENABLED and self.training and torch.is_grad_enabled() and x.is_cuda
                and x.dtype == torch.float32 and self._fast_eligible())

    def forward_points(self, x0, x1=None, addend=None, gidx=None):
        """x0 [B,C0,P] (+ x1 [B,C1,P] virtually concatenated on channels) -> [B,Cout,P].
        Dense layers run on tcgen05 (csrc/pointwise_tc.cu, fp16 hi/lo split); thin ones on the
        exact-fp32 CUDA-core kernel (csrc/pointwise.cu)."""
        norm = self.norm if self.normalization == 'batch' else None
        cin = x0.shape[1] + (0 if x1 is None else x1.shape[1])
        cout = self.conv.out_channels
        rows = x0.shape[0] * x0.shape[2]
        # tcgen05 eligibility: a dense enough contraction (cin >= 32) over at least two 128-row
        # tiles; thin OUTPUTS (the 128 -> 3 ConvToPC heads) qualify too — the layer is then a
        # streaming read of the activations, which the TMA-fed kernel does at several times the
        # rate of the register-tiled fp32 kernel
        if cin >= 32 and rows >= 256 and os.environ.get("SONET_TC", "1") != "0":
            w, shift = self._folded.get(self._conv_weight2d(), self.conv.bias, norm,
                                        transpose=False)
            key = (self._folded.version, w.data_ptr())
            if getattr(self, "_tc_key", None) != key:
                blob, inv = ops.pointwise_tc_pack(w)
                self._tc_pack = (blob.to(w.device), inv)
                self._tc_key = key
            blob, inv = self._tc_pack
            return ops.pointwise_layer_tc(x0, blob, inv, shift, cout, self.activation == 'relu',
                                          x1=x1, addend=addend, gidx=gidx)
        w, shift = self._folded.get(self._conv_weight2d(), self.conv.bias, norm, transpose=True)
        return ops.pointwise_layer(x0, w, None, shift, self.activation == 'relu', x1=x1,
                                   addend=addend, gidx=gidx)


class MyConv2d(_PointwiseConvBase):
    """Conv2d + BN2d + act (models/layers.py:169-211); 1x1 kernels run on the point-wise kernel."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 activation=None, momentum=0.1, normalization=None, bn_momentum_decay_step=None,
                 bn_momentum_decay=1):
        super().__init__()
        self.activation = activation
        self.normalization = normalization
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        if normalization == 'batch':
            self.norm = MyBatchNorm2d(out_channels, momentum=momentum, affine=True,
                                      momentum_decay_step=bn_momentum_decay_step,
                                      momentum_decay=bn_momentum_decay)
        elif normalization == 'instance':
            self.norm = nn.InstanceNorm2d(out_channels, momentum=momentum, affine=True)
        act = _make_act(activation)
        if act is not None:
            self.act = act
        self._folded = _FoldedParams()
        self.weight_init()

    def weight_init(self):
        c = self.conv
        n = c.kernel_size[0] * c.kernel_size[1] * c.in_channels
        nn.init.normal_(c.weight, 0, math.sqrt(2. / n))
        if c.bias is not None:
            nn.init.zeros_(c.bias)
        if self.normalization in ('batch', 'instance'):
            nn.init.ones_(self.norm.weight)
            nn.init.zeros_(self.norm.bias)

    def forward(self, x, epoch=None):
        if _fast_ok(self, x) and x.dim() == 4 and self._fast_eligible():
            B, C, H, W = x.shape
            y = self.forward_points(x.contiguous().view(B, C, H * W))
            return y.view(B, y.shape[1], H, W)
        if x.dim() == 4 and self._train_kernels(x):
            B, C, H, W = x.shape
            y = train_ops.conv_bn_act_train(
                x.contiguous().view(B, C, H * W), self._conv_weight2d(), self.conv.bias,
                self.norm if self.normalization == 'batch' else None, self.activation == 'relu', epoch)
            return y.view(B, y.shape[1], H, W)
        # PyTorch/cuDNN path (3x3 decoder convs, training): strict fp32 — cuDNN's default TF32
        # convolution is ~1e-3 relative, outside the 1e-4 parity bar of the reference's fp32 math
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            x = self.conv(x)
        if self.normalization == 'batch':
            x = self.norm(x, epoch)
        elif self.normalization is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.act(x)
        return x


def _pick_splits(kchunks, base_items, target=120, cap=32):
    """K split of a grouped tcgen05 launch: the smallest power of two dividing the number of
    64-channel K chunks that brings the item count to ~one wave of SMs."""
    s = 1
    while s * 2 <= cap and kchunks % (s * 2) == 0 and base_items * s < target:
        s *= 2
    return s


class UpConv(nn.Module):
    """Nearest x2 upsample + 3x3 conv + BN + act (models/layers.py:214-240), the decoder's
    building block. Eval fast path: four parity GEMMs over the LOW-resolution
    map on tcgen05 (csrc/upconv.cu, csrc/pointwise_tc.cu) — the up-sampled image is never built
    and the 3x3 taps that coincide after up-sampling are pre-summed (K = 4*Cin, not 9*Cin)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0,
                 output_padding=0, bias=True, activation=None, normalization=None):
        super().__init__()
        self.activation = activation
        self.normalization = normalization
        self.up_sample = nn.Upsample(scale_factor=2)
        self.conv = MyConv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1,
                             bias=True, activation=activation, normalization=normalization)
        self._tc_key = None
        self._tc_pack = None
        self._scratch = {}
        self.weight_init()

    def weight_init(self):
        c = self.conv.conv
        n = c.kernel_size[0] * c.kernel_size[1] * c.out_channels
        nn.init.normal_(c.weight, 0, math.sqrt(2. / n))
        if c.bias is not None:
            nn.init.constant_(c.bias, 0.001)

    # ---- eval fast path --------------------------------------------------------------------------
    def _fast(self, x):
        c = self.conv.conv
        return (_fast_ok(self, x) and x.dim() == 4 and c.kernel_size == (3, 3)
                and c.stride == (1, 1) and c.padding == (1, 1) and c.dilation == (1, 1)
                and c.groups == 1 and c.out_channels >= 64 and c.in_channels >= 16
                and self.normalization in (None, 'batch') and self.activation in (None, 'relu')
                and os.environ.get("SONET_TC", "1") != "0")

    def _packed(self):
        """BN-folded, parity-combined weights as tcgen05 blobs; re-packed when a source tensor
        changes. Wg[g=(py,px)][co][(a,c)*Cin + ci] = sum of the 3x3 taps (dy,dx) that read low-res
        neighbour (a,c) for output parity (py,px): rows {0}|{1,2} (py=0) or {0,1}|{2} (py=1)."""
        c = self.conv.conv
        norm = self.conv.norm if self.normalization == 'batch' else None
        srcs = [c.weight, c.bias]
        if norm is not None:
            srcs += [norm.weight, norm.bias, norm.running_mean, norm.running_var]
        key = tuple((t.data_ptr(), t._version, t.device) for t in srcs if t is not None)
        if key != self._tc_key:
            with torch.no_grad():
                w = c.weight.detach().float()                              # [Cout,Cin,3,3]
                shift = c.bias.detach().float() if c.bias is not None else \
                    torch.zeros(w.shape[0], device=w.device)
                if norm is not None:
                    sc = norm.weight.detach() / torch.sqrt(norm.running_var.detach() + norm.eps)
                    shift = (shift - norm.running_mean.detach()) * sc + norm.bias.detach()
                    w = w * sc[:, None, None, None]
                sets = (((0,), (1, 2)), ((0, 1), (2,)))                    # [parity][tap] -> kernel rows
                Cout, Cin = w.shape[0], w.shape[1]
                Wg = torch.zeros(4, Cout, 4, Cin, device=w.device)
                for py in range(2):
                    for px in range(2):
                        for a in range(2):
                            for cc in range(2):
                                acc = 0
                                for dy in sets[py][a]:
                                    for dx in sets[px][cc]:
                                        acc = acc + w[:, :, dy, dx]
                                Wg[py * 2 + px, :, a * 2 + cc, :] = acc
                blob4, per4, inv4 = ops.pointwise_tc_pack_groups(Wg.view(4, Cout, 4 * Cin))
                # 1x1 input: only the tap that lands on the single pixel survives -> one dense
                # layer Cin -> 4*Cout (output channel co*4 + py*2 + px)
                W1 = torch.stack([Wg[g, :, (1 - (g >> 1)) * 2 + (1 - (g & 1)), :] for g in range(4)],
                                 dim=1).reshape(1, Cout * 4, Cin)
                blob1, per1, inv1 = ops.pointwise_tc_pack_groups(W1)
                dev = w.device
                self._tc_pack = dict(blob4=blob4.to(dev), per4=per4, inv4=inv4, blob1=blob1.to(dev),
                                     per1=per1, inv1=inv1, shift=shift.contiguous(),
                                     shift1=shift.repeat_interleave(4).contiguous())
            self._tc_key = key
        return self._tc_pack

    def _scratch_for(self, n, dev):
        buf = self._scratch.get(dev)
        if buf is None or buf.numel() < n:
            buf = self._scratch[dev] = torch.empty(n, dtype=torch.float32, device=dev)
        return buf

    def forward(self, x):
        if not self._fast(x):
            return self.conv(self.up_sample(x))
        pk = self._packed()
        B, Cin, H, W = x.shape
        Cout = self.conv.conv.out_channels
        relu = self.activation == 'relu'
        x = x.contiguous()
        if H == 1 and W == 1:
            kch = (Cin + 63) // 64
            splits = _pick_splits(kch, ((B + 127) // 128) * ((4 * Cout + 255) // 256))
            scratch = self._scratch_for(splits * B * 4 * Cout, x.device) if splits > 1 else None
            y = ops.pointwise_tc_grouped(x.view(B, Cin, 1), pk["blob1"], pk["per1"], pk["inv1"],
                                         pk["shift1"], 4 * Cout, relu, groups=1, splits=splits,
                                         scratch=scratch)
            return y.view(B, Cout, 2, 2)
        P = H * W
        kch = (4 * Cin + 63) // 64
        splits = _pick_splits(kch, 4 * ((B * P + 127) // 128) * ((Cout + 255) // 256))
        scratch = self._scratch_for(4 * splits * B * Cout * P, x.device) if splits > 1 else None
        # H*W % 64 == 0 and Cin % 64 == 0: no im2col — three horizontally shifted copies, the
        # vertical shifts are TMA coordinate offsets inside the GEMM (3x the input instead of 16x)
        conv = P % 64 == 0 and Cin % 64 == 0 and os.environ.get("SONET_UPCONV_IM2COL", "0") != "1"
        xin = ops.upconv_hshift(x) if conv else ops.upconv_im2col(x)      # [3B,Cin,HW] | [4B,4Cin,HW]
        y = ops.pointwise_tc_grouped(xin, pk["blob4"], pk["per4"], pk["inv4"], pk["shift"], Cout,
                                     relu, groups=4, splits=splits, scat_w=W, scratch=scratch,
                                     conv=conv)
        return y.view(B, Cout, 2 * H, 2 * W)


class EquivariantLayer(_PointwiseConvBase):
    """Conv1d(k=1) + BN1d + act on [B,C,P] (models/layers.py:243-296)."""

    def __init__(self, num_in_channels, num_out_channels, activation='relu', normalization=None,
                 momentum=0.1, bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        self.num_in_channels = num_in_channels
        self.num_out_channels = num_out_channels
        self.activation = activation
        self.normalization = normalization
        self.conv = nn.Conv1d(num_in_channels, num_out_channels, kernel_size=1, stride=1,
                              padding=0)
        if normalization == 'batch':
            self.norm = MyBatchNorm1d(num_out_channels, momentum=momentum, affine=True,
                                      momentum_decay_step=bn_momentum_decay_step,
                                      momentum_decay=bn_momentum_decay)
        elif normalization == 'instance':
            self.norm = nn.InstanceNorm1d(num_out_channels, momentum=momentum, affine=True)
        act = _make_act(activation)
        if act is not None:
            self.act = act
        self._folded = _FoldedParams()
        self.weight_init()

    def weight_init(self):
        nn.init.normal_(self.conv.weight, 0,
                        math.sqrt(2. / (self.conv.kernel_size[0] * self.conv.in_channels)))
        if self.conv.bias is not None:
            nn.init.zeros_(self.conv.bias)
        if self.normalization in ('batch', 'instance'):
            nn.init.ones_(self.norm.weight)
            nn.init.zeros_(self.norm.bias)

    def fast(self, *tensors):
        return _fast_ok(self, *tensors) and self._fast_eligible()

    def forward(self, x, epoch=None):
        if self.fast(x) and x.dim() == 3:
            return self.forward_points(x.contiguous())
        if self._train_kernels(x) and x.dim() == 3:
            return train_ops.conv_bn_act_train(
                x, self._conv_weight2d(), self.conv.bias,
                self.norm if self.normalization == 'batch' else None, self.activation == 'relu', epoch)
        y = self.conv(x)
        if self.normalization == 'batch':
            y = self.norm(y, epoch)
        elif self.normalization is not None:
            y = self.norm(y)
        if self.activation is not None:
            y = self.act(y)
        return y


class KNNModule(nn.Module):
    """Group K neighbour nodes, decentre, 1x1-conv stack, max over K (models/layers.py:299-367)."""

    def __init__(self, in_channels, out_channels_list, activation, normalization, momentum=0.1,
                 bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        self.layers = nn.ModuleList()
        prev = in_channels
        for c_out in out_channels_list:
            self.layers.append(MyConv2d(prev, c_out, kernel_size=1, stride=1, padding=0, bias=True,
                                        activation=activation, normalization=normalization,
                                        momentum=momentum,
                                        bn_momentum_decay_step=bn_momentum_decay_step,
                                        bn_momentum_decay=bn_momentum_decay))
            prev = c_out

    def _fast(self, coordinate, x, precomputed_knn_I, center_type):
        return (_fast_ok(self, coordinate, x) and center_type in ('avg', 'center')
                and all(l._fast_eligible() for l in self.layers)
                and (precomputed_knn_I is None or precomputed_knn_I.is_cuda))

    def forward_pooled(self, coordinate, pool, precomputed_knn_I, K, center_type, epoch=None,
                       owner=None):
        """forward() for the fused-pool path: `pool` = (keys [B,C,M] i32, p0 [B,C]) as left by
        ops.pointresnet_tc_pool(finalize=False). Returns (center, feature, masked_max [B,C,M]) or
        None when this module cannot take the fused route (the caller then finalizes the pool)."""
        keys, p0 = pool
        B, C, M = keys.shape
        if not (self._fast(coordinate, p0, precomputed_knn_I, center_type)
                and M <= 256 and M * K <= 2304):
            return None
        coord = coordinate.detach().contiguous()
        if precomputed_knn_I is not None:
            if precomputed_knn_I.size()[2] < K:
                return None      # the unfused forward() raises the reference's assertion
            knn_I = precomputed_knn_I.contiguous()
        else:
            knn_I = ops.node_knn(coord, K)
        center, h, masked_max = ops.knn_assemble_pool(coord, keys, p0, knn_I, K, center_type,
                                                      owner=owner)
        for layer in self.layers:
            h = layer.forward_points(h)
        feature = ops.rowmax(h.view(B, h.shape[1], M, K))
        return center, feature, masked_max

    def forward(self, coordinate, x, precomputed_knn_I, K, center_type, epoch=None):
        """coordinate [B,3,M], x [B,C,M], precomputed_knn_I [B,M,K'] -> (center [B,3,M],
        feature [B,Cout,M])."""
        if self._fast(coordinate, x, precomputed_knn_I, center_type):
            coord = coordinate.detach().contiguous()
            if precomputed_knn_I is not None:
                assert precomputed_knn_I.size()[2] >= K
                knn_I = precomputed_knn_I.contiguous()  # kernel reads the first K columns
            else:
                knn_I = ops.node_knn(coord, K)
            B, C, M = x.shape
            center, h = ops.knn_assemble(coord, x.contiguous(), knn_I, K, center_type)
            for layer in self.layers:
                h = layer.forward_points(h)                     # [B, C', M*K]
            feature = ops.rowmax(h.view(B, h.shape[1], M, K))   # max over K
            return center, feature

        coordinate_tensor = coordinate.data
        if precomputed_knn_I is not None:
            assert precomputed_knn_I.size()[2] >= K
            knn_I = precomputed_knn_I[:, :, 0:K]
        else:
            d = torch.sum((coordinate_tensor.unsqueeze(3) - coordinate_tensor.unsqueeze(2)) ** 2,
                          dim=1)
            _, knn_I = torch.topk(d, k=K, dim=2, largest=False, sorted=True)
        neighbors = operations.knn_gather_wrapper(coordinate_tensor, knn_I)  # Bx3xMxK
        if center_type == 'avg':
            neighbors_center = torch.mean(neighbors, dim=3, keepdim=True)
        elif center_type == 'center':
            neighbors_center = coordinate_tensor.unsqueeze(3)
        neighbors_decentered = (neighbors - neighbors_center).detach()
        neighbors_center = neighbors_center.squeeze(3).detach()
        x_neighbors = operations.knn_gather_by_indexing(x, knn_I)
        x_augmented = torch.cat((neighbors_decentered, x_neighbors), dim=1)
        for layer in self.layers:
            x_augmented = layer(x_augmented, epoch)
        feature, _ = torch.max(x_augmented, dim=3, keepdim=False)
        return neighbors_center, feature


def _chain(in_channels, out_channels_list, activation, normalization, momentum,
           bn_momentum_decay_step, bn_momentum_decay, last_extra_in=0):
    layers = nn.ModuleList()
    prev = in_channels
    last = len(out_channels_list) - 1
    for i, c_out in enumerate(out_channels_list):
        if i != last:
            layers.append(EquivariantLayer(prev, c_out, activation, normalization, momentum,
                                           bn_momentum_decay_step, bn_momentum_decay))
        else:
            layers.append(EquivariantLayer(prev + last_extra_in, c_out, None, None))
        prev = c_out
    return layers


class PointNet(nn.Module):
    """Chain of EquivariantLayers, last one bare (models/layers.py:370-387)."""

    def __init__(self, in_channels, out_channels_list, activation, normalization, momentum=0.1,
                 bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        self.layers = _chain(in_channels, out_channels_list, activation, normalization, momentum,
                             bn_momentum_decay_step, bn_momentum_decay)

    def forward(self, x, epoch=None):
        for layer in self.layers:
            x = layer(x, epoch)
        return x

    def forward_pair(self, x0, x1, epoch=None):
        """forward(torch.cat((x0, x1), dim=1)) without materialising the concat in eval mode."""
        first = self.layers[0]
        if first.fast(x0, x1):
            x = first.forward_points(x0.contiguous(), x1.contiguous())
            for layer in self.layers[1:]:
                x = layer(x, epoch)
            return x
        return self.forward(torch.cat((x0, x1), dim=1), epoch)


class PointResNet(nn.Module):
    """in -> out[0] -> ... -> out[k-2]; cat(out[0], out[k-2]) -> out[k-1] (models/layers.py:390-432)."""

    def __init__(self, in_channels, out_channels_list, activation, normalization, momentum=0.1,
                 bn_momentum_decay_step=None, bn_momentum_decay=1):
        super().__init__()
        self.out_channels_list = out_channels_list
        self.layers = _chain(in_channels, out_channels_list, activation, normalization, momentum,
                             bn_momentum_decay_step, bn_momentum_decay,
                             last_extra_in=out_channels_list[0])

    # ---- fused tcgen05 path (csrc/pointmlp_tc.cu): the encoder's first PointResNet -------------
    def _tc_eligible(self, cin, x1):
        if x1 is not None or os.environ.get("SONET_TC", "1") == "0":
            return False
        if list(self.out_channels_list) != [64, 128, 256, 384] or cin > 6:
            return False
        ls = self.layers
        return (all(l.normalization == 'batch' and l.activation == 'relu' for l in ls[:3])
                and ls[3].normalization is None and ls[3].activation is None)

    def _tc_params(self):
        """(blob, fparams) on the weights' device, re-packed only when a parameter changes."""
        folded = [l._folded.get(l._conv_weight2d(), l.conv.bias,
                                l.norm if l.normalization == 'batch' else None, transpose=False)
                  for l in self.layers]
        key = tuple(l._folded.version for l in self.layers) + \
            tuple(w.data_ptr() for w, _ in folded)
        if getattr(self, "_tc_key", None) != key:
            dev = folded[0][0].device
            blob, fpar = ops.pointresnet_tc_pack([w for w, _ in folded], [s for _, s in folded],
                                                 folded[0][0].shape[1])
            self._tc_pack = (blob.to(dev), fpar.to(dev))
            self._tc_key = key
        return self._tc_pack

    def forward(self, x, epoch=None, x1=None):
        """x [B,C,P]; x1: optional second tensor virtually concatenated to x (eval fast path)."""
        n = len(self.out_channels_list)
        first, last = self.layers[0], self.layers[n - 1]
        if first.fast(x) and x.dim() == 3 and all(l.fast(x) for l in self.layers) \
                and self._tc_eligible(x.shape[1], x1):
            blob, fpar = self._tc_params()
            return ops.pointresnet_tc(x.contiguous(), blob, fpar)
        if first.fast(x) and x.dim() == 3 and all(l.fast(x) for l in self.layers):
            layer0_out = first.forward_points(x.contiguous(),
                                              None if x1 is None else x1.contiguous())
            x_tmp = layer0_out
            for l in range(1, n - 1):
                x_tmp = self.layers[l].forward_points(x_tmp)
            return last.forward_points(layer0_out, x_tmp)   # the skip-concat, never materialised
        if x1 is not None:
            x = torch.cat((x, x1), dim=1)
        layer0_out = first(x, epoch)
        x_tmp = layer0_out
        for l in range(1, n - 1):
            x_tmp = self.layers[l](x_tmp, epoch)
        return last(torch.cat((layer0_out, x_tmp), dim=1), epoch)
