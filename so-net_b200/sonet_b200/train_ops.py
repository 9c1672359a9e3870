"""Train-mode building blocks on the sm_100a kernels (csrc/train.cu): autograd
Functions that the layer classes use in train() mode instead of the PyTorch composition
conv1d -> batch_norm(training=True) -> relu and gather (models/layers.py:22-70, 282-296;
models/networks.py:185).

    ConvTC        y = W x + b, dx = W^T dy AND dW = dy x^T on the generic tcgen05 layer kernel
                  (weights / activations packed on the device each step; gradients pre-scaled by a
                  power of two before the fp16 hi/lo split)
    BNActTrain    batch-statistics BatchNorm + ReLU, forward and backward, fused elementwise passes
                  and two-stage deterministic reductions
    IndexMaxGather  first_pn_out_masked_max = first_pn_out.gather(2, idx * mask_row_max) with the
                  arg-max kernel in the forward and a deterministic scatter in the backward

ENABLED = False switches every layer back to the PyTorch composition (the tests compare the two).
"""
import torch

from . import _C, ops

ENABLED = True


def _scratch(dev, n, dtype, cache={}):
    key = (dev, dtype)
    buf = cache.get(key)
    if buf is None or buf.numel() < n:
        buf = cache[key] = torch.empty(max(n, 1), dtype=dtype, device=dev)
    return buf


class BNActTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, relu):
        x = x.contiguous()
        B, C, P = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            y = torch.empty_like(x)
            mean = torch.empty(C, dtype=torch.float32, device=dev)
            var = torch.empty(C, dtype=torch.float32, device=dev)
            invstd = torch.empty(C, dtype=torch.float32, device=dev)
            part = _scratch(dev, _C.lib().sonet_bn_partial_slots(B, C), torch.float64)
            ops._call("sonet_bn_train_forward_f32", _C.ptr(x), _C.ptr(gamma), _C.ptr(beta), B, C, P,
                      float(eps), int(bool(relu)), _C.ptr(part), _C.ptr(y), _C.ptr(mean), _C.ptr(var),
                      _C.ptr(invstd), ops._stream(x), kernels=3)
        ctx.save_for_backward(x, mean, invstd, gamma, beta)
        ctx.relu = bool(relu)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        x, mean, invstd, gamma, beta = ctx.saved_tensors
        dy = dy.contiguous()
        B, C, P = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            dx = torch.empty_like(x)
            dgamma = torch.empty(C, dtype=torch.float32, device=dev)
            dbeta = torch.empty(C, dtype=torch.float32, device=dev)
            part = _scratch(dev, _C.lib().sonet_bn_partial_slots(B, C), torch.float64)
            ops._call("sonet_bn_train_backward_f32", _C.ptr(dy), _C.ptr(x), _C.ptr(mean),
                      _C.ptr(invstd), _C.ptr(gamma), _C.ptr(beta), B, C, P, int(ctx.relu),
                      _C.ptr(part), _C.ptr(dx), _C.ptr(dgamma), _C.ptr(dbeta), ops._stream(x),
                      kernels=3)
        return dx, dgamma, dbeta, None, None


def _pack_device(W, transpose):
    """W [Cout,Cin] (device) -> (blob, scale2) for the tcgen05 layer; W^T when transpose."""
    Cout, Cin = W.shape
    n, k = (Cin, Cout) if transpose else (Cout, Cin)
    dev = W.device
    with torch.cuda.device(dev):
        blob = torch.empty(int(_C.lib().sonet_pointwise_tc_blob_bytes(n, k)), dtype=torch.uint8,
                           device=dev)
        scale2 = torch.empty(2, dtype=torch.float32, device=dev)
        bits = torch.empty(1, dtype=torch.int32, device=dev)
        ops._call("sonet_pointwise_tc_pack_device", _C.ptr(W), Cout, Cin, int(transpose),
                  _C.ptr(blob), _C.ptr(scale2), _C.ptr(bits), ops._stream(W), kernels=3)
    return blob, scale2


def _absmax_scale(t):
    """(scale, 1/scale) device tensor: the power of two that puts max|t| into [256,512) — gradients
    are far below fp16's normal range, so they are pre-scaled before the fp16 hi/lo split."""
    dev = t.device
    with torch.cuda.device(dev):
        scale2 = torch.empty(2, dtype=torch.float32, device=dev)
        bits = torch.empty(1, dtype=torch.int32, device=dev)
        ops._call("sonet_absmax_scale_f32", _C.ptr(t), t.numel(), _C.ptr(scale2), _C.ptr(bits),
                  ops._stream(t), kernels=2)
    return scale2


def _tc_dev(x, blob, inv, shift, cout, act_scale=None):
    """Generic tcgen05 layer with a device-packed blob: inv (1-element device tensor) = the total
    inverse pre-scale, act_scale (1-element) = activation pre-scale or None."""
    B, C, P = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((B, cout, P), dtype=torch.float32, device=x.device)
        ops._call("sonet_pointwise_tc_forward_dev", _C.ptr(x), C, B, P, _C.ptr(blob), _C.ptr(inv),
                  _C.ptr(act_scale), _C.ptr(shift), int(cout), 0, 1, _C.ptr(out), None,
                  ops._stream(x))
    return out


def tc_eligible(cin, cout, rows):
    return cin >= 32 and cout >= 32 and rows >= 256


def wgrad_tc(dy, x):
    """dW [Cout,Cin] = sum_{b,p} dy[b,:,p] x[b,:,p]^T on tcgen05 (csrc/train.cu: sonet_wgrad_tc_f32)."""
    B, Cout, P = dy.shape
    Cin = x.shape[1]
    lib = _C.lib()
    base = ((Cout + 127) // 128) * (((Cin + 63) // 64 * 64 + 255) // 256)
    splits = max(1, min(64, 296 // base, (B * P + 63) // 64))
    Kpad = int(lib.sonet_wgrad_kpad(B, P, splits))
    dev = dy.device
    with torch.cuda.device(dev):
        dyT = _scratch(dev, Kpad * Cout, torch.float32, _WGRAD_CACHE.setdefault("dyT", {}))
        blob = _scratch(dev, int(lib.sonet_wgrad_blob_bytes(Cin, Kpad)), torch.uint8,
                        _WGRAD_CACHE.setdefault("blob", {}))
        part = _scratch(dev, splits * Cin * Cout, torch.float32, _WGRAD_CACHE.setdefault("part", {}))
        small = torch.empty(8, dtype=torch.float32, device=dev)
        dWT = torch.empty((Cin, Cout), dtype=torch.float32, device=dev)
        ops._call("sonet_wgrad_tc_f32", _C.ptr(dy), _C.ptr(x), B, Cout, Cin, P, splits, _C.ptr(dyT),
                  _C.ptr(blob), _C.ptr(part), _C.ptr(small), _C.ptr(dWT), ops._stream(dy), kernels=9)
    return dWT.t()


_WGRAD_CACHE = {}


class ConvTC(torch.autograd.Function):
    """x [B,Cin,P], W [Cout,Cin], b [Cout] or None -> W x + b. Forward, dgrad and wgrad all run on
    the generic tcgen05 layer kernel (fp16 hi/lo split, fp32 accumulation)."""

    @staticmethod
    def forward(ctx, x, W, b):
        x = x.contiguous()
        Wc = W.contiguous()
        blob, scale2 = _pack_device(Wc, False)
        y = _tc_dev(x, blob, scale2[1:], None if b is None else b.contiguous(), Wc.shape[0])
        ctx.save_for_backward(x, Wc)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            blob, scale2 = _pack_device(W, True)                     # dgrad: dx = W^T dy
            s_dy = _absmax_scale(dy)
            inv = scale2[1:] * s_dy[1:]
            dx = _tc_dev(dy, blob, inv, None, W.shape[1], act_scale=s_dy[:1])
        if ctx.needs_input_grad[1]:
            if W.shape[0] % 64 == 0 and W.shape[1] >= 16 and WGRAD_TC:
                dW = wgrad_tc(dy, x)
            else:                                                    # thin layers: cuBLAS fp32
                prev = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                try:
                    dW = torch.einsum("bop,bip->oi", dy, x)
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = prev
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 2))
        return dx, dW, db


WGRAD_TC = True


class IndexMaxGather(torch.autograd.Function):
    """data [B,C,N], index [B,N] int32, K -> data.gather(2, idx * mask_row_max) [B,C,K]
    (models/networks.py:181-185), differentiable w.r.t. data."""

    @staticmethod
    def forward(ctx, data, index, K):
        data = data.contiguous()
        idx, val = ops.index_max(data.detach(), index, K, with_values=True)
        ctx.save_for_backward(idx)
        ctx.N = data.shape[2]
        return val

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        B, C, K = g.shape
        with torch.cuda.device(g.device):
            gd = torch.empty((B, C, ctx.N), dtype=torch.float32, device=g.device)
            ops._call("sonet_index_max_backward_f32", _C.ptr(g), _C.ptr(idx), B, C, ctx.N, K,
                      _C.ptr(gd), ops._stream(g))
        return gd, None, None


def conv_bn_act_train(x3, weight2d, bias, norm, relu, epoch=None):
    """Train-mode EquivariantLayer / MyConv2d(1x1) body on [B,Cin,P]: conv (tcgen05 when dense
    enough) -> batch-stat BN (+ running-stat update with the reference's momentum schedule,
    models/layers.py:57-65) -> ReLU."""
    B, Cin, P = x3.shape
    Cout = weight2d.shape[0]
    if tc_eligible(Cin, Cout, B * P):
        y0 = ConvTC.apply(x3, weight2d, bias)
    else:
        y0 = torch.nn.functional.conv1d(x3, weight2d.unsqueeze(2), bias)
    if norm is None:
        return torch.relu(y0) if relu else y0
    step = norm.momentum_decay_step
    if epoch is not None and epoch >= 1 and step is not None and step > 0:
        norm.momentum = max(norm.momentum_original * (norm.momentum_decay ** (epoch // step)), 0.01)
    y, mean, var = BNActTrain.apply(y0, norm.weight, norm.bias, norm.eps, relu)
    with torch.no_grad():
        n = B * P
        m = norm.momentum
        norm.running_mean.mul_(1 - m).add_(mean, alpha=m)
        norm.running_var.mul_(1 - m).add_(var * (n / max(n - 1, 1)), alpha=m)
        if norm.num_batches_tracked is not None:
            norm.num_batches_tracked += 1
    return y
