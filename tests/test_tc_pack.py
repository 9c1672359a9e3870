"""Host-side packing of the tcgen05 point-MLP weights (no GPU needed): the bf16 hi/lo K-major
core-matrix images must reproduce the fp32 weights to ~2^-16 and sit at the documented offsets."""
import numpy as np
import torch


def _bf16(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def test_pack_images_round_trip(lib_built):
    from sonet_b200 import ops
    rs = np.random.RandomState(0)
    W = [torch.from_numpy(rs.normal(size=s).astype(np.float32))
         for s in ((64, 6), (128, 64), (256, 128), (384, 320))]
    sh = [torch.from_numpy(rs.normal(size=s).astype(np.float32)) for s in (64, 128, 256, 384)]
    blob, fpar = ops.pointresnet_tc_pack(W, sh, 6)
    assert blob.numel() == 32768 + 4 * 32768 + 6 * (32768 * 2 + 16384)
    b = blob.numpy().view(np.uint16)

    def unpack(base, rows, kt, hi_bytes):
        out = np.zeros((rows, kt), np.float32)
        for r in range(rows):
            for k in range(kt):
                off = (r >> 3) * kt * 16 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2
                hi = b[(base + off) // 2]
                lo = b[(base + hi_bytes + off) // 2]
                out[r, k] = _bf16(np.array([hi]))[0] + _bf16(np.array([lo]))[0]
        return out

    w1 = unpack(0, 128, 64, 16384)
    assert np.abs(w1 - W[1].numpy()).max() <= 2.0 ** -15 * np.abs(W[1].numpy()).max()
    # layer-2 chunk 3 (rows 192..255) is the 4th streamed stage
    w2c = unpack(32768 + 3 * 32768, 64, 128, 16384)
    assert np.abs(w2c - W[2].numpy()[192:256]).max() <= 2.0 ** -15 * np.abs(W[2].numpy()).max()
    # layer-3 chunk 1, K slab 2 (rows 64..127, k 256..319): stage 4 + 3 + 2
    base = 32768 + 4 * 32768 + (32768 * 2 + 16384) + 2 * 32768
    w3c = unpack(base, 64, 64, 8192)
    assert np.abs(w3c - W[3].numpy()[64:128, 256:320]).max() <= 2.0 ** -15 * np.abs(W[3].numpy()).max()
    f = fpar.numpy()
    assert np.array_equal(f[:384].reshape(64, 6), W[0].numpy())
    assert np.array_equal(f[384:448], sh[0].numpy()) and np.array_equal(f[448 + 128 + 256:], sh[3].numpy())
