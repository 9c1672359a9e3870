"""-m gpu: model-level parity of the CUDA path against (a) the golden vectors produced by the
reference itself and (b) the oracle on fresh seeded inputs; plus full-size properties."""
import numpy as np
import pytest
import torch

from helpers import (assert_close, assert_golden, build_states, golden, golden_case,
                     to_ref_slot_order, to_slot_order)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gpu_opt(opt):
    opt.device = torch.device(DEV)
    opt.gpu_id = 0
    return opt


def _classifier(opt, st):
    from sonet_b200 import classifier
    m = classifier.Model(_gpu_opt(opt))
    m.encoder.load_state_dict(st["encoder"])
    m.classifier.load_state_dict(st["head"])
    return m


def _check_encoder_against_golden(g, enc, opt):
    from oracle import oracle
    assert np.array_equal(oracle.canon_sets(enc.min_idx.cpu(), opt.k).numpy(), g["knn_sets"])
    mask = enc.mask
    assert mask.dtype == torch.int32
    assert np.array_equal(torch.max(mask, dim=1)[0].cpu().numpy(), g["mask_row_max"])
    assert np.array_equal(torch.sum(mask, dim=1).cpu().numpy(), g["mask_row_sum"])
    for n in ("som_node", "first_pn_out_masked_max", "final_pn_out", "feature"):
        assert_golden(g, n, getattr(enc, n))
    assert_golden(g, "first_pn_out", to_ref_slot_order(enc.first_pn_out, enc.min_idx, g, opt.k))
    if opt.som_k >= 2:
        assert_golden(g, "knn_center_1", enc.knn_center_1)
        assert_golden(g, "knn_feature_1", enc.knn_feature_1)


@pytest.mark.parametrize("name", ["classifier_b2_n256", "classifier_b2_n200_emptynodes",
                                  "classifier_b2_n256_somk0"])
def test_classifier_vs_reference_golden(name):
    g = golden(name)
    opt, inp, seed = golden_case(g, "classifier")
    m = _classifier(opt, build_states("classifier", opt, seed))
    m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
    m.test_model()
    _check_encoder_against_golden(g, m.encoder, opt)
    assert_golden(g, "score", m.score)


def test_classifier_vs_oracle_cfg1_shape(oracle_mod):
    """BASELINE.json configs[0]: B=8, N=1024, 8x8 SOM."""
    from sonet_b200 import synth
    opt = synth.make_opt("classifier", batch_size=8, input_pc_num=1024)
    st = build_states("classifier", opt, seed=11)
    inp = synth.synth_inputs(8, 1024, seed=11)
    m = _classifier(opt, st)
    m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
    m.test_model()
    o = oracle_mod.encoder_forward(st["encoder"], opt, inp["pc"], inp["sn"], inp["node"],
                                   inp["node_knn_I"])
    enc = m.encoder
    assert torch.equal(oracle_mod.canon_sets(enc.min_idx.cpu(), 3),
                       oracle_mod.canon_sets(o["min_idx"], 3))
    # bit-exact arg-max indices on identical inputs: pool the ORACLE's activations on the GPU
    from sonet_b200 import ops
    gi = ops.index_max(o["first_pn_out"].to(DEV), enc.min_idx, 64)
    ref_gi = oracle_mod.index_max(o["first_pn_out"], enc.min_idx.cpu(), 64)
    assert torch.equal(gi.cpu(), ref_gi)
    for n in ("som_node", "first_pn_out_masked_max", "knn_center_1", "knn_feature_1",
              "final_pn_out", "feature"):
        assert_close(getattr(enc, n), o[n], n)
    assert_close(m.score, oracle_mod.classifier_forward(st["head"], o["feature"]), "score")
    # per-copy tensors: in the oracle's slot order (torch.topk(sorted=False) order is unspecified)
    for n in ("first_pn_out", "centers", "x_decentered"):
        assert_close(to_slot_order(getattr(enc, n), enc.min_idx, o["min_idx"], 3), o[n], n)


def test_segmenter_vs_reference_golden_and_dropin_signature(oracle_mod):
    from sonet_b200 import segmenter
    g = golden("segmenter_b2_n128")
    opt, inp, seed = golden_case(g, "segmenter")
    st = build_states("segmenter", opt, seed)
    m = segmenter.Model(_gpu_opt(opt))
    m.encoder.load_state_dict(st["encoder"])
    m.segmenter.load_state_dict(st["head"])
    seg = torch.zeros(int(g["B"]), int(g["N"]), dtype=torch.int64)
    m.set_input(inp["pc"], inp["sn"], inp["label"], seg, inp["node"], inp["node_knn_I"])
    m.test_model()
    _check_encoder_against_golden(g, m.encoder, opt)
    assert_golden(g, "centers", to_ref_slot_order(m.encoder.centers, m.encoder.min_idx, g, opt.k))
    assert_golden(g, "x_decentered",
                  to_ref_slot_order(m.encoder.x_decentered, m.encoder.min_idx, g, opt.k))
    assert_golden(g, "score_segmenter", m.score_segmenter)
    # the reference call signature (per-point tensors gathered by the caller,
    # models/segmenter.py:90-109) gives the same scores as the node-level fast entry
    enc = m.encoder
    with torch.no_grad():
        B, kN = enc.min_idx.shape
        idx = torch.max(enc.mask, dim=2)[1].unsqueeze(1)
        gat = lambda t: torch.gather(t, 2, idx.expand(B, t.shape[1], kN))  # noqa: E731
        s2 = m.segmenter(enc.x_decentered, m.pc, enc.centers, m.sn, m.input_label,
                         enc.first_pn_out, gat(enc.first_pn_out_masked_max),
                         gat(enc.knn_feature_1), gat(enc.final_pn_out), m.feature)
    assert_golden(g, "score_segmenter", s2)
    # two numeric paths (dense fp32 concat GEMM vs split tcgen05 GEMMs): both inside the 1e-4 bar
    assert_close(s2, m.score_segmenter, "forward vs forward_nodes")


def test_autoencoder_vs_reference_golden():
    from sonet_b200 import autoencoder
    g = golden("autoencoder_b2_n256")
    opt, inp, seed = golden_case(g, "autoencoder")
    st = build_states("autoencoder", opt, seed)
    m = autoencoder.Model(_gpu_opt(opt))
    m.encoder.load_state_dict(st["encoder"])
    m.decoder.load_state_dict(st["head"])
    m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
    m.test_model()
    _check_encoder_against_golden(g, m.encoder, opt)
    # the up-convolution decoder runs on the tcgen05 kernels (csrc/upconv.cu): 1e-4 like
    # everything else (round 1 ran it through cuDNN and needed 5e-4)
    assert_golden(g, "predicted_pc", m.predicted_pc)
    assert_golden(g, "conv_pc4", m.decoder.conv_pc4)
    assert_close(m.loss_chamfer, g["loss_chamfer"], "loss_chamfer")
    assert_close(m.loss_chamfer_conv4, g["loss_chamfer_conv4"], "loss_chamfer_conv4")
    assert_close(m.loss, g["loss"], "loss")
    assert_close(m.chamfer_criteria.loss_array, g["loss_array"], "loss_array")


def test_full_size_shard_invariance_and_determinism():
    """cfg-2 (B=64, N=5000): the forward is per-cloud, so running the two halves of the batch
    separately must reproduce the full-batch logits BIT-EXACTLY (the multi-GPU parity definition,
    SURVEY.md §8e), and repeated runs are bit-identical (fixed-order reductions)."""
    from sonet_b200 import synth
    B, N = 64, 5000
    opt = synth.make_opt("classifier", batch_size=B, input_pc_num=N)
    st = build_states("classifier", opt, seed=21)
    inp = synth.synth_inputs(B, N, seed=21)
    m = _classifier(opt, st)
    keys = ("pc", "sn", "label", "node", "node_knn_I")
    m.set_input(*[inp[k] for k in keys])
    m.test_model()
    full = m.score.clone()
    m.test_model()
    assert torch.equal(full, m.score)
    halves = []
    for lo in (0, 32):
        m.set_input(*[inp[k][lo:lo + 32] for k in keys])
        m.test_model()
        halves.append(m.score.clone())
    assert torch.equal(full, torch.cat(halves))
    assert torch.isfinite(full).all()


def _assert_encoder_slice_vs_oracle(oracle_mod, enc, o, k, n_clouds):
    """Every node-level tensor of the first n_clouds clouds against the oracle (1e-4), k-sets exact."""
    assert torch.equal(oracle_mod.canon_sets(enc.min_idx[:n_clouds].cpu(), k),
                       oracle_mod.canon_sets(o["min_idx"], k))
    for n in ("som_node", "first_pn_out_masked_max", "knn_center_1", "knn_feature_1",
              "final_pn_out", "feature"):
        assert_close(getattr(enc, n)[:n_clouds], o[n], n)


def test_cfg2_benchmarked_pipeline_vs_oracle(oracle_mod):
    """The EXACT pipeline bench.py times — classifier.Model with enable_cuda_graph(True) at
    BASELINE.json configs[1] (B=64, N=5000: som_group + fused tcgen05 PointResNet/pool at
    kN=15000 + graph replay) — against the oracle on a 3-cloud slice: k-sets exact, som_node,
    first_pn_out_masked_max, knn_feature_1, final_pn_out, feature and score within 1e-4
    (|a-b| <= 1e-4 * max(|b|,1)). Checked on the capture run AND on a replay with new inputs."""
    from sonet_b200 import synth
    B, N, S = 64, 5000, 3
    opt = synth.make_opt("classifier", batch_size=B, input_pc_num=N)
    st = build_states("classifier", opt, seed=61)
    m = _classifier(opt, st)
    m.enable_cuda_graph(True)
    keys = ("pc", "sn", "label", "node", "node_knn_I")
    cpu_opt = synth.make_opt("classifier", batch_size=S, input_pc_num=N)
    # seeds 61/62 -> buffer set 1, then 0 (capture runs); 63/64 -> replays of both graphs.
    # 'uniform' nodes on the last one: empty nodes through the fused pool at full size
    for i, (seed, mode) in enumerate(((61, "sampled"), (62, "sampled"), (63, "sampled"),
                                      (64, "uniform"))):
        inp = synth.synth_inputs(B, N, seed=seed, node_mode=mode)
        if mode == "uniform":
            inp["node"] = inp["node"] * 2.5     # nodes far outside the cloud: many stay empty
        m.set_input(*[inp[k] for k in keys])
        m.test_model()
        if i in (1, 2):
            continue                     # oracle time: check the first capture and both replays' ends
        # sorted_slots: with empty nodes the reference gathers "the feature of stacked copy 0",
        # whose node depends on topk(sorted=False)'s implementation-defined slot order; the
        # oracle instance with ascending slots is the one the CUDA path reproduces
        o = oracle_mod.encoder_forward(st["encoder"], cpu_opt, inp["pc"][:S], inp["sn"][:S],
                                       inp["node"][:S], inp["node_knn_I"][:S],
                                       sorted_slots=(mode == "uniform"))
        _assert_encoder_slice_vs_oracle(oracle_mod, m.encoder, o, 3, S)
        assert_close(m.score[:S], oracle_mod.classifier_forward(st["head"], o["feature"]), "score")
        if mode == "uniform":
            assert int((o["mask_row_max"] == 0).sum()) > 0, "case must contain empty nodes"


def test_training_step_runs_on_gpu_and_changes_weights():
    """train() mode composes differentiable PyTorch ops around the kernels' indices."""
    from sonet_b200 import synth
    opt = synth.make_opt("classifier", batch_size=4, input_pc_num=256)
    st = build_states("classifier", opt, seed=31)
    inp = synth.synth_inputs(4, 256, seed=31)
    m = _classifier(opt, st)
    m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
    w0 = m.encoder.first_pointnet.layers[0].conv.weight.detach().clone()
    m.optimize()
    assert torch.isfinite(m.loss)
    assert not torch.equal(w0, m.encoder.first_pointnet.layers[0].conv.weight.detach())
    m.test_model()     # folded weights are re-packed after the in-place optimizer update
    assert torch.isfinite(m.score).all()


def test_segmenter_cfg3_full_size(oracle_mod):
    """BASELINE.json configs[2]: segmenter forward, B=32, N=1024. Full batch: shard invariance
    (bit-exact halves) + finiteness; a 2-cloud slice against the oracle (1e-4)."""
    from sonet_b200 import segmenter, synth
    B, N = 32, 1024
    opt = synth.make_opt("segmenter", batch_size=B, input_pc_num=N)
    st = build_states("segmenter", opt, seed=41)
    inp = synth.synth_inputs(B, N, seed=41)
    m = segmenter.Model(_gpu_opt(opt))
    m.encoder.load_state_dict(st["encoder"])
    m.segmenter.load_state_dict(st["head"])
    seg = torch.zeros(B, N, dtype=torch.int64)
    keys = ("pc", "sn", "label")
    m.set_input(inp["pc"], inp["sn"], inp["label"], seg, inp["node"], inp["node_knn_I"])
    m.test_model()
    full = m.score_segmenter.clone()
    assert full.shape == (B, 50, N) and torch.isfinite(full).all()
    halves = []
    for lo in (0, 16):
        sl = slice(lo, lo + 16)
        m.set_input(inp["pc"][sl], inp["sn"][sl], inp["label"][sl], seg[sl], inp["node"][sl],
                    inp["node_knn_I"][sl])
        m.test_model()
        halves.append(m.score_segmenter.clone())
    assert torch.equal(full, torch.cat(halves))
    cpu_opt = synth.make_opt("segmenter", batch_size=2, input_pc_num=N)
    o = oracle_mod.encoder_forward(st["encoder"], cpu_opt, inp["pc"][:2], inp["sn"][:2],
                                   inp["node"][:2], inp["node_knn_I"][:2])
    want = oracle_mod.segmenter_forward(st["head"], cpu_opt, o, inp["pc"][:2], inp["sn"][:2],
                                        inp["label"][:2])
    assert_close(full[:2], want, "segmenter cfg-3 slice vs oracle")


def test_autoencoder_cfg4_full_size(oracle_mod):
    """BASELINE.json configs[3]: AE forward + Chamfer, B=32, N=5000: finite, deterministic,
    shard-consistent per-cloud losses; Chamfer of a 2-cloud slice against the oracle."""
    from sonet_b200 import autoencoder, synth
    B, N = 32, 5000
    opt = synth.make_opt("autoencoder", batch_size=B, input_pc_num=N)
    st = build_states("autoencoder", opt, seed=43)
    inp = synth.synth_inputs(B, N, seed=43)
    m = autoencoder.Model(_gpu_opt(opt))
    m.encoder.load_state_dict(st["encoder"])
    m.decoder.load_state_dict(st["head"])
    keys = ("pc", "sn", "label", "node", "node_knn_I")
    m.set_input(*[inp[k] for k in keys])
    m.test_model()
    loss, arr = float(m.loss), m.chamfer_criteria.loss_array.clone()
    pred = m.predicted_pc.clone()
    assert pred.shape == (B, 3, 1280) and np.isfinite(loss) and arr.shape == (B,)
    m.test_model()
    assert float(m.loss) == loss and torch.equal(arr, m.chamfer_criteria.loss_array)
    m.set_input(*[inp[k][:16] for k in keys])
    m.test_model()
    # per-cloud losses are shard-consistent (the decoder's K split depends on the batch size, so
    # the partial-sum grouping — not the order — differs: ~1e-7, not bit-exact)
    assert_close(m.chamfer_criteria.loss_array, arr[:16], "cfg-4 shard consistency", 1e-5)
    o = oracle_mod.chamfer(pred[:2].cpu(), inp["pc"][:2])
    assert_close(arr[:2], o["loss_array"], "cfg-4 chamfer loss_array slice vs oracle")
    # the N=5000 encoder of this config against the oracle (2-cloud slice, 1e-4)
    cpu_opt = synth.make_opt("autoencoder", batch_size=2, input_pc_num=N)
    m.set_input(*[inp[k] for k in keys])
    m.test_model()
    oe = oracle_mod.encoder_forward(st["encoder"], cpu_opt, inp["pc"][:2], inp["sn"][:2],
                                    inp["node"][:2], inp["node_knn_I"][:2])
    _assert_encoder_slice_vs_oracle(oracle_mod, m.encoder, oe, 3, 2)


def test_cuda_graph_replay_matches_eager_and_tracks_changes():
    """classifier.Model.enable_cuda_graph: bit-identical scores, per-buffer-set graphs, re-capture
    after a weight update, launch accounting."""
    from sonet_b200 import ops, synth
    opt = synth.make_opt("classifier", batch_size=4, input_pc_num=1024)
    st = build_states("classifier", opt, seed=51)
    m = _classifier(opt, st)
    a, b = synth.synth_inputs(4, 1024, seed=51), synth.synth_inputs(4, 1024, seed=52)
    keys = ("pc", "sn", "label", "node", "node_knn_I")
    want = []
    for inp in (a, b):
        m.set_input(*[inp[k] for k in keys])
        m.test_model()
        want.append(m.score.clone())
    m.enable_cuda_graph(True)
    for rep in range(3):                       # alternates the two input sets / graphs
        for inp, w in zip((a, b), want):
            m.set_input(*[inp[k] for k in keys])
            k0 = ops.KERNEL_LAUNCHES
            m.test_model()
            assert ops.KERNEL_LAUNCHES - k0 >= 12
            assert torch.equal(m.score, w), rep
    with torch.no_grad():
        m.classifier.fc3.linear.bias.add_(1.0)  # parameter version changes -> re-capture
    m.set_input(*[a[k] for k in keys])
    m.test_model()
    assert torch.allclose(m.score, want[0] + 1.0, atol=1e-5)


def test_segmenter_training_gradients_reach_the_encoder():
    """in train()/grad mode Segmenter.forward_nodes must gather the node-level feature
    maps with a differentiable op (models/segmenter.py:96-98 uses torch.gather), so that the
    encoder is trained through first_pn_out_masked_max, knn_feature_1 and final_pn_out — compared
    against the reference call signature (per-point torch.gather by the caller)."""
    from sonet_b200 import segmenter, synth
    B, N = 4, 256
    opt = synth.make_opt("segmenter", batch_size=B, input_pc_num=N)
    st = build_states("segmenter", opt, seed=91)
    inp = synth.synth_inputs(B, N, seed=91)
    seg = (torch.arange(B * N).view(B, N) * 7) % 50
    grads = {}
    for mode in ("nodes", "reference_signature"):
        m = segmenter.Model(_gpu_opt(opt))
        m.encoder.load_state_dict(st["encoder"])
        m.segmenter.load_state_dict(st["head"])
        m.set_input(inp["pc"], inp["sn"], inp["label"], seg, inp["node"], inp["node_knn_I"])
        m.encoder.train()
        m.segmenter.train()
        torch.manual_seed(0)
        if mode == "nodes":
            m.forward(is_train=True)
        else:
            enc = m.encoder
            m.feature = enc(m.pc, m.sn, m.input_node, m.input_node_knn_I, True, None)
            kN = enc.min_idx.shape[1]
            idx = enc.min_idx.long().unsqueeze(1)
            g = lambda t: torch.gather(t, 2, idx.expand(B, t.shape[1], kN))   # noqa: E731
            m.score_segmenter = m.segmenter(enc.x_decentered, m.pc, enc.centers, m.sn, m.input_label,
                                            enc.first_pn_out, g(enc.first_pn_out_masked_max),
                                            g(enc.knn_feature_1), g(enc.final_pn_out), m.feature)
        loss = m.softmax_segmenter(m.score_segmenter, m.seg)
        m.encoder.zero_grad()
        loss.backward()
        grads[mode] = {n: p.grad.clone() for n, p in m.encoder.named_parameters() if p.grad is not None}
    for name in ("final_pointnet.layers.1.conv.weight", "knnlayer.layers.1.conv.weight",
                 "first_pointnet.layers.3.conv.weight"):
        a, b = grads["nodes"][name], grads["reference_signature"][name]
        assert float(b.abs().max()) > 0, name
        err = float((a - b).abs().max()) / float(b.abs().max())
        assert err <= 2e-3, "%s: %.3e" % (name, err)


def test_pool_keys_survive_an_aborted_forward():
    """a forward that dies between the fused pool launch and its finalisation must not
    poison the next one (the keys are per-Encoder and re-initialised when left dirty)."""
    from sonet_b200 import synth
    opt = synth.make_opt("classifier", batch_size=2, input_pc_num=512)
    st = build_states("classifier", opt, seed=93)
    m = _classifier(opt, st)
    keys = ("pc", "sn", "label", "node", "node_knn_I")
    a, b = synth.synth_inputs(2, 512, seed=93), synth.synth_inputs(2, 512, seed=94)
    m.set_input(*[b[k] for k in keys])
    m.test_model()
    want = m.score.clone()
    m.set_input(*[a[k] for k in keys])
    bad_knn = a["node_knn_I"][:, :, :4].contiguous()        # fewer than som_k columns
    with pytest.raises(AssertionError):                      # the reference's own assertion (layers.py:330)
        with torch.no_grad():
            m.encoder(m.pc, m.sn, m.input_node, bad_knn.to(DEV))
    m.set_input(*[b[k] for k in keys])
    m.test_model()
    assert torch.equal(m.score, want)
