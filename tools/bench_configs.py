"""Throughput of the other BASELINE.json configs (not bench lines, context for DESIGN.md):
cfg-1 classifier B=8 N=1024, cfg-3 segmenter B=32 N=1024, cfg-4 autoencoder+Chamfer B=32 N=5000."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "so-net_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from sonet_b200 import autoencoder, classifier, networks, ops, segmenter, synth  # noqa: E402


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def build(task, B, N):
    opt = synth.make_opt(task, batch_size=B, input_pc_num=N, device="cuda:0")
    cpu = synth.make_opt(task, batch_size=B, input_pc_num=N)
    inp = synth.synth_inputs(B, N, seed=0)
    mod = {"classifier": classifier, "segmenter": segmenter, "autoencoder": autoencoder}[task]
    m = mod.Model(opt)
    m.encoder.load_state_dict(synth.synth_state_dict(networks.Encoder(cpu), seed=1))
    if task == "segmenter":
        m.segmenter.load_state_dict(synth.synth_state_dict(networks.Segmenter(cpu), seed=2))
        m.set_input(inp["pc"], inp["sn"], inp["label"], torch.zeros(B, N, dtype=torch.int64),
                    inp["node"], inp["node_knn_I"])
    else:
        m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
    return m


out = {}
for name, task, B, N in (("cfg1_classifier_B8_N1024", "classifier", 8, 1024),
                         ("cfg3_segmenter_B32_N1024", "segmenter", 32, 1024),
                         ("cfg4_autoencoder_chamfer_B32_N5000", "autoencoder", 32, 5000)):
    m = build(task, B, N)
    ms = timed(m.test_model)
    if task == "classifier":          # the small config is launch-bound when run eagerly
        m.enable_cuda_graph(True)
        ms_graph = timed(m.test_model)
        m.enable_cuda_graph(False)
        print(name, "with CUDA-graph replay:", round(ms_graph, 4), "ms", round(B / ms_graph * 1e3, 1),
              "clouds/s")
    ops.PROFILE = []
    m.test_model()
    torch.cuda.synchronize()
    prof = [(n.replace("sonet_", ""), round(a.elapsed_time(b), 4)) for n, a, b, _ in ops.PROFILE]
    ops.PROFILE = None
    out[name] = {"ms_per_step": round(ms, 4), "clouds_per_s": round(B / ms * 1e3, 1), "kernels": prof}
    print(name, out[name]["ms_per_step"], "ms", out[name]["clouds_per_s"], "clouds/s")
    print("   ", prof)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)
