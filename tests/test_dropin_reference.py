"""Drop-in boundary: the reference's own models/classifier.py (byte-identical, imported from
/root/reference) constructs on top of sonet_b200's networks. Only runs where the reference tree
exists (the build container); the GPU box has no /root/reference."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

CODE = r"""
import sys, torch
sys.path.insert(0, %(pkg)r)
import sonet_b200.install as inst
from sonet_b200 import synth, networks
inst.install(%(ref)r)
from models import classifier, segmenter, autoencoder   # the reference's files
import models.networks as n
assert n is networks and classifier.networks is networks
assert classifier.__file__.startswith(%(ref)r)
opt = synth.make_opt("classifier", batch_size=2, input_pc_num=64)
m = classifier.Model(opt)
assert type(m.encoder).__module__ == "sonet_b200.networks"
inp = synth.synth_inputs(2, 64)
m.set_input(inp["pc"], inp["sn"], inp["label"], inp["node"], inp["node_knn_I"])
try:
    m.test_model()
    raise SystemExit("expected the CUDA-only encoder to refuse CPU tensors")
except RuntimeError as e:
    assert "no CPU fallback" in str(e)
opt = synth.make_opt("segmenter", batch_size=2, input_pc_num=64)
s = segmenter.Model(opt)
assert type(s.segmenter).__module__ == "sonet_b200.networks"
opt = synth.make_opt("autoencoder", batch_size=2, input_pc_num=64)
a = autoencoder.Model(opt)
assert type(a.chamfer_criteria).__module__ == "sonet_b200.losses"
print("DROPIN_OK")
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="no reference tree here")
def test_reference_model_files_run_on_our_networks():
    code = CODE % dict(pkg=os.path.join(ROOT, "so-net_b200"), ref=REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "DROPIN_OK" in r.stdout, r.stdout + r.stderr
