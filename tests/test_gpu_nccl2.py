"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): world_size-2 NCCL run
of the batch-sharded classifier forward — the all-gathered logits of the two ranks equal the
single-GPU forward of the full batch BIT FOR BIT (parity definition 5)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(B, N, dev):
    import sys
    for p in (ROOT, os.path.join(ROOT, "so-net_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from helpers import build_states
    from sonet_b200 import classifier, synth
    opt = synth.make_opt("classifier", batch_size=B, input_pc_num=N, device=dev,
                         gpu_id=torch.device(dev).index)
    st = build_states("classifier", opt, seed=71)
    m = classifier.Model(opt)
    m.encoder.load_state_dict(st["encoder"])
    m.classifier.load_state_dict(st["head"])
    return m, synth.synth_inputs(B, N, seed=71)


def _worker(rank, world, port, B, N, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.join(ROOT, "so-net_b200"))
    import torch.distributed as dist
    from sonet_b200 import dist as sdist
    sdist.init_from_env(backend="nccl")
    dev = "cuda:%d" % rank
    lo, hi = sdist.shard_bounds(B, rank, world)
    m, inp = _build(hi - lo, N, dev)
    keys = ("pc", "sn", "label", "node", "node_knn_I")
    m.enable_cuda_graph(True)
    for _ in range(2):                          # capture + replay
        m.set_input(*[inp[k][lo:hi] for k in keys])
        m.test_model()
    out = sdist.all_gather_rows(m.score, B)
    # the C-ABI collective (sonet_comm_init / sonet_allgather): same rows, bit for bit
    comm = sdist.SonetComm()
    out2 = torch.empty_like(out)
    comm.all_gather(m.score, out2)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    comm.destroy()
    if rank == 0:
        q.put(out.cpu())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_nccl_world2_gathered_logits_equal_single_gpu_bitwise():
    B, N = 16, 2048
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    m, inp = _build(B, N, "cuda:0")
    m.set_input(*[inp[k] for k in ("pc", "sn", "label", "node", "node_knn_I")])
    m.test_model()
    assert torch.equal(got, m.score.cpu())
