"""world_size-2 gloo tests (CPU) of the data-parallel plumbing used by bench.py / N > 1 runs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mt3_b200 import distributed as D


def test_shard_range_covers_everything_in_order():
    for s in (0, 1, 5, 64, 88, 512, 513):
        for n in (1, 2, 3, 4, 8):
            spans = [D.shard_range(s, r, n) for r in range(n)]
            flat = [i for lo, hi in spans for i in range(lo, hi)]
            assert flat == list(range(s))
            assert max(hi - lo for lo, hi in spans) == (-(-s // n) if s else 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, s_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weight broadcast: rank 0 holds the blob, the other starts empty
        blob = torch.arange(1000, dtype=torch.float32) * 0.5 if rank == 0 else torch.zeros(1000)
        D.broadcast_weights(blob, src=0)
        ok_w = bool(torch.equal(blob, torch.arange(1000, dtype=torch.float32) * 0.5))
        # parameter-dict form used by InferenceModel.restore_from_checkpoint: only rank 0 has the weights
        from mt3_b200 import network, weights
        cfg = network.T5Config(vocab_size=128, emb_dim=64, num_heads=2, num_encoder_layers=1, num_decoder_layers=1,
                               head_dim=64, mlp_dim=128, mlp_activations=('gelu', 'linear'))
        want = weights.synthetic_params(cfg, 3)
        got = D.broadcast_params(want if rank == 0 else None, cfg, torch.device("cpu"), src=0)
        ok_w = ok_w and list(got) == list(want) and all(np.array_equal(got[k], want[k]) for k in want)
        # each rank "decodes" its shard: token[i, :] = global segment index
        lo, hi = D.shard_range(s_total, rank, world)
        local = torch.arange(lo, hi, dtype=torch.int32)[:, None].repeat(1, 16)
        allt = D.gather_tokens(local, s_total)
        ok_t = bool(torch.equal(allt, torch.arange(s_total, dtype=torch.int32)[:, None].repeat(1, 16)))
        q.put((rank, ok_w, ok_t, tuple(allt.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("s_total", [8, 5])
def test_broadcast_and_gather_world2(s_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, s_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_w, ok_t, shape in res:
        assert ok_w and ok_t and shape == (s_total, 16), (rank, ok_w, ok_t, shape)


def test_single_process_passthrough():
    t = torch.arange(12, dtype=torch.int32).reshape(3, 4)
    assert torch.equal(D.gather_tokens(t, 3), t)
    assert D.world() == (0, 1)
