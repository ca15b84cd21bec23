"""torchrun check of the data-parallel API on real GPUs (NCCL): rank 0 owns the weights, one broadcast
at load; every rank transcribes its shard; one all-gather; rank 0 compares with a single-GPU run of the
whole list (must be identical: sequences never interact)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mt3_b200 import inference  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    im = inference.InferenceModel("synthetic:0" if rank == 0 else None, "mt3", device=dev, batch_size=8)
    rng = np.random.default_rng(7)
    S = 4 * world + 1                                   # ragged: the last rank gets a short shard
    audio = (0.1 * rng.standard_normal((S, 256 * 128))).astype(np.float32)
    toks = im.transcribe_segments_sharded(audio, num_steps=48, stop_at_eos=False, decoded=False)
    assert toks.shape == (S, 1024), toks.shape
    ok = True
    if rank == 0:
        # single-GPU run of everything, outside the process group's sharding
        ref = im.transcribe_segments(audio, num_steps=48, stop_at_eos=False, decoded=False)
        ok = bool(np.array_equal(ref, toks))
        print(f"dist_check world={world} S={S}: sharded == single-GPU: {ok}; first tokens {toks[0, :6].tolist()}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("DIST_CHECK FAILED")
    if rank == 0:
        print("DIST_CHECK PASSED")


if __name__ == "__main__":
    main()
