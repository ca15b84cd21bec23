import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mt3_b200 import _lib, inference
dev = torch.device("cuda:0")
im = inference.InferenceModel("synthetic:0", "mt3", device=dev, batch_size=64)
audio = torch.from_numpy((0.1 * np.random.default_rng(0).standard_normal((64, 32768))).astype(np.float32))
im.transcribe_segments(audio, num_steps=4, stop_at_eos=False)
lib = _lib.load(); h = im.model._h; st = torch.cuda.current_stream(dev).cuda_stream
def t(kind, pos, iters=64):
    _lib.check(lib.mt3_debug_launch(h, kind, pos, 8, st)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(lib.mt3_debug_launch(h, kind, pos, iters, st)); e1.record(); torch.cuda.synchronize()
    return 1000 * e0.elapsed_time(e1) / iters
for rows in ("64", "56", "32", "8"):
    os.environ["MT3_CHAIN_ROWS"] = rows
    print("rows", rows, "full %.2f us  stream-only %.2f us  compute-only %.2f us" % (t(5, 0), t(5, 1), t(5, 2)))
