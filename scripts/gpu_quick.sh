#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in 0 1; do
echo "== MT3_LOGMEL_CFG=$cfg"
MT3_LOGMEL_CFG=$cfg timeout 300 python -m pytest tests -q -m gpu -x -k "logmel and not other_fft" 2>&1 | tail -2
MT3_LOGMEL_CFG=$cfg timeout 300 python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
from mt3_b200 import spectrograms
cfg = spectrograms.SpectrogramConfig()
for S in (64, 293):
    a = torch.from_numpy((0.1 * np.random.default_rng(0).standard_normal((S, 32768))).astype(np.float32)).cuda()
    spectrograms.compute_spectrogram(a, cfg); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): spectrograms.compute_spectrogram(a, cfg)
    e1.record(); torch.cuda.synchronize()
    print(f"  logmel {S} segments: {1000 * e0.elapsed_time(e1) / 20:.1f} us")
PY
done
