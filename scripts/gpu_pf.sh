#!/bin/bash
# One gpurun call for the L2-prefetch study: (1) scripts/ab_prefetch.py -- every knob setting in one process, token streams
# must equal the baseline's; (2) the best setting (>= 1 % faster than base, else none) is exported and the whole `-m gpu`
# suite, the default bench line and a decode-step timeline run under it.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== A/B"; timeout 240 python scripts/ab_prefetch.py > gpurun_out/ab_prefetch.txt 2> gpurun_out/ab_prefetch.err
tail -22 gpurun_out/ab_prefetch.txt; tail -3 gpurun_out/ab_prefetch.err
python - <<'PY' > gpurun_out/chosen_env.sh
import json
rows = []
for line in open("gpurun_out/ab_prefetch.txt"):
    if line.startswith("{"):
        rows.append(json.loads(line))
ok = [r for r in rows if r["tokens_equal_base"]]
base = [r for r in rows if r["setting"].startswith("base")]
if ok and base:
    b = min(r["ms_mean"] for r in base)
    best = min(ok, key=lambda r: r["ms_mean"])
    if best["ms_mean"] < 0.99 * b and best["env"]:
        print("# chosen: %s (%.2f ms vs base %.2f)" % (best["setting"], best["ms_mean"], b))
        for kv in best["env"].split(","):
            print("export " + kv)
    else:
        print("# no setting is 1 %% faster than base (%.2f ms); best was %s %.2f" % (b, best["setting"], best["ms_mean"]))
else:
    print("# A/B produced no usable rows")
PY
cat gpurun_out/chosen_env.sh
source gpurun_out/chosen_env.sh
echo "== pytest -m gpu (under the chosen setting)"; timeout 420 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench (under the chosen setting)"
timeout 300 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
cut -c1-300 gpurun_out/bench_default.json; grep -E "timed|e2e|cpu port" gpurun_out/bench_default.err
echo "== decode-step timeline"; TRACE_POS=512 timeout 120 python scripts/trace_step.py > gpurun_out/trace_step.log 2>&1; tail -14 gpurun_out/trace_step.log
