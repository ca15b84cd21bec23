#!/bin/bash
TRACE_POS=512 TRACE_ROWS=9 timeout 600 python scripts/trace_step.py 2>&1 | tee gpurun_out/trace_step_smid.log | head -24
