#!/bin/bash
# First GPU call of round 2 (about 4 minutes of box time): is ABW_SMALL_COPIES=kernel correct on the whole GPU suite,
# and what does it do to the end-to-end step?  Usage:  gpurun --timeout 420 -- 'bash scripts/round2_first_call.sh'
# Reads afterwards: gpurun_out/r2a_*.log  (value / e2e / step_ms of the two bench lines, pytest tails).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_default.log 2>&1
ABW_SMALL_COPIES=kernel python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_kernel.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_default.log 2> gpurun_out/r2a_bench_default.err
ABW_SMALL_COPIES=kernel python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_kernel.log 2> gpurun_out/r2a_bench_kernel.err
tail -1 gpurun_out/r2a_pytest_default.log gpurun_out/r2a_pytest_kernel.log
python - <<'PY'
import json
for f in ("default", "kernel"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2a_bench_{f}.log") if l.startswith("{")][0])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "ms", d["e2e"]["ms_per_step"], "pcie", d.get("pcie"))
    except Exception as e:
        print(f, "no bench line:", e)
PY
