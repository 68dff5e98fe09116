#!/bin/bash
# One GPU call: the GPU test suite, then the bench (no CPU baseline).  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_check.sh TAG [pytest args]'
TAG=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1
tail -n 15 gpurun_out/${TAG}_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
tail -n 5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench.log") if l.startswith("{")][0])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "pcie", d.get("pcie"))
    print(d["phase_wall_ms"])
    for k, v in list(d["kernels"].items())[:14]:
        print(f"  {k:40s} {v['launches']:4d} {v['ms']:8.4f}")
except Exception as e:
    print("no bench line:", e)
PY
