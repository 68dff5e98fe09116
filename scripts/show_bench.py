"""Print the interesting parts of a bench.py JSON line: python scripts/show_bench.py gpurun_out/x_bench.log [nkernels]"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
nk = int(sys.argv[2]) if len(sys.argv) > 2 else 24
print("value", d["value"], "ms", d["ms_per_step"], "| e2e", d["e2e"], "| launches", d.get("gpu_launches"), "| pcie", d.get("pcie"))
print("phase", d.get("phase_wall_ms"))
print("search_profile", d.get("search_profile_ms"))
print("resident steps", d["step_ms"]["resident"]); print("e2e steps", d["step_ms"]["e2e"])
print("roofline", {k: v for k, v in d.get("roofline", {}).items() if k not in ("peak_source", "traffic_source")})
for k, v in list(d["kernels"].items())[:nk]:
    print(f"  {k:44s} {v['launches']:4d} {v['ms']:8.4f}")
print("  total kernel ms", round(sum(v["ms"] for v in d["kernels"].values()), 3))
for key in ("cpu_baseline", "clocks", "config"):
    if key in d:
        print(key, d[key])
