"""Driver for profiling the coverage kernels alone on the configs[1] shape: python scripts/prof_coverage.py [n_scaffolds] [n_samples] [order]
order: scaffold (generator order) | coordinate (sorted by scaffold, position) | shuffled.  Prints CUDA-event times of abw_coverage_batch per format."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
from abawaca_b200 import capi, pipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 3
order = sys.argv[3] if len(sys.argv) > 3 else "scaffold"
mg = synth.make_metagenome(n, ns, 32, synth.MASTER_SEED + 2, q6_reads=True, shuffle_reads=(order == "shuffled"))
reads = mg.reads
if order == "coordinate":
    reads = [r[np.lexsort((r["pos0"], r["scaf"]))] for r in reads]
ctx = capi.Context(0)
import torch
ext = torch.cuda.ExternalStream(ctx.stream, device=0)
for fmt in ("full", "compact"):
    rd = reads if fmt == "full" else [pipeline.compact_reads(r, mg.nscaf) for r in reads]
    # device-resident records
    samples = []
    for r in rd:
        if fmt == "full":
            d = ctx.alloc(max(r.nbytes, 16)); ctx.to_device(d, r)
            samples.append(pipeline.ReadSample(capi.READS_FULL, r.size, d_recs=d))
        else:
            d = ctx.alloc(max(r.recs.nbytes, 16)); ctx.to_device(d, r.recs)
            samples.append(pipeline.ReadSample(capi.READS_COMPACT, r.n, length=r.length, d_recs=d))
    for i in range(3):
        t = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fb = pipeline.build_features(ctx, mg.seq, mg.offsets, samples, timings=t)
        print(fmt, order, i, fb.nseg, sum(s.n for s in samples), {k: round(v, 3) for k, v in t.items()}, flush=True)
        fb.close()
    ctx.profile(True)
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, samples)
    fb.close()
    rep = ctx.profile_report()
    ctx.profile(False)
    print("  " + "  ".join(f"{k}:{c}x{ms * 1000:.0f}us" for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]) if "cov" in k or "rs_" in k or "scan" in k))
    for s in samples:
        ctx.free(s.d_recs)
ctx.close()
