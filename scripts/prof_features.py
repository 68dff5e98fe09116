"""Small driver for profiling the feature-stage kernels alone (no reads), sequence resident on the device: python scripts/prof_features.py [n_scaffolds]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abawaca_b200 import capi, pipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
mg = synth.make_metagenome(n, 0, 16, 99, with_reads=False)
ctx = capi.Context(0)
d_seq = ctx.alloc(mg.seq.size + 64)
ctx.to_device(d_seq, mg.seq)
for i in range(3):
    t = {}
    fb = pipeline.build_features(ctx, d_seq, mg.offsets, [], timings=t, seq_on_device=True)
    print(i, fb.nseg, mg.seq.size, {k: round(v, 3) for k, v in t.items()})
    fb.close()
ctx.profile(True)
fb = pipeline.build_features(ctx, d_seq, mg.offsets, [], seq_on_device=True)
fb.close()
rep = ctx.profile_report()
ctx.profile(False)
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:44s} {c:4d} {ms * 1000:9.1f} us")
