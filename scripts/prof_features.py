"""Small driver for profiling the feature-stage kernels alone (no reads): python scripts/prof_features.py [n_scaffolds]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abawaca_b200 import capi, pipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
mg = synth.make_metagenome(n, 0, 16, 99, with_reads=False)
ctx = capi.Context(0)
for i in range(3):
    t = {}
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, [], timings=t)
    print(i, fb.nseg, {k: round(v, 3) for k, v in t.items()})
    fb.close()
