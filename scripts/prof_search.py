"""Driver for profiling the split search alone on the configs[1] shape: python scripts/prof_search.py [n_scaffolds] [n_samples] [reps]
Features are built once; abw_search_create + abw_search_run are repeated and timed (host wall clock, stream synchronised)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abawaca_b200 import capi, pipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
genomes = int(sys.argv[4]) if len(sys.argv) > 4 else 32
mg = synth.make_metagenome(n, ns, genomes, synth.MASTER_SEED + 2, q6_reads=True)
ctx = capi.Context(0)
fb = pipeline.build_features(ctx, mg.seq, mg.offsets, [pipeline.compact_reads(r, mg.nscaf) for r in mg.reads])
counts = np.diff(fb.seg_first_host().astype(np.int64))
row_of_dp, T, kept, N = pipeline.search_rows_from_counts(counts)
if kept is None:
    kept = slice(None)
lengths = np.diff(mg.offsets.astype(np.int64)).astype(np.uint64)[kept]
masks = mg.scg_masks()[kept]
buf = {}
for i in range(reps):
    t = {}
    res = pipeline.search(ctx, fb.d_rows, None, T, lengths, masks, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols,
                          row_of_dp=row_of_dp, timings=t, buffers=buf)
    print(i, N, len(res.recs), res.profile.levels, {k: round(v, 3) for k, v in t.items()}, flush=True)
if os.environ.get("ABW_PROF_KERNELS"):
    ctx.profile(True)
    res = pipeline.search(ctx, fb.d_rows, None, T, lengths, masks, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols, row_of_dp=row_of_dp, buffers=buf)
    rep = ctx.profile_report()
    ctx.profile(False)
    for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:44s} {c:4d} {ms * 1000:9.1f} us")
fb.close()
ctx.close()
