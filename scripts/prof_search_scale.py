"""Cost of abw_search_create / abw_search_run on ONE GPU for the problem a rank of an N-GPU run sees: all datapoints of N x 50k scaffolds,
D/N dimensions.  The cfg2 k-mer rows are replicated N times (different scaffold ids), which is enough for timing.
  ABW_TRACE=1 python scripts/prof_search_scale.py [N]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abawaca_b200 import capi, pipeline, synth
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mg = synth.make_metagenome(50000, 0, 32, 20261020, with_reads=False)
ctx = capi.Context(0)
fb = pipeline.build_features(ctx, mg.seq, mg.offsets, [])
rows = fb.rows_host()
keep, dp2scaf, T, kept = pipeline.search_problem_from_counts(np.diff(fb.seg_first_host().astype(np.int64)))
fb.close()
D = max(1, 189 // rep)
vals = np.ascontiguousarray(np.tile(rows[keep][:, :D], (rep, 1)))
S = T.size
dp2scaf_all = np.concatenate([dp2scaf + r * S for r in range(rep)]).astype(np.uint32)
T_all = np.tile(T, rep)
length = np.tile(np.diff(mg.offsets.astype(np.int64)).astype(np.uint64)[kept], rep)
mask = np.tile(mg.scg_masks()[kept], (rep, 1))
d_vals = ctx.alloc(vals.nbytes)
ctx.to_device(d_vals, vals)
print("N", vals.shape[0], "D", D, "S", T_all.size, flush=True)
for i in range(3):
    t = {}
    t0 = time.perf_counter()
    res = pipeline.search(ctx, d_vals, dp2scaf_all, T_all, length, mask, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=vals.shape[0], D=D, ld=D, timings=t)
    print(i, "total %.2f ms" % (1000 * (time.perf_counter() - t0)), {k: round(v, 2) for k, v in t.items()}, "clusters", len(res.recs), flush=True)
