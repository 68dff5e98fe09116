"""GPU parity at BASELINE.json's full configs[1] size (50 000 scaffolds, 10 samples, 0.5 Gbp, 145 M read records): the CUDA path through
the C ABI against the flat-array oracle (itself pinned bit for bit to the unmodified reference, tests/test_oracle_vs_reference.py),
everything bit-exact -- window table, all 227 k x 189 feature values, every evaluated cluster record, final bins -- plus
size-independent properties of the result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2():
    from abawaca_b200 import synth
    return synth.make_metagenome(**synth.CONFIGS["cfg2"], q6_reads=True)


def test_cfg2_features_and_search_match_the_oracle(cfg2, oracle):
    from abawaca_b200 import capi, pipeline
    mg = cfg2
    ctx = capi.Context(0)
    try:
        fb = pipeline.build_features(ctx, mg.seq, mg.offsets, mg.reads, this_sample=0)
        rows = fb.rows_host()
        sg = fb.segments_host()
        f = oracle.build_features(mg.seq, mg.offsets, mg.reads, this_sample=0)
        assert np.array_equal(sg["seg_scaf"], f["seg_scaf"]) and np.array_equal(sg["seg_start"], f["seg_start"]) and np.array_equal(sg["seg_end"], f["seg_end"])
        assert rows.shape == f["rows"].shape
        assert np.array_equal(rows, f["rows"])
        st = fb.scaffold_stats_host(np.diff(mg.offsets.astype(np.int64)))
        assert np.array_equal(st["cvg"], f["info_cvg"]) and np.array_equal(st["gc"], f["info_gc"]) and np.array_equal(st["Ns"], f["info_Ns"])
        # properties that hold at any size: every value is a multiple of 0.001 in [0, 1] (k-mer) / >= 0 (coverage); C + (A) = 1 before truncation
        assert np.all(rows[:, :179] >= 0) and np.all(rows[:, :179] <= 1) and np.all(rows[:, 179:] >= 0)
        assert np.array_equal(rows, np.round(rows * 1000) / 1000)
        keep, dp2scaf, T, kept = pipeline.search_problem_from_counts(np.diff(fb.seg_first_host().astype(np.int64)))
        length = np.diff(mg.offsets.astype(np.int64)).astype(np.uint64)[kept]
        mask = mg.scg_masks()[kept]
        row_of_dp = None if keep.all() else np.nonzero(keep)[0].astype(np.uint64)
        res = pipeline.search(ctx, fb.d_rows, dp2scaf, T, length, mask, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols,
                              row_of_dp=row_of_dp)
        fb.close()
        vals = np.ascontiguousarray(rows[keep].T)
        orecs, odp, osc = oracle.Search(vals, dp2scaf, T, length, mask).run()
        assert len(orecs) == len(res.recs)
        for o, r in zip(orecs, res.recs):
            assert (o.id, o.parent, o.ndps, o.nscafs, o.split) == (r.id, r.parent, r.ndps, r.nscafs, r.split)
            if o.split:
                assert (o.best.dim, o.best.value, o.best.a, o.best.b, o.best.legal) == (r.best.dim, r.best.value, r.best.a, r.best.b, r.best.legal)
                assert (o.child1, o.child2, o.child1_ndps, o.child2_ndps, o.child1_nscafs, o.child2_nscafs, o.child1_raw, o.child2_raw) == \
                       (r.child1, r.child2, r.child1_ndps, r.child2_ndps, r.child1_nscafs, r.child2_nscafs, r.child1_raw, r.child2_raw)
            else:
                assert (o.total_size, o.scg_unique, o.scg_avg) == (r.total_size, r.scg_unique, r.scg_avg)
        assert odp.tolist() == res.dp2cluster.tolist()
        assert osc.tolist() == res.scaf2cluster.tolist()
        # the 32 synthetic genomes come back as 32 bins, each scaffold in the bin of its genome
        bins = res.scaf2cluster
        genome = mg.genome[kept]
        assert len(set(bins.tolist()) - {0}) == 32
        for b in set(bins.tolist()) - {0}:
            assert len(set(genome[bins == b].tolist())) == 1
        # idempotence: a final bin, searched on its own, is terminal (no legal separation left inside it)
        b = int(np.bincount(bins).argmax())
        sel_scaf = np.nonzero(bins == b)[0]
        sel_dp = np.nonzero(np.isin(dp2scaf, sel_scaf))[0]
        remap = np.full(T.size, -1, dtype=np.int64)
        remap[sel_scaf] = np.arange(sel_scaf.size)
        sub = pipeline.search(ctx, np.ascontiguousarray(vals[:, sel_dp]), remap[dp2scaf[sel_dp]].astype(np.uint32), T[sel_scaf], length[sel_scaf], mask[sel_scaf])
        assert len(sub.recs) == 1 and sub.recs[0].split == 0
    finally:
        ctx.close()
