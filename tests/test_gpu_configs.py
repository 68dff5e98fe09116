"""GPU parity on problems SHAPED like BASELINE.json configs[2], configs[3] and configs[4], scaled so that the flat-array oracle (pinned bit for bit to the
unmodified reference, tests/test_oracle_vs_reference.py) finishes in seconds on the host:

  configs[2]  500k scaffolds x 20 samples, 128 genomes   ->  40 000 scaffolds, 20 samples, 128 genomes, reads thinned (deep tree: >= 9 levels)
  configs[3]  1M scaffolds x 50 samples, 256 genomes,
              deep recursive splitting with SCG checks    ->  20 000 scaffolds, 50 samples, 256 genomes, a third of them with two coverage modes: the SCG
                                                              test (ClusterQuality::is_split_better) decides which candidates may win; hundreds of clusters
  configs[4]  k-mer feature build over 10 Mbp - 10 Gbp    ->  k-mer rows of a 100 Mbp assembly against the oracle; a 1 Gbp assembly against the oracle on a
                                                              random subsample of its scaffolds plus the sharding invariance (halves == whole)

Everything is compared with ==: window table, all feature values, every evaluated cluster record, final bins.  The full-size configurations are run by
scripts/run_config.py (logs under profiles/)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from abawaca_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _depth(recs):
    parent = {r.id: r.parent for r in recs}
    best = 0
    for r in recs:
        d, i = 1, r.id
        while parent.get(i, 0):
            i = parent[i]
            d += 1
        best = max(best, d)
    return best


def _features_and_search(ctx, oracle, mg, use_compact):
    from abawaca_b200 import capi, pipeline
    reads = [pipeline.compact_reads(r, mg.nscaf) for r in mg.reads] if use_compact else mg.reads
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, reads, this_sample=0)
    rows = fb.rows_host()
    sg = fb.segments_host()
    f = oracle.build_features(mg.seq, mg.offsets, mg.reads, this_sample=0)
    assert np.array_equal(sg["seg_scaf"], f["seg_scaf"]) and np.array_equal(sg["seg_start"], f["seg_start"]) and np.array_equal(sg["seg_end"], f["seg_end"])
    assert np.array_equal(rows, f["rows"])
    keep, dp2scaf, T, kept = pipeline.search_problem_from_counts(np.diff(fb.seg_first_host().astype(np.int64)))
    length = np.diff(mg.offsets.astype(np.int64)).astype(np.uint64)[kept]
    mask = mg.scg_masks()[kept]
    row_of_dp = None if keep.all() else np.nonzero(keep)[0].astype(np.uint64)
    res = pipeline.search(ctx, fb.d_rows, dp2scaf, T, length, mask, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols,
                          row_of_dp=row_of_dp)
    fb.close()
    vals = np.ascontiguousarray(rows[keep].T)
    S = oracle.Search(vals, dp2scaf, T, length, mask)
    orecs, odp, osc = S.run()
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
    return res, (vals, dp2scaf, T, length, mask), orecs, osc


def test_configs2_shape_500k_x_20_scaled(ctx, oracle):
    from abawaca_b200 import synth
    mg = synth.make_metagenome(40000, 20, 128, synth.MASTER_SEED + 3, q6_reads=True, cov_lo=0.05, cov_hi=0.6)
    res, _, _, osc = _features_and_search(ctx, oracle, mg, use_compact=True)
    assert _depth(res.recs) >= 9
    assert len(set(osc.tolist()) - {0}) >= 64                 # most of the 128 genomes come back as bins of their own


def test_configs3_shape_1m_x_50_scaled_scg_checks_decide(ctx, oracle):
    from abawaca_b200 import synth
    # a third of the genomes carry two coverage modes (synth.make_metagenome: bimodal_frac): clean separations inside one genome that only the SCG test rejects
    mg = synth.make_metagenome(20000, 50, 256, synth.MASTER_SEED + 4, q6_reads=True, cov_lo=0.05, cov_hi=0.8, mean_extra=9000, bimodal_frac=0.3)
    res, (vals, dp2scaf, T, length, mask), orecs, osc = _features_and_search(ctx, oracle, mg, use_compact=False)
    assert _depth(res.recs) >= 9
    # the SCG test is decisive on this set: without SCG information (every candidate passes on the size rule or fails it) the tree is a different one
    norecs, _, nosc = oracle.Search(vals, dp2scaf, T, length, np.zeros_like(mask)).run()
    assert [(r.id, r.split, r.best.dim if r.split else 0) for r in norecs] != [(r.id, r.split, r.best.dim if r.split else 0) for r in orecs]


def test_configs4_kmer_rows_of_100_mbp(ctx, oracle):
    from abawaca_b200 import pipeline, synth
    mg = synth.make_metagenome(10000, 0, 64, synth.MASTER_SEED + 5, with_reads=False, n_run_frac=0.02)
    assert 90e6 < mg.seq.size < 115e6
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, [])
    rows = fb.rows_host()
    sg = fb.segments_host()
    fb.close()
    f = oracle.build_features(mg.seq, mg.offsets, [], this_sample=0)
    assert np.array_equal(sg["seg_start"], f["seg_start"]) and np.array_equal(sg["seg_end"], f["seg_end"])
    assert np.array_equal(rows, f["rows"])


def test_configs4_kmer_rows_of_1_gbp_subsample_and_sharding_invariance(ctx, oracle):
    from abawaca_b200 import pipeline, synth
    mg = synth.make_metagenome(100000, 0, 64, synth.MASTER_SEED + 6, with_reads=False, n_run_frac=0.01)
    assert mg.seq.size > 0.9e9
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, [])
    rows = fb.rows_host()
    first = fb.seg_first_host().astype(np.int64)
    fb.close()
    # the oracle on a random subsample of the scaffolds (windows and k-mer rows of a scaffold depend on that scaffold alone)
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(mg.nscaf, 3000, replace=False))
    lens = np.diff(mg.offsets.astype(np.int64))[pick]
    off = np.zeros(pick.size + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    seq = np.concatenate([mg.scaffold(int(i)) for i in pick])
    f = oracle.build_features(seq, off, [], this_sample=0)
    got = np.concatenate([rows[first[i]:first[i + 1]] for i in pick])
    assert np.array_equal(got, f["rows"])
    # scaffold-sharded feature build: the rows of each half, built on its own, are the rows of the whole
    half = mg.nscaf // 2
    cut = int(mg.offsets[half])
    fa = pipeline.build_features(ctx, mg.seq[:cut], mg.offsets[:half + 1], [])
    ra = fa.rows_host()
    fa.close()
    fbh = pipeline.build_features(ctx, mg.seq[cut:], mg.offsets[half:] - mg.offsets[half], [])
    rb = fbh.rows_host()
    fbh.close()
    assert np.array_equal(np.concatenate([ra, rb]), rows)
    # size-independent properties: every value is a multiple of 0.001 in [0, 1]
    assert np.array_equal(rows, np.round(rows * 1000) / 1000) and rows.min() >= 0 and rows.max() <= 1
