"""Host-side logic of the drop-in path that needs no GPU: the read filter a host-side SAM parser may apply before it hands 8-byte records to
abw_coverage_batch (abawaca-build.cpp:546-551), the search problem derived from the window counts (ScafDpData.cpp:92-93), the
round-robin dimension shards, and the launch-list summariser used for profiles/."""
import gzip
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_side_read_filter_keeps_exactly_the_reads_the_reference_counts(oracle):
    """compact_reads drops unmapped reads, secondary alignments, reads with more than max_snps mismatches and reads on unknown scaffolds: the oracle
    (pinned on the reference) gives the same feature rows and the same per-scaffold read bases for the kept reads as for all reads."""
    from abawaca_b200 import capi, pipeline, synth
    mg = synth.make_metagenome(300, 2, 3, 17, shuffle_reads=True, bad_read_frac=0.2)
    reads = [r.copy() for r in mg.reads]
    reads[1]["scaf"][::53] = 0xFFFFFFFF                       # reference name not in the assembly
    reads[1]["len"] = np.random.default_rng(3).integers(30, 300, reads[1].size)
    full = oracle.build_features(mg.seq, mg.offsets, reads, this_sample=1)
    kept = []
    for r in reads:
        c = pipeline.compact_reads(r, mg.nscaf)
        back = np.zeros(c.n, dtype=capi.READ_DTYPE)
        back["scaf"], back["pos0"] = c.recs["scaf"], c.recs["pos0"]
        back["len"] = c.len16 if c.len16 is not None else c.length
        kept.append(back)                                       # flag 0, no mismatches: every kept read counts
        flag, nsnps = r["flag_nsnps"] & 0xFFFF, r["flag_nsnps"] >> 16
        assert c.n == int((((flag & 0x104) == 0) & (nsnps <= 15) & (r["scaf"] < mg.nscaf)).sum()) and 0 < c.n < r.size
    filtered = oracle.build_features(mg.seq, mg.offsets, kept, this_sample=1)
    assert np.array_equal(full["rows"], filtered["rows"]) and np.array_equal(full["info_cvg"], filtered["info_cvg"])
    # one length for the sample -> no per-read lengths travel; mixed lengths -> uint16 per read; too long -> the full format is demanded
    assert pipeline.compact_reads(reads[0], mg.nscaf).len16 is None and pipeline.compact_reads(reads[1], mg.nscaf).len16.dtype == np.uint16
    big = reads[0][:4].copy()
    big["flag_nsnps"] = 0
    big["len"] = [100, 70000, 100, 100]
    try:
        pipeline.compact_reads(big, mg.nscaf)
        raise AssertionError("a 70 kb read was accepted into the 16-bit length format")
    except ValueError:
        pass


def test_search_problem_drops_single_window_scaffolds():
    from abawaca_b200 import pipeline
    counts = np.array([3, 1, 2, 0, 5, 1])
    keep, dp2scaf, T, kept = pipeline.search_problem_from_counts(counts)
    assert kept.tolist() == [0, 2, 4] and T.tolist() == [3, 2, 5]
    assert keep.tolist() == [True] * 3 + [False] + [True] * 2 + [True] * 5 + [False]
    assert dp2scaf.tolist() == [0] * 3 + [1] * 2 + [2] * 5


def test_round_robin_shards_cover_every_dimension_once():
    for D in (1, 7, 180, 189, 229):
        for world in (1, 2, 3, 4, 8):
            owned = [list(range(r, D, world)) for r in range(world)]
            assert sorted(d for o in owned for d in o) == list(range(D))
            assert [len(o) for o in owned] == [(D - r + world - 1) // world if D > r else 0 for r in range(world)]     # cnt(q) of csrc/peer.cu
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_launch_shares_of_the_committed_launch_list(tmp_path):
    src = os.path.join(ROOT, "profiles", "r02_launches_ncu_gputime.csv.gz")
    csv = tmp_path / "launches.csv"
    csv.write_bytes(gzip.open(src).read())
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_shares.py"), str(csv), "1"], capture_output=True, text=True, check=True).stdout
    assert "launches," in out.splitlines()[0] and "| k_partition2 | 13 |" in out and "| k_sweep_ss | 13 |" in out
    shares = [float(l.split("|")[4]) for l in out.splitlines() if l.startswith("| k_")]
    assert 0.95 < sum(shares) < 1.02                        # shares are printed with three decimals


def test_run_ahead_rule_of_the_sharded_search_enqueues_the_same_levels_on_every_rank():
    """The rule of abw_search_run_sharded (csrc/search.cu): level l is enqueued iff the search had not ended by level l - lead, decided only after
    level l - lead has been reported.  Restated here and driven with ranks that observe the device's progress words (ticks, done_at) at arbitrary
    moments: whatever the timing, every rank enqueues exactly (index of the last level) + lead levels, so the collective counts match."""
    rng = np.random.default_rng(7)

    def levels_enqueued(last_level, lead, observe):
        # device: level i completes at time i + 1 (ticks = i + 1); done_at = last_level + 1 from then on
        lvl = 0
        while True:
            lvl += 1                                            # level lvl - 1 has just been enqueued
            if lvl < lead:
                continue
            need = lvl - lead + 1
            t = max(observe(), need)                            # the host polls until ticks >= need; it may look later than that
            done_at = last_level + 1 if t >= last_level + 1 else 0
            if done_at != 0 and done_at <= need:
                return lvl
    for lead in (1, 2, 3, 5, 8):
        for last_level in (0, 1, 2, 7, 12, 30):
            counts = {levels_enqueued(last_level, lead, lambda: int(rng.integers(0, 50))) for _ in range(200)}
            assert counts == {last_level + lead}, (lead, last_level, counts)
