"""The debug allocator that stands in for compute-sanitizer memcheck (csrc/context.cu, ABW_REDZONE): it must see a write one element past a block and a
write before it, and stay silent for a clean feature build + search.  Runs in a child process because the mode is read once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
from abawaca_b200 import capi, pipeline, synth
L = capi.load()
ctx = capi.Context(0)
# clean work: nothing may be flagged
mg = synth.make_metagenome(300, 2, 3, 5)
fb = pipeline.build_features(ctx, mg.seq, mg.offsets, mg.reads, this_sample=0)
res, _kept = pipeline.search_features(ctx, fb, np.diff(mg.offsets.astype(np.int64)).astype(np.uint64), mg.scg_masks())
fb.close()
ctx.synchronize()
clean = int(L.abw_redzone_violations())
# a fresh block is handed out filled with 0xFF
d = ctx.alloc(1000)
back = np.zeros(1000, dtype=np.uint8)
ctx.to_host(back, d)
fresh = bool((back == 0xFF).all())
# one byte past the end (the payload is rounded up to 16 bytes: 1008), then one byte before the start
ctx.memset(d, 0, 1009)
ctx.free(d)
after = int(L.abw_redzone_violations())
d = ctx.alloc(64)
ctx.memset(d - 1, 0, 1)
ctx.free(d)
before = int(L.abw_redzone_violations())
ctx.close()
print("RESULT", clean, fresh, after, before, len(res.recs))
"""


def test_redzone_allocator_flags_overruns_and_passes_clean_work():
    env = dict(os.environ, ABW_REDZONE="1")
    out = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "0" and line[2] == "True" and line[3] == "1" and line[4] == "2" and int(line[5]) >= 1
    assert out.stderr.count("ABW_REDZONE: block") == 2
