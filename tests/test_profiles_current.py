"""The committed ncu captures that bench.py reads (`roofline.traffic`, hardware FP64 counts) must have been taken from the
sources in the tree: bench.py refuses a stale one at run time; this makes the same check part of the CPU suite."""
import json
import os

import pytest

import bench


@pytest.mark.parametrize("workload", ["c3", "c2", "c4", "c5"])
def test_committed_ncu_capture_matches_the_sources(workload):
    p = os.path.join(bench.ROOT, "profiles", "r02_ncu_%s.json" % workload)
    d = json.load(open(p))
    assert d["source_hash"] == bench.source_hash(), "re-run scripts/ncu_profile.py %s on a B200" % workload
    prof, src = bench.load_profile(workload)
    assert prof is not None and src.endswith("r02_ncu_%s.json" % workload)
    W = bench.Workload(workload)
    assert W.kernel in d["kernel"] and d["chains"] == W.chains
    assert d["dram_bytes"] > 0 and d["duration_ms"] > 0
