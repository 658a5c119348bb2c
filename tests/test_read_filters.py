"""Parity of the read-level filter cascade: device labels vs. the oracle's per-stage labels (bit-exact)."""
import numpy as np
import pytest
import worldutil


def check_read_filters(world, lib_path):
    ctx = worldutil.context_from_oracle(world, lib_path)
    ctx.run_read_filters()
    got, early = ctx.fragment_filters()
    want = world.stage("rf_low_entropy")["frag_filter"]
    want_early = world.stage("rf_low_coverage_viral_contigs")["frag_filter"]
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, "label mismatch at fragments %s: got %s want %s" % (bad[:10], got[bad[:10]], want[bad[:10]])
    assert np.array_equal(early, want_early)
    counts = ctx.filter_counts()
    assert np.array_equal(counts, np.bincount(want, minlength=38)[:38])
    assert int(counts[0]) == int(world.stage("rf_low_entropy")["remaining"][0])
    # the world must actually exercise the rules
    assert (want == 1).sum() > 0 and (want == 10).sum() > 0 and (want == 36).sum() > 0
    ctx.close()


def test_read_filters_hostsim(worlds, hostsim_lib):
    check_read_filters(worlds.get("small"), hostsim_lib)


def test_read_filters_hostsim_l151_shuffled(worlds, hostsim_lib):
    check_read_filters(worlds.get("l151", read_length=151, seed=7, extra=("--shuffle", "--varnames")), hostsim_lib)


@pytest.mark.gpu
def test_read_filters_cuda(worlds, cuda_lib):
    check_read_filters(worlds.get("small"), cuda_lib)


@pytest.mark.gpu
def test_read_filters_cuda_l151(worlds, cuda_lib):
    check_read_filters(worlds.get("l151", read_length=151, seed=7, extra=("--shuffle", "--varnames")), cuda_lib)
