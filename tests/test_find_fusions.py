"""Parity of candidate generation (find_fusions): candidate table, supporting lists and mate canonicalisation vs. the oracle."""
import numpy as np
import pytest
import worldutil


def key_of(t, i):
    return (int(t["gene1"][i]), int(t["gene2"][i]), int(t["contig1"][i]), int(t["contig2"][i]), int(t["breakpoint1"][i]), int(t["breakpoint2"][i]),
            int(t["direction1"][i]), int(t["direction2"][i]))


def check_find_fusions(world, lib_path):
    ctx = worldutil.context_from_oracle(world, lib_path)
    # start from the oracle's labels so that this test isolates find_fusions
    ctx.set_fragment_filters(world.stage("rf_low_entropy")["frag_filter"])
    ff = world.stage("find_fusions")
    ctx.find_fusions(int(ff["max_mate_gap"][0]))
    got = ctx.candidates()
    assert got["n"] == len(ff["gene1"])
    want_index = {key_of(ff, i): i for i in range(len(ff["gene1"]))}
    assert len(want_index) == len(ff["gene1"])
    n_checked_lists = 0
    for i in range(got["n"]):
        j = want_index[key_of(got, i)]
        ctxmsg = "candidate %s" % (key_of(got, i),)
        for name in ("split_reads1", "split_reads2", "discordant_mates", "filter"):
            assert int(got[name][i]) == int(ff[name][j]), (ctxmsg, name, int(got[name][i]), int(ff[name][j]))
        assert int(got["anchor_start1"][i]) == int(ff["anchor_start1"][j]) and int(got["anchor_start2"][i]) == int(ff["anchor_start2"][j]), ctxmsg
        b = int(got["bits"][i])
        want_bits = (int(ff["exonic1"][j]) | int(ff["exonic2"][j]) << 1 | int(ff["spliced1"][j]) << 2 | int(ff["spliced2"][j]) << 3 |
                     int(ff["predicted_strand1"][j]) << 4 | int(ff["predicted_strand2"][j]) << 5 | int(ff["predicted_strands_ambiguous"][j]) << 6 |
                     int(ff["transcript_start"][j]) << 7)
        assert b == want_bits, (ctxmsg, bin(b), bin(want_bits))
        assert int(got["bits2"][i]) == int(ff["transcript_start_ambiguous"][j]), ctxmsg
        for lname, oname in (("list1", "list1_off"), ("list2", "list2_off"), ("listd", "listd_off")):
            g = got[lname][got[oname][i]:got[oname][i + 1]]
            w = ff[lname][ff[oname][j]:ff[oname][j + 1]]
            assert np.array_equal(g, w), (ctxmsg, lname, g[:10], w[:10])
            n_checked_lists += len(w)
    assert n_checked_lists > 0
    # candidates are numbered in the reference's first-insertion order: first supporting fragment ascending
    # mate canonicalisation (fusions.cpp:416-421)
    ann = world.stage("annotated")
    slot0_before = ann["start"][ann["aln_off"][:-1]]
    swapped_want = (ff["slot0_start"] != slot0_before) | (ff["slot0_contig"] != ann["contig"][ann["aln_off"][:-1]])
    swapped_got = ctx.slot_swaps().astype(bool)
    # a swap of two mates with identical start and contig is not observable in the dump
    assert np.all(swapped_got[swapped_want]), "missing canonicalisation"
    extra = np.nonzero(swapped_got & ~swapped_want)[0]
    for i in extra:
        a = ann["aln_off"][i]
        assert ann["start"][a] == ann["start"][a + 1] and ann["contig"][a] == ann["contig"][a + 1]
    assert int((got["filter"] == 0).sum()) == int(ff["remaining"][0])
    ctx.close()


def test_find_fusions_hostsim(worlds, hostsim_lib):
    check_find_fusions(worlds.get("small"), hostsim_lib)


def test_find_fusions_hostsim_l151_shuffled(worlds, hostsim_lib):
    check_find_fusions(worlds.get("l151", read_length=151, seed=7, extra=("--shuffle", "--varnames")), hostsim_lib)


@pytest.mark.gpu
def test_find_fusions_cuda(worlds, cuda_lib):
    check_find_fusions(worlds.get("small"), cuda_lib)


@pytest.mark.gpu
def test_find_fusions_cuda_l151(worlds, cuda_lib):
    check_find_fusions(worlds.get("l151", read_length=151, seed=7, extra=("--shuffle", "--varnames")), cuda_lib)
