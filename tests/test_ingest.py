"""Parity of the host front end (BAM ingest, annotation, dummy genes, strandedness, fragment length) with the oracle,
and of the whole file-to-candidates path."""
import numpy as np
import pytest
import worldutil
from arriba_b200 import lib as L


def frag_tables_equal(got, d, check_genes, check_pred):
    """got: Pipeline.fragments(); d: oracle fragment dump."""
    n = len(d["n_aln"])
    assert got["n_fragments"] == n, (got["n_fragments"], n)
    names_want = bytes(d["names"])
    assert got["names_blob"] == names_want, "fragment names / name order differ"
    assert np.array_equal(got["name_off"], d["name_off"].astype(np.uint64))
    want = worldutil.chunk_from_dump(d)
    assert np.array_equal(got["n_aln"], want["n_aln"])
    assert np.array_equal(got["fflags"], want["fflags"]), np.nonzero(got["fflags"] != want["fflags"])[0][:10]
    valid = np.zeros(3 * n, bool)
    for s in range(3):
        valid[s * n:(s + 1) * n] = want["n_aln"] > s
    for k in ("contig", "start", "end", "cigar_cnt"):
        bad = np.nonzero((got[k] != want[k]) & valid)[0]
        assert len(bad) == 0, (k, bad[:10] % n, bad[:10] // n, got[k][bad[:10]], want[k][bad[:10]])
    mask = 0b001011 | (0b000100 if check_genes else 0) | (0b100000 if check_pred else 0)
    ga, wa = got["aflags"] & mask, want["aflags"] & mask
    bad = np.nonzero((ga != wa) & valid)[0]
    assert len(bad) == 0, ("aflags", bad[:10] % n, bad[:10] // n, got["aflags"][bad[:10]], want["aflags"][bad[:10]])
    if check_pred:  # predicted strand only meaningful when not ambiguous
        unamb = valid & ((want["aflags"] & 32) == 0)
        bad = np.nonzero(((got["aflags"] ^ want["aflags"]) & 16 != 0) & unamb)[0]
        assert len(bad) == 0, ("predicted strand", bad[:10] % n, bad[:10] // n)
    # CIGARs and sequences, alignment by alignment
    for x in np.nonzero(valid)[0]:
        a = got["cigar"][got["cigar_off"][x]:got["cigar_off"][x] + got["cigar_cnt"][x]]
        b = want["cigar"][want["cigar_off"][x]:want["cigar_off"][x] + want["cigar_cnt"][x]]
        if not np.array_equal(a, b):
            raise AssertionError(("cigar", x % n, x // n, a, b))
    v2 = valid[:2 * n]
    assert np.array_equal(got["seq_len"][v2], want["seq_len"][v2])
    for x in np.nonzero(v2)[0][:20000]:
        nb = (int(want["seq_len"][x]) + 1) // 2
        a = got["seq"][int(got["seq_off"][x]) * 16:int(got["seq_off"][x]) * 16 + nb]
        b = want["seq"][int(want["seq_off"][x]) * 16:int(want["seq_off"][x]) * 16 + nb]
        if not np.array_equal(a, b):
            raise AssertionError(("seq", x % n, x // n))
    if check_genes:
        assert np.array_equal(got["genes_cnt"][valid], want["genes_cnt"][valid]), "gene set sizes differ"
        for x in np.nonzero(valid)[0]:
            a = got["genes"][got["genes_off"][x]:got["genes_off"][x] + got["genes_cnt"][x]]
            b = want["genes"][want["genes_off"][x]:want["genes_off"][x] + want["genes_cnt"][x]]
            if not np.array_equal(a, b):
                raise AssertionError(("genes", x % n, x // n, a, b))


def check_front_end(world, lib_path, threads):
    p = L.Pipeline(world.prefix + ".bam", world.prefix + ".gtf", world.prefix + ".fa", threads=threads, lib_path=lib_path)
    p.step(L.STEP_LOAD_REFERENCE)
    p.step(L.STEP_INGEST)
    st = p.stats()
    rca = world.stage("read_chimeric_alignments")
    assert st.mapped_reads == int(rca["mapped_reads"][0])
    assert st.n_fragments == int(rca["total"][0])
    frag_tables_equal(p.fragments(), world.stage("ingest"), check_genes=False, check_pred=False)
    # coverage windows
    off = rca["coverage_off"]
    for c in range(len(off) - 1):
        cov, starts, ends = p.coverage(c)
        want = rca["coverage"][off[c]:off[c + 1]]
        assert len(cov) == len(want), (c, len(cov), len(want))
        assert np.array_equal(cov, want), ("coverage", c, np.nonzero(cov != want)[0][:10])
        assert np.array_equal(starts, rca["fragment_starts"][off[c]:off[c + 1]]), ("starts", c)
        assert np.array_equal(ends, rca["fragment_ends"][off[c]:off[c + 1]]), ("ends", c)
    p.step(L.STEP_ANNOTATE)
    assert p.stats().strandedness == int(world.stage("strandedness")["strandedness"][0])
    frag_tables_equal(p.fragments(), world.stage("annotated"), check_genes=True, check_pred=True)
    fl = world.stage("fragment_length")
    g = p.genes()
    assert g["n_genes"] == len(fl["gene_id"])
    for k, w in (("gene_contig", "gene_contig"), ("gene_start", "gene_start"), ("gene_end", "gene_end"), ("gene_strand", "gene_strand"), ("gene_exonic_length", "gene_exonic_length")):
        assert np.array_equal(g[k], fl[w]), k
    assert np.array_equal(g["gene_flags"], fl["gene_is_dummy"] | fl["gene_is_protein_coding"] << 1)
    want_ann = worldutil.annotation_from_dump(fl, g["n_contigs"])
    for k in ("exon_gene", "exon_start", "exon_end", "exon_cds_start", "exon_cds_end", "exon_flags"):
        assert np.array_equal(g[k], want_ann[k]), k
    # device stages fed by the product's own front end
    p.step(L.STEP_UPLOAD); p.step(L.STEP_READ_FILTERS); p.step(L.STEP_FRAGMENT_LENGTH)
    ctx = p.context()
    labels, early = ctx.fragment_filters()
    assert np.array_equal(labels, world.stage("rf_low_entropy")["frag_filter"])
    st = p.stats()
    assert st.fragment_length_ok == int(fl["ok"][0])
    if st.fragment_length_ok:
        assert st.mate_gap_mean == fl["gap_mean_stddev_readlen"][0] and st.mate_gap_stddev == fl["gap_mean_stddev_readlen"][1]
        assert st.read_length_mean == fl["gap_mean_stddev_readlen"][2]
    ff = world.stage("find_fusions")
    assert st.max_mate_gap == int(ff["max_mate_gap"][0])
    p.step(L.STEP_FIND_FUSIONS)
    cand = ctx.candidates()
    assert cand["n"] == len(ff["gene1"])
    assert int((cand["filter"] == 0).sum()) == int(ff["remaining"][0])
    p.close()


def test_front_end_hostsim(worlds, hostsim_lib):
    check_front_end(worlds.get("small"), hostsim_lib, threads=1)


def test_front_end_hostsim_threads_shuffled(worlds, hostsim_lib):
    check_front_end(worlds.get("l151", read_length=151, seed=7, extra=("--shuffle", "--varnames")), hostsim_lib, threads=5)


@pytest.mark.gpu
def test_front_end_cuda(worlds, cuda_lib):
    check_front_end(worlds.get("small"), cuda_lib, threads=8)


@pytest.mark.parametrize("threads,min_bytes", [(7, "20000"), (16, "1000"), (3, "300")])
def test_front_end_hostsim_parallel_record_scan(worlds, hostsim_lib, monkeypatch, threads, min_bytes):
    """Record boundaries found piecewise from guessed starts (verified against the true chain; csrc/bamscan.cu), down to pieces smaller than a record."""
    monkeypatch.setenv("ARB_SCAN_MIN_BYTES", min_bytes)
    monkeypatch.setenv("ARB_BAM_SCAN_PIECE", str(max(50, int(min_bytes) // 4)))
    check_front_end(worlds.get("l151", read_length=151, seed=7, extra=("--shuffle", "--varnames")), hostsim_lib, threads=threads)


@pytest.mark.parametrize("threads,min_bytes", [(6, "5000"), (13, "100")])
def test_front_end_hostsim_parallel_block_table_compressed(worlds, hostsim_lib, monkeypatch, threads, min_bytes):
    """The BGZF block table is found piecewise too (guessed block starts, verified against the true chain): deflate-compressed blocks, whose payload is
    arbitrary bytes, and pieces smaller than one block."""
    monkeypatch.setenv("ARB_SCAN_MIN_BYTES", min_bytes)
    check_front_end(worlds.get("zsmall", seed=5, extra=("--compress", "6")), hostsim_lib, threads=threads)


def test_block_table_errors_are_the_serial_ones(worlds, hostsim_lib, monkeypatch, tmp_path):
    """A file cut inside a block, or with a damaged block header in the middle, fails with the message of the one-thread walk."""
    from arriba_b200 import lib as L
    world = worlds.get("small")
    raw = open(world.prefix + ".bam", "rb").read()
    damaged = bytearray(raw); mid = raw.find(b"\x1f\x8b\x08\x04", len(raw) // 2); damaged[mid + 1] = 0
    cases = {"cut": raw[:len(raw) * 2 // 3], "damaged": bytes(damaged)}
    for name, data in cases.items():
        path = str(tmp_path / (name + ".bam")); open(path, "wb").write(data)
        messages = []
        for min_bytes, threads in (("1000000000", 1), ("1000", 9)):
            monkeypatch.setenv("ARB_SCAN_MIN_BYTES", min_bytes)
            p = L.Pipeline(path, world.prefix + ".gtf", world.prefix + ".fa", threads=threads, lib_path=hostsim_lib)
            p.step(L.STEP_LOAD_REFERENCE)
            with pytest.raises(L.ArbError) as err:
                p.step(L.STEP_INGEST)
            messages.append(str(err.value)); p.close()
        assert messages[0] == messages[1] and "failed to load alignments" in messages[0], (name, messages)


@pytest.mark.parametrize("threads,chunk_bytes,extra", [(4, "70000", ()), (3, "300000", ("--compress", "6")), (9, "1", ("--shuffle", "--varnames"))])
def test_front_end_hostsim_many_chunks(worlds, hostsim_lib, monkeypatch, threads, chunk_bytes, extra):
    """The file is read in chunks (128 MiB in production): records that straddle a chunk boundary are carried over, mates that wait for their partner
    survive the recycling of the buffer, and the workers' pools are reserved ahead after the first chunk. Here: chunks of one to a few BGZF blocks."""
    monkeypatch.setenv("ARB_CHUNK_BYTES", chunk_bytes)
    name = "chunks_" + "_".join(x.strip("-") for x in extra) if extra else "small"
    kw = dict(seed=5, extra=extra) if extra else {}
    check_front_end(worlds.get(name, **kw), hostsim_lib, threads=threads)


def test_pools_are_kept_between_samples_and_can_be_released(worlds, hostsim_lib, tmp_path):
    """the ingest keeps the workers' pools of a finished sample (csrc/host/ingest.cpp, take_workers / put_workers): a second sample in the same process, a third
    after arb_release_host_memory, and one with another number of workers must all read the same table"""
    import numpy as np
    from arriba_b200 import lib as L
    world = worlds.get("small")
    def table(threads):
        p = L.Pipeline(world.prefix + ".bam", world.prefix + ".gtf", world.prefix + ".fa", threads=threads, lib_path=hostsim_lib,
                       output=str(tmp_path / "f.tsv"), discarded=str(tmp_path / "d.tsv"))
        p.step(L.STEP_LOAD_REFERENCE); p.step(L.STEP_INGEST)
        f = p.fragments()
        mine = ("n_aln", "fflags", "contig", "start", "end", "cigar_off", "cigar_cnt", "seq_off", "seq_len", "cigar", "seq", "name_off")   # what the ingest itself fills (the gene columns belong to the annotation)
        out = {k: np.array(f[k], copy=True) for k in mine}
        out["names"] = np.frombuffer(f["names_blob"], dtype=np.uint8).copy()
        p.close()
        return out
    first = table(4)
    again = table(4)
    L.load(hostsim_lib).arb_release_host_memory()
    after_release = table(4)
    other_width = table(3)
    for other in (again, after_release, other_width):
        assert sorted(other) == sorted(first)
        for k in first:
            assert np.array_equal(first[k], other[k]), k
