"""Test infrastructure: synthetic worlds, oracle runs, and conversion of oracle dumps into C-ABI inputs."""
import os, subprocess, hashlib, json
import numpy as np
import arbdump
from arriba_b200 import _build

NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
INTERESTING = set([str(i) for i in range(1, 23)] + ["X", "Y"])

# small, feature-dense default world (seconds for the oracle)
SMALL = dict(scale=0.001, genes=400, breakpoints=200, fragments=20000, read_length=101)


def run_synth(prefix, seed=0xA881BA, scale=0.001, genes=400, breakpoints=200, fragments=20000, read_length=101, extra=()):
    synth = _build.build_tools()
    cmd = [synth, "--prefix", prefix, "--seed", str(seed), "--scale", str(scale), "--genes", str(genes), "--breakpoints", str(breakpoints),
           "--fragments", str(fragments), "--read-length", str(read_length)] + list(extra)
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)


def run_oracle(prefix, outdir, dump=True, det_alloc=True, args=("-f", "blacklist"), level=2):
    oracle = _build.build_oracle()
    os.makedirs(outdir, exist_ok=True)
    env = dict(os.environ)
    if dump:
        d = os.path.join(outdir, "dump"); os.makedirs(d, exist_ok=True)
        env["ARB_DUMP_DIR"] = d; env["ARB_DUMP_LEVEL"] = str(level)
    if det_alloc:
        env["ARB_DET_ALLOC"] = "1"
    cmd = [oracle, "-x", prefix + ".bam", "-g", prefix + ".gtf", "-a", prefix + ".fa", "-o", os.path.join(outdir, "fusions.tsv"),
           "-O", os.path.join(outdir, "fusions.discarded.tsv")] + list(args)
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    with open(os.path.join(outdir, "stdout.txt"), "w") as f:
        f.write(r.stdout)
    with open(os.path.join(outdir, "stderr.txt"), "w") as f:
        f.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle failed: " + r.stderr[-2000:])
    return outdir


class World:
    def __init__(self, prefix, outdir):
        self.prefix = prefix; self.outdir = outdir
        self._dumps = None; self._fasta = None

    @property
    def dumps(self):
        if self._dumps is None:
            self._dumps = arbdump.read_dir(os.path.join(self.outdir, "dump"))
        return self._dumps

    def stage(self, name, occurrence=0):
        return arbdump.stage(self.dumps, name, occurrence)

    @property
    def fasta(self):
        if self._fasta is None:
            self._fasta = read_fasta(self.prefix + ".fa")
        return self._fasta


class WorldCache:
    def __init__(self, root):
        self.root = root; self.cache = {}

    def get(self, name="small", oracle_args=("-f", "blacklist"), **kw):
        params = dict(SMALL); params.update(kw)
        key = name + "_" + hashlib.md5(json.dumps([params, list(oracle_args)], sort_keys=True, default=str).encode()).hexdigest()[:10]
        if key not in self.cache:
            d = os.path.join(self.root, key); os.makedirs(d, exist_ok=True)
            prefix = os.path.join(d, "w")
            extra = params.pop("extra", ())
            run_synth(prefix, extra=extra, **params)
            run_oracle(prefix, os.path.join(d, "oracle"), args=oracle_args)
            self.cache[key] = World(prefix, os.path.join(d, "oracle"))
        return self.cache[key]


def read_fasta(path):
    names, seqs = [], []
    cur = []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if names:
                    seqs.append(b"".join(cur).upper())
                names.append(line[1:].split()[0].decode()); cur = []
            else:
                cur.append(line.strip())
    seqs.append(b"".join(cur).upper())
    return names, seqs


def strip_chr(name):
    if name.startswith("chr"):
        name = name[3:]
    return "MT" if name == "M" else name


def contigs_from_world(world):
    """contig id = order of first appearance (assembly first), as in the reference (assembly.cpp:42-44)."""
    names, seqs = world.fasta
    flags = np.array([1 if strip_chr(n) in INTERESTING else 0 for n in names], np.uint8)
    return names, flags, seqs


def chunk_from_dump(d):
    """oracle fragment dump (stage `annotated`) -> arb_soa_chunk arrays."""
    n = len(d["n_aln"])
    aln_off = d["aln_off"].astype(np.int64)
    A = int(aln_off[-1])
    frag_of = np.repeat(np.arange(n), np.diff(aln_off))
    slot_of = np.arange(A) - aln_off[frag_of]
    dst = slot_of * n + frag_of
    ch = {"n_fragments": n, "n_aln": d["n_aln"].copy(), "filter": np.zeros(n, np.uint8),
          "fflags": (d["single_end"] | d["multimapper"] << 1 | d["duplicate"] << 2).astype(np.uint8)}
    if "names" in d and "name_off" in d:   # bit 3: same read name (up to the last comma) as the fragment before it
        blob = bytes(d["names"]); off = d["name_off"].astype(np.int64)
        stems = [blob[off[i]:off[i + 1]].rpartition(b",")[0] or blob[off[i]:off[i + 1]] for i in range(n)]
        same = np.array([i > 0 and stems[i] == stems[i - 1] for i in range(n)], bool)
        ch["fflags"] |= (same.astype(np.uint8) << 3)
    def slot_array(src, dtype):
        out = np.zeros(3 * n, dtype); out[dst] = src; return out
    ch["contig"] = slot_array(d["contig"], np.uint16); ch["start"] = slot_array(d["start"], np.int32); ch["end"] = slot_array(d["end"], np.int32)
    af = (d["supplementary"] | d["first_in_pair"] << 1 | d["exonic"] << 2 | d["strand"] << 3 | d["predicted_strand"] << 4 | d["predicted_strand_ambiguous"] << 5)
    ch["aflags"] = slot_array(af, np.uint8)
    ch["cigar_off"] = slot_array(d["cigar_off"][:-1], np.uint32); ch["cigar_cnt"] = slot_array(np.diff(d["cigar_off"].astype(np.int64)), np.uint16)
    ch["cigar"] = d["cigar"].copy()
    ch["genes_off"] = slot_array(d["genes_off"][:-1], np.uint32); ch["genes_cnt"] = slot_array(np.diff(d["genes_off"].astype(np.int64)), np.uint16)
    ch["genes"] = d["genes"].copy() if len(d["genes"]) else np.zeros(1, np.uint32)
    # sequences: ASCII -> nt16 nibbles, each sequence 16-byte aligned; slots 0 and 1 only
    seq_off_src = d["seq_off"].astype(np.int64)
    lens = np.diff(seq_off_src)
    seq_len = slot_array(lens, np.uint32)
    assert np.all(seq_len[2 * n:] == 0), "supplementary alignments carry no sequence"
    lut = np.full(256, 15, np.uint8)
    for c, v in NT16.items():
        lut[ord(c)] = v
    codes = lut[d["seq"]]
    nbytes16 = ((lens + 1) // 2 + 15) // 16
    start16 = np.concatenate([[0], np.cumsum(nbytes16)])
    total = int(start16[-1]) * 16
    packed = np.zeros(total + 16, np.uint8)
    # position of every base
    base_aln = np.repeat(np.arange(A), lens)
    base_pos = np.arange(len(codes)) - seq_off_src[base_aln]
    byte_idx = start16[base_aln] * 16 + base_pos // 2
    hi = (base_pos % 2) == 0
    np.add.at(packed, byte_idx[hi], codes[hi] << 4)
    np.add.at(packed, byte_idx[~hi], codes[~hi])
    so = np.zeros(3 * n, np.uint32); so[dst] = start16[:-1]
    ch["seq_off"] = so[:2 * n].copy(); ch["seq_len"] = seq_len[:2 * n].astype(np.uint16); ch["seq"] = packed
    return ch


def annotation_from_dump(fl, n_contigs):
    """oracle gene/exon dump (stage `fragment_length`, level 2) -> arb_annotation arrays."""
    a = {"n_genes": len(fl["gene_id"]), "n_contigs": n_contigs}
    assert np.all(fl["gene_id"] == np.arange(a["n_genes"]))
    a["gene_contig"] = fl["gene_contig"]; a["gene_start"] = fl["gene_start"]; a["gene_end"] = fl["gene_end"]; a["gene_strand"] = fl["gene_strand"]
    a["gene_exonic_length"] = fl["gene_exonic_length"]; a["gene_flags"] = (fl["gene_is_dummy"] | fl["gene_is_protein_coding"] << 1).astype(np.uint8)
    ne = len(fl["exon_gene"]); a["n_exons"] = ne
    a["exon_gene"] = fl["exon_gene"]; a["exon_start"] = fl["exon_start"]; a["exon_end"] = fl["exon_end"]
    a["exon_cds_start"] = fl["exon_cds_start"]; a["exon_cds_end"] = fl["exon_cds_end"]
    nxt = fl["exon_next"]; a["exon_next_start"] = np.where(nxt >= 0, fl["exon_start"][np.maximum(nxt, 0)], -1).astype(np.int32)
    a["exon_flags"] = ((fl["exon_prev"] >= 0).astype(np.uint8) | (nxt >= 0).astype(np.uint8) << 1)
    rc = fl["region_contig"].astype(np.int64)
    a["exon_region_begin"] = np.searchsorted(rc, np.arange(n_contigs + 1)).astype(np.uint32)
    a["exon_region_end"] = fl["region_end"]; a["exon_region_off"] = fl["region_off"]
    a["exon_region_items"] = fl["region_exons"] if len(fl["region_exons"]) else np.zeros(1, np.uint32)
    # the gene index is not needed by the device stages under test
    a["gene_region_begin"] = np.zeros(n_contigs + 1, np.uint32); a["gene_region_end"] = np.zeros(1, np.int32)
    a["gene_region_off"] = np.zeros(1, np.uint32); a["gene_region_items"] = np.zeros(1, np.uint32)
    return a


def context_from_oracle(world, lib_path, device=0):
    """Context loaded with the oracle's annotated fragments, annotation and the world's contigs."""
    from arriba_b200 import lib
    ctx = lib.Context(device, lib_path)
    names, flags, seqs = contigs_from_world(world)
    fl = world.stage("fragment_length")
    n_contigs = max(len(names), int(fl["gene_contig"].max()) + 1)
    assert n_contigs == len(names)
    ctx.set_contigs(flags, seqs)
    ctx.set_annotation(annotation_from_dump(fl, n_contigs))
    ctx.push_chunk(chunk_from_dump(world.stage("annotated")))
    return ctx
