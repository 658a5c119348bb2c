"""Test infrastructure: damage to a GTF file that the reference's reader has defined behaviour for (annotation.cpp:113-377: comment lines, lines it cannot
parse, missing attributes, attribute order, gene ids that appear twice (another contig, or too far apart), genes beyond the end of their contig, CDS records
of unknown transcripts, features it does not read, lines in any order). The oracle says what comes out; the product's gene / exon tables and both output
files must equal it."""
import random

KINDS = ("comment", "blank", "short_line", "bad_number", "no_gene_name", "no_gene_id", "no_transcript_id", "attribute_order", "extra_attributes", "other_feature",
         "gene_id_on_other_contig", "gene_id_far_away", "beyond_contig_end", "cds_unknown_transcript", "unquoted", "dot_strand", "unknown_contig", "duplicate_exon")


def damage_gtf(src, dst, seed, rate, contig_lengths, kinds=KINDS):
    """contig_lengths: {name: length} of the assembly. One kind of damage per chosen line; returns {kind: count}."""
    rng = random.Random(seed)
    lines = open(src).read().split("\n")
    out = []; tally = {}
    contigs = sorted(contig_lengths)
    for line in lines:
        f = line.split("\t")
        if len(f) < 9 or line.startswith("#") or rng.random() >= rate:
            out.append(line); continue
        kind = rng.choice(kinds); tally[kind] = tally.get(kind, 0) + 1
        attrs = [a.strip() for a in f[8].split(";") if a.strip()]
        g = list(f)
        if kind == "comment":
            out += ["#" + line, line]
        elif kind == "blank":
            out += ["", line]
        elif kind == "short_line":
            out += ["\t".join(f[:rng.randint(1, 6)]), line]
        elif kind == "bad_number":
            g[3] = "x" + g[3]; out += ["\t".join(g), line]
        elif kind in ("no_gene_name", "no_gene_id", "no_transcript_id"):
            key = kind[3:]
            g[8] = "; ".join(a for a in attrs if not a.startswith(key + " ")) + ";"; out.append("\t".join(g))
        elif kind == "attribute_order":
            rng.shuffle(attrs); g[8] = "; ".join(attrs) + ";"; out.append("\t".join(g))
        elif kind == "extra_attributes":
            extra = ['gene_type "protein_coding"', 'tag "basic"', 'note "gene_name in a note; with gene_id too"', 'level 2', 'transcript_name "T-201"']
            rng.shuffle(extra); attrs = attrs + extra[: rng.randint(1, len(extra))]
            if rng.random() < 0.5: rng.shuffle(attrs)
            g[8] = "; ".join(attrs) + ";"; out.append("\t".join(g))
        elif kind == "other_feature":
            g[2] = rng.choice(["UTR", "start_codon", "stop_codon", "transcript", "Selenocysteine", "five_prime_utr"]); out += [line, "\t".join(g)]
        elif kind == "gene_id_on_other_contig":
            g[0] = rng.choice([c for c in contigs if c != f[0]] or contigs); out += [line, "\t".join(g)]
        elif kind == "gene_id_far_away":
            shift = 2000000; length = contig_lengths.get(f[0], 0)
            if int(f[4]) + shift + 10 < length:
                g[3] = str(int(f[3]) + shift); g[4] = str(int(f[4]) + shift); out += [line, "\t".join(g)]
            else:
                out.append(line)
        elif kind == "beyond_contig_end":
            length = contig_lengths.get(f[0], 0)
            g[3] = str(length - 50); g[4] = str(length + rng.choice([0, 1, 500])); out += [line, "\t".join(g)]
        elif kind == "cds_unknown_transcript":
            g[2] = "CDS"; g[8] = g[8].replace('transcript_id "', 'transcript_id "NOSUCH'); g[7] = "0"
            out += [line] + (["\t".join(g)] if "transcript_id" in g[8] else [])
        elif kind == "unquoted":
            g[8] = g[8].replace('gene_name "', "gene_name ").replace('";', ";", 1); out += ["\t".join(g), line]
        elif kind == "dot_strand":
            g[6] = "."; out.append("\t".join(g))
        elif kind == "unknown_contig":
            g[0] = "GL000" + str(rng.randint(100, 999)) + ".1"; out += [line, "\t".join(g)]
        elif kind == "duplicate_exon":
            out += [line, line]
    if rng.random() < 0.5:   # the reader does not depend on the order of the lines
        head = [l for l in out if l.startswith("#")]; body = [l for l in out if not l.startswith("#")]
        rng.shuffle(body); out = head + body
    open(dst, "w").write("\n".join(out))
    return tally


def rewrite_fasta(names, seqs, dst, seed):
    """The assembly as another tool might have written it (assembly.cpp:29-60 reads the name up to the first blank, upper-cases, skips empty lines): descriptions
    after the names, ragged line lengths, empty lines, lower / mixed case, IUPAC codes, unplaced contigs before and after, no newline at the end."""
    rng = random.Random(seed)
    with open(dst, "wb") as f:
        f.write(b">GL000220.1 unplaced scaffold\n" + b"ACGT" * 500 + b"\n")
        for name, s in zip(names, seqs):
            b = bytearray(s)
            for _ in range(len(b) // 3000):
                b[rng.randrange(len(b))] = rng.choice(b"RYSWKMBDHV")
            lo = rng.randrange(len(b)); b[lo: lo + len(b) // 3] = bytes(b[lo: lo + len(b) // 3]).lower()
            f.write(b">" + name.encode() + rng.choice([b"", b" dna:chromosome REF", b"\tAC:CM000663.2  gi:568336023"]) + b"\n")
            at = 0
            while at < len(b):
                n = rng.randint(1, 200); f.write(bytes(b[at: at + n]) + b"\n"); at += n
                if rng.random() < 0.02:
                    f.write(b"\n")
        f.write(b">KI270728.1\n" + b"TTTTGGGG" * 300 + b"\n>chrUn_x\n\n>last\nACGTNNNN")
