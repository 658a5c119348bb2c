"""Test infrastructure: a spec-level BAM reader / writer in plain Python (SAMv1 sections 4.1, 4.2, 4.2.4), independent of tools/synth.cpp's writer,
of the oracle's htslib shim and of the product's reader. `transcode` rewrites a BAM into one that must mean the same to the reference
(read_chimeric_alignments.cpp:611-756 looks at the core fields, the CIGAR, the sequence and the tags HI and SA only) while changing everything a reader
can trip over: auxiliary fields of every type in every order (integer tags in every width that holds the value, strings whose bytes look like other tags,
B arrays, empty arrays), qualities, bin / mapq, the header text, and the BGZF framing (members of 1 byte to 64 KiB of payload, stored and deflated at
every level, empty members in the middle, records and even their 4-byte length words split over members)."""
import struct, zlib, random

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_members(raw):
    """yields the inflated payload of every BGZF member, checking the framing as the specification words it"""
    at = 0
    while at < len(raw):
        if len(raw) - at < 18:
            raise ValueError("truncated member header")
        id1, id2, cm, flg, _mtime, _xfl, _os, xlen = struct.unpack_from("<BBBBIBBH", raw, at)
        if (id1, id2, cm) != (31, 139, 8) or not flg & 4:
            raise ValueError("not a BGZF member at %d" % at)
        extra = raw[at + 12: at + 12 + xlen]; bsize = None; e = 0
        while e + 4 <= len(extra):
            si1, si2, slen = struct.unpack_from("<BBH", extra, e)
            if (si1, si2) == (66, 67) and slen == 2:
                bsize = struct.unpack_from("<H", extra, e + 4)[0]
            e += 4 + slen
        if bsize is None:
            raise ValueError("no BC subfield")
        total = bsize + 1
        cdata = raw[at + 12 + xlen: at + total - 8]
        crc, isize = struct.unpack_from("<II", raw, at + total - 8)
        data = zlib.decompress(cdata, -15) if cdata else b""
        if len(data) != isize or zlib.crc32(data) != crc:
            raise ValueError("CRC / ISIZE mismatch at %d" % at)
        yield data
        at += total


def read_bam(path):
    """-> (header text, [(contig name, length)], [record bodies as bytes])"""
    data = b"".join(bgzf_members(open(path, "rb").read()))
    if data[:4] != b"BAM\x01":
        raise ValueError("magic")
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8: 8 + l_text]; at = 8 + l_text
    n_ref = struct.unpack_from("<i", data, at)[0]; at += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, at)[0]; at += 4
        name = data[at: at + l_name - 1].decode(); at += l_name
        refs.append((name, struct.unpack_from("<i", data, at)[0])); at += 4
    records = []
    while at < len(data):
        size = struct.unpack_from("<i", data, at)[0]; at += 4
        records.append(data[at: at + size]); at += size
    if at != len(data):
        raise ValueError("trailing bytes")
    return text, refs, records


def split_record(body):
    """record body -> dict of its parts (aux as a list of (tag, type, value bytes as stored))"""
    tid, pos, l_qname, mapq, bin_, n_cigar, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", body, 0)
    at = 32
    qname = body[at: at + l_qname]; at += l_qname
    cigar = body[at: at + 4 * n_cigar]; at += 4 * n_cigar
    seq = body[at: at + (l_seq + 1) // 2]; at += (l_seq + 1) // 2
    qual = body[at: at + l_seq]; at += l_seq
    aux = []
    while at < len(body):
        tag = body[at: at + 2]; t = body[at + 2: at + 3]; at += 3
        if t in b"AcC":
            n = 1
        elif t in b"sS":
            n = 2
        elif t in b"iIf":
            n = 4
        elif t in b"ZH":
            n = body.index(b"\0", at) - at + 1
        elif t == b"B":
            sub = body[at: at + 1]; cnt = struct.unpack_from("<I", body, at + 1)[0]
            n = 5 + cnt * {b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}[sub]
        else:
            raise ValueError("aux type %r" % t)
        aux.append((tag, t, body[at: at + n])); at += n
    return dict(tid=tid, pos=pos, mapq=mapq, bin=bin_, flag=flag, l_seq=l_seq, mtid=mtid, mpos=mpos, tlen=tlen, qname=qname, cigar=cigar, seq=seq, qual=qual, aux=aux)


def join_record(r):
    head = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(r["qname"]), r["mapq"], r["bin"], len(r["cigar"]) // 4, r["flag"], r["l_seq"], r["mtid"], r["mpos"], r["tlen"])
    return head + r["qname"] + r["cigar"] + r["seq"] + r["qual"] + b"".join(tag + t + v for tag, t, v in r["aux"])


def aux_int_value(t, v):
    return struct.unpack("<" + {b"c": "b", b"C": "B", b"s": "h", b"S": "H", b"i": "i", b"I": "I"}[t], v)[0]


def int_encodings(x):
    """every (type, bytes) of the specification that holds the integer x"""
    out = []
    for t, fmt, lo, hi in ((b"c", "b", -128, 127), (b"C", "B", 0, 255), (b"s", "h", -32768, 32767), (b"S", "H", 0, 65535), (b"i", "i", -2 ** 31, 2 ** 31 - 1), (b"I", "I", 0, 2 ** 32 - 1)):
        if lo <= x <= hi:
            out.append((t, struct.pack("<" + fmt, x)))
    return out


def noise_tags(rng):
    """auxiliary fields the reference never asks for, of every type; some are built to mislead a reader that searches for the bytes "HI" or "SA" """
    pool = [
        (b"XA", b"A", bytes([rng.choice(b"!~HIS")])),
        (b"Xc", b"c", struct.pack("<b", rng.randint(-128, 127))),
        (b"XC", b"C", struct.pack("<B", rng.randint(0, 255))),
        (b"Xs", b"s", struct.pack("<h", rng.randint(-32768, 32767))),
        (b"XS", b"S", struct.pack("<H", rng.randint(0, 65535))),
        (b"Xi", b"i", struct.pack("<i", rng.randint(-2 ** 31, 2 ** 31 - 1))),
        (b"XI", b"I", struct.pack("<I", rng.randint(0, 2 ** 32 - 1))),
        (b"Xf", b"f", struct.pack("<f", rng.random())),
        (b"XZ", b"Z", rng.choice([b"", b"HIC\x07", b"SAZ1,5,+,3S7M,255,0;", b"HIi\x01\x01\x01\x01", b"x" * rng.randint(1, 300)]) + b"\0"),
        (b"XH", b"H", rng.choice([b"", b"1AE301", b"4849"]) + b"\0"),
        (b"IH", b"C", b"\x09"), (b"HJ", b"C", b"\x09"), (b"Sa", b"Z", b"9,9,-,5M,1,1;\0"), (b"AS", b"i", struct.pack("<i", rng.randint(0, 300))),
        (b"MD", b"Z", b"50A50\0"), (b"NM", b"C", b"\x01"), (b"jM", b"B", b"c" + struct.pack("<I", 2) + b"\x01\xff"), (b"jI", b"B", b"i" + struct.pack("<Iii", 2, 100, 200)),
    ]
    for sub, width in ((b"c", 1), (b"C", 1), (b"s", 2), (b"S", 2), (b"i", 4), (b"I", 4), (b"f", 4)):
        cnt = rng.choice([0, 1, 2, 7, 40])
        body = bytes(rng.getrandbits(8) for _ in range(cnt * width))
        if sub == b"C" and cnt >= 2:
            body = b"HIC\x05SAZ"[: cnt] + body[len(b"HIC\x05SAZ"[: cnt]):]   # the bytes of an array may spell a tag
        pool.append((b"B" + sub.upper() if sub.islower() else b"b" + sub, b"B", sub + struct.pack("<I", cnt) + body))
    rng.shuffle(pool)
    return pool[: rng.randint(0, len(pool))]


def transcode_record(body, rng):
    r = split_record(body)
    kept = []
    for tag, t, v in r["aux"]:
        if tag == b"HI":                       # bam_aux2i reads every integer type (read_chimeric_alignments.cpp:619-621)
            t, v = rng.choice(int_encodings(aux_int_value(t, v)))
        if tag == b"NH":                       # never read by the reference: other widths, or gone
            if rng.random() < 0.2:
                continue
            t, v = rng.choice(int_encodings(aux_int_value(t, v)))
        kept.append((tag, t, v))
    aux = kept + noise_tags(rng)
    rng.shuffle(aux)
    r["aux"] = aux
    r["qual"] = bytes([0xff]) * r["l_seq"] if rng.random() < 0.1 else bytes(rng.randint(0, 41) for _ in range(r["l_seq"]))
    r["mapq"] = rng.choice([0, 1, 3, 255]); r["bin"] = rng.randint(0, 37449); r["tlen"] = rng.randint(-10 ** 6, 10 ** 6)
    return join_record(r)


def bgzf_member(payload, level):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = c.compress(payload) + c.flush()
    if len(cdata) + 26 > 65536:
        raise ValueError("member too large")
    return (struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, len(cdata) + 25) + cdata + struct.pack("<II", zlib.crc32(payload), len(payload)))


def write_bam(path, text, refs, bodies, rng, eof=True):
    data = [b"BAM\x01", struct.pack("<i", len(text)), text, struct.pack("<i", len(refs))]
    for name, length in refs:
        data += [struct.pack("<i", len(name) + 1), name.encode() + b"\0", struct.pack("<i", length)]
    for b in bodies:
        data += [struct.pack("<i", len(b)), b]
    data = b"".join(data)
    with open(path, "wb") as f:
        at = 0
        while at < len(data):
            mode = rng.random()
            n = rng.randint(1, 40) if mode < 0.05 else rng.randint(1, 3000) if mode < 0.3 else rng.randint(20000, 65280)
            level = rng.choice([0, 0, 1, 6, 9])
            if level and n > 60000:
                n = 60000   # deflate of incompressible bytes may grow a little: stay inside the 64 KiB member
            f.write(bgzf_member(data[at: at + n], level)); at += n
            if rng.random() < 0.03:
                f.write(bgzf_member(b"", rng.choice([0, 6])))   # an empty member in the middle is legal (and is what an EOF marker of a concatenated file is)
        if eof:
            f.write(EOF_BLOCK)


def transcode(src, dst, seed, eof=True):
    rng = random.Random(seed)
    text, refs, bodies = read_bam(src)
    text = text + b"@CO\ttranscoded by tests/bamtools.py\n@CO\t" + b"HI:i:1 SA:Z:x " * rng.randint(0, 50) + b"\n"
    write_bam(dst, text, refs, [transcode_record(b, rng) for b in bodies], rng, eof)
    return len(bodies)


# ---- record-level mutations: files the aligner would not write, which the reference nevertheless has defined behaviour for (flag tests, the malformed-
# fragment rules of read_chimeric_alignments.cpp:451-495, supplementary clips at the wrong end, missing HI, orphans). The oracle says what comes out.
C_M, C_S, C_H, C_EQ, C_X = 0, 4, 5, 7, 8


def _cigar_ops(r):
    return list(struct.unpack("<%dI" % (len(r["cigar"]) // 4), r["cigar"]))


def _set_cigar(r, ops):
    r["cigar"] = struct.pack("<%dI" % len(ops), *ops)


def _replace_op(r, which, src, dst):
    """turns the first / last CIGAR operation from `src` into `dst` (same length); False if it is not a `src`"""
    ops = _cigar_ops(r)
    if not ops or (ops[which] & 15) != src:
        return False
    ops[which] = (ops[which] >> 4) << 4 | dst
    _set_cigar(r, ops)
    return True


MUTATIONS = ("unmapped", "mate_unmapped", "clear_proper", "set_proper", "drop_record", "double_supplementary", "clip_to_other_end", "hard_clip_supplementary",
             "hard_clip_primary", "single_end", "secondary_without_hi", "other_hi", "match_ops", "supplementary_elsewhere", "flip_strand", "swap_mate_flags", "third_primary",
             "drop_sa", "negative_hi")


def mutate_fragment(recs, kind, rng, refs):
    """recs: list of split records of one read name (modified in place); returns the new list"""
    pick = rng.randrange(len(recs)); r = recs[pick]
    supp = [x for x in recs if x["flag"] & 0x800]; prim = [x for x in recs if not x["flag"] & 0x800]
    if kind == "unmapped":
        r["flag"] |= 0x4
    elif kind == "mate_unmapped":
        r["flag"] |= 0x8
    elif kind == "clear_proper":
        for x in recs: x["flag"] &= ~0x2
    elif kind == "set_proper":
        for x in recs: x["flag"] |= 0x2
    elif kind == "drop_record":
        del recs[pick]
    elif kind == "double_supplementary" and supp:
        recs.append(dict(supp[0]))
    elif kind == "clip_to_other_end" and supp:
        ops = _cigar_ops(supp[0]); ops.reverse(); _set_cigar(supp[0], ops)
    elif kind == "hard_clip_supplementary" and supp:
        _replace_op(supp[0], 0, C_S, C_H); _replace_op(supp[0], -1, C_S, C_H)   # legal (STAR's HardClip); the sequence is kept, which the reference never looks at
    elif kind == "hard_clip_primary" and prim:
        x = rng.choice(prim)
        if not _replace_op(x, 0, C_S, C_H): _replace_op(x, -1, C_S, C_H)
    elif kind == "single_end":
        for x in recs: x["flag"] &= ~(0x1 | 0x2 | 0x8 | 0x20 | 0x40 | 0x80)
    elif kind == "secondary_without_hi":
        r["flag"] |= 0x100; r["aux"] = [a for a in r["aux"] if a[0] != b"HI"]
    elif kind == "other_hi":
        r["aux"] = [a for a in r["aux"] if a[0] != b"HI"] + [(b"HI", b"C", bytes([rng.randint(2, 9)]))]
    elif kind == "negative_hi":
        for x in recs: x["aux"] = [a for a in x["aux"] if a[0] != b"HI"] + [(b"HI", b"i", struct.pack("<i", -3))]
    elif kind == "match_ops":
        ops = _cigar_ops(r); _set_cigar(r, [(o >> 4) << 4 | rng.choice([C_EQ, C_X]) if (o & 15) == C_M else o for o in ops])
    elif kind == "supplementary_elsewhere" and supp:
        supp[0]["tid"] = rng.randrange(len(refs)); supp[0]["pos"] = rng.randint(1000, max(1000, refs[supp[0]["tid"]][1] - 2000))   # inside the contig: past its end the reference reads beyond its coverage vector (undefined)
    elif kind == "flip_strand":
        r["flag"] ^= 0x10
    elif kind == "swap_mate_flags":
        for x in recs:
            if x["flag"] & 0xC0 in (0x40, 0x80): x["flag"] ^= 0xC0
    elif kind == "third_primary" and prim:
        recs.append(dict(rng.choice(prim)))
    elif kind == "drop_sa":
        for x in recs: x["aux"] = [a for a in x["aux"] if a[0] != b"SA"]
    return recs


def mutate(src, dst, seed, rate=0.15, kinds=MUTATIONS):
    """rewrites `src` with one mutation applied to a fraction `rate` of its read names; returns {kind: how many}"""
    rng = random.Random(seed)
    text, refs, bodies = read_bam(src)
    recs = [split_record(b) for b in bodies]
    by_name = {}
    for i, r in enumerate(recs):
        by_name.setdefault(r["qname"], []).append(i)
    replaced = {}; tally = {}
    for name, idx in by_name.items():
        if rng.random() >= rate:
            continue
        kind = rng.choice(kinds)
        replaced[idx[0]] = mutate_fragment([recs[i] for i in idx], kind, rng, refs)
        for i in idx[1:]:
            replaced[i] = []
        tally[kind] = tally.get(kind, 0) + 1
    out = []
    for i, r in enumerate(recs):
        out += [join_record(x) for x in replaced[i]] if i in replaced else [bodies[i]]
    write_bam(dst, text, refs, out, rng)
    return tally
