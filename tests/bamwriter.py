"""Test helper: write a small coordinate-sorted BAM and its BAI from record tuples (stdlib only), so
the two BAM readers of this repo (svtyper_b200/bamio.py and libsvgt_pack.so) can be compared on
constructed edge cases.  Not a general writer: one record never spans BGZF blocks' 64 KB limit."""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"


def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    payload = comp.compress(data) + comp.flush()
    bsize = len(payload) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize)
            + payload + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def _parse_cigar(text):
    out, num = [], ""
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((CIGAR_OPS.index(ch), int(num)))
            num = ""
    return out


def encode_record(tid, pos, qname, flag, mapq, cigar, l_seq, tags, tlen=0):
    """tags: list of (key, 'Z'|'i'|'A', value)."""
    cig = _parse_cigar(cigar) if isinstance(cigar, str) else list(cigar)
    end = pos + sum(n for op, n in cig if op in (0, 2, 3, 7, 8))
    name = qname.encode("ascii") + b"\x00"
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, _reg2bin(pos, max(end, pos + 1)), len(cig), flag,
                       l_seq, -1, -1, tlen)
    body += name + b"".join(struct.pack("<I", (n << 4) | op) for op, n in cig)
    body += b"\x00" * ((l_seq + 1) // 2) + b"\xff" * l_seq
    for key, typ, val in tags:
        body += key.encode("ascii") + typ.encode("ascii")
        if typ == "Z":
            body += val.encode("ascii") + b"\x00"
        elif typ == "i":
            body += struct.pack("<i", val)
        elif typ == "A":
            body += val.encode("ascii")
    return struct.pack("<i", len(body)) + body, end


def write_bam(path, references, header_text, records, block_bytes=3000):
    """references: [(name, length)]; records: dicts with tid,pos,qname,flag,mapq,cigar,l_seq,tags
    (must already be coordinate-sorted).  Writes `path` and `path + '.bai'`."""
    head = b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode("ascii")
    head += struct.pack("<i", len(references))
    for name, length in references:
        head += struct.pack("<i", len(name) + 1) + name.encode("ascii") + b"\x00" + struct.pack("<i", length)
    blocks = [_bgzf_block(head)]
    coff = len(blocks[0])
    # small blocks on purpose: many block boundaries inside and between records' chunks
    cur, cur_recs, placed = b"", [], []
    def flush():
        nonlocal cur, cur_recs, coff
        if not cur:
            return
        blk = _bgzf_block(cur)
        u = 0
        for rec, enc, end in cur_recs:
            placed.append((rec, (coff << 16) | u, len(enc), end, coff, len(blk), len(cur)))
            u += len(enc)
        blocks.append(blk)
        coff += len(blk)
        cur, cur_recs = b"", []
    for rec in records:
        enc, end = encode_record(rec["tid"], rec["pos"], rec["qname"], rec["flag"], rec["mapq"], rec["cigar"],
                                 rec.get("l_seq", 0), rec.get("tags", []), rec.get("tlen", 0))
        if len(cur) + len(enc) > block_bytes:
            flush()
        cur += enc
        cur_recs.append((rec, enc, end))
    flush()
    blocks.append(_bgzf_block(b""))
    with open(path, "wb") as f:
        f.write(b"".join(blocks))
    # ---- BAI: bins + 16 kb linear index ----
    n_ref = len(references)
    bins = [dict() for _ in range(n_ref)]
    linear = [dict() for _ in range(n_ref)]
    for rec, voff, size, end, coff_b, blk_len, ulen in placed:
        tid = rec["tid"]
        if tid < 0:
            continue
        u_end = (voff & 0xFFFF) + size
        vend = ((coff_b + blk_len) << 16) if u_end >= ulen else ((coff_b << 16) | u_end)
        e = max(end, rec["pos"] + 1)
        b = _reg2bin(rec["pos"], e)
        chunks = bins[tid].setdefault(b, [])
        if chunks and chunks[-1][1] == voff:
            chunks[-1][1] = vend
        else:
            chunks.append([voff, vend])
        for w in range(rec["pos"] >> 14, ((e - 1) >> 14) + 1):
            if w not in linear[tid] or voff < linear[tid][w]:
                linear[tid][w] = voff
    out = b"BAI\x01" + struct.pack("<i", n_ref)
    for tid in range(n_ref):
        out += struct.pack("<i", len(bins[tid]))
        for b, chunks in sorted(bins[tid].items()):
            out += struct.pack("<Ii", b, len(chunks))
            for cb, ce in chunks:
                out += struct.pack("<QQ", cb, ce)
        n_intv = (max(linear[tid]) + 1) if linear[tid] else 0
        out += struct.pack("<i", n_intv)
        last = 0
        for w in range(n_intv):
            last = linear[tid].get(w, last)
            out += struct.pack("<Q", last)
    with open(path + ".bai", "wb") as f:
        f.write(out)
