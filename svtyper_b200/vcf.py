"""VCF header / record model for the drop-in entry points.

Host-side text plumbing only (nothing here is accelerated); it exists because the output
text is part of the parity contract: `sv_genotype` / `sso_genotype` must emit a VCF that is
`diff -I '^##fileDate='`-clean against the reference's (reference tests/test_svtyper.py:66-89).
Behaviour follows the reference's `Vcf` / `Variant` / `Genotype` classes
(svtyper/parsers.py:18-399), including the quirks that shape the text:

  * header re-emitted by category: fileformat, fileDate (today), reference, INFO, ALT,
    FORMAT, everything else, column line (parsers.py:75-95);
  * FORMAT `GT` is always defined first, with the description "Genotype" (parsers.py:27);
  * structured header fields are read positionally as ID, Number, Type, Description and
    each value is what follows the FIRST '=' up to the next '=' (parsers.py:53-66);
  * INFO keys that the header does not declare are dropped on output (parsers.py:310-318);
  * FORMAT keys are emitted in header order; floats print as %0.2f; QUAL as %0.2f
    (parsers.py:320-346, :386-398).
"""
from __future__ import annotations

import re
import sys
import time

_FIELD_RE = re.compile(r'(?:[^,"]|"[^"]*")+')

SVTYPER_FORMATS = (
    ("GQ", 1, "Integer", "Genotype quality"),
    ("SQ", 1, "Float", "Phred-scaled probability that this site is variant (non-reference in this sample"),
    ("GL", "G", "Float", "Genotype Likelihood, log10-scaled likelihoods of the data given the called genotype "
                         "for each possible genotype generated from the reference and alternate alleles given "
                         "the sample ploidy"),
    ("DP", 1, "Integer", "Read depth"),
    ("RO", 1, "Integer", "Reference allele observation count, with partial observations recorded fractionally"),
    ("AO", "A", "Integer", "Alternate allele observations, with partial observations recorded fractionally"),
    ("QR", 1, "Integer", "Sum of quality of reference observations"),
    ("QA", "A", "Integer", "Sum of quality of alternate observations"),
    ("RS", 1, "Integer", "Reference allele split-read observation count, with partial observations recorded "
                         "fractionally"),
    ("AS", "A", "Integer", "Alternate allele split-read observation count, with partial observations recorded "
                           "fractionally"),
    ("ASC", "A", "Integer", "Alternate allele clipped-read observation count, with partial observations "
                            "recorded fractionally"),
    ("RP", 1, "Integer", "Reference allele paired-end observation count, with partial observations recorded "
                         "fractionally"),
    ("AP", "A", "Integer", "Alternate allele paired-end observation count, with partial observations recorded "
                           "fractionally"),
    ("AB", "A", "Float", "Allele balance, fraction of observations from alternate allele, QA/(QR+QA)"),
)


def _unquote(text):
    text = str(text)
    if text.startswith('"') and text.endswith('"'):
        return text[1:-1]
    return text


class MetaLine(object):
    """One structured ##INFO / ##FORMAT / ##ALT definition."""
    __slots__ = ("kind", "id", "number", "type", "desc")

    def __init__(self, kind, id_, number, type_, desc):
        self.kind, self.id = kind, str(id_)
        self.number = None if number is None else str(number)
        self.type = None if type_ is None else str(type_)
        self.desc = _unquote(desc)

    def render(self):
        if self.kind == "ALT":
            return '##ALT=<ID=%s,Description="%s">' % (self.id, self.desc)
        return '##%s=<ID=%s,Number=%s,Type=%s,Description="%s">' % (
            self.kind, self.id, self.number, self.type, self.desc)


class VcfHeader(object):
    def __init__(self):
        self.file_format = "VCFv4.2"
        self.reference = ""
        self.infos, self.alts, self.formats = [], [], []
        self.misc = []
        self.samples = []
        self.filename = None
        self.define("FORMAT", "GT", 1, "String", "Genotype")

    # ---- definitions -------------------------------------------------------------------
    def _table(self, kind):
        return {"INFO": self.infos, "ALT": self.alts, "FORMAT": self.formats}[kind]

    def define(self, kind, id_, number=None, type_=None, desc=""):
        table = self._table(kind)
        if any(m.id == str(id_) for m in table):
            return                      # first definition wins
        table.append(MetaLine(kind, id_, number, type_, desc))
        self._format_ids = None

    def format_ids(self):
        ids = getattr(self, "_format_ids", None)
        if ids is None or len(ids) != len(self.formats):
            ids = self._format_ids = [m.id for m in self.formats]
            self._format_rank = {k: i for i, k in enumerate(ids)}
        return ids

    def format_rank(self):
        """{FORMAT id: position in the header}: the order FORMAT keys are written in."""
        self.format_ids()
        return self._format_rank

    def info_ids(self):
        return [m.id for m in self.infos]

    def ensure_svtyper_fields(self):
        self.define("INFO", "SVTYPE", 1, "String", "Type of structural variant")
        for fid, num, typ, desc in SVTYPER_FORMATS:
            self.define("FORMAT", fid, num, typ, desc)

    # ---- parsing -----------------------------------------------------------------------
    def parse(self, lines):
        for line in lines:
            key = line.split("=")[0]
            if key == "##fileformat":
                self.file_format = line.rstrip().split("=")[1]
            elif key == "##reference":
                self.reference = line.rstrip().split("=")[1]
            elif key in ("##INFO", "##FORMAT", "##ALT"):
                body = line[line.find("<") + 1:line.rfind(">")]
                vals = [piece.split("=")[1] for piece in _FIELD_RE.findall(body)]
                if key == "##ALT":
                    self.define("ALT", vals[0], None, None, vals[1])
                else:
                    self.define(key[2:], vals[0], vals[1], vals[2], vals[3])
            elif line[0] == "#" and line[1] != "#":
                self.samples = line.rstrip().split("\t")[9:]
            elif line.startswith("##fileDate="):
                pass
            else:
                self.misc.append(line.rstrip())
        return self

    def add_sample(self, name):
        self.samples.append(name)

    def sample_column(self, name):
        return self.samples.index(name) + 9

    # ---- output ------------------------------------------------------------------------
    def render(self):
        lines = ["##fileformat=" + self.file_format,
                 "##fileDate=" + time.strftime("%Y%m%d"),
                 "##reference=" + self.reference]
        lines += [m.render() for m in self.infos]
        lines += [m.render() for m in self.alts]
        lines += [m.render() for m in self.formats]
        lines += self.misc
        lines.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"]
                               + self.samples))
        return "\n".join(lines)


class SampleCall(object):
    """FORMAT values of one sample at one record; shares the record's active-key list."""
    __slots__ = ("record", "values")

    def __init__(self, record, gt):
        self.record = record
        self.values = {}
        self.set("GT", gt)

    def set(self, key, value):
        rank = self.record.header.format_rank()
        if key not in rank:
            sys.stderr.write('Error: invalid FORMAT field, "' + key + '"\n')
            sys.exit(1)
        self.values[key] = value
        active = self.record.active_formats
        if key not in active:
            active.append(key)
            active.sort(key=rank.__getitem__)

    def set_many(self, items):
        """set() for a list of (key, value) pairs with one ordering pass (the scored row of a site)."""
        rank = self.record.header.format_rank()
        active = self.record.active_formats
        seen = set(active)
        grew = False
        for key, value in items:
            if key not in rank:
                sys.stderr.write('Error: invalid FORMAT field, "' + key + '"\n')
                sys.exit(1)
            self.values[key] = value
            if key not in seen:
                seen.add(key)
                active.append(key)
                grew = True
        if grew:
            active.sort(key=rank.__getitem__)

    def get(self, key):
        return self.values[key]

    def render(self):
        out = []
        for key in self.record.active_formats:
            if key in self.values:
                v = self.values[key]
                out.append("%0.2f" % v if type(v) == float else v)
            else:
                out.append(".")
        return ":".join(map(str, out))


class VcfRecord(object):
    def __init__(self, cols, header):
        self.header = header
        self.chrom = cols[0]
        self.pos = int(cols[1])
        self.var_id = cols[2]
        self.ref = cols[3]
        self.alt = cols[4]
        self.qual = 0 if cols[5] == "." else float(cols[5])
        self.filter = cols[6]
        self.active_formats = []
        self.calls = {}
        if len(cols) < 8:
            sys.stderr.write("Error: VCF file must have at least 8 columns\n")
            sys.exit(1)
        if len(cols) < 9:
            cols.append("GT")
        for name in header.samples:
            try:
                col = cols[header.sample_column(name)]
                call = SampleCall(self, col.split(":")[0])
                self.calls[name] = call
                for key, value in zip(cols[8].split(":"), col.split(":")):
                    call.set(key, value)
            except IndexError:
                self.calls[name] = SampleCall(self, "./.")
        self.info = {}
        for item in cols[7].split(";"):
            kv = item.split("=")
            self.info[kv[0]] = True if len(kv) == 1 else kv[1]

    # ---- accessors -----------------------------------------------------------------------
    def has_svtype(self):
        return "SVTYPE" in self.info

    def svtype(self):
        return self.info["SVTYPE"]

    def call(self, sample):
        if sample in self.header.samples:
            return self.calls[sample]
        sys.stderr.write('Error: invalid sample name, "' + sample + '"\n')
        return None

    def adopt_calls(self, other):
        """BND mates share one genotype: the second record prints the first one's calls
        (reference classic.py:517-521, singlesample.py:648-652)."""
        self.qual = other.qual
        self.active_formats = other.active_formats
        self.calls = other.calls

    # ---- output --------------------------------------------------------------------------
    def info_string(self):
        parts = []
        for meta in self.header.infos:
            if meta.id in self.info:
                parts.append(meta.id if meta.type == "Flag" else "%s=%s" % (meta.id, self.info[meta.id]))
        return ";".join(parts)

    def format_string(self):
        active = set(self.active_formats)
        return ":".join(f for f in self.header.format_ids() if f in active)

    def render(self):
        calls = "\t".join(self.calls[s].render() for s in self.header.samples)
        return "\t".join(map(str, [self.chrom, self.pos, self.var_id, self.ref, self.alt, "%0.2f" % self.qual,
                                   self.filter, self.info_string(), self.format_string(), calls]))


def confidence_interval(record, tag, alt_tag, max_ci_dist):
    """CIPOS/CIEND, or the 95% interval when the full one is wider than max_ci_dist
    (reference parsers.py:11-15)."""
    ci = [int(x) for x in record.info[tag].split(",")]
    if ci[1] - ci[0] > max_ci_dist:
        return [int(x) for x in record.info[alt_tag].split(",")]
    return ci


def _strand_is_reverse(alt):
    return not (alt[-1] == "[" or alt[-1] == "]")


def simple_breakpoint(rec, max_ci_dist):
    """Breakpoint dict of a DEL/DUP/INV record (reference parsers.py:171-209)."""
    svtype = rec.svtype()
    posA, posB = rec.pos, int(rec.info["END"])
    bp = {"id": rec.var_id, "svtype": svtype}
    if svtype == "DEL":
        bp["var_length"] = posB - posA
        rev = (False, True)
    elif svtype == "DUP":
        rev = (True, False)
    else:
        rev = (False, False)
    bp["A"] = {"chrom": rec.chrom, "pos": posA + int(rev[0]),
               "ci": confidence_interval(rec, "CIPOS", "CIPOS95", max_ci_dist), "is_reverse": rev[0]}
    bp["B"] = {"chrom": rec.chrom, "pos": posB + int(rev[1]),
               "ci": confidence_interval(rec, "CIEND", "CIEND95", max_ci_dist), "is_reverse": rev[1]}
    return bp


def bnd_breakpoint(first, second, max_ci_dist):
    """Breakpoint dict of a BND mate pair (reference parsers.py:125-169, classic.py:233-258)."""
    revA, revB = _strand_is_reverse(first.alt), _strand_is_reverse(second.alt)
    return {
        "id": first.var_id, "svtype": "BND",
        "A": {"chrom": first.chrom, "pos": first.pos + int(revA),
              "ci": confidence_interval(first, "CIPOS", "CIPOS95", max_ci_dist), "is_reverse": revA},
        "B": {"chrom": second.chrom, "pos": second.pos + int(revB),
              "ci": confidence_interval(second, "CIPOS", "CIPOS95", max_ci_dist), "is_reverse": revB},
    }


def split_header_and_body(lines):
    """(header lines incl. the #CHROM line, record lines) of a VCF text stream."""
    header, body = [], []
    for line in lines:
        if line.startswith("#"):
            header.append(line)
        elif line.strip():
            body.append(line)
    return header, body
