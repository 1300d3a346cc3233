"""Per-sample library statistics: the producers of the kernel's insert-size tables.

Host-side, one-off per BAM.  Follows the reference's `Library` /
`Sample` (svtyper/parsers.py:406-719) and Counter statistics (svtyper/statistics.py:40-121):
either read back from the `-l` JSON or measured from the BAM, and written out in the
reference's JSON layout (svtyper/utils.py:25-51) so the two tools can share the cache file.
"""
from __future__ import annotations

import json
import os
import sys
from collections import OrderedDict

from . import evidence as ev

MIN_LIB_PREVALENCE = 1e-3      # reference classic.py:134, singlesample.py:68
OUTLIER_MADS = 10              # reference parsers.py:538


def open_alignment(path, ref_fasta=None):
    """pysam.AlignmentFile if pysam is installed, else this repo's BAM reader."""
    if not (path.endswith(".bam") or path.endswith(".cram")):
        sys.stderr.write("Error: %s is not a valid alignment file (*.bam or *.cram)\n" % path)
        sys.exit(1)
    try:
        import pysam
        if getattr(pysam, "__standin__", False):
            raise ImportError
        if path.endswith(".cram"):
            return pysam.AlignmentFile(path, mode="rc", reference_filename=ref_fasta)
        return pysam.AlignmentFile(path, mode="rb")
    except ImportError:
        from . import bamio
        if path.endswith(".cram"):
            sys.stderr.write("Error: CRAM input needs pysam; only BAM is readable without it\n")
            sys.exit(1)
        return bamio.AlignmentFile(path, "rb")


# ---- statistics over {value: count} tables (reference statistics.py:40-121) --------------
def _total(hist):
    return sum(hist.values())


def counter_median(hist):
    n = _total(hist)
    limit = 0.5 * n
    keys = sorted(hist)
    seen, i, v = 0, 0, keys[0]
    while seen < limit:
        v = keys[i]
        seen += hist[v]
        i += 1
    if seen == limit:
        return (v + keys[i]) / 2.0
    return v


def counter_upper_mad(hist, med):
    resid = {}
    for x, c in hist.items():
        if x > med:
            d = abs(x - med)
            resid[d] = resid.get(d, 0) + c
    return counter_median(resid)


def counter_mean(hist):
    s = 0.0
    for x in sorted(hist):
        s += x * float(hist[x])
    return s / _total(hist)


def counter_stdev(hist):
    u = counter_mean(hist)
    acc = 0.0
    for x in sorted(hist):           # CPython 2 iterates small-int dict keys in ascending order
        acc += hist[x] * (x - u) ** 2
    return (float(acc) / _total(hist)) ** 0.5


class LibraryInfo(object):
    def __init__(self, name, readgroups, read_length, hist, mean, sd, prevalence):
        self.name, self.readgroups = name, list(readgroups)
        self.read_length, self.hist = read_length, hist
        self.mean, self.sd, self.prevalence = mean, sd, prevalence

    @classmethod
    def from_json(cls, entry):
        hist = {int(k): int(v) for k, v in entry["histogram"].items()}
        return cls(entry["library_name"], entry["readgroups"], int(entry["read_length"]), hist,
                   float(entry["mean"]), float(entry["sd"]), float(entry["prevalence"]))

    @classmethod
    def from_bam(cls, bam, name, num_samp):
        groups = []
        for rg in bam.header["RG"]:
            in_lib = (rg["LB"] == name) if "LB" in rg else (name == "")
            if in_lib:
                groups.append(rg["ID"])
        lib = cls(name, groups, None, None, None, None, None)
        lib._measure_read_length(bam)
        lib._measure_inserts(bam, num_samp)
        lib._measure_prevalence(bam)
        return lib

    def _measure_read_length(self, bam, limit=10000):
        longest, seen = 0, 0
        for read in bam.fetch():
            if read.get_tag("RG") not in self.readgroups:
                continue
            n = read.infer_query_length()
            if n is not None and n > longest:
                longest = n
            if seen == limit:
                break
            seen += 1
        self.read_length = longest

    def _measure_inserts(self, bam, num_samp):
        hist, taken = {}, 0
        for read in bam.fetch():
            if (read.is_reverse or not read.mate_is_reverse or read.is_unmapped or read.mate_is_unmapped
                    or read.is_supplementary or read.is_secondary or read.template_length <= 0
                    or read.get_tag("RG") not in self.readgroups):
                continue
            hist[read.template_length] = hist.get(read.template_length, 0) + 1
            taken += 1
            if taken == num_samp:
                break
        self._finish_hist(hist, bam)

    def _finish_hist(self, hist, bam):
        """Outlier trimming + moments of the raw {template length: count} table (parsers.py:536-553)."""
        if not hist:
            sys.stderr.write("Error: failed to build insert size histogram for paired-end reads.\n"
                             "Please ensure BAM file (%s) has inward facing, paired-end reads.\n" % bam.filename)
            sys.exit(1)
        med = counter_median(hist)
        cut = med + OUTLIER_MADS * counter_upper_mad(hist, med)
        for x in [x for x in hist if x > cut]:
            del hist[x]
        self.hist = hist
        self.mean, self.sd = counter_mean(hist), counter_stdev(hist)

    def _measure_prevalence(self, bam, limit=100000):
        mine = seen = 0
        for read in bam.fetch():
            if seen == limit:
                break
            if read.get_tag("RG") in self.readgroups:
                mine += 1
            seen += 1
        self.prevalence = float(mine) / seen

    def to_json(self):
        return OrderedDict([("library_name", self.name), ("readgroups", self.readgroups),
                            ("read_length", self.read_length), ("mean", self.mean), ("sd", self.sd),
                            ("prevalence", self.prevalence),
                            ("histogram", {str(k): v for k, v in self.hist.items()})])


class SampleInfo(object):
    """One BAM: its sample name, libraries (in table order) and read-group map."""

    def __init__(self, name, bam, libraries, mapped, unmapped, shadowed=()):
        self.name, self.bam = name, bam
        self.libraries = libraries
        self.mapped, self.unmapped = mapped, unmapped
        self.rg_to_lib = {}
        for i, lib in enumerate(libraries):
            for rg in lib.readgroups:
                self.rg_to_lib[rg] = i
        # `shadowed`: entries of a -l JSON whose library name is listed again later.  The reference keys its
        # library dict by name (parsers.py:636), so the later entry replaces the earlier one there, while the
        # earlier one's read groups keep pointing at the replaced object (parsers.py:643-644): their reads are
        # still scored with the earlier entry's statistics, are active iff the name is (the active list holds
        # names: parsers.py:614-617, taken from the surviving entry), and the entry no longer counts for the
        # fetch flank (parsers.py:689).
        self.shadowed = set(shadowed)
        last = {lib.name: lib for i, lib in enumerate(libraries) if i not in self.shadowed}
        self.active = set(i for i, lib in enumerate(libraries)
                          if last.get(lib.name, lib).prevalence >= MIN_LIB_PREVALENCE)
        self._table = None

    @classmethod
    def from_lib_info(cls, bam, lib_info):
        name = bam.header["RG"][0]["SM"]
        try:
            entry = lib_info[name]
            libs = [LibraryInfo.from_json(e) for e in entry["libraryArray"]]
        except KeyError:
            sys.stderr.write("Error: sample %s not found in JSON library file.\n" % name)
            sys.exit(1)
        last = {}
        for i, lib in enumerate(libs):
            last[lib.name] = i
        shadowed = [i for i, lib in enumerate(libs) if last[lib.name] != i]
        return cls(name, bam, libs, entry["mapped"], entry["unmapped"], shadowed=shadowed)

    @classmethod
    def from_bam(cls, bam, num_samp, native=None):
        """Libraries measured from the BAM.  With an indexed .bam on disk the three passes per library
        (read length, insert sizes, prevalence) are one native scan (libsvgt_pack.so,
        svgt_bam_scan_libraries); `native=False` forces the Python passes (its parity checker)."""
        name = bam.header["RG"][0]["SM"]
        by_name = OrderedDict()
        for rg in bam.header["RG"]:
            lib_name = rg.get("LB", "")
            if lib_name not in by_name:
                by_name[lib_name] = None
        if native is None:
            from . import packer
            native = packer.usable_path(bam)
        if native:
            from . import packer
            libs = []
            for lib_name in by_name:
                groups = [rg["ID"] for rg in bam.header["RG"]
                          if ((rg["LB"] == lib_name) if "LB" in rg else (lib_name == ""))]
                libs.append(LibraryInfo(lib_name, groups, None, None, None, None, None))
            scans = packer.scan_libraries(bam, [lib.readgroups for lib in libs], num_samp)
            for lib, (read_length, mine, seen, hist) in zip(libs, scans):
                lib.read_length = read_length
                lib._finish_hist(hist, bam)
                lib.prevalence = float(mine) / seen
            return cls(name, bam, libs, bam.mapped, bam.unmapped)
        for lib_name in by_name:
            by_name[lib_name] = LibraryInfo.from_bam(bam, lib_name, num_samp)
        return cls(name, bam, list(by_name.values()), bam.mapped, bam.unmapped)

    @classmethod
    def open(cls, bam_path, lib_info_path, ref_fasta, num_samp):
        bam = open_alignment(bam_path, ref_fasta)
        if lib_info_path is not None and os.path.isfile(lib_info_path):
            with open(lib_info_path) as f:
                return cls.from_lib_info(bam, json.load(f))
        return cls.from_bam(bam, num_samp)

    def fetch_flank(self, z=3):
        return max(lib.mean + lib.sd * z for i, lib in enumerate(self.libraries) if i not in self.shadowed)

    def library_table(self):
        if self._table is None:
            self._table = ev.LibraryTable([(lib.mean, lib.sd, lib.hist) for lib in self.libraries])
        return self._table

    def to_json(self):
        return OrderedDict([("sample_name", self.name), ("bam", self.bam.filename),
                            ("libraryArray", [lib.to_json() for lib in self.libraries]),
                            ("mapped", self.mapped), ("unmapped", self.unmapped)])

    def close(self):
        nb = getattr(self, "native_bam", None)
        if nb is not None:
            nb.close()
            self.native_bam = None
        self.bam.close()


def write_sample_json(samples, path):
    """The `-l` cache file, in the reference's layout (svtyper/utils.py:25-51)."""
    with open(path, "w") as f:
        json.dump(OrderedDict((s.name, s.to_json()) for s in samples), f, indent=4)
