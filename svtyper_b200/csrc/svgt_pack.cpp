/*
 * svgt_pack.cpp -- native evidence packer (libsvgt_pack.so): BGZF + BAI region reader, read gathering,
 * split-candidate QC and row packing for a batch of breakpoints.  C ABI and reference citations:
 * include/svgt_pack.h.  Behaviour is pinned to svtyper_b200/gather.py + evidence.BatchPacker (themselves
 * pinned to the reference through the golden VCF) by tests/test_pack_native.py.
 *
 * Host code only (no CUDA): read gathering stays on the CPU by design.
 */
#include "../../include/svgt_pack.h"

#include <zlib.h>
#include <math.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

/* ---- schema constants (svtyper_b200/evidence.py) ---- */
enum {
    F_HAS_A = 1 << 0, F_HAS_B = 1 << 1, F_REV_A = 1 << 2, F_REV_B = 1 << 3, F_PAIRED = 1 << 4,
    F_CONT = 1 << 5, F_EXTRA = 1 << 6, F_MULTI_A = 1 << 7, F_MULTI_B = 1 << 8
};
enum { S_SOFT_CLIP = 1 << 0, S_FIRST = 1 << 1 };
const int TID_NONE = -2;
enum { FUNMAP = 0x4, FREVERSE = 0x10, FSECONDARY = 0x100, FQCFAIL = 0x200, FDUP = 0x400, FSUPPLEMENTARY = 0x800 };
/* reference parsers.py:960-962 */
const int MIN_NON_OVERLAP = 20, MIN_INDEL = 50, MAX_UNMAPPED_BASES = 50;

inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

/* CIGAR op classes, MIDNSHP=X */
inline bool consumes_ref(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
inline bool is_aligned(int op) { return op == 0 || op == 7 || op == 8; }
inline bool consumes_query_aln(int op) { return op == 0 || op == 1 || op == 7 || op == 8; }
inline bool is_clip(int op) { return op == 4 || op == 5; }

/* ------------------------------------------------------------------------------------------ */
/* BGZF: random access by virtual offset, a small cache of inflated blocks                      */
/* ------------------------------------------------------------------------------------------ */
struct Block { std::vector<uint8_t> data; uint32_t csize = 0; uint64_t stamp = 0; };

class Bgzf {
public:
    ~Bgzf() { if (f_) fclose(f_); }
    bool open(const char *path) { f_ = fopen(path, "rb"); return f_ != nullptr; }
    bool seek(uint64_t voff) { if (!load(voff >> 16)) return false; uoff_ = (uint32_t)(voff & 0xFFFF); return true; }
    uint64_t tell() const
    {
        if (cur_ && uoff_ >= cur_->data.size() && cur_->csize) return (coff_ + cur_->csize) << 16;
        return (coff_ << 16) | uoff_;
    }
    /* read up to n bytes; returns the number read, or -1 on a corrupt file */
    long read(uint8_t *dst, size_t n)
    {
        size_t got = 0;
        while (n > 0) {
            if (!cur_) return -1;
            const size_t avail = cur_->data.size() > uoff_ ? cur_->data.size() - uoff_ : 0;
            if (avail == 0) {
                if (cur_->csize == 0) break;                  /* end of file */
                if (!load(coff_ + cur_->csize)) return -1;
                uoff_ = 0;
                if (cur_->csize == 0) break;
                continue;
            }
            const size_t take = avail < n ? avail : n;
            memcpy(dst + got, cur_->data.data() + uoff_, take);
            uoff_ += (uint32_t)take; got += take; n -= take;
        }
        return (long)got;
    }

private:
    bool load(uint64_t coff)
    {
        auto it = cache_.find(coff);
        if (it != cache_.end()) { cur_ = &it->second; cur_->stamp = ++clock_; coff_ = coff; return true; }
        if (cache_.size() >= 512) {                           /* drop the older half */
            std::vector<std::pair<uint64_t, uint64_t>> age;
            for (auto &kv : cache_) age.push_back({kv.second.stamp, kv.first});
            std::sort(age.begin(), age.end());
            for (size_t i = 0; i < age.size() / 2; ++i) cache_.erase(age[i].second);
        }
        Block b;
        uint8_t hdr[18];
        if (fseeko(f_, (off_t)coff, SEEK_SET) != 0) return false;
        const size_t nh = fread(hdr, 1, 18, f_);
        if (nh == 18) {
            if (!(hdr[0] == 0x1f && hdr[1] == 0x8b && hdr[2] == 8 && hdr[3] == 4)) return false;
            const unsigned xlen = rd16(hdr + 10);
            std::vector<uint8_t> extra(xlen);
            memcpy(extra.data(), hdr + 12, xlen < 6 ? xlen : 6);
            if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f_) != xlen - 6) return false;
            int bsize = -1;
            for (size_t p = 0; p + 4 <= extra.size();) {
                const unsigned slen = rd16(extra.data() + p + 2);
                if (extra[p] == 66 && extra[p + 1] == 67 && p + 6 <= extra.size()) bsize = rd16(extra.data() + p + 4);
                p += 4 + slen;
            }
            if (bsize < 0) return false;
            b.csize = (uint32_t)bsize + 1;
            const long plen = (long)b.csize - 12 - (long)xlen - 8;
            if (plen < 0) return false;
            std::vector<uint8_t> payload((size_t)plen);
            uint8_t tail[8];
            if (plen && fread(payload.data(), 1, (size_t)plen, f_) != (size_t)plen) return false;
            if (fread(tail, 1, 8, f_) != 8) return false;
            const uint32_t isize = rd32(tail + 4);
            b.data.resize(isize);
            if (isize) {
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) return false;
                zs.next_in = payload.data(); zs.avail_in = (uInt)plen;
                zs.next_out = b.data.data(); zs.avail_out = isize;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.total_out != isize) return false;
            }
        }                                                     /* short read: end of file, empty block */
        b.stamp = ++clock_;
        auto ins = cache_.emplace(coff, std::move(b));
        cur_ = &ins.first->second; coff_ = coff;
        return true;
    }

    FILE *f_ = nullptr;
    std::unordered_map<uint64_t, Block> cache_;
    Block *cur_ = nullptr;
    uint64_t coff_ = 0, clock_ = 0;
    uint32_t uoff_ = 0;
};

/* ------------------------------------------------------------------------------------------ */
/* BAI                                                                                          */
/* ------------------------------------------------------------------------------------------ */
struct RefIndex {
    std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear;
};

struct Bai {
    std::vector<RefIndex> refs;

    bool load(const char *path)
    {
        FILE *f = fopen(path, "rb");
        if (!f) return false;
        std::vector<uint8_t> b;
        uint8_t tmp[65536];
        size_t n;
        while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) b.insert(b.end(), tmp, tmp + n);
        fclose(f);
        if (b.size() < 8 || memcmp(b.data(), "BAI\1", 4) != 0) return false;
        size_t p = 4;
        const int n_ref = (int)rd32(&b[p]); p += 4;
        refs.resize(n_ref < 0 ? 0 : n_ref);
        for (auto &r : refs) {
            if (p + 4 > b.size()) return false;
            const int n_bin = (int)rd32(&b[p]); p += 4;
            for (int i = 0; i < n_bin; ++i) {
                if (p + 8 > b.size()) return false;
                const uint32_t bin = rd32(&b[p]);
                const int n_chunk = (int)rd32(&b[p + 4]); p += 8;
                if (n_chunk < 0 || p + 16 * (size_t)n_chunk > b.size()) return false;
                if (bin != 37450) {                               /* 37450: htslib's per-reference counts */
                    auto &v = r.bins[bin];
                    for (int k = 0; k < n_chunk; ++k) v.push_back({rd64(&b[p + 16 * k]), rd64(&b[p + 16 * k + 8])});
                }
                p += 16 * (size_t)n_chunk;
            }
            if (p + 4 > b.size()) return false;
            const int n_intv = (int)rd32(&b[p]); p += 4;
            if (n_intv < 0 || p + 8 * (size_t)n_intv > b.size()) return false;
            r.linear.resize(n_intv);
            for (int k = 0; k < n_intv; ++k) r.linear[k] = rd64(&b[p + 8 * k]);
            p += 8 * (size_t)n_intv;
        }
        return true;
    }

    /* merged, sorted chunks that may hold records overlapping [beg, end) */
    std::vector<std::pair<uint64_t, uint64_t>> chunks(int tid, int64_t beg, int64_t end) const
    {
        std::vector<std::pair<uint64_t, uint64_t>> out;
        if (tid < 0 || tid >= (int)refs.size()) return out;
        const RefIndex &r = refs[tid];
        const size_t w = (size_t)(beg >> 14);
        const uint64_t min_off = w < r.linear.size() ? r.linear[w] : (r.linear.empty() ? 0 : r.linear.back());
        auto add_bin = [&](uint32_t bin) {
            auto it = r.bins.find(bin);
            if (it == r.bins.end()) return;
            for (auto &c : it->second) if (c.second > min_off) out.push_back(c);
        };
        const int64_t e = end - 1;
        add_bin(0);
        const int shift[5] = {26, 23, 20, 17, 14};
        const uint32_t base[5] = {1, 9, 73, 585, 4681};
        for (int l = 0; l < 5; ++l)
            for (int64_t k = base[l] + (beg >> shift[l]); k <= base[l] + (e >> shift[l]); ++k) add_bin((uint32_t)k);
        std::sort(out.begin(), out.end());
        std::vector<std::pair<uint64_t, uint64_t>> merged;
        for (auto &c : out) {
            if (!merged.empty() && c.first <= merged.back().second) {
                if (c.second > merged.back().second) merged.back().second = c.second;
            } else merged.push_back(c);
        }
        return merged;
    }
};

/* ------------------------------------------------------------------------------------------ */
/* one BAM record (what the gatherer keeps of it)                                               */
/* ------------------------------------------------------------------------------------------ */
struct Read {
    int32_t tid = -1, pos = 0, end = 0, mapq = 0, flag = 0, l_seq = 0, tlen = 0;
    std::string qname;
    std::vector<std::pair<int, int>> cigar;       /* (op, len) */
    std::string rg, sa;
    bool has_rg = false, has_sa = false;
    bool is_reverse() const { return (flag & FREVERSE) != 0; }
};

/* scan the aux block for the RG and SA strings; false on a malformed block */
bool parse_tags(const uint8_t *b, size_t n, Read &r)
{
    size_t p = 0;
    while (p + 3 <= n) {
        const char k0 = (char)b[p], k1 = (char)b[p + 1], t = (char)b[p + 2];
        p += 3;
        size_t sz = 0;
        switch (t) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'Z': case 'H': {
            const void *e = memchr(b + p, 0, n - p);
            if (!e) return false;
            const size_t len = (const uint8_t *)e - (b + p);
            if (t == 'Z' && k0 == 'R' && k1 == 'G') { r.rg.assign((const char *)b + p, len); r.has_rg = true; }
            if (t == 'Z' && k0 == 'S' && k1 == 'A') { r.sa.assign((const char *)b + p, len); r.has_sa = true; }
            p += len + 1;
            continue;
        }
        case 'B': {
            if (p + 5 > n) return false;
            const char sub = (char)b[p];
            const uint32_t cnt = rd32(b + p + 1);
            const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            sz = 5 + es * (size_t)cnt;
            break;
        }
        default: return false;
        }
        if (p + sz > n) return false;
        p += sz;
    }
    return true;
}

/* what every reader of one BAM shares (read-only after open) */
struct Shared {
    std::string path;
    Bai bai;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::unordered_map<std::string, int> tid_of;
    uint64_t first_record = 0;
};

/* one reader: its own file handle, block cache and record buffer (one per worker thread) */
/* insert-size table of one library in first-seen order (the order a Python dict would keep) */
struct LibScan {
    std::vector<int32_t> keys;
    std::vector<int64_t> counts;
    std::unordered_map<int32_t, size_t> slot;
};

struct svgt_bam_impl {
    Bgzf bgzf;
    std::shared_ptr<Shared> sh;
    std::vector<uint8_t> buf;
    std::vector<int32_t> frags, splits;               /* rows of the last svgt_pack_sites() */
    std::vector<LibScan> scans;                       /* tables of the last svgt_bam_scan_libraries() */
};

/* next record at the reader's position: 1 = ok, 0 = end of file, -1 = corrupt */
int next_record(svgt_bam_impl &B, Read &r, bool want_tags)
{
    uint8_t szb[4];
    const long g = B.bgzf.read(szb, 4);
    if (g < 0) return -1;
    if (g < 4) return 0;
    const uint32_t sz = rd32(szb);
    if (sz < 32 || sz > (1u << 28)) return -1;
    B.buf.resize(sz);
    if (B.bgzf.read(B.buf.data(), sz) != (long)sz) return -1;
    const uint8_t *b = B.buf.data();
    r.tid = (int32_t)rd32(b); r.pos = (int32_t)rd32(b + 4);
    const unsigned l_name = b[8];
    r.mapq = b[9];
    const unsigned n_cig = rd16(b + 12);
    r.flag = rd16(b + 14);
    r.l_seq = (int32_t)rd32(b + 16);
    r.tlen = (int32_t)rd32(b + 28);
    size_t p = 32;
    if (p + l_name + 4 * (size_t)n_cig > sz || l_name == 0) return -1;
    r.qname.assign((const char *)b + p, l_name - 1);
    p += l_name;
    r.cigar.resize(n_cig);
    int64_t e = r.pos;
    for (unsigned i = 0; i < n_cig; ++i) {
        const uint32_t v = rd32(b + p + 4 * i);
        r.cigar[i] = {(int)(v & 0xF), (int)(v >> 4)};
        if (consumes_ref((int)(v & 0xF))) e += (int)(v >> 4);
    }
    r.end = (int32_t)e;
    p += 4 * (size_t)n_cig;
    p += ((size_t)r.l_seq + 1) / 2 + (size_t)r.l_seq;
    r.has_rg = r.has_sa = false;
    if (want_tags) {
        if (p > sz) return -1;
        if (!parse_tags(b + p, sz - p, r)) return -1;
    }
    return 1;
}

/* pysam fetch(): file-order records of contig `tid` with pos < end and reference_end > beg.
 * `fn(read)` returns false to stop.  Returns 0 or a negative error. */
template <class Fn>
int fetch(svgt_bam_impl &B, int tid, int64_t beg, int64_t end, bool want_tags, Fn fn)
{
    if (tid < 0 || tid >= (int)B.sh->ref_names.size()) return fail(SVGT_PACK_ERR_ARG, "invalid contig id %d", tid);
    if (beg < 0) beg = 0;
    if (end <= beg) return 0;
    Read r;
    for (auto &c : B.sh->bai.chunks(tid, beg, end)) {
        if (!B.bgzf.seek(c.first)) return fail(SVGT_PACK_ERR_IO, "BGZF seek failed");
        while (B.bgzf.tell() < c.second) {
            const int rc = next_record(B, r, want_tags);
            if (rc < 0) return fail(SVGT_PACK_ERR_IO, "corrupt BAM record");
            if (rc == 0) break;
            if (r.tid != tid || r.pos >= end) return 0;       /* ends the whole query */
            int64_t rend = r.end;
            if ((r.flag & FUNMAP) || r.cigar.empty() || rend <= r.pos) rend = (int64_t)r.pos + 1;
            if (rend > beg) { if (!fn(r)) return 0; }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* split candidates (gather.py: Piece, split_candidate)                                         */
/* ------------------------------------------------------------------------------------------ */
struct Piece {
    int tid = TID_NONE;            /* TID_NONE: chrom None / not in the header */
    std::string chrom;             /* name as the reference compares it        */
    bool has_chrom = false;
    int start = 0, end = 0, mapq = 0;
    bool rev = false;
    int qstart = 0, qend = 0, qlen = 0;

    void span(const std::vector<std::pair<int, int>> &cigar)
    {
        int s = 0, e = 0, l = 0;
        const int n = (int)cigar.size();
        for (int i = 0; i < n; ++i) {
            const auto &c = rev ? cigar[n - 1 - i] : cigar[i];
            if (is_clip(c.first)) { if (i == 0) { s += c.second; e += c.second; } l += c.second; }
            else if (consumes_query_aln(c.first)) { e += c.second; l += c.second; }
        }
        qstart = s; qend = e; qlen = l;
    }
    int start_diagonal() const { return start - (rev ? (qlen - qend) : qstart); }
    int end_diagonal() const { return end - (rev ? (qlen - qstart) : qend); }
};

struct Split { Piece left, right; bool soft = false; };

bool left_clipped(const std::vector<std::pair<int, int>> &cigar)
{
    const bool l = is_clip(cigar.front().first), r = is_clip(cigar.back().first);
    return (l && !r) || (l && r && cigar.front().second > cigar.back().second);
}

/* '36M2D64M' -> ops, as the regex (\d+)([MIDNSHP=X]) finds them */
std::vector<std::pair<int, int>> cigar_from_string(const std::string &t)
{
    static const char ops[] = "MIDNSHP=X";
    std::vector<std::pair<int, int>> out;
    size_t i = 0;
    while (i < t.size()) {
        if (t[i] < '0' || t[i] > '9') { ++i; continue; }
        long v = 0;
        size_t j = i;
        while (j < t.size() && t[j] >= '0' && t[j] <= '9') { v = v * 10 + (t[j] - '0'); ++j; }
        if (j < t.size()) {
            const char *q = strchr(ops, t[j]);
            if (q && t[j]) { out.push_back({(int)(q - ops), (int)v}); i = j + 1; continue; }
        }
        /* digits not followed by an op: the regex restarts one character further */
        i = i + 1;
    }
    return out;
}

/* 1 = candidate in `out`, 0 = none, negative = error */
int split_candidate(const svgt_bam_impl &B, const Read &r, Split &out)
{
    if (r.cigar.empty()) return fail(SVGT_PACK_ERR_RECORD, "mapped read %s has no CIGAR", r.qname.c_str());
    Piece own;
    own.tid = r.tid; own.has_chrom = r.tid >= 0 && r.tid < (int)B.sh->ref_names.size();
    if (own.has_chrom) own.chrom = B.sh->ref_names[r.tid];
    own.start = r.pos; own.end = r.end; own.rev = r.is_reverse(); own.mapq = r.mapq;
    own.span(r.cigar);
    if (!r.has_sa) {
        const bool first = is_clip(r.cigar.front().first), last = is_clip(r.cigar.back().first);
        if (!(first || last)) return 0;
        const int longest = std::max(first ? r.cigar.front().second : 0, last ? r.cigar.back().second : 0);
        int qal = 0;
        for (auto &c : r.cigar) if (consumes_query_aln(c.first)) qal += c.second;
        if (longest > 0 && (r.l_seq - qal) <= MAX_UNMAPPED_BASES) {
            Piece ghost;
            ghost.tid = TID_NONE; ghost.start = 1; ghost.end = 1; ghost.rev = own.rev; ghost.mapq = 0;
            ghost.span(r.cigar);
            out.soft = true;
            if (left_clipped(r.cigar)) { out.left = ghost; out.right = own; }
            else { out.left = own; out.right = ghost; }
            return 1;
        }
        return 0;
    }
    /* entries = SA.rstrip(';').split(';'); more than one supplementary alignment: not a candidate */
    std::string sa = r.sa;
    while (!sa.empty() && sa.back() == ';') sa.pop_back();
    if (sa.find(';') != std::string::npos) return 0;
    std::vector<std::string> f;
    size_t s0 = 0;
    while (f.size() < 5) {
        const size_t c = sa.find(',', s0);
        if (c == std::string::npos) { f.push_back(sa.substr(s0)); break; }
        f.push_back(sa.substr(s0, c - s0));
        s0 = c + 1;
    }
    if (f.size() < 5) return fail(SVGT_PACK_ERR_RECORD, "malformed SA tag on read %s", r.qname.c_str());
    char *endp = nullptr;
    const long sa_pos = strtol(f[1].c_str(), &endp, 10);
    if (endp == f[1].c_str()) return fail(SVGT_PACK_ERR_RECORD, "malformed SA position on read %s", r.qname.c_str());
    const long sa_mapq = strtol(f[4].c_str(), &endp, 10);
    if (endp == f[4].c_str()) return fail(SVGT_PACK_ERR_RECORD, "malformed SA mapq on read %s", r.qname.c_str());
    const auto mcig = cigar_from_string(f[3]);
    Piece mate;
    mate.chrom = f[0]; mate.has_chrom = true;
    auto it = B.sh->tid_of.find(f[0]);
    mate.tid = it == B.sh->tid_of.end() ? TID_NONE : it->second;
    mate.start = (int)sa_pos - 1;                     /* SA is one-based */
    int e = mate.start;
    for (auto &c : mcig) if (consumes_ref(c.first)) e += c.second;
    mate.end = e; mate.rev = f[2] == "-"; mate.mapq = (int)sa_mapq;
    mate.span(mcig);
    bool mate_left;
    if (own.has_chrom && own.chrom == mate.chrom) mate_left = own.start > mate.start;
    else mate_left = left_clipped(r.cigar);
    const Piece &L = mate_left ? mate : own, &R = mate_left ? own : mate;
    /* the two pieces must each cover enough of the read that the other does not */
    const int overlap = std::max(0, 1 + std::min(L.qend, R.qend) - std::max(L.qstart, R.qstart));
    if (std::min(1 + L.qend - L.qstart - overlap, 1 + R.qend - R.qstart - overlap) < MIN_NON_OVERLAP) return 0;
    if (L.has_chrom && R.has_chrom && L.chrom == R.chrom && L.rev == R.rev) {
        const int ins = L.rev ? R.end_diagonal() - L.start_diagonal() : L.end_diagonal() - R.start_diagonal();
        if (std::abs(ins) < MIN_INDEL) return 0;
        const int desert = R.qstart - L.qend - 1;
        if (desert > 0 && desert - std::max(0, ins) > MAX_UNMAPPED_BASES) return 0;
    }
    out.soft = false; out.left = L; out.right = R;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* fragments and rows (gather.py: Fragment, _collect; evidence.py: BatchPacker._pack_fragment)   */
/* ------------------------------------------------------------------------------------------ */
struct Prim {
    int32_t tid, pos, end, mapq;
    bool rev;
    std::vector<std::pair<int, int>> iv;          /* merged gap-free aligned intervals */
};

struct Fragment {
    int lib = 0;
    std::vector<Prim> prim;
    std::vector<Split> splits;
    std::unordered_set<int> seen_flags;           /* (query_name, flag) seen; the name is the map key */
};

Prim make_prim(const Read &r)
{
    Prim p;
    p.tid = r.tid; p.pos = r.pos; p.end = r.end; p.mapq = r.mapq; p.rev = r.is_reverse();
    int pos = r.pos;
    for (auto &c : r.cigar) {
        if (is_aligned(c.first)) {
            const int s = pos, e = pos + c.second;
            if (!p.iv.empty() && s == p.iv.back().second) p.iv.back().second = e;
            else if (e > s) p.iv.push_back({s, e});
        }
        if (consumes_ref(c.first)) pos += c.second;
    }
    return p;
}

struct Gather {
    svgt_bam_impl &B;
    const std::unordered_map<std::string, int> &rg_lib;
    const uint8_t *lib_active;
    int n_lib;
    std::map<std::string, Fragment> frags;        /* std::map: iteration = sorted(query_name), bytewise */
    int err = 0;

    /* one fetched record (gather.py _collect body); false = stop with `err` set */
    bool add(const Read &r)
    {
        if (r.flag & (FUNMAP | FDUP)) return true;
        if (!r.has_rg) { err = fail(SVGT_PACK_ERR_RG, "read %s has no RG tag", r.qname.c_str()); return false; }
        auto it = rg_lib.find(r.rg);
        if (it == rg_lib.end() || it->second < 0 || it->second >= n_lib) {
            err = fail(SVGT_PACK_ERR_RG, "read group %s is not in the library table", r.rg.c_str());
            return false;
        }
        if (!lib_active[it->second]) return true;
        return keep(r, it->second);
    }
    /* the part after the library filter; split out so the classic limit test can sit between them */
    bool keep(const Read &r, int lib)
    {
        auto ins = frags.find(r.qname);
        if (ins == frags.end()) { ins = frags.emplace(r.qname, Fragment()).first; ins->second.lib = lib; }
        Fragment &f = ins->second;
        if (!f.seen_flags.insert(r.flag).second) return true;
        if (r.flag & (FSUPPLEMENTARY | FSECONDARY)) return true;
        f.prim.push_back(make_prim(r));
        Split sp;
        const int rc = split_candidate(B, r, sp);
        if (rc < 0) { err = rc; return false; }
        if (rc > 0) f.splits.push_back(sp);
        return true;
    }
};

void push_row(std::vector<int32_t> &v, const int32_t w[8]) { v.insert(v.end(), w, w + 8); }

void pack_fragment(const Fragment &f, int ordinal, std::vector<int32_t> &frows, std::vector<int32_t> &srows, int &nf,
                   int &ns)
{
    struct Group { const Prim *a, *b; int fl; };
    std::vector<Group> groups;
    if (f.prim.size() == 2) groups.push_back({&f.prim[0], &f.prim[1], F_PAIRED});
    else for (size_t i = 0; i < f.prim.size(); ++i) groups.push_back({&f.prim[i], nullptr, i > 0 ? F_CONT : 0});
    const int lib = f.lib & 0xFFFF;
    for (auto &g : groups) {
        static const std::vector<std::pair<int, int>> none;
        auto multi = [](const Prim *p) {
            return p && !(p->iv.size() == 1 && p->iv[0].first == p->pos && p->iv[0].second == p->end);
        };
        const bool ma = multi(g.a), mb = multi(g.b);
        const auto &xa = ma ? g.a->iv : none;
        const auto &xb = mb ? g.b->iv : none;
        for (size_t k = 0; k < std::max(xa.size(), xb.size()); ++k) {
            int32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int efl = F_EXTRA | (g.fl & F_CONT);
            if (k < xa.size()) { efl |= F_HAS_A; w[0] = xa[k].first; w[1] = xa[k].second; w[4] = g.a->tid; }
            if (k < xb.size()) { efl |= F_HAS_B; w[2] = xb[k].first; w[3] = xb[k].second; w[5] = g.b->tid; }
            w[6] = (int32_t)((uint32_t)lib << 16);
            w[7] = efl;
            push_row(frows, w); ++nf;
        }
        int32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int fl = g.fl;
        uint32_t w6 = 0;
        if (g.a) {
            fl |= F_HAS_A | (g.a->rev ? F_REV_A : 0) | (ma ? F_MULTI_A : 0);
            w[0] = g.a->pos; w[1] = g.a->end; w[4] = g.a->tid;
            w6 |= (uint32_t)std::min(g.a->mapq, 255);
        }
        if (g.b) {
            fl |= F_HAS_B | (g.b->rev ? F_REV_B : 0) | (mb ? F_MULTI_B : 0);
            w[2] = g.b->pos; w[3] = g.b->end; w[5] = g.b->tid;
            w6 |= (uint32_t)std::min(g.b->mapq, 255) << 8;
        }
        w6 |= (uint32_t)lib << 16;
        w[6] = (int32_t)w6; w[7] = fl;
        push_row(frows, w); ++nf;
    }
    for (size_t i = 0; i < f.splits.size(); ++i) {
        const Split &s = f.splits[i];
        const int sfl = (s.soft ? S_SOFT_CLIP : 0) | (i == 0 ? S_FIRST : 0);
        const uint32_t meta = (uint32_t)std::min(s.left.mapq, 255) | ((uint32_t)std::min(s.right.mapq, 255) << 8) |
                              ((uint32_t)sfl << 16);
        const int32_t w[8] = {s.left.tid < 0 ? TID_NONE : s.left.tid, s.left.start, s.left.end,
                              s.right.tid < 0 ? TID_NONE : s.right.tid, s.right.start, s.right.end, (int32_t)meta,
                              ordinal};
        push_row(srows, w); ++ns;
    }
}

int64_t count_region(svgt_bam_impl &B, int tid, int64_t beg, int64_t end, bool filter_all)
{
    int64_t n = 0;
    const int rc = fetch(B, tid, beg, end, false, [&](const Read &r) {
        if (!filter_all || !(r.flag & (FUNMAP | FSECONDARY | FQCFAIL | FDUP))) ++n;
        return true;
    });
    return rc < 0 ? rc : n;
}

struct PackJob {
    const std::unordered_map<std::string, int> &rgmap;
    const uint8_t *lib_active;
    int n_lib, mode;
    int64_t max_reads;
};

/* gather + pack one breakpoint with reader B; rows are appended to frows / srows */
int pack_one_site(svgt_bam_impl &B, const PackJob &J, const svgt_pack_site_t &S, svgt_pack_count_t &C,
                  std::vector<int32_t> &frows, std::vector<int32_t> &srows)
{
    C.n_frag_rows = C.n_split_rows = C.skip = C.n_fragments = 0;
    const int tid[2] = {S.tidA, S.tidB};
    const int64_t beg[2] = {S.begA, S.begB}, end[2] = {S.endA, S.endB};
    Gather G{B, J.rgmap, J.lib_active, J.n_lib, {}, 0};
    bool too_many = false;
    if (J.mode == SVGT_PACK_MODE_SSO) {
        if (J.max_reads >= 0) {
            for (int s = 0; s < 2 && !too_many; ++s) {
                const int64_t n = count_region(B, tid[s], beg[s], end[s], true);
                if (n < 0) return (int)n;
                if (n > J.max_reads) too_many = true;
            }
        }
        for (int s = 0; s < 2 && !too_many; ++s) {
            const int rc = fetch(B, tid[s], beg[s], end[s], true, [&](const Read &r) { return G.add(r); });
            if (rc < 0) return rc;
            if (G.err) return G.err;
        }
    } else {
        /* classic.py:79-91: the limit is on the record's index in the fetch, tested after the
         * duplicate / library filters, so a filtered record never trips it */
        for (int s = 0; s < 2 && !too_many; ++s) {
            int64_t i = 0;
            const int rc = fetch(B, tid[s], beg[s], end[s], true, [&](const Read &r) {
                const int64_t idx = i++;
                if (r.flag & (FUNMAP | FDUP)) return true;
                if (!r.has_rg) { G.err = fail(SVGT_PACK_ERR_RG, "read %s has no RG tag", r.qname.c_str()); return false; }
                auto it = J.rgmap.find(r.rg);
                if (it == J.rgmap.end() || it->second < 0 || it->second >= J.n_lib) {
                    G.err = fail(SVGT_PACK_ERR_RG, "read group %s is not in the library table", r.rg.c_str());
                    return false;
                }
                if (!J.lib_active[it->second]) return true;
                if (J.max_reads >= 0 && idx > J.max_reads) { too_many = true; return false; }
                return G.keep(r, it->second);
            });
            if (rc < 0) return rc;
            if (G.err) return G.err;
        }
    }
    if (too_many) { C.skip = 1; return 0; }
    int ordinal = 0, nf = 0, ns = 0;
    for (auto &kv : G.frags) pack_fragment(kv.second, ordinal++, frows, srows, nf, ns);
    C.n_frag_rows = nf; C.n_split_rows = ns; C.n_fragments = ordinal;
    return 0;
}

}  // namespace

struct svgt_bam { svgt_bam_impl impl; };

extern "C" {

int svgt_pack_abi_version(void) { return SVGT_PACK_ABI_VERSION; }
const char *svgt_pack_last_error(void) { return g_err; }

int svgt_bam_open(const char *bam_path, const char *bai_path, svgt_bam_t **out)
{
    if (!bam_path || !out) return fail(SVGT_PACK_ERR_ARG, "null argument");
    *out = nullptr;
    svgt_bam *h = new (std::nothrow) svgt_bam();
    if (!h) return fail(SVGT_PACK_ERR_IO, "out of memory");
    svgt_bam_impl &B = h->impl;
    B.sh = std::make_shared<Shared>();
    Shared &S = *B.sh;
    S.path = bam_path;
    auto bail = [&](int code, const char *msg) { delete h; return fail(code, "%s: %s", msg, bam_path); };
    if (!B.bgzf.open(bam_path)) return bail(SVGT_PACK_ERR_IO, "cannot open");
    if (!B.bgzf.seek(0)) return bail(SVGT_PACK_ERR_IO, "not a BGZF file");
    uint8_t w[8];
    if (B.bgzf.read(w, 8) != 8 || memcmp(w, "BAM\1", 4) != 0) return bail(SVGT_PACK_ERR_IO, "not a BAM file");
    const uint32_t l_text = rd32(w + 4);
    std::vector<uint8_t> text(l_text);
    if (l_text && B.bgzf.read(text.data(), l_text) != (long)l_text) return bail(SVGT_PACK_ERR_IO, "truncated header");
    if (B.bgzf.read(w, 4) != 4) return bail(SVGT_PACK_ERR_IO, "truncated header");
    const int n_ref = (int)rd32(w);
    for (int i = 0; i < n_ref; ++i) {
        if (B.bgzf.read(w, 4) != 4) return bail(SVGT_PACK_ERR_IO, "truncated reference list");
        const uint32_t l_name = rd32(w);
        std::vector<uint8_t> nm(l_name + 4);
        if (l_name == 0 || l_name > 65536 || B.bgzf.read(nm.data(), l_name + 4) != (long)(l_name + 4))
            return bail(SVGT_PACK_ERR_IO, "truncated reference list");
        S.ref_names.emplace_back((const char *)nm.data(), l_name - 1);
        S.ref_lens.push_back((int32_t)rd32(nm.data() + l_name));
        S.tid_of[S.ref_names.back()] = i;         /* a repeated name keeps its last index, like the dict the Python reader builds */
    }
    S.first_record = B.bgzf.tell();
    std::string cand[2];
    if (bai_path) cand[0] = bai_path;
    else {
        cand[0] = std::string(bam_path) + ".bai";
        std::string stem(bam_path);
        const size_t dot = stem.rfind('.'), slash = stem.rfind('/');
        if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) stem.resize(dot);
        cand[1] = stem + ".bai";
    }
    bool ok = false;
    for (auto &c : cand) if (!c.empty() && S.bai.load(c.c_str())) { ok = true; break; }
    if (!ok) return bail(SVGT_PACK_ERR_IO, "no readable .bai index for");
    *out = h;
    return SVGT_PACK_OK;
}

int svgt_bam_close(svgt_bam_t *bam)
{
    delete bam;
    return SVGT_PACK_OK;
}

int svgt_bam_n_references(const svgt_bam_t *bam) { return bam ? (int)bam->impl.sh->ref_names.size() : SVGT_PACK_ERR_ARG; }

const char *svgt_bam_reference_name(const svgt_bam_t *bam, int tid)
{
    if (!bam || tid < 0 || tid >= (int)bam->impl.sh->ref_names.size()) return nullptr;
    return bam->impl.sh->ref_names[tid].c_str();
}

int64_t svgt_bam_reference_length(const svgt_bam_t *bam, int tid)
{
    if (!bam || tid < 0 || tid >= (int)bam->impl.sh->ref_lens.size()) return SVGT_PACK_ERR_ARG;
    return bam->impl.sh->ref_lens[tid];
}

int64_t svgt_bam_count(svgt_bam_t *bam, int tid, int64_t beg, int64_t end, int filter_all)
{
    if (!bam) return fail(SVGT_PACK_ERR_ARG, "null bam");
    return count_region(bam->impl, tid, beg, end, filter_all != 0);
}

int svgt_pack_sites(svgt_bam_t *bam, const svgt_pack_site_t *sites, int64_t n_sites, const char *const *rg_names,
                    const int32_t *rg_lib, int32_t n_rg, const uint8_t *lib_active, int32_t n_lib, int32_t mode,
                    int64_t max_reads, int32_t n_threads, svgt_pack_count_t *counts)
{
    if (!bam || n_sites < 0 || (n_sites && (!sites || !counts)) || n_rg < 0 || (n_rg && (!rg_names || !rg_lib)) ||
        n_lib < 0 || (n_lib && !lib_active))
        return fail(SVGT_PACK_ERR_ARG, "bad argument");
    if (mode != SVGT_PACK_MODE_SSO && mode != SVGT_PACK_MODE_CLASSIC) return fail(SVGT_PACK_ERR_ARG, "bad mode %d", mode);
    svgt_bam_impl &B = bam->impl;
    B.frags.clear(); B.splits.clear();
    std::unordered_map<std::string, int> rgmap;
    for (int i = 0; i < n_rg; ++i) if (rg_names[i]) rgmap[rg_names[i]] = rg_lib[i];
    PackJob job{rgmap, lib_active, n_lib, mode, max_reads};

    /* sites are independent: blocks of consecutive sites go to worker threads, each with its own reader
     * (file handle + block cache); rows are stitched back in site order */
    const int64_t kBlock = 16;
    const int64_t n_blocks = (n_sites + kBlock - 1) / kBlock;
    int nt = n_threads <= 0 ? (int)std::thread::hardware_concurrency() : n_threads;
    if (nt < 1) nt = 1;
    if ((int64_t)nt > n_blocks) nt = (int)(n_blocks < 1 ? 1 : n_blocks);
    if (nt == 1) {
        for (int64_t si = 0; si < n_sites; ++si) {
            const int rc = pack_one_site(B, job, sites[si], counts[si], B.frags, B.splits);
            if (rc < 0) return rc;
        }
        return SVGT_PACK_OK;
    }
    struct BlockOut { std::vector<int32_t> frags, splits; };
    std::vector<BlockOut> outs((size_t)n_blocks);
    std::atomic<int64_t> cursor(0);
    std::atomic<int> first_err(0);
    std::vector<std::string> errs((size_t)nt);
    auto worker = [&](int w) {
        svgt_bam_impl R;
        R.sh = B.sh;
        if (!R.bgzf.open(B.sh->path.c_str())) {
            int z = 0;
            if (first_err.compare_exchange_strong(z, SVGT_PACK_ERR_IO)) errs[w] = "cannot reopen " + B.sh->path;
            return;
        }
        for (;;) {
            const int64_t b = cursor.fetch_add(1);
            if (b >= n_blocks || first_err.load() != 0) return;
            for (int64_t si = b * kBlock; si < std::min(n_sites, (b + 1) * kBlock); ++si) {
                const int rc = pack_one_site(R, job, sites[si], counts[si], outs[b].frags, outs[b].splits);
                if (rc < 0) {
                    int z = 0;
                    if (first_err.compare_exchange_strong(z, rc)) errs[w] = g_err;    /* g_err is thread-local */
                    return;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int w = 0; w < nt; ++w) pool.emplace_back(worker, w);
    for (auto &t : pool) t.join();
    if (first_err.load() != 0) {
        for (auto &e : errs) if (!e.empty()) return fail(first_err.load(), "%s", e.c_str());
        return fail(first_err.load(), "worker failed");
    }
    size_t nf = 0, ns = 0;
    for (auto &o : outs) { nf += o.frags.size(); ns += o.splits.size(); }
    B.frags.reserve(nf); B.splits.reserve(ns);
    for (auto &o : outs) {
        B.frags.insert(B.frags.end(), o.frags.begin(), o.frags.end());
        B.splits.insert(B.splits.end(), o.splits.begin(), o.splits.end());
    }
    return SVGT_PACK_OK;
}

int svgt_bam_scan_libraries(svgt_bam_t *bam, const char *const *rg_names, const int32_t *rg_lib, int32_t n_rg,
                            int32_t n_lib, int64_t num_samp, int64_t read_length_reads, int64_t prevalence_records,
                            svgt_lib_scan_t *out)
{
    if (!bam || n_rg < 0 || (n_rg && (!rg_names || !rg_lib)) || n_lib <= 0 || !out)
        return fail(SVGT_PACK_ERR_ARG, "bad argument");
    svgt_bam_impl &B = bam->impl;
    std::unordered_map<std::string, int> rgmap;
    for (int i = 0; i < n_rg; ++i) if (rg_names[i] && rg_lib[i] >= 0 && rg_lib[i] < n_lib) rgmap[rg_names[i]] = rg_lib[i];
    B.scans.assign((size_t)n_lib, LibScan());
    struct St { int64_t rl_seen = 0, longest = 0, taken = 0, mine = 0; bool rl_done = false, ins_done = false; };
    std::vector<St> st((size_t)n_lib);
    for (auto &s : st) { s.ins_done = num_samp <= 0; }
    int64_t seen = 0;                                 /* records examined by the prevalence pass */
    if (!B.bgzf.seek(B.sh->first_record)) return fail(SVGT_PACK_ERR_IO, "BGZF seek failed");
    Read r;
    for (;;) {
        bool all_done = seen >= prevalence_records;
        for (auto &s : st) all_done = all_done && s.rl_done && s.ins_done;
        if (all_done) break;
        const int rc = next_record(B, r, true);
        if (rc < 0) return fail(SVGT_PACK_ERR_IO, "corrupt BAM record");
        if (rc == 0 || r.tid < 0) break;              /* fetch() without a region stops at the unplaced reads */
        if (!r.has_rg) return fail(SVGT_PACK_ERR_RG, "read %s has no RG tag", r.qname.c_str());
        auto it = rgmap.find(r.rg);
        const int lib = it == rgmap.end() ? -1 : it->second;
        if (seen < prevalence_records) {              /* Library.calc_lib_prevalence, parsers.py:555-576 */
            if (lib >= 0) ++st[lib].mine;
            ++seen;
        }
        if (lib < 0) continue;
        St &s = st[lib];
        if (!s.rl_done) {                             /* Library.calc_read_length, parsers.py:501-516 */
            int64_t n = 0;
            for (auto &c : r.cigar) if (c.first == 0 || c.first == 1 || c.first == 4 || c.first == 7 || c.first == 8) n += c.second;
            if (n > s.longest) s.longest = n;
            if (s.rl_seen == read_length_reads) s.rl_done = true;
            else ++s.rl_seen;
        }
        if (!s.ins_done) {                            /* Library.calc_insert_hist, parsers.py:518-553 */
            const bool skip = (r.flag & 0x10) || !(r.flag & 0x20) || (r.flag & 0x4) || (r.flag & 0x8) ||
                              (r.flag & FSUPPLEMENTARY) || (r.flag & FSECONDARY) || r.tlen <= 0;
            if (!skip) {
                LibScan &L = B.scans[lib];
                auto f = L.slot.find(r.tlen);
                if (f == L.slot.end()) { L.slot.emplace(r.tlen, L.keys.size()); L.keys.push_back(r.tlen); L.counts.push_back(1); }
                else ++L.counts[f->second];
                if (++s.taken == num_samp) s.ins_done = true;
            }
        }
    }
    for (int l = 0; l < n_lib; ++l) {
        out[l].read_length = st[l].longest;
        out[l].lib_records = st[l].mine;
        out[l].records_seen = seen;
        out[l].n_hist = (int64_t)B.scans[l].keys.size();
    }
    return SVGT_PACK_OK;
}

int svgt_bam_scan_hist(const svgt_bam_t *bam, int32_t lib, const int32_t **keys, const int64_t **counts, int64_t *n)
{
    if (!bam || !keys || !counts || !n || lib < 0 || lib >= (int)bam->impl.scans.size())
        return fail(SVGT_PACK_ERR_ARG, "bad argument");
    const LibScan &L = bam->impl.scans[lib];
    *keys = L.keys.data(); *counts = L.counts.data(); *n = (int64_t)L.keys.size();
    return SVGT_PACK_OK;
}

int svgt_pack_rows(const svgt_bam_t *bam, const int32_t **frags, int64_t *n_frag, const int32_t **splits,
                   int64_t *n_split)
{
    if (!bam || !frags || !n_frag || !splits || !n_split) return fail(SVGT_PACK_ERR_ARG, "null argument");
    *frags = bam->impl.frags.data(); *n_frag = (int64_t)(bam->impl.frags.size() / 8);
    *splits = bam->impl.splits.data(); *n_split = (int64_t)(bam->impl.splits.size() / 8);
    return SVGT_PACK_OK;
}


/* ------------------------------------------------------------------------------------------------ */
/* wide rows -> compact rows (svtyper_b200/compact.py is the specification and the parity checker)   */
/* ------------------------------------------------------------------------------------------------ */
}  /* extern "C" */
namespace {
enum : uint32_t {
    CW_CLS_A_ON_A = 1u << 28, CW_CLS_A_ON_B = 1u << 29, CW_CLS_B_ON_A = 1u << 30, CW_CLS_B_ON_B = 1u << 31,
    CW_PAIRED = 1u << 25, CW_REV_A = 1u << 26, CW_REV_B = 1u << 27, CW_CONT = 1u << 28, CW_MULTI_A = 1u << 30,
    CW_MULTI_B = 1u << 31, CW_SOFT = 1u << 16, CW_FIRST = 1u << 17, CW_WIDE = 1u << 18, CW_XEND = 1u << 19
};
const int CW_LEN_BITS = 14, CW_LEN_MAX = (1 << 14) - 1, CW_LIB_MAX = 511, CW_SLEN_MAX = 0xFFFF;

inline int64_t wide_off(const int32_t *s, int lo) { return (int64_t)(uint32_t)s[lo] | ((int64_t)s[lo + 1] << 32); }

/* one gap-free interval [s, e) on `tid` against both is_ref_seq windows (the oracle's ref_seq_hit()) */
inline bool interval_hit(int32_t tid, int64_t s, int64_t e, const int32_t *site, int m)
{
    const int64_t pA = site[0], pB = site[1];
    return (tid == site[6] && pA - m >= 0 && s <= pA - m && e >= pA + m) ||
           (tid == site[7] && pB - m >= 0 && s <= pB - m && e >= pB + m);
}

/* rows the compact form of one site needs; fills them when `out` is given */
struct SiteRows { int64_t nf, ns; };
int compact_site(const int32_t *site, const int32_t *frags, int64_t n_frag, const int32_t *splits, int64_t n_split, int m,
                 int32_t *out, SiteRows &cnt)
{
    cnt.nf = cnt.ns = 0;
    if (site[9] & 16) return 0;                          /* SKIP: no rows the path may look at */
    const int64_t foff = wide_off(site, 10), soff = wide_off(site, 13);
    const int64_t nf = site[12], ns = site[15];
    if (nf < 0 || ns < 0 || foff < 0 || soff < 0 || foff + nf > n_frag || soff + ns > n_split) return -1;
    int32_t *o = out;
    bool pendA = false, pendB = false;
    for (int64_t j = 0; j < nf; ++j) {
        const int32_t *f = frags + (foff + j) * 8;
        const int fl = f[7];
        const bool hasA = fl & F_HAS_A, hasB = fl & F_HAS_B, paired = fl & F_PAIRED;
        if (fl & F_EXTRA) {
            if (hasA && interval_hit(f[4], f[0], f[1], site, m)) pendA = true;
            if (hasB && interval_hit(f[5], f[2], f[3], site, m)) pendB = true;
            continue;
        }
        if (!hasA || (paired && !hasB)) return -2;
        const uint32_t lib = ((uint32_t)f[6] >> 16) & 0xFFFFu;
        if (lib > (uint32_t)CW_LIB_MAX) return -3;
        if (out) {
            const int64_t lenA = (int64_t)f[1] - f[0], lenB = (int64_t)f[3] - f[2];
            const bool mA = (fl & F_MULTI_A) || lenA < 0 || lenA > CW_LEN_MAX;
            const bool mB = hasB && ((fl & F_MULTI_B) || lenB < 0 || lenB > CW_LEN_MAX);
            const bool hitA = (fl & F_MULTI_A) ? pendA : interval_hit(f[4], f[0], f[1], site, m);
            const bool hitB = (fl & F_MULTI_B) ? pendB : (hasB && interval_hit(f[5], f[2], f[3], site, m));
            const uint32_t la = mA ? (hitA ? 1u : 0u) : (uint32_t)lenA;
            const uint32_t lb = mB ? (hitB ? 1u : 0u) : (hasB ? (uint32_t)lenB : 0u);
            uint32_t cls = (f[4] == site[6] ? CW_CLS_A_ON_A : 0u) | (f[4] == site[7] ? CW_CLS_A_ON_B : 0u);
            if (hasB) cls |= (f[5] == site[6] ? CW_CLS_B_ON_A : 0u) | (f[5] == site[7] ? CW_CLS_B_ON_B : 0u);
            uint32_t w3 = ((uint32_t)f[6] & 0xFFu) | (hasB ? ((uint32_t)f[6] & 0xFF00u) : 0u) | (lib << 16);
            w3 |= (paired ? CW_PAIRED : 0u) | ((fl & F_REV_A) ? CW_REV_A : 0u) | ((fl & F_REV_B) ? CW_REV_B : 0u) |
                  ((fl & F_CONT) ? CW_CONT : 0u) | (mA ? CW_MULTI_A : 0u) | (mB ? CW_MULTI_B : 0u);
            o[0] = f[0]; o[1] = hasB ? f[3] : 0;
            o[2] = (int32_t)(la | (lb << CW_LEN_BITS) | cls);
            o[3] = (int32_t)w3;
            o += 4;
        }
        pendA = pendB = false;
        ++cnt.nf;
    }
    for (int64_t j = 0; j < ns; ++j) {
        const int32_t *q = splits + (soff + j) * 8;
        const int64_t lenL = (int64_t)q[2] - q[1], lenR = (int64_t)q[5] - q[4];
        const bool wide = lenL < 0 || lenL > CW_SLEN_MAX || lenR < 0 || lenR > CW_SLEN_MAX;
        if (wide && cnt.ns % 32 == 31) {                 /* a WIDE row never sits in the last slot of a chunk */
            if (out) { o[0] = o[1] = o[2] = o[3] = 0; o += 4; }
            ++cnt.ns;
        }
        if (out) {
            const uint32_t sfl = ((uint32_t)q[6] >> 16) & 0xFFFFu;
            const uint32_t cls = (q[0] == site[6] ? 1u << 28 : 0u) | (q[0] == site[7] ? 1u << 29 : 0u) |
                                 (q[3] == site[6] ? 1u << 30 : 0u) | (q[3] == site[7] ? 1u << 31 : 0u);
            o[0] = q[1]; o[1] = q[4];
            o[2] = wide ? 0 : (int32_t)((uint32_t)lenL | ((uint32_t)lenR << 16));
            o[3] = (int32_t)(((uint32_t)q[6] & 0xFFFFu) | ((sfl & S_SOFT_CLIP) ? CW_SOFT : 0u) | ((sfl & S_FIRST) ? CW_FIRST : 0u) |
                             (wide ? CW_WIDE : 0u) | cls);
            o += 4;
        }
        ++cnt.ns;
        if (wide) {
            if (out) { o[0] = q[2]; o[1] = q[5]; o[2] = 0; o[3] = (int32_t)CW_XEND; o += 4; }
            ++cnt.ns;
        }
    }
    return 0;
}

template <typename Fn>
void parallel_for(int64_t n, int n_threads, Fn fn)
{
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    const int64_t block = 2048;
    if (n_threads == 1 || n <= block) { fn(0, n); return; }
    std::atomic<int64_t> next(0);
    std::vector<std::thread> th;
    auto work = [&]() {
        for (;;) {
            const int64_t lo = next.fetch_add(block);
            if (lo >= n) break;
            fn(lo, std::min(n, lo + block));
        }
    };
    for (int t = 0; t < n_threads - 1; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}
}  // namespace

extern "C" {

int svgt_compact_count(const int32_t *sites, int64_t n_sites, const int32_t *frags, int64_t n_frag, const int32_t *splits,
                       int64_t n_split, int32_t min_aligned, int32_t n_threads, int64_t *row_off, int32_t *counts)
{
    if (n_sites < 0 || (n_sites > 0 && (!sites || !row_off || !counts))) return fail(SVGT_PACK_ERR_ARG, "bad argument");
    std::atomic<int> err(0);
    parallel_for(n_sites, n_threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            SiteRows c;
            const int rc = compact_site(sites + i * 16, frags, n_frag, splits, n_split, min_aligned, nullptr, c);
            if (rc) err = rc;
            counts[2 * i] = (int32_t)c.nf; counts[2 * i + 1] = (int32_t)c.ns;
        }
    });
    if (err == -3) return fail(SVGT_PACK_ERR_ARG, "compact schema holds library indices up to %d", CW_LIB_MAX);
    if (err) return fail(SVGT_PACK_ERR_ARG, "malformed wide batch (row offsets out of range, or a fragment row without read A / a PAIRED row without read B)");
    int64_t off = 0;
    for (int64_t i = 0; i < n_sites; ++i) { row_off[i] = off; off += (int64_t)counts[2 * i] + counts[2 * i + 1]; }
    row_off[n_sites] = off;
    return SVGT_PACK_OK;
}

int svgt_compact_fill(const int32_t *sites, int64_t n_sites, const int32_t *frags, int64_t n_frag, const int32_t *splits,
                      int64_t n_split, int32_t min_aligned, int32_t n_threads, const int64_t *row_off, const int32_t *counts,
                      int32_t *out_sites, int32_t *out_rows)
{
    if (n_sites < 0 || (n_sites > 0 && (!sites || !row_off || !counts || !out_sites))) return fail(SVGT_PACK_ERR_ARG, "bad argument");
    if (n_sites > 0 && row_off[n_sites] > 0 && !out_rows) return fail(SVGT_PACK_ERR_ARG, "null row buffer");
    std::atomic<int> err(0);
    parallel_for(n_sites, n_threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const int32_t *s = sites + i * 16;
            SiteRows c;
            const int rc = compact_site(s, frags, n_frag, splits, n_split, min_aligned, out_rows + row_off[i] * 4, c);
            if (rc || c.nf != counts[2 * i] || c.ns != counts[2 * i + 1]) err = -1;
            int32_t *o = out_sites + i * 12;
            for (int k = 0; k < 6; ++k) o[k] = s[k];
            o[6] = s[8];
            o[7] = (s[9] & 0x1F) | (s[6] == s[7] ? 32 : 0);
            o[8] = (int32_t)((uint64_t)row_off[i] & 0xFFFFFFFFu); o[9] = (int32_t)((uint64_t)row_off[i] >> 32);
            o[10] = counts[2 * i]; o[11] = counts[2 * i + 1];
        }
    });
    if (err) return fail(SVGT_PACK_ERR_ARG, "wide batch changed between svgt_compact_count and svgt_compact_fill");
    return SVGT_PACK_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* FORMAT text of scored rows (reference parsers.py:375-398, singlesample.py:406-473)                */
/* ------------------------------------------------------------------------------------------------ */
}  /* extern "C" */
namespace {
struct OutRow { double gl[3]; double sq; int32_t gt, gq, dp, ro, ao, qr, qa, rs, as_, asc, rp, ap; };
enum { FK_GT, FK_GQ, FK_SQ, FK_GL, FK_DP, FK_RO, FK_AO, FK_QR, FK_QA, FK_RS, FK_AS, FK_ASC, FK_RP, FK_AP, FK_AB, FK_COUNT };

/* one sample column; style 0 = every field, 1 = the blank row, 2 = "./." and '.' for every other field */
size_t format_call(const OutRow &r, const int32_t *order, int n_fields, int style, char *o)
{
    char *p = o;
    for (int k = 0; k < n_fields; ++k) {
        if (k) *p++ = ':';
        const int f = order[k];
        if (style == 2) { if (f == FK_GT) { memcpy(p, "./.", 3); p += 3; } else *p++ = '.'; continue; }
        if (style == 1) {
            if (f == FK_GT) { memcpy(p, "./.", 3); p += 3; }
            else if (f == FK_GQ || f == FK_SQ || f == FK_GL || f == FK_AB) *p++ = '.';
            else *p++ = '0';
            continue;
        }
        const bool called = r.gt >= 0;
        switch (f) {
        case FK_GT: memcpy(p, !called ? "./." : r.gt == 0 ? "0/0" : r.gt == 1 ? "0/1" : "1/1", 3); p += 3; break;
        case FK_GQ: if (called) p += sprintf(p, "%d", r.gq); else *p++ = '.'; break;
        case FK_SQ: if (called) p += sprintf(p, "%0.2f", r.sq); else *p++ = '.'; break;
        case FK_GL: p += sprintf(p, "%.0f,%.0f,%.0f", r.gl[0], r.gl[1], r.gl[2]); break;
        case FK_AB: {
            const int64_t tot = (int64_t)r.qr + r.qa;
            if (tot) p += sprintf(p, "%.2g", (double)r.qa / (double)tot); else *p++ = '.';
            break;
        }
        default: {
            const int32_t *ints = &r.gt;               /* gt gq dp ro ao qr qa rs as asc rp ap */
            static const int slot[FK_COUNT] = {0, 1, -1, -1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, -1};
            p += sprintf(p, "%d", ints[slot[f]]);
        }
        }
    }
    return (size_t)(p - o);
}
}  // namespace

extern "C" {

int svgt_format_calls(const void *rows, int64_t n, const int32_t *order, int32_t n_fields, const uint8_t *style,
                      int32_t n_threads, char *out, int64_t stride, int32_t *lengths)
{
    if (n < 0 || n_fields < 1 || n_fields > FK_COUNT || !order || stride < 16 * FK_COUNT + 64 ||
        (n > 0 && (!rows || !out || !lengths || !style)))
        return fail(SVGT_PACK_ERR_ARG, "bad argument");
    for (int k = 0; k < n_fields; ++k)
        if (order[k] < 0 || order[k] >= FK_COUNT) return fail(SVGT_PACK_ERR_ARG, "bad field id");
    const OutRow *r = (const OutRow *)rows;
    parallel_for(n, n_threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) lengths[i] = (int32_t)format_call(r[i], order, n_fields, style[i], out + i * stride);
    });
    return SVGT_PACK_OK;
}

/* SQ / GQ / GT from the (bit-exact) GL values with the HOST libm, exactly as CPython evaluates
 * singlesample.py:447-471: gt_sum = sum(10 ** gl), SQ = abs(-10 * (gl[0] - math.log(gt_sum, 10))); all three 10 ** gl
 * underflowing to 0.0 gives the "./." row.  The device computes the same with CUDA's pow / log, which are not
 * correctly rounded: this pass makes the printed SQ and QUAL independent of that. */
int svgt_host_sq(void *rows, int64_t n, int32_t n_threads)
{
    if (n < 0 || (n > 0 && !rows)) return fail(SVGT_PACK_ERR_ARG, "bad argument");
    OutRow *r = (OutRow *)rows;
    parallel_for(n, n_threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            OutRow &o = r[i];
            if (o.gt < -1) continue;                     /* blank / skipped rows carry no likelihoods */
            double gt_sum = 0.0;
            for (int g = 0; g < 3; ++g) gt_sum += pow(10.0, o.gl[g]);
            if (gt_sum > 0) {
                int best = 0;
                for (int g = 1; g < 3; ++g) if (o.gl[g] > o.gl[best]) best = g;
                int second = -1;
                for (int g = 0; g < 3; ++g) {
                    if (g == best) continue;
                    if (second < 0 || o.gl[g] > o.gl[second]) second = g;
                }
                const double gt_sum_log = log(gt_sum) / log(10.0);
                o.sq = fabs(-10 * (o.gl[0] - gt_sum_log));
                double phred = -10 * (o.gl[second] - o.gl[best]);
                if (phred > 200) phred = 200;
                o.gq = (int32_t)phred;
                o.gt = best;
            } else {
                o.gq = -1; o.sq = 0.0; o.gt = -1;
            }
        }
    });
    return SVGT_PACK_OK;
}

int svgt_format_quals(const double *qual, int64_t n, char *out, int64_t stride, int32_t *lengths)
{
    if (n < 0 || stride < 32 || (n > 0 && (!qual || !out || !lengths))) return fail(SVGT_PACK_ERR_ARG, "bad argument");
    for (int64_t i = 0; i < n; ++i) lengths[i] = (int32_t)snprintf(out + i * stride, (size_t)stride, "%0.2f", qual[i]);
    return SVGT_PACK_OK;
}

}  // extern "C"
