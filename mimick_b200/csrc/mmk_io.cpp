// mmk_io.cpp — host-side text readers: FASTA (FASTX.FASTAReader equivalent) and VCF
// (the VariantCallFormat fields the reference touches).  gzip, BGZF and plain text all go
// through zlib's gzFile, which reads concatenated gzip members (BGZF) and passes plain text
// through — the three cases of safe_read, mimick/MimickLinkedReads.jl/src/Common.jl:6-40.
#include <zlib.h>

#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "mmk_host.h"

namespace mmkio {

struct LineReader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    explicit LineReader(const char *path) : buf(1 << 20) {
        f = gzopen(path, "rb");
        if (f) gzbuffer(f, 1 << 20);
    }
    ~LineReader() {
        if (f) gzclose(f);
    }
    bool ok() const { return f != nullptr; }
    // next line without the terminator; false at end of file
    bool next(std::string &line) {
        line.clear();
        for (;;) {
            if (pos == end) {
                if (eof) return !line.empty();
                int n = gzread(f, buf.data(), (unsigned)buf.size());
                if (n <= 0) {
                    eof = true;
                    return !line.empty();
                }
                pos = 0;
                end = (size_t)n;
            }
            char *nl = (char *)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                line.append(buf.data() + pos, nl - (buf.data() + pos));
                pos = (size_t)(nl - buf.data()) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(buf.data() + pos, end - pos);
            pos = end;
        }
    }
};

int read_fasta(const char *path, std::vector<FastaRecord> &out, std::string &err) {
    FILE *probe = fopen(path, "rb");
    if (!probe) {
        err = std::string(path) + " does not exist.";
        return MMK_ERR_IO;
    }
    fclose(probe);
    LineReader lr(path);
    if (!lr.ok()) {
        err = std::string("cannot open ") + path;
        return MMK_ERR_IO;
    }
    std::string line;
    bool have = false;
    while (lr.next(line)) {
        if (!line.empty() && line[0] == '>') {
            out.emplace_back();
            have = true;
            size_t e = 1;
            while (e < line.size() && line[e] != ' ' && line[e] != '\t') ++e;  // identifier = up to first whitespace
            out.back().name = line.substr(1, e - 1);
        } else if (have) {
            size_t b = 0, e = line.size();
            while (b < e && (line[b] == ' ' || line[b] == '\t')) ++b;
            while (e > b && (line[e - 1] == ' ' || line[e - 1] == '\t')) --e;
            out.back().seq.append(line, b, e - b);
        }
    }
    return MMK_OK;
}

static void split(const std::string &s, char sep, std::vector<std::string> &out) {
    out.clear();
    size_t b = 0;
    for (;;) {
        size_t e = s.find(sep, b);
        if (e == std::string::npos) {
            out.emplace_back(s, b);
            return;
        }
        out.emplace_back(s, b, e - b);
        b = e + 1;
    }
}

int read_vcf(const char *path, VcfData &v, std::string &err) {
    FILE *probe = fopen(path, "rb");
    if (!probe) {
        err = std::string(path) + " does not exist.";
        return MMK_ERR_IO;
    }
    fclose(probe);
    LineReader lr(path);
    if (!lr.ok()) {
        err = std::string("cannot open ") + path;
        return MMK_ERR_IO;
    }
    std::string line;
    std::vector<std::string> f;
    std::vector<const char *> fb, fe;
    v.ploidy = 0;
    while (lr.next(line)) {
        if (line.empty()) continue;
        if (line[0] == '#') {
            if (line.compare(0, 10, "##contig=<") == 0) {
                // ##contig=<ID=chr1,length=...>
                size_t b = line.find("ID=", 10);
                if (b != std::string::npos) {
                    b += 3;
                    size_t e = line.find_first_of(",>", b);
                    v.contigs.push_back(line.substr(b, e - b));
                }
            } else if (line.compare(0, 6, "#CHROM") == 0) {
                split(line, '\t', f);
                for (size_t i = 9; i < f.size(); ++i) v.samples.push_back(f[i]);
            }
            continue;
        }
        // tab-separated fields as (begin, end) views into the line: no per-field allocations
        const char *L0 = line.data(), *Lend = L0 + line.size();
        fb.clear();
        fe.clear();
        for (const char *p = L0;;) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(Lend - p));
            fb.push_back(p);
            fe.push_back(t ? t : Lend);
            if (!t) break;
            p = t + 1;
        }
        if (fb.size() < 10 || fb.size() != 9 + v.samples.size()) {
            err = "malformed VCF record (column count)";
            return MMK_ERR_VCF;
        }
        VcfRecord r;
        r.chrom.assign(fb[0], fe[0]);
        const std::string pos_text(fb[1], fe[1]);
        r.pos = strtoll(pos_text.c_str(), nullptr, 10);
        r.ref.assign(fb[3], fe[3]);
        {
            const std::string alt_text(fb[4], fe[4]);
            split(alt_text, ',', r.alts);
        }
        // vartype, ProcessVariants.jl:61-69
        bool snp = r.ref.size() == 1;
        for (auto &a : r.alts) snp = snp && a.size() == 1;
        r.type = snp ? 0 : 1;
        int gi = -1;
        {   // index of GT among the ':'-separated FORMAT keys
            int k = 0;
            for (const char *p = fb[8];; ++k) {
                const char *t = (const char *)memchr(p, ':', (size_t)(fe[8] - p));
                const char *e = t ? t : fe[8];
                if (e - p == 2 && p[0] == 'G' && p[1] == 'T') gi = k;
                if (!t) break;
                p = t + 1;
            }
        }
        if (gi < 0) {
            err = "VCF record without GT";
            return MMK_ERR_VCF;
        }
        r.alleles.reserve(v.samples.size() * (v.ploidy > 0 ? (size_t)v.ploidy : 2u));
        r.ploidy_of.reserve(v.samples.size());
        for (size_t s = 0; s < v.samples.size(); ++s) {
            // the gi-th ':'-separated subfield of the sample column
            const char *p = fb[9 + s], *fend = fe[9 + s];
            int k = 0;
            while (k < gi) {
                const char *t = (const char *)memchr(p, ':', (size_t)(fend - p));
                if (!t) break;
                p = t + 1;
                ++k;
            }
            if (k < gi) {
                err = "VCF sample field without GT";
                return MMK_ERR_VCF;
            }
            const char *gend = (const char *)memchr(p, ':', (size_t)(fend - p));
            if (!gend) gend = fend;
            // split(gt, ['/', '|']) then parse(UInt8, ·) — a missing allele "." is an error there too
            int n = 0;
            for (const char *q = p;;) {
                const char *e = q;
                long a = 0;
                while (e < gend && *e >= '0' && *e <= '9') {
                    a = a < 100000 ? a * 10 + (*e - '0') : a;
                    ++e;
                }
                if (e == q || (e < gend && *e != '/' && *e != '|')) {
                    err = "cannot parse genotype '" + std::string(p, gend) + "' at " + r.chrom + ":" + pos_text;
                    return MMK_ERR_VCF;
                }
                if (a < 0 || a > (long)r.alts.size() || a > 255) {
                    err = "allele index out of range at " + r.chrom + ":" + pos_text;
                    return MMK_ERR_VCF;
                }
                r.alleles.push_back((uint8_t)a);
                ++n;
                if (e == gend) break;
                q = e + 1;
                if (q == gend) {  // trailing separator: an empty allele
                    err = "cannot parse genotype '" + std::string(p, gend) + "' at " + r.chrom + ":" + pos_text;
                    return MMK_ERR_VCF;
                }
            }
            if (v.ploidy == 0) v.ploidy = n;  // get_ploidy_and_contigs: sample 1 of the first record
            if (s == 0 && v.records.empty()) v.ploidy = n;
            r.ploidy_of.push_back((uint8_t)n);
        }
        v.records.push_back(std::move(r));
    }
    return MMK_OK;
}

}  // namespace mmkio
