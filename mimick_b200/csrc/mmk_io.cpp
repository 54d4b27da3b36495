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
    std::vector<std::string> f, keys, sub;
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
        split(line, '\t', f);
        if (f.size() < 10 || f.size() != 9 + v.samples.size()) {
            err = "malformed VCF record (column count)";
            return MMK_ERR_VCF;
        }
        VcfRecord r;
        r.chrom = f[0];
        r.pos = strtoll(f[1].c_str(), nullptr, 10);
        r.ref = f[3];
        split(f[4], ',', r.alts);
        // vartype, ProcessVariants.jl:61-69
        bool snp = r.ref.size() == 1;
        for (auto &a : r.alts) snp = snp && a.size() == 1;
        r.type = snp ? 0 : 1;
        split(f[8], ':', keys);
        int gi = -1;
        for (size_t i = 0; i < keys.size(); ++i)
            if (keys[i] == "GT") gi = (int)i;
        if (gi < 0) {
            err = "VCF record without GT";
            return MMK_ERR_VCF;
        }
        for (size_t s = 0; s < v.samples.size(); ++s) {
            split(f[9 + s], ':', sub);
            if ((size_t)gi >= sub.size()) {
                err = "VCF sample field without GT";
                return MMK_ERR_VCF;
            }
            const std::string &gt = sub[gi];
            // split(gt, ['/', '|']) then parse(UInt8, ·) — a missing allele "." is an error there too
            int n = 0;
            size_t b = 0;
            for (;;) {
                size_t e = gt.find_first_of("/|", b);
                std::string tok = gt.substr(b, e == std::string::npos ? std::string::npos : e - b);
                if (tok.empty() || tok.find_first_not_of("0123456789") != std::string::npos) {
                    err = "cannot parse genotype '" + gt + "' at " + r.chrom + ":" + f[1];
                    return MMK_ERR_VCF;
                }
                long a = strtol(tok.c_str(), nullptr, 10);
                if (a < 0 || a > (long)r.alts.size() || a > 255) {
                    err = "allele index out of range at " + r.chrom + ":" + f[1];
                    return MMK_ERR_VCF;
                }
                r.alleles.push_back((uint8_t)a);
                ++n;
                if (e == std::string::npos) break;
                b = e + 1;
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
