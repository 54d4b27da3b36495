// mmk_host.h — host-side declarations shared by mmk_host.cpp and mmk_io.cpp.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "mimick_b200.h"

namespace mmkio {

struct FastaRecord {
    std::string name;
    std::string seq;
};
int read_fasta(const char *path, std::vector<FastaRecord> &out, std::string &err);

struct VcfRecord {
    std::string chrom;
    int64_t pos = 0;
    std::string ref;
    std::vector<std::string> alts;
    int type = 0;                     // 0 :snp, 1 :indel
    std::vector<uint8_t> alleles;     // all samples' GT alleles, concatenated
    std::vector<uint8_t> ploidy_of;   // alleles per sample in this record
};
struct VcfData {
    std::vector<std::string> contigs;  // ##contig IDs, header order
    std::vector<std::string> samples;
    int ploidy = 0;
    std::vector<VcfRecord> records;
};
int read_vcf(const char *path, VcfData &out, std::string &err);

}  // namespace mmkio

struct mmk_vcf {
    mmkio::VcfData d;
};
