// btgpu_kmc.hpp — sequential reader of KMC databases (<prefix>.kmc_pre / .kmc_suf) for the host programs.
//
// Restates what the reference does through the vendored KMC API 2.3.0, CKMCFile::OpenForListing / ReadNextKmer
// (external/kmc_api/kmc_file.cpp:66-99,177-292,428-515), for the two database layouts that API accepts:
//   KMC1 (kmc_version 0)     .kmc_pre = "KMCP" | LUT[4^p] u64 | header 5 x u64 | header_offset u32 | "KMCP"
//   KMC2 (kmc_version 0x200) .kmc_pre = "KMCP" | LUT[bins][4^p] u64 (+ 1) | signature map u32[4^s + 1] | header | version u32 | header_offset u32 | "KMCP"
//   .kmc_suf = "KMCS" | records ( (k - p)/4 suffix bytes, 4 nt per byte MSB first, A0 C1 G2 T3 | counter, counter_size bytes LE ) | "KMCS"
// LUT[i] is the index of the first record whose k-mer starts with prefix i (KMC2: of bin i / 4^p, prefix i % 4^p); records of one
// prefix are sorted, so a KMC1 database lists its k-mers in lexicographic order and a KMC2 database bin by bin, each bin sorted.
// K-mers come out packed in the C ABI's layout (2 x uint64, nucleotide i at bits [2i, 2i+1]), ready for btg_bloom_insert /
// btg_table_add_sample_kmers.  Host-side only.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace btg {

class KmcReader {
  public:
    uint32_t kmer_length = 0, mode = 0, counter_size = 0, lut_prefix_length = 0, signature_len = 0, min_count = 0, kmc_version = 0;
    uint64_t max_count = 0, total_kmers = 0;
    bool both_strands = true;

    explicit KmcReader(const std::string &prefix, size_t buffer_bytes = 1 << 24) : buf_(buffer_bytes) {
        std::vector<uint8_t> pre = slurp(prefix + ".kmc_pre", "KMCP");
        parse_prefix_file(pre);
        suf_ = std::fopen((prefix + ".kmc_suf").c_str(), "rb");
        if (!suf_) throw std::runtime_error("cannot open " + prefix + ".kmc_suf");
        char m[4];
        if (std::fread(m, 1, 4, suf_) != 4 || std::memcmp(m, "KMCS", 4) != 0) throw std::runtime_error(prefix + ".kmc_suf: bad marker");
        suffix_bytes_ = (kmer_length - lut_prefix_length) / 4;
        if ((kmer_length - lut_prefix_length) % 4 != 0 || counter_size == 0 || counter_size > 8) throw std::runtime_error(prefix + ": unsupported record layout");
        record_bytes_ = suffix_bytes_ + counter_size;
        if (kmer_length > 64) throw std::runtime_error("k-mers longer than 64 nucleotides do not fit the two-word ABI layout");
    }
    ~KmcReader() { if (suf_) std::fclose(suf_); }
    KmcReader(const KmcReader &) = delete;
    KmcReader &operator=(const KmcReader &) = delete;

    // Appends up to max_records (k-mer, count) pairs that pass the database's [min_count, max_count] window (ReadNextKmer's filter);
    // kmers gets 2 words per k-mer.  Returns the number appended; 0 at the end of the database.
    size_t read(std::vector<uint64_t> &kmers, std::vector<uint32_t> &counts, size_t max_records) {
        size_t n = 0;
        const uint64_t prefix_mask = (1ull << (2 * lut_prefix_length)) - 1;
        while (n < max_records && record_ < total_kmers) {
            while (record_ == lut_[prefix_index_ + 1]) prefix_index_++;      // next non-empty prefix (kmc_file.cpp:447-453)
            const uint8_t *rec = next_record();
            uint64_t w[2] = {0, 0};
            uint32_t nt = 0;
            const uint64_t p = prefix_index_ & prefix_mask;
            for (uint32_t i = 0; i < lut_prefix_length; i++, nt++) put(w, nt, (p >> (2 * (lut_prefix_length - 1 - i))) & 3u);
            for (uint32_t b = 0; b < suffix_bytes_; b++)
                for (int sft = 6; sft >= 0; sft -= 2, nt++) put(w, nt, (rec[b] >> sft) & 3u);
            uint64_t c = 0;
            for (uint32_t b = 0; b < counter_size; b++) c |= (uint64_t)rec[suffix_bytes_ + b] << (8 * b);
            record_++;
            if (c < min_count || c > max_count) continue;
            kmers.push_back(w[0]); kmers.push_back(w[1]);
            counts.push_back((uint32_t)c);
            n++;
        }
        return n;
    }

  private:
    std::vector<uint64_t> lut_;      // record index at which every (bin,) prefix starts, + sentinel total_kmers + 1
    std::vector<uint8_t> buf_;
    size_t buf_len_ = 0, buf_pos_ = 0;
    FILE *suf_ = nullptr;
    uint32_t suffix_bytes_ = 0, record_bytes_ = 0;
    uint64_t record_ = 0, prefix_index_ = 0;
    uint8_t carry_[64];

    static void put(uint64_t w[2], uint32_t nt, uint64_t code) { w[nt >> 5] |= code << (2 * (nt & 31u)); }

    static std::vector<uint8_t> slurp(const std::string &path, const char *marker) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path);
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::rewind(f);
        std::vector<uint8_t> d((size_t)std::max<long>(n, 0));
        const size_t got = d.empty() ? 0 : std::fread(d.data(), 1, d.size(), f);
        std::fclose(f);
        if (got != d.size() || d.size() < 8 || std::memcmp(d.data(), marker, 4) != 0 || std::memcmp(d.data() + d.size() - 4, marker, 4) != 0)
            throw std::runtime_error(path + ": not a KMC file (marker)");
        return d;
    }
    template <class T> static T rd(const std::vector<uint8_t> &d, size_t off) {
        if (off + sizeof(T) > d.size()) throw std::runtime_error("KMC prefix file truncated");
        T v;
        std::memcpy(&v, d.data() + off, sizeof(T));
        return v;
    }

    void parse_prefix_file(const std::vector<uint8_t> &d) {
        const size_t end = d.size() - 4;                       // before the terminal marker
        kmc_version = rd<uint32_t>(d, end - 8);
        const uint32_t header_offset = d[end - 4];             // the API reads one byte of it (fgetc)
        if (kmc_version == 0x200) {
            const size_t h = end - 4 - header_offset;          // the header (its last word is the version) ends at the offset field (kmc_file.cpp:189-193: seek to -(header_offset + 8) from the end)
            kmer_length = rd<uint32_t>(d, h); mode = rd<uint32_t>(d, h + 4); counter_size = rd<uint32_t>(d, h + 8); lut_prefix_length = rd<uint32_t>(d, h + 12);
            signature_len = rd<uint32_t>(d, h + 16); min_count = rd<uint32_t>(d, h + 20); max_count = rd<uint32_t>(d, h + 24); total_kmers = rd<uint64_t>(d, h + 28);
            both_strands = !rd<uint8_t>(d, h + 36);
            const size_t map_bytes = ((1ull << (2 * signature_len)) + 1) * 4;
            if (h < 4 + 8 + map_bytes) throw std::runtime_error("KMC2 prefix file truncated");
            const size_t lut_bytes = h - map_bytes - 4 - 8;    // LUT area: after the initial marker, before one extra word and the signature map (kmc_file.cpp:206-214)
            lut_.resize(lut_bytes / 8 + 1);
            std::memcpy(lut_.data(), d.data() + 4, lut_bytes / 8 * 8);
            lut_[lut_bytes / 8] = total_kmers + 1;
        } else if (kmc_version == 0) {
            const size_t h = end - 4 - header_offset;          // KMC1: header_offset bytes of header right before the offset field
            const uint64_t d0 = rd<uint64_t>(d, h), d1 = rd<uint64_t>(d, h + 8), d2 = rd<uint64_t>(d, h + 16), d3 = rd<uint64_t>(d, h + 24), d4 = rd<uint64_t>(d, h + 32);
            kmer_length = (uint32_t)d0; mode = (uint32_t)(d0 >> 32); counter_size = (uint32_t)d1; lut_prefix_length = (uint32_t)(d1 >> 32);
            min_count = (uint32_t)d2; max_count = (d2 >> 32) + (d4 & 0xFFFFFFFF00000000ull); total_kmers = d3; both_strands = !((d4 & 0xF) == 1);
            const size_t n_lut = (h - 4) / 8;
            lut_.resize(n_lut + 1);
            std::memcpy(lut_.data(), d.data() + 4, n_lut * 8);
            lut_[n_lut] = total_kmers + 1;
        } else {
            throw std::runtime_error("unsupported KMC database version");
        }
        if (mode != 0) throw std::runtime_error("KMC databases with quality-aware counters (mode 1) are not supported (the reference asserts mode 0, KmerCounter.cpp:448-449)");
    }

    const uint8_t *next_record() {
        if (buf_len_ - buf_pos_ < record_bytes_) {             // refill, keeping a partial record
            const size_t rest = buf_len_ - buf_pos_;
            std::memcpy(carry_, buf_.data() + buf_pos_, rest);
            std::memcpy(buf_.data(), carry_, rest);
            const size_t got = std::fread(buf_.data() + rest, 1, buf_.size() - rest, suf_);
            buf_len_ = rest + got; buf_pos_ = 0;
            if (buf_len_ < record_bytes_) throw std::runtime_error("KMC suffix file truncated");
        }
        const uint8_t *r = buf_.data() + buf_pos_;
        buf_pos_ += record_bytes_;
        return r;
    }
};

}  // namespace btg
