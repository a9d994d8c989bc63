// btkmc — KMC databases on the host side of the k-mer path.
//
//   btkmc info <kmc_prefix>                 parameters of the database (CKMCFile::Info)
//   btkmc list <kmc_prefix>                 every (k-mer, count) in listing order, as text "<55-mer>\t<count>" (CKMCFile::ReadNextKmer)
//   btkmc makebloom <kmc_prefix> [fpr]      `bayesTyperTools makeBloom` (src/bayesTyperTools/MakeBloom.cpp:200-295): KmerBloom(total_kmers, fpr),
//                                           every k-mer of the database inserted in batches of 1,000,000 on the GPU, saved as
//                                           <kmc_prefix>.bloomMeta / .bloomData (byte-identical to the reference's filter: insertion order
//                                           does not matter).  Needs a B200 and libbtgpu.so; info / list are host-only.
#include <iostream>

#include "btgpu.hpp"
#include "btgpu_kmc.hpp"

int main(int argc, char **argv) {
    if (argc < 3) { std::cerr << "usage: btkmc info|list|makebloom <kmc_prefix> [false_positive_rate]\n"; return 2; }
    const std::string cmd = argv[1], prefix = argv[2];
    try {
        btg::KmcReader db(prefix);
        if (cmd == "info") {
            std::cout << "kmer_length " << db.kmer_length << "\nmode " << db.mode << "\ncounter_size " << db.counter_size << "\nlut_prefix_length " << db.lut_prefix_length
                      << "\nsignature_len " << db.signature_len << "\nmin_count " << db.min_count << "\nmax_count " << db.max_count << "\ntotal_kmers " << db.total_kmers
                      << "\nboth_strands " << db.both_strands << "\nkmc_version " << db.kmc_version << "\n";
            return 0;
        }
        std::vector<uint64_t> kmers;
        std::vector<uint32_t> counts;
        if (cmd == "list") {
            std::string line(db.kmer_length, 'A');
            while (true) {
                kmers.clear(); counts.clear();
                const size_t n = db.read(kmers, counts, 1 << 16);
                if (!n) break;
                for (size_t i = 0; i < n; i++) {
                    for (uint32_t j = 0; j < db.kmer_length; j++) line[j] = "ACGT"[(kmers[2 * i + (j >> 5)] >> (2 * (j & 31u))) & 3u];
                    std::cout << line << '\t' << counts[i] << '\n';
                }
            }
            return 0;
        }
        if (cmd == "makebloom") {
            if (db.kmer_length != BTG_KMER_SIZE) throw btg::Error("the database's k-mer length is not the library's (55)");
            const float fpr = argc > 3 ? std::stof(argv[3]) : 0.001f;      // src/bayesTyperTools/main.cpp:127
            btg::Library lib(0);
            btg_bloom *b = btg::check_ptr(btg_bloom_create(db.total_kmers, fpr, BTG_KMER_SIZE));
            uint64_t parsed = 0;
            while (true) {
                kmers.clear(); counts.clear();
                const size_t n = db.read(kmers, counts, 1000000);
                if (!n) break;
                btg::check(btg_bloom_insert(b, kmers.data(), n));
                parsed += n;
            }
            btg::check(btg_bloom_save(b, prefix.c_str()));
            btg_bloom_free(b);
            std::cout << "btkmc: bloom filter of " << parsed << " k-mers (false positive rate " << fpr << ") written to " << prefix << ".bloomMeta/.bloomData" << std::endl;
            return 0;
        }
        throw btg::Error("unknown command " + cmd);
    } catch (const std::exception &e) {
        std::cerr << "\nERROR: " << e.what() << "\n" << std::endl;
        return 1;
    }
}
