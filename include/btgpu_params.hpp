// btgpu_params.hpp — the two parameter files `bayesTyper genotype` writes next to the VCF:
//   <out>_genomic_parameters.txt   "Sample\tMean\tVariance", the fitted negative binomial of one haploid copy per sample
//                                  (CountDistribution::setGenomicCountDistributions, src/bayesTyper/CountDistribution.cpp:70-78,128-137;
//                                   NegativeBinomialDistribution::mean / var: size (1 - p) / p and mean / p)
//   <out>_noise_parameters.txt     "Chain\tIteration\t<samples>", one row per recorded noise-rate draw and the final "0\t0\t<means>"
//                                  (InferenceEngine.cpp:165-172,205,229,266; rows = what btg_estimate_noise returns in trace_out)
// Numbers go through operator<< of a default ostream as in the reference.  Host-side only.
#pragma once
#include <cstdint>
#include <ostream>
#include <string>
#include <vector>

namespace btg {

inline void writeGenomicParameters(std::ostream &os, const std::vector<std::string> &sample_names, const double *nb_p, const double *nb_size) {
    os << "Sample\tMean\tVariance" << std::endl;
    for (size_t s = 0; s < sample_names.size(); s++) {
        const double mean = nb_size[s] * (1 - nb_p[s]) / nb_p[s];
        os << sample_names[s] << "\t" << mean << "\t" << mean / nb_p[s] << std::endl;
    }
}

// trace: rows x (2 + S) doubles (chain, iteration, rates...)
inline void writeNoiseParameters(std::ostream &os, const std::vector<std::string> &sample_names, const double *trace, size_t rows) {
    const size_t S = sample_names.size();
    os << "Chain\tIteration";
    for (auto &n : sample_names) os << "\t" << n;
    os << std::endl;
    for (size_t r = 0; r < rows; r++) {
        const double *row = trace + r * (2 + S);
        os << (uint64_t)row[0] << "\t" << (uint64_t)row[1];
        for (size_t s = 0; s < S; s++) os << "\t" << row[2 + s];
        os << std::endl;
    }
}

}  // namespace btg
