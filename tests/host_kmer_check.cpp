// Host-side compile of bayestyper_b200/csrc/kmer.cuh (plain g++) exposing its
// integer arithmetic to the CPU test-suite.  Test infrastructure only.
#include <cstddef>
#include <cstdint>
#include "../bayestyper_b200/csrc/kmer.cuh"

using namespace btg;

static uint64_t T[256];
static bool t_init = false;
static void init() {
    if (!t_init) {
        for (unsigned b = 0; b < 256; b++) T[b] = hash_table_entry(b);
        t_init = true;
    }
}

extern "C" {

uint64_t hk_hash(uint64_t w0, uint64_t w1) {
    init();
    return ntp64(from_boundary(w0, w1), T);
}
void hk_canonical(uint64_t w0, uint64_t w1, uint64_t *out) {
    Kmer128 f = from_boundary(w0, w1), r = revcomp(f);
    Kmer128 c = forward_is_canonical(f, r) ? f : r;
    to_boundary(c, out[0], out[1]);
}
void hk_roundtrip(uint64_t w0, uint64_t w1, uint64_t *out) {
    Kmer128 f = from_boundary(w0, w1);
    to_boundary(f, out[0], out[1]);
}
uint64_t hk_mod(uint64_t h, uint64_t m) { return mod_m(h, m, ~0ULL / m); }
int hk_contains(const uint8_t *bits, uint64_t m, uint32_t nh, uint64_t w0, uint64_t w1, unsigned *probes) {
    init();
    BloomView b{bits, m, ~0ULL / m, nh};
    return bloom_contains(b, ntp64(from_boundary(w0, w1), T), probes);
}
void hk_locs(uint64_t m, uint32_t nh, uint64_t w0, uint64_t w1, uint64_t *locs) {
    init();
    BloomView b{nullptr, m, ~0ULL / m, nh};
    uint64_t h = ntp64(from_boundary(w0, w1), T);
    for (unsigned i = 0; i < nh; i++) locs[i] = probe_loc(b, h, i);
}
unsigned hk_root(uint64_t w0, uint64_t w1) {
    init();
    return (unsigned)(ntp64_seeded(ntp64(from_boundary(w0, w1), T), kThreadedSeed) % kThreadedRoots);
}
// rolling scan: same contract as bto_scan_sequence, plus the rolled canonical hash
size_t hk_scan(const char *seq, size_t len, uint64_t *out, uint64_t *hash_out, uint32_t *pos_out, size_t cap) {
    Roller r;
    r.reset();
    size_t n = 0;
    for (size_t p = 0; p < len; p++) {
        unsigned c = nt_code(seq[p]);
        bool complete = false;
        if (c > 3) r.reset();
        else complete = r.push(c);
        if (complete) {
            if (n < cap) {
                Kmer128 cn = r.canonical();
                to_boundary(cn, out[2 * n], out[2 * n + 1]);
                hash_out[n] = r.canonical_hash();
                pos_out[n] = (uint32_t)p;
            }
            n++;
        }
    }
    return n;
}
}
