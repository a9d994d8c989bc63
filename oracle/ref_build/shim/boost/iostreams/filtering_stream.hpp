// Stand-in for boost::iostreams::filtering_[io]stream as the reference uses them:
//   in : push(gzip_decompressor()) [optional]; push(boost::ref(ifstream))  -> whole file is
//        read at push time and inflated with zlib when it carries the gzip magic
//   out: push(gzip_compressor()) [optional]; push(boost::ref(ofstream))    -> written through
//        UNCOMPRESSED (oracle-R outputs are read back by this same shim or by tests)
#pragma once
#include <zlib.h>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
using std::ref;
namespace iostreams {
struct gzip_decompressor {};
struct gzip_compressor {};

class filtering_istream : public std::istream {
   public:
    filtering_istream() : std::istream(&buf_), n_(0), complete_(false) {}
    void push(const gzip_decompressor &) { n_++; }
    void push(std::reference_wrapper<std::ifstream> f) {
        std::stringstream ss;
        ss << f.get().rdbuf();
        std::string raw = ss.str();
        if (raw.size() > 2 && (unsigned char)raw[0] == 0x1f && (unsigned char)raw[1] == 0x8b) raw = inflate_all(raw);
        buf_.str(raw);
        clear();
        n_++;
        complete_ = true;
        dev_ = &f.get();
    }
    bool is_complete() const { return complete_; }
    bool empty() const { return n_ == 0; }
    // real Boost pops the chain with auto-close: the underlying fstream is closed
    // (the reference relies on it: VariantFileParser.cpp:101-124)
    void reset() { buf_.str(""); clear(); n_ = 0; complete_ = false; if (dev_) { dev_->close(); dev_->clear(); dev_ = nullptr; } }
   private:
    static std::string inflate_all(const std::string &in) {
        z_stream zs{};
        if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) throw std::runtime_error("inflateInit2");
        zs.next_in = (Bytef *)in.data();
        zs.avail_in = in.size();
        std::string out;
        char tmp[1 << 16];
        int rc;
        do {
            zs.next_out = (Bytef *)tmp;
            zs.avail_out = sizeof(tmp);
            rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); throw std::runtime_error("inflate"); }
            out.append(tmp, sizeof(tmp) - zs.avail_out);
            if (rc == Z_STREAM_END && zs.avail_in > 0) { inflateReset(&zs); rc = Z_OK; }  // concatenated members
        } while (rc != Z_STREAM_END && (zs.avail_in > 0 || zs.avail_out == 0));
        inflateEnd(&zs);
        return out;
    }
    std::stringbuf buf_;
    int n_;
    bool complete_;
    std::ifstream *dev_ = nullptr;
};

class filtering_ostream : public std::ostream {
   public:
    filtering_ostream() : std::ostream(nullptr), n_(0), complete_(false) {}
    void push(const gzip_compressor &) { n_++; }
    void push(std::reference_wrapper<std::ofstream> f) { rdbuf(f.get().rdbuf()); clear(); n_++; complete_ = true; dev_ = &f.get(); }
    bool is_complete() const { return complete_; }
    bool empty() const { return n_ == 0; }
    void reset() { flush(); rdbuf(nullptr); n_ = 0; complete_ = false; if (dev_) { dev_->close(); dev_->clear(); dev_ = nullptr; } }
    ~filtering_ostream() { if (dev_) flush(); }
   private:
    int n_;
    bool complete_;
    std::ofstream *dev_ = nullptr;
};
}  // namespace iostreams
}  // namespace boost
