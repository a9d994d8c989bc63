// btd.hpp — BTD1 named-array container (bayestyper_b200/btd.py) for the C++ host programs.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <string>
#include <vector>

#include "btgpu.hpp"

namespace btd {

struct Array {
    uint8_t dtype = 0;  // 0 u8, 1 u16, 2 u32, 3 u64, 4 i32, 5 f32, 6 f64, 7 i64
    std::vector<uint64_t> dims;
    std::vector<uint8_t> bytes;
    template <class T> const T *as() const { return reinterpret_cast<const T *>(bytes.data()); }
    uint64_t count() const { uint64_t n = 1; for (auto d : dims) n *= d; return n; }
};
static const size_t kItem[8] = {1, 2, 4, 8, 4, 4, 8, 8};

inline std::map<std::string, Array> read_btd(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw btg::Error("cannot open " + path);
    std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.size() < 4 || memcmp(buf.data(), "BTD1", 4) != 0) throw btg::Error(path + ": not a BTD1 file");
    std::map<std::string, Array> out;
    size_t off = 4;
    auto need = [&](size_t n) { if (off + n > buf.size()) throw btg::Error(path + ": truncated"); };
    while (off < buf.size()) {
        need(4); uint32_t nl; memcpy(&nl, &buf[off], 4); off += 4;
        need(nl); std::string name(&buf[off], nl); off += nl;
        need(2); Array a; a.dtype = (uint8_t)buf[off]; const uint8_t nd = (uint8_t)buf[off + 1]; off += 2;
        if (a.dtype > 7) throw btg::Error(path + ": bad dtype");
        need(8ull * nd); a.dims.resize(nd); memcpy(a.dims.data(), &buf[off], 8ull * nd); off += 8ull * nd;
        const size_t nb = a.count() * kItem[a.dtype];
        need(nb); a.bytes.assign(buf.begin() + off, buf.begin() + off + nb); off += nb;
        out.emplace(name, std::move(a));
    }
    return out;
}

struct BtdWriter {
    std::ofstream f;
    explicit BtdWriter(const std::string &path) : f(path, std::ios::binary) { if (!f) throw btg::Error("cannot write " + path); f.write("BTD1", 4); }
    template <class T> void put(const std::string &name, uint8_t dtype, const T *data, std::vector<uint64_t> dims) {
        const uint32_t nl = (uint32_t)name.size();
        f.write((const char *)&nl, 4); f.write(name.data(), nl);
        const uint8_t hd[2] = {dtype, (uint8_t)dims.size()};
        f.write((const char *)hd, 2); f.write((const char *)dims.data(), 8 * dims.size());
        uint64_t n = 1; for (auto d : dims) n *= d;
        f.write((const char *)data, n * sizeof(T));
    }
    template <class T> void put(const std::string &name, uint8_t dtype, const std::vector<T> &v) { put(name, dtype, v.data(), {(uint64_t)v.size()}); }
};


}  // namespace btd
