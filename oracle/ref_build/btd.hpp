// btd.hpp — "BTD1" named-array container used for oracle-R dumps and test fixtures.
// Layout: magic "BTD1", then records: u32 name_len, name, u8 dtype, u8 ndim, u64 dims[ndim], raw data.
// dtype: 0=u8 1=u16 2=u32 3=u64 4=i32 5=f32 6=f64 7=i64.  Read by bayestyper_b200/btd.py.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

namespace btd {

template <class T> struct dtype_of;
template <> struct dtype_of<uint8_t> { static const uint8_t v = 0; };
template <> struct dtype_of<uint16_t> { static const uint8_t v = 1; };
template <> struct dtype_of<uint32_t> { static const uint8_t v = 2; };
template <> struct dtype_of<uint64_t> { static const uint8_t v = 3; };
template <> struct dtype_of<int32_t> { static const uint8_t v = 4; };
template <> struct dtype_of<float> { static const uint8_t v = 5; };
template <> struct dtype_of<double> { static const uint8_t v = 6; };
template <> struct dtype_of<int64_t> { static const uint8_t v = 7; };

class Writer {
   public:
    explicit Writer(const std::string &path) : f_(fopen(path.c_str(), "wb")) {
        if (f_) fwrite("BTD1", 1, 4, f_);
    }
    ~Writer() { if (f_) fclose(f_); }
    bool ok() const { return f_ != nullptr; }
    template <class T> void put(const std::string &name, const T *data, const std::vector<uint64_t> &dims) {
        uint32_t nl = (uint32_t)name.size();
        fwrite(&nl, 4, 1, f_);
        fwrite(name.data(), 1, nl, f_);
        uint8_t dt = dtype_of<T>::v, nd = (uint8_t)dims.size();
        fwrite(&dt, 1, 1, f_);
        fwrite(&nd, 1, 1, f_);
        uint64_t n = 1;
        for (uint64_t d : dims) { fwrite(&d, 8, 1, f_); n *= d; }
        if (n) fwrite(data, sizeof(T), n, f_);
    }
    template <class T> void put(const std::string &name, const std::vector<T> &v) { put(name, v.data(), {(uint64_t)v.size()}); }
    void put_str(const std::string &name, const std::string &s) { put(name, (const uint8_t *)s.data(), {(uint64_t)s.size()}); }
   private:
    FILE *f_;
};

}  // namespace btd
