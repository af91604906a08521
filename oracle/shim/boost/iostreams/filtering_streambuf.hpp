// Minimal stand-in for the Boost.Iostreams gzip pipeline used by utils/utils.h:35-91
// (OBS tag compression), backed by zlib. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <zlib.h>

namespace boost {
namespace iostreams {

struct output {};
struct input {};
struct gzip_compressor {};
struct gzip_decompressor {};

template <class Ch>
struct basic_array_source {
    const Ch* data;
    size_t size;
    basic_array_source(const Ch* d, size_t n) : data(d), size(n) {}
};

template <class Mode>
class filtering_streambuf {
public:
    void push(const gzip_compressor&) { compress_ = true; }
    void push(const gzip_decompressor&) { decompress_ = true; }
    void push(std::stringstream& s) { sink_ = &s; }
    void push(const basic_array_source<char>& src) { src_ = std::string(src.data, src.size); }
    bool compress_ = false, decompress_ = false;
    std::stringstream* sink_ = nullptr;
    std::string src_;
};

inline std::string zdeflate_gzip(const std::string& in)
{
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { throw std::runtime_error("deflateInit2"); }
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
    zs.avail_in = in.size();
    std::string out;
    char buf[32768];
    int ret;
    do {
        zs.next_out = reinterpret_cast<Bytef*>(buf);
        zs.avail_out = sizeof(buf);
        ret = deflate(&zs, Z_FINISH);
        out.append(buf, sizeof(buf) - zs.avail_out);
    } while (ret == Z_OK);
    deflateEnd(&zs);
    return out;
}

inline std::string zinflate_gzip(const std::string& in)
{
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 16) != Z_OK) { throw std::runtime_error("inflateInit2"); }
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
    zs.avail_in = in.size();
    std::string out;
    char buf[32768];
    int ret;
    do {
        zs.next_out = reinterpret_cast<Bytef*>(buf);
        zs.avail_out = sizeof(buf);
        ret = inflate(&zs, Z_NO_FLUSH);
        out.append(buf, sizeof(buf) - zs.avail_out);
    } while (ret == Z_OK);
    inflateEnd(&zs);
    return out;
}

// copy(array_source, out-chain): compress into the sink pushed on the chain
inline void copy(const basic_array_source<char>& src, filtering_streambuf<output>& out)
{
    std::string data(src.data, src.size);
    (*out.sink_) << (out.compress_ ? zdeflate_gzip(data) : data);
}

// copy(in-chain, stringstream): decompress the source pushed on the chain
inline void copy(filtering_streambuf<input>& in, std::stringstream& dst)
{
    dst << (in.decompress_ ? zinflate_gzip(in.src_) : in.src_);
}

template <class T>
inline void close(T&) {}

} // namespace iostreams
} // namespace boost
