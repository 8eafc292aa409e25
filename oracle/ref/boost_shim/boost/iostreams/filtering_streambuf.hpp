// TEST INFRASTRUCTURE ONLY (oracle/).  The reference's read loader (src/ReadData.cpp:8-10,95-104,
// 165-174) reads gzip input through boost::iostreams:
//     filtering_streambuf<input> *inbuf = new ...;  inbuf->push(gzip_decompressor());
//     inbuf->push(infile);  fin = new std::istream(inbuf);
// Boost is a network download of the reference's build and is not in this image.  This stand-in
// gives exactly those names on top of zlib (which IS here): a std::streambuf that inflates the
// pushed std::istream (gzip or zlib framing, concatenated members included), so the UNMODIFIED
// src/ReadData.cpp compiles where it lies (oracle/Makefile, target readdata).
#pragma once
#include <zlib.h>

#include <cstring>
#include <istream>
#include <stdexcept>
#include <streambuf>
#include <vector>

namespace boost {
namespace iostreams {

struct input {};
struct gzip_decompressor {};

template <typename Mode>
class filtering_streambuf : public std::streambuf {
public:
    filtering_streambuf() : in_(1 << 16), out_(1 << 16) { std::memset(&zs_, 0, sizeof zs_); }
    ~filtering_streambuf() override {
        if (gz_) inflateEnd(&zs_);
    }
    void push(const gzip_decompressor &) {
        if (inflateInit2(&zs_, 15 + 32) != Z_OK) throw std::runtime_error("boost shim: inflateInit2 failed");
        gz_ = true;
    }
    void push(std::istream &src) { src_ = &src; }

protected:
    int_type underflow() override {
        if (gptr() < egptr()) return traits_type::to_int_type(*gptr());
        if (!src_ || done_) return traits_type::eof();
        if (!gz_) {
            src_->read(out_.data(), (std::streamsize)out_.size());
            std::streamsize got = src_->gcount();
            if (got <= 0) { done_ = true; return traits_type::eof(); }
            setg(out_.data(), out_.data(), out_.data() + got);
            return traits_type::to_int_type(*gptr());
        }
        for (;;) {
            if (zs_.avail_in == 0) {
                src_->read(in_.data(), (std::streamsize)in_.size());
                std::streamsize got = src_->gcount();
                if (got <= 0) { done_ = true; return traits_type::eof(); }
                zs_.next_in = reinterpret_cast<Bytef *>(in_.data());
                zs_.avail_in = (uInt)got;
            }
            zs_.next_out = reinterpret_cast<Bytef *>(out_.data());
            zs_.avail_out = (uInt)out_.size();
            int rc = inflate(&zs_, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR)
                throw std::runtime_error("boost shim: inflate failed");
            size_t produced = out_.size() - zs_.avail_out;
            if (rc == Z_STREAM_END) inflateReset(&zs_);   // next gzip member, if any
            if (produced) {
                setg(out_.data(), out_.data(), out_.data() + produced);
                return traits_type::to_int_type(*gptr());
            }
        }
    }

private:
    std::istream *src_ = nullptr;
    bool gz_ = false, done_ = false;
    z_stream zs_;
    std::vector<char> in_, out_;
};

}  // namespace iostreams
}  // namespace boost
