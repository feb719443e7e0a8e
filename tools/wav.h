// Minimal RIFF/WAVE reader / writer for the mcbeam tool: PCM 16 / 24 / 32-bit and IEEE float 32 / 64-bit, interleaved,
// delivered as doubles normalised like libsndfile's sf_readf_double (integer PCM scaled by 1 / 2^(bits-1), floats as they
// are).  libsndfile, which the reference CLI links (src/programs/mcabeamf.cpp:186-190), is not part of this build.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace wav {

struct Info { int channels = 0, sample_rate = 0, bits = 0, format = 0; long long frames = 0; };   // format 1 = PCM, 3 = IEEE float

class Reader {
 public:
  explicit Reader(const std::string &path) {
    f_ = std::fopen(path.c_str(), "rb");
    if (!f_) throw std::runtime_error("cannot open " + path);
    char id[4]; uint32_t sz;
    if (std::fread(id, 1, 4, f_) != 4 || std::memcmp(id, "RIFF", 4) || std::fread(&sz, 4, 1, f_) != 1 || std::fread(id, 1, 4, f_) != 4 || std::memcmp(id, "WAVE", 4))
      throw std::runtime_error(path + ": not a RIFF/WAVE file");
    bool have_fmt = false;
    while (std::fread(id, 1, 4, f_) == 4 && std::fread(&sz, 4, 1, f_) == 1) {
      if (!std::memcmp(id, "fmt ", 4)) {
        unsigned char b[40] = {0};
        const size_t n = sz < sizeof(b) ? sz : sizeof(b);
        if (std::fread(b, 1, n, f_) != n) break;
        if (sz > n) std::fseek(f_, long(sz - n), SEEK_CUR);
        uint16_t fmt, ch, bits; uint32_t rate;
        std::memcpy(&fmt, b, 2); std::memcpy(&ch, b + 2, 2); std::memcpy(&rate, b + 4, 4); std::memcpy(&bits, b + 14, 2);
        if (fmt == 0xFFFE && n >= 26) std::memcpy(&fmt, b + 24, 2);   // WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the code
        info_.format = fmt; info_.channels = ch; info_.sample_rate = int(rate); info_.bits = bits;
        have_fmt = true;
      } else if (!std::memcmp(id, "data", 4)) {
        if (!have_fmt) break;
        info_.frames = (long long)sz / (info_.channels * (info_.bits / 8));
        ok_ = true;
        break;
      } else {
        std::fseek(f_, long(sz + (sz & 1)), SEEK_CUR);
      }
    }
    if (!ok_ || !(info_.format == 1 || info_.format == 3) || info_.channels < 1) throw std::runtime_error(path + ": unsupported WAVE layout");
  }
  ~Reader() { if (f_) std::fclose(f_); }
  const Info &info() const { return info_; }

  /** up to nframes interleaved frames as doubles; returns frames read */
  int readf_double(double *dst, int nframes) {
    const long long left = info_.frames - pos_;
    if (nframes > left) nframes = int(left);
    if (nframes <= 0) return 0;
    const int bps = info_.bits / 8, n = nframes * info_.channels;
    raw_.resize(size_t(n) * bps);
    const size_t got = std::fread(raw_.data(), size_t(bps) * info_.channels, size_t(nframes), f_);
    const int m = int(got) * info_.channels;
    const unsigned char *p = raw_.data();
    for (int i = 0; i < m; ++i, p += bps) {
      if (info_.format == 3) {
        if (bps == 4) { float v; std::memcpy(&v, p, 4); dst[i] = v; } else { double v; std::memcpy(&v, p, 8); dst[i] = v; }
      } else if (bps == 2) { int16_t v; std::memcpy(&v, p, 2); dst[i] = v / 32768.0; }
      else if (bps == 3) { int32_t v = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24); dst[i] = v / 2147483648.0; }
      else if (bps == 4) { int32_t v; std::memcpy(&v, p, 4); dst[i] = v / 2147483648.0; }
      else dst[i] = (int(p[0]) - 128) / 128.0;
    }
    pos_ += (long long)got;
    return int(got);
  }

 private:
  std::FILE *f_ = nullptr;
  Info info_;
  bool ok_ = false;
  long long pos_ = 0;
  std::vector<unsigned char> raw_;
};

class Writer {
 public:
  Writer(const std::string &path, Info info) : info_(info) {
    f_ = std::fopen(path.c_str(), "wb");
    if (!f_) throw std::runtime_error("cannot create " + path);
    header();
  }
  ~Writer() { if (f_) { header(); std::fclose(f_); } }

  void writef_double(const double *src, int nframes) {
    const int n = nframes * info_.channels, bps = info_.bits / 8;
    raw_.resize(size_t(n) * bps);
    unsigned char *p = raw_.data();
    for (int i = 0; i < n; ++i, p += bps) {
      if (info_.format == 3) {
        if (bps == 4) { float v = float(src[i]); std::memcpy(p, &v, 4); } else std::memcpy(p, &src[i], 8);
      } else {
        double s = src[i] * (bps == 2 ? 32768.0 : 2147483648.0);
        const double hi = bps == 2 ? 32767.0 : 2147483647.0, lo = bps == 2 ? -32768.0 : -2147483648.0;
        s = s > hi ? hi : (s < lo ? lo : s);
        const long long q = (long long)(s < 0 ? s - 0.5 : s + 0.5);
        if (bps == 2) { int16_t v = int16_t(q); std::memcpy(p, &v, 2); }
        else if (bps == 3) { int32_t v = int32_t(q); p[0] = (unsigned char)(v >> 8); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 24); }
        else { int32_t v = int32_t(q); std::memcpy(p, &v, 4); }
      }
    }
    std::fwrite(raw_.data(), 1, raw_.size(), f_);
    bytes_ += (long long)raw_.size();
  }

 private:
  void header() {
    std::fseek(f_, 0, SEEK_SET);
    const uint32_t data = uint32_t(bytes_), riff = 36 + data, rate = uint32_t(info_.sample_rate), fmtsz = 16;
    const uint16_t fmt = uint16_t(info_.format), ch = uint16_t(info_.channels), bits = uint16_t(info_.bits), align = uint16_t(ch * bits / 8);
    const uint32_t brate = rate * align;
    std::fwrite("RIFF", 1, 4, f_); std::fwrite(&riff, 4, 1, f_); std::fwrite("WAVEfmt ", 1, 8, f_); std::fwrite(&fmtsz, 4, 1, f_);
    std::fwrite(&fmt, 2, 1, f_); std::fwrite(&ch, 2, 1, f_); std::fwrite(&rate, 4, 1, f_); std::fwrite(&brate, 4, 1, f_);
    std::fwrite(&align, 2, 1, f_); std::fwrite(&bits, 2, 1, f_); std::fwrite("data", 1, 4, f_); std::fwrite(&data, 4, 1, f_);
    std::fseek(f_, 0, SEEK_END);
  }
  std::FILE *f_ = nullptr;
  Info info_;
  long long bytes_ = 0;
  std::vector<unsigned char> raw_;
};

}  // namespace wav
