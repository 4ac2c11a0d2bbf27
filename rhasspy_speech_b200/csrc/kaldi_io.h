// Reader for Kaldi's on-disk object streams (binary "\0B" and text mode).
//
// Follows the format written/read by the reference's kaldi/src/base/io-funcs{.cc,-inl.h}
// (tokens end with one space in both modes; binary basic types carry a size byte),
// kaldi/src/matrix/kaldi-matrix.cc / kaldi-vector.cc / packed-matrix.cc (FM/DM/FV/DV/FP/DP
// headers, text "[ ... ]").  Host-only C++; no Kaldi code is linked.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace rs {

struct Error : std::runtime_error {
  explicit Error(const std::string &m) : std::runtime_error(m) {}
};

#define RS_FAIL(msg)                                   \
  do {                                                 \
    std::ostringstream rs_oss_;                        \
    rs_oss_ << msg;                                    \
    throw ::rs::Error(rs_oss_.str());                  \
  } while (0)

struct MatrixF {
  int rows = 0, cols = 0;
  std::vector<float> d;  // row-major, stride == cols
  float &operator()(int r, int c) { return d[(size_t)r * cols + c]; }
  float operator()(int r, int c) const { return d[(size_t)r * cols + c]; }
};
struct MatrixD {
  int rows = 0, cols = 0;
  std::vector<double> d;
  double &operator()(int r, int c) { return d[(size_t)r * cols + c]; }
  double operator()(int r, int c) const { return d[(size_t)r * cols + c]; }
};

class KaldiReader {
 public:
  explicit KaldiReader(const std::string &path) : path_(path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) RS_FAIL("cannot open " << path);
    std::stringstream ss;
    ss << f.rdbuf();
    buf_ = ss.str();
    if (buf_.size() >= 2 && buf_[0] == '\0' && buf_[1] == 'B') {
      binary_ = true;
      pos_ = 2;
    }
  }
  bool binary() const { return binary_; }
  bool eof() const { return pos_ >= buf_.size(); }
  const std::string &path() const { return path_; }

  void SkipWs() {
    while (pos_ < buf_.size() && isspace((unsigned char)buf_[pos_])) pos_++;
  }
  int Peek() {
    if (!binary_) SkipWs();
    return pos_ < buf_.size() ? (unsigned char)buf_[pos_] : -1;
  }
  // first character of the next token, ignoring a leading '<' (Kaldi's PeekToken)
  int PeekToken() {
    if (!binary_) SkipWs();
    size_t p = pos_;
    if (p < buf_.size() && buf_[p] == '<') p++;
    return p < buf_.size() ? (unsigned char)buf_[p] : -1;
  }
  std::string ReadToken() {
    if (!binary_) SkipWs();
    size_t s = pos_;
    while (pos_ < buf_.size() && !isspace((unsigned char)buf_[pos_])) pos_++;
    if (pos_ == s) RS_FAIL(path_ << ": expected a token at byte " << s);
    std::string t = buf_.substr(s, pos_ - s);
    if (pos_ < buf_.size()) pos_++;  // the terminating space
    return t;
  }
  void ExpectToken(const char *tok) {
    std::string t = ReadToken();
    if (t != tok) RS_FAIL(path_ << ": expected token " << tok << ", got " << t);
  }
  // Kaldi's ExpectOneOrTwoTokens: [tok1] tok2
  void ExpectOneOrTwo(const char *tok1, const char *tok2) {
    std::string t = ReadToken();
    if (t == tok1) t = ReadToken();
    if (t != tok2) RS_FAIL(path_ << ": expected token " << tok2 << ", got " << t);
  }
  std::string ReadLine() {  // raw text up to and excluding '\n'
    size_t s = pos_;
    while (pos_ < buf_.size() && buf_[pos_] != '\n') pos_++;
    std::string l = buf_.substr(s, pos_ - s);
    if (pos_ < buf_.size()) pos_++;
    if (!l.empty() && l.back() == '\r') l.pop_back();
    return l;
  }

  int32_t ReadInt32() {
    if (binary_) {
      int sz = (signed char)Byte();
      if (sz != 4) RS_FAIL(path_ << ": expected a 4-byte integer, size byte is " << sz);
      int32_t v;
      Raw(&v, 4);
      return v;
    }
    return (int32_t)strtol(TextWord().c_str(), nullptr, 10);
  }
  float ReadFloat() {
    if (binary_) {
      int sz = (signed char)Byte();
      if (sz == 4) {
        float v;
        Raw(&v, 4);
        return v;
      } else if (sz == 8) {
        double v;
        Raw(&v, 8);
        return (float)v;
      }
      RS_FAIL(path_ << ": expected a float, size byte is " << sz);
    }
    return TextFloat<float>();
  }
  double ReadDouble() {
    if (binary_) {
      int sz = (signed char)Byte();
      if (sz == 8) {
        double v;
        Raw(&v, 8);
        return v;
      } else if (sz == 4) {
        float v;
        Raw(&v, 4);
        return v;
      }
      RS_FAIL(path_ << ": expected a double, size byte is " << sz);
    }
    return TextFloat<double>();
  }
  bool ReadBool() {
    if (!binary_) SkipWs();
    int c = Byte();
    if (c != 'T' && c != 'F') RS_FAIL(path_ << ": expected T or F");
    if (!binary_ && pos_ < buf_.size()) pos_++;
    return c == 'T';
  }
  std::vector<int32_t> ReadIntVector() {
    std::vector<int32_t> v;
    if (binary_) {
      int sz = (signed char)Byte();
      if (sz != 4) RS_FAIL(path_ << ": integer vector with element size " << sz);
      int32_t n;
      Raw(&n, 4);
      if (n < 0) RS_FAIL(path_ << ": negative vector size");
      v.resize(n);
      if (n) Raw(v.data(), (size_t)n * 4);
    } else {
      SkipWs();
      if (Byte() != '[') RS_FAIL(path_ << ": expected [ at start of integer vector");
      while (true) {
        SkipWs();
        if (Peek() == ']') {
          pos_++;
          break;
        }
        v.push_back((int32_t)strtol(TextWord().c_str(), nullptr, 10));
      }
    }
    return v;
  }

  // float or double vector -> double storage (exact for both)
  std::vector<double> ReadVectorD() {
    std::vector<double> v;
    if (binary_) {
      std::string t = ReadToken();
      if (t != "FV" && t != "DV") RS_FAIL(path_ << ": expected FV/DV, got " << t);
      int32_t n = ReadInt32();
      v.resize(n);
      if (t == "FV") {
        std::vector<float> tmp(n);
        if (n) Raw(tmp.data(), (size_t)n * 4);
        for (int i = 0; i < n; i++) v[i] = tmp[i];
      } else if (n) {
        Raw(v.data(), (size_t)n * 8);
      }
    } else {
      SkipWs();
      if (Byte() != '[') RS_FAIL(path_ << ": expected [ at start of vector");
      while (true) {
        SkipWs();
        if (Peek() == ']') {
          pos_++;
          break;
        }
        v.push_back(TextFloat<double>());
      }
    }
    return v;
  }
  std::vector<float> ReadVectorF() {
    std::vector<double> d = ReadVectorD();
    return std::vector<float>(d.begin(), d.end());
  }
  MatrixD ReadMatrixD() {
    MatrixD m;
    if (binary_) {
      std::string t = ReadToken();
      if (t == "CM" || t == "CM2" || t == "CM3") return ReadCompressed(t);
      if (t != "FM" && t != "DM") RS_FAIL(path_ << ": expected FM/DM/CM matrix header, got " << t);
      m.rows = ReadInt32();
      m.cols = ReadInt32();
      size_t n = (size_t)m.rows * m.cols;
      m.d.resize(n);
      if (t == "FM") {
        std::vector<float> tmp(n);
        if (n) Raw(tmp.data(), n * 4);
        for (size_t i = 0; i < n; i++) m.d[i] = tmp[i];
      } else if (n) {
        Raw(m.d.data(), n * 8);
      }
    } else {
      SkipWs();
      if (Byte() != '[') RS_FAIL(path_ << ": expected [ at start of matrix");
      std::vector<double> row;
      int cols = -1;
      while (true) {
        // skip blanks but stop at newlines, which end a row
        while (pos_ < buf_.size() && (buf_[pos_] == ' ' || buf_[pos_] == '\t' || buf_[pos_] == '\r')) pos_++;
        if (pos_ >= buf_.size()) RS_FAIL(path_ << ": unterminated matrix");
        char c = buf_[pos_];
        if (c == '\n' || c == ';' || c == ']') {
          pos_++;
          if (!row.empty()) {
            if (cols < 0) cols = (int)row.size();
            if ((int)row.size() != cols) RS_FAIL(path_ << ": ragged text matrix");
            m.d.insert(m.d.end(), row.begin(), row.end());
            m.rows++;
            row.clear();
          }
          if (c == ']') break;
          continue;
        }
        row.push_back(TextFloat<double>());
      }
      m.cols = cols < 0 ? 0 : cols;
    }
    return m;
  }
  // CompressedMatrix (kaldi/src/matrix/compressed-matrix.cc:565-660, read through Matrix::Read,
  // kaldi-matrix.cc:1475-1513): global header {min, range, rows, cols} without the format word, then
  //   CM  : per-column header of four uint16 percentiles, then one byte per element, column-major;
  //   CM2 : uint16 per element, row-major;   CM3 : uint8 per element, row-major.
  // Expanded with the reference's float expressions (CopyToMat :600-660, CharToFloat :490-500).
  MatrixD ReadCompressed(const std::string &tok) {
    MatrixD m;
    float min_value, range;
    int32_t rows, cols;
    Raw(&min_value, 4);
    Raw(&range, 4);
    Raw(&rows, 4);
    Raw(&cols, 4);
    if (rows < 0 || cols < 0) RS_FAIL(path_ << ": bad compressed-matrix header");
    m.rows = rows;
    m.cols = cols;
    if (cols == 0) {
      m.rows = 0;
      return m;
    }
    const size_t n = (size_t)rows * cols;
    m.d.resize(n);
    if (tok == "CM") {
      std::vector<uint16_t> hdr((size_t)cols * 4);
      Raw(hdr.data(), hdr.size() * 2);
      std::vector<uint8_t> bytes(n);
      Raw(bytes.data(), n);
      auto u16 = [&](uint16_t v) { return min_value + range * 1.52590218966964e-05F * v; };
      for (int c = 0; c < cols; c++) {
        const float p0 = u16(hdr[4 * c]), p25 = u16(hdr[4 * c + 1]), p75 = u16(hdr[4 * c + 2]), p100 = u16(hdr[4 * c + 3]);
        for (int r = 0; r < rows; r++) {
          const uint8_t v = bytes[(size_t)c * rows + r];
          float f;
          if (v <= 64) f = p0 + (p25 - p0) * v * (1 / 64.0);
          else if (v <= 192) f = p25 + (p75 - p25) * (v - 64) * (1 / 128.0);
          else f = p75 + (p100 - p75) * (v - 192) * (1 / 63.0);
          m.d[(size_t)r * cols + c] = f;
        }
      }
    } else if (tok == "CM2") {
      std::vector<uint16_t> data(n);
      Raw(data.data(), n * 2);
      const float increment = range * (1.0 / 65535.0);
      for (size_t i = 0; i < n; i++) m.d[i] = min_value + data[i] * increment;
    } else {
      std::vector<uint8_t> data(n);
      Raw(data.data(), n);
      const float increment = range * (1.0 / 255.0);
      for (size_t i = 0; i < n; i++) m.d[i] = min_value + data[i] * increment;
    }
    return m;
  }
  MatrixF ReadMatrixF() {
    MatrixD d = ReadMatrixD();
    MatrixF m;
    m.rows = d.rows;
    m.cols = d.cols;
    m.d.assign(d.d.begin(), d.d.end());
    return m;
  }
  // packed symmetric matrix (lower triangle, row by row) -> packed double vector, dim returned
  std::vector<double> ReadSpMatrixD(int *dim) {
    std::vector<double> v;
    if (binary_) {
      std::string t = ReadToken();
      if (t != "FP" && t != "DP") RS_FAIL(path_ << ": expected FP/DP, got " << t);
      int32_t n = ReadInt32();
      *dim = n;
      size_t num = (size_t)n * (n + 1) / 2;
      v.resize(num);
      if (t == "FP") {
        std::vector<float> tmp(num);
        if (num) Raw(tmp.data(), num * 4);
        for (size_t i = 0; i < num; i++) v[i] = tmp[i];
      } else if (num) {
        Raw(v.data(), num * 8);
      }
    } else {
      SkipWs();
      if (Byte() != '[') RS_FAIL(path_ << ": expected [ at start of packed matrix");
      while (true) {
        SkipWs();
        if (Peek() == ']') {
          pos_++;
          break;
        }
        v.push_back(TextFloat<double>());
      }
      int n = 0;
      while ((size_t)n * (n + 1) / 2 < v.size()) n++;
      if ((size_t)n * (n + 1) / 2 != v.size()) RS_FAIL(path_ << ": bad packed matrix size");
      *dim = n;
    }
    return v;
  }

 private:
  int Byte() {
    if (pos_ >= buf_.size()) RS_FAIL(path_ << ": unexpected end of file");
    return (unsigned char)buf_[pos_++];
  }
  void Raw(void *dst, size_t n) {
    if (pos_ + n > buf_.size()) RS_FAIL(path_ << ": unexpected end of file");
    memcpy(dst, buf_.data() + pos_, n);
    pos_ += n;
  }
  std::string TextWord() {
    SkipWs();
    size_t s = pos_;
    while (pos_ < buf_.size() && !isspace((unsigned char)buf_[pos_]) && buf_[pos_] != ']') pos_++;
    if (pos_ == s) RS_FAIL(path_ << ": expected a number at byte " << s);
    return buf_.substr(s, pos_ - s);
  }
  template <typename T>
  T TextFloat() {
    std::string w = TextWord();
    // Kaldi prints inf/nan in several spellings (io-funcs-inl.h ReadBasicType<float>)
    std::string l;
    for (char c : w) l.push_back((char)tolower(c));
    if (l == "inf" || l == "infinity" || l == "+inf") return (T)HUGE_VAL;
    if (l == "-inf" || l == "-infinity") return (T)-HUGE_VAL;
    if (l == "nan" || l == "-nan" || l == "1.#qnan") return (T)strtod("nan", nullptr);
    char *end = nullptr;
    double v = strtod(w.c_str(), &end);
    if (end == w.c_str()) RS_FAIL(path_ << ": bad number '" << w << "'");
    return (T)v;
  }

  std::string path_, buf_;
  size_t pos_ = 0;
  bool binary_ = false;
};

}  // namespace rs
