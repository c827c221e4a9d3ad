// vtu_writer.h - host-side writer for particle fields held by the device path.
//
// Produces, byte for byte, the `.vtu` file the reference's Points<S>::write_vtk writes (reference src/Points.h:851-1039
// through src/VtkXmlWriter.h:36-148 and the tinyxml2 XMLPrinter it drives, lib/tinyxml2/tinyxml2.cpp:2486-2760), so that
// fields advanced on the GPU can be diffed against the reference's output files with `cmp`. Nothing of tinyxml2 or
// cppcodec is used: the printer below reproduces the handful of formatting rules that matter here -
//   * declaration `<?xml version="1.0"?>` first, no newline before it;
//   * every element opens on a new line indented 4 spaces per depth, unless text has been pushed into its parent;
//   * attributes print in the order of a std::map<std::string,...> (the reference collects them in one), i.e. sorted
//     by key, after the one explicit `format="binary"` pushed by writeDataArray (which comes LAST: it is pushed after
//     the map's entries);
//   * a DataArray's payload is " " + base64(uint32 = length of the ENCODED payload) + base64(raw bytes) + " "
//     (VtkXmlWriter.h:101-110 - the header holds the encoded length, not the raw byte count; kept as is);
//   * an element that received text closes on the same line; others close on their own line; a final newline
//     follows the root's close.
// Pure host I/O: no device work and no arithmetic of the path happens here.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace o3d {

inline std::string base64_encode(const void* data, size_t n) {
  static const char* tbl = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  const unsigned char* p = static_cast<const unsigned char*>(data);
  std::string out;
  out.reserve(((n + 2) / 3) * 4);
  size_t i = 0;
  for (; i + 2 < n; i += 3) {
    const unsigned v = (p[i] << 16) | (p[i + 1] << 8) | p[i + 2];
    out.push_back(tbl[v >> 18]); out.push_back(tbl[(v >> 12) & 63]); out.push_back(tbl[(v >> 6) & 63]); out.push_back(tbl[v & 63]);
  }
  if (i + 1 == n) {
    const unsigned v = p[i] << 16;
    out.push_back(tbl[v >> 18]); out.push_back(tbl[(v >> 12) & 63]); out.push_back('='); out.push_back('=');
  } else if (i + 2 == n) {
    const unsigned v = (p[i] << 16) | (p[i + 1] << 8);
    out.push_back(tbl[v >> 18]); out.push_back(tbl[(v >> 12) & 63]); out.push_back(tbl[(v >> 6) & 63]); out.push_back('=');
  }
  return out;
}

// the subset of tinyxml2::XMLPrinter (non-compact mode) that VtkXmlWriter exercises
class XmlOut {
 public:
  explicit XmlOut(std::FILE* fp) : fp_(fp) {}
  void declaration(const char* v) {
    seal();
    if (text_depth_ < 0 && !first_) { std::fputc('\n', fp_); space(depth_); }
    first_ = false;
    std::fputs("<?", fp_); std::fputs(v, fp_); std::fputs("?>", fp_);
  }
  void open(const char* name) {
    seal();
    stack_.push_back(name);
    if (text_depth_ < 0 && !first_) std::fputc('\n', fp_);
    space(depth_);
    std::fputc('<', fp_); std::fputs(name, fp_);
    just_opened_ = true;
    first_ = false;
    ++depth_;
  }
  void attribute(const char* name, const char* value) {
    std::fputc(' ', fp_); std::fputs(name, fp_); std::fputs("=\"", fp_); std::fputs(value, fp_); std::fputc('"', fp_);
  }
  void text(const char* t) {
    text_depth_ = depth_ - 1;
    seal();
    std::fputs(t, fp_);   // payloads here are base64 and blanks: nothing to escape
  }
  void close() {
    --depth_;
    const std::string name = stack_.back();
    stack_.pop_back();
    if (just_opened_) {
      std::fputs("/>", fp_);
    } else {
      if (text_depth_ < 0) { std::fputc('\n', fp_); space(depth_); }
      std::fputs("</", fp_); std::fputs(name.c_str(), fp_); std::fputc('>', fp_);
    }
    if (text_depth_ == depth_) text_depth_ = -1;
    if (depth_ == 0) std::fputc('\n', fp_);
    just_opened_ = false;
  }
  int depth() const { return depth_; }

 private:
  void seal() {
    if (!just_opened_) return;
    just_opened_ = false;
    std::fputc('>', fp_);
  }
  void space(int d) { for (int i = 0; i < d; ++i) std::fputs("    ", fp_); }
  std::FILE* fp_;
  std::vector<std::string> stack_;
  int depth_ = 0, text_depth_ = -1;
  bool just_opened_ = false, first_ = true;
};

class VtuWriter {
 public:
  explicit VtuWriter(std::FILE* fp) : x_(fp) {
    x_.declaration("xml version=\"1.0\"");
    x_.open("VTKFile");
    x_.attribute("type", "UnstructuredGrid");
    x_.attribute("version", "0.1");
    x_.attribute("byte_order", "LittleEndian");
    x_.attribute("header_type", "UInt32");
    x_.open("UnstructuredGrid");
  }
  void element(const char* name, const std::map<std::string, std::string>& attrs = {}) {
    x_.open(name);
    for (const auto& kv : attrs) x_.attribute(kv.first.c_str(), kv.second.c_str());
  }
  void data(const void* bytes, size_t n) {
    x_.attribute("format", "binary");
    const std::string enc = base64_encode(bytes, n);
    const uint32_t len = (uint32_t)enc.size();
    const std::string hdr = base64_encode(&len, sizeof len);
    x_.text(" "); x_.text(hdr.c_str()); x_.text(enc.c_str()); x_.text(" ");
  }
  void close() { x_.close(); }
  void finish() { while (x_.depth() > 0) x_.close(); }

 private:
  XmlOut x_;
};

// Points<S>::write_vtk (src/Points.h:851-1039). s == nullptr and r == nullptr: inert points ("fldpt_" files).
// Arrays are SoA (one pointer per component), as the containers hold them. Returns false if the file cannot be opened.
inline bool write_points_vtu(const char* path, int64_t n, const float* const* x, const float* const* s, const float* r,
                             const float* const* u, double time) {
  std::FILE* fp = std::fopen(path, "wb");
  if (!fp) return false;
  VtuWriter w(fp);
  auto interleave = [n](const float* const* a) {
    std::vector<float> v(3 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) { v[3 * i] = a[0][i]; v[3 * i + 1] = a[1][i]; v[3 * i + 2] = a[2][i]; }
    return v;
  };
  w.element("FieldData");
  w.element("DataArray", {{"type", "Float64"}, {"Name", "TimeValue"}, {"NumberOfTuples", "1"}});
  w.data(&time, sizeof time);
  w.close();
  w.close();
  w.element("Piece", {{"NumberOfPoints", std::to_string(n)}, {"NumberOfCells", std::to_string(n)}});
  w.element("Points");
  w.element("DataArray", {{"NumberOfComponents", "3"}, {"Name", "position"}, {"type", "Float32"}});
  { const std::vector<float> v = interleave(x); w.data(v.data(), v.size() * 4); }
  w.close();
  w.close();
  w.element("Cells");
  {
    std::vector<int32_t> v((size_t)n);
    w.element("DataArray", {{"Name", "connectivity"}, {"type", "Int32"}});
    for (int64_t i = 0; i < n; ++i) v[i] = (int32_t)i;
    w.data(v.data(), v.size() * 4);
    w.close();
    w.element("DataArray", {{"Name", "offsets"}, {"type", "Int32"}});
    for (int64_t i = 0; i < n; ++i) v[i] = (int32_t)(i + 1);
    w.data(v.data(), v.size() * 4);
    w.close();
    std::vector<uint8_t> t((size_t)n, 1);
    w.element("DataArray", {{"Name", "types"}, {"type", "UInt8"}});
    w.data(t.data(), t.size());
    w.close();
  }
  w.close();
  {
    std::map<std::string, std::string> a = {{"Vectors", "velocity"}};
    if (r) a.insert({"Scalars", "radius"});
    w.element("PointData", a);
  }
  if (s) {
    w.element("DataArray", {{"NumberOfComponents", "3"}, {"Name", "circulation"}, {"type", "Float32"}});
    const std::vector<float> v = interleave(s);
    w.data(v.data(), v.size() * 4);
    w.close();
  }
  if (r) {
    w.element("DataArray", {{"Name", "radius"}, {"type", "Float32"}});
    w.data(r, (size_t)n * 4);
    w.close();
  }
  w.element("DataArray", {{"NumberOfComponents", "3"}, {"Name", "velocity"}, {"type", "Float32"}});
  { const std::vector<float> v = interleave(u); w.data(v.data(), v.size() * 4); }
  w.close();
  w.close();   // PointData
  w.close();   // Piece
  w.finish();
  std::fclose(fp);
  return true;
}

}  // namespace o3d
