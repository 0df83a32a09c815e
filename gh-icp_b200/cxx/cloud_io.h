// cloud_io.h — point-cloud file I/O for the command-line driver: the formats the reference's
// DataIo reads through PCL (include/dataio.hpp:26-139, 490-585) without PCL:
//   .pcd  ASCII and binary (uncompressed) — any field list that contains x, y, z as 4-byte floats (pcl::io::loadPCDFile)
//   .ply  ASCII and binary_little_endian — vertex element with float x, y, z properties (pcl::io::loadPLYFile)
//   .txt  "x y z" per line, read as doubles and narrowed to float like readTxtFile (include/dataio.hpp:508-533)
// Not provided: .las (needs libLAS and the reference's interactive shift prompt, :36-54), binary_compressed PCD.
// Points are float32 x, y, z like pcl::PointXYZ; written clouds use the same formats (PCD binary, PLY binary, TXT %.6f
// like writeTxtFile :535-560).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ghicp {

struct Cloud {                   // flat float32 x, y, z triples
  std::vector<float> xyz;
  size_t size() const { return xyz.size() / 3; }
  void push(float x, float y, float z) { xyz.push_back(x); xyz.push_back(y); xyz.push_back(z); }
};

inline std::string file_extension(const std::string &name) {   // include/dataio.hpp:28-29
  const size_t p = name.find_last_of('.');
  return p == std::string::npos ? std::string() : name.substr(p + 1);
}

inline void read_txt(const std::string &name, Cloud &c) {
  std::ifstream in(name.c_str());
  if (!in) throw std::runtime_error("cannot open " + name);
  double x, y, z;
  while (in >> x >> y >> z) {      // include/dataio.hpp:517-528 (extra columns are not supported there either)
    c.push((float)x, (float)y, (float)z);
    std::string rest;
    std::getline(in, rest);
  }
}
inline void write_txt(const std::string &name, const Cloud &c) {
  std::ofstream ofs(name.c_str());
  if (!ofs) throw std::runtime_error("cannot write " + name);
  ofs << std::setiosflags(std::ios::fixed) << std::setprecision(6);
  for (size_t i = 0; i < c.size(); ++i) ofs << c.xyz[3 * i] << "  " << c.xyz[3 * i + 1] << "  " << c.xyz[3 * i + 2] << "\n";
}

struct FieldLayout { int off[3] = {-1, -1, -1}; int stride = 0; int col[3] = {-1, -1, -1}; int ncols = 0; };

inline void read_pcd(const std::string &name, Cloud &c) {
  std::ifstream in(name.c_str(), std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + name);
  std::vector<std::string> fields; std::vector<int> sizes, counts; std::vector<char> types;
  size_t points = 0, width = 0, height = 1;
  std::string data, line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key; ls >> key;
    if (key == "FIELDS") { std::string f; while (ls >> f) fields.push_back(f); }
    else if (key == "SIZE") { int v; while (ls >> v) sizes.push_back(v); }
    else if (key == "TYPE") { char v; while (ls >> v) types.push_back(v); }
    else if (key == "COUNT") { int v; while (ls >> v) counts.push_back(v); }
    else if (key == "WIDTH") ls >> width;
    else if (key == "HEIGHT") ls >> height;
    else if (key == "POINTS") ls >> points;
    else if (key == "DATA") { ls >> data; break; }
  }
  if (fields.empty() || sizes.size() != fields.size()) throw std::runtime_error(name + ": malformed PCD header");
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (types.empty()) types.assign(fields.size(), 'F');
  if (counts.size() != fields.size() || types.size() != fields.size()) throw std::runtime_error(name + ": malformed PCD header (TYPE / COUNT shorter than FIELDS)");
  for (size_t f = 0; f < fields.size(); ++f)
    if (sizes[f] <= 0 || counts[f] <= 0) throw std::runtime_error(name + ": malformed PCD header (non-positive SIZE / COUNT)");
  if (points == 0) points = width * height;
  FieldLayout L;
  for (size_t f = 0; f < fields.size(); ++f) {
    for (int a = 0; a < 3; ++a)
      if (fields[f] == std::string(1, "xyz"[a])) {
        if (sizes[f] != 4 || types[f] != 'F') throw std::runtime_error(name + ": x, y, z must be 4-byte floats");
        L.off[a] = L.stride; L.col[a] = L.ncols;
      }
    L.stride += sizes[f] * counts[f];
    L.ncols += counts[f];
  }
  if (L.off[0] < 0 || L.off[1] < 0 || L.off[2] < 0) throw std::runtime_error(name + ": PCD without x, y, z fields");
  c.xyz.reserve(3 * points);
  if (data == "ascii") {
    std::vector<double> row(L.ncols);
    for (size_t i = 0; i < points; ++i) {
      for (int k = 0; k < L.ncols; ++k) if (!(in >> row[k])) throw std::runtime_error(name + ": truncated ASCII PCD");
      c.push((float)row[L.col[0]], (float)row[L.col[1]], (float)row[L.col[2]]);
    }
  } else if (data == "binary") {
    std::vector<char> buf((size_t)L.stride * points);
    in.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)in.gcount() != buf.size()) throw std::runtime_error(name + ": truncated binary PCD");
    for (size_t i = 0; i < points; ++i) {
      float v[3];
      for (int a = 0; a < 3; ++a) std::memcpy(&v[a], &buf[i * L.stride + L.off[a]], 4);
      c.push(v[0], v[1], v[2]);
    }
  } else {
    throw std::runtime_error(name + ": PCD DATA '" + data + "' is not supported (ascii and binary are)");
  }
}
inline void write_pcd(const std::string &name, const Cloud &c) {   // pcl::io::savePCDFileBinary layout for PointXYZ
  std::ofstream ofs(name.c_str(), std::ios::binary);
  if (!ofs) throw std::runtime_error("cannot write " + name);
  ofs << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
      << "WIDTH " << c.size() << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << c.size() << "\nDATA binary\n";
  ofs.write((const char *)c.xyz.data(), (std::streamsize)(c.xyz.size() * sizeof(float)));
}

inline int ply_type_size(const std::string &t) {
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  throw std::runtime_error("PLY: unknown property type " + t);
}
inline void read_ply(const std::string &name, Cloud &c) {
  std::ifstream in(name.c_str(), std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + name);
  std::string line, format;
  size_t nvert = 0;
  bool in_vertex = false;
  FieldLayout L;
  std::vector<int> psize; std::vector<std::string> ptype;
  if (!std::getline(in, line) || line.substr(0, 3) != "ply") throw std::runtime_error(name + ": not a PLY file");
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string key; ls >> key;
    if (key == "format") ls >> format;
    else if (key == "element") { std::string what; size_t n; ls >> what >> n; in_vertex = (what == "vertex"); if (in_vertex) nvert = n; }
    else if (key == "property" && in_vertex) {
      std::string t, nm; ls >> t;
      if (t == "list") throw std::runtime_error(name + ": list property in the vertex element");
      ls >> nm;
      const int sz = ply_type_size(t);
      for (int a = 0; a < 3; ++a)
        if (nm == std::string(1, "xyz"[a])) {
          if (!(t == "float" || t == "float32")) throw std::runtime_error(name + ": x, y, z must be float properties");
          L.off[a] = L.stride; L.col[a] = L.ncols;
        }
      L.stride += sz; L.ncols += 1; psize.push_back(sz); ptype.push_back(t);
    } else if (key == "end_header") break;
  }
  if (L.off[0] < 0 || L.off[1] < 0 || L.off[2] < 0) throw std::runtime_error(name + ": PLY vertex element without x, y, z");
  c.xyz.reserve(3 * nvert);
  if (format == "ascii") {
    std::vector<double> row(L.ncols);
    for (size_t i = 0; i < nvert; ++i) {
      for (int k = 0; k < L.ncols; ++k) if (!(in >> row[k])) throw std::runtime_error(name + ": truncated ASCII PLY");
      c.push((float)row[L.col[0]], (float)row[L.col[1]], (float)row[L.col[2]]);
    }
  } else if (format == "binary_little_endian") {
    std::vector<char> buf((size_t)L.stride * nvert);
    in.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)in.gcount() != buf.size()) throw std::runtime_error(name + ": truncated binary PLY");
    for (size_t i = 0; i < nvert; ++i) {
      float v[3];
      for (int a = 0; a < 3; ++a) std::memcpy(&v[a], &buf[i * L.stride + L.off[a]], 4);
      c.push(v[0], v[1], v[2]);
    }
  } else {
    throw std::runtime_error(name + ": PLY format '" + format + "' is not supported");
  }
}
inline void write_ply(const std::string &name, const Cloud &c) {
  std::ofstream ofs(name.c_str(), std::ios::binary);
  if (!ofs) throw std::runtime_error("cannot write " + name);
  ofs << "ply\nformat binary_little_endian 1.0\nelement vertex " << c.size()
      << "\nproperty float x\nproperty float y\nproperty float z\nend_header\n";
  ofs.write((const char *)c.xyz.data(), (std::streamsize)(c.xyz.size() * sizeof(float)));
}

// DataIo::readCloudFile / writeCloudFile dispatch on the suffix (include/dataio.hpp:26-119)
inline void read_cloud(const std::string &name, Cloud &c) {
  const std::string e = file_extension(name);
  if (e == "pcd") read_pcd(name, c);
  else if (e == "ply") read_ply(name, c);
  else if (e == "txt") read_txt(name, c);
  else if (e == "las") throw std::runtime_error(name + ": .las needs libLAS (and the reference asks for a shift interactively, include/dataio.hpp:36-54): convert to pcd / ply / txt");
  else throw std::runtime_error(name + ": undefined point cloud format");
  if (c.size() == 0) throw std::runtime_error(name + ": no points");
}
inline void write_cloud(const std::string &name, const Cloud &c) {
  const std::string e = file_extension(name);
  if (e == "pcd") write_pcd(name, c);
  else if (e == "ply") write_ply(name, c);
  else if (e == "txt") write_txt(name, c);
  else throw std::runtime_error(name + ": undefined point cloud format");
}

}  // namespace ghicp
