#include "io.h"

#include <cstdlib>
#include <cstring>
#include <fstream>

#include "log.h"

namespace adapm {

namespace {
std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  ADAPM_CHECK(f.good(), "cannot open '" << path << "'");
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0, std::ios::beg);
  std::string s((size_t)n, '\0');
  f.read(&s[0], n);
  return s;
}
inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p;
  return p;
}
inline const char* line_end(const char* p, const char* e) {
  const void* q = memchr(p, '\n', (size_t)(e - p));
  return q ? (const char*)q : e;
}
}  // namespace

std::vector<int64_t> read_triples_file(const std::string& path) {
  const std::string buf = slurp(path);
  const char* p = buf.data();
  const char* e = p + buf.size();
  std::vector<int64_t> out;
  out.reserve(buf.size() / 4);
  while (true) {
    p = skip_ws(p, e);
    if (p >= e) break;
    char* q = nullptr;
    const long long v = strtoll(p, &q, 10);
    ADAPM_CHECK(q != p, path << ": not an integer near byte " << (p - buf.data()));
    out.push_back((int64_t)v);
    p = q;
  }
  ADAPM_CHECK(out.size() % 3 == 0, path << ": the number of values (" << out.size() << ") is not a multiple of 3");
  return out;
}

CooMatrix read_matrix_market_coo_file(const std::string& path) {
  const std::string buf = slurp(path);
  const char* p = buf.data();
  const char* e = p + buf.size();
  ADAPM_CHECK(buf.compare(0, 32, "%%MatrixMarket matrix coordinate") == 0, path << ": not a MatrixMarket coordinate file");
  // comment lines
  while (p < e && *p == '%') p = line_end(p, e) + 1;
  CooMatrix m;
  char* q = nullptr;
  m.rows = strtoll(p, &q, 10); p = q;
  m.cols = strtoll(p, &q, 10); p = q;
  const long long nnz = strtoll(p, &q, 10); p = q;
  ADAPM_CHECK(m.rows > 0 && m.cols > 0 && nnz >= 0, path << ": bad size line");
  m.i.reserve((size_t)nnz); m.j.reserve((size_t)nnz); m.x.reserve((size_t)nnz);
  for (long long k = 0; k < nnz; ++k) {
    p = skip_ws(p, e);
    ADAPM_CHECK(p < e, path << ": " << nnz << " entries announced, " << k << " found");
    const long long i = strtoll(p, &q, 10); p = q;
    const long long j = strtoll(p, &q, 10); p = q;
    const double x = strtod(p, &q);
    ADAPM_CHECK(q != p, path << ": malformed entry " << k);
    p = q;
    ADAPM_CHECK(i >= 1 && i <= m.rows && j >= 1 && j <= m.cols, path << ": entry " << k << " is outside the matrix");
    m.i.push_back(i - 1); m.j.push_back(j - 1); m.x.push_back((float)x);
  }
  return m;
}

}  // namespace adapm
