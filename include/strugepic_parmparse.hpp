// strugepic_parmparse.hpp -- reader of the reference's input decks.
//
// Every shipped StrugePIC driver reads its deck through `amrex::ParmParse pp; pp.get("key", value);`
// (e.g. test/single_particle/main.cpp:41-77, examples/full/bernstein_main.cpp:41-86 of MoPHA/strugepic)
// after `amrex::Initialize(argc, argv)` took the deck path from argv[1] and `key=value` overrides from the
// remaining arguments.  This header re-provides that surface without AMReX so that the decks
// (`*.input`: `key = v1 v2 v3`, `#` comments, optional quotes, a stray `;` after a value as in
// `output_interval = 100 ;`) run unchanged against strugepic_b200:
//
//   strugepic::Initialize(argc, argv);         // amrex::Initialize
//   strugepic::ParmParse pp;                   // amrex::ParmParse
//   pp.get("n_cell", n_cell);                  // std::array<int,3>, int, double, std::string, std::vector<T>
//   if (pp.query("order", order)) ...          // optional keys
//   strugepic::Finalize();
//
// Semantics kept from AMReX: the LAST definition of a key wins (so command-line overrides beat the
// file), `get` on a missing key or a value of the wrong type is fatal (amrex::Abort -> here a
// strugepic::ParmParseError), `query` returns whether the key was present, an optional prefix makes
// `ParmParse pp("amr")` look up `amr.key`.
#ifndef STRUGEPIC_PARMPARSE_HPP
#define STRUGEPIC_PARMPARSE_HPP

#include <array>
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace strugepic {

class ParmParseError : public std::runtime_error {
 public:
  explicit ParmParseError(const std::string& what) : std::runtime_error("ParmParse: " + what) {}
};

namespace detail {

struct ParmTable {
  std::vector<std::pair<std::string, std::vector<std::string>>> entries;  // in definition order
  static ParmTable& global() {
    static ParmTable t;
    return t;
  }
};

// Splits deck text into tokens: whitespace separates, `#` starts a comment, "..." is one token,
// `=` is a token of its own.
inline std::vector<std::string> tokenize(const std::string& text) {
  std::vector<std::string> out;
  std::string cur;
  bool have = false;
  auto flush = [&]() {
    if (have) out.push_back(cur);
    cur.clear();
    have = false;
  };
  for (std::size_t p = 0; p < text.size(); ++p) {
    const char ch = text[p];
    if (ch == '#') {
      flush();
      while (p < text.size() && text[p] != '\n') ++p;
    } else if (ch == '"') {
      flush();
      std::size_t e = text.find('"', p + 1);
      if (e == std::string::npos) throw ParmParseError("unterminated string");
      out.push_back(text.substr(p + 1, e - p - 1));
      p = e;
    } else if (ch == '=') {
      flush();
      out.push_back("=");
    } else if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') {
      flush();
    } else {
      cur.push_back(ch);
      have = true;
    }
  }
  flush();
  return out;
}

// `name = v v v name2 = v ...`: a token followed by `=` opens a new definition
inline void add_definitions(ParmTable& t, const std::vector<std::string>& tok) {
  std::size_t p = 0;
  while (p < tok.size()) {
    if (p + 1 >= tok.size() || tok[p + 1] != "=" || tok[p] == "=")
      throw ParmParseError("expected `name = value` near `" + tok[p] + "`");
    const std::string name = tok[p];
    p += 2;
    std::vector<std::string> vals;
    while (p < tok.size() && !(p + 1 < tok.size() && tok[p + 1] == "=")) {
      if (tok[p] == "=") throw ParmParseError("misplaced `=` after `" + name + "`");
      if (tok[p] != ";") vals.push_back(tok[p]);  // `output_interval = 100 ;` in the shipped decks
      ++p;
    }
    t.entries.emplace_back(name, vals);
  }
}

template <class T>
inline bool convert(const std::string& s, T& out);
template <>
inline bool convert<std::string>(const std::string& s, std::string& out) {
  out = s;
  return true;
}
template <>
inline bool convert<double>(const std::string& s, double& out) {
  char* end = nullptr;
  errno = 0;
  out = std::strtod(s.c_str(), &end);
  return end != s.c_str() && *end == '\0';
}
template <>
inline bool convert<float>(const std::string& s, float& out) {
  double d;
  const bool ok = convert<double>(s, d);
  out = (float)d;
  return ok;
}
template <>
inline bool convert<long>(const std::string& s, long& out) {
  char* end = nullptr;
  errno = 0;
  out = std::strtol(s.c_str(), &end, 10);
  if (end != s.c_str() && *end == '\0' && errno == 0) return true;
  double d;  // AMReX accepts `1e3` for an integer when it is integral
  if (convert<double>(s, d) && d == (double)(long)d) {
    out = (long)d;
    return true;
  }
  return false;
}
template <>
inline bool convert<int>(const std::string& s, int& out) {
  long l;
  const bool ok = convert<long>(s, l);
  out = (int)l;
  return ok && l == (long)(int)l;
}
template <>
inline bool convert<bool>(const std::string& s, bool& out) {
  if (s == "true" || s == "t" || s == "True" || s == "1") return out = true, true;
  if (s == "false" || s == "f" || s == "False" || s == "0") return out = false, true;
  return false;
}

}  // namespace detail

class ParmParse {
 public:
  explicit ParmParse(std::string prefix = std::string()) : prefix_(std::move(prefix)) {}

  // amrex::Initialize's part of the job: argv[1] without `=` is the deck, the rest are overrides
  static void Initialize(int argc, char** argv) {
    detail::ParmTable& t = detail::ParmTable::global();
    t.entries.clear();
    int first = 1;
    if (argc > 1 && std::string(argv[1]).find('=') == std::string::npos) {
      addfile(argv[1]);
      first = 2;
    }
    std::string rest;
    for (int a = first; a < argc; ++a) rest += std::string(argv[a]) + "\n";
    detail::add_definitions(t, detail::tokenize(rest));
  }
  static void addfile(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw ParmParseError("cannot open deck `" + path + "`");
    std::stringstream ss;
    ss << f.rdbuf();
    addtext(ss.str());
  }
  static void addtext(const std::string& text) {
    detail::add_definitions(detail::ParmTable::global(), detail::tokenize(text));
  }
  static void Finalize() { detail::ParmTable::global().entries.clear(); }

  bool contains(const std::string& name) const { return find(name) != nullptr; }
  int countval(const std::string& name) const {
    const std::vector<std::string>* v = find(name);
    return v ? (int)v->size() : 0;
  }

  // ---- scalars: the ival-th value of the last definition --------------------------------------
  template <class T>
  int query(const std::string& name, T& ref, int ival = 0) const {
    const std::vector<std::string>* v = find(name);
    if (!v) return 0;
    if (ival < 0 || ival >= (int)v->size())
      throw ParmParseError("`" + full(name) + "` has no value #" + std::to_string(ival));
    T tmp;
    if (!detail::convert<T>((*v)[ival], tmp))
      throw ParmParseError("`" + full(name) + "`: cannot read `" + (*v)[ival] + "` as the requested type");
    ref = tmp;
    return 1;
  }
  template <class T>
  void get(const std::string& name, T& ref, int ival = 0) const {
    if (!query(name, ref, ival)) throw ParmParseError("`" + full(name) + "` not found in the inputs");
  }

  // ---- arrays ---------------------------------------------------------------------------------
  template <class T>
  int queryarr(const std::string& name, std::vector<T>& ref) const {
    const std::vector<std::string>* v = find(name);
    if (!v) return 0;
    std::vector<T> tmp(v->size());
    for (std::size_t i = 0; i < v->size(); ++i)
      if (!detail::convert<T>((*v)[i], tmp[i]))
        throw ParmParseError("`" + full(name) + "`: cannot read `" + (*v)[i] + "` as the requested type");
    ref.swap(tmp);
    return 1;
  }
  template <class T>
  void getarr(const std::string& name, std::vector<T>& ref) const {
    if (!queryarr(name, ref)) throw ParmParseError("`" + full(name) + "` not found in the inputs");
  }
  template <class T, std::size_t N>
  int query(const std::string& name, std::array<T, N>& ref) const {
    std::vector<T> v;
    if (!queryarr(name, v)) return 0;
    if (v.size() < N)
      throw ParmParseError("`" + full(name) + "` needs " + std::to_string(N) + " values, found " +
                           std::to_string(v.size()));
    for (std::size_t i = 0; i < N; ++i) ref[i] = v[i];
    return 1;
  }
  template <class T, std::size_t N>
  void get(const std::string& name, std::array<T, N>& ref) const {
    if (!query(name, ref)) throw ParmParseError("`" + full(name) + "` not found in the inputs");
  }
  template <class T>
  int query(const std::string& name, std::vector<T>& ref) const { return queryarr(name, ref); }
  template <class T>
  void get(const std::string& name, std::vector<T>& ref) const { getarr(name, ref); }

 private:
  std::string full(const std::string& name) const { return prefix_.empty() ? name : prefix_ + "." + name; }
  const std::vector<std::string>* find(const std::string& name) const {
    const std::string key = full(name);
    const auto& e = detail::ParmTable::global().entries;
    for (std::size_t i = e.size(); i-- > 0;)
      if (e[i].first == key) return &e[i].second;
    return nullptr;
  }
  std::string prefix_;
};

}  // namespace strugepic
#endif
