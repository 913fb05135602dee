// boda_base.h -- boost-free mirrors of the Boda types that appear in the rtc_fwd plug-in surface, so that
// b200_compute_t / b200_conv_fwd_t keep the reference's names and argument meaning:
//   rt_err / unsup_err + exceptions     src/boda_base.H:98-105, :1077-1090
//   dim_t / dims_t                      src/boda_base.H:498-690
//   nda_t                               src/boda_base.H:751-810
//   lexp text grammar                   src/lexp.cc:22-30 (escape), :603-621
//   op_base_t                           src/op_base.H:9-41, src/op_base.cc:16-23 (ordering)
//   nda / dims / op text <-> object     src/nesi.cc:661-785, src/boda_base.cc:404-420 (canonical printer)
// Both op-line syntaxes (current str_vals/nda_vals and the stale type/dims_vals one) are accepted: SURVEY Appendix A.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace boda {
using std::map;
using std::shared_ptr;
using std::string;
using std::vector;
typedef vector<string> vect_string;
typedef map<string, string> map_str_str;

struct rt_exception : public std::runtime_error {
  explicit rt_exception(string const &m) : std::runtime_error(m) {}
};
struct unsup_exception : public rt_exception {
  explicit unsup_exception(string const &m) : rt_exception(m) {}
};
[[noreturn]] inline void rt_err(string const &m) { throw rt_exception("error: " + m); }
[[noreturn]] inline void unsup_err(string const &m) { throw unsup_exception("unsupported: " + m); }
#define assert_st(x) do { if (!(x)) { ::boda::rt_err(string("assertion failed: " #x " at ") + __FILE__ + ":" + std::to_string(__LINE__)); } } while (0)

template <typename T> inline string str(T const &v) { std::ostringstream o; o << v; return o.str(); }

// ---- dims_t ---------------------------------------------------------------------------------------------------
struct dim_t {
  uint32_t sz;
  uint32_t stride;
  string name;
};

inline uint32_t tn_bytes(string const &tn) {
  if (tn == "float" || tn == "uint32_t" || tn == "int32_t") { return 4; }
  if (tn == "double" || tn == "uint64_t" || tn == "int64_t") { return 8; }
  if (tn == "half" || tn == "uint16_t" || tn == "int16_t") { return 2; }
  if (tn == "uint8_t" || tn == "int8_t") { return 1; }
  if (tn == "none") { return 0; }
  rt_err("unknown type name '" + tn + "'");
}

struct dims_t : public vector<dim_t> {
  string tn;  // element type name; "float" by default, "none" for dims-only params (kern_sz, stride, in_pad)
  dims_t() : tn("float") {}
  dims_t(vector<uint32_t> const &szs, vect_string const &names, string const &tn_) : tn(tn_) {
    assert_st(szs.size() == names.size());
    for (size_t i = 0; i < szs.size(); ++i) { push_back(dim_t{szs[i], 0, names[i]}); }
    calc_strides();
  }
  void add_dim(string const &name, uint32_t sz) { push_back(dim_t{sz, 0, name}); }
  void calc_strides() {
    uint32_t s = 1;
    for (size_t i = size(); i-- > 0;) { (*this)[i].stride = s; s *= (*this)[i].sz; }
  }
  uint64_t dims_prod() const { uint64_t p = 1; for (auto const &d : *this) { p *= d.sz; } return p; }
  uint64_t bytes_sz() const { return dims_prod() * tn_bytes(tn); }
  dim_t const *get_dim_by_name(string const &n) const { for (auto const &d : *this) { if (d.name == n) { return &d; } } return nullptr; }
  bool has_dim(string const &n) const { return get_dim_by_name(n) != nullptr; }
  uint32_t dsz(string const &n) const { dim_t const *d = get_dim_by_name(n); if (!d) { rt_err("dims " + pretty() + " have no dim named '" + n + "'"); } return d->sz; }
  uint32_t dstride(string const &n) const { dim_t const *d = get_dim_by_name(n); if (!d) { rt_err("no dim named '" + n + "'"); } return d->stride; }
  uint32_t dims(size_t i) const { return at(i).sz; }
  uint32_t strides(size_t i) const { return at(i).stride; }
  bool operator==(dims_t const &o) const {  // exact equality incl. names and type (src/nvrtc_util.cc:302-303 relies on it)
    if (tn != o.tn || size() != o.size()) { return false; }
    for (size_t i = 0; i < size(); ++i) { if (at(i).sz != o.at(i).sz || at(i).name != o.at(i).name) { return false; } }
    return true;
  }
  bool operator!=(dims_t const &o) const { return !(*this == o); }
  bool operator<(dims_t const &o) const { return param_str() < o.param_str(); }
  // canonical text form (src/boda_base.cc:404-420): float is implicit, other types via tn=
  string dims_list_str() const {
    string r = "(";
    for (size_t i = 0; i < size(); ++i) { if (i) { r += ","; } r += at(i).name + "=" + std::to_string(at(i).sz); }
    return r + ")";
  }
  string param_str() const {
    string r = "(";
    if (tn != "float") { r += "tn=" + tn + (empty() ? "" : ","); }
    if (!empty()) { r += "dims=" + dims_list_str(); }
    return r + ")";
  }
  string pretty() const { return (tn == "float" ? string() : tn + ":") + dims_list_str(); }
};

// ---- nda_t ----------------------------------------------------------------------------------------------------
struct nda_t {
  dims_t dims;
  shared_ptr<void> elems;  // may be null: dims-only nda (or a raw-pointer wrapper via rp)
  void *rp = nullptr;      // non-owning raw pointer (get_var_raw_native_pointer)
  nda_t() {}
  explicit nda_t(dims_t const &d, bool alloc = true) : dims(d) {
    if (alloc && d.bytes_sz()) { elems = shared_ptr<void>(::operator new(d.bytes_sz()), [](void *p) { ::operator delete(p); }); std::memset(elems.get(), 0, d.bytes_sz()); }
  }
  void *rp_elems() const { return elems ? elems.get() : rp; }
  uint64_t elems_sz() const { return dims.dims_prod(); }
  bool has_data() const { return rp_elems() != nullptr; }
};
typedef shared_ptr<nda_t> p_nda_t;
typedef map<string, p_nda_t> map_str_p_nda_t;
typedef nda_t nda_float_t;  // the fwd map of has_conv_fwd_t holds float ndas
typedef shared_ptr<nda_float_t> p_nda_float_t;
typedef map<string, p_nda_float_t> map_str_p_nda_float_t;
typedef shared_ptr<map_str_p_nda_float_t> p_map_str_p_nda_float_t;

template <typename T> inline p_nda_t make_scalar_nda(T const &v, string const &tn) {
  dims_t d; d.tn = tn; d.calc_strides();
  p_nda_t r = std::make_shared<nda_t>(d);
  *static_cast<T *>(r->rp_elems()) = v;
  return r;
}
inline p_nda_t make_dims_nda(dims_t const &d) { return std::make_shared<nda_t>(d, false); }
// scalar-or-first-element read with type conversion (SNE<T> in the reference)
inline double nda_scalar_as_double(nda_t const &n) {
  if (!n.has_data()) { rt_err("nda has no value"); }
  string const &tn = n.dims.tn;
  void *p = n.rp_elems();
  if (tn == "float") { return *static_cast<float *>(p); }
  if (tn == "double") { return *static_cast<double *>(p); }
  if (tn == "uint32_t") { return *static_cast<uint32_t *>(p); }
  if (tn == "int32_t") { return *static_cast<int32_t *>(p); }
  if (tn == "uint64_t") { return static_cast<double>(*static_cast<uint64_t *>(p)); }
  rt_err("nda_scalar_as_double: unhandled type " + tn);
}

// ---- lexp -----------------------------------------------------------------------------------------------------
struct lexp_t;
typedef shared_ptr<lexp_t> p_lexp_t;
struct lexp_t {
  bool is_leaf = true;
  string leaf;
  vector<std::pair<string, p_lexp_t>> kids;
  p_lexp_t find(string const &k) const { for (auto const &kv : kids) { if (kv.first == k) { return kv.second; } } return p_lexp_t(); }
};

namespace detail {
struct lexp_parser_t {
  string const &s;
  size_t pos = 0;
  explicit lexp_parser_t(string const &s_) : s(s_) {}
  p_lexp_t parse_val() {
    p_lexp_t r = std::make_shared<lexp_t>();
    if (pos < s.size() && s[pos] == '(') {
      ++pos;
      r->is_leaf = false;
      while (true) {
        if (pos >= s.size()) { rt_err("lexp: unterminated list in '" + s + "'"); }
        if (s[pos] == ')') { ++pos; return r; }
        string k;
        while (pos < s.size() && s[pos] != '=' && s[pos] != ',' && s[pos] != '(' && s[pos] != ')') {
          if (s[pos] == '\\') { ++pos; if (pos >= s.size()) { rt_err("lexp: dangling escape"); } }
          k.push_back(s[pos++]);
        }
        if (pos >= s.size() || s[pos] != '=') { rt_err("lexp: expected '=' after key '" + k + "' in '" + s + "'"); }
        ++pos;
        if (r->find(k)) { rt_err("lexp: duplicate key '" + k + "'"); }
        r->kids.push_back(std::make_pair(k, parse_val()));
        if (pos < s.size() && s[pos] == ',') { ++pos; }
      }
    }
    while (pos < s.size() && s[pos] != ',' && s[pos] != '(' && s[pos] != ')') {
      if (s[pos] == '\\') { ++pos; if (pos >= s.size()) { rt_err("lexp: dangling escape"); } }
      r->leaf.push_back(s[pos++]);
    }
    return r;
  }
};
inline string strip(string const &s) {
  size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
  return (b == string::npos) ? string() : s.substr(b, e - b + 1);
}
}  // namespace detail

inline p_lexp_t parse_lexp(string const &s_) {
  string const s = detail::strip(s_);
  detail::lexp_parser_t p(s);
  p_lexp_t r = p.parse_val();
  if (p.pos != s.size()) { rt_err("lexp: trailing characters in '" + s + "'"); }
  return r;
}

// dims list `(name=sz,...)`, optional pseudo-dim __tn__ (src/nesi.cc:661-681)
inline dims_t dims_from_lexp(lexp_t const &l, string const &default_tn) {
  if (l.is_leaf) { rt_err("dims must be a list"); }
  dims_t d;
  d.tn = default_tn;
  for (auto const &kv : l.kids) {
    if (!kv.second->is_leaf) { rt_err("dims entry '" + kv.first + "' must be a leaf"); }
    if (kv.first == "__tn__") { d.tn = kv.second->leaf; continue; }
    d.add_dim(kv.first, static_cast<uint32_t>(std::stoul(kv.second->leaf)));
  }
  d.calc_strides();
  return d;
}

// nda `(tn=..,dims=(..),v=a:b:c)` (src/nesi.cc:720-785): tn defaults to float; v optional; scalar when no dims
inline p_nda_t nda_from_lexp(lexp_t const &l) {
  if (l.is_leaf) { rt_err("nda must be a list"); }
  string tn = "float";
  p_lexp_t dims_l, v_l;
  for (auto const &kv : l.kids) {
    if (kv.first == "tn") { tn = kv.second->leaf; }
    else if (kv.first == "dims") { dims_l = kv.second; }
    else if (kv.first == "v") { v_l = kv.second; }
    else { rt_err("nda: unused field '" + kv.first + "'"); }
  }
  dims_t d;
  if (dims_l) { d = dims_from_lexp(*dims_l, tn); } else { d.tn = tn; d.calc_strides(); }
  if (!v_l) { return make_dims_nda(d); }
  vect_string parts;
  { std::stringstream ss(v_l->leaf); string item; while (std::getline(ss, item, ':')) { parts.push_back(item); } }
  if (parts.size() != d.dims_prod()) { rt_err("nda: value count " + str(parts.size()) + " != dims_prod " + str(d.dims_prod())); }
  p_nda_t r = std::make_shared<nda_t>(d);
  for (size_t i = 0; i < parts.size(); ++i) {
    if (d.tn == "float") { static_cast<float *>(r->rp_elems())[i] = std::stof(parts[i]); }
    else if (d.tn == "double") { static_cast<double *>(r->rp_elems())[i] = std::stod(parts[i]); }
    else if (d.tn == "uint32_t") { static_cast<uint32_t *>(r->rp_elems())[i] = static_cast<uint32_t>(std::stoul(parts[i])); }
    else if (d.tn == "int32_t") { static_cast<int32_t *>(r->rp_elems())[i] = static_cast<int32_t>(std::stol(parts[i])); }
    else { rt_err("nda: values of type '" + d.tn + "' unsupported"); }
  }
  return r;
}

inline string nda_param_str(nda_t const &n) {
  string r = "(";
  bool need_comma = false;
  if (n.dims.tn != "float" || n.dims.empty()) { r += "tn=" + n.dims.tn; need_comma = true; }
  if (!n.dims.empty()) { r += string(need_comma ? "," : "") + "dims=" + n.dims.dims_list_str(); need_comma = true; }
  if (n.has_data() && n.dims.tn != "none") {
    r += string(need_comma ? "," : "") + "v=";
    for (uint64_t i = 0; i < n.elems_sz(); ++i) {
      if (i) { r += ":"; }
      if (n.dims.tn == "float") { r += str(static_cast<float *>(n.rp_elems())[i]); }
      else if (n.dims.tn == "uint32_t") { r += str(static_cast<uint32_t *>(n.rp_elems())[i]); }
      else if (n.dims.tn == "int32_t") { r += str(static_cast<int32_t *>(n.rp_elems())[i]); }
      else if (n.dims.tn == "double") { r += str(static_cast<double *>(n.rp_elems())[i]); }
    }
  }
  return r + ")";
}

// ---- op_base_t ------------------------------------------------------------------------------------------------
struct op_base_t {
  map_str_str str_vals;
  map_str_p_nda_t nda_vals;
  op_base_t() {}
  bool has(string const &an) const { return nda_vals.count(an) != 0; }
  void set_dims(string const &an, dims_t const &dims) { if (has(an)) { rt_err("op: '" + an + "' already set"); } nda_vals[an] = make_dims_nda(dims); }
  void set(string const &an, p_nda_t const &nda) { if (has(an)) { rt_err("op: '" + an + "' already set"); } nda_vals[an] = nda; }
  void erase(string const &an) { if (!has(an)) { rt_err("op: '" + an + "' not set"); } nda_vals.erase(an); }
  void reset_dims(string const &an, dims_t const &dims) { erase(an); set_dims(an, dims); }
  dims_t const &get_dims(string const &an) const { return get(an)->dims; }
  p_nda_t const &get(string const &an) const { auto i = nda_vals.find(an); if (i == nda_vals.end()) { rt_err("op: missing nda/dims parameter '" + an + "'"); } return i->second; }
  string const &get_str(string const &an) const { auto i = str_vals.find(an); if (i == str_vals.end()) { rt_err("op: missing str parameter '" + an + "'"); } return i->second; }
  uint32_t get_u32(string const &an) const { return static_cast<uint32_t>(nda_scalar_as_double(*get(an))); }
  float get_float(string const &an) const { return static_cast<float>(nda_scalar_as_double(*get(an))); }
  void set_u32(string const &an, uint32_t const &v) { set(an, make_scalar_nda<uint32_t>(v, "uint32_t")); }
  bool has_type() const { return str_vals.count("type") != 0; }
  string const &get_type() const { return get_str("type"); }
  void set_type(string const &t) { str_vals["type"] = t; }
  bool has_func_name() const { return str_vals.count("func_name") != 0; }
  string const &get_func_name() const { return get_str("func_name"); }
  void set_func_name(string const &f) { str_vals["func_name"] = f; }
  void erase_func_name() { str_vals.erase("func_name"); }
  // {y,x} accessors for the conv params (conv_op_base_t::kern_sz()/stride()/in_pad(), src/conv_util.H:75-110)
  uint32_t yx(string const &an, string const &d, uint32_t dflt) const { return has(an) ? get_dims(an).dsz(d) : dflt; }
  string param_str() const {  // canonical op-line text (current syntax)
    string r = "(str_vals=(";
    bool first = true;
    for (auto const &kv : str_vals) { r += string(first ? "" : ",") + kv.first + "=" + kv.second; first = false; }
    r += "),nda_vals=(";
    first = true;
    for (auto const &kv : nda_vals) { r += string(first ? "" : ",") + kv.first + "=" + nda_param_str(*kv.second); first = false; }
    return r + "))";
  }
  bool operator<(op_base_t const &o) const { return param_str() < o.param_str(); }  // total order for caches (src/op_base.cc:16-23)
};
typedef shared_ptr<op_base_t> p_op_base_t;

// op-line -> op_base_t. NESI rejects unknown fields (src/nesi.cc:25-35); so do we.
inline void fill_op_base_from_lexp(op_base_t &op, lexp_t const &l, std::set<string> const &extra_ok = std::set<string>()) {
  if (l.is_leaf) { rt_err("op line must be a list"); }
  bool const stale = l.find("dims_vals") || l.find("type");
  for (auto const &kv : l.kids) {
    string const &k = kv.first;
    if (extra_ok.count(k)) { continue; }
    if (k == "str_vals") {
      if (kv.second->is_leaf) { if (!kv.second->leaf.empty()) { rt_err("str_vals must be a list"); } continue; }
      for (auto const &sv : kv.second->kids) {
        if (stale && sv.first == "out_chans") { op.nda_vals["out_chans"] = make_scalar_nda<uint32_t>(static_cast<uint32_t>(std::stoul(sv.second->leaf)), "uint32_t"); }
        else { op.str_vals[sv.first] = sv.second->leaf; }
      }
    } else if (k == "nda_vals" && !stale) {
      if (kv.second->is_leaf) { continue; }
      for (auto const &nv : kv.second->kids) { op.nda_vals[nv.first] = nda_from_lexp(*nv.second); }
    } else if (k == "type" && stale) {
      op.str_vals["type"] = kv.second->leaf;
    } else if (k == "dims_vals" && stale) {
      for (auto const &dv : kv.second->kids) {
        bool const is_param = (dv.first == "in_pad" || dv.first == "kern_sz" || dv.first == "stride");
        op.nda_vals[dv.first] = make_dims_nda(dims_from_lexp(*dv.second, is_param ? "none" : "float"));
      }
    } else {
      rt_err("op: unused field '" + k + "'");
    }
  }
}
// canonical text of an nda / op (the form the reference prints and its op-list files hold: src/boda_base.cc:392-420 for ndas -- tn omitted
// when it is float and dims are given, dims omitted for scalars, values ':'-separated -- inside NESI's "(str_vals=(..),nda_vals=(..))")
inline string nda_text(nda_t const &o) {
  std::ostringstream out;
  out << "(";
  bool const show_dims = !o.dims.empty();
  bool const show_tn = !show_dims || o.dims.tn != "float";
  if (show_tn) { out << "tn=" << o.dims.tn; }
  if (show_dims) {
    if (show_tn) { out << ","; }
    out << "dims=(";
    for (size_t i = 0; i < o.dims.size(); ++i) { out << (i ? "," : "") << o.dims[i].name << "=" << o.dims[i].sz; }
    out << ")";
  }
  if (o.has_data()) {
    out << ",v=";
    string const &tn = o.dims.tn;
    void *p = o.rp_elems();
    for (uint64_t i = 0; i < o.elems_sz(); ++i) {
      if (i) { out << ":"; }
      if (tn == "float") { out << static_cast<float *>(p)[i]; }
      else if (tn == "double") { out << static_cast<double *>(p)[i]; }
      else if (tn == "uint32_t") { out << static_cast<uint32_t *>(p)[i]; }
      else if (tn == "int32_t") { out << static_cast<int32_t *>(p)[i]; }
      else if (tn == "uint64_t") { out << static_cast<uint64_t *>(p)[i]; }
      else { rt_err("nda_text: unhandled type " + tn); }
    }
  }
  out << ")";
  return out.str();
}
inline string op_base_text(op_base_t const &op) {
  string out = "(str_vals=(";
  bool first = true;
  for (auto const &kv : op.str_vals) { out += (first ? "" : ",") + kv.first + "=" + kv.second; first = false; }
  out += "),nda_vals=(";
  first = true;
  for (auto const &kv : op.nda_vals) { out += (first ? "" : ",") + kv.first + "=" + nda_text(*kv.second); first = false; }
  return out + "))";
}
inline p_op_base_t make_p_op_base_t_from_str(string const &line) {
  p_op_base_t op = std::make_shared<op_base_t>();
  fill_op_base_from_lexp(*op, *parse_lexp(line));
  return op;
}

}  // namespace boda
