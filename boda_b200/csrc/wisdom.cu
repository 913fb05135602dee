// wisdom.cu -- see wisdom.h. Host-only code.
#include "wisdom.h"
#include <cmath>
#include <functional>
#include <regex>
#include <limits>
#include <random>
#include <set>
#include <sstream>

namespace boda {

namespace {

// boost::random::uniform_int_distribution<uint64_t>(0, hi) driven by a 32-bit mt19937 (boost/random/uniform_int_distribution.hpp,
// generate_uniform_int): a zero range consumes nothing, a full 32-bit range is one draw, smaller ranges use bucketed rejection.
// std::mt19937 is the same engine as boost::random::mt19937; only the distribution differs from libstdc++'s, hence this restatement.
uint64_t boost_uniform_u64(std::mt19937 &gen, uint64_t hi) {
  uint64_t const brange = 0xFFFFFFFFull;
  if (hi == 0) { return 0; }
  if (hi == brange) { return gen(); }
  if (hi > brange) { rt_err("nda_digest: tensors of 2^32 or more elements are not supported (dims_t sizes are uint32_t, src/boda_base.H:608)"); }
  uint64_t bucket = brange / (hi + 1);
  if (brange % (hi + 1) == hi) { ++bucket; }
  while (true) {
    uint64_t const r = gen() / bucket;
    if (r <= hi) { return r; }
  }
}

int floor_log2_u64(uint64_t v) { int r = -1; while (v) { v >>= 1; ++r; } return r; }

template <typename T> void bw(std::string &o, T const &v) { o.append(reinterpret_cast<char const *>(&v), sizeof(T)); }
void bw_str(std::string &o, string const &s) { bw<uint32_t>(o, (uint32_t)s.size()); o += s; }

}  // namespace

string nda_digest_hex(string const &var_name, dims_t const &dims_in, float const *ve) {
  dims_t dims = dims_in;
  dims.calc_strides();
  uint64_t const sz = dims.dims_prod();
  uint64_t const seed = std::hash<std::string>()(var_name);
  // min / max (set_from_nda, src/boda_base.cc:249-256)
  float min_v = std::numeric_limits<float>::max(), max_v = std::numeric_limits<float>::lowest();
  for (uint64_t i = 0; i < sz; ++i) { if (ve[i] < min_v) { min_v = ve[i]; } if (ve[i] > max_v) { max_v = ve[i]; } }
  // sample strides: small primes <= size, every dim stride, the size itself (get_samp_strides, :221-233)
  std::set<uint64_t> strides;
  for (uint32_t p : {1u, 2u, 3u, 5u, 7u, 11u, 13u, 17u, 19u, 23u, 29u}) { if (p <= sz) { strides.insert(p); } }
  for (size_t i = 0; i < dims.size(); ++i) { strides.insert(dims[i].stride); }
  strides.insert(sz);
  std::mt19937 gen(static_cast<uint32_t>(seed));  // boost::random::mt19937 gen( seed ): the 64-bit seed is truncated to the engine's 32 bits
  vector<float> samps;
  for (uint64_t stride : strides) {
    if (!stride || stride > sz) { rt_err("nda_digest: bad sample stride"); }
    int const num_offsets = floor_log2_u64(stride + 1);
    std::set<uint64_t> seen;
    for (int k = 0; k < num_offsets; ++k) {
      uint64_t const offset = boost_uniform_u64(gen, stride - 1);
      if (!seen.insert(offset).second) { continue; }  // duplicate offsets are skipped (:243)
      float sv = 0.0f;
      for (uint32_t i = (uint32_t)offset; i < sz; i += (uint32_t)stride) { sv += ve[i]; }  // sequential float checksum, uint32_t index (:268-270)
      samps.push_back(sv);
    }
  }
  // bwrite (:329-338; string / vector / dims_t writers src/boda_base.H:319-417,728-740): SURVEY Appendix D
  std::string b;
  bw<uint8_t>(b, 1);            // non-null pointer
  bw_str(b, "float");           // element type the reader dispatches on
  bw<uint32_t>(b, 0xDADA0101u);
  bw<double>(b, 0.0);           // self_cmp_mrd
  bw<uint32_t>(b, (uint32_t)dims.size());
  for (size_t i = 0; i < dims.size(); ++i) { bw<uint32_t>(b, dims[i].sz); bw<uint32_t>(b, dims[i].stride); bw_str(b, dims[i].name); }
  bw_str(b, "float");           // dims.tn
  bw<uint64_t>(b, sz);          // strides_sz
  bw<uint8_t>(b, 1);            // strides valid
  bw<uint64_t>(b, seed);
  bw<float>(b, min_v);
  bw<float>(b, max_v);
  bw<uint32_t>(b, (uint32_t)samps.size());
  for (float s : samps) { bw<float>(b, s); }
  static char const *const hexd = "0123456789ABCDEF";
  string h;
  h.reserve(b.size() * 2);
  for (unsigned char c : b) { h += hexd[c >> 4]; h += hexd[c & 15]; }
  return h;
}

string wisdom_record_text(string const &op_text, vector<std::pair<string, string>> const &kgs, vector<wisdom_run_t> const &runs) {
  std::ostringstream out;
  out << "op_wisdom_t\n" << op_text << "\n";
  for (auto const &kg : kgs) { out << "kg\n" << kg.first << "\n" << kg.second << "\n"; }
  for (auto const &r : runs) {
    out << "op_tune_wisdom_t\n" << r.op_tune_text << "\n";
    out << "op_run_t\n" << r.be_plat_tag << "\n" << r.rt_secs << "\n" << r.err << "\n";
    if (r.err.empty()) { out << r.run_op_text << "\n"; }
    out << "/op_tune_wisdom_t\n";
  }
  out << "/op_wisdom_t\n";
  return out.str();
}

// ---- reader (src/op-tuner.cc:17-93) ----------------------------------------------------------------------------------
namespace {
struct line_reader_t {
  string const &t;
  size_t pos = 0;
  explicit line_reader_t(string const &t_) : t(t_) {}
  bool getline(string &out) {  // false at EOF
    if (pos >= t.size()) { return false; }
    size_t const e = t.find('\n', pos);
    out = t.substr(pos, (e == string::npos ? t.size() : e) - pos);
    pos = (e == string::npos) ? t.size() : e + 1;
    return true;
  }
  string must_getline() {
    string r;
    if (!getline(r)) { rt_err("error reading line-oriented text stream. expected a non-empty line, but got EOF."); }
    return r;
  }
};
double parse_secs(string const &s) {
  if (s == "nan" || s == "-nan") { return std::nan(""); }
  try { size_t n = 0; double const v = std::stod(s, &n); if (n != s.size()) { throw 0; } return v; }
  catch (...) { rt_err("can't convert '" + s + "' to double."); }
  return 0;
}
}  // namespace

vector<wis_op_t> read_wisdom_text(string const &text) {
  vector<wis_op_t> ret;
  line_reader_t in(text);
  string line;
  while (in.getline(line)) {
    if (line != "op_wisdom_t") { rt_err("error reading line-oriented text stream. expected a line with 'op_wisdom_t', but saw '" + line + "'."); }
    wis_op_t op;
    op.op_text = in.must_getline();
    make_p_op_base_t_from_str(op.op_text);  // must parse as an op (throws otherwise)
    while (true) {
      line = in.must_getline();
      if (line == "/op_wisdom_t") { break; }
      else if (line == "kg") { string const n = in.must_getline(); op.kgs.push_back({n, in.must_getline()}); }
      else if (line == "op_tune_wisdom_t") {
        wis_tune_t tune;
        tune.tune_text = in.must_getline();
        while (true) {
          line = in.must_getline();
          if (line == "/op_tune_wisdom_t") { break; }
          else if (line == "op_run_t") {
            wis_run_t r;
            r.be_plat_tag = in.must_getline();
            r.rt_secs = parse_secs(in.must_getline());
            r.err = in.must_getline();
            if (r.err.empty()) { r.op_text = in.must_getline(); }
            for (auto const &o : tune.runs) { if (o.be_plat_tag == r.be_plat_tag) { rt_err("duplicate run for platform '" + r.be_plat_tag + "' in one op_tune_wisdom_t"); } }
            tune.runs.push_back(r);
          } else { rt_err("unknown op_tune_wisdom_t text format stream command read '" + line + "'"); }
        }
        op.tunes.push_back(tune);
      } else { rt_err("unknown op_wisdom_t text format stream command read '" + line + "'"); }
    }
    ret.push_back(op);
  }
  return ret;
}

uint64_t op_text_flops(string const &op_text) {
  p_op_base_t op = make_p_op_base_t_from_str(op_text);
  if (op->has_type() && op->get_type() == "sgemm") {
    dims_t const &a = op->get_dims("a"), &b = op->get_dims("b");
    return 2ull * a.dsz("M") * b.dsz("N") * a.dsz("K");
  }
  dims_t const &dout = op->get_dims("out"), &din = op->get_dims("in"), &filts = op->get_dims("filts");
  if (din.dsz("img") != dout.dsz("img")) { rt_err("op flops: in / out disagree in img"); }
  uint64_t const M = (uint64_t)dout.dsz("img") * dout.dsz("x") * dout.dsz("y"), K = (uint64_t)filts.dsz("in_chan") * filts.dsz("x") * filts.dsz("y"), N = filts.dsz("out_chan");
  return M * N * K * 2;
}

wis_ana_res_t wis_ana(vector<wis_op_t> const &ops_in, wis_ana_opts_t const &opts) {
  std::regex const r_plat(opts.s_plat);
  struct score_t { double tot_rt_secs = 0; uint32_t tot_num = 0; };
  map<string, score_t> scores;
  // select ops, drop runs with errors / of other platforms (filter_runs)
  vector<wis_op_t> ops;
  for (auto const &o : ops_in) {
    p_op_base_t op = make_p_op_base_t_from_str(o.op_text);
    if (opts.s_img && op->get_dims("in").dsz("img") != opts.s_img) { continue; }
    if (!((double)op_text_flops(o.op_text) >= opts.min_flops)) { continue; }
    for (auto const &p : ops) { if (p.op_text == o.op_text) { rt_err("wis-ana: duplicate op in wisdom input: " + o.op_text); } }
    wis_op_t f = o;
    for (auto &t : f.tunes) {
      vector<wis_run_t> keep;
      for (auto const &r : t.runs) { if (r.err.empty() && std::regex_search(r.be_plat_tag, r_plat)) { keep.push_back(r); } }
      t.runs = keep;
    }
    ops.push_back(f);
  }
  std::sort(ops.begin(), ops.end(), [](wis_op_t const &a, wis_op_t const &b) { return a.op_text < b.op_text; });
  wis_ana_res_t res;
  double const nan = std::nan("");
  for (auto const &o : ops) {
    wis_ana_row_t row;
    row.op_text = o.op_text; row.flops = op_text_flops(o.op_text); row.aom = row.pom = row.ref = nan;
    double min_time = std::numeric_limits<double>::max();
    bool have_ref = false;
    for (auto const &t : o.tunes) {
      for (auto const &r : t.runs) {
        if (!opts.ref_tune.empty() && t.tune_text == opts.ref_tune) {
          if (have_ref) { rt_err("wis-ana: more than one reference-tune run for an op (only one-platform filters are supported)"); }
          have_ref = true; row.ref = r.rt_secs;
        } else {
          ++res.tot_runs;
          score_t &s = scores[t.tune_text]; ++s.tot_num; s.tot_rt_secs += r.rt_secs;
          if (r.rt_secs < min_time) { min_time = r.rt_secs; row.pom = r.rt_secs; row.pom_tune = t.tune_text; }
        }
      }
    }
    res.rows.push_back(row);
  }
  // best overall tune: most cases handled without error first, least total time second
  score_t const *best = nullptr;
  for (auto const &kv : scores) {
    if (!best || kv.second.tot_num > best->tot_num || (kv.second.tot_num == best->tot_num && kv.second.tot_rt_secs < best->tot_rt_secs)) { best = &kv.second; res.aom_tune = kv.first; }
  }
  for (size_t i = 0; i < ops.size(); ++i) {
    for (auto const &t : ops[i].tunes) { if (t.tune_text == res.aom_tune) { for (auto const &r : t.runs) { res.rows[i].aom = r.rt_secs; } } }
  }
  return res;
}

string wis_ana_csv(wis_ana_res_t const &res, wis_ana_opts_t const &opts) {
  auto num = [](double v) { if (std::isnan(v)) { return string("nan"); } std::ostringstream o; o << v; return o.str(); };
  string out = "OP FLOPS " + opts.aom_tag + " " + opts.pom_tag + " " + opts.ref_tag + "\n";
  for (auto const &r : res.rows) { out += r.op_text + " " + str(r.flops) + " " + num(r.aom) + " " + num(r.pom) + " " + num(r.ref) + "\n"; }
  return out;
}

}  // namespace boda
