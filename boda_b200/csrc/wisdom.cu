// wisdom.cu -- see wisdom.h. Host-only code.
#include "wisdom.h"
#include <functional>
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

}  // namespace boda
