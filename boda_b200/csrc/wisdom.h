// wisdom.h -- writer for Boda's wisdom-file records: the `nda_digest_t` known-good vectors (src/boda_base.cc:210-383,
// binary layout SURVEY Appendix D) and the text records of src/op-tuner.cc:98-130 (write_op_wisdom / write_op_tune_wisdom / write_op_run),
// so per-op results measured with be=b200 can be dropped into Boda's own ops-prof / wis-ana flow and compared with its database.
// Host-only. std::hash<std::string> supplies the digest seed exactly as in the reference (same libstdc++ algorithm).
#pragma once
#include "boda_base.h"

namespace boda {

// hex(bwrite(nda_digest_T<float>)) of a float tensor with the given dims, seeded by std::hash<string>(var_name) (src/rtc_prof.cc:297-310)
string nda_digest_hex(string const &var_name, dims_t const &dims, float const *data);

struct wisdom_run_t { string op_tune_text, be_plat_tag, err, run_op_text; double rt_secs = 0; };
// one op_wisdom_t record: op line, kg entries (name, digest hex), one op_tune_wisdom_t block per run
string wisdom_record_text(string const &op_text, vector<std::pair<string, string>> const &kgs, vector<wisdom_run_t> const &runs);

// ---- reader + analysis (the reference's wis-ana mode, src/op-tuner.cc:68-93 read_next_wisdom, :135-143 filter_runs, :204-396 wis_ana_t) ----
struct wis_run_t { string be_plat_tag, err, op_text; double rt_secs = 0; };
struct wis_tune_t { string tune_text; vector<wis_run_t> runs; };
struct wis_op_t { string op_text; vector<std::pair<string, string>> kgs; vector<wis_tune_t> tunes; };
// line-oriented text records as written by write_op_wisdom (and by wisdom_record_text); format errors throw rt_exception with the
// reference's wording ("unknown op_wisdom_t text format stream command read '...'")
vector<wis_op_t> read_wisdom_text(string const &text);

struct wis_ana_opts_t {
  uint32_t s_img = 0;       // 0 == all # of imgs; otherwise only ops with this batch size
  string s_plat = ".*";     // regex selecting the platform tag of the runs that take part
  string ref_tune;          // if non-empty: the tune whose times form the REF column (and which is excluded from the minima)
  double min_flops = 0;     // only ops with >= min_flops
  string aom_tag = "boda-manual-tune", pom_tag = "boda-autotuned", ref_tag = "REF";
};
struct wis_ana_row_t { string op_text; uint64_t flops = 0; double aom = 0, pom = 0, ref = 0; string pom_tune; };  // NaN = no such run
struct wis_ana_res_t { vector<wis_ana_row_t> rows; string aom_tune; uint64_t tot_runs = 0; };
// AOM = the time of the single best overall tune (most ops handled without error, then least total time), POM = the per-op minimum over
// all non-reference tunes, REF = the reference tune's time. Rows are ordered by op text (the reference orders by its op_base_t comparison).
wis_ana_res_t wis_ana(vector<wis_op_t> const &ops, wis_ana_opts_t const &opts);
// the csv the reference writes for wis-plot.py: header "OP FLOPS <aom_tag> <pom_tag> <ref_tag>", then one row per op
string wis_ana_csv(wis_ana_res_t const &res, wis_ana_opts_t const &opts);
// algorithmic FLOPs of a Convolution / sgemm op text: 2*M*N*K (src/op-tuner.cc:242-264, src/latex-util.H:116-133)
uint64_t op_text_flops(string const &op_text);

}  // namespace boda
