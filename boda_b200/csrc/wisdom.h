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

}  // namespace boda
