// caffe_prototxt.cu -- see caffe_prototxt.h. Host-only code (compiled by nvcc with the rest of the library).
#include "caffe_prototxt.h"
#include <cctype>

namespace boda {

namespace {

// ---- protobuf text format: message := field*, field := name ':' scalar | name [':'] '{' message '}' ------------------------------------
struct pb_msg_t;
typedef shared_ptr<pb_msg_t> p_pb_msg_t;
struct pb_field_t { string name, scalar; p_pb_msg_t msg; };  // msg set for nested messages, scalar otherwise (strings unquoted)
struct pb_msg_t {
  vector<pb_field_t> fields;
  vector<pb_field_t const *> all(string const &n) const { vector<pb_field_t const *> r; for (auto const &f : fields) { if (f.name == n) { r.push_back(&f); } } return r; }
  pb_field_t const *first(string const &n) const { for (auto const &f : fields) { if (f.name == n) { return &f; } } return nullptr; }
  bool has(string const &n) const { return first(n) != nullptr; }
  string str(string const &n, string const &dflt = string()) const { pb_field_t const *f = first(n); return (f && !f->msg) ? f->scalar : dflt; }
  p_pb_msg_t sub(string const &n) const { pb_field_t const *f = first(n); return (f && f->msg) ? f->msg : p_pb_msg_t(); }
  vect_string strs(string const &n) const { vect_string r; for (auto const *f : all(n)) { if (!f->msg) { r.push_back(f->scalar); } } return r; }
};

struct pb_parser_t {
  string const &s;
  size_t i = 0;
  int line = 1;
  explicit pb_parser_t(string const &s_) : s(s_) {}
  [[noreturn]] void err(string const &m) { rt_err("prototxt line " + std::to_string(line) + ": " + m); }
  void skip_ws() {
    while (i < s.size()) {
      char const c = s[i];
      if (c == '\n') { ++line; ++i; }
      else if (isspace((unsigned char)c) || c == ',' || c == ';') { ++i; }
      else if (c == '#') { while (i < s.size() && s[i] != '\n') { ++i; } }
      else { break; }
    }
  }
  string ident() {
    size_t const b = i;
    while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '.' || s[i] == '-' || s[i] == '+')) { ++i; }
    if (i == b) { err(string("expected a name or value, got '") + (i < s.size() ? string(1, s[i]) : string("end of input")) + "'"); }
    return s.substr(b, i - b);
  }
  string quoted() {
    char const q = s[i++];
    string r;
    while (true) {
      if (i >= s.size()) { err("unterminated string"); }
      char const c = s[i++];
      if (c == q) { break; }
      if (c == '\\' && i < s.size()) { char const e = s[i++]; r += (e == 'n') ? '\n' : (e == 't') ? '\t' : e; }
      else { if (c == '\n') { ++line; } r += c; }
    }
    return r;
  }
  p_pb_msg_t message(bool top) {
    p_pb_msg_t m = std::make_shared<pb_msg_t>();
    while (true) {
      skip_ws();
      if (i >= s.size()) { if (!top) { err("missing '}'"); } break; }
      if (s[i] == '}') { if (top) { err("unmatched '}'"); } ++i; break; }
      pb_field_t f;
      f.name = ident();
      skip_ws();
      bool const colon = (i < s.size() && s[i] == ':');
      if (colon) { ++i; skip_ws(); }
      if (i < s.size() && (s[i] == '{' || s[i] == '<')) { ++i; f.msg = message(false); }
      else if (!colon) { err("expected ':' or '{' after '" + f.name + "'"); }
      else if (i < s.size() && (s[i] == '"' || s[i] == '\'')) { f.scalar = quoted(); }
      else { f.scalar = ident(); }
      m->fields.push_back(f);
    }
    return m;
  }
};

uint32_t to_u32(string const &v, string const &what) {
  try { size_t pos = 0; unsigned long const r = std::stoul(v, &pos); if (pos != v.size()) { throw std::invalid_argument(v); } return (uint32_t)r; }
  catch (std::exception const &) { rt_err("prototxt: '" + what + "' wants an unsigned integer, got '" + v + "'"); }
}

// NetStateRule matching for NetState{phase: TEST} (layer_included_for_state): a layer with include rules needs one that matches; without
// include rules it is kept unless an exclude rule matches. Only `phase` is interpreted (the nets use nothing else).
bool rule_matches_test(pb_msg_t const &rule) { return !rule.has("phase") || rule.str("phase") == "TEST"; }
bool layer_included_for_test(pb_msg_t const &l) {
  auto const inc = l.all("include"), exc = l.all("exclude");
  if (!inc.empty()) { for (auto const *r : inc) { if (r->msg && rule_matches_test(*r->msg)) { return true; } } return false; }
  for (auto const *r : exc) { if (r->msg && rule_matches_test(*r->msg)) { return false; } }
  return true;
}

// V1 `layers { type: CONVOLUTION }` enum names -> layer type strings (Caffe's upgrade_proto UpgradeV1LayerType)
string norm_type(string const &t) {
  static map<string, string> const v1{{"CONVOLUTION", "Convolution"}, {"RELU", "ReLU"}, {"POOLING", "Pooling"}, {"LRN", "LRN"}, {"DROPOUT", "Dropout"},
                                      {"CONCAT", "Concat"}, {"SOFTMAX", "Softmax"}, {"SOFTMAX_LOSS", "SoftmaxWithLoss"}, {"ACCURACY", "Accuracy"},
                                      {"DATA", "Data"}, {"INNER_PRODUCT", "InnerProduct"}, {"ELTWISE", "Eltwise"}, {"SPLIT", "Split"}};
  auto i = v1.find(t);
  return i == v1.end() ? t : i->second;
}

string yx_nda(string const &name, uint32_t y, uint32_t x) { return name + "=(tn=none,dims=(y=" + std::to_string(y) + ",x=" + std::to_string(x) + "))"; }

// kernel / stride / pad of a ConvolutionParameter (repeated fields: 1 value = both axes, 2 = y,x; or the _h/_w pair) -- fill_in_conv_op_from_param
void conv_geom(pb_msg_t const &p, string const &rep, string const &h, string const &w, string const &nda_name, vect_string &ndas, string const &tag) {
  vect_string const v = p.strs(rep);
  if (p.has(h) || p.has(w)) {
    if (!(p.has(h) && p.has(w)) || !v.empty()) { rt_err("layer '" + tag + "': give either " + rep + " or both " + h + " and " + w); }
    ndas.push_back(yx_nda(nda_name, to_u32(p.str(h), h), to_u32(p.str(w), w)));
  } else if (v.size() == 1) { ndas.push_back(yx_nda(nda_name, to_u32(v[0], rep), to_u32(v[0], rep))); }
  else if (v.size() == 2) { ndas.push_back(yx_nda(nda_name, to_u32(v[0], rep), to_u32(v[1], rep))); }
  else if (!v.empty()) { rt_err("layer '" + tag + "': saw " + std::to_string(v.size()) + " " + rep + " values; use 1 or 2 (N-d convolutions are not supported)"); }
}

string join(vect_string const &v, char const *sep) { string r; for (size_t i = 0; i < v.size(); ++i) { r += (i ? sep : "") + v[i]; } return r; }

}  // namespace

string conv_pipe_text_from_prototxt(string const &prototxt, prototxt_opts_t const &opts) {
  pb_parser_t parser(prototxt);
  p_pb_msg_t net = parser.message(true);
  string out;
  map<string, uint32_t> unused_overrides = opts.in_dims;
  auto add_source = [&](string const &name, uint32_t d[4]) {
    char const *const names[4] = {"img", "chan", "y", "x"};
    for (int i = 0; i < 4; ++i) { auto o = opts.in_dims.find(names[i]); if (o != opts.in_dims.end()) { d[i] = o->second; unused_overrides.erase(names[i]); } }
    out += "(node=" + name + ",dims=(img=" + std::to_string(d[0]) + ",chan=" + std::to_string(d[1]) + ",y=" + std::to_string(d[2]) + ",x=" + std::to_string(d[3]) + "))\n";
  };
  // old-style top-level input blobs (src/caffepb.cc:178-201): `input_dim` x4 per blob, or one `input_shape { dim ... }` per blob
  vect_string const inputs = net->strs("input"), input_dims = net->strs("input_dim");
  auto const input_shapes = net->all("input_shape");
  if (!inputs.empty()) {
    bool const use_shape = (inputs.size() == input_shapes.size()) && input_dims.empty();
    if (!use_shape && !(inputs.size() * 4 == input_dims.size() && input_shapes.empty())) {
      rt_err("prototxt: top-level input blobs need either one input_shape each or four input_dim values each");
    }
    for (size_t b = 0; b < inputs.size(); ++b) {
      uint32_t d[4];
      if (use_shape) {
        vect_string const dims = input_shapes[b]->msg ? input_shapes[b]->msg->strs("dim") : vect_string();
        if (dims.size() != 4) { rt_err("prototxt: input blob '" + inputs[b] + "' does not have 4 dims"); }
        for (int i = 0; i < 4; ++i) { d[i] = to_u32(dims[i], "dim"); }
      } else { for (int i = 0; i < 4; ++i) { d[i] = to_u32(input_dims[4 * b + i], "input_dim"); } }
      add_source(inputs[b], d);
    }
  }
  vector<pb_field_t const *> layers = net->all("layer");
  if (layers.empty()) { layers = net->all("layers"); }  // V1 spelling
  if (layers.empty()) { rt_err("prototxt: no layer { } entries"); }
  bool found_out = false;
  for (auto const *lf : layers) {
    if (!lf->msg) { rt_err("prototxt: 'layer' must be a message"); }
    pb_msg_t const &l = *lf->msg;
    if (!l.has("name") || !l.has("type")) { rt_err("prototxt: layer without name or type"); }
    if (!layer_included_for_test(l)) { continue; }
    string const tag = l.str("name"), type = norm_type(l.str("type"));
    vect_string const bots = l.strs("bottom"), tops = l.strs("top");
    string op_type = type;
    vect_string ndas;
    bool emit = true;
    if (type == "Convolution") {
      p_pb_msg_t p = l.sub("convolution_param");
      if (!p) { rt_err("layer '" + tag + "': Convolution without convolution_param"); }
      if (p->has("group") && to_u32(p->str("group"), "group") != 1) { rt_err("layer '" + tag + "': grouped convolutions are not supported (nor by the reference)"); }
      if (p->has("dilation") && to_u32(p->str("dilation"), "dilation") != 1) { rt_err("layer '" + tag + "': dilated convolutions are not supported"); }
      conv_geom(*p, "kernel_size", "kernel_h", "kernel_w", "kern_sz", ndas, tag);
      if (ndas.empty()) { rt_err("layer '" + tag + "': Convolution needs a kernel size"); }
      conv_geom(*p, "stride", "stride_h", "stride_w", "stride", ndas, tag);
      conv_geom(*p, "pad", "pad_h", "pad_w", "in_pad", ndas, tag);
      ndas.push_back("out_chans=(tn=uint32_t,v=" + std::to_string(to_u32(p->str("num_output"), "num_output")) + ")");
      if (p->str("bias_term", "true") == "false") { ndas.push_back("bias_term=(tn=uint32_t,v=0)"); }
    } else if (type == "InnerProduct") {
      p_pb_msg_t p = l.sub("inner_product_param");
      if (!p) { rt_err("layer '" + tag + "': InnerProduct without inner_product_param"); }
      ndas.push_back("out_chans=(tn=uint32_t,v=" + std::to_string(to_u32(p->str("num_output"), "num_output")) + ")");
      if (p->str("bias_term", "true") == "false") { ndas.push_back("bias_term=(tn=uint32_t,v=0)"); }
    } else if (type == "Pooling") {
      p_pb_msg_t p = l.sub("pooling_param");
      if (!p) { rt_err("layer '" + tag + "': Pooling without pooling_param"); }
      if (p->has("round_mode") && p->str("round_mode") != "CEIL") { rt_err("layer '" + tag + "': only the ceil output-size rule is supported (src/conv_util.cc:198-204), got round_mode " + p->str("round_mode")); }
      string const method = p->str("pool", "MAX");
      if (method != "MAX" && method != "AVE") { rt_err("layer '" + tag + "': unhandled pooling method " + method); }
      ndas.push_back(string("avg_pool=(tn=uint32_t,v=") + (method == "AVE" ? "1" : "0") + ")");
      bool const global = (p->str("global_pooling", "false") == "true");
      vect_string k;
      conv_geom(*p, "kernel_size", "kernel_h", "kernel_w", "kern_sz", k, tag);
      if (global == !k.empty()) { rt_err("layer '" + tag + "': global pooling iff no kernel size (src/caffepb.cc:274)"); }
      if (!global) {
        ndas.push_back(k[0]);
        size_t const n0 = ndas.size();
        conv_geom(*p, "stride", "stride_h", "stride_w", "stride", ndas, tag);
        if (ndas.size() == n0) { ndas.push_back(yx_nda("stride", 1, 1)); }
        size_t const n1 = ndas.size();
        conv_geom(*p, "pad", "pad_h", "pad_w", "in_pad", ndas, tag);
        if (ndas.size() == n1) { ndas.push_back(yx_nda("in_pad", 0, 0)); }
      }
    } else if (type == "LRN") {
      p_pb_msg_t p = l.sub("lrn_param");
      pb_msg_t const empty;
      pb_msg_t const &pp = p ? *p : empty;
      if (pp.str("norm_region", "ACROSS_CHANNELS") != "ACROSS_CHANNELS") { rt_err("layer '" + tag + "': only ACROSS_CHANNELS LRN is supported"); }
      ndas.push_back("local_size=(tn=uint32_t,v=" + std::to_string(to_u32(pp.str("local_size", "5"), "local_size")) + ")");
      ndas.push_back("alpha=(tn=float,v=" + pp.str("alpha", "1") + ")");
      ndas.push_back("beta=(tn=float,v=" + pp.str("beta", "0.75") + ")");
      ndas.push_back("k=(tn=float,v=" + pp.str("k", "1") + ")");
    } else if (type == "ReLU") {
      p_pb_msg_t p = l.sub("relu_param");
      if (p && p->has("negative_slope") && std::stod(p->str("negative_slope")) != 0.0) { rt_err("layer '" + tag + "': leaky ReLU (negative_slope != 0) is not supported"); }
    } else if (type == "Concat") {
      p_pb_msg_t p = l.sub("concat_param");
      if (p && ((p->has("axis") && p->str("axis") != "1") || (p->has("concat_dim") && p->str("concat_dim") != "1"))) { rt_err("layer '" + tag + "': only channel (axis 1) Concat is supported"); }
    } else if (type == "Eltwise") {
      p_pb_msg_t p = l.sub("eltwise_param");
      if (p && p->str("operation", "SUM") != "SUM") { rt_err("layer '" + tag + "': only Eltwise SUM is supported"); }
      if (p && p->has("coeff")) { rt_err("layer '" + tag + "': Eltwise coefficients are not supported"); }
    } else if (type == "BatchNorm") {
      p_pb_msg_t p = l.sub("batch_norm_param");
      if (p && p->str("use_global_stats", "true") != "true") { rt_err("layer '" + tag + "': BatchNorm needs use_global_stats (inference)"); }
      ndas.push_back("eps=(tn=float,v=" + (p ? p->str("eps", "1e-5") : string("1e-5")) + ")");
    } else if (type == "Scale") {
      p_pb_msg_t p = l.sub("scale_param");
      if (!p || p->str("bias_term", "false") != "true") { rt_err("layer '" + tag + "': Scale without bias_term is not supported"); }
      if ((p->has("axis") && p->str("axis") != "1") || (p->has("num_axes") && p->str("num_axes") != "1")) { rt_err("layer '" + tag + "': only per-channel Scale (axis 1, num_axes 1) is supported"); }
    } else if (type == "Dropout") {
      if (tops != bots) { rt_err("layer '" + tag + "': a Dropout that is not in place becomes `clone`, which rtc_fwd cannot run (src/caffepb.cc:235-238)"); }
      p_pb_msg_t p = l.sub("dropout_param");
      ndas.push_back("dropout_ratio=(tn=float,v=" + (p ? p->str("dropout_ratio", "0.5") : string("0.5")) + ")");
    } else if (type == "Softmax") {
      if (!opts.keep_softmax) { emit = false; }  // the reference ignores Softmax layers in forward graphs (src/caffepb.cc:250-254)
    } else if (type == "SoftmaxWithLoss" || type == "Accuracy") {
      emit = false;
    } else if (type == "Data") {  // top(0) becomes a source node {batch_size, 3, crop_size, crop_size}; the label top is dropped (src/caffepb.cc:281-306)
      p_pb_msg_t dp = l.sub("data_param"), tp = l.sub("transform_param");
      if (!dp || !tp) { rt_err("layer '" + tag + "': Data layer needs data_param and transform_param"); }
      if (!bots.empty() || tops.size() != 2) { rt_err("layer '" + tag + "': unhandled Data layer (wants no bottoms and two tops)"); }
      uint32_t d[4] = {to_u32(dp->str("batch_size"), "batch_size"), 3, to_u32(tp->str("crop_size"), "crop_size"), to_u32(tp->str("crop_size"), "crop_size")};
      add_source(tops[0], d);
      emit = false;
    } else {
      rt_err("layer '" + tag + "': unsupported layer type '" + type + "'");
    }
    bool has_out = false;
    for (auto const &t : tops) { if (!opts.out_node_name.empty() && t == opts.out_node_name) { has_out = true; found_out = true; } }
    if (found_out && !has_out) { break; }  // layers after the one(s) producing out_node_name are not read (src/caffepb.cc:311-315)
    if (!emit) { continue; }
    if (bots.empty() || tops.empty()) { rt_err("layer '" + tag + "': needs at least one bottom and one top"); }
    out += "(tag=" + tag + ",str_vals=(type=" + op_type + ")" + (ndas.empty() ? string() : ",nda_vals=(" + join(ndas, ",") + ")") + ",bots=" + join(bots, ":") +
           ",tops=" + join(tops, ":") + ")\n";
  }
  if (!unused_overrides.empty()) { rt_err("prototxt: unused/unknown dims in in_dims: " + unused_overrides.begin()->first); }
  if (!opts.out_node_name.empty() && !found_out) { rt_err("prototxt: node '" + opts.out_node_name + "' is not produced by any layer"); }
  return out;
}

}  // namespace boda
