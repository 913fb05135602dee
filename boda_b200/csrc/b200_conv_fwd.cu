// b200_conv_fwd.cu -- see b200_conv_fwd.h. Graph walk follows conv_pipe_fwd_t::init / gen_ops_rec / gen_op
// (src/rtc_fwd.cc:469-527, :436-465, :263-405); every node is an fp32 NCHW var as in the reference, so
// test_compute-style per-node comparison works on any node (src/test_compute.cc:165-169).
#include "b200_conv_fwd.h"
#include <set>
#include <algorithm>

namespace boda {

#define CU_CHK(x) do { cudaError_t const e_ = (x); if (e_ != cudaSuccess) { rt_err(string("CUDA error: ") + cudaGetErrorString(e_) + " in " #x " at " + __FILE__ + ":" + std::to_string(__LINE__)); } } while (0)

namespace {
vect_string split_colon(string const &s) {
  vect_string r;
  std::stringstream ss(s);
  string item;
  while (std::getline(ss, item, ':')) { if (!item.empty()) { r.push_back(item); } }
  return r;
}
uint32_t ceil_div_u32(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
}  // namespace

p_conv_node_t conv_pipe_t::get_or_make_node(string const &n) {
  auto i = nodes.find(n);
  if (i != nodes.end()) { return i->second; }
  p_conv_node_t node = std::make_shared<conv_node_t>();
  node->name = n;
  node->dims = dims_t();
  nodes[n] = node;
  return node;
}

void conv_pipe_t::add_op_from_lexp(lexp_t const &l) {
  if (p_lexp_t nl = l.find("node")) {  // source node declaration
    p_conv_node_t node = get_or_make_node(nl->leaf);
    p_lexp_t dl = l.find("dims");
    if (!dl) { rt_err("pipe: node line needs dims"); }
    node->dims = dims_from_lexp(*dl, "float");
    data_node_names.push_back(node->name);
    return;
  }
  p_conv_op_t op = std::make_shared<conv_op_t>();
  fill_op_base_from_lexp(*op, l, {"tag", "bots", "tops"});
  p_lexp_t tl = l.find("tag"), bl = l.find("bots"), tpl = l.find("tops");
  if (!tl || !bl || !tpl) { rt_err("pipe: op line needs tag, bots and tops"); }
  op->tag = tl->leaf;
  op->bots = split_colon(bl->leaf);
  op->tops = split_colon(tpl->leaf);
  if (!op->has_type()) { rt_err("Operation has no type field; can't determine type."); }
  for (auto const &o : ops) { if (o->tag == op->tag) { rt_err("pipe: duplicate op tag '" + op->tag + "'"); } }
  op->in_place = (op->bots.size() == 1 && op->tops.size() == 1 && op->bots[0] == op->tops[0]);
  if (op->is("InnerProduct")) {  // an inner product is the convolution whose window is the whole input (the reference reads it as such, src/caffepb.cc:276-279)
    op->set_type("Convolution");
    op->set_u32("is_inner_product", 1);
  }
  if (op->is("Convolution")) {  // filts / biases are implicit parameter nodes named <tag>_filts / <tag>_biases (src/conv_util.cc:270-293)
    bool const bias_term = !op->has("bias_term") || op->get_u32("bias_term") != 0;  // Caffe convolution_param.bias_term (ResNet: false)
    if (op->bots.size() == 1) { op->bots.push_back(op->tag + "_filts"); if (bias_term) { op->bots.push_back(op->tag + "_biases"); } }
    if (op->bots.size() != (bias_term ? 3u : 2u)) { rt_err("Convolution '" + op->tag + "' needs bots in[:filts[:biases]]"); }
  }
  // BatchNorm (use_global_stats) and Scale are in-place per-channel affine ops with their own parameter nodes (Caffe blob order):
  //   BatchNorm: <tag>_mean, <tag>_var (chan), <tag>_sf (1: the moving-average scale factor); Scale: <tag>_gamma, <tag>_beta (chan)
  if (op->is("BatchNorm") && op->bots.size() == 1) { for (char const *sfx : {"_mean", "_var", "_sf"}) { op->bots.push_back(op->tag + sfx); } }
  if (op->is("Scale") && op->bots.size() == 1) { for (char const *sfx : {"_gamma", "_beta"}) { op->bots.push_back(op->tag + sfx); } }
  if (op->is("BatchNorm") || op->is("Scale")) { op->in_place = (op->tops.size() == 1 && op->bots[0] == op->tops[0]); }
  for (auto const &b : op->bots) { get_or_make_node(b)->bot_for.push_back(op->tag); }
  for (auto const &t : op->tops) {
    p_conv_node_t tn = get_or_make_node(t);
    if (op->in_place) { tn->in_place_ops.push_back(op); }
    else { if (!tn->top_for.empty()) { rt_err("pipe: node '" + t + "' has multiple writers"); } tn->top_for.push_back(op->tag); }
  }
  ops.push_back(op);
}

void conv_pipe_t::calc_dims() {
  for (auto const &op : ops) {
    if (op->in_place) {
      dims_t const &d = must_get_node(op->bots[0])->dims;
      if (d.empty()) { rt_err("pipe: in-place op '" + op->tag + "' on node without dims"); }
      if (op->is("BatchNorm") || op->is("Scale")) {  // per-channel parameter nodes
        for (size_t i = 1; i < op->bots.size(); ++i) {
          p_conv_node_t pn = must_get_node(op->bots[i]);
          bool const scalar = (op->is("BatchNorm") && i == 3);
          pn->dims = dims_t({scalar ? 1u : d.dsz("chan")}, {scalar ? "v" : "chan"}, "float");
          if (!pn->is_param) { pn->is_param = true; param_names.push_back(pn->name); }
        }
      }
      continue;
    }
    if (op->is("BatchNorm") || op->is("Scale")) { rt_err("'" + op->tag + "': only in-place BatchNorm / Scale (folded into the producing Convolution) are supported"); }
    dims_t const &din = must_get_node(op->bots[0])->dims;
    if (din.empty()) { rt_err("pipe: op '" + op->tag + "' reads node '" + op->bots[0] + "' before it has dims (ops must be in topological order)"); }
    uint32_t const N = din.dsz("img"), C = din.dsz("chan"), H = din.dsz("y"), W = din.dsz("x");
    dims_t dout;
    if (op->is("Convolution")) {
      if (op->has("is_inner_product")) { op->set_dims("kern_sz", dims_t({H, W}, {"y", "x"}, "none")); }
      uint32_t const KH = op->yx("kern_sz", "y", 0), KW = op->yx("kern_sz", "x", 0);
      if (!KH || !KW) { rt_err("Convolution '" + op->tag + "' needs kern_sz"); }
      uint32_t const sy = op->yx("stride", "y", 1), sx = op->yx("stride", "x", 1), py = op->yx("in_pad", "y", 0), px = op->yx("in_pad", "x", 0);
      uint32_t const OC = op->get_u32("out_chans");
      if (H + 2 * py < KH || W + 2 * px < KW) { rt_err("Convolution '" + op->tag + "': padded input smaller than kernel"); }
      dout = dims_t({N, OC, (H + 2 * py - KH) / sy + 1, (W + 2 * px - KW) / sx + 1}, {"img", "chan", "y", "x"}, "float");
      p_conv_node_t fn = must_get_node(op->bots[1]);
      fn->dims = dims_t({OC, C, KH, KW}, {"out_chan", "in_chan", "y", "x"}, "float");
      if (!fn->is_param) { fn->is_param = true; param_names.push_back(fn->name); }
      if (op->bots.size() > 2) {
        p_conv_node_t bn = must_get_node(op->bots[2]);
        bn->dims = dims_t({OC}, {"out_chan"}, "float");
        if (!bn->is_param) { bn->is_param = true; param_names.push_back(bn->name); }
      }
    } else if (op->is("Pooling")) {
      if (op->has("kern_sz")) {  // Caffe: any partial window makes an output (src/conv_util.cc:198-204)
        uint32_t const KH = op->yx("kern_sz", "y", 1), KW = op->yx("kern_sz", "x", 1), sy = op->yx("stride", "y", 1), sx = op->yx("stride", "x", 1);
        uint32_t const py = op->yx("in_pad", "y", 0), px = op->yx("in_pad", "x", 0);
        uint32_t const OH = (H + 2 * py < KH) ? 1 : ceil_div_u32(H + 2 * py - KH, sy) + 1, OW = (W + 2 * px < KW) ? 1 : ceil_div_u32(W + 2 * px - KW, sx) + 1;
        dout = dims_t({N, C, OH, OW}, {"img", "chan", "y", "x"}, "float");
      } else { dout = dims_t({N, C, 1, 1}, {"img", "chan", "y", "x"}, "float"); }
    } else if (op->is("LRN") || op->is("ReLU") || op->is("Dropout") || op->is("Softmax")) {
      dout = din;
    } else if (op->is("Concat")) {
      uint32_t oc = 0;
      for (auto const &b : op->bots) {
        dims_t const &d = must_get_node(b)->dims;
        if (d.empty() || d.dsz("img") != N || d.dsz("y") != H || d.dsz("x") != W) { rt_err("Concat '" + op->tag + "': inputs disagree in img/y/x"); }
        oc += d.dsz("chan");
      }
      dout = dims_t({N, oc, H, W}, {"img", "chan", "y", "x"}, "float");
    } else if (op->is("Eltwise") || op->is("Reduce")) {
      for (auto const &b : op->bots) { if (!(must_get_node(b)->dims == din)) { rt_err("Eltwise '" + op->tag + "': input dims differ"); } }
      dout = din;
    } else {
      rt_err("calc_dims: unhandled op of type: " + op->get_type());
    }
    for (auto const &t : op->tops) { must_get_node(t)->dims = dout; }
  }
}

uint64_t conv_pipe_t::total_conv_flops() const {  // 2*B*OC*OH*OW*IC*KH*KW (src/latex-util.H:116-120)
  uint64_t fl = 0;
  for (auto const &op : ops) {
    if (!op->is("Convolution")) { continue; }
    dims_t const &o = must_get_node(op->tops[0])->dims, &f = must_get_node(op->bots[1])->dims;
    fl += 2ull * o.dims_prod() * f.dsz("in_chan") * f.dsz("y") * f.dsz("x");
  }
  return fl;
}

p_conv_pipe_t make_conv_pipe_from_text(string const &pipe_text) {
  p_conv_pipe_t cp = std::make_shared<conv_pipe_t>();
  std::stringstream ss(pipe_text);
  string line;
  while (std::getline(ss, line)) {
    line = detail::strip(line);
    if (line.empty() || line[0] == '#') { continue; }
    cp->add_op_from_lexp(*parse_lexp(line));
  }
  if (cp->data_node_names.empty()) { rt_err("pipe: no source (node=...) lines"); }
  cp->calc_dims();
  return cp;
}

// ---- b200_conv_fwd_t ------------------------------------------------------------------------------------------
b200_conv_fwd_t::b200_conv_fwd_t() {}
b200_conv_fwd_t::~b200_conv_fwd_t() {
  if (rtc && rtc->stream()) { cudaStreamSynchronize(rtc->stream()); }
  if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
  for (auto &sl : slots) { for (auto &kv : sl.staging) { cudaFree(kv.second); } if (sl.h2d_done) { cudaEventDestroy(sl.h2d_done); } if (sl.freed) { cudaEventDestroy(sl.freed); } }
  for (auto &e : ticket_ev) { if (e) { cudaEventDestroy(e); } }
  if (flush_buf) { cudaFree(flush_buf); }
  if (graph_exec) { cudaGraphExecDestroy(graph_exec); }
  if (graph) { cudaGraphDestroy(graph); }
}

string conv_pipe_op_sigs_text(conv_pipe_t const &cp) {
  std::set<string> sigs;
  for (auto const &op : cp.ops) {
    if (!op->is("Convolution") || op->has("is_inner_product")) { continue; }
    op_base_t sig;
    sig.str_vals = op->str_vals;
    sig.nda_vals = op->nda_vals;
    dims_t const &fd = cp.must_get_node(op->bots[1])->dims;
    for (char const *an : {"in", "filts", "biases", "out"}) { if (sig.has(an)) { sig.erase(an); } }
    sig.set_dims("in", cp.must_get_node(op->bots[0])->dims);
    sig.set_dims("filts", fd);
    if (op->bots.size() > 2) { sig.set_dims("biases", dims_t({fd.dsz("out_chan")}, {"out_chan"}, "float")); }
    sig.set_dims("out", cp.must_get_node(op->tops[0])->dims);
    sigs.insert(op_base_text(sig));
  }
  string out;
  for (auto const &l : sigs) { out += l + "\n"; }
  return out;
}

void b200_conv_fwd_t::add_call(string const &fn_base, conv_op_t const &op, op_base_t const &fop, map_str_rtc_arg_t const &args) {
  fwd_call_t c;
  c.tag = op.tag;
  c.func_name = fn_base + "__" + op.tag + "__" + str(fwd_calls.size());
  rtc_func_info_t fi;
  fi.func_name = c.func_name;
  fi.op = fop;
  fi.op.set_func_name(fn_base);
  rtc->compile({fi}, rtc_compile_opts_t());
  c.rfc.rtc_func_name = c.func_name;
  c.rfc.arg_map = args;
  fwd_calls.push_back(c);
  double fl = 0.0;  // algorithmic FLOPs: 2*B*OC*OH*OW*IC*KH*KW for conv (src/latex-util.H:116-120); bandwidth ops count none
  if (fn_base == "conv") { dims_t const &o = fop.get_dims("out"), &f = fop.get_dims("filts"); fl = 2.0 * o.dims_prod() * f.dsz("in_chan") * f.dsz("y") * f.dsz("x"); }
  call_flops.push_back(fl);
}

static char const *const absmax_cells_vn = "__absmax_cells";
void b200_conv_fwd_t::add_absmax_args(map_str_rtc_arg_t &args, string const &which, string const &node) {
  auto i = absmax_ix.find(node);
  if (i == absmax_ix.end()) { return; }
  args[which + "_absmax_cells"] = rtc_arg_t(string(absmax_cells_vn));
  args[which + "_absmax_ix"] = rtc_arg_t(make_scalar_nda<uint32_t>(i->second, "uint32_t"));
}

op_base_t b200_conv_fwd_t::conv_fop(conv_op_t const &op) const {
  op_base_t fop;
  fop.str_vals = op.str_vals;
  fop.nda_vals = op.nda_vals;
  dims_t const &fd = cp->must_get_node(op.bots[1])->dims;
  fop.set_dims("in", cp->must_get_node(op.bots[0])->dims);
  fop.set_dims("filts", fd);
  fop.set_dims("biases", dims_t({fd.dsz("out_chan")}, {"out_chan"}, "float"));
  fop.set_dims("out", cp->must_get_node(op.tops[0])->dims);
  return fop;
}

// A destination node (a convolution's own output, or the Concat output its channels are written into) gets its NHWC plane from its
// producers when: some Convolution with more than 8 input channels reads it (the row-merged small-channel path packs differently); every
// producer is a convolution that can write the plane (b200_compute_t::conv_plane_writable) at a channel offset that is a multiple of 8;
// for a Concat, every input is written in place (else the plane would have holes); and nothing modifies the node afterwards except the
// ops the convolution already absorbs (folded BatchNorm / Scale, fused ReLU) or Dropout (identity).
bool b200_conv_fwd_t::dst_plane_by_producers(string const &dst) {
  if (!pack_by_producers) { return false; }
  p_conv_node_t dn = cp->must_get_node(dst);
  bool feeds = false;
  for (auto const &o : cp->ops) { if (o->is("Convolution") && !o->bots.empty() && o->bots[0] == dst && dn->dims.dsz("chan") > 8) { feeds = true; } }
  if (!feeds) { return false; }
  vector<p_conv_op_t> producers;
  bool is_cat = false;
  for (auto const &o : cp->ops) {
    if (o->is("Concat") && o->tops[0] == dst) {
      is_cat = true;
      for (auto const &b : o->bots) {
        auto al = concat_alias.find(b);
        if (al == concat_alias.end() || (al->second.ocix % 8) != 0) { return false; }
        for (auto const &w : cp->ops) { if (w->is("Convolution") && w->tops[0] == b) { producers.push_back(w); } }
      }
    }
  }
  if (!is_cat) {
    for (auto const &w : cp->ops) {
      if (!w->is("Convolution")) { continue; }
      auto rf = res_fuse.find(w->tag);  // a convolution with a residual join in its epilogue produces the join's output node
      if (rf != res_fuse.end() ? rf->second.out_node == dst : w->tops[0] == dst) { producers.push_back(w); }
    }
  }
  if (producers.empty()) { return false; }
  for (auto const &w : producers) { if (!rtc->conv_plane_writable(conv_fop(*w), is_cat)) { return false; } }
  for (auto const &ip : dn->in_place_ops) {
    if (ip->is("Dropout") || ip->is("BatchNorm") || ip->is("Scale")) { continue; }
    if (ip->is("ReLU") && !is_cat && ip == dn->in_place_ops[0]) { continue; }  // the one the convolution fuses
    if (ip->is("ReLU") && !is_cat) {  // ReLU right after the folded BatchNorm / Scale ops is fused too
      size_t k = 0;
      while (k < dn->in_place_ops.size() && (dn->in_place_ops[k]->is("BatchNorm") || dn->in_place_ops[k]->is("Scale"))) { ++k; }
      if (k < dn->in_place_ops.size() && dn->in_place_ops[k] == ip) { continue; }
    }
    return false;
  }
  return true;
}

// The layout a destination node's producer-written planes take: the shared-padding layout (py, px) of igemm4.cuh's halo mode when EVERY
// Convolution reading the node runs in that mode with the same padding, else plain pixel-major NHWC (a halo-mode reader then packs for itself).
bool b200_conv_fwd_t::dst_plane_pad(string const &dst, int &py, int &px) {
  bool any = false;
  for (auto const &o : cp->ops) {
    if (!o->is("Convolution") || o->bots.empty() || o->bots[0] != dst) { continue; }
    int y = 0, x = 0;
    if (!rtc->conv_halo_pad(conv_fop(*o), y, x)) { return false; }
    if (any && (y != py || x != px)) { return false; }
    py = y; px = x; any = true;
  }
  return any;
}
void b200_conv_fwd_t::add_out_pack_args(map_str_rtc_arg_t &args, string const &dst) {
  args["out_pack"] = rtc_arg_t(make_scalar_nda<uint32_t>(1, "uint32_t"));
  int py = 0, px = 0;
  if (dst_plane_pad(dst, py, px)) {
    args["out_pack_py"] = rtc_arg_t(make_scalar_nda<uint32_t>((uint32_t)py, "uint32_t"));
    args["out_pack_px"] = rtc_arg_t(make_scalar_nda<uint32_t>((uint32_t)px, "uint32_t"));
  }
}

void b200_conv_fwd_t::gen_op(p_conv_op_t const &op) {
  if (op->fused) { return; }  // folded into its producer (src/rtc_fwd.cc:266)
  op_base_t fop;              // function signature: op params + the dims of every argument (conv_op_t::set_arg_dims_and_map_from_pipe)
  fop.str_vals = op->str_vals;
  fop.nda_vals = op->nda_vals;
  if (op->is("Convolution")) {
    dims_t const &fd = cp->must_get_node(op->bots[1])->dims;
    dims_t const bd({fd.dsz("out_chan")}, {"out_chan"}, "float");
    fop.set_dims("in", cp->must_get_node(op->bots[0])->dims);
    fop.set_dims("filts", fd);
    fop.set_dims("biases", bd);
    fop.set_dims("out", cp->must_get_node(op->tops[0])->dims);
    p_conv_node_t on = cp->must_get_node(op->tops[0]);
    // Leading in-place BatchNorm / Scale ops on the conv's output are per-out-channel affine maps: fold them into the filters and biases
    // (the reference only stubs these layers, src/caffepb.cc:231-232). The fold runs on the device whenever the
    // parameters change (prep_calls), writing <tag>_filts__folded / <tag>_biases__folded, which the convolution then reads.
    size_t n_aff = 0;
    p_conv_op_t bn_op, sc_op;
    while (n_aff < on->in_place_ops.size() && (on->in_place_ops[n_aff]->is("BatchNorm") || on->in_place_ops[n_aff]->is("Scale"))) {
      p_conv_op_t const &a = on->in_place_ops[n_aff];
      if (a->is("BatchNorm")) { if (bn_op || sc_op) { break; } bn_op = a; } else { if (sc_op) { break; } sc_op = a; }
      a->fused = true;
      ++n_aff;
    }
    string filts_vn = op->bots[1], biases_vn = op->bots.size() > 2 ? op->bots[2] : string();
    if (bn_op || sc_op) {
      string const ff = op->tag + "_filts__folded", bf = op->tag + "_biases__folded";
      rtc->create_var_with_dims(ff, fd);
      rtc->create_var_with_dims(bf, bd);
      op_base_t pop;
      pop.set_type("bn_fold");
      pop.set_dims("filts", fd);
      float const eps = (bn_op && bn_op->has("eps")) ? (float)nda_scalar_as_double(*bn_op->get("eps")) : 1e-5f;  // Caffe batch_norm_param.eps default
      pop.set("eps", make_scalar_nda<float>(eps, "float"));
      map_str_rtc_arg_t pargs{{"filts", filts_vn}, {"out_filts", ff}, {"out_biases", bf}};
      if (!biases_vn.empty()) { pargs["biases"] = biases_vn; }
      if (bn_op) { pargs["mean"] = bn_op->bots[1]; pargs["var"] = bn_op->bots[2]; pargs["sf"] = bn_op->bots[3]; }
      if (sc_op) { pargs["gamma"] = sc_op->bots[1]; pargs["beta"] = sc_op->bots[2]; }
      fwd_call_t pc;
      pc.tag = op->tag;
      pc.func_name = "bn_fold__" + op->tag;
      rtc_func_info_t fi;
      fi.func_name = pc.func_name;
      fi.op = pop;
      fi.op.set_func_name("bn_fold");
      rtc->compile({fi}, rtc_compile_opts_t());
      pc.rfc.rtc_func_name = pc.func_name;
      pc.rfc.arg_map = pargs;
      prep_calls.push_back(pc);
      filts_vn = ff; biases_vn = bf;
    }
    // conv+ReLU fusion: only if ReLU is the FIRST remaining in-place op on the conv's output node (src/rtc_fwd.cc:486-494)
    bool relu = false;
    if (on->in_place_ops.size() > n_aff && on->in_place_ops[n_aff]->is("ReLU")) { relu = true; on->in_place_ops[n_aff]->fused = true; }
    fop.set_u32("conv_has_relu", relu ? 1 : 0);
    map_str_rtc_arg_t args{{"in", op->bots[0]}, {"filts", filts_vn}, {"out", op->tops[0]}};
    if (!biases_vn.empty()) { args["biases"] = biases_vn; }
    add_absmax_args(args, "in", op->bots[0]);
    auto rf = res_fuse.find(op->tag);
    if (rf != res_fuse.end()) {
      // Residual join in the epilogue: this convolution adds the join's other input and writes the Eltwise output
      // directly, with the join's ReLU. Its own output node is bypassed; a plain call that computes it is kept for readers of that node.
      fwd_call_t pc;
      pc.tag = op->tag;
      pc.func_name = "conv__" + op->tag + "__plain";
      rtc_func_info_t fi;
      fi.func_name = pc.func_name;
      fi.op = fop;
      fi.op.set_func_name("conv");
      rtc->compile({fi}, rtc_compile_opts_t());
      pc.rfc.rtc_func_name = pc.func_name;
      pc.rfc.arg_map = args;
      elided_calls[op->tops[0]] = pc.rfc;
      p_conv_node_t jn = cp->must_get_node(rf->second.out_node);
      bool jrelu = false;
      if (!jn->in_place_ops.empty() && jn->in_place_ops[0]->is("ReLU")) { jrelu = true; jn->in_place_ops[0]->fused = true; }
      fop.erase("conv_has_relu");
      fop.set_u32("conv_has_relu", jrelu ? 1 : 0);
      args["out"] = rtc_arg_t(rf->second.out_node);
      args["res"] = rtc_arg_t(rf->second.res_node);
      add_absmax_args(args, "out", rf->second.out_node);
      add_absmax_args(args, "res", rf->second.res_node);
      if (dst_plane_by_producers(rf->second.out_node)) { add_out_pack_args(args, rf->second.out_node); }
      add_call("conv", *op, fop, args);
      return;
    }
    auto al = concat_alias.find(op->tops[0]);
    if (al != concat_alias.end()) {  // write into the Concat output at this input's channel offset; its abs-max cell is the Concat output's
      args["out_concat"] = rtc_arg_t(al->second.cat_node);
      args["out_ocix"] = rtc_arg_t(make_scalar_nda<uint32_t>(al->second.ocix, "uint32_t"));
      add_absmax_args(args, "out", al->second.cat_node);
    } else { add_absmax_args(args, "out", op->tops[0]); }
    // layout-transform elimination (bf16 storage mode): also write the NHWC plane the consuming convolutions read (they then skip their pack)
    {
      string const dst = (al != concat_alias.end()) ? al->second.cat_node : op->tops[0];
      if (dst_plane_by_producers(dst)) { add_out_pack_args(args, dst); }
    }
    add_call("conv", *op, fop, args);
  } else if (op->is("Pooling")) {
    map_str_rtc_arg_t args{{"in", op->bots[0]}, {"out", op->tops[0]}};
    add_absmax_args(args, "out", op->tops[0]);
    string pool_in = op->bots[0];
    auto lf = lrn_fuse.find(op->tag);
    if (lf != lrn_fuse.end()) {
      // the LRN in front runs inside this pool's kernel: the pool reads the LRN's input and takes its parameters by value; a plain lrn call is
      // kept for readers of the (now bypassed) LRN node
      p_conv_op_t l;
      for (auto const &o : cp->ops) { if (o->tag == lf->second.lrn_tag) { l = o; } }
      op_base_t lop;
      lop.str_vals = l->str_vals;
      lop.nda_vals = l->nda_vals;
      rtc_func_info_t fi;
      fi.func_name = "lrn__" + l->tag + "__plain";
      fi.op = lop;
      fi.op.set_func_name("lrn");
      rtc->compile({fi}, rtc_compile_opts_t());
      rtc_func_call_t plain;
      plain.rtc_func_name = fi.func_name;
      plain.arg_map = map_str_rtc_arg_t{{"in", lf->second.in_node}, {"out", lf->second.lrn_node}};
      elided_calls[lf->second.lrn_node] = plain;
      pool_in = lf->second.in_node;
      args["in"] = rtc_arg_t(pool_in);
      args["lrn_local_size"] = rtc_arg_t(make_scalar_nda<uint32_t>(5, "uint32_t"));
      for (char const *pn : {"alpha", "beta", "k"}) {
        if (l->has(pn)) { args[string("lrn_") + pn] = rtc_arg_t(make_scalar_nda<float>((float)nda_scalar_as_double(*l->get(pn)), "float")); }
      }
    }
    {  // layout-transform elimination: the pool kernel also writes the NHWC planes of a Convolution that reads its output (no in-place ops between)
      p_conv_node_t on = cp->must_get_node(op->tops[0]);
      bool feeds = false, clean = true;
      for (auto const &o : cp->ops) { if (o->is("Convolution") && !o->bots.empty() && o->bots[0] == op->tops[0] && on->dims.dsz("chan") > 8) { feeds = true; } }
      for (auto const &ip : on->in_place_ops) { if (!ip->is("Dropout")) { clean = false; } }
      if (pack_by_producers && feeds && clean) {
        add_absmax_args(args, "in", pool_in);
        add_out_pack_args(args, op->tops[0]);
      }
    }
    add_call("pool", *op, fop, args);
  } else if (op->is("LRN")) {
    map_str_rtc_arg_t args{{"in", op->bots[0]}, {"out", op->tops[0]}};
    add_absmax_args(args, "out", op->tops[0]);
    add_call("lrn", *op, fop, args);
  } else if (op->is("ReLU")) {
    if (!op->in_place) { rt_err("ReLU '" + op->tag + "' must be in-place (src/rtc_fwd.cc:337)"); }
    add_call("relu", *op, fop, {{"inout", op->bots[0]}});
  } else if (op->is("Dropout")) {
    if (!op->in_place) { rt_err("non-in-place Dropout becomes `clone`, which rtc_fwd cannot run (src/caffepb.cc:235-238)"); }
    // test-phase forward: identity
  } else if (op->is("Softmax")) {
    add_call("softmax", *op, fop, {{"in", op->bots[0]}, {"prob", op->tops[0]}});
  } else if (op->is("Concat")) {  // one copy per input at a running channel offset (src/rtc_fwd.cc:267-280)
    uint32_t chans_out_done = 0;
    for (auto const &b : op->bots) {
      if (concat_alias.count(b)) { chans_out_done += cp->must_get_node(b)->dims.dsz("chan"); continue; }  // already written in place by its convolution
      op_base_t cop = fop;
      cop.set_u32("ocix", chans_out_done);
      map_str_rtc_arg_t args{{"in", b}, {"out", op->tops[0]}};
      add_absmax_args(args, "out", op->tops[0]);
      add_call("copy", *op, cop, args);
      chans_out_done += cp->must_get_node(b)->dims.dsz("chan");
    }
  } else if (op->is("BatchNorm") || op->is("Scale")) {
    rt_err("'" + op->tag + "': BatchNorm / Scale must directly follow (in place) the Convolution they are folded into");
  } else if (op->is("Eltwise") || op->is("Reduce")) {
    map_str_rtc_arg_t args{{"out", op->tops[0]}};
    for (size_t i = 0; i < op->bots.size(); ++i) { args["ins_" + str(i)] = op->bots[i]; }
    fop.set_u32("ins_num", (uint32_t)op->bots.size());
    // sum + ReLU fusion, same rule as conv + ReLU: the ReLU must be the first in-place op on the output node (ResNet residual joins)
    p_conv_node_t on = cp->must_get_node(op->tops[0]);
    bool relu = false;
    if (!on->in_place_ops.empty() && on->in_place_ops[0]->is("ReLU")) { relu = true; on->in_place_ops[0]->fused = true; }
    fop.set_u32("relu", relu ? 1 : 0);
    add_absmax_args(args, "out", op->tops[0]);
    add_call("reduce", *op, fop, args);
  } else {
    rt_err("gen_op: unhandled op of type: " + op->get_type());  // src/rtc_fwd.cc:402-404
  }
}

void b200_conv_fwd_t::init(p_conv_pipe_t const &cp_, string const &opts) {
  cp = cp_;
  rtc = std::make_shared<b200_compute_t>();
  if (!opts.empty()) {
    p_lexp_t l = parse_lexp(opts);
    if (!l->is_leaf) {
      for (auto const &kv : l->kids) {
        string const &k = kv.first, &v = kv.second->leaf;
        if (k == "use_graph") { use_graph = (uint32_t)std::stoul(v); }
        else if (k == "enable_prof") { enable_prof = (uint32_t)std::stoul(v); }
        else if (k == "concat_by_offset") { concat_by_offset = (uint32_t)std::stoul(v); }
        else if (k == "fuse_eltwise") { fuse_eltwise = (uint32_t)std::stoul(v); }
        else if (k == "fuse_lrn_pool") { fuse_lrn_pool = (uint32_t)std::stoul(v); }
        else if (k == "fuse_fc_chain") { fuse_fc_chain = (uint32_t)std::stoul(v); }
        else if (k == "pack_by_producers") { pack_by_producers = (uint32_t)std::stoul(v); }
        else if (k == "op_tune" && !kv.second->is_leaf) {  // the reference's nested form, op_tune=(k1conv=1,tconv=1,...) (src/rtc_fwd.cc:36)
          for (auto const &tk : kv.second->kids) { if (!rtc->set_option(tk.first, tk.second->is_leaf ? tk.second->leaf : string())) { rt_err("mode=b200: unused op_tune option '" + tk.first + "'"); } }
        }
        else if (rtc->set_option(k, v)) {}  // back-end options, and the reference's op_tune_t knob names (accepted, ignored)
        else { rt_err("mode=b200: unused option '" + k + "'"); }  // NESI rejects unused keys (src/nesi.cc:25-35)
      }
    }
  }
  rtc->init();
  for (auto const &kv : cp->nodes) {  // one zero-filled fp32 var per pipe node
    if (kv.second->dims.empty()) { rt_err("pipe: node '" + kv.first + "' has no dims (unused / unreachable?)"); }
    rtc->create_var_with_dims(kv.first, kv.second->dims);
  }
  // Concat by offset (the reference copies every Concat input, src/rtc_fwd.cc:267-280): a Concat input qualifies when it is
  // written by one Convolution, read by nothing but this Concat, and carries no in-place op other than the ReLU the convolution fuses.
  if (concat_by_offset) {
    for (auto const &op : cp->ops) {
      if (!op->is("Concat")) { continue; }
      uint32_t ocix = 0;
      for (auto const &b : op->bots) {
        p_conv_node_t bn = cp->must_get_node(b);
        // (a node listed twice among this Concat's bottoms has two destinations: it keeps its own var and is copied both times, as the reference does)
        bool ok = bn->top_for.size() == 1 && !concat_alias.count(b) && std::count(op->bots.begin(), op->bots.end(), b) == 1;
        for (auto const &reader : bn->bot_for) {  // readers: this Concat, or the node's own in-place ops (they list the node as their bottom)
          bool in_place_reader = false;
          for (auto const &ip : bn->in_place_ops) { if (ip->tag == reader) { in_place_reader = true; } }
          if (reader != op->tag && !in_place_reader) { ok = false; }
        }
        p_conv_op_t writer;
        if (ok) { for (auto const &o : cp->ops) { if (o->tag == bn->top_for[0]) { writer = o; } } }
        ok = ok && writer && writer->is("Convolution") && !writer->has("is_inner_product");
        if (ok) { for (auto const &ip : bn->in_place_ops) { if (!(ip->is("ReLU") && ip == bn->in_place_ops[0]) && !ip->is("Dropout")) { ok = false; } } }
        if (ok) { concat_alias[b] = concat_alias_t{op->tops[0], ocix, string()}; }
        ocix += bn->dims.dsz("chan");
      }
    }
  }
  // Residual joins: Eltwise SUM of two inputs, one of them written by a Convolution (folded BatchNorm / Scale allowed, no
  // other in-place op), read by nothing but this Eltwise, with the other input complete before that convolution runs.
  if (fuse_eltwise) {
    map<string, size_t> op_ix;
    for (size_t i = 0; i < cp->ops.size(); ++i) { op_ix[cp->ops[i]->tag] = i; }
    for (auto const &e : cp->ops) {
      if (!(e->is("Eltwise") || e->is("Reduce")) || e->bots.size() != 2 || e->bots[0] == e->bots[1] || e->in_place) { continue; }
      for (int side = 1; side >= 0 && !e->fused; --side) {
        string const &n = e->bots[side], &other = e->bots[1 - side];
        p_conv_node_t nn = cp->must_get_node(n), on = cp->must_get_node(other);
        if (nn->top_for.size() != 1 || concat_alias.count(n) || e->tops[0] == n || e->tops[0] == other) { continue; }
        bool ok = true;
        for (auto const &reader : nn->bot_for) {
          bool in_place_reader = false;
          for (auto const &ip : nn->in_place_ops) { if (ip->tag == reader) { in_place_reader = true; } }
          if (reader != e->tag && !in_place_reader) { ok = false; }
        }
        for (auto const &ip : nn->in_place_ops) { if (!(ip->is("BatchNorm") || ip->is("Scale"))) { ok = false; } }
        p_conv_op_t const &w = cp->ops[op_ix.at(nn->top_for[0])];
        ok = ok && w->is("Convolution") && !w->has("is_inner_product") && !res_fuse.count(w->tag);
        size_t const wi = op_ix.at(w->tag);
        for (auto const &t : on->top_for) { if (op_ix.at(t) > wi) { ok = false; } }
        for (auto const &ip : on->in_place_ops) { if (op_ix.at(ip->tag) > wi) { ok = false; } }
        if (ok && rtc->conv_res_fusable(conv_fop(*w))) { res_fuse[w->tag] = res_fuse_t{e->tops[0], other}; e->fused = true; }
      }
    }
  }
  // LRN in front of a max pool (AlexNet norm1 -> pool1, norm2 -> pool2; GoogLeNet norm2 -> pool2): when the LRN output is read by nothing but
  // one 3x3 / 2 max pool without padding and carries no in-place op, the pool's kernel normalises the values it stages (lrn_maxpool_kernel).
  if (fuse_lrn_pool) {
    for (auto const &l : cp->ops) {
      if (!l->is("LRN") || l->in_place || l->bots.size() != 1 || l->tops.size() != 1) { continue; }
      uint32_t const ls = l->has("local_size") ? (uint32_t)nda_scalar_as_double(*l->get("local_size")) : 5;
      p_conv_node_t ln = cp->must_get_node(l->tops[0]);
      if (ls != 5 || !ln->in_place_ops.empty() || ln->top_for.size() != 1 || ln->bot_for.size() != 1) { continue; }
      p_conv_op_t pool;
      for (auto const &o : cp->ops) { if (o->tag == ln->bot_for[0]) { pool = o; } }
      if (!pool || !pool->is("Pooling") || pool->in_place || !pool->has("kern_sz")) { continue; }
      bool const avg = pool->has("avg_pool") && nda_scalar_as_double(*pool->get("avg_pool")) != 0;
      if (avg || pool->yx("kern_sz", "y", 0) != 3 || pool->yx("kern_sz", "x", 0) != 3 || pool->yx("stride", "y", 1) != 2 || pool->yx("stride", "x", 1) != 2 ||
          pool->yx("in_pad", "y", 0) != 0 || pool->yx("in_pad", "x", 0) != 0) { continue; }
      if (pool->has("emit_out_in_yx") && nda_scalar_as_double(*pool->get("emit_out_in_yx")) != 0) { continue; }
      dims_t const &d = ln->dims;
      if ((uint64_t)d.dims_prod() * 4 > (64ull << 20)) { continue; }  // (measured: no gain over the two kernels on a 154 MB map, see lrn_maxpool_kernel)
      if ((uint64_t)16 * 9 * d.dsz("x") * 4 + (uint64_t)16 * 4 * cp->must_get_node(pool->tops[0])->dims.dsz("x") * 4 > 200 * 1024) { continue; }  // the kernel's staging must fit shared memory
      lrn_fuse[pool->tag] = lrn_fuse_t{l->tag, l->bots[0], l->tops[0]};
      l->fused = true;
    }
  }
  // abs-max side channel: a node gets a cell when its (single) writer can publish max|x| and some Convolution reads it.
  // In-place ReLU/Dropout on the node only shrink max|x|, so the published value stays a valid bound.
  for (auto const &kv : cp->nodes) {
    conv_node_t const &n = *kv.second;
    if (n.top_for.size() != 1) { continue; }
    p_conv_op_t writer;
    for (auto const &o : cp->ops) { if (o->tag == n.top_for[0]) { writer = o; } }
    if (!writer || !(writer->is("Convolution") || writer->is("Pooling") || writer->is("LRN") || writer->is("Concat") || writer->is("Eltwise") || writer->is("Reduce"))) { continue; }
    bool feeds_conv = false;
    for (auto const &o : cp->ops) {
      if (o->is("Convolution") && !o->bots.empty() && o->bots[0] == n.name) { feeds_conv = true; }
      // ... or a Pooling op whose output a Convolution reads: the pool kernel scales the planes it writes for that convolution by max|in|
      if (o->is("Pooling") && !o->bots.empty() && o->bots[0] == n.name) {
        for (auto const &o2 : cp->ops) { if (o2->is("Convolution") && !o2->bots.empty() && o2->bots[0] == o->tops[0]) { feeds_conv = true; } }
      }
    }
    for (auto const &rf : res_fuse) { if (rf.second.res_node == n.name) { feeds_conv = true; } }  // a residual input: its max bounds the join's output
    for (auto const &lf : lrn_fuse) {  // the input of an LRN that runs inside a pool kernel: that kernel scales the planes it writes by max|in| (the LRN factor is <= 1)
      if (lf.second.in_node != n.name) { continue; }
      for (auto const &o : cp->ops) {
        if (o->tag != lf.first) { continue; }
        for (auto const &o2 : cp->ops) { if (o2->is("Convolution") && !o2->bots.empty() && o2->bots[0] == o->tops[0]) { feeds_conv = true; } }
      }
    }
    if (feeds_conv) { uint32_t const ix = (uint32_t)absmax_ix.size(); absmax_ix[n.name] = ix; }
  }
  rtc->create_var_with_dims(absmax_cells_vn, dims_t({(uint32_t)std::max<size_t>(absmax_ix.size(), 1)}, {"cell"}, "uint32_t"));
  for (auto const &op : cp->ops) { gen_op(op); }
  if (fuse_fc_chain) { fuse_fc_chains(); }
  for (auto &kv : concat_alias) {  // read-back functions: node = Concat output[:, ocix : ocix + chan]
    op_base_t cop;
    cop.set_type("Concat");
    cop.set_u32("ocix", kv.second.ocix);
    cop.set_u32("reverse", 1);
    rtc_func_info_t fi;
    fi.func_name = kv.second.extract_func = "copy__extract__" + kv.first;
    fi.op = cop;
    fi.op.set_func_name("copy");
    rtc->compile({fi}, rtc_compile_opts_t());
  }
  info_log = "mode=b200 plat=" + rtc->get_plat_tag() + " nodes=" + str(cp->nodes.size()) + " ops=" + str(cp->ops.size()) + " fwd_calls=" + str(fwd_calls.size()) +
             " conv_flops=" + str(cp->total_conv_flops());
}

// Runs of consecutive conv calls that are inner-product shaped at this batch (b200_compute_t::func_fc_chainable), each reading the previous
// one's output, become one "fc_chain" call (fcchain.cuh). The conv functions stay compiled: the chain uses their plans and filter packs, and
// every layer's fp32 node is still written, so other readers of those nodes are unaffected.
void b200_conv_fwd_t::fuse_fc_chains() {
  auto arg_var = [](fwd_call_t const &c, char const *an) -> string {
    auto i = c.rfc.arg_map.find(an);
    return (i != c.rfc.arg_map.end() && i->second.is_var()) ? i->second.get_var() : string();
  };
  auto plain = [&](fwd_call_t const &c) {  // a conv call with nothing but in / filts / biases / out and abs-max cells
    if (c.func_name.compare(0, 6, "conv__") != 0 || !rtc->func_fc_chainable(c.func_name)) { return false; }
    for (auto const &kv : c.rfc.arg_map) {
      string const &k = kv.first;
      if (k != "in" && k != "filts" && k != "biases" && k != "out" && k != "in_absmax_cells" && k != "in_absmax_ix" && k != "out_absmax_cells" && k != "out_absmax_ix") { return false; }
    }
    return true;
  };
  vector<fwd_call_t> calls;
  vector<double> flops;
  for (size_t i = 0; i < fwd_calls.size();) {
    size_t j = i;
    if (plain(fwd_calls[i])) {
      j = i + 1;
      while (j < fwd_calls.size() && j - i < 4 && plain(fwd_calls[j]) && arg_var(fwd_calls[j], "in") == arg_var(fwd_calls[j - 1], "out")) { ++j; }
    }
    if (j - i < 2) { calls.push_back(fwd_calls[i]); flops.push_back(call_flops[i]); ++i; continue; }
    fwd_call_t c;
    c.tag = fwd_calls[i].tag;
    c.func_name = "fc_chain__" + c.tag + "__" + str(calls.size());
    op_base_t cop;
    string layers;
    vector<string> tags;
    double fl = 0.0;
    for (size_t k = i; k < j; ++k) {
      fwd_call_t const &l = fwd_calls[k];
      string const si = str(k - i);
      layers += (k == i ? "" : ":") + l.func_name;
      tags.push_back(l.tag);
      fl += call_flops[k];
      for (auto const &kv : l.rfc.arg_map) {
        if (kv.first == "in" || kv.first == "in_absmax_cells" || kv.first == "in_absmax_ix") { if (k == i) { c.rfc.arg_map[kv.first] = kv.second; } continue; }
        string an = kv.first;  // filts -> filts<i>, out_absmax_ix -> out<i>_absmax_ix
        size_t const us = an.find('_');
        an = (us == string::npos) ? an + si : an.substr(0, us) + si + an.substr(us);
        c.rfc.arg_map[an] = kv.second;
      }
    }
    cop.str_vals["layers"] = layers;
    rtc_func_info_t fi;
    fi.func_name = c.func_name;
    fi.op = cop;
    fi.op.set_func_name("fc_chain");
    rtc->compile({fi}, rtc_compile_opts_t());
    c.rfc.rtc_func_name = c.func_name;
    calls.push_back(c);
    flops.push_back(fl);
    fc_chains.push_back(tags);
    i = j;
  }
  fwd_calls.swap(calls);
  call_flops.swap(flops);
}

string b200_conv_fwd_t::plan_text() const {
  string out;
  auto put_call = [&](char const *kind, fwd_call_t const &c) {
    out += string(kind) + " " + c.func_name;
    for (auto const &kv : c.rfc.arg_map) {
      out += " " + kv.first + "=";
      if (kv.second.is_var()) { out += kv.second.get_var(); } else { out += str((uint64_t)nda_scalar_as_double(*kv.second.get_nda())); }
    }
    string const lp = rtc->func_plan_text(c.func_name);  // convolutions: the launch plan, as "plan:<key>=<value>" items
    size_t b = 0;
    while (b < lp.size()) { size_t e = lp.find(' ', b); if (e == string::npos) { e = lp.size(); } out += " plan:" + lp.substr(b, e - b); b = e + 1; }
    out += "\n";
  };
  for (auto const &c : prep_calls) { put_call("prep", c); }
  for (auto const &c : fwd_calls) { put_call("call", c); }
  for (auto const &kv : concat_alias) { out += "alias " + kv.first + " " + kv.second.cat_node + " " + str(kv.second.ocix) + "\n"; }
  for (auto const &kv : res_fuse) { out += "join " + kv.first + " " + kv.second.out_node + " " + kv.second.res_node + "\n"; }
  for (auto const &ch : fc_chains) { out += "fcchain"; for (auto const &t : ch) { out += " " + t; } out += "\n"; }
  for (auto const &kv : lrn_fuse) { out += "lrnpool " + kv.first + " " + kv.second.lrn_tag + " " + kv.second.in_node + "\n"; }
  for (auto const &kv : absmax_ix) { out += "absmax " + kv.first + " " + str(kv.second) + "\n"; }
  return out;
}

void b200_conv_fwd_t::set_param(string const &node_name, float const *src, uint64_t n_elems) {
  p_conv_node_t n = cp->must_get_node(node_name);
  if (n->dims.dims_prod() != n_elems) { rt_err("set_param '" + node_name + "': got " + str(n_elems) + " elements, node holds " + str(n->dims.dims_prod())); }
  rtc->copy_raw_to_var(node_name, src, n_elems * 4);
  // the captured graph skips weight packing (packed once per weight version): new weights need a fresh warm-up + capture
  if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
  if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
  warmed = false;
}

// the same from a device buffer (a slice of the flat buffer the weights were broadcast in, b200_shard.cu): device-to-device on the back-end's stream
void b200_conv_fwd_t::set_param_device(string const &node_name, void const *dev_src, uint64_t n_elems) {
  p_conv_node_t n = cp->must_get_node(node_name);
  if (n->dims.dims_prod() != n_elems) { rt_err("set_param '" + node_name + "': got " + str(n_elems) + " elements, node holds " + str(n->dims.dims_prod())); }
  rtc->copy_device_to_var_async(node_name, dev_src, n_elems * 4);
  rtc->finish_and_sync();
  if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
  if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
  warmed = false;
}

bool b200_conv_fwd_t::attach_gather(string const &node, b200_gather_desc_t const *d) {
  bool ok = false;
  if (d && !fwd_calls.empty() && fwd_calls.back().func_name.compare(0, 10, "fc_chain__") == 0) {
    string last_out;
    for (auto const &kv : fwd_calls.back().rfc.arg_map) {  // out<i> with the largest i
      if (kv.first.compare(0, 3, "out") == 0 && kv.first.find('_') == string::npos && kv.second.is_var() && kv.first >= "out0") { last_out = kv.second.get_var(); }
    }
    if (last_out == node) { ok = rtc->set_tail_gather(node, d); }
  }
  if (!ok) { rtc->set_tail_gather(string(), nullptr); }
  if (ok || gather_attached) {  // the last call's launch changes: capture again
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
    if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
    warmed = false;
  }
  gather_attached = ok;
  return ok;
}

void b200_conv_fwd_t::run_calls() {
  rtc->set_var_to_zero(absmax_cells_vn);  // one memset re-arms every abs-max cell for this forward
  for (auto const &c : fwd_calls) { rtc->run(c.rfc); }
}

void b200_conv_fwd_t::ensure_graph() {
  // source nodes are written from outside the call list: mark them modified so that the warm-up passes and, above all, the captured
  // graph contain the pack of the first convolution's input (derived operands are cached on the source var's write generation)
  auto touch_sources = [&]() { for (auto const &dn : cp->data_node_names) { rtc->get_var_raw_native_pointer(dn); } };
  if (!warmed) {
    touch_sources();  // first pass eager: allocates packed-operand buffers, packs weights, sets kernel attributes
    uint64_t const l0 = rtc->launches();
    rtc->set_timing(false);
    for (auto const &c : prep_calls) { rtc->run(c.rfc); }  // parameter-only work (BatchNorm / Scale folding): once per weight version
    run_calls();
    rtc->finish_and_sync();
    run_calls();  // second pass = steady state (weights cached): count kernels per forward
    rtc->finish_and_sync();
    uint64_t const l1 = rtc->launches();
    run_calls();
    rtc->finish_and_sync();
    kernels_per_fwd = rtc->launches() - l1;
    (void)l0;
    warmed = true;
  }
  if (use_graph && !graph_exec) {
    rtc->set_timing(false);
    touch_sources();
    uint64_t const l0 = rtc->launches();
    CU_CHK(cudaStreamBeginCapture(rtc->stream(), cudaStreamCaptureModeThreadLocal));
    try { run_calls(); } catch (...) { cudaGraph_t g = nullptr; cudaStreamEndCapture(rtc->stream(), &g); if (g) { cudaGraphDestroy(g); } throw; }
    CU_CHK(cudaStreamEndCapture(rtc->stream(), &graph));
    CU_CHK(cudaGraphInstantiate(&graph_exec, graph, 0));
    kernels_per_fwd = rtc->launches() - l0;
  }
}

int b200_conv_fwd_t::submit(int n_set, char const *const *set_names, float const *const *set_bufs, uint64_t const *set_elems, int n_get,
                            char const *const *get_names, float *const *get_bufs, uint64_t const *get_elems) {
  for (int i = 0; i < n_set; ++i) {
    p_conv_node_t n = cp->must_get_node(set_names[i]);
    if (n->dims.dims_prod() != set_elems[i]) { rt_err(string("run_fwd: input '") + set_names[i] + "' has " + str(set_elems[i]) + " elements, node holds " + str(n->dims.dims_prod())); }
  }
  for (int i = 0; i < n_get; ++i) {
    p_conv_node_t n = cp->must_get_node(get_names[i]);
    if (n->dims.dims_prod() != get_elems[i]) { rt_err(string("run_fwd: output '") + get_names[i] + "' has " + str(get_elems[i]) + " elements, node holds " + str(n->dims.dims_prod())); }
  }
  ensure_graph();
  if (!copy_stream) {
    CU_CHK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (auto &sl : slots) { CU_CHK(cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming)); CU_CHK(cudaEventCreateWithFlags(&sl.freed, cudaEventDisableTiming)); }
    for (auto &e : ticket_ev) { CU_CHK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
  }
  int const ticket = (int)(n_submitted % kTickets);
  slot_t &sl = slots[n_submitted % kSlots];
  ++n_submitted;
  cudaStream_t const st = rtc->stream();
  // copy stream: wait until the slot's previous contents were consumed, then H2D into the staging slot
  if (sl.used) { CU_CHK(cudaStreamWaitEvent(copy_stream, sl.freed, 0)); }
  for (int i = 0; i < n_set; ++i) {
    void *&stg = sl.staging[set_names[i]];
    if (!stg) { CU_CHK(cudaMalloc(&stg, set_elems[i] * 4)); }
    CU_CHK(cudaMemcpyAsync(stg, set_bufs[i], set_elems[i] * 4, cudaMemcpyHostToDevice, copy_stream));
  }
  CU_CHK(cudaEventRecord(sl.h2d_done, copy_stream));
  // compute stream: staging -> node vars (device copy), forward, D2H of the requested nodes
  CU_CHK(cudaStreamWaitEvent(st, sl.h2d_done, 0));
  for (int i = 0; i < n_set; ++i) { rtc->copy_device_to_var_async(set_names[i], sl.staging[set_names[i]], set_elems[i] * 4); }
  CU_CHK(cudaEventRecord(sl.freed, st));
  sl.used = true;
  if (use_graph) { CU_CHK(cudaGraphLaunch(graph_exec, st)); graph_launches += kernels_per_fwd; }
  else { rtc->set_timing(false); run_calls(); }
  for (int i = 0; i < n_get; ++i) {
    materialise_aliased(get_names[i]);
    rtc->copy_var_to_raw_async(get_bufs[i], get_names[i], get_elems[i] * 4);
  }
  CU_CHK(cudaEventRecord(ticket_ev[ticket], st));
  return ticket;
}

void b200_conv_fwd_t::materialise_aliased(string const &node) {
  auto el = elided_calls.find(node);
  if (el != elided_calls.end()) {  // bypassed by a residual join in its convolution's epilogue: compute it now (inputs are still in their vars)
    rtc->set_timing(false);
    rtc->run(el->second);
    return;
  }
  auto al = concat_alias.find(node);
  if (al == concat_alias.end()) { return; }
  rtc_func_call_t rfc;
  rfc.rtc_func_name = al->second.extract_func;
  rfc.arg_map = map_str_rtc_arg_t{{"in", node}, {"out", al->second.cat_node}};
  rtc->set_timing(false);
  rtc->run(rfc);
}

void b200_conv_fwd_t::wait(int ticket) {
  if (ticket < 0 || ticket >= kTickets || !ticket_ev[ticket]) { rt_err("wait: invalid ticket"); }
  CU_CHK(cudaEventSynchronize(ticket_ev[ticket]));
}

void b200_conv_fwd_t::run_fwd_raw(int n_set, char const *const *set_names, float const *const *set_bufs, uint64_t const *set_elems, int n_get,
                                  char const *const *get_names, float *const *get_bufs, uint64_t const *get_elems) {
  if (enable_prof) { rtc->profile_start(); }  // src/rtc_fwd.cc:537
  wait(submit(n_set, set_names, set_bufs, set_elems, n_get, get_names, get_bufs, get_elems));
  if (enable_prof) { rtc->profile_stop(); }   // src/rtc_fwd.cc:559
}

void b200_conv_fwd_t::run_fwd(vect_string const &to_set_vns, p_map_str_p_nda_float_t const &fwd, vect_string const &to_get_vns) {
  vector<char const *> sn, gn;
  vector<float const *> sb;
  vector<float *> gb;
  vector<uint64_t> se, ge;
  for (auto const &vn : to_set_vns) {
    auto i = fwd->find(vn);
    if (i == fwd->end()) { rt_err("run_fwd: input '" + vn + "' not in fwd map"); }
    if (!(i->second->dims == cp->must_get_node(vn)->dims)) { rt_err("run_fwd: dims mismatch for input '" + vn + "'"); }
    sn.push_back(vn.c_str()); sb.push_back(static_cast<float *>(i->second->rp_elems())); se.push_back(i->second->elems_sz());
  }
  for (auto const &vn : to_get_vns) {  // created if absent; overwritten if dims match (src/rtc_compute.cc:92-97)
    dims_t const &d = cp->must_get_node(vn)->dims;
    auto i = fwd->find(vn);
    if (i == fwd->end()) { (*fwd)[vn] = std::make_shared<nda_float_t>(d); i = fwd->find(vn); }
    else if (!(i->second->dims == d)) { rt_err("run_fwd: dims mismatch for output '" + vn + "'"); }
    gn.push_back(vn.c_str()); gb.push_back(static_cast<float *>(i->second->rp_elems())); ge.push_back(d.dims_prod());
  }
  run_fwd_raw((int)sn.size(), sn.data(), sb.data(), se.data(), (int)gn.size(), gn.data(), gb.data(), ge.data());
}

float b200_conv_fwd_t::run_device_only(int iters) {
  ensure_graph();
  cudaEvent_t b, e;
  CU_CHK(cudaEventCreate(&b));
  CU_CHK(cudaEventCreate(&e));
  rtc->set_timing(false);
  CU_CHK(cudaEventRecord(b, rtc->stream()));
  for (int i = 0; i < iters; ++i) {
    if (use_graph) { CU_CHK(cudaGraphLaunch(graph_exec, rtc->stream())); graph_launches += kernels_per_fwd; }
    else { run_calls(); }
  }
  CU_CHK(cudaEventRecord(e, rtc->stream()));
  CU_CHK(cudaEventSynchronize(e));
  float ms = 0;
  CU_CHK(cudaEventElapsedTime(&ms, b, e));
  cudaEventDestroy(b);
  cudaEventDestroy(e);
  return ms / std::max(iters, 1);
}

vector<b200_conv_fwd_t::prof_row_t> b200_conv_fwd_t::profile(int iters) {
  ensure_graph();
  vector<prof_row_t> res;
  for (size_t i = 0; i < fwd_calls.size(); ++i) { res.push_back(prof_row_t{fwd_calls[i].func_name, 0.0f, 0.0f, i < call_flops.size() ? call_flops[i] : 0.0}); }
  rtc->set_timing(true);
  for (int it = 0; it < iters; ++it) {
    rtc->release_per_call_id_data();
    vector<uint32_t> ids;
    for (auto const &dn : cp->data_node_names) { rtc->get_var_raw_native_pointer(dn); }  // inputs count as freshly written: include their pack
    rtc->set_var_to_zero(absmax_cells_vn);
    for (auto const &c : fwd_calls) { ids.push_back(rtc->run(c.rfc)); }
    rtc->finish_and_sync();
    for (size_t i = 0; i < ids.size(); ++i) { res[i].call_ms += rtc->get_dur(ids[i], ids[i]) / iters; res[i].kernel_ms += rtc->get_kernel_dur(ids[i]) / iters; }
  }
  rtc->release_per_call_id_data();
  rtc->set_timing(false);
  return res;
}

void b200_conv_fwd_t::flush_l2(uint64_t bytes) {
  if (!bytes) { return; }
  if (bytes > flush_bytes) {
    if (flush_buf) { CU_CHK(cudaStreamSynchronize(rtc->stream())); cudaFree(flush_buf); flush_buf = nullptr; }
    CU_CHK(cudaMalloc(&flush_buf, bytes));
    flush_bytes = bytes;
  }
  CU_CHK(cudaMemsetAsync(flush_buf, (int)(flush_count++ & 0xff), bytes, rtc->stream()));
}

void b200_conv_fwd_t::enqueue_fwd() {
  ensure_graph();
  rtc->set_timing(false);
  if (use_graph) { CU_CHK(cudaGraphLaunch(graph_exec, rtc->stream())); graph_launches += kernels_per_fwd; }
  else { run_calls(); }
}

vector<float> b200_conv_fwd_t::run_timed(int iters, uint64_t l2_flush_bytes) {
  ensure_graph();
  rtc->set_timing(false);
  vector<cudaEvent_t> evb(iters), eve(iters);
  for (int i = 0; i < iters; ++i) { CU_CHK(cudaEventCreate(&evb[i])); CU_CHK(cudaEventCreate(&eve[i])); }
  for (int i = 0; i < iters; ++i) {
    flush_l2(l2_flush_bytes);
    CU_CHK(cudaEventRecord(evb[i], rtc->stream()));
    enqueue_fwd();
    CU_CHK(cudaEventRecord(eve[i], rtc->stream()));
  }
  rtc->finish_and_sync();
  vector<float> ms(iters, 0.0f);
  for (int i = 0; i < iters; ++i) { CU_CHK(cudaEventElapsedTime(&ms[i], evb[i], eve[i])); cudaEventDestroy(evb[i]); cudaEventDestroy(eve[i]); }
  return ms;
}

}  // namespace boda
