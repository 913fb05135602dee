// b200_conv_fwd.h -- `mode=b200`: whole-net forward over be=b200, mirroring has_conv_fwd_t (src/has_conv_fwd.H:16-25)
// as implemented by conv_pipe_fwd_t (src/rtc_fwd.cc:43-577), plus the slice of the conv_pipe graph IR it needs
// (conv_op_t / conv_node_t / conv_pipe_t, src/conv_util.H:75-243; dims rules src/conv_util.cc:167-226, :405-529).
#pragma once
#include "b200_compute.h"

namespace boda {

struct conv_op_t : public op_base_t {  // src/conv_util.H:112-141
  string tag;
  vect_string bots, tops;
  bool in_place = false;  // single bot == single top (ReLU / Dropout written onto their input node)
  bool fused = false;     // folded into the producer (conv+ReLU, src/rtc_fwd.cc:486-494)
  bool is(string const &t) const { return has_type() && get_type() == t; }
};
typedef shared_ptr<conv_op_t> p_conv_op_t;

struct conv_node_t {  // src/conv_util.H:152-170
  string name;
  dims_t dims;
  vect_string top_for, bot_for;
  vector<p_conv_op_t> in_place_ops;
  bool is_param = false;  // filts / biases (conv_pipe_t::op_params)
};
typedef shared_ptr<conv_node_t> p_conv_node_t;

struct conv_pipe_t {  // src/conv_util.H:172-243
  map<string, p_conv_node_t> nodes;
  vector<p_conv_op_t> ops;  // topological (file) order
  vect_string data_node_names, param_names;
  p_conv_node_t must_get_node(string const &n) const { auto i = nodes.find(n); if (i == nodes.end()) { rt_err("pipe: no node named '" + n + "'"); } return i->second; }
  p_conv_node_t get_or_make_node(string const &n);
  void add_op_from_lexp(lexp_t const &l);
  void calc_dims();  // src/conv_util.cc:405-529
  uint64_t total_conv_flops() const;
};
typedef shared_ptr<conv_pipe_t> p_conv_pipe_t;
p_conv_pipe_t make_conv_pipe_from_text(string const &pipe_text);

// The unique Convolution signatures of a pipe -- op parameters plus the dims of in / filts / biases / out, without tags -- one canonical op
// line per signature, sorted: what the reference's write_op_sigs option collects (src/rtc_fwd.cc:246-264) and its per-op flows read back
// (test/conv-ops-*.txt). Host-only.
string conv_pipe_op_sigs_text(conv_pipe_t const &cp);

struct fwd_call_t { string func_name, tag; rtc_func_call_t rfc; };

struct b200_conv_fwd_t {
  string mode = "b200";
  // options (conv_pipe_fwd_t fields, src/rtc_fwd.cc:48-66, reduced to what applies)
  uint32_t use_graph = 1;      // replay the forward calls as one CUDA graph (launch-bound at B200 speeds)
  uint32_t enable_prof = 0;
  uint32_t pack_by_producers = 1; // bf16 storage mode: convolutions also write the NHWC bf16 plane their consumers read (no activation pack kernel)
  uint32_t fuse_lrn_pool = 1;     // an LRN (local_size 5) read only by a 3x3 / 2 max pool runs inside that pool's kernel: no round trip of the normalised map, one
                                  // launch fewer; the LRN node is computed on demand (plain lrn call) when run_fwd is asked for it
  uint32_t fuse_fc_chain = 1;     // consecutive inner-product-shaped convolutions at batch <= 32 (AlexNet fc6 -> fc7 -> fc8), each reading the previous one's
                                  // output, run as ONE persistent kernel (fcchain.cuh): the weight stream does not stop between the layers
  uint32_t fuse_eltwise = 1;      // residual joins: a two-input Eltwise SUM (+ReLU) whose one input is written by a Convolution and read by nothing else is
                                  // folded into that convolution's epilogue; the bypassed node is recomputed on demand when read
  uint32_t concat_by_offset = 1;  // Convolutions that only feed a Concat write straight into its output at their channel offset (no copy kernel);
                                  // their own node is materialised from that slice only when run_fwd is asked for it
  p_b200_compute_t rtc;
  p_conv_pipe_t cp;
  vector<fwd_call_t> fwd_calls;
  vector<fwd_call_t> prep_calls;  // run once per weight version, outside the forward (BatchNorm / Scale folding into the conv parameters)
  string info_log;

  b200_conv_fwd_t();
  ~b200_conv_fwd_t();
  void init(p_conv_pipe_t const &cp_, string const &opts);  // has_conv_fwd_t::init
  void set_det_drop_seed(uint32_t const &) {}              // dropout is stripped from fwd graphs (src/caffepb.cc:235-238)
  void run_fwd(vect_string const &to_set_vns, p_map_str_p_nda_float_t const &fwd, vect_string const &to_get_vns);
  string get_info_log() { return info_log; }

  // raw-buffer variants used by the C ABI
  void set_param(string const &node_name, float const *src, uint64_t n_elems);
  void set_param_device(string const &node_name, void const *dev_src, uint64_t n_elems);
  void run_fwd_raw(int n_set, char const *const *set_names, float const *const *set_bufs, uint64_t const *set_elems, int n_get,
                   char const *const *get_names, float *const *get_bufs, uint64_t const *get_elems);
  // Pipelined form of run_fwd for serving loops: submit() enqueues H2D of this batch's inputs (on a copy stream, into one of two staging
  // slots, so it overlaps the previous batch's forward), the forward, and D2H of the requested nodes; wait() blocks until that batch's
  // outputs are in the caller's host buffers. Host buffers must stay valid until wait(). run_fwd_raw == submit + wait.
  int submit(int n_set, char const *const *set_names, float const *const *set_bufs, uint64_t const *set_elems, int n_get,
             char const *const *get_names, float *const *get_bufs, uint64_t const *get_elems);
  void wait(int ticket);
  float run_device_only(int iters);  // ms per forward, CUDA-event timed on the back-end's stream
  struct prof_row_t { string func_name; float call_ms, kernel_ms; double flops; };
  vector<prof_row_t> profile(int iters);
  // `iters` forwards, each bracketed by its own CUDA events on the back-end's stream; when l2_flush_bytes > 0 a scratch buffer of that size is
  // overwritten before every forward (outside the events) so no iteration starts with a warm L2. Returns ms per forward.
  vector<float> run_timed(int iters, uint64_t l2_flush_bytes);
  // Asynchronous pieces for a caller that orders its own device work (e.g. an NCCL gather of the logits on another stream) against the
  // forward: stream() is the back-end's cudaStream_t; enqueue_fwd() queues one forward on device-resident inputs and returns without a
  // sync; flush_l2() queues an overwrite of a scratch buffer of `bytes`.
  void *stream() const { return (void *)rtc->stream(); }
  void enqueue_fwd();
  // multi-GPU: let the forward's last call (an fc_chain ending in `node`) do the logits gather inside its kernel; false = not that shape (the caller
  // keeps launching b200_shard_gather_push_wait). nullptr detaches. Either way the captured graph is dropped.
  bool attach_gather(string const &node, b200_gather_desc_t const *d);
  bool gather_attached = false;
  void flush_l2(uint64_t bytes);
  uint64_t launches() const { return rtc->launches() + graph_launches; }
  // The forward as planned at init, one line per item, for inspection and host-only tests (also on a plan_only=1 instance, which needs no
  // device):  "call <func_name> <arg>=<var or by-value scalar> ... [plan:<key>=<value> ...]" in launch order (convolutions carry their launch plan,
  // b200_compute_t::func_plan_text), "prep <func_name>" (parameter-only calls),
  // "alias <node> <concat node> <channel offset>", "join <conv tag> <join node> <residual node>", "absmax <node> <cell>".
  string plan_text() const;

 private:
  void gen_op(p_conv_op_t const &op);
  void add_call(string const &fn_base, conv_op_t const &op, op_base_t const &fop, map_str_rtc_arg_t const &args);
  void run_calls();
  void ensure_graph();
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_launches = 0, kernels_per_fwd = 0;
  bool warmed = false;
  static constexpr int kSlots = 2, kTickets = 8;
  cudaStream_t copy_stream = nullptr;
  struct slot_t { map<string, void *> staging; cudaEvent_t h2d_done = nullptr, freed = nullptr; bool used = false; } slots[kSlots];
  cudaEvent_t ticket_ev[kTickets] = {};
  uint64_t n_submitted = 0;
  void *flush_buf = nullptr;
  uint64_t flush_bytes = 0, flush_count = 0;
  vector<double> call_flops;
  map<string, uint32_t> absmax_ix;  // nodes whose producer publishes max|x| for the consuming convolution's operand scaling
  struct concat_alias_t { string cat_node; uint32_t ocix; string extract_func; };
  struct res_fuse_t { string out_node, res_node; };
  map<string, res_fuse_t> res_fuse;          // Convolution tag -> (Eltwise output it writes, the join's other input)
  struct lrn_fuse_t { string lrn_tag, in_node, lrn_node; };
  map<string, lrn_fuse_t> lrn_fuse;          // Pooling tag -> the LRN in front of it that runs inside the pool kernel (lrn_maxpool_kernel)
  vector<vector<string>> fc_chains;          // per fc_chain call: the tags of the convolutions it runs
  void fuse_fc_chains();
  map<string, rtc_func_call_t> elided_calls; // bypassed node -> the plain convolution call that materialises it when somebody reads it
  map<string, concat_alias_t> concat_alias;  // node -> (Concat output that holds its channels, channel offset, name of the read-back function)
  void materialise_aliased(string const &node);
  op_base_t conv_fop(conv_op_t const &op) const;  // function signature of a Convolution op: its params + the dims of in / filts / biases / out
  bool dst_plane_by_producers(string const &dst);
  bool dst_plane_pad(string const &dst, int &py, int &px);
  void add_out_pack_args(map_str_rtc_arg_t &args, string const &dst);  // bf16 mode: every producer of node `dst` can write its NHWC plane and some conv reads it  // enqueue the slice copy Concat output -> node var (before reading an aliased node)
  void add_absmax_args(map_str_rtc_arg_t &args, string const &which, string const &node);
};

}  // namespace boda
