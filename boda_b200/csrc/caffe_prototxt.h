// caffe_prototxt.h -- Caffe net description (protobuf TEXT format) -> conv_pipe text, without protobuf.
//
// The reference builds its conv_pipe_t from a prototxt through libprotobuf + Caffe's upgrade code (create_pipe_from_param,
// src/caffepb.cc:166-326; fill_in_conv_op_from_param :79-141; phase rules via layer_included_for_state). This is a from-scratch reader for
// the subset those nets use: a generic text-format parser (fields, nested messages, strings, numbers, enums, '#' comments) and the same
// layer-by-layer translation for the TEST-phase forward graph: top-level `input` / `input_dim` / `input_shape` blobs and Data layers become
// source nodes ({batch, 3, crop, crop}, overridable through in_dims), Convolution / Pooling / LRN / ReLU / Dropout / Concat / InnerProduct
// map one to one, Accuracy / SoftmaxWithLoss are dropped, Softmax is dropped unless keep_softmax (src/caffepb.cc:250-261), and -- beyond
// the reference, whose readers for them are stubs (:231-232, :308) -- BatchNorm (eps), Scale, Eltwise and convolution_param.bias_term are
// read. Both `layer { type: "Convolution" }` and the V1 `layers { type: CONVOLUTION }` spelling are accepted.
// The output is the pipe text b200_fwd_create / make_conv_pipe_from_text take, so everything downstream is unchanged.
#pragma once
#include "boda_base.h"

namespace boda {

struct prototxt_opts_t {
  map<string, uint32_t> in_dims;  // overrides for the source nodes' dims by name, e.g. {"img": 32} (maybe_override_dims_and_calc_strides)
  string out_node_name;           // stop after the layer that produces this node ("" = read all layers)
  bool keep_softmax = false;      // the reference drops Softmax layers from forward graphs
};

// throws rt_exception on syntax errors and on layer kinds / parameters outside the supported subset
string conv_pipe_text_from_prototxt(string const &prototxt, prototxt_opts_t const &opts);

}  // namespace boda
