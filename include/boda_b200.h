/* boda_b200.h -- C ABI of libboda_b200.so: the drop-in boundary of the B200-native rtc_fwd back-end.
 *
 * Plain pointers, sizes and NUL-terminated strings only; no C++/torch types; no exceptions cross the boundary:
 * every function returns 0 on success (or a non-negative id / handle where stated) and a negative code on
 * failure, with the message available from b200_last_error(). Codes: -1 = rt_exception (reference: rt_err,
 * src/boda_base.H:98-105), -2 = unsup_exception (unsup_err: shape/feature this back-end does not handle).
 *
 * Each entry point names the reference interface it replaces (paths relative to moskewcz/boda):
 *   tier A = rtc_compute_t      src/rtc_compute.H:35-97   (implemented for `be=nvrtc` by src/nvrtc_util.cc:174-395)
 *   tier B = has_conv_fwd_t     src/has_conv_fwd.H:16-25  (implemented for `mode=rtc` by src/rtc_fwd.cc:43-577)
 * INTEGRATION.md shows the C++ adaptor classes a Boda maintainer adds to bind these (NESI type_id "b200").
 */
#ifndef BODA_B200_H_
#define BODA_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_rtc b200_rtc;  /* one back-end instance == one rtc_compute_t object (one device, one stream) */
typedef struct b200_fwd b200_fwd;  /* one has_conv_fwd_t object */

/* ---- library ---- */
const char *b200_last_error(void);                 /* message of the last failure on this thread */
const char *b200_version(void);
int b200_device_count(void);                       /* number of CUDA devices visible (0 on a CPU-only box) */

/* ---- tier A: rtc_compute_t ---- */
b200_rtc *b200_rtc_create(void);                   /* NESI construction of `(be=b200)`; never touches the GPU */
void b200_rtc_destroy(b200_rtc *r);
int b200_rtc_set_option(b200_rtc *r, const char *key, const char *val); /* "prec"=fp32|fp16|bf16, "acc_chunk_kblks"=N, "device"=N */
int b200_rtc_init(b200_rtc *r);                    /* rtc_compute_t::init()            src/rtc_compute.H:45 */
const char *b200_rtc_get_plat_tag(b200_rtc *r);    /* ::get_plat_tag()                 :46 */
/* ::create_var_with_dims(vn, dims)  :48  -- dims as element type name + named sizes; new vars are zero-filled */
int b200_rtc_create_var(b200_rtc *r, const char *vn, const char *tn, int ndims, const char *const *dim_names, const uint32_t *dim_sizes);
/* ::create_var_with_dims_as_reshaped_view_of_var  :49 */
int b200_rtc_create_view(b200_rtc *r, const char *vn, const char *tn, int ndims, const char *const *dim_names, const uint32_t *dim_sizes, const char *src_vn);
int b200_rtc_release_var(b200_rtc *r, const char *vn);           /* ::release_var       :50 */
/* ::get_var_dims  :51 -- returns ndims (<= max_dims filled), dim names are written as a ':'-joined string */
int b200_rtc_get_var_dims(b200_rtc *r, const char *vn, int max_dims, uint32_t *dim_sizes, char *names_buf, int names_buf_len);
int b200_rtc_set_var_to_zero(b200_rtc *r, const char *vn);       /* ::set_var_to_zero   :52 */
/* ::compile(func_infos, opts)  :55 -- binds `func_name` to a precompiled sm_100a kernel chosen from the op text
 * (current or stale op-line syntax). No source is compiled at run time. */
int b200_rtc_compile(b200_rtc *r, const char *func_name, const char *op_text);
int b200_rtc_release_func(b200_rtc *r, const char *func_name);   /* ::release_func      :56 */
/* ::run(rtc_func_call_t)  :59 -- arg_vals[i] is a var name, or an nda literal "(tn=uint32_t,v=5)" passed by value.
 * Returns the call id (>= 0). tpb/blks of the reference call struct are not needed. */
int b200_rtc_run(b200_rtc *r, const char *func_name, int nargs, const char *const *arg_names, const char *const *arg_vals);
int b200_rtc_finish_and_sync(b200_rtc *r);                       /* ::finish_and_sync   :60 */
int b200_rtc_release_per_call_id_data(b200_rtc *r);              /* ::release_per_call_id_data :61 */
int b200_rtc_release_all_funcs(b200_rtc *r);                     /* ::release_all_funcs :62 */
int b200_rtc_get_dur(b200_rtc *r, uint32_t b, uint32_t e, float *ms_out); /* ::get_dur(b,e) in ms  :70 */
int b200_rtc_get_kernel_dur(b200_rtc *r, uint32_t call_id, float *ms_out); /* the call's contraction kernel alone (no operand packing) */
/* ::copy_nda_to_var / ::copy_var_to_nda  :79-81 -- host buffers; `bytes` must equal the var's size */
int b200_rtc_copy_to_var(b200_rtc *r, const char *vn, const void *host_src, uint64_t bytes);
int b200_rtc_copy_from_var(b200_rtc *r, void *host_dst, const char *vn, uint64_t bytes);
/* ::get_var_raw_native_pointer  :80 -- device pointer (for external low-level libs) */
int b200_rtc_get_var_raw_native_pointer(b200_rtc *r, const char *vn, void **dev_ptr_out);
uint64_t b200_rtc_launches(b200_rtc *r);           /* kernels launched so far by this instance */

/* ---- host-only: the conv_pipe graph IR (needs no GPU) ---- */
/* Parse `pipe_text` (same format as b200_fwd_create) and run the dims inference of conv_pipe_t::calc_dims (src/conv_util.cc:405-529):
 * writes one line per node, "<name> <dim>=<sz>:... [param]" in name order, then "ops <n> conv_flops <f>", into buf (NUL-terminated,
 * truncated to buf_len). Returns the full length needed (excluding the NUL), or <0 on a malformed pipe (see b200_last_error). */
int64_t b200_pipe_describe(const char *pipe_text, char *buf, uint64_t buf_len);

/* Host-only: an op line in either syntax of the reference's op-list files (current "(str_vals=(type=..),nda_vals=(..))" or the stale
 * "(type=..,dims_vals=(..),str_vals=(out_chans=N))") -> its canonical current-syntax line (the reference's printer:
 * src/boda_base.cc:392-420 inside NESI's struct dump). Returns the text length or <0. */
int64_t b200_op_canonical(const char *op_text, char *buf, uint64_t buf_len);
/* Host-only: the unique Convolution signatures of a pipe (op parameters + dims of in / filts / biases / out, no tags), one canonical op line
 * each, sorted -- what the reference's write_op_sigs option collects (src/rtc_fwd.cc:246-264) and what its per-op test files hold
 * (test/conv-ops-*.txt); feed them to b200_rtc_compile / tools/ops_prof.py. Returns the text length or <0. */
int64_t b200_pipe_op_sigs(const char *pipe_text, char *buf, uint64_t buf_len);
/* Host-only: the forward as b200_fwd_create would plan it for this pipe and these options -- graph passes (concat by offset, residual joins
 * in the convolution, producer-written planes, abs-max cells) and the launch plan of every convolution for a 148-SM device -- without touching
 * a device ("plan_only=1" instance: it cannot compute). Text, one item per line: "call <func> <arg>=<var|scalar> ... plan:<key>=<value> ..."
 * (convolutions carry their launch plan: kernel=pair|single, bn, kblks, splits, swapped, rowmerge, im2col, grid), "prep <func> ...",
 * "alias <node> <concat node> <chan offset>", "join <conv tag> <join node> <residual node>", "absmax <node> <cell>".
 * Returns the text length (the buffer gets at most buf_len-1 chars) or <0. */
int64_t b200_fwd_plan(const char *pipe_text, const char *opts, char *buf, uint64_t buf_len);
/* Caffe prototxt (protobuf text format) -> pipe text, host-only, no protobuf: the translation of create_pipe_from_param
 * (src/caffepb.cc:166-326) for TEST-phase forward graphs. `opts` is a lexp list: "(img=32)" style dims overrides for the source nodes
 * (the reference's in_dims), "out_node_name=<node>" to stop after the layer producing it, "keep_softmax=1" to keep Softmax layers (the
 * reference drops them). Same buffer convention as b200_pipe_describe: returns the length needed, or <0 on error. */
int64_t b200_pipe_from_prototxt(const char *prototxt_text, const char *opts, char *buf, uint64_t buf_len);

/* Boda wisdom-file records (host-only; src/op-tuner.cc:98-130, src/boda_base.cc:210-383): hex(bwrite(nda_digest_t)) of a host float tensor
 * with named dims, seeded by std::hash<string>(var_name) as ops-prof does (src/rtc_prof.cc:297-310); and one op_wisdom_t text record with
 * `n_kgs` (name, digest hex) pairs and one op_tune_wisdom_t/op_run_t block. Same buffer convention as b200_pipe_describe. */
int64_t b200_nda_digest_hex(const char *var_name, int ndims, const char *const *dim_names, const uint32_t *dim_sizes, const float *host_data,
                            char *buf, uint64_t buf_len);
int64_t b200_wisdom_record(const char *op_text, int n_kgs, const char *const *kg_names, const char *const *kg_hex, const char *op_tune_text,
                           const char *be_plat_tag, double rt_secs, const char *err, const char *run_op_text, char *buf, uint64_t buf_len);
/* The reference's wis-ana analysis (src/op-tuner.cc:204-396) over wisdom text in its own format (test/wisdom-merged.wis, or records written
 * by b200_wisdom_record): runs with errors or of platforms not matching the regex s_plat are dropped, ops are filtered by batch size s_img
 * (0 = all) and min_flops; per op it reports FLOPs, AOM (time of the single best overall tune), POM (per-op minimum over the non-reference
 * tunes) and REF (time of the tune equal to ref_tune, "" = none). detail == 0: the csv wis-plot.py reads ("OP FLOPS boda-manual-tune
 * boda-autotuned REF"); detail == 1: "#aom_tune\t<tune>\ttot_runs\t<n>" then tab-separated flops, aom, pom, ref, pom tune, op text per op.
 * Host-only. Returns the text length or <0. */
int64_t b200_wis_ana(const char *wisdom_text, uint32_t s_img, const char *s_plat, const char *ref_tune, double min_flops, int detail, char *buf,
                     uint64_t buf_len);

/* ---- tier B: has_conv_fwd_t ---- */
/* has_conv_fwd_t::init(conv_pipe, nia)  src/has_conv_fwd.H:21 (conv_pipe_fwd_t::init, src/rtc_fwd.cc:469-527).
 * `pipe_text`: one conv_op_t per line in NESI text form,
 *   (tag=conv1,str_vals=(type=Convolution),nda_vals=(kern_sz=..,stride=..,in_pad=..,out_chans=..),bots=(data,conv1_filts,conv1_biases),tops=(conv1))
 * preceded by one line per source node: (node=data,dims=(img=32,chan=3,y=227,x=227)).
 * `opts`: lexp list of conv_pipe_fwd_t-style options, e.g. "(prec=fp32,use_graph=1,acc_chunk_kblks=4,device=0)". */
b200_fwd *b200_fwd_create(const char *pipe_text, const char *opts);
void b200_fwd_destroy(b200_fwd *f);
/* parameter (filts/biases) upload: conv_pipe_t::op_params -> copy_ndas_to_vars (src/rtc_fwd.cc:524) */
int b200_fwd_set_param(b200_fwd *f, const char *node_name, const float *host_src, uint64_t n_elems);
/* has_conv_fwd_t::run_fwd(to_set_vns, fwd, to_get_vns)  src/has_conv_fwd.H:23 (src/rtc_fwd.cc:529-577): host fp32 NCHW
 * buffers in, host buffers out; synchronous. n_elems arrays are checked against the node dims. */
int b200_fwd_run(b200_fwd *f, int n_set, const char *const *set_names, const float *const *set_bufs, const uint64_t *set_elems,
                 int n_get, const char *const *get_names, float *const *get_bufs, const uint64_t *get_elems);
/* Pipelined form of run_fwd for serving loops (run_fwd == submit + wait): submit() enqueues H2D of this batch's inputs on a copy
 * stream (two staging slots, so the copy overlaps the previous batch's forward), the forward and the D2H of the requested nodes, and
 * returns a ticket (>= 0); wait() blocks until that batch's outputs are in the caller's host buffers. At most 2 batches may be
 * in flight; host buffers (pinned for overlap) must stay valid until the matching wait(). */
int b200_fwd_submit(b200_fwd *f, int n_set, const char *const *set_names, const float *const *set_bufs, const uint64_t *set_elems,
                    int n_get, const char *const *get_names, float *const *get_bufs, const uint64_t *get_elems);
int b200_fwd_wait(b200_fwd *f, int ticket);
/* device-resident variant for benchmarking the kernels alone: inputs must have been set by a previous b200_fwd_run /
 * b200_fwd_set_param; runs the forward calls only and returns after a stream sync. */
int b200_fwd_run_device_only(b200_fwd *f, int iters, float *ms_per_iter_out);
int b200_fwd_set_det_drop_seed(b200_fwd *f, uint32_t seed);      /* ::set_det_drop_seed  :22 (no-op: dropout is stripped from fwd graphs) */
const char *b200_fwd_get_info_log(b200_fwd *f);                  /* ::get_info_log       :24 */
/* node dims: returns ndims (img,chan,y,x order) or <0 */
int b200_fwd_get_node_dims(b200_fwd *f, const char *node_name, uint32_t *dims4);
int b200_fwd_num_calls(b200_fwd *f);                             /* forward calls per run_fwd (size of fwd_calls) */
uint64_t b200_fwd_launches(b200_fwd *f);                         /* kernels launched so far (graph replays counted per node) */
/* per-call profile of `iters` eager runs (the per_call_fn dump of src/rtc_fwd.cc:560-572): function names ('\n'-joined), ms per
 * call, ms of the call's contraction kernel alone (conv calls also pack their operands), algorithmic FLOPs. Returns #calls. */
int b200_fwd_profile(b200_fwd *f, int iters, char *tags_buf, int tags_buf_len, float *call_ms_out, float *kernel_ms_out, double *flops_out, int max_calls);
/* `iters` forwards on device-resident inputs, each bracketed by its own CUDA events on the back-end's stream; a scratch
 * buffer of l2_flush_bytes (0 = off) is overwritten before each one, outside the events. ms_each_out[iters]. */
int b200_fwd_run_timed(b200_fwd *f, int iters, uint64_t l2_flush_bytes, float *ms_each_out);
/* Asynchronous pieces for a caller that orders its own device work against the forward (bench.py overlaps the NCCL gather of batch
 * i's logits with batch i+1's forward): the back-end's cudaStream_t; one forward on device-resident inputs queued without a sync; an
 * overwrite of a scratch buffer of `bytes` queued on the same stream (L2 flush between timed iterations). */
int b200_fwd_get_stream(b200_fwd *f, void **stream_out);
int b200_fwd_enqueue(b200_fwd *f);
int b200_fwd_flush_l2(b200_fwd *f, uint64_t bytes);
/* device pointer of a node's fp32 NCHW var (e.g. to hand the logits to NCCL) */
int b200_fwd_get_node_raw_native_pointer(b200_fwd *f, const char *node_name, void **dev_ptr_out);

/* ---- batch sharding of run_fwd over the GPUs of one box (the reference has no multi-GPU path, north_star asks for
 * "a single NCCL broadcast of weights and gather of logits over NVLink" with a C++ host side). One process per GPU, one b200_shard per
 * process. Rendezvous bytes (the 128-byte ncclUniqueId, the 64-byte CUDA IPC handles) are plain memory: the caller moves them between the
 * processes with whatever it has (bench.py: torch.distributed object collectives; a Boda driver: files or MPI). ---- */
typedef struct b200_shard b200_shard;
b200_shard *b200_shard_create(int device, int rank, int world);       /* needs a CUDA device (no CPU fallback); NULL + b200_last_error() on failure */
void b200_shard_destroy(b200_shard *s);
/* weights: rank 0 makes the id, every rank joins (ncclCommInitRank; the NCCL C API is resolved from libnccl.so.2 at run time), then ONE
 * in-place ncclBroadcast of a flat device buffer; b200_fwd_set_param_device slices it into the parameter vars device-to-device */
int b200_shard_nccl_unique_id(b200_shard *s, void *id_out_128);
int b200_shard_nccl_init(b200_shard *s, const void *id_128);
int b200_shard_broadcast(b200_shard *s, void *dev_buf, uint64_t bytes, int root, void *stream);
int b200_fwd_set_param_device(b200_fwd *f, const char *node_name, const void *dev_src, uint64_t n_elems);
/* logits: every rank owns a gather buffer [2][world][bytes_per_rank] (two halves by step parity: the result of step i stays readable while step
 * i+1 arrives) that its peers map through CUDA IPC. b200_shard_gather_push queues ONE small
 * kernel on `stream` (normally the forward's own stream, b200_fwd_get_stream) that writes this rank's output straight into slot [rank] of every
 * peer's buffer over NVLink and then publishes its step number there; b200_shard_gather_wait queues a one-CTA kernel that returns once every
 * rank's step number in the LOCAL buffer has reached `step`. No collective kernel occupies SMs beside the forward. push returns the step (>= 1)
 * or a negative code. b200_shard_all_gather_nccl is the ncclAllGather of the same data, kept for A/B runs. */
int b200_shard_gather_export(b200_shard *s, uint64_t bytes_per_rank, void *ipc_handle_out_64);
int b200_shard_gather_import(b200_shard *s, const void *ipc_handles_world_x_64);
int64_t b200_shard_gather_push(b200_shard *s, const void *dev_src, void *stream);
int b200_shard_gather_wait(b200_shard *s, uint32_t step, void *stream);
/* the per-step form: push this step's logits AND wait for `wait_step` (0 = none) in one launch (programmatic dependent launch: the launch and the
 * wait overlap the forward's tail); returns the published step */
int64_t b200_shard_gather_push_wait(b200_shard *s, const void *dev_src, uint32_t wait_step, void *stream);
/* Fused form: when the forward ends in an fc_chain call that writes `node_name` (AlexNet: fc8), that kernel stores the logits into every rank's
 * gather buffer itself, publishes the step and waits for the previous one -- no gather launch at all, and the captured CUDA graph replays it
 * (the step counter lives in device memory). Returns 1 when attached, 0 when this net does not end that way (keep calling
 * b200_shard_gather_push_wait), negative on error; s == NULL detaches. After the last fused forward, b200_shard_step_from_device() returns the
 * step it published (and re-synchronises the host-side counter that the separate push / wait calls use). */
int b200_fwd_attach_gather(b200_fwd *f, b200_shard *s, const char *node_name);
int64_t b200_shard_step_from_device(b200_shard *s);
int b200_shard_gather_ptr(b200_shard *s, uint32_t step, void **dev_ptr_out); /* device pointer of the local [world][bytes_per_rank] result of `step` */
int b200_shard_all_gather_nccl(b200_shard *s, const void *dev_src, void *dev_dst, uint64_t bytes_per_rank, void *stream);
uint64_t b200_shard_launches(b200_shard *s);                          /* kernels this module has launched so far */

#ifdef __cplusplus
}
#endif
#endif /* BODA_B200_H_ */
