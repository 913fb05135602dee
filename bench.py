#!/usr/bin/env python
"""bench.py -- AlexNet-ng forward images/sec (BASELINE.json metric, config C2: nets/alexnet_ng_conv batch=32 fp32) on N B200s.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                   (the reference arm: the path's CPU implementation on host cores)

A step = one whole-net forward of one batch (32 images per GPU) of synthetic NCHW input with hash-synthetic weights.
`value`   : device-resident whole-job throughput (inputs already in HBM), CUDA-event timed per step on the back-end's
            stream with an L2 flush between steps, max over ranks.
`e2e`     : the same metric through the reference-facing call (has_conv_fwd_t::run_fwd via the C ABI) with HOST pinned
            buffers: H2D of the batch and D2H of the logits inside the timed region of every step.
`roofline`: the dominant kernel (the tcgen05 implicit-GEMM contraction behind every Convolution): algorithmic conv FLOPs
            (2*B*OC*OH*OW*IC*KH*KW, src/latex-util.H:116-120) / its CUDA-event launch durations, vs the measured bf16 peak.
`cpu_baseline` / `--impl reference`: the oracle port of the reference's operator semantics (Boda itself cannot be built
            here: boost/protobuf/python2/Caffe missing) on all host cores, on a bounded sample of the same workload (>= 10 s of CPU
            work). `cpu_baseline.torch_cpu` adds SURVEY 8(d)'s second figure, torch's CPU operators (oneDNN), from a child process.
Multi-GPU: images are independent units of the path, so ranks shard the batch dimension (weak scaling, 32 images per
GPU); weights are broadcast once from rank 0 over NCCL at init (b200_shard_broadcast) and every step's logits are gathered
into every rank's peer-mapped buffer over NVLink, awaited one step late: inside the forward's last kernel when the net ends
in an inner-product chain (b200_fwd_attach_gather), else with one gather launch per step (b200_shard_gather_push_wait).
Each form is checked against the NCCL all-gather of the same data before it is timed. B200_BENCH_GATHER=peer_launch forces
the one-launch form, =nccl the host-synchronised NCCL all-gather of round 1, =none times the loop without any gather (A/B only).
Also run by default: `other_configs` (BASELINE C4 GoogLeNet B=64 bf16, C5 ResNet-50 B=32/GPU fp32) and a >= 2 s `sustained` run
(N=1); --no-other-configs / --no-cpu-baseline skip them.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "alexnet_ng_conv_fwd_images_per_sec"
UNIT = "images/s"
PER_GPU_BATCH = 32
GATHER_FORM = ["b200_shard_gather_push_wait: one launch per step"]  # how the multi-GPU timed loop gathered the logits (set by bench_one)
L2_FLUSH_BYTES = 256 << 20


def committed_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/roofline_traffic.json, written by
    tools/ncu_summarise.py runs); None when no capture of this build's kernel is committed."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    return d.get("dram_bytes_per_launch"), d.get("source")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML every few ms while the timed regions run."""

    def __init__(self, dev_index):
        super().__init__(daemon=True)
        self.dev_index, self.stop_flag, self.samples, self.reasons, self.max_mhz = dev_index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
                r = int(get(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.003)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


NET_NAME = "alexnet_ng_conv"
NET_IN_SZ = 227


class CpuReference:
    """The reference arm / cpu_baseline: the oracle port of Boda's operator semantics, whole-net forward on the host cores (OpenMP over
    ALL of them: the team size is set explicitly, because torch.distributed.run exports OMP_NUM_THREADS=1 to its children). The net text,
    the hash-synthetic parameters and the input are built ONCE here, outside every timed region. The one place bench.py executes oracle/."""

    def __init__(self, n_images):
        from boda_b200 import nets
        from oracle import boda_oracle as bo, net_oracle
        self.n_images, self.bo, self.net_oracle = n_images, bo, net_oracle
        bo.set_num_threads(os.cpu_count() or 1)
        self.txt, self.in_node, self.out_node = nets.NETS[NET_NAME](n_images)
        self.params = nets.synth_params(self.txt)
        self.x = nets.synth_input((n_images, 3, NET_IN_SZ, NET_IN_SZ))
        self.cores = bo.num_threads()

    def forward(self):
        return self.net_oracle.run_pipe(self.txt, {self.in_node: self.x}, self.params)


def cpu_reference_forward(n_images, reps=1, min_seconds=0.0, max_reps=16):
    """cpu_baseline leg: at least `reps` forwards, then until min_seconds of CPU work. Returns (images/s, cores, seconds, forwards)."""
    ref = CpuReference(n_images)
    t0 = time.perf_counter()
    done = 0
    while done < reps or (time.perf_counter() - t0 < min_seconds and done < max_reps):
        ref.forward()
        done += 1
    dt = time.perf_counter() - t0
    return n_images * done / dt, ref.cores, dt, done


def torch_cpu_forward(n_images, min_seconds=3.0, check=False):
    """Second CPU figure of SURVEY section 8(d): the same net through torch's CPU operators (oneDNN convolutions) on all host cores, in the
    same run. A LIBRARY number for scale only -- not the reference's algorithm, not the reported baseline (that is the oracle port above).
    Walks the same conv_pipe text. Returns (images/s, threads, seconds, forwards) [+ the output node when check=True]."""
    import torch
    import torch.nn.functional as F
    from boda_b200 import nets
    from oracle import boda_oracle as bo
    txt, i, o = nets.NETS[NET_NAME](n_images)
    params = {k: torch.from_numpy(v) for k, v in nets.synth_params(txt).items()}
    x0 = torch.from_numpy(nets.synth_input((n_images, 3, NET_IN_SZ, NET_IN_SZ)))
    ops = []
    for line in txt.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        d = bo.parse_lexp(line)
        if "node" in d:
            continue
        nv = {k: bo._nda_from_lexp(v) for k, v in (d.get("nda_vals") or {}).items()}
        ops.append((d["str_vals"]["type"], d["tag"], [b for b in d["bots"].split(":") if b], [t for t in d["tops"].split(":") if t], nv))

    def yx(nv, name, dflt):
        return (nv[name].dims["y"], nv[name].dims["x"]) if name in nv else dflt

    def fwd():
        nodes = {i: x0}
        for typ, tag, bots, tops, nv in ops:
            x = nodes[bots[0]]
            if typ in ("Convolution", "InnerProduct"):
                y = F.conv2d(x, params[tag + "_filts"], params.get(tag + "_biases"), stride=yx(nv, "stride", (1, 1)), padding=yx(nv, "in_pad", (0, 0)))
            elif typ == "BatchNorm":
                sf = float(params[tag + "_sf"].reshape(-1)[0])
                sc = 0.0 if sf == 0 else 1.0 / sf
                eps = float(nv["eps"].v) if "eps" in nv else 1e-5
                y = (x - (params[tag + "_mean"] * sc)[None, :, None, None]) / torch.sqrt(params[tag + "_var"] * sc + eps)[None, :, None, None]
            elif typ == "Scale":
                y = x * params[tag + "_gamma"][None, :, None, None] + params[tag + "_beta"][None, :, None, None]
            elif typ == "ReLU":
                y = F.relu(x)
            elif typ == "Dropout":
                y = x
            elif typ == "LRN":
                y = F.local_response_norm(x, int(nv["local_size"].v), float(nv["alpha"].v), float(nv["beta"].v), float(nv["k"].v))
            elif typ == "Pooling":
                avg = bool(int(nv["avg_pool"].v)) if "avg_pool" in nv else False
                if "kern_sz" not in nv:
                    y = x.mean(dim=(2, 3), keepdim=True) if avg else x.amax(dim=(2, 3), keepdim=True)
                elif avg:
                    y = F.avg_pool2d(x, yx(nv, "kern_sz", None), yx(nv, "stride", (1, 1)), yx(nv, "in_pad", (0, 0)), ceil_mode=True, count_include_pad=False)
                else:
                    y = F.max_pool2d(x, yx(nv, "kern_sz", None), yx(nv, "stride", (1, 1)), yx(nv, "in_pad", (0, 0)), ceil_mode=True)
            elif typ == "Concat":
                y = torch.cat([nodes[b] for b in bots], dim=1)
            elif typ in ("Eltwise", "Reduce"):
                y = sum(nodes[b] for b in bots)
            elif typ == "Softmax":
                y = F.softmax(x, dim=1)
            else:
                raise ValueError("torch_cpu_forward: unhandled op type " + typ)
            nodes[tops[0]] = y
        return nodes[o]

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with torch.no_grad():
        out = fwd()  # warm-up (oneDNN primitive creation, weight reorders)
        t0 = time.perf_counter()
        done = 0
        while done < 2 or (time.perf_counter() - t0 < min_seconds and done < 64):
            out = fwd()
            done += 1
        dt = time.perf_counter() - t0
    res = (n_images * done / dt, threads, dt, done)
    return res + (out.numpy(),) if check else res


def run_reference_arm(args, rank, world):
    """`--impl reference`: rank 0 alone times the path's CPU implementation (the oracle port; Boda itself is not buildable here) on all host
    cores; the other ranks exit without work. Each step = one forward of a bounded sample of the batch, sized from a calibration forward so
    that the whole warm-up + K-step run stays within ~150 s of wall time. Setup (net text, parameters, input) is outside the timed loop."""
    if rank != 0:
        return
    budget_s = float(os.environ.get("B200_REF_BUDGET_S", "150"))
    probe_n = min(4, PER_GPU_BATCH)
    probe = CpuReference(probe_n)
    probe.forward()  # library load, page faults
    t0 = time.perf_counter()
    probe.forward()
    per_img = (time.perf_counter() - t0) / probe_n
    steps_total = max(args.steps + args.warmup, 1)
    sample = int(max(1, min(PER_GPU_BATCH, budget_s / steps_total / max(per_img, 1e-6))))
    ref = probe if sample == probe_n else CpuReference(sample)
    for _ in range(args.warmup):
        ref.forward()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.forward()
    dt = time.perf_counter() - t0
    cores = ref.cores
    val = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "nets/%s fwd, batch=%d per GPU, fp32, %dx%d%s" % (NET_NAME, PER_GPU_BATCH, NET_IN_SZ, NET_IN_SZ, " (BASELINE configs[1])" if NET_NAME == "alexnet_ng_conv" else ""),
                       "parallelism": "host cores only"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "each step = %s forward of %d images (a bounded sample of the %d-image batch; parameters and input synthesised once, outside the timed loop) through the oracle port, OpenMP on all %d host cores; Boda's own binary / Caffe CPU path is not buildable here" % (NET_NAME, sample, PER_GPU_BATCH, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def bind_to_gpu_numa_node(dev_index):
    """Run this rank's host threads (and so first-touch its pinned buffers) on the NUMA node its GPU hangs off: the e2e path moves 19.8 MB per
    step per GPU from host memory, and at 4-8 ranks a remote-socket source halves it. Best effort: silently does nothing without sysfs / NVML."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(dev_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def _emit(line: dict):
    """Exactly one JSON line on the real stdout (fd 1 is pointed at stderr while libraries such as NCCL are chatty)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def bench_one(args, rank, local_rank, world, dist, sh_nccl, is_main):
    """One configuration (args.net / args.batch / args.prec) on this process's GPU: device-resident steps, end-to-end steps, roofline of the
    dominant kernel; the CPU baseline only for the main configuration. Returns the JSON line as a dict on rank 0 (None elsewhere)."""
    import torch
    import boda_b200 as bb
    from boda_b200 import nets
    global NET_NAME, NET_IN_SZ, METRIC
    NET_NAME = args.net
    NET_IN_SZ = 224 if args.net in ("googlenet_conv", "resnet50") else 227
    METRIC = "%s_fwd_images_per_sec" % args.net
    B = args.batch
    txt, in_node, out_node = nets.NETS[args.net](B)
    extra = os.environ.get("B200_FWD_OPTS", "")  # e.g. "use_2cta=0,use_graph=0" for A/B experiments
    fwd = bb.B200ConvFwd(txt, "(prec=%s,device=%d%s)" % (args.prec, local_rank, ("," + extra) if extra else ""))

    # ---- weights: rank 0 synthesises; ONE NCCL broadcast of a flat DEVICE buffer over NVLink (north_star: "single NCCL broadcast of weights"),
    # issued through the C ABI (b200_shard_broadcast); every rank then slices it into its parameter vars device-to-device
    from boda_b200 import shard
    shapes = nets.conv_param_shapes(txt)
    sh = None
    if world > 1:
        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        sh = bb.B200Shard(local_rank, rank, world)  # this configuration's gather buffers (the NCCL communicator of sh_nccl serves every configuration)
        layout = shard.param_layout(shapes)
        total = layout[-1][1] + layout[-1][2]
        flat = torch.empty(total, dtype=torch.float32, device="cuda")
        if rank == 0:
            params0 = nets.synth_params(txt)
            host = np.empty(total, np.float32)
            for n, off, sz, shape in layout:
                host[off:off + sz] = np.asarray(params0[n], np.float32).ravel()
            flat.copy_(torch.from_numpy(host))
        torch.cuda.synchronize()
        sh_nccl.broadcast(flat.data_ptr(), total * 4, 0, 0)
        torch.cuda.synchronize()
        for n, off, sz, shape in layout:
            fwd.set_param_device(n, flat.data_ptr() + 4 * off, sz)
        del flat
    else:
        params = nets.synth_params(txt)
        for n in sorted(shapes):
            fwd.set_param(n, params[n])

    # ---- inputs: each rank owns its shard of the global batch (images [rank*B, (rank+1)*B)), pinned host memory
    x_host = torch.from_numpy(nets.synth_input((B, 3, NET_IN_SZ, NET_IN_SZ), seed=rank)).pin_memory()
    x_host2 = torch.from_numpy(nets.synth_input((B, 3, NET_IN_SZ, NET_IN_SZ), seed=rank + 1000)).pin_memory()
    logits_host = torch.empty((B, 1000, 1, 1), dtype=torch.float32).pin_memory()
    logits_host2 = torch.empty((B, 1000, 1, 1), dtype=torch.float32).pin_memory()
    in_elems, out_elems = x_host.numel(), logits_host.numel()
    xs, ls = (x_host, x_host2), (logits_host, logits_host2)

    def e2e_step():
        fwd.run_fwd_ptrs([in_node], [x_host.data_ptr()], [in_elems], [out_node], [logits_host.data_ptr()], [out_elems])

    def e2e_pipelined(steps):
        """Serving loop over the public submit()/wait() API: batch i+1's H2D overlaps batch i's forward; every batch's input is copied
        from pinned host memory and its logits are read back to the host inside the timed region."""
        prev = None
        for i in range(steps):
            t = fwd.submit_ptrs([in_node], [xs[i & 1].data_ptr()], [in_elems], [out_node], [ls[i & 1].data_ptr()], [out_elems])
            if prev is not None:
                fwd.wait(prev)
            prev = t
        fwd.wait(prev)

    # the logits node as a torch view of the back-end's device var, for the NCCL gather
    class _DevView:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
    n_logits = int(np.prod(fwd.node_dims(out_node)[1:]))
    logits_dev = torch.as_tensor(_DevView(fwd.node_device_ptr(out_node), (B, n_logits)), device="cuda")
    gathered = None
    if world > 1:  # every rank's gather buffer [world][B * n_logits] is mapped by its peers (CUDA IPC): logits are written into it over NVLink by one small kernel
        sh.gather_setup(B * n_logits * 4, exchange)

    sampler = ClockSampler(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also: first eager pass, weight packing, CUDA-graph capture)
    for _ in range(args.warmup):
        e2e_step()
    e2e_pipelined(args.warmup)
    fwd.run_timed(args.warmup, L2_FLUSH_BYTES)
    if dist:  # one checked gather: every rank must see every rank's logits, image order = rank order
        st_ptr = fwd.stream_ptr()
        fwd.enqueue()
        step = sh.gather_push(fwd.node_device_ptr(out_node), st_ptr)
        sh.gather_wait(step, st_ptr)
        barrier()
        ref = torch.empty((world * B, n_logits), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(ref, logits_dev.contiguous())
        torch.cuda.synchronize()
        gathered = torch.as_tensor(_DevView(sh.gather_ptr(step), (world * B, n_logits)), device="cuda")
        if not torch.equal(ref, gathered):
            raise RuntimeError("peer-memory logits gather disagrees with the NCCL all-gather of the same data")
        # the per-step form of the timed loop (push step i + wait for step i-1 in one launch), checked the same way
        fwd.enqueue()
        step2 = sh.gather_push_wait(fwd.node_device_ptr(out_node), step, st_ptr)
        sh.gather_wait(step2, st_ptr)
        torch.cuda.synchronize()
        gathered2 = torch.as_tensor(_DevView(sh.gather_ptr(step2), (world * B, n_logits)), device="cuda")
        if not torch.equal(ref, gathered2):
            raise RuntimeError("peer-memory logits gather (push + wait in one launch) disagrees with the NCCL all-gather of the same data")
    barrier()

    sampler.start()
    # ---- timed region 1: device-resident steps (value)
    launches0 = fwd.launches()
    barrier()
    if world == 1:
        ms_each = fwd.run_timed(args.steps, L2_FLUSH_BYTES)
        dev_ms = float(sum(ms_each))
    elif os.environ.get("B200_BENCH_GATHER", "peer") == "nccl":
        # A/B: the round-1 form -- forward (events on the back-end's stream), then the NCCL all-gather of its logits, host-synchronised
        dev_ms = 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(args.steps):
            dev_ms += fwd.run_timed(1, L2_FLUSH_BYTES)[0]
            ev0.record()
            shard.gather_logits(dist, logits_dev, out=ref)
            ev1.record()
            ev1.synchronize()
            dev_ms += ev0.elapsed_time(ev1)
    else:
        # Serving pipeline on ONE stream (the back-end's): step i = forward i, push of its logits into every rank's gather buffer, then the wait
        # for the logits of step i-1 from all ranks (one step late, so a rank never idles on its slowest peer inside a step). K forwards +
        # K gathers = K+1 event-bracketed windows (the first has no wait, the last no forward), the L2 flush sits between windows.
        s_fwd = torch.cuda.ExternalStream(fwd.stream_ptr(), device=torch.device("cuda", local_rank))
        st_ptr = fwd.stream_ptr()
        out_ptr = fwd.node_device_ptr(out_node)
        # Nets that end in an inner-product chain (AlexNet: fc8) gather INSIDE the chain's kernel: its last layer's CTAs store the logits into every
        # rank's buffer, the last CTA publishes the step and waits for the previous one (b200_fwd_attach_gather) -- no gather launch, and part of
        # the captured graph. Checked against the NCCL all-gather before it is timed. B200_BENCH_GATHER=peer_launch keeps the separate launch.
        fused = os.environ.get("B200_BENCH_GATHER", "peer") == "peer" and fwd.attach_gather(sh, out_node)
        fused_all = shard.max_over_ranks(dist, 0.0 if fused else 1.0, device="cuda") == 0.0  # every rank or none
        if fused and not fused_all:
            fwd.attach_gather(None, out_node)
        fused = fused and fused_all
        GATHER_FORM[0] = ("inside the fc_chain kernel of the forward: b200_fwd_attach_gather, no gather launch" if fused
                          else "b200_shard_gather_push_wait: one launch per step")
        if fused:
            with torch.cuda.stream(s_fwd):
                for _ in range(3):  # warm-up forward, capture, one replay: all ranks the same count
                    fwd.enqueue()
                s_fwd.synchronize()
            step_f = sh.step_from_device()
            sh.gather_wait(step_f, st_ptr)
            torch.cuda.synchronize()
            barrier()
            dist.all_gather_into_tensor(ref, logits_dev.contiguous())
            torch.cuda.synchronize()
            gathered_f = torch.as_tensor(_DevView(sh.gather_ptr(step_f), (world * B, n_logits)), device="cuda")
            if not torch.equal(ref, gathered_f):
                raise RuntimeError("in-kernel logits gather (fc_chain) disagrees with the NCCL all-gather of the same data")
            barrier()
        windows = []
        sh_launches0 = sh.launches()
        torch.cuda.synchronize()
        with torch.cuda.stream(s_fwd):
            prev_step = None
            for i in range(args.steps + 1):
                fwd.flush_l2(L2_FLUSH_BYTES)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                step = None
                if fused:
                    if i < args.steps:
                        fwd.enqueue()  # forward i + push of its logits + wait for step i-1, one graph
                elif i < args.steps:
                    fwd.enqueue()
                    if os.environ.get("B200_BENCH_GATHER") != "none":  # ("none": experiment -- the same loop without any gather)
                        step = sh.gather_push_wait(out_ptr, prev_step or 0, st_ptr)  # one launch: push step i, wait for step i-1
                elif prev_step is not None:
                    sh.gather_wait(prev_step, st_ptr)
                prev_step = step
                ev1.record()
                windows.append((ev0, ev1))
            s_fwd.synchronize()
            if fused:  # the last step's gather completes with an explicit wait (outside the windows' forwards, inside the timed sum)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                last = sh.step_from_device()
                ev0.record()
                sh.gather_wait(last, st_ptr)
                ev1.record()
                windows.append((ev0, ev1))
                s_fwd.synchronize()
        dev_ms = float(sum(a.elapsed_time(b) for a, b in windows))
        shard_launches = sh.launches() - sh_launches0
        if fused:
            fwd.attach_gather(None, out_node)  # the e2e and profile passes below run the plain forward ...
            fwd.run_timed(3, 0)                # ... whose graph is captured again here, outside every timed region
            torch.cuda.synchronize()
    barrier()
    launches = fwd.launches() - launches0 + (shard_launches if world > 1 and os.environ.get("B200_BENCH_GATHER", "peer") != "nccl" else 0)
    if dist:
        dev_ms = shard.max_over_ranks(dist, dev_ms, device="cuda")

    # ---- a sustained run beside the burst figure: the same device-resident steps (same per-step events, same L2 flush) for >= 2 s of GPU time, so the
    # whole-step number is also seen under sustained clocks / power (the K-step region above is only a few milliseconds long)
    sustained = None
    if is_main and world == 1:
        per = max(dev_ms / max(args.steps, 1), 1e-3)
        n_sus = int(min(20000, max(args.steps, 2000.0 / per)))
        s_smp = ClockSampler(local_rank)
        s_smp.start()
        t0 = time.perf_counter()
        sus_ms = float(sum(fwd.run_timed(n_sus, L2_FLUSH_BYTES)))
        wall = time.perf_counter() - t0
        s_smp.stop_flag = True
        s_smp.join(timeout=2)
        sustained = {"value": B * n_sus / (sus_ms * 1e-3), "unit": UNIT, "steps": n_sus, "ms_per_step": sus_ms / n_sus, "gpu_seconds": sus_ms * 1e-3, "wall_seconds": wall,
                     "clocks": s_smp.summary()}

    # ---- timed region 2: end to end through run_fwd with host buffers (e2e)
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(min(args.steps, 20)):
        e2e_step()
    e2e_sync_ms = 1e3 * (time.perf_counter() - t0) / min(args.steps, 20)
    if dist:
        e2e_s = shard.max_over_ranks(dist, e2e_s, device="cuda")
    barrier()

    # ---- roofline of the dominant kernel: per-launch CUDA events around every contraction kernel (eager profile pass)
    prof = fwd.profile(max(5, min(args.steps, 20)))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    conv_rows = [r for r in prof if r[3] > 0]
    per_op = nets.conv_algorithmic_elems(txt, per_op=True)  # tag -> ((in + out + filts + biases) elements, output pixels per image)
    def _tag(func):
        return func.split("__")[1]
    # the DOMINANT kernel = the persistent CTA-pair contraction kernel (igemm2.cuh) behind every convolution with a spatial output; the
    # inner-product-shaped layers (1x1 output: M = batch rows) run the one-CTA split-K kernel and are bound by streaming their weights from HBM
    dom_rows = [r for r in conv_rows if per_op.get(_tag(r[0]), (0, 2))[1] > 1]
    fc_rows = [r for r in conv_rows if per_op.get(_tag(r[0]), (0, 2))[1] <= 1]
    conv_flops = sum(r[3] for r in dom_rows)
    conv_bytes = 4.0 * sum(per_op.get(_tag(r[0]), (0, 0))[0] for r in dom_rows)  # each tensor once, fp32 (src/latex-util.H:119)
    conv_kernel_ms = sum(r[2] for r in dom_rows)
    peaks = measured_peaks()
    achieved_tf = conv_flops / (conv_kernel_ms * 1e-3) / 1e12 if conv_kernel_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops"]
    # (an fc_chain call runs several inner-product layers: its bytes are those of every layer in the chain)
    chains = {c[0]: c for c in bb.fwd_plan(txt, "(prec=%s%s)" % (args.prec, ("," + extra) if extra else "")).get("fcchain", [])}
    def _fc_tags(func):
        return chains.get(_tag(func), [_tag(func)]) if func.startswith("fc_chain__") else [_tag(func)]
    fc_bytes = 4.0 * sum(per_op.get(t, (0, 0))[0] for r in fc_rows for t in _fc_tags(r[0]))
    fc_chained = any(r[0].startswith("fc_chain__") for r in fc_rows)
    fc_ms = sum(r[2] for r in fc_rows)
    all_flops, all_ms = sum(r[3] for r in conv_rows), sum(r[2] for r in conv_rows)

    if rank == 0:
        global_batch = B * world
        value = global_batch * args.steps / (dev_ms * 1e-3)
        e2e_val = global_batch * args.steps / e2e_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32 (fp16 hi/lo split, 3 tcgen05.mma per k-step, fp32 accumulate)", "fp16": "f16", "bf16": "bf16"}[args.prec], "data": "synthetic",
            "config": {"workload": "nets/%s fwd, batch=%d per GPU, %s, %dx%d%s" % (args.net, B, args.prec, NET_IN_SZ, NET_IN_SZ, " (BASELINE configs[1])" if (args.net, B, args.prec) == ("alexnet_ng_conv", 32, "fp32") else ""), "global_batch": global_batch,
                       "parallelism": ("batch-shard x%d, one NCCL broadcast of the weights at init (device-resident) + per-step logits gather into every rank's peer-mapped buffer "
                                       "over NVLink (%s, awaited one step late)" % (world, GATHER_FORM[0])) if world > 1 else "single GPU",
                       "l2": "256 MiB scratch buffer overwritten before every timed step (outside the events)", "cuda_graph": True},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": in_elems * 4, "d2h_bytes_per_step": out_elems * 4, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "b200_fwd_submit / b200_fwd_wait (pipelined run_fwd, depth 2: H2D of batch i+1 overlaps the forward of batch i)",
                    "sync_run_fwd_ms_per_step": e2e_sync_ms},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                         "traffic": committed_traffic()[0] if (args.net, args.prec) == ("alexnet_ng_conv", "fp32") else None, "traffic_source": committed_traffic()[1],
                         "algorithmic_flops_per_launch": conv_flops / max(len(dom_rows), 1), "algorithmic_bytes_per_launch": conv_bytes / max(len(dom_rows), 1),
                         "avg_launch_ms": conv_kernel_ms / max(len(dom_rows), 1),
                         "kernel": "b200::igemm_umma_2cta_kernel (persistent CTA pairs, tcgen05 cta_group::2): %d launches per forward, one per Convolution with a spatial output" % len(dom_rows),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernels timed one by one with CUDA events), %s" % peaks["source"],
                         "mma_passes": 3 if args.prec == "fp32" else 1,
                         "tensor_work_frac": (achieved_tf * (3 if args.prec == "fp32" else 1) / peak_tf) if peak_tf else None,
                         "note": "achieved = algorithmic conv FLOPs / summed CUDA-event launch durations of the dominant kernel's launches; the fp32-parity mode issues "
                                 "3 fp16 MMAs per product (mma_passes), so tensor_work_frac = achieved x mma_passes / peak is the share of the tensor pipe's peak actually kept busy",
                         "all_contraction_launches": {"launches": len(conv_rows), "achieved": all_flops / (all_ms * 1e-3) / 1e12 if all_ms > 0 else None, "unit": "TFLOP/s"},
                         "fc_shaped_layers": {"bound": "hbm", "launches": len(fc_rows),
                                              "kernel": ("b200::fc_chain_kernel (all inner-product layers of a chain in one persistent launch; weights as the 128-row operand, split-K, "
                                                         "grid barriers between the layers)" if fc_chained else "b200::igemm_umma_kernel (weights as the 128-row operand, split-K)"),
                                              "achieved": fc_bytes / (fc_ms * 1e-3) / 1e9 if fc_ms > 0 else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                              "frac": (fc_bytes / (fc_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if fc_ms > 0 else None}},
            "clocks": sampler.summary(),
            "sustained": sustained,
            "per_call": [{"func": r[0], "call_ms": round(r[1], 5), "kernel_ms": round(r[2], 5), "gflop": round(r[3] / 1e9, 3)} for r in prof],
        }
        if is_main and not args.no_cpu_baseline and world == 1:
            v, cores, secs, n_fwd = cpu_reference_forward(B, reps=2, min_seconds=10.0)  # a bounded sample: >= 10 s of CPU work, at most 16 forwards
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d x %s forward of the full %d-image batch through the oracle port (OpenMP, all host cores), %.1f s" % (n_fwd, args.net, B, secs)}
            try:  # torch's CPU operators (oneDNN) on the same box in the same run -- a library figure, for scale only.
                # In a child process (no GPU, its own thread pools, bounded time) so that nothing it does can cost the bench line.
                import subprocess
                env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
                cp_ = subprocess.run([sys.executable, os.path.abspath(__file__), "--torch-cpu-child", "--net", args.net, "--batch", str(B)], env=env, capture_output=True, text=True, timeout=180)
                tc = json.loads(cp_.stdout.strip().splitlines()[-1])
                line["cpu_baseline"]["torch_cpu"] = {"value": tc["value"], "unit": UNIT, "cores": tc["cores"], "kind": "library (torch CPU / oneDNN; not the reference's algorithm)",
                                                     "sample": "%d x %s forward of the full %d-image batch, %.1f s" % (tc["forwards"], args.net, B, tc["seconds"])}
            except Exception as e:  # never let the extra figure break the bench line
                line["cpu_baseline"]["torch_cpu"] = {"unavailable": repr(e)[:200]}
        return line
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C4 / C5 lines that the default run appends as `other_configs`")
    ap.add_argument("--prec", default="fp32", choices=["fp32", "fp16", "bf16"])
    ap.add_argument("--net", default="alexnet_ng_conv", choices=["alexnet_ng_conv", "nin_imagenet", "googlenet_conv", "resnet50"],
                    help="default = BASELINE configs[1]; googlenet_conv --batch 64 --prec bf16 = configs[3]; resnet50 --batch 32 (x8 GPUs = 256) = configs[4]")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step")
    ap.add_argument("--torch-cpu-child", action="store_true", help=argparse.SUPPRESS)  # internal: the torch-CPU figure of cpu_baseline, run as a child process
    args = ap.parse_args()
    global NET_NAME, NET_IN_SZ, METRIC
    NET_NAME = args.net
    NET_IN_SZ = 224 if args.net in ("googlenet_conv", "resnet50") else 227
    METRIC = "%s_fwd_images_per_sec" % args.net
    if args.torch_cpu_child:
        v, cores, secs, n = torch_cpu_forward(args.batch)
        _emit({"value": v, "cores": cores, "seconds": secs, "forwards": n})
        return
    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import boda_b200 as bb
    from boda_b200 import nets

    if bb.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the B200 back-end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = None if os.environ.get("B200_BENCH_NO_NUMA_BIND") else bind_to_gpu_numa_node(local_rank)  # before any pinned allocation
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sh_nccl = None
    if world > 1:
        def exchange0(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        sh_nccl = bb.B200Shard(local_rank, rank, world)
        sh_nccl.nccl_init(exchange0)
    line = bench_one(args, rank, local_rank, world, dist, sh_nccl, True)
    # ---- the other BASELINE configurations, measured after the headline region with the same event timing (fewer steps), so that they are
    # driver-visible numbers: C4 = nets/googlenet_conv batch 64 bf16, C5 = nets/resnet-50 batch 32 per GPU (x8 GPUs = 256) fp32
    others = {}
    if (args.net, args.batch, args.prec) == ("alexnet_ng_conv", PER_GPU_BATCH, "fp32") and not args.no_other_configs:
        for key, net, batch, prec in (("C4_googlenet_conv_b64_bf16", "googlenet_conv", 64, "bf16"), ("C5_resnet50_b32_per_gpu_fp32", "resnet50", 32, "fp32")):
            a2 = argparse.Namespace(**vars(args))
            a2.net, a2.batch, a2.prec, a2.steps, a2.warmup = net, batch, prec, min(args.steps, 20), max(3, min(args.warmup, 5))
            try:
                l2 = bench_one(a2, rank, local_rank, world, dist, sh_nccl, False)
                if l2 is not None:
                    others[key] = {k: l2[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "dtype", "config", "e2e", "gpu_launches", "clocks") if k in l2}
                    others[key]["roofline"] = {k: l2["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "mma_passes") if k in l2["roofline"]}
            except Exception as e:  # never let an extra configuration cost the headline line
                if rank == 0:
                    others[key] = {"unavailable": repr(e)[:300]}
    if rank == 0:
        if others:
            line["other_configs"] = others
        _emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
