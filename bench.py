#!/usr/bin/env python
"""bench.py -- IA-SSD set-abstraction backbone throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one forward of the full KITTI IA-SSD SA stack (BASELINE.json configs[1]: D-FPS 16384->4096->1024,
ctr-aware top-k ->512->256, vote layer, MSG ball query + shared MLP) over one batch of 16 synthetic
16384-point scenes, eval mode.  Scenes are independent: with N > 1 every rank runs its own batches
(weak scaling, no data-path collective); NCCL carries only the timing reduction.

Prints ONE JSON line (rank 0).  `value` = whole-job scenes/s with inputs resident in HBM (a pool of
distinct batches larger than L2 is cycled); `e2e` = the same metric through `BackbonePipeline.submit_host`
(pinned host input -> H2D -> forward -> D2H of the centre features, every step); `roofline` describes the
dominant kernel of the step (CUDA-event timed in an instrumented eager pass of the same run);
`cpu_baseline` = the CPU oracle port (oracle/) timed on a bounded sample on the host cores.

--impl reference runs the UNMODIFIED reference (its pointnet2_batch CUDA ops rebuilt for sm_100a +
its own pointnet2_modules.py + IASSD_backbone.py, installed by oracle/build_ref.sh into oracle/_ref) on the
same GPU, same weights, same inputs; when that module is unavailable it falls back to the CPU oracle port
(--impl reference-cpu forces the CPU port).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "IA-SSD SA-backbone scenes/s (16k pts)"
UNIT = "scenes/s"
BATCH = 16
NPTS = 16384
NCOLS = 5  # [batch_idx, x, y, z, intensity]
L2_BYTES = 126 * 1024 * 1024
WORKLOAD = ("IA-SSD KITTI cfg full SA stack (D-FPS 16384->4096->1024, ctr-aware top-k ->512->256, vote, "
            "MSG ball query + shared MLP), batch 16 x 16384 pts per GPU, eval")
KIND = "kitti"

# BASELINE.json configs: [1] = kitti (the headline, default), [2] = spsnet (stability-aware top-k), [3] = waymo
WORKLOADS = {
    "kitti": dict(batch=16, npts=16384, ncols=5, kind="kitti", cfg="kitti_iassd_cfg", bb="IASSD_Backbone", text=WORKLOAD),
    "spsnet": dict(batch=16, npts=16384, ncols=5, kind="kitti", cfg="kitti_spsnet_cfg", bb="PAGNet_Backbone",
                   text="SPSNet-IA KITTI cfg: stability-score (sss_aware) top-k sampling in SA layers 2,3 with per-point stds, "
                        "otherwise the IA-SSD SA stack, batch 16 x 16384 pts per GPU, eval"),
    "spsnet_sf": dict(batch=16, npts=16384, ncols=5, kind="kitti", cfg="kitti_spsnet_surface_cfg", bb="PAGNet_Backbone",
                      text="SPSNet-IA backbone exactly as shipped (SPSNet.yaml): stability-score top-k with per-point stds, USE_SURFACE "
                           "(4 DenseEdgeConv units on all 16384 points -> 60 surface channels into the vote layer), 124-wide layer-1 "
                           "MLPs, batch 16 x 16384 pts per GPU, eval"),
    "spsnet_full": dict(batch=16, npts=16384, ncols=5, kind="kitti", cfg="kitti_spsnet_surface_cfg", bb="SPSNET_DET",
                        text="SPSNet-IA detector as shipped (SPSNet.yaml), inference: stability generator (-> stds) -> PAGNet_Backbone with "
                             "stability-score top-k + USE_SURFACE -> MLT_SSD_Head -> decode + rotated-IoU NMS, batch 16 x 16384 pts per GPU"),
    "spsnet_e2e": dict(batch=16, npts=16384, ncols=5, kind="kitti", cfg="kitti_spsnet_cfg", bb="SPSNetIA",
                       text="SPSNet-IA end to end on the path: stability generator (SA layer with M = N = 16384 centres + logvar "
                            "head -> stds) feeding the PAGNet backbone with stability-score top-k, batch 16 x 16384 pts per GPU, eval"),
    "kitti_det": dict(batch=16, npts=16384, ncols=5, kind="kitti", cfg="kitti_iassd_cfg", bb="IASSD_DET",
                      text="IA-SSD KITTI detector, inference: the full SA stack + IASSD_Head (2 x 3-layer FC on 256 centres x 512 ch) + box "
                           "decode + score filter + rotated-IoU NMS -> final boxes (SURVEY.md 8f rank 3), batch 16 x 16384 pts per GPU"),
    "waymo_det": dict(batch=8, npts=65536, ncols=6, kind="waymo", cfg="waymo_iassd_cfg", bb="IASSD_DET",
                      text="IA-SSD Waymo detector, inference: the Waymo SA stack (65536 -> 16384 -> 4096 -> 2048 -> 1024 centres) + IASSD_Head + "
                           "box decode + score filter + rotated-IoU NMS over 1024 boxes per scene, batch 8 x 65536 pts per GPU"),
    "waymo": dict(batch=8, npts=65536, ncols=6, kind="waymo", cfg="waymo_iassd_cfg", bb="IASSD_Backbone",
                  text="IA-SSD Waymo cfg full SA stack (D-FPS 65536->16384->4096, ctr-aware top-k ->2048->1024, vote, MSG ball "
                       "query + shared MLP), batch 8 x 65536 pts per GPU, eval"),
}
_WL = WORKLOADS["kitti"]


def set_workload(name: str):
    global BATCH, NPTS, NCOLS, WORKLOAD, KIND, _WL, METRIC
    _WL = WORKLOADS[name]
    if _WL["bb"] == "IASSD_DET":
        METRIC = "IA-SSD detector (SA backbone + head + NMS) scenes/s (%s pts)" % ("64k" if _WL["kind"] == "waymo" else "16k")
    if _WL["bb"] == "SPSNET_DET":
        METRIC = "SPSNet-IA detector (stability generator + SA backbone + head + NMS) scenes/s (16k pts)"
    BATCH, NPTS, NCOLS, WORKLOAD, KIND = _WL["batch"], _WL["npts"], _WL["ncols"], _WL["text"], _WL["kind"]


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = Path(f"/tmp/spsk_clocks_{os.getpid()}.csv")

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in self.path.read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            self.path.unlink()
        except OSError:
            pass
        return out


def build_net(seed=0):
    from spsnet_b200 import backbone as bb

    torch.manual_seed(seed)
    if _WL["bb"] == "IASSD_DET":
        from spsnet_b200 import detector

        if KIND == "waymo":
            from spsnet_b200 import dense_head as dh

            net = detector.IASSD({"BACKBONE_3D": bb.waymo_iassd_cfg(), "POINT_HEAD": dh.waymo_iassd_head_cfg(),
                                  "POST_PROCESSING": dh.waymo_post_processing()}, num_class=3, input_channels=NCOLS - 1)
        else:
            net = detector.IASSD(num_class=3, input_channels=NCOLS - 1)
    elif _WL["bb"] == "SPSNET_DET":
        from spsnet_b200 import detector
        from spsnet_b200 import stability as st

        net = detector.SPSNetIA(num_class=3, input_channels=NCOLS - 1, generator=st.Generate_center(st.sf_unc_cfg()))
    elif _WL["bb"] == "SPSNetIA":
        from spsnet_b200 import stability as st

        net = st.SPSNetIAFrontEnd(st.Generate_center(st.sf_unc_cfg()),
                                  bb.PAGNet_Backbone(getattr(bb, _WL["cfg"])(), num_class=3, input_channels=NCOLS - 1))
    else:
        net = getattr(bb, _WL["bb"])(getattr(bb, _WL["cfg"])(), num_class=3, input_channels=NCOLS - 1)
    bb.randomize_bn_stats(net, seed=seed)
    return net.eval()


def extra_inputs(device):
    """Per-point stability `stds` for the SPSNet workload (generator output stand-in)."""
    if _WL["bb"] != "PAGNet_Backbone":
        return None
    from spsnet_b200 import scenes

    return {"stds": torch.from_numpy(scenes.make_stds(77, BATCH, NPTS)).to(device)}


def make_pool(rank: int, n_batches: int):
    """n_batches distinct batches in OpenPCDet `points` layout, pinned host tensors."""
    from spsnet_b200 import scenes

    pool = []
    for i in range(n_batches):
        arr = scenes.to_points(scenes.make_batch(100000 * rank + i * BATCH, BATCH, NPTS, KIND))
        pool.append(torch.from_numpy(arr).pin_memory())
    return pool


def cpu_baseline(net_cpu, sample_scenes: int):
    """The oracle port (oracle/oracle.py + oracle.c, all host threads) on a bounded sample of the workload."""
    from oracle import oracle as O
    from spsnet_b200 import scenes

    pts = scenes.make_batch(0, sample_scenes, NPTS, KIND)
    det = hasattr(net_cpu, "point_head")
    backbone = net_cpu.backbone_3d if det else net_cpu

    def run(p):
        out = O.backbone_forward(backbone, p, dtype=torch.float32)
        if det:  # head + decode + post-processing restatement (oracle.py: head_forward, post_processing)
            import types

            from spsnet_b200 import dense_head as dh

            h = net_cpu.point_head
            ns = types.SimpleNamespace(cls_center_layers=h.cls_center_layers, box_center_layers=h.box_center_layers,
                                       mean_size=h.box_coder._mean_np, bin_size=h.box_coder.bin_size)
            centers = np.concatenate([np.repeat(np.arange(p.shape[0]), out["centers"].shape[1])[:, None].astype(np.float32),
                                      out["centers"].reshape(-1, 3)], axis=1)
            cls, _, boxes = O.head_forward(ns, out["centers_features"], centers, dtype=torch.float32)
            pp = dh.waymo_post_processing() if KIND == "waymo" else dh.KITTI_POST_PROCESSING
            nc = pp["NMS_CONFIG"]
            O.post_processing(cls, boxes, p.shape[0], pp["SCORE_THRESH"], nc["NMS_THRESH"], nc["NMS_PRE_MAXSIZE"], nc["NMS_POST_MAXSIZE"])

    run(pts[:1])  # warm (page in, build)
    t0 = time.perf_counter()
    for s0 in range(0, sample_scenes, 8):   # 8 scenes at a time: the torch-CPU conv outputs of a whole sample would take GBs
        run(pts[s0:s0 + 8])
    dt = time.perf_counter() - t0
    cores = max(O.num_threads(), torch.get_num_threads())
    return {"value": sample_scenes / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_scenes} scenes of the same workload (full SA stack, fp32), oracle/oracle.c ops (OpenMP) + torch-CPU "
                      f"conv/BN, {dt:.1f} s; host has {os.cpu_count()} logical cpus"}


# --------------------------------------------------------------------------------------------------
# kernel table: algorithmic work per launch for the roofline object
# --------------------------------------------------------------------------------------------------

def profile_kernels(net, dev_points, steps: int):
    """Instrumented eager pass: CUDA events around every libspsk call on the launching stream."""
    from spsnet_b200 import _lib

    records = []
    originals = {}
    names = [n for n in _lib.SIGNATURES if n not in ("spsk_last_error", "spsk_abi_version", "spsk_built_for_sm", "spsk_launch_count")]

    def wrap(name, fn):
        def inner(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            if name == "spsk_grouped_linear":  # a[0] is byref(GroupDesc): read the row count while it is alive
                g = a[0]._obj
                a = (int(g.b) * int(g.m) * int(g.nsample),) + tuple(a[1:])
            elif name == "spsk_sa_mma_forward":
                g = a[0]._obj
                nl = int(g.nlayers)
                a = (int(g.b) * int(g.m) * int(g.nsample), [int(g.kpad[i]) for i in range(nl)], [int(g.cpad[i]) for i in range(nl)],
                     int(g.cout_last), int(g.split), int(g.c_feat) + (3 if g.use_xyz else 0))
            elif name == "spsk_pw_mma_forward":
                g = a[0]._obj
                a = (int(g.rows), int(g.k), int(g.n))
            elif name.startswith("spsk_ball_query_msg"):
                a = tuple(a[:5]) + ([int(a[5][i]) for i in range(int(a[3]))],)
            elif name == "spsk_detect_postprocess":
                g = a[0]._obj
                a = (int(g.batch), int(g.m), int(g.post_max))
            records.append((name, a, e0, e1))
            return rc
        return inner

    for n in names:
        originals[n] = getattr(_lib.lib, n)
        setattr(_lib.lib, n, wrap(n, originals[n]))
    # per-kernel durations must not include a concurrent kernel: keep the layer-1 FPS prefetch (side stream) off in this pass
    prev_prefetch = os.environ.get("SPSK_FPS_PREFETCH")
    os.environ["SPSK_FPS_PREFETCH"] = "0"
    try:
        with torch.no_grad():
            for i in range(steps):
                d = {"batch_size": BATCH, "points": dev_points[i % len(dev_points)]}
                d.update(extra_inputs("cuda") or {})
                net.forward_padded(d) if hasattr(net, "forward_padded") else net(d)
        torch.cuda.synchronize()
    finally:
        for n in names:
            setattr(_lib.lib, n, originals[n])
        if prev_prefetch is None:
            os.environ.pop("SPSK_FPS_PREFETCH", None)
        else:
            os.environ["SPSK_FPS_PREFETCH"] = prev_prefetch
    table = {}
    for name, a, e0, e1 in records:
        key = name
        if name in ("spsk_farthest_point_sampling",):
            key = f"{name}[n={a[1]},m={a[2]}]"
        elif name == "spsk_ball_query_msg":
            key = f"{name}[n={a[1]},m={a[2]}]"
        elif name == "spsk_grouped_linear":
            key = f"{name}[cin={a[3]},cout={a[6]}]"
        elif name == "spsk_sa_mma_forward":
            key = f"{name}[rows={a[0]},cout={a[3]}{',split' if a[4] else ''}]"
        elif name == "spsk_pw_mma_forward":
            key = f"{name}[rows={a[0]},k={a[1]},n={a[2]}]"
        ms = e0.elapsed_time(e1)
        t = table.setdefault(key, {"ms": 0.0, "launches": 0, "args": a})
        t["ms"] += ms
        t["launches"] += 1
    return table, steps


N_SMS = 148


def _fps_sms(b, n):
    return b * (1 if n <= 16384 else (2 if n <= 32768 else (4 if n <= 65536 else 8)))


def _roof_entry(key, t, peaks, traffic):
    """Roofline of one kernel class: ALGORITHMIC work per launch / mean CUDA-event duration."""
    avg_s = t["ms"] / t["launches"] / 1e3
    a = t["args"]
    hbm = peaks.get("hbm_gbs", 6650.0)          # measured copy bandwidth, GB/s
    tens = peaks.get("bf16_tflops", 1500.0)     # measured dense bf16 (burst: these kernels are timed alone), TFLOP/s
    sms = N_SMS
    e = {"kernel": key, "avg_us": avg_s * 1e6}
    if key.startswith("spsk_farthest_point_sampling"):
        b, n, m = a[0], a[1], a[2]
        sms = _fps_sms(b, n)
        alg_bytes = b * (n * 12 + m * 4)  # read xyz once + write idx
        pairs = b * n * (m - 1)
        lane_peak = N_SMS * 128 * 1.965e9  # fp32 lane-ops/s at max clock
        e.update(bound="hbm", achieved=alg_bytes / avg_s / 1e9, peak=hbm, unit="GB/s",
                 note="FPS is latency-bound (m-1 strictly sequential arg-max steps, one CTA or CTA-cluster per scene): neither "
                      "the HBM nor the tensor roofline applies; see us_per_iter / cycles_per_iter / lane_frac",
                 us_per_iter=avg_s * 1e6 / max(m - 1, 1), cycles_per_iter=avg_s * 1.965e9 / max(m - 1, 1),
                 pair_evals_per_s=pairs / avg_s, lane_frac=pairs * 10 / avg_s / lane_peak)
    elif key.startswith("spsk_sa_mma_forward"):
        rows_n, kpad, cpad, cout_last, split, k0 = a
        widths = [k0] + list(cpad[:-1]) + [cout_last]
        flops = 2.0 * rows_n * sum(x * y for x, y in zip(widths[:-1], widths[1:]))
        e.update(bound="tensor", achieved=flops / avg_s / 1e12, peak=tens, unit="TFLOP/s",
                 note="fused gather + shared MLP + max-pool; algorithmic flops from the true layer widths (split-fp16 chains "
                      "issue 3x these on the tensor cores); peak = measured dense bf16 cuBLAS throughput")
    elif key.startswith("spsk_pw_mma_forward"):
        rows_n, k, n = a
        e.update(bound="tensor", achieved=2.0 * rows_n * k * n / avg_s / 1e12, peak=tens, unit="TFLOP/s",
                 note="point-wise Conv1d GEMM (hi+lo fp16 operands: 3x these flops issued)")
    elif key.startswith("spsk_grouped_linear") or key.startswith("spsk_pointwise_linear"):
        if key.startswith("spsk_grouped_linear"):
            rows_n, cin, cout = a[0], a[3], a[6]
        else:
            rows_n, cin, cout = a[0] * a[1], a[3], a[6]
        e.update(bound="tensor", achieved=2.0 * rows_n * cin * cout / avg_s / 1e12, peak=74.4, unit="TFLOP/s",
                 note="exact-fp32 FFMA path: peak is the fp32 CUDA-core peak (148 SM x 128 lanes x 2 x 1.965 GHz)")
    elif key.startswith("spsk_ball_query_msg"):
        b, n, m = a[0], a[1], a[2]
        nsum = sum(int(x) for x in a[5][:a[3]]) if len(a) > 5 else 48
        alg_bytes = b * (n * 12 + m * 12 + m * nsum * 4)
        e.update(bound="hbm", achieved=alg_bytes / avg_s / 1e9, peak=hbm, unit="GB/s",
                 note="algorithmic bytes = xyz + centres + index lists; the kernel is fp32-issue / latency bound, not HBM bound")
    elif key.startswith("spsk_edge_conv_point"):
        rows_n, cin = int(a[1]), int(a[3])
        e.update(bound="tensor", achieved=2.0 * rows_n * (cin * 24 + 24 * 48) / avg_s / 1e12, peak=74.4, unit="TFLOP/s",
                 note="per-point transform FC + [P|Q|R2|R3] projections of a DenseEdgeConv unit, fp32 FFMA with constant-bank "
                      "weights; peak = fp32 CUDA-core peak")
    elif key.startswith("spsk_edge_conv_aggregate"):
        b, n, k = int(a[1]), int(a[2]), int(a[3])
        e.update(bound="tensor", achieved=2.0 * b * n * k * 432 / avg_s / 1e12, peak=74.4, unit="TFLOP/s",
                 note="per-(point, neighbour) dense edge MLP + max over neighbours, fp32 FFMA (432 MACs per row; repeated padding "
                      "indices are skipped, so this is an upper bound of the work done); peak = fp32 CUDA-core peak")
    elif key.startswith("spsk_detect_postprocess"):
        b, m, post = a
        alg_bytes = b * m * (33 + 3 + 7 + 2) * 4 + b * post * (7 + 1 + 2 + 2) * 4   # logits + encodings + centres in, boxes/scores/labels out
        e.update(bound="hbm", achieved=alg_bytes / avg_s / 1e9, peak=hbm, unit="GB/s",
                 note="3 launches (decode + score sort, IoU bit mask, greedy pass); latency-bound: 256 boxes per scene, the greedy "
                      "pass is sequential over 64-box blocks -- neither roofline applies")
    else:
        e.update(bound="hbm", achieved=None, peak=hbm, unit="GB/s")
    e["frac"] = (e["achieved"] / e["peak"]) if e.get("achieved") else None
    e["sms_used"] = sms
    e["sm_time_ms_per_step"] = None
    tr = traffic.get(key.split("[")[0]) if traffic else None
    e["traffic"] = tr
    e["peak_source"] = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    return e


def roofline_from_table(table, steps, peaks):
    """Rooflines of the step's kernels.  Returned: (dominant, per-kernel rows, list of rooflines).  "Dominant" = largest
    SM-time (duration x SMs held / 148): with several batches in flight the step is bound by SM-time, and the
    latency-bound FPS (16 of 148 SMs) would otherwise hide the kernels that actually fill the GPU."""
    total = sum(t["ms"] for t in table.values())
    traffic = {}
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text())
        except Exception:
            traffic = {}
    rows, roofs = [], []
    for key, t in sorted(table.items(), key=lambda kv: -kv[1]["ms"]):
        rows.append({"kernel": key, "ms_per_step": t["ms"] / steps, "launches_per_step": t["launches"] / steps,
                     "share": t["ms"] / total})
        if t["ms"] / total >= 0.01:
            e = _roof_entry(key, t, peaks, traffic)
            e["sm_time_ms_per_step"] = t["ms"] / steps * e["sms_used"] / N_SMS
            roofs.append(e)
    roofs.sort(key=lambda e: -e["sm_time_ms_per_step"])
    return roofs[0], rows, roofs


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return {}
    return {}


# --------------------------------------------------------------------------------------------------
# arms
# --------------------------------------------------------------------------------------------------

def timed_region(pipe, inputs, steps, warmup, host: bool, world: int):
    """W warm-up steps, then EXACTLY K steps bracketed by barrier + synchronize; device-side events."""
    import torch.distributed as dist

    submit = pipe.submit_host if host else pipe.submit_device
    for i in range(warmup):
        submit(inputs[i % len(inputs)])
    pipe.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    pipe.fork(main)
    t0 = time.perf_counter()
    for i in range(steps):
        submit(inputs[(warmup + i) % len(inputs)])
    pipe.join(main)
    e1.record(main)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    from spsnet_b200 import sharding

    # the only communication of the path: SUM of scenes, MAX of the device-timed milliseconds (spsnet_b200/sharding.py)
    _, ms = sharding.reduce_report(BATCH * steps, e0.elapsed_time(e1), device="cuda")
    return ms, wall


class EagerRef:
    """Adapter giving the reference backbone the same submit/sync surface (no graphs: its forward host-syncs)."""

    def __init__(self, net, outputs=("centers_features", "centers"), extra=None):
        self.net, self.outputs = net, outputs
        self.extra = extra or {}
        self.stream = torch.cuda.current_stream()
        self.host_outs = None
        self.dev_in = torch.zeros((BATCH * NPTS, NCOLS), dtype=torch.float32, device="cuda")
        self.last = None

    def _fwd(self, pts):
        with torch.no_grad():
            out = self.net({"batch_size": BATCH, "points": pts, **self.extra})
        self.last = {k: out[k] for k in self.outputs}

    def submit_device(self, dev_points):
        self._fwd(dev_points)

    def submit_host(self, host_points):
        self.dev_in.copy_(host_points, non_blocking=True)
        self._fwd(self.dev_in)
        if self.host_outs is None or any(self.host_outs[k].shape != v.shape for k, v in self.last.items()):
            # (the reference detector returns variable-length results: fresh pinned mirrors when the shapes change)
            self.host_outs = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in self.last.items()}
        for k, v in self.last.items():
            self.host_outs[k].copy_(v, non_blocking=True)

    def sync(self):
        torch.cuda.synchronize()

    def fork(self, stream):
        pass

    def join(self, stream):
        pass

    def h2d_bytes(self):
        return BATCH * NPTS * NCOLS * 4

    def d2h_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.last.values())


class RefDetector(torch.nn.Module):
    """The reference's own detector pieces, unmodified, wired the way Detector3DTemplate does: IASSD_Backbone ->
    IASSD_Head (pcdet/models/dense_heads/IASSD_head.py) -> post_processing (detector3d_template.py:207-290, the
    class-agnostic branch, restated here around the reference's class_agnostic_nms / nms_gpu because the template class
    itself imports the whole of pcdet)."""

    def __init__(self, backbone, head, post_cfg, nms_utils):
        super().__init__()
        self.backbone_3d, self.point_head = backbone, head
        self.post_cfg, self.nms_utils = post_cfg, nms_utils

    def forward(self, batch_dict):
        batch_dict = self.point_head(self.backbone_3d(batch_dict))
        cfg = self.post_cfg
        boxes_out, scores_out, labels_out = [], [], []
        for index in range(batch_dict["batch_size"]):
            batch_mask = batch_dict["batch_index"] == index
            box_preds = batch_dict["batch_box_preds"][batch_mask]
            cls_preds = torch.sigmoid(batch_dict["batch_cls_preds"][batch_mask])
            cls_preds, label_preds = torch.max(cls_preds, dim=-1)
            label_preds = label_preds + 1
            selected, selected_scores = self.nms_utils.class_agnostic_nms(
                box_scores=cls_preds, box_preds=box_preds, nms_config=cfg.NMS_CONFIG, score_thresh=cfg.SCORE_THRESH)
            boxes_out.append(box_preds[selected])
            scores_out.append(selected_scores)
            labels_out.append(label_preds[selected])
        counts = torch.tensor([b.shape[0] for b in boxes_out], dtype=torch.int32)
        return {"det_boxes": torch.cat(boxes_out), "det_scores": torch.cat(scores_out), "det_labels": torch.cat(labels_out),
                "det_count": counts}


def load_reference_detector(state_dict):
    import importlib
    import types
    import warnings

    from spsnet_b200 import backbone as bb
    from spsnet_b200 import dense_head as dh

    backbone = load_reference_backbone({k[len("backbone_3d."):]: v for k, v in state_dict.items() if k.startswith("backbone_3d.")})
    sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))  # absent dependency of pcdet.utils.common_utils, unused here
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hm = importlib.import_module("pcdet.models.dense_heads.IASSD_head")
        nu = importlib.import_module("pcdet.models.model_utils.model_nms_utils")
    import copy

    hcfg = copy.deepcopy(dict(dh.waymo_iassd_head_cfg()) if KIND == "waymo" else dh.KITTI_IASSD_HEAD)
    hcfg["LOSS_CONFIG"] = {"LOSS_CLS": "WeightedCrossEntropy", "LOSS_REG": "WeightedSmoothL1Loss", "LOSS_INS": "WeightedCrossEntropy",
                           "CORNER_LOSS_REGULARIZATION": False, "CENTERNESS_REGULARIZATION": False, "IOU3D_REGULARIZATION": False,
                           "LOSS_WEIGHTS": {"code_weights": [1.0] * 6}}
    head = hm.IASSD_Head(3, 512, bb.Cfg(hcfg))
    head.load_state_dict({k[len("point_head."):]: v for k, v in state_dict.items() if k.startswith("point_head.")}, strict=False)
    return RefDetector(backbone, head.eval(), bb.Cfg(dh.waymo_post_processing() if KIND == "waymo" else dh.KITTI_POST_PROCESSING), nu).eval()


def load_reference_backbone(state_dict):
    ref_root = ROOT / "oracle" / "_ref"
    so = ref_root / "pcdet" / "ops" / "pointnet2" / "pointnet2_batch" / "pointnet2_batch_cuda.so"
    if not so.exists():
        raise RuntimeError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    import importlib
    import warnings

    sys.path.insert(0, str(ref_root))
    from spsnet_b200 import backbone as bb

    if _WL["bb"] not in ("IASSD_Backbone", "IASSD_DET", "PAGNet_Backbone"):
        raise RuntimeError("the reference arm runs the IASSD_Backbone / PAGNet_Backbone workloads (kitti, waymo, kitti_det, spsnet, spsnet_sf)")
    cls_name = "PAGNet_Backbone" if _WL["bb"] == "PAGNet_Backbone" else "IASSD_Backbone"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("pcdet.models.backbones_3d." + cls_name.replace("_Backbone", "_backbone"))
    net = getattr(mod, cls_name)(getattr(bb, _WL["cfg"])(), num_class=3, input_channels=NCOLS - 1)
    net.load_state_dict(state_dict)
    return net.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--depth", type=int, default=8, help="pipeline slots (streams) of BackbonePipeline: batches in flight; hides the\n                    latency-bound FPS (16 of 148 SMs for 2.5 ms per batch) behind the GEMM-heavy kernels of other batches")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pool", type=int, default=0, help="distinct input batches (0 = enough to exceed L2)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="scenes in the cpu_baseline sample (0 = skip): 64 scenes = 10-15 s of host work")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS), help="kitti = BASELINE.json configs[1] (headline)")
    args = ap.parse_args()
    set_workload(args.workload)
    rank, world, local = dist_env()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference-cpu" or (args.impl == "reference" and not torch.cuda.is_available()):
        return run_reference_cpu(args, rank, world)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    net = build_net()
    state = {k: v.clone() for k, v in net.state_dict().items()}
    n_pool = args.pool or (L2_BYTES // (BATCH * NPTS * NCOLS * 4) + 2)
    host_pool = make_pool(rank, n_pool)
    dev_pool = [t.cuda() for t in host_pool]

    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    cfg = {"workload": WORKLOAD, "batch_per_gpu": BATCH, "points_per_scene": NPTS,
           "l2_policy": f"input pool of {n_pool} distinct batches ({n_pool * BATCH * NPTS * NCOLS * 4 / 2**20:.0f} MiB > 126 MiB L2) cycled",
           "parallelism": f"scene-sharded x{world}, no data-path collective"}

    if args.impl == "reference":
        try:
            ref = (load_reference_detector(state) if _WL["bb"] == "IASSD_DET" else load_reference_backbone(state)).cuda()
        except Exception as e:  # fall back to the CPU port of the oracle
            if rank == 0:
                sys.stderr.write(f"[bench] reference CUDA module unavailable ({e}); using the CPU oracle port\n")
            return run_reference_cpu(args, rank, world)
        torch.backends.cudnn.allow_tf32 = True  # the reference's stock setting
        pipe = EagerRef(ref, outputs=("det_boxes", "det_scores", "det_labels", "det_count") if _WL["bb"] == "IASSD_DET"
                        else ("centers_features", "centers"), extra=extra_inputs("cuda"))
        pipe.submit_device(dev_pool[0])
        pipe.sync()
        sampler = ClockSampler(local)
        sampler.start()
        ms, _ = timed_region(pipe, dev_pool, args.steps, args.warmup, host=False, world=world)
        ms_e2e, _ = timed_region(pipe, host_pool, args.steps, args.warmup, host=True, world=world)
        clocks = sampler.stop()
        value = world * BATCH * args.steps / (ms / 1e3)
        e2e = world * BATCH * args.steps / (ms_e2e / 1e3)
        line.update(impl="reference", value=value, ms_per_step=ms / args.steps, clocks=clocks, gpu_launches=None,
                    e2e={"value": e2e, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes(), "d2h_bytes_per_step": pipe.d2h_bytes()},
                    cpu_baseline={"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                  "sample": "not a CPU run: the reference's own CUDA ops (pointnet2_batch rebuilt for sm_100a) + its "
                                            "pointnet2_modules.py / IASSD_backbone.py on the same B200, cudnn.allow_tf32=True (stock); "
                                            "use --impl reference-cpu for the CPU oracle port"})
        cfg["reference"] = "unmodified reference from oracle/_ref on GPU (eager, as shipped)"
        line["config"] = cfg
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
        return

    # ---- our arm
    from spsnet_b200 import _lib
    from spsnet_b200.runtime import BackbonePipeline

    net = net.cuda()
    outputs = ("det_boxes", "det_scores", "det_labels", "det_count") if _WL["bb"] in ("IASSD_DET", "SPSNET_DET") else ("centers_features", "centers")
    pipe = BackbonePipeline(net, BATCH, NPTS, NCOLS, depth=args.depth, use_graph=not args.no_graph, extra_inputs=extra_inputs("cuda"),
                            outputs=outputs)
    pipe.prepare(dev_pool[0])
    sampler = ClockSampler(local)
    sampler.start()
    ms, wall = timed_region(pipe, dev_pool, args.steps, args.warmup, host=False, world=world)
    ms_e2e, _ = timed_region(pipe, host_pool, args.steps, args.warmup, host=True, world=world)
    clocks = sampler.stop()
    value = world * BATCH * args.steps / (ms / 1e3)
    e2e = world * BATCH * args.steps / (ms_e2e / 1e3)
    line.update(impl="ours", value=value, ms_per_step=ms / args.steps, clocks=clocks,
                gpu_launches=int(pipe.launches_per_step * args.steps),
                e2e={"value": e2e, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes(), "d2h_bytes_per_step": pipe.d2h_bytes()})
    cfg.update(pipeline_depth=args.depth, cuda_graph=not args.no_graph, launches_per_step=int(pipe.launches_per_step))

    if rank == 0 and not args.no_profile:
        table, psteps = profile_kernels(net, dev_pool, steps=min(args.steps, 5))
        roof, rows, roofs = roofline_from_table(table, psteps, peaks)
        line["roofline"] = roof
        line["rooflines"] = roofs[:8]
        line["kernels"] = rows[:12]
        line["eager_ms_per_step_sum_of_kernels"] = sum(r["ms_per_step"] for r in rows)
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        try:
            line["cpu_baseline"] = cpu_baseline(build_net(), args.cpu_sample)
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    line["config"] = cfg
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def run_reference_cpu(args, rank, world):
    """CPU oracle port as the reference arm (only when the rebuilt reference cannot run): rank 0 only."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    net = build_net()
    from oracle import oracle as O
    from spsnet_b200 import scenes

    sample = 2
    pts = [scenes.make_batch(i * sample, sample, NPTS, KIND) for i in range(2)]
    for _ in range(min(args.warmup, 1)):
        O.backbone_forward(net, pts[0][:1], dtype=torch.float32)
    steps = min(args.steps, 5)
    t0 = time.perf_counter()
    for i in range(steps):
        O.backbone_forward(net, pts[i % 2], dtype=torch.float32)
    dt = time.perf_counter() - t0
    v = sample * steps / dt
    cores = max(O.num_threads(), torch.get_num_threads())
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"each step = {sample} scenes of the workload on the host CPU"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} scenes per step x {steps} steps, oracle/oracle.c (OpenMP) + torch-CPU conv/BN"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
