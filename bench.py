#!/usr/bin/env python
"""Benchmark of the geometry hot path (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path on the box's host cores

One "step" = the reference's ChamferEMD reconstruction loss (src/train/metrics_and_losses.py:70-79) forward AND
backward w.r.t. the reconstruction on one batch of B=32 cloud pairs of 2048 points per GPU (BASELINE.json configs
[0]+[2], "Chamfer+EMD fwd/bwd clouds/sec (B=32,N=2048)").  The batch is sharded over ranks (weak scaling: 32 clouds
per GPU); the only collective is one all-reduce of the loss.  The kNN graph throughput (configs[1]) and the
Chamfer-only / EMD-only throughputs are reported in the same JSON line under "sub_metrics" with their rooflines.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

B_PER_GPU = 32
GLOBAL_BATCH = 32  # strong scaling: BASELINE's B=32 split over the GPUs
N_POINTS = 2048
KNN_N, KNN_K, KNN_C = 1024, 20, 64
METRIC = "chamfer_emd_fwd_bwd_clouds_per_sec"
WORKLOAD = ("ChamferEMD recon loss fwd+bwd (pykeops_chamfer + match_cost), B=32 x 2048 xyz points per GPU, "
            "S1 synthetic ShapeNet-shaped clouds (BASELINE configs[0]+[2]); kNN graphs/s k=20 N=1024 C=3/64 "
            "(configs[1]) in sub_metrics")


# ----------------------------------------------------------------------------------------------------------------
def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--eager", action="store_true", help="time the steps with eager launches instead of GraphedLossStep")
    p.add_argument("--no-pipeline", action="store_true", help="end-to-end leg without the copy-stream pipelining")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--no-sub", action="store_true", help="skip sub-metrics (Chamfer-only, EMD-only, kNN)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                   help="weak: 32 clouds per GPU (default); strong: the global batch of 32 clouds split over the GPUs")
    return p.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks() -> dict:
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            d = json.loads(f.read_text())
            peaks.update(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], sm_max_mhz=d.get("sm_max_mhz"),
                         source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return peaks


def measure_pipe_peaks() -> dict:
    """FP32 / MUFU pipe peaks measured on this box by tools/pipe_peaks (built in-tree); derived figures otherwise."""
    derived = {"ffma_tflops": 148 * 128 * 2 * 1.965e9 / 1e12, "mufu_ex2_gops": 148 * 16 * 1.965e9 / 1e9,
               "source": "derived: 148 SM x 128 FMA lanes (16 SFU lanes) x 1.965 GHz"}
    exe = ROOT / "tools" / "pipe_peaks"
    if not exe.exists():
        return derived
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        d["ffma_tflops"] = max(d["ffma_tflops"], d.get("ffma2_tflops", 0.0))
        d["source"] = "measured on this GPU by tools/pipe_peaks (FFMA/FFMA2 and MUFU.EX2 saturation loops)"
        return d
    except Exception:
        return derived


# ----------------------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    from pointcloudcounterfactual_b200 import _lib, edgeconv, losses, neighbour_ops, sharding, synthetic
    from pointcloudcounterfactual_b200.structural_losses import match_cost
    from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import NNDistance

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl b200) needs CUDA GPUs; there is no CPU fallback")
    rank, world, local = sharding.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    K, W = args.steps, max(args.warmup, 3)
    global B_PER_GPU
    if args.scaling == "strong":
        if GLOBAL_BATCH % world:
            raise RuntimeError(f"--scaling strong needs a GPU count that divides {GLOBAL_BATCH}")
        B_PER_GPU = GLOBAL_BATCH // world

    # identical bits on CPU and GPU: generated on the CPU, this rank's slice of the global batch
    recon_h, ref_h = synthetic.s1_near(B_PER_GPU, N_POINTS, first=rank * B_PER_GPU)
    recon_h, ref_h = recon_h.pin_memory(), ref_h.pin_memory()
    recon_d, ref_d = recon_h.to(dev), ref_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The reported global mean is ONE all-reduce of (sum, count) per logging interval -- here: per timed region, issued and
    # waited for INSIDE it -- not one per step: sharding.LossAccumulator (the reference's drytorch metrics work the same way).
    loss_acc = sharding.LossAccumulator(dev)
    graphed = graphed_dev = None
    if not args.eager:
        try:  # the public fixed-shape step: [H2D copies +] loss forward/backward + D2H of the loss as ONE graph launch
            graphed = losses.GraphedLossStep(losses.chamfer_emd, recon_h, ref_h, dev)
            graphed_dev = losses.GraphedLossStep(losses.chamfer_emd, recon_d, ref_d, dev)
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            graphed = graphed_dev = None
            print(f"[bench] graph capture refused ({type(e).__name__}: {e}); eager steps", file=sys.stderr)

    def step_device():
        if graphed_dev is not None:
            _, grad = graphed_dev()
            loss = graphed_dev.loss_device
        else:
            r = recon_d.detach().requires_grad_(True)
            loss = losses.chamfer_emd(r, ref_d)
            (grad,) = torch.autograd.grad(loss.sum(), r)
        loss_acc.add(loss)
        return loss, grad

    def step_e2e():
        if graphed is not None:
            local, grad = graphed()          # pinned host clouds -> device -> loss + gradient -> loss on the host
            loss = graphed.loss_device
        else:
            r = recon_h.to(dev, non_blocking=True).requires_grad_(True)
            t = ref_h.to(dev, non_blocking=True)
            loss = losses.chamfer_emd(r, t)
            (grad,) = torch.autograd.grad(loss.sum(), r)
            local = loss.cpu()
        loss_acc.add(loss)
        torch.cuda.current_stream(dev).synchronize()  # the step's result is on the host now
        host_value = float(local[0])
        return host_value, grad

    def timed(fn, steps, warm):
        """per-step CUDA events on the launching stream; L2 flushed between steps outside the events; the global mean of
        the interval's losses is reduced over the ranks inside the last step's events."""
        for _ in range(warm):
            fn()
        loss_acc.reduce()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i, (e0, e1) in enumerate(evs):
            flush.zero_()
            e0.record()
            fn()
            if i == steps - 1:
                timed.last_mean = loss_acc.reduce()  # the one collective of the interval
            e1.record()
        barrier()
        timed.last_mean = float(timed.last_mean)
        return [e0.elapsed_time(e1) for e0, e1 in evs]

    def reduce_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pipe = measure_pipe_peaks() if rank == 0 else {}
    peaks = load_peaks()
    barrier()

    # ---- headline: device-resident ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    times = timed(step_device, K, W)
    launches = (_lib.launch_count() - l0) * K // (K + W)
    if graphed_dev is not None:  # replays do not pass through the C entry points: count what the capture recorded
        launches = graphed_dev.kernels_per_replay * K
    total_ms = reduce_max(sum(times))
    value = world * B_PER_GPU * K / (total_ms * 1e-3)

    # ---- e2e: pinned host inputs -> device -> loss back on the host, every step ------------------------------
    e2e_pipe = None
    if graphed is not None and not args.no_pipeline:
        try:  # the same graphed step fed through a copy stream: step i+1's clouds travel while step i computes
            e2e_pipe = losses.PipelinedLossStep(losses.chamfer_emd, recon_h, ref_h, dev)
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            print(f"[bench] pipelined e2e refused ({type(e).__name__}: {e}); unpipelined graph", file=sys.stderr)

    def run_e2e(nsteps):
        if e2e_pipe is None:
            for _ in range(nsteps):
                step_e2e()
            return float(loss_acc.reduce().cpu())  # the interval's global mean arrives on the host inside the timed region
        e2e_pipe.prefetch()  # H2D of the first step's clouds
        host_values = []
        for s_i in range(nsteps):
            e2e_pipe.wait_prefetch()  # the copy that reads the pinned buffers has finished: refill them for the next step
            nxt_r, nxt_t = host_batches[(s_i + 1) & 1]
            recon_h.copy_(nxt_r)
            ref_h.copy_(nxt_t)
            prev = e2e_pipe.step()  # launches this step (graph), starts the next step's H2D, returns the previous loss
            loss_acc.add(e2e_pipe.loss_device)
            if prev is not None:
                host_values.append(float(prev[0]))  # the previous step's per-cloud loss, read on the host
        host_values.append(float(e2e_pipe.drain()[0]))
        assert len(host_values) == nsteps
        return float(loss_acc.reduce().cpu())  # the interval's global mean, on the host inside the timed region

    # two host batches alternate in the pinned buffers (a data loader's role): every step's inputs are rewritten on the host
    host_batches = [(recon_h.clone(), ref_h.clone()),
                    tuple(t.clone() for t in synthetic.s1_near(B_PER_GPU, N_POINTS, first=(world + rank) * B_PER_GPU))]
    run_e2e(W)
    barrier()
    t0 = time.perf_counter()
    run_e2e(K)
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_value = world * B_PER_GPU * K / e2e_s
    h2d = recon_h.numel() * 4 + ref_h.numel() * 4
    d2h = B_PER_GPU * 4 + 4

    # ---- dominant kernel: one full approxmatch sweep (18 of the 26 per step) --------------------------------------
    def ev_time(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ones = torch.ones(B_PER_GPU, N_POINTS, device=dev)
    ratio = torch.empty(B_PER_GPU, N_POINTS, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def sweep():
        _lib.check(lib.pcc_approxmatch_sweep(B_PER_GPU, N_POINTS, N_POINTS, recon_d.data_ptr(), ref_d.data_ptr(),
                                             ones.data_ptr(), ones.data_ptr(), ratio.data_ptr(), -16.0, 0, st), "sweep")

    sweep_ms = ev_time(sweep, 50)
    pairs = B_PER_GPU * N_POINTS * N_POINTS
    mufu_peak = pipe.get("mufu_ex2_gops", 148 * 16 * 1.965) if rank == 0 else 1.0
    fp32_peak = pipe.get("ffma_tflops", 74.4) if rank == 0 else 1.0
    roofline = {
        "kernel": "am_sweep_kernel (approxmatch solver sweep: 26 sweeps per step in 18 launches, 8 of them two sweeps fused; the 8 "
                  "sweeps of the three steepest levels run as culled kernels that skip exactly-zero partners, 18 as this one)",
        "bound": "sfu", "unit": "Gexp/s", "achieved": pairs / (sweep_ms * 1e-3) / 1e9, "peak": mufu_peak,
        "frac": pairs / (sweep_ms * 1e-3) / 1e9 / mufu_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one am_sweep_kernel launch, ncu --set full capture
        # (profiles/r01b_ncu_full_summary.md): the clouds and the scaling vectors, everything else stays on chip
        "traffic": 2108416, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one am_sweep_kernel launch, "
                                              "profiles/r02_ncu_full_summary.md (ncu --set full of this command)",
        "hbm_view": {"algorithmic_bytes": 3 * B_PER_GPU * N_POINTS * 4 * 2 + 3 * B_PER_GPU * N_POINTS * 4,
                     "achieved_gbs": 2108416 / (sweep_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                     "frac": 2108416 / (sweep_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "note": "not HBM-bound: 2 MB per launch against 134 M exponentials"},
        "peak_source": pipe.get("source", ""), "launch_ms": sweep_ms,
        # the sweep's own instruction mix (8 packed FMA-pipe + 2 MUFU.EX2 per partner and thread) on otherwise idle SMs:
        # what the hardware sustains when both pipes are fed together, with many warps and with the 2 warps per SM
        # sub-partition that 65 536 points / (32 lanes x 2 points) = 1024 warps leave on 592 sub-partitions
        "mix_view": None if "emd_mix_gexp_per_s" not in pipe else {
            "mix_peak_gexp_s": pipe["emd_mix_gexp_per_s_2warps_per_smsp"],
            "mix_peak_many_warps_gexp_s": pipe["emd_mix_gexp_per_s"],
            "frac_of_mix_peak": pairs / (sweep_ms * 1e-3) / 1e9 / pipe["emd_mix_gexp_per_s_2warps_per_smsp"],
            "note": "tools/pipe_peaks k_emd_mix (no shared-memory loads, independent accumulators); with the kernel's "
                    "broadcast LDS.128 and ordered accumulation tools/emd_mix_probe measures 3101 Gexp/s at 2 warps per "
                    "sub-partition, x 1.73/2 occupancy quantisation = 2682: the kernel is at 97 % of that"},
        "fp32_tflops": 11 * pairs / (sweep_ms * 1e-3) / 1e12, "fp32_frac": 11 * pairs / (sweep_ms * 1e-3) / 1e12 / fp32_peak,
        "share_of_step": 18 * sweep_ms / (sum(times) / K),
        "algorithmic_unit": "exp-pair evaluations: B*n*m = 134.2 M per sweep launch (DESIGN.md section 4)",
        "note": "1 MUFU.EX2 + 11 flop per pair; the SFU pipe (16 lanes/SM) bounds the kernel, not HBM or tensor cores; "
                "share_of_step counts the 18 full sweeps as single-sweep durations (the fused launches run two sweeps in "
                "~1.7 of them); the culled sweeps of levels j = 7, 6, 5 and the cost/gradient kernel are the rest",
    }

    # ---- sub-metrics ----------------------------------------------------------------------------------------
    sub = {}
    if not args.no_sub:
        def graph_or_eager(fn, reps=50):
            """CUDA-graph replay of one op (removes Python launch latency); eager if capture is refused."""
            try:
                fn()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                ms = ev_time(g.replay, reps)
                return ms, True
            except Exception as e:  # noqa: BLE001
                torch.cuda.synchronize()
                print(f"[bench] graph capture refused ({type(e).__name__}); eager timing", file=sys.stderr)
                return ev_time(fn, reps), False

        rr = recon_d.detach().requires_grad_(True)

        def chamfer_fb():
            loss = losses.pykeops_chamfer(rr, ref_d)
            torch.autograd.grad(loss.sum(), rr)

        def chamfer_f():
            NNDistance(recon_d, ref_d)

        def emd_fb():
            loss = match_cost(rr, ref_d)
            torch.autograd.grad(loss.sum(), rr)

        x3 = synthetic.knn_xyz(B_PER_GPU, KNN_N, first=rank * B_PER_GPU).to(dev)
        xf = synthetic.knn_features(B_PER_GPU, KNN_C, KNN_N, seed=2000 + rank).to(dev)
        x25 = synthetic.knn_xyz(B_PER_GPU, N_POINTS, first=rank * B_PER_GPU).to(dev)

        f64a_ = synthetic.knn_features(B_PER_GPU, 64, N_POINTS, seed=3000 + rank).to(dev)

        ms, gr = graph_or_eager(chamfer_f)
        flops = 8.0 * 2 * B_PER_GPU * N_POINTS * N_POINTS
        sub["chamfer_fwd"] = {"ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr,
                              "roofline": {"bound": "fp32", "unit": "TFLOP/s", "achieved": flops / (ms * 1e-3) / 1e12,
                                           "peak": fp32_peak, "frac": flops / (ms * 1e-3) / 1e12 / fp32_peak,
                                           "gpairs_per_s": flops / 8 / (ms * 1e-3) / 1e9}}
        ms, gr = graph_or_eager(chamfer_fb)
        sub["chamfer_fwd_bwd"] = {"ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr,
                                  "roofline": {"bound": "fp32", "unit": "TFLOP/s", "achieved": flops / (ms * 1e-3) / 1e12,
                                               "peak": fp32_peak, "frac": flops / (ms * 1e-3) / 1e12 / fp32_peak}}
        ms, gr = graph_or_eager(emd_fb, reps=10)
        sub["emd_fwd_bwd"] = {
            "ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr,
            # the reference's solver evaluates 27 sweeps x B*n*m exponentials; here 26 run (the last one only fills the
            # scratch vector) and 8 of them skip the partners whose exponential is exactly zero
            "algorithmic_gexp_per_s": 27 * pairs / (ms * 1e-3) / 1e9,
            "note": "27*B*n*m exp-pair evaluations of the reference's solver / this time (cost+gradient kernel included in "
                    "the time, its 6 special-function ops per pair not counted); MUFU.EX2 peak in roofline.peak"}
        tf32_peak = peaks["bf16_tflops"] / 2.0  # dense TF32 = half the bf16 tensor rate (nominal 1.1 vs 2.25 PFLOP/s)
        for name, x, k, c in (("knn_xyz_k20_n1024", x3, KNN_K, 3), ("knn_feat64_k20_n1024", xf, KNN_K, KNN_C),
                              ("knn_xyz_k25_n2048", x25, 25, 3), ("knn_xyz_k4_n2048", x25, 4, 3)):
            ms, gr = graph_or_eager(lambda x=x, k=k: neighbour_ops.knn(x, k))
            n = x.shape[2]
            if c == 3:
                fl = 8.0 * B_PER_GPU * n * n
                rl = {"bound": "fp32", "unit": "TFLOP/s", "achieved": fl / (ms * 1e-3) / 1e12, "peak": fp32_peak,
                      "frac": fl / (ms * 1e-3) / 1e12 / fp32_peak,
                      "note": "warp-cooperative exact fp32 kernel (references in registers); 8 flop per pair, selection "
                              "(threshold, compaction, ranking) is not counted as work"}
            else:
                fl = 2.0 * c * B_PER_GPU * n * n  # algorithmic: one -2 X X^T contraction
                rl = {"bound": "tensor", "unit": "TFLOP/s", "achieved": fl / (ms * 1e-3) / 1e12, "peak": tf32_peak,
                      "frac": fl / (ms * 1e-3) / 1e12 / tf32_peak, "peak_source": "MEASURED_PEAKS bf16 / 2 (TF32)",
                      "note": "tcgen05 kind::tf32 candidate generator (issues the contraction twice) + exact fp32 re-rank "
                              "from the shared-memory key tiles; time includes the transpose/norm prep launch"}
            sub[name] = {"ms": ms, "graphs_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr, "roofline": rl}

        # ---- the same operators reached the way the reference's UNCHANGED scripts reach them: install() registers the KeOps
        # shim, and the reference's own expressions (src/utils/neighbour_ops.py:35-40,77-82: transpose + LazyTensor on (x, x)
        # + argKmin; src/train/metrics_and_losses.py:32-41: two argmin of one expression + gathers) run on top of it --------
        from pointcloudcounterfactual_b200 import install as pcc_install

        pcc_install.install(patch_reference=False)
        from pykeops.torch import LazyTensor  # the shim

        def ref_pykeops_square_distance(t1, t2):
            return ((LazyTensor(t1[:, :, None, :]) - LazyTensor(t2[:, None, :, :])) ** 2).sum(-1)

        def ref_pykeops_knn(x, k):
            x = x.transpose(2, 1).contiguous()
            return ref_pykeops_square_distance(x, x).argKmin(k, dim=2)

        def ref_pykeops_chamfer(t1, t2):
            dist = ref_pykeops_square_distance(t1, t2)
            idx1 = dist.argmin(axis=1).expand(-1, -1, t1.shape[2])
            squared1 = ((t2 - t1.gather(1, idx1)) ** 2).sum(2).mean(1)
            idx2 = dist.argmin(axis=2).expand(-1, -1, t1.shape[2])
            squared2 = ((t1 - t2.gather(1, idx2)) ** 2).sum(2).mean(1)
            return squared1 + squared2

        def chamfer_fb_shim():
            torch.autograd.grad(ref_pykeops_chamfer(rr, ref_d).sum(), rr)

        for name, x, k in (("knn_xyz_k20_n1024", x3, KNN_K), ("knn_feat64_k20_n1024", xf, KNN_K)):
            r0 = _lib.route_counts()
            ms, gr = graph_or_eager(lambda x=x, k=k: ref_pykeops_knn(x, k))
            r1 = _lib.route_counts()
            sub[name + "_via_install"] = {
                "ms": ms, "graphs_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr, "direct_ms": sub[name]["ms"],
                "ratio_to_direct": ms / sub[name]["ms"],
                "kernel_families": sorted(k2 for k2 in r1 if r1[k2] > r0[k2]),
                "note": "the reference's pykeops_knn body (transpose(2,1).contiguous() + LazyTensor expression + argKmin) on "
                        "the KeOps shim: includes torch's transposition, reaches the point-major self-kNN route"}
        r0 = _lib.route_counts()
        ms, gr = graph_or_eager(chamfer_fb_shim)
        r1 = _lib.route_counts()
        sub["chamfer_fwd_bwd_via_install"] = {
            "ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr,
            "direct_ms": sub["chamfer_fwd_bwd"]["ms"], "ratio_to_direct": ms / sub["chamfer_fwd_bwd"]["ms"],
            "kernel_families": sorted(k2 for k2 in r1 if r1[k2] > r0[k2]),
            "note": "the reference's pykeops_chamfer BODY on the KeOps shim: one fused nn_sym launch serves both argmin calls, "
                    "the gathers, differences, means and their autograd backward stay torch's (about 20 small kernels); the "
                    "post-import hook of install() replaces the whole function by the fused operator = the direct number"}

        # ---- EdgeConv front-end: get_graph_features forward / backward, HBM-bound -----------------------
        idx25 = neighbour_ops.knn(f64a_, 25)
        fa = f64a_.detach().requires_grad_(True)
        gf_bytes = 2 * 64 * B_PER_GPU * N_POINTS * 25 * 4  # the (B,2C,N,k) tensor: written once fwd, read once bwd

        def gf_fwd():
            neighbour_ops.get_graph_features(f64a_, idx25, 25)

        gfeat = torch.ones((B_PER_GPU, 128, N_POINTS, 25), device=dev)

        def gf_fwd_bwd():  # forward and backward inside one capture (a backward of a graph built outside cannot be captured)
            feat = neighbour_ops.get_graph_features(fa, idx25, 25)[1]
            torch.autograd.grad(feat, fa, gfeat)

        ms_f, gr_f = graph_or_eager(gf_fwd, reps=10)
        ms_fb, gr_fb = graph_or_eager(gf_fwd_bwd, reps=10)
        for nm, ms, gr in (("graph_features_c64_n2048_k25_fwd", ms_f, gr_f),
                           ("graph_features_c64_n2048_k25_bwd", ms_fb - ms_f, gr_f and gr_fb)):
            sub[nm] = {"ms": ms, "cuda_graph": gr,
                       "roofline": {"bound": "hbm", "unit": "GB/s", "achieved": gf_bytes / (ms * 1e-3) / 1e9,
                                    "peak": peaks["hbm_gbs"], "frac": gf_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "algorithmic_bytes": gf_bytes,
                                    "note": "the (B,2C,N,k) feature tensor crosses HBM once; x and idx stay in L2"
                                            + ("; backward = (forward+backward) - forward" if nm.endswith("bwd") else "")}}
        del gfeat

        # ---- decoder-output smoothing: kNN k=4 + fused graph_filtering forward and backward -------------
        xg = x25.detach().requires_grad_(True)

        def gfilt():
            out = neighbour_ops.graph_filtering(xg, 4)
            torch.autograd.grad(out, xg, out)

        ms, gr = graph_or_eager(gfilt, reps=20)
        sub["graph_filtering_n2048_fwd_bwd"] = {"ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr,
                                                "note": "kNN k=4 (knn3w_kernel) + one launch forward + one launch backward"}

        # ---- BASELINE configs[3] / [4]: the reference's models cannot be instantiated here (drytorch / hydra are not
        # installed), so these are the hot-path op sequences of one training step / one latent-optimisation
        # iteration, per GPU, with the step's real collective ------------------------------------------------------
        grad_buf = torch.zeros(45 * (1 << 20) // 4, device=dev)  # ~45 MB of fp32 autoencoder gradients

        # ---- auction EMD (external/emd): clouds in [0,1]^3, eps = 0.005, 50 rounds -------------------------
        from pointcloudcounterfactual_b200.emd import emdModule

        ua, uc = (t.to(dev) for t in synthetic.auction_clouds(B_PER_GPU, N_POINTS))
        auction = emdModule()
        ms = ev_time(lambda: auction(ua, uc, 0.005, 50), 5, warm=2)
        sub["auction_emd_fwd_n2048_eps0.005_iters50"] = {
            "ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": False,
            "note": "one launch: a 4-CTA thread-block cluster per cloud runs all rounds (the reference launches 7 kernels "
                    "per round; its CUDA kernels recompiled for sm_100a take 6.07 ms on the same GPU, tools/ref_cuda_compare.py)"}

        # ---- fused EdgeConv layer (full row) and the DGCNN edge-convolution stack built from it ----------
        class _EdgeConv(torch.nn.Module):
            """Attribute names and forward of the reference's EdgeConvLayer (src/module/layers.py:159-203)."""

            def __init__(self, cin, cout, act):
                super().__init__()
                self.dense = torch.nn.Conv2d(cin, cout, kernel_size=1, bias=False)
                self.bn = torch.nn.BatchNorm2d(cout)
                self.act, self.residual = act, False

            def forward(self, x):
                y = self.bn(self.dense(x))
                return self.act(y) if self.act is not None else y

        torch.manual_seed(7)
        h_dim = (64, 64, 128, 256)  # DGCNN.h_dim (src/module/encoders.py:36); first layer without activation (:37)
        enc = torch.nn.ModuleList(
            [_EdgeConv(6, h_dim[0], None)] +
            [_EdgeConv(2 * i, o, torch.nn.LeakyReLU(0.2, inplace=True)) for i, o in zip(h_dim[:-1], h_dim[1:])]).to(dev)
        enc_params = [p for p in enc.parameters()]
        x_enc = x25.detach().requires_grad_(True)

        def encoder(fused: bool):
            xs, h = [], x_enc
            for layer in enc:
                if fused:
                    h = edgeconv.fused_edge_conv(layer, h, torch.empty(0), 25)[1]
                else:  # the reference's three lines (encoders.py:49-54) on top of the fused gather
                    h = layer(neighbour_ops.get_graph_features(h, torch.empty(0), 25)[1]).max(dim=3, keepdim=False)[0]
                xs.append(h)
            return torch.cat(xs, dim=1)  # (B, 512, N), the input of DGCNN.final_conv

        def encoder_fb(fused: bool = True):
            feat = encoder(fused)
            torch.autograd.grad(feat, [x_enc] + enc_params, feat)

        def layer_fb(fused: bool):
            x = f64a_.detach().requires_grad_(True)
            if fused:
                out = edgeconv.fused_edge_conv(enc[1], x, idx25, 25)[1]
            else:
                out = enc[1](neighbour_ops.get_graph_features(x, idx25, 25)[1]).max(dim=3, keepdim=False)[0]
            torch.autograd.grad(out, [x] + list(enc[1].parameters()), out)

        ms, gr = graph_or_eager(lambda: layer_fb(True), reps=10)
        ms_t = ev_time(lambda: layer_fb(False), 5)
        edge_bytes = B_PER_GPU * N_POINTS * 25 * 64 * 4  # one 64-channel row of u per edge, gathered from L2
        sub["edgeconv_layer_c64_cout64_n2048_k25_fwd_bwd"] = {
            "ms": ms, "cuda_graph": gr, "torch_composition_ms": ms_t, "speedup_vs_torch_composition": ms_t / ms,
            "note": "kNN given; graph features -> Conv2d 1x1 -> BatchNorm2d (batch statistics) -> LeakyReLU -> max over k, "
                    "forward and backward w.r.t. input, weight, gamma, beta.  Fused: conv applied to the points (one GEMM), "
                    f"edge pass gathers {edge_bytes / 1e6:.0f} MB of rows from L2; torch composition: the reference's op "
                    "sequence (cudnn TF32 conv on the (B,2C,N,k) tensor) on top of the fused gather"}
        ms, gr = graph_or_eager(lambda: encoder_fb(True), reps=10)
        ms_t = ev_time(lambda: encoder_fb(False), 3)
        tf32_was = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True  # what the reference's cudnn convolution uses by default
        ms_tf32 = graph_or_eager(lambda: encoder_fb(True), reps=10)[0]
        torch.backends.cuda.matmul.allow_tf32 = tf32_was
        sub["dgcnn_edgeconv_stack_fwd_bwd"] = {
            "ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr, "torch_composition_ms": ms_t,
            "speedup_vs_torch_composition": ms_t / ms, "ms_with_tf32_point_gemms": ms_tf32,
            "note": "the four chained EdgeConv layers of DGCNN (3->64->64->128->256, N=2048, k=25, dynamic kNN per layer), "
                    "forward and backward; torch composition = same kNN and gather kernels, reference op sequence after "
                    "(its cudnn convolution runs TF32 by default; the fused path's point GEMMs are fp32 unless "
                    "torch.backends.cuda.matmul.allow_tf32 is set -- ms_with_tf32_point_gemms)"}

        # gradient exchange of the training step: 45 MB of fp32, bucketed the way DistributedDataParallel would bucket the
        # reference's autoencoder (25 MB cap, reverse parameter order).  Where the parameters are (configs/experiment/
        # autoencoder/model: w_dim 1024, PCGen with 8 components of conv_dims [1024, 256, 16]): decoder 10.5 M parameters =
        # 42 MB, encoder final_conv 0.5 M = 2 MB, the four EdgeConv layers 0.09 M = 0.36 MB.  The decoder's gradients are
        # complete once its backward is (loss -> graph_filtering -> MLPs), i.e. BEFORE the encoder's backward starts: buckets
        # 0 (25 MB) and 1 (19.6 MB: rest of the decoder + final_conv) are all-reduced on a communication stream under the
        # encoder's backward, and only bucket 2 (the EdgeConv weights, 0.4 MB) is exposed at the end.  The whole step -- kernels
        # of this library, torch glue and the NCCL calls -- is captured as ONE CUDA graph per rank when the capture is accepted
        n_all = grad_buf.numel()
        n0, n2 = int(n_all * 25.0 / 45.0), int(n_all * 0.4 / 45.0)
        buckets = [grad_buf[:n0], grad_buf[n0:n_all - n2], grad_buf[n_all - n2:]]
        comm = torch.cuda.Stream(dev)

        def reduce_bucket(i):
            if world > 1:
                comm.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(comm):
                    dist.all_reduce(buckets[i])

        def ae_step():
            loss = losses.chamfer_emd(rr, ref_d)  # ChamferEMD forward + backward: the decoder's gradients come first
            torch.autograd.grad(loss.sum(), rr)
            gfilt()  # decoder graph_filtering (kNN k=4 + smoothing forward / backward); the decoder MLPs would follow
            reduce_bucket(0)
            reduce_bucket(1)
            encoder_fb(True)  # encoder: per layer kNN graph (k=25) + fused EdgeConv layer, forward and backward
            reduce_bucket(2)
            if world > 1:
                torch.cuda.current_stream(dev).wait_stream(comm)

        if world > 1:  # NCCL communicator warm-up outside any capture
            for i in range(3):
                reduce_bucket(i)
            torch.cuda.current_stream(dev).wait_stream(comm)
            torch.cuda.synchronize()
        ms, ae_graphed = graph_or_eager(ae_step, reps=10)
        sub["ae_step_hotpath"] = {
            "ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": ae_graphed,
            "note": "stand-in for configs[3]: DGCNN edge-convolution stack (4 layers, dynamic kNN k=25, fused EdgeConv) "
                    "forward and backward + decoder graph_filtering (kNN k=4) fwd+bwd + ChamferEMD fwd+bwd"
                    + (" + NCCL all-reduce of 45 MB fp32 gradients in DDP-style buckets (25 + 19.6 MB = decoder and final_conv, "
                       "reduced on a communication stream under the encoder's backward; 0.4 MB of EdgeConv weights exposed)"
                       if world > 1 else "") + "; 32 clouds per GPU.  "
                    "Not included (plain torch layers of the reference): final_conv, the PCGen decoder MLPs, optimizer"}

        leaf = recon_d.detach().clone().requires_grad_(True)
        opt = torch.optim.Adam([leaf], lr=1e-3, capturable=True)

        def generate_iter():
            smooth = neighbour_ops.graph_filtering(leaf.transpose(1, 2).contiguous(), 4).transpose(1, 2)
            loss = losses.pykeops_chamfer(smooth.contiguous(), ref_d)
            opt.zero_grad(set_to_none=False)
            loss.sum().backward()
            opt.step()

        ms, gr = graph_or_eager(generate_iter, reps=20)
        sub["generate_loop_iter"] = {
            "ms": ms, "clouds_per_s": world * B_PER_GPU / (ms * 1e-3), "cuda_graph": gr,
            "note": "synthetic stand-in for configs[4] (the reference's generate.py has no optimisation loop): graph_filtering "
                    "(kNN k=4) + Chamfer, forward and backward w.r.t. the cloud, + Adam on a (32,2048,3) leaf per GPU"}

    # ---- strong scaling (BASELINE.md protocol): the GLOBAL batch of 32 clouds split over the ranks; reported in
    # every weak-scaling run beside the headline (the driver launches the default mode only) --------------------------
    strong = None
    if args.scaling == "weak" and GLOBAL_BATCH % world == 0:
        bs = GLOBAL_BATCH // world
        lo = rank * bs
        g_strong = losses.GraphedLossStep(losses.chamfer_emd, recon_d[:bs].contiguous(), ref_d[:bs].contiguous(), dev) \
            if world > 1 else graphed_dev
        if g_strong is not None:
            def step_strong():
                g_strong()
                loss_acc.add(g_strong.loss_device)

            st_times = timed(step_strong, K, W)
            st_ms = reduce_max(sum(st_times))
            strong = {"value": GLOBAL_BATCH * K / (st_ms * 1e-3), "unit": "clouds/s", "ms_per_step": st_ms / K,
                      "global_batch": GLOBAL_BATCH, "batch_per_gpu": bs, "first_cloud": lo,
                      "note": "same graphed step and timing as the headline, 32 clouds in total: at 8 GPUs a rank holds 4 "
                              "clouds, i.e. 32 CTAs of EMD sweep per launch on 148 SMs -- launch- and occupancy-bound"}

    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_cuda = ref_cuda_timings(dev, ev_time)

    clocks = sampler.stop() if rank == 0 else {}  # sampled from the headline region through the sub-metrics
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_reference_value(sample_clouds=B_PER_GPU, reps=1)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "global_batch": world * B_PER_GPU, "points": N_POINTS,
                       "parallelism": f"batch-sharded x{world}, no data-path collective; ONE all-reduce of the (sum, count) of the losses per "
                                      "timed region, inside it (sharding.LossAccumulator)",
                       "l2": "flushed between steps (256 MiB memset outside the per-step CUDA events); inputs are 1.5 MB",
                       "timing": "sum of per-step CUDA-event durations on the launching stream, max over ranks",
                       "launch": "losses.GraphedLossStep (one CUDA-graph launch per step)" if graphed_dev is not None
                       else "eager"},
            "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "pinned host clouds -> H2D -> chamfer_emd fwd+bwd -> loss D2H each step, host wall clock; "
                            + ("the step is losses.PipelinedLossStep: one CUDA-graph launch per step, step i+1's H2D on a "
                               "copy stream while step i computes, step i's loss read on the host during step i+1"
                               if e2e_pipe is not None else
                               "the step is losses.GraphedLossStep (copies + kernels captured as one CUDA graph)"
                               if graphed is not None else "eager launches")},
            "gpu_launches": int(launches),
            # the kNN half of BASELINE's metric, at top level (sub_metrics holds the rooflines)
            "knn_xyz_k20_n1024_graphs_per_s": sub.get("knn_xyz_k20_n1024", {}).get("graphs_per_s"),
            "knn_feat64_k20_n1024_graphs_per_s": sub.get("knn_feat64_k20_n1024", {}).get("graphs_per_s"),
            "chamfer_fwd_bwd_clouds_per_s": sub.get("chamfer_fwd_bwd", {}).get("clouds_per_s"),
            "emd_fwd_bwd_clouds_per_s": sub.get("emd_fwd_bwd", {}).get("clouds_per_s"),
            "strong_scaling": strong,
            "ref_cuda": ref_cuda,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "peaks": {**peaks, "pipe": pipe},
            "sub_metrics": sub,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_value(sample_clouds: int, reps: int, with_knn: bool = True) -> dict:
    """The reference's CPU path on the host cores, bounded sample of the same workload.

    Chamfer: the reference's torch CPU path (torch_chamfer, metrics_and_losses.py:44-47) forward+backward, restated
    in oracle/torch_ref.py.  EMD: the reference has NO CPU implementation of match_cost (it is CUDA-only,
    metrics_and_losses.py:76-79), so its algorithm is timed through the C restatement oracle/geom_oracle.c
    (approxmatch + matchcost + matchcostgrad, OpenMP over all cores)."""
    import torch

    import oracle
    from oracle import torch_ref
    from pointcloudcounterfactual_b200 import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle.set_num_threads(cores)
    recon, ref = synthetic.s1_near(sample_clouds, N_POINTS)
    torch_ref.chamfer_fwd_bwd(recon, ref)  # warm-up
    oracle.approxmatch(recon[:1].numpy(), ref[:1].numpy(), want_match=False)
    t0 = time.perf_counter()
    for _ in range(reps):
        torch_ref.chamfer_fwd_bwd(recon, ref)
        match, _ = oracle.approxmatch(recon.numpy(), ref.numpy())
        oracle.matchcost(recon.numpy(), ref.numpy(), match)
        oracle.matchcostgrad(recon.numpy(), ref.numpy(), match)
    dt = (time.perf_counter() - t0) / reps
    # the kNN half of the metric: the reference's torch CPU path (torch_knn, neighbour_ops.py:71-74), full batch
    knn = {}
    for name, x in (() if not with_knn else (("knn_xyz_k20_n1024_graphs_per_s", synthetic.knn_xyz(B_PER_GPU, KNN_N)),
                    ("knn_feat64_k20_n1024_graphs_per_s", synthetic.knn_features(B_PER_GPU, KNN_C, KNN_N)))):
        torch_ref.torch_knn(x, KNN_K)
        t1 = time.perf_counter()
        for _ in range(3):
            torch_ref.torch_knn(x, KNN_K)
        knn[name] = B_PER_GPU / ((time.perf_counter() - t1) / 3)
    return {"value": sample_clouds / dt, "unit": "clouds/s", "cores": cores, "kind": "port", **knn,
            "sample": f"{sample_clouds} clouds x {N_POINTS} points, {reps} repetitions: torch CPU torch_chamfer fwd+bwd "
                      f"(reference path) + C oracle approxmatch/matchcost/matchcostgrad (no CPU EMD exists in the reference)"}


def ref_cuda_timings(dev, ev_time) -> dict:
    """BASELINE configs[2] names the reference's CUDA op as the bar: the reference's own kernels, compiled UNMODIFIED from
    /root/reference into oracle/_ref (oracle/build_ref.py; test infrastructure), timed with the same CUDA-event harness on
    the same GPU in the same run, beside this library's operators.  Baseline leg only: never on the product path."""
    import torch

    from oracle import build_ref
    from pointcloudcounterfactual_b200 import synthetic
    from pointcloudcounterfactual_b200.structural_losses.structural_losses_backend import (MatchCostFused, NNDistance,
                                                                                          NNDistanceGrad)

    if not build_ref.available("structural_losses_backend_ref"):
        return {"unavailable": "oracle/_ref/structural is not built (needs /root/reference at build time)"}
    ref = build_ref.load_ref("structural_losses_backend_ref")
    a, c = (t.to(dev) for t in synthetic.s1_near(B_PER_GPU, N_POINTS))
    g1 = torch.full((B_PER_GPU, N_POINTS), 1.0 / N_POINTS, device=dev)
    _, ri1, _, ri2 = ref.NNDistance(a, c)

    def ref_emd():
        match, _ = ref.ApproxMatch(a, c)
        ref.MatchCost(a, c, match)
        ref.MatchCostGrad(a, c, match)

    out = {"unit": "ms", "shape": f"B={B_PER_GPU} x {N_POINTS}", "timing": "CUDA events, eager launches",
           "nn_distance_fwd": {"reference_cuda": ev_time(lambda: ref.NNDistance(a, c), 10),
                               "b200": ev_time(lambda: NNDistance(a, c), 10)},
           "nn_distance_bwd": {"reference_cuda": ev_time(lambda: ref.NNDistanceGrad(a, c, ri1, ri2, g1, g1), 10),
                               "b200": ev_time(lambda: NNDistanceGrad(a, c, ri1, ri2, g1, g1), 10)},
           "approxmatch_matchcost_matchcostgrad": {"reference_cuda": ev_time(ref_emd, 3, warm=1),
                                                   "b200": ev_time(lambda: MatchCostFused(a, c, True, True), 5, warm=1)}}
    if build_ref.available("emd_backend_ref"):
        from pointcloudcounterfactual_b200.emd import emdModule

        refe = build_ref.load_ref("emd_backend_ref")
        ua, uc = (t.to(dev) for t in synthetic.auction_clouds(B_PER_GPU, N_POINTS))
        bn = (B_PER_GPU, N_POINTS)

        def buf(shape, dtype, fill=0):
            return torch.full(shape, fill, dtype=dtype, device=dev)

        def ref_auction():  # work buffers initialised per call exactly as emd_module.py:34-45 does
            refe.forward(ua, uc, buf(bn, torch.float32), buf(bn, torch.int32, -1), buf(bn, torch.float32),
                         buf(bn, torch.int32, -1), buf(bn, torch.int32), buf(bn, torch.float32), buf(bn, torch.float32),
                         buf((bn[0] * bn[1],), torch.int32), buf((512,), torch.int32), buf((512,), torch.int32),
                         buf((512,), torch.int32), buf((bn[0] * bn[1],), torch.int32), 0.005, 50)

        mod = emdModule()
        out["auction_emd_fwd_eps0.005_iters50"] = {"reference_cuda": ev_time(ref_auction, 3, warm=1),
                                                   "b200": ev_time(lambda: mod(ua, uc, 0.005, 50), 5, warm=1)}
    for v in out.values():
        if isinstance(v, dict) and "reference_cuda" in v:
            v["speedup"] = v["reference_cuda"] / v["b200"]
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    # every step is the FULL workload (32 clouds): approxmatch is OpenMP-parallel over the points of a cloud, matchcost /
    # matchcostgrad over (cloud, point), the torch Chamfer path over the batch -- all host cores work, same config as the
    # GPU arm
    sample = B_PER_GPU
    base = None
    t_all0 = time.perf_counter()
    vals = []
    budget_s = 200.0  # the whole run must end within a few minutes: if one full batch is too slow on this host, fewer clouds
    i = 0
    while i < W + K:
        t_step = time.perf_counter()
        r = cpu_reference_value(sample_clouds=sample, reps=1, with_knn=(i == W + K - 1))
        t_step = time.perf_counter() - t_step
        if i == 0 and sample > 4 and t_step * (W + K) > budget_s:  # re-size once, after the first (untimed) step
            while sample > 4 and t_step * (W + K) * sample / B_PER_GPU > budget_s:
                sample //= 2
            if W == 0:
                continue  # the first step was sized wrong and would be a timed one: redo it
        if i >= W:
            vals.append(r["value"])
        base = r
        i += 1
    total = sum(sample / v for v in vals)
    value = sample * K / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "points": N_POINTS,
                   "sample": f"each step = {sample} clouds" + (" = the full batch" if sample == B_PER_GPU else
                                                                 f" of the {B_PER_GPU} (host too slow for the full batch in the time budget)")},
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": base["cores"], "kind": "port", "sample": base["sample"]},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all0,
    }
    emit(out)


def _claim_stdout() -> None:
    """Libraries print to stdout (NCCL's version banner under torchrun): keep fd 1 for the ONE JSON line by pointing it
    at stderr for everything else; print() below writes to the saved descriptor."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_JSON_OUT = None


def emit(obj: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


if __name__ == "__main__":
    _claim_stdout()
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
