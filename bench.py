#!/usr/bin/env python
"""bench.py -- b_sae 512->32768 4-bit forward tokens/s on N B200s (+ roofline, CPU baseline).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (oracle port, rank 0 only)

A "step" is one forward of the hot path over one batch of synthetic activations:
x [B,512] fp32 (resident in HBM) -> bf16 cast -> fused tcgen05 encoder + top-k -> merge -> packed
int4 sparse decode -> (values, indices) [B,k] + reconstruction [B,512] (one call: qsae_bsae_forward). Weights are pre-packed
(one-time cost, not in the step). Rows are independent, so N GPUs shard the batch with
replicated weights and no collective on the data path ("scaling": "weak").
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D, H, N_BITS, GAMMA = 512, 32768, 4, 4.0
E2E_CHUNK = int(os.environ.get("QSAE_E2E_CHUNK", "8192"))   # largest row chunk of the host-buffer pipeline
METRIC = "b_sae 512->32768 4-bit fwd tokens/s"
UNIT = "tokens/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="rows per GPU per step")
    ap.add_argument("--k", type=int, default=32, help="latents kept per row (model.k = k / 32768)")
    ap.add_argument("--cpu-batch", type=int, default=4096, help="rows per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step from CUDA graphs (one per rotating input): removes the launch gaps that matter at "
                         "small batches (--batch 4096); the default run launches the kernels directly")
    ap.add_argument("--heavy-tail", action="store_true",
                    help="SURVEY 8d heavy-tail variant of the inputs: 8 of the 512 dimensions scaled by 20")
    ap.add_argument("--variant", default="batch-sharded", choices=["batch-sharded", "dict-sharded"],
                    help="dict-sharded: BASELINE config 5, b_sae 512->2^20 with the dictionary split over the GPUs "
                         "(NCCL all-gather of top-k candidates + reduce-scatter of partial reconstructions)")
    ap.add_argument("--hidden", type=int, default=2 ** 20, help="dictionary size of the dict-sharded variant")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p"],
                    help="dict-sharded variant: candidate / partial exchange through NCCL collectives or CUDA IPC peer memory")
    return ap.parse_args()


def workload_name(batch, k):
    return (f"b_sae input_dim=512 hidden_dim=32768 n_bits=4 gamma=4.0 k={k} forward, "
            f"synthetic Pythia-70m-shaped activations{' (heavy-tail variant: 8 dims x 20)' if HEAVY_TAIL else ''}, "
            f"batch {batch} per GPU")


# ---------------------------------------------------------------------------------------------
# synthetic weights / inputs: xavier encoder rounded to bf16-representable
# fp32, polarised (+-110) decoder logits, N(0,1) decoder bias, N(0,1) x rounded to bf16-representable
# ---------------------------------------------------------------------------------------------
def make_weights(torch, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    bound = (6.0 / (H + D)) ** 0.5
    We = ((torch.rand((H, D), device=device, generator=g) * 2 - 1) * bound).bfloat16().float()
    be = torch.zeros(H, device=device)
    logits = torch.where(torch.rand((H, D * N_BITS), device=device, generator=g) < 0.5, 110.0, -110.0)
    bd = torch.randn(D, device=device, generator=g)
    return We, be, logits.float().contiguous(), bd


HEAVY_TAIL = False   # set from --heavy-tail


def make_x(torch, device, batch, seed):
    g = torch.Generator(device=device).manual_seed(1000 + seed)
    x = torch.randn((batch, D), device=device, generator=g)
    if HEAVY_TAIL:       # a few outlier channels, like the residual stream of a transformer
        x[:, torch.arange(8, device=device) * 61 + 5] *= 20.0
    return x.bfloat16().float()


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread (a timed
    region is only tens of milliseconds long), nvidia-smi -lms as the fallback when pynvml is unavailable."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in bits.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._nvml = N
            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 2 ms polling"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            p = [t.strip() for t in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------
# CPU arm: op-for-op port of the reference forward (oracle/), torch CPU, all host threads
# ---------------------------------------------------------------------------------------------
def cpu_forward_rate(batch, k, steps, warmup):
    import torch

    from oracle import qsae_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu = torch.device("cpu")
    We, be, logits, bd = make_weights(torch, cpu)
    x = make_x(torch, cpu, batch, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.bsae_forward_dense_port_torch(x, We, be, logits, bd, n_bits=N_BITS, gamma=GAMMA, k=k)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    rate, sec, cores = cpu_forward_rate(args.cpu_batch, args.k, args.steps, max(1, args.warmup))
    sample = (f"{args.cpu_batch} rows per step of the same workload (the reference's dense fp32 forward needs "
              f"3 x [B,32768] fp32 temporaries); torch CPU, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, args.k), "cpu_rows_per_step": args.cpu_batch},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"tflops": float(j.get("bf16_tflops_sustained", j.get("bf16_tflops"))), "hbm_gbs": float(j["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def run_b200(args, rank, world, local_rank):
    import torch

    from quantizedsae_b200 import _lib as L

    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    lib = L.load()
    L.check(lib.qsae_check_device())
    B, k = args.batch, args.k

    We, be, logits, bd = make_weights(torch, device)
    w_bf16 = L.cast_bf16(We)
    sample = L.prepare_sample(w_bf16, be)
    packed, _, gap = L.pack_bitplanes(logits, D, N_BITS)
    assert gap == 0.0
    del logits
    qstep = GAMMA / 2 ** (N_BITS - 1)
    xs = [make_x(torch, device, B, s + 10 * rank) for s in range(3)]   # 3 x 134 MB rotating inputs (> L2)

    def step(i):
        x = xs[i % len(xs)]
        vals, idx, _, recon = L.bsae_forward(x, w_bf16, None, be, k, packed, N_BITS, qstep, bd, sample=sample)
        return vals, idx, recon

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled meanwhile
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        k_ev[i][0].record(); k_ev[i][1].record()           # materialise the handles before handing them over
        L.check(lib.qsae_set_encode_kernel_events(k_ev[i][0].cuda_event, k_ev[i][1].cuda_event))
        step(i)
    ev1.record()
    L.check(lib.qsae_set_encode_kernel_events(None, None))
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = L.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in k_ev)
    if dist is not None:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- the same step in exact mode (fp32 re-scoring of k + 16 tensor-core candidates from the fp32 weights: what the
    #      drop-in modules do by default for arbitrary fp32 operands); reported next to the headline, not instead of it
    exact_steps = max(3, min(args.steps, 10))
    for i in range(2):
        L.bsae_forward(xs[i % len(xs)], w_bf16, We, be, k, packed, N_BITS, qstep, bd, exact=True, sample=sample)
    barrier()
    evx0, evx1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evx0.record()
    for i in range(exact_steps):
        L.bsae_forward(xs[i % len(xs)], w_bf16, We, be, k, packed, N_BITS, qstep, bd, exact=True, sample=sample)
    evx1.record()
    barrier()
    exact_ms = evx0.elapsed_time(evx1) / exact_steps
    if dist is not None:
        t = torch.tensor([exact_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exact_ms = float(t[0])

    # ---- optional: the same step replayed from CUDA graphs (after the headline measurement, which stays undisturbed)
    graphs = None
    graph_ms = None
    if args.graph:
        # capture one graph per rotating input on a side stream (ctypes launches go to torch's current stream; the
        # workspace is already allocated by the warm-up, outputs come from the graph's private pool)
        graphs = []
        for i in range(len(xs)):
            g_i = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_i):
                out_i = step(i)
            graphs.append((g_i, out_i))
        for g_i, _ in graphs:
            g_i.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            graphs[i % len(graphs)][0].replay()
        e1.record()
        barrier()
        graph_ms = e0.elapsed_time(e1) / args.steps
        if dist is not None:
            t = torch.tensor([graph_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            graph_ms = float(t[0])

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (H2D + kernels + D2H inside)
    e2e = None
    plan = C.c_void_p()
    g = torch.Generator(device=device).manual_seed(5)
    logits2 = torch.where(torch.rand((H, D * N_BITS), device=device, generator=g) < 0.5, 110.0, -110.0).float()
    L.check(lib.qsae_bsae_plan_create(We.data_ptr(), be.data_ptr(), logits2.data_ptr(), bd.data_ptr(), H, D, N_BITS,
                                      C.c_float(GAMMA), k, E2E_CHUNK, C.byref(plan)))
    del logits2
    try:
        hx = [x.cpu().pin_memory() for x in xs[:2]]
        hv = torch.empty((B, k), dtype=torch.float32).pin_memory()
        hi = torch.empty((B, k), dtype=torch.int32).pin_memory()
        hr = torch.empty((B, D), dtype=torch.float32).pin_memory()
        e2e_steps = max(3, min(args.steps, 10))
        for i in range(2):
            L.check(lib.qsae_bsae_forward_host(plan, hx[i % 2].data_ptr(), B, hv.data_ptr(), hi.data_ptr(), hr.data_ptr()))
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            L.check(lib.qsae_bsae_forward_host(plan, hx[i % 2].data_ptr(), B, hv.data_ptr(), hi.data_ptr(), hr.data_ptr()))
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        e2e = {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * D * 4,
               "d2h_bytes_per_step": B * k * 8 + B * D * 4, "steps": e2e_steps,
               "api": "qsae_bsae_forward_host (pinned host x in; values, indices, reconstruction to pinned host)"}
    finally:
        lib.qsae_bsae_plan_destroy(plan)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flops = 2.0 * B * H * D
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("encode_topk_kernel_dram_bytes_per_launch")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(B, k), "latents_out": "sparse (values, indices) [B,k]; dense [B,H] not written",
                   "l2": "inputs 134 MB/step exceed the 126 MB L2; 3 rotating input buffers",
                   "weights": "pre-packed once (bf16 encoder, int4 dictionary); not in the step",
                   "precision": "synthetic x and encoder weights are bf16-representable (SURVEY 8d config 1), so the bf16 "
                                "tensor-core products are exact and the result equals the fp32 reference up to fp32 "
                                "accumulation order; int4 decode accumulates exact integers. exact_mode = same step with "
                                "fp32 re-scoring from the fp32 weights (valid for arbitrary fp32 operands)",
                   "parallelism": f"batch-sharded x{world}, replicated weights, no collective"},
        "roofline": {"bound": "tensor", "kernel": "encode_topk_kernel<8>", "achieved": achieved, "peak": peaks["tflops"],
                     "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                     "algorithmic_flops_per_launch": flops, "kernel_ms": kernel_ms, "peak_source": peaks["source"]},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "exact_mode": {"value": world * B / (exact_ms * 1e-3), "unit": UNIT, "ms_per_step": exact_ms, "steps": exact_steps},
    }
    if graph_ms is not None:
        out["cuda_graph"] = {"value": world * B / (graph_ms * 1e-3), "unit": UNIT, "ms_per_step": graph_ms,
                             "note": "the same step replayed from CUDA graphs (one per rotating input)"}
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores = cpu_forward_rate(args.cpu_batch, k, 2, 1)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{args.cpu_batch} rows x 2 timed forwards (1 warm-up) of the same workload, "
                                         f"oracle/qsae_oracle.bsae_forward_dense_port_torch"}
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_dict_sharded(args, rank, world, local_rank):
    """BASELINE config 5: one 2^20-latent dictionary split over the ranks; x replicated; strong scaling."""
    import torch

    from quantizedsae_b200 import _lib as L
    from quantizedsae_b200.sharded import DictionaryShardedBinarySAE

    dist = None
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=device)
    L.check(L.load().qsae_check_device())
    Hh, B, k = args.hidden, min(args.batch, 4096) if args.batch == 65536 else args.batch, args.k
    with torch.device(device):
        m = DictionaryShardedBinarySAE(D, Hh, GAMMA, N_BITS, rank=rank, world_size=world)
    g = torch.Generator(device=device).manual_seed(100 + rank)
    with torch.no_grad():
        m.encoder[0].weight.copy_(m.encoder[0].weight.bfloat16().float())
        hs = m.plan.shard_latents
        for a in range(0, hs, 65536):          # polarised logits, generated in slices
            b = min(hs, a + 65536)
            m.decoder.weight[a:b] = torch.where(torch.rand((b - a, D * N_BITS), device=device, generator=g) < 0.5, 110.0, -110.0)
        m.decoder.bias.copy_(torch.randn(D, device=device, generator=torch.Generator(device=device).manual_seed(7)))
    m.eval()
    m.exact = False
    m.transport = args.transport
    m.k = k / Hh
    gx = torch.Generator(device=device).manual_seed(1000)     # the same x on every rank (replicated input)
    xs = [torch.randn((B, D), device=device, generator=gx).bfloat16().float() for _ in range(3)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(3, args.warmup)):
            m(xs[i % 3])
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = L.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            lat, rows, _ = m(xs[i % 3])
        ev1.record()
        barrier()
    if m._peer is not None:
        m._peer.check()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = L.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
    if rank == 0:
        peaks = load_peaks()
        flops_per_gpu = 2.0 * B * (Hh // world) * D
        ms = elapsed_ms / args.steps
        print(json.dumps({
            "metric": f"b_sae 512->{Hh} 4-bit dictionary-sharded fwd tokens/s", "value": B * args.steps / (elapsed_ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"b_sae input_dim=512 hidden_dim={Hh} n_bits=4 gamma=4.0 k={k} forward, batch {B} "
                                   f"replicated, dictionary split over {world} GPU(s)",
                       "l2": f"encoder shard {(Hh // world) * D * 2 / 1e6:.0f} MB bf16 per GPU streams from HBM every step (> 126 MB L2 when > 1); 3 rotating inputs",
                       "parallelism": (f"dictionary-sharded x{world}: NCCL all-gather of [B,k] candidates, reduce-scatter of [B,512] partials"
                                       if args.transport == "nccl" else
                                       f"dictionary-sharded x{world}: candidates and [B,512] partials exchanged through CUDA IPC peer "
                                       f"memory (flag-synchronised P2P loads inside the merge / reduce kernels)")},
            "roofline": {"bound": "tensor", "kernel": "encode_topk_kernel<8> over the local shard", "achieved": flops_per_gpu / (ms * 1e-3) / 1e12,
                         "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": flops_per_gpu / (ms * 1e-3) / 1e12 / peaks["tflops"],
                         "traffic": None, "note": "whole step time used (upper bound on kernel time): fraction is a lower bound",
                         "peak_source": peaks["source"]},
            "gpu_launches": launches, "clocks": clocks}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    global HEAVY_TAIL
    args = parse()
    HEAVY_TAIL = bool(args.heavy_tail)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: spawn it ourselves so `python bench.py --gpus N` also works
        port = os.environ.get("MASTER_PORT", "29531")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", port, __file__, "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup), "--batch", str(args.batch), "--k", str(args.k),
               "--variant", args.variant, "--hidden", str(args.hidden), "--transport", args.transport] + (["--heavy-tail"] if args.heavy_tail else []) + (["--graph"] if args.graph else [])
        raise SystemExit(subprocess.call(cmd))
    if args.variant == "dict-sharded":
        run_dict_sharded(args, rank, world, local_rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
