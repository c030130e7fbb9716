#!/usr/bin/env python
"""bench.py -- G+D training-step throughput (Step-A of trainer.py) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference algorithm on the host CPU cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...  # data parallel, one rank per GPU

Prints ONE JSON line (rank 0).  A "step" = one Step-A pass (G forward, 3 PatD updates, G update + EMA) over a
synthetic COCO-shaped batch of 16 images per GPU at 256x256 (BASELINE.json configs[1]; weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "G+D training-step images/sec @256x256 (Step-A: G fwd+bwd, 3 PatD updates, G update through 3 PatD + KL)"
UNIT = "images/s"
GMAC_PER_IMG_STEP_A = 330.0  # SURVEY.md 8d / BASELINE.md section 2
GMAC_PER_IMG_STEP_B = 765.0  # same table; includes DAMSM + per-step Inception (48 GMAC) which this step does not run
DTYPE = "f32 (contractions: 3xFP16 error-compensated tcgen05 MMA, 22-bit mantissa operands, fp32 accumulate)"


def ncu_traffic(kernel_key):
    """DRAM read+write bytes per launch of a kernel, from the committed ncu --set full capture summary
    (profiles/ncu_traffic.json, written by profiles/summarise_ncu.py from the .ncu-rep of this bench command).
    Returns (bytes or None, provenance string)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no ncu capture summary committed"
    d = json.load(open(p))
    e = d.get(kernel_key)
    if not e:
        return None, f"{kernel_key} not in profiles/ncu_traffic.json"
    return e["dram_bytes"], f"dram__bytes_read.sum + dram__bytes_write.sum per launch, {e['source']}"



def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active")
                                                          for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_setup(n):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world):
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    barrier(world)
    return max_over_ranks(ms, world)


def conv_roofline(hbm, bf16_tf, src):
    """Dominant kernel: the stage-3 HmapResBlock first conv (194->388, 3x3 reflect, 128x128, B=16): 11.1 GMAC/img
    (SURVEY.md 8d).  Timed alone with CUDA events on the launching stream; operands (210 MB in, 420 MB out) exceed
    the 126 MB L2, so every iteration is HBM-cold."""
    from objgan_b200 import model
    from objgan_b200.lib import PAD_REFLECT
    B, C, H = 16, 194, 128
    m = model.Conv2dP(C, 2 * C, 3, 1, 1, mode=PAD_REFLECT, split=C).cuda()
    x = torch.randn(B, H, H, 200, device="cuda")
    x[..., 194:] = 0
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * B * H * H * (9 * C) * (2 * C)          # algorithmic, unpadded
    achieved = flops / (ms * 1e-3) / 1e12
    peak = bf16_tf                                         # kind::f16 runs at the bf16 rate
    from objgan_b200 import ops as _ops
    eng = _ops.CONV_ENGINE
    traffic, tsrc = ncu_traffic("conv_tc2_res3_conv1") if eng == "f16x3" else (None, "")
    kname = {"simt": "conv_gemm_kernel<8> (fp32 FMA)", "f16": "conv_tc2_kernel<208> (tcgen05 kind::f16, 1 product)",
             "f16x3": "conv_tc2_kernel<208> (tcgen05 kind::f16, 3xFP16 error-compensated hi/lo operands)"}[eng]
    note = {"simt": "CUDA-core fp32 path", "f16": "single fp16 product (not the parity mode)",
            "f16x3": "3 MMAs per algorithmic product, so frac <= 1/3 by construction; amax + prep_split passes included"}[eng]
    return {"bound": "tensor", "kernel": kname + " + prep_split: res-block conv1 194->388 3x3 @128x128, B=16",
            "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
            "traffic": traffic, "ms_per_launch": round(ms, 3),
            "algorithmic_flops_per_launch": flops,
            "issued_mma_tflops": round(achieved * (3 if eng == "f16x3" else 1), 1) if eng != "simt" else None,
            "frac_of_peak_counting_issued_mmas": round(achieved * (3 if eng == "f16x3" else 1) / peak, 4) if eng != "simt" else None,
            "peak_source": f"{src} bf16 cuBLAS burst (kind::f16 = bf16 rate); {note}; traffic = {tsrc} "
                           "(algorithmic: 635 MB)"}


def attn_roofline(hbm, src):
    """GlobalAttentionGeneral forward, Q=16384 regions x L=18 words, C=48, B=16: algorithmic bytes 4*Q*(2C+L) per image
    = 7.47 MB."""
    from objgan_b200 import ops
    B, Q, C, L = 16, 16384, 48, 18
    # 8 independent input sets (8 x 119 MB >> 126 MB L2): launched back to back between two events, so every launch
    # reads HBM-cold data and the host's launch latency is hidden behind the previous kernel
    sets = [(torch.randn(B, 128, 128, C, device="cuda"), torch.randn(B, C, L, device="cuda")) for _ in range(8)]
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    # the 8 launches are captured once in a CUDA graph (a launch lasts ~33 us: issued one by one from Python the host,
    # not the kernel, would be timed) and the graph is replayed between two events on the capture stream
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for h, srcw in sets:
            ops.att_general(h, srcw, None, C)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for h, srcw in sets:
                ops.att_general(h, srcw, None, C)
        times = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1) / len(sets))
    torch.cuda.current_stream().wait_stream(side)
    ms = sum(times) / len(times)
    byts = 4.0 * Q * (2 * C + L) * B
    ach = byts / (ms * 1e-3) / 1e9
    traffic, tsrc = ncu_traffic("att_general_fwd")
    return {"bound": "hbm", "kernel": "att_general_fwd_tc_kernel<20> (tcgen05, fp16 hi/lo operands; Q=16384, L=18, C=48, B=16)", "achieved": round(ach, 1),
            "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4), "traffic": traffic, "ms_per_launch": round(ms, 4),
            "algorithmic_bytes_per_launch": byts,
            "peak_source": f"{src} copy bandwidth; traffic = {tsrc} (written lines may still be dirty in the 126 MB L2 "
                           "when the kernel ends)"}


def _best_cpu_threads():
    """The reference's CPU path is torch CPU ops; on a many-core host more threads is not always faster, so give it
    the best of a few thread counts (measured on one G forward at batch 2 each) -- the fairest CPU arm we can build."""
    from objgan_b200 import model, synth
    from oracle import objgan_oracle as O
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    g = model.G_NET(80)
    sd = g.state_dict()
    inp = synth.make_inputs(2, seed=1, parity=False)
    best, best_t = cores, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            O.g_net_forward({k: v.clone() for k, v in sd.items()}, inp)      # warm
            t0 = time.perf_counter()
            O.g_net_forward({k: v.clone() for k, v in sd.items()}, inp)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


CPU_BATCH = 16     # the CPU arm runs the SAME configuration as the GPU arm (BASELINE configs[1]: batch 16)


def _reference_modules():
    """The reference's own image_generation modules, importable only where /root/reference is mounted (never on the
    GPU box).  Returns the namespace of tests/refimport.py or None."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refimport
        return refimport.load() if refimport.available() else None
    except Exception:
        return None


def cpu_step_a(batch, steps, warmup, seed=1234):
    """The reference's algorithm for Step-A on the host cores at batch `batch`.  kind "reference": the reference's own
    model.py / losses.py modules + torch.optim.Adam (only where /root/reference is importable); kind "port": the
    oracle restatement (oracle/objgan_oracle.py), which is what can travel to the GPU box.
    Returns (img/s, s/step, threads, kind)."""
    from objgan_b200 import model, synth
    from oracle import objgan_oracle as O
    threads = _best_cpu_threads()
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    inp = synth.make_inputs(batch, seed=seed, parity=False)
    ref = _reference_modules()
    if ref is not None:
        g = ref.model.G_NET(80)
        ds = [ref.model.PAT_D_NET64(), ref.model.PAT_D_NET128(), ref.model.PAT_D_NET256()]
        for m in [g, *ds]:
            m.apply(ref.utils.weights_init)
        opt_g = torch.optim.Adam(g.parameters(), lr=2e-4, betas=(0.5, 0.999))
        opt_d = [torch.optim.Adam(d.parameters(), lr=2e-4, betas=(0.5, 0.999)) for d in ds]
        avg = [p.data.clone() for p in g.parameters()]
        bce = torch.nn.BCELoss()

        def one():  # ref: trainer.py:388-406, 444-462 restricted to G + PatD + KL (Step-A)
            out = g(inp["z"], inp["sent_emb"], inp["words_embs"], inp["glove_words_embs"], inp["slabels_feat"],
                    inp["mask"], inp["hmaps"], inp["rois"], inp["fm_rois"], inp["num_rois"], inp["bt_masks"],
                    inp["fm_bt_masks"], inp["glb_max_num_roi"])
            fake, mu, logvar = out[0], out[4], out[5]
            for i, d in enumerate(ds):
                d.zero_grad()
                ref.losses.patD_loss(d, inp["imgs"][i], fake[i], inp["sent_emb"]).backward()
                opt_d[i].step()
            g.zero_grad()
            total = ref.losses.KL_loss(mu, logvar)
            for i, d in enumerate(ds):
                f = d(fake[i])
                c, u = d.COND_DNET(f, inp["sent_emb"]), d.UNCOND_DNET(f)
                total = total + bce(u, torch.ones_like(u)) * 1.0 + bce(c, torch.ones_like(c)) * 0.1
            total.backward()
            opt_g.step()
            for p, a in zip(g.parameters(), avg):
                a.mul_(0.999).add_(p.data, alpha=0.001)
        kind = "reference"
    else:
        g = model.G_NET(80)
        ds = [model.PAT_D_NET64(), model.PAT_D_NET128(), model.PAT_D_NET256()]
        state = O.StepAState(g.state_dict(), [d.state_dict() for d in ds])
        one = lambda: O.step_a(state, inp)
        kind = "port"
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = CPU_BATCH
    ips, spstep, cores, kind = cpu_step_a(B, args.steps, args.warmup)
    what = ("the reference's own model.py / losses.py modules" if kind == "reference"
            else "oracle/objgan_oracle.py, the CPU restatement of the reference (the Python reference itself cannot "
                 "travel to the GPU box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(ips, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(spstep * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "step_a_b16_256 (BASELINE configs[1]: full G_NET 64/128/256 + 3 patch Ds, batch 16)",
                   "global_batch": B, "words": 18, "rois": 10, "same_config_as_gpu_arm": True},
        "cpu_baseline": {"value": round(ips, 4), "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} Step-A steps at batch {B} ({what}; torch CPU fp32, "
                                   f"{cores} threads = best of 8/16/32/64/all {os.cpu_count()} host cores)"},
        "e2e": {"value": round(ips, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def step_b_section(args, local, hbm):
    """The reference's COMPLETE step (trainer.py:385-462: + 3 shape Ds, 2 object Ds at 512^2 and their G-loss terms)
    on the same batch-16 inputs; eager launches (the kept-roi count and permute_seg's shuffles change the tensor
    shapes every step, so it is not replayed from a graph), no host synchronisation inside the step."""
    from objgan_b200 import synth, trainer
    import random
    random.seed(1234)
    B = args.batch_per_gpu
    tr = trainer.StepBTrainer(device=f"cuda:{local}", seed=1234)
    full = synth.make_inputs(B, seed=1234, parity=False)
    full.pop("eps")
    host = trainer.pin(synth.compact(full))
    dev = tr.to_device(host)
    for _ in range(2):
        tr.step(dev)
    torch.cuda.synchronize()
    n = max(3, args.steps // 2)
    t0 = time.perf_counter()
    ms = timed(lambda: tr.step(dev), n, 1)
    wall = time.perf_counter() - t0
    ips = B * n / (ms * 1e-3)
    return {"workload": "step_b_b16_256: G fwd, 3 PatD + 3 ShpD + ObjSS + ObjLS updates, G update through all eight "
                        "+ KL, Adam x9, EMA (no DAMSM image encoder, no per-step Inception score)",
            "value": round(ips, 3), "unit": UNIT, "ms_per_step": round(ms / n, 2), "steps": n, "cuda_graph": False,
            "host_wall_ms_per_step": round(wall * 1e3 / n, 2),
            "algorithmic_gmac_per_image": GMAC_PER_IMG_STEP_B - 48.0,
            "algorithmic_tflops": round(2 * (GMAC_PER_IMG_STEP_B - 48.0) * 1e9 * ips / 1e12, 2)}


def run_b200(args):
    from objgan_b200 import lib, synth, trainer
    world, rank, local = dist_setup(args.gpus)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun"
    hbm, bf16_tf, bf16_sus, src = peaks()
    B = args.batch_per_gpu
    tr = trainer.StepATrainer(device=f"cuda:{local}", seed=1234)
    tr.broadcast_parameters()
    full = synth.make_inputs(B, seed=1234 + rank, parity=False)
    full.pop("eps")  # CA_NET draws its own noise on the device, like the reference
    # what crosses PCIe every step: the batch WITHOUT the class heat maps / label embeddings (rebuilt on the device
    # from the per-roi masks by trainer.prepare_data; the reference copies them, 86 % of its input bytes)
    host = trainer.pin(synth.compact(full))
    h2d_reference_format = synth.input_bytes(full)
    del full
    dev = tr.to_device(host)
    torch.cuda.synchronize()
    h2d = synth.input_bytes(host)

    graphed = False
    if not args.no_graph:
        try:
            tr.capture(dev)          # whole step (fwd, bwd, all-reduce, Adam+EMA) as one CUDA graph
            graphed = True
        except Exception as e:       # report and fall back to eager launches
            print(f"[bench] CUDA graph capture failed, running eager: {e!r}", file=sys.stderr)
            tr._graph = None
    for _ in range(args.warmup):
        tr.step(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.get().launches
    ms = timed(lambda: tr.step(dev), args.steps, world)
    launches = lib.get().launches - n0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # end-to-end through the public call: pinned host batch -> device every step, loss read back every step
    tr.step_from_host(host)
    ms_e2e = timed(lambda: tr.step_from_host(host), args.steps, world)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        return
    roof = conv_roofline(hbm, bf16_tf, src)
    roof_att = attn_roofline(hbm, src)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, sps, thr, kind = cpu_step_a(CPU_BATCH, 1, 0)
        cpu = {"value": round(ips, 4), "unit": UNIT, "cores": thr, "kind": kind,
               "sample": f"1 Step-A step at batch {CPU_BATCH} (the GPU arm's configuration; {sps:.1f} s) of "
                         f"{'the reference modules' if kind == 'reference' else 'oracle/objgan_oracle.py'}, torch CPU "
                         f"fp32, {thr} threads (best of 8/16/32/64/all {os.cpu_count()} host cores)"}
    ops_lines = step_b = None
    if world == 1 and not args.no_extras:
        import bench_ops
        ops_lines = bench_ops.collect()          # BASELINE configs[2] / configs[3] operator lines
        del tr
        torch.cuda.empty_cache()
        step_b = step_b_section(args, local, hbm)
    step_tflops = 2 * GMAC_PER_IMG_STEP_A * 1e9 * value / 1e12
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": "step_a_b16_256 (BASELINE configs[1]: full G_NET 64/128/256 + 3 patch Ds, batch 16/GPU)",
                   "global_batch": world * B, "words": 18, "rois": 10, "parallelism": f"dp{world}",
                   "l2": "inputs (0.47 GB/step) and activations (>10 GB) exceed the 126 MB L2; no explicit flush",
                   "conv_engine": __import__("objgan_b200.ops", fromlist=["x"]).CONV_ENGINE, "cuda_graph": graphed},
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step_reference_format": h2d_reference_format,
                "note": "inputs cross PCIe without the 80-channel class heat maps and label embeddings; both are "
                        "rebuilt on the device (og_form_hmaps / og_form_clabels_feat) inside the timed region"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_attention": roof_att,
        "algorithmic_tflops": round(step_tflops, 2),
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if ops_lines is not None:
        line["ops"] = ops_lines
    if step_b is not None:
        line["step_b"] = step_b
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the operator lines (configs 3/4) and the Step-B section")
    args = ap.parse_args()
    if os.environ.get("OBJGAN_BENCH_WATCHDOG"):      # debugging aid: dump every thread's Python stack after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["OBJGAN_BENCH_WATCHDOG"]), repeat=False, file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl == "b200":
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
