#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 thermodynamic scoring path.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on
that fits one GPU): a synthetic batch of 10^7 random 18-30 nt oligo pairs; one
"step" scores the whole batch: calcHeterodimer(A,B) + calcHairpin(A) +
calcHairpin(B) (3 thal evaluations per pair, the metric's unit) + calcTm(A) +
calcTm(B).  Under torchrun every rank scores its own batch of the same shape
(independent shards, no data-path collective: weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line on rank 0 (contract in the task description):
  value        whole-job thal evaluations/s with inputs resident in HBM
  e2e          same metric through the public API with pinned HOST buffers
  roofline     dominant kernel (heterodimer DP) against the measured FP64 peak
  cpu_baseline the CPU oracle on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "oligo_pair_thal_evaluations_per_sec"
UNIT = "thal evaluations/s"
EVALS_PER_PAIR = 3
FLOP_PER_RELAXATION = 24  # SURVEY.md 8d convention


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU per step")
    p.add_argument("--cpu-sample", type=int, default=150_000, help="pairs in the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-markers", type=int, default=64, help="markers in the config-5 CPU baseline sample")
    p.add_argument("--markers", type=int, default=8192,
                   help="markers per GPU of the secondary config-5 (KASP design) measurement; 0 skips it")
    return p.parse_args()


def workload_config(args, world):
    return {"workload": "config4: synthetic random 18-30 nt oligo pairs, heterodimer + 2x hairpin + 2x Tm per pair",
            "pairs_per_gpu": args.pairs, "global_pairs": args.pairs * world, "evals_per_pair": EVALS_PER_PAIR,
            "conditions": "mv 50 mM, dv 0, dNTP 0.8 mM, DNA 50 nM, 37 C, max_loop 30 (ThermoAnalysis defaults)",
            "parallelism": "independent shards, %d rank(s), no collective" % world,
            "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2" % (args.pairs * (32 + 136) / 1e6)}


# ----------------------------------------------------------------------------
# CPU side (oracle): cpu_baseline leg and the --impl reference arm
# ----------------------------------------------------------------------------
def cpu_score(orc, a_str, b_str, threads):
    """One CPU pass over a sample: same work as one GPU step."""
    from oracle.oracle import ANY, HAIRPIN
    t0 = time.perf_counter()
    orc.thal_batch(a_str, b_str, ANY, nthreads=threads)
    orc.thal_batch(a_str, None, HAIRPIN, nthreads=threads)
    orc.thal_batch(b_str, None, HAIRPIN, nthreads=threads)
    orc.tm_batch(a_str, nthreads=threads)
    orc.tm_batch(b_str, nthreads=threads)
    return time.perf_counter() - t0


def cpu_sample(args):
    from polyoligo_b200 import synth
    a, b = synth.config4(args.cpu_sample)
    return synth.to_strings(a), synth.to_strings(b)


def run_reference(args, rank, world):
    """The reference's CPU path: libprimer3 cannot be installed offline, so this is
    the oracle port of it, on all host threads, on a bounded sample per step."""
    if rank != 0:
        return
    from oracle.oracle import Oracle, build
    build()
    orc = Oracle()
    threads = os.cpu_count() or 1
    a_str, b_str = cpu_sample(args)
    for _ in range(min(args.warmup, 1)):
        cpu_score(orc, a_str[:2000], b_str[:2000], threads)
    times = [cpu_score(orc, a_str, b_str, threads) for _ in range(args.steps)]
    dt = sum(times)
    value = EVALS_PER_PAIR * args.cpu_sample * args.steps / dt
    sample = "config-4 batch of %d pairs per step (same generator and distribution, smaller n)" % args.cpu_sample
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_config5(args, rank, local_rank, barrier):
    """BASELINE config 5: the whole KASP design (enumeration, validity, common-primer picking,
    merge, reporter-tailed heterodimers, goodness, top-10) of synthetic markers through the public
    array API with HOST buffers; wall clock of one pass after a small warm-up."""
    from polyoligo_b200 import kasp_pipeline as kp, lib_primer3 as lp, primer3_bindings as p3b
    p3b.resetP3Globals()
    lp.PRIMER3_GLOBALS.clear()
    # reference src/polyoligo/data/PRIMER3_KASP.yaml + cli_kasp.py defaults (--depth 100, -n 10, dyes)
    lp.set_globals(PRIMER_OPT_SIZE=20, PRIMER_MIN_SIZE=18, PRIMER_MAX_SIZE=25, PRIMER_OPT_TM=60.0,
                   PRIMER_MIN_TM=55.0, PRIMER_MAX_TM=65.0, PRIMER_MIN_GC=20.0, PRIMER_MAX_GC=80.0,
                   PRIMER_MAX_POLY_X=100, PRIMER_SALT_MONOVALENT=50.0, PRIMER_DNA_CONC=50.0,
                   PRIMER_MAX_NS_ACCEPTED=0, PRIMER_PAIR_MAX_DIFF_TM=6.0,
                   PRIMER_PRODUCT_SIZE_RANGE=[[90, 150], [100, 200], [50, 250]], PRIMER_NUM_RETURN=100)
    lp.set_tm_delta(5.0)
    dyes = ("GAAGGTCGGAGTCAACGGATT", "GAAGGTGACCAAGTTCATGCT")
    workers = 2
    kctxs = [kp._worker_context(local_rank, slot) for slot in range(workers)]
    kp.design_markers_arrays(kp.synth_config5(args.markers, seed=1), n_primers=10, reporters=dyes, device=local_rank,
                             workers=2)  # warm-up
    data = kp.synth_config5(args.markers, seed=kp.CONFIG5_SEED + rank)
    stats = {}
    items0 = sum(c.thal_item_count() for c in kctxs)
    launches0 = sum(c.launch_count() for c in kctxs)
    barrier()
    t0 = time.perf_counter()
    res = kp.design_markers_arrays(data, n_primers=10, reporters=dyes, device=local_rank, stats=stats,
                                   workers=workers)
    dt = time.perf_counter() - t0
    barrier()
    c5 = {"workload": "synthetic KASP design, %d markers per GPU x 500 nt, SNP at 249, depth 100, top 10, "
                      "no off-targets (BASELINE config 5 shape)" % args.markers,
          "timing": "host wall clock of one pass through kasp_pipeline.design_markers_arrays (host buffers)",
          "thal_evaluations_rank0": int(sum(c.thal_item_count() for c in kctxs) - items0),
          "gpu_launches": int(sum(c.launch_count() for c in kctxs) - launches0),
          "host_threads": workers,
          "picker_tasks": int(stats.get("p3_tasks", 0)), "picker_pairs": int(stats.get("p3_pairs", 0)),
          "markers_with_pairs": int((res["n_sel"] > 0).sum()), "pairs_ranked": int(len(res["f"]))}
    return c5, dt * 1e3


def run_b200(args, rank, world, local_rank):
    import torch
    import polyoligo_b200 as po
    from polyoligo_b200 import synth

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = po.Context(local_rank)   # raises if the CUDA library is missing: no CPU fallback
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    n = args.pairs

    # ---- workload: pinned host records and device-resident copies ------------
    a_np, b_np = synth.config4(n, seed=synth.CONFIG4_SEED + rank)
    h_a = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n * 16, dtype=torch.uint8, pin_memory=True)
    h_a.numpy()[:] = a_np.view(np.uint8)
    h_b.numpy()[:] = b_np.view(np.uint8)
    d_a = h_a.cuda(non_blocking=True)
    d_b = h_b.cuda(non_blocking=True)
    d_het = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    d_hpa = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    d_hpb = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    d_tma = torch.empty(n, dtype=torch.float64, device="cuda")
    d_tmb = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    # a dedicated non-default stream: its handle is passed to the C ABI, and the CUDA
    # events below are recorded on the same stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def step_dev(events=None):
        """One step on resident inputs; optional per-kernel CUDA events."""
        def mark(k):
            if events is not None:
                events[k].record(stream)
        mark(0)
        ctx.thal_batch_dev(po.THAL_ANY, d_a.data_ptr(), d_b.data_ptr(), n, d_het.data_ptr(), sp)
        mark(1)
        ctx.thal_batch_dev(po.THAL_HAIRPIN, d_a.data_ptr(), None, n, d_hpa.data_ptr(), sp)
        ctx.thal_batch_dev(po.THAL_HAIRPIN, d_b.data_ptr(), None, n, d_hpb.data_ptr(), sp)
        mark(2)
        ctx.tm_batch_dev(d_a.data_ptr(), n, d_tma.data_ptr(), None, sp)
        ctx.tm_batch_dev(d_b.data_ptr(), n, d_tmb.data_ptr(), None, sp)
        mark(3)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed: algorithmic work of one step and the compute peak ------------
    relax_het = ctx.count_relaxations(po.THAL_ANY, a_np, b_np)
    fp64_peak = ctx.measure_fp64_peak()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()

    # ---- timed region 1: device-resident -----------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record(stream)
    for k in range(args.steps):
        step_dev(ev[k])
    t_end.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    dev_ms = t_begin.elapsed_time(t_end)
    het_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    hp_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    tm_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))

    # ---- timed region 2: end to end through the public host-buffer API -------------
    h_out = {"het": torch.empty(n * 40, dtype=torch.uint8, pin_memory=True),
             "hairpin_a": torch.empty(n * 40, dtype=torch.uint8, pin_memory=True),
             "hairpin_b": torch.empty(n * 40, dtype=torch.uint8, pin_memory=True),
             "tm_a": torch.empty(n, dtype=torch.float64, pin_memory=True),
             "tm_b": torch.empty(n, dtype=torch.float64, pin_memory=True)}
    out_np = {k: (v.numpy().view(po.THAL_OUT_DTYPE) if v.dtype == torch.uint8 else v.numpy()) for k, v in h_out.items()}
    a_pin = h_a.numpy().view(po.OLIGO_DTYPE)
    b_pin = h_b.numpy().view(po.OLIGO_DTYPE)
    ctx.score_pairs(a_pin, b_pin, out=out_np)  # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        ctx.score_pairs(a_pin, b_pin, out=out_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    # results of the two paths must agree (same kernels): cheap integrity check
    het_dev = d_het.cpu().numpy().view(po.THAL_OUT_DTYPE)
    assert het_dev.tobytes() == out_np["het"].tobytes(), "device-resident and host-buffer paths disagree"

    # ---- secondary: BASELINE config 5, KASP design of synthetic markers (markers/s) ------------
    c5 = None
    c5_ms = 0.0
    if args.markers > 0:
        c5, c5_ms = run_config5(args, rank, local_rank, barrier)

    # ---- reduce over ranks: max time ------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3, het_ms, hp_ms, tm_ms, c5_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, het_ms, hp_ms, tm_ms, c5_ms = [float(x) for x in times.cpu()]

    if rank == 0:
        evals = EVALS_PER_PAIR * n * world * args.steps
        value = evals / (dev_ms * 1e-3)
        e2e_value = evals / (e2e_ms * 1e-3)
        achieved = relax_het * FLOP_PER_RELAXATION / (het_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = n * (32 + 40)  # two records in, one result out per heterodimer
        roofline = {"bound": "fp64", "kernel": "k_thal<dimer> (heterodimer DP)",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                    "peak_source": "measured here: register-resident DFMA microbenchmark (po_measure_fp64_peak)",
                    "relaxations_per_launch": relax_het, "flop_per_relaxation": FLOP_PER_RELAXATION,
                    "kernel_ms": het_ms, "share_of_step": het_ms / (dev_ms / args.steps),
                    # ncu --set full, 400k-pair launch of this kernel: dram__bytes_read.sum 12.889 MB +
                    # dram__bytes_write.sum 0.019 MB (profiles/r1_v10_ncu_full_summary.txt) = 32.3 B per
                    # pair, i.e. the two 16-byte input records; scaled to this launch's pair count
                    "traffic": 32.31 * n,
                    "traffic_note": "DRAM bytes per launch scaled from the ncu capture of a 400k-pair launch "
                                    "(32.3 B/pair: inputs only, results stay in L2 within the capture); "
                                    "algorithmic bytes are 72 B/pair (32 in + 40 out)",
                    "hbm": {"achieved_gbs": hbm_bytes / (het_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "frac": hbm_bytes / (het_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650"},
                    # the binding resource: warp-instruction issue.  Instructions per pair come from the
                    # ncu capture (smsp__inst_executed.sum 23,692,710,893 for a 400k-pair launch,
                    # profiles/r1_v10_ncu_full_summary.txt); peak = 4 schedulers x SMs x SM clock
                    "issue_slots": {"warp_instructions_per_pair": 59230,
                                    "achieved_ginst_s": 59230 * n / (het_ms * 1e-3) / 1e9,
                                    "peak_ginst_s": 4 * sm_count * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                                    "frac": 59230 * n / (het_ms * 1e-3) /
                                            (4 * sm_count * (clocks.get("sm_mhz") or 1965.0) * 1e6),
                                    "source": "instruction count from the ncu capture of the same kernel, time from "
                                              "this run's CUDA events"},
                    "note": "compute-bound max-plus style DP on CUDA cores; not an HBM or tensor-core kernel (SURVEY 8d); "
                            "the kernel is bound by instruction issue (issue_slots), the FP64 share of the "
                            "instructions is small by construction"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 136 * n,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "kernel_ms": {"heterodimer": het_ms, "hairpin_x2": hp_ms, "tm_x2": tm_ms},
                "heterodimers_per_s": n * world / (het_ms * 1e-3), "hairpins_per_s": 2 * n * world / (hp_ms * 1e-3)}
        if c5 is not None:
            c5["markers_per_s"] = args.markers * world / (c5_ms * 1e-3)
            c5["ms"] = c5_ms
            c5["thal_evaluations_per_s"] = c5["thal_evaluations_rank0"] * world / (c5_ms * 1e-3)
            line["config5_kasp_design"] = c5
        if world == 1 and not args.no_cpu_baseline:
            from oracle.oracle import Oracle, build
            build()
            orc = Oracle()
            threads = os.cpu_count() or 1
            a_str, b_str = cpu_sample(args)
            cpu_score(orc, a_str[:2000], b_str[:2000], threads)
            dt = cpu_score(orc, a_str, b_str, threads)
            # cross-check a prefix of the GPU batch against the oracle
            ref = orc.thal_batch(synth.to_strings(a_np[:20000]), synth.to_strings(b_np[:20000]), 1, nthreads=threads)
            got = out_np["het"][:20000]
            parity = bool(np.array_equal(got["tm"], ref["tm"]) and np.array_equal(got["dg"], ref["dg"]))
            if c5 is not None and args.cpu_markers > 0:
                # the same pipeline with the CPU oracle's context (C-oracle thal on all cores, lazy picker)
                from oracle.cpu_pipeline import CpuContext
                from polyoligo_b200 import kasp_pipeline as kp
                cdata = kp.synth_config5(args.cpu_markers, seed=kp.CONFIG5_SEED)
                cctx = CpuContext(nthreads=threads)
                t0 = time.perf_counter()
                cres = kp.design_markers_arrays(cdata, n_primers=10, ctx=cctx,
                                                reporters=("GAAGGTCGGAGTCAACGGATT", "GAAGGTGACCAAGTTCATGCT"))
                cdt = time.perf_counter() - t0
                c5["cpu_baseline"] = {"value": args.cpu_markers / cdt, "unit": "markers/s", "cores": threads,
                                      "kind": "port", "thal_evaluations": int(cctx.thal_item_count()),
                                      "pairs_ranked": int(len(cres["f"])),
                                      "sample": "first %d markers of the same generator, one pass of the same "
                                                "pipeline through oracle/cpu_pipeline.py" % args.cpu_markers}
            line["cpu_baseline"] = {"value": EVALS_PER_PAIR * args.cpu_sample / dt, "unit": UNIT, "cores": threads,
                                    "kind": "port",
                                    "sample": "config-4 batch of %d pairs, one pass (same generator, smaller n)" % args.cpu_sample,
                                    "gpu_matches_oracle_on_sample": parity}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
