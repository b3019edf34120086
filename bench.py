#!/usr/bin/env python3
"""bench.py -- genome pairs/sec for ANI+AF behind skDER's `skani triangle` call site.

  python bench.py --gpus N --steps K --warmup W            (ours; one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W    (CPU arm: the oracle port on all host threads)

A step = one pass of the hot path over one synthetic genome set (BASELINE.json configs):
  value : pairs/s with sketches already resident in HBM (prescreen + ANI/AF + edge gather), device-timed
  e2e   : pairs/s through the C-ABI from packed genomes in pinned HOST memory to edges on the host
          (H2D upload + sketch + index + prescreen + ANI/AF + D2H), every step.
The workload is the configuration BASELINE.json quotes its metric on (N=5k x 5 Mbp = configs[2], `config3`; it fits one GPU);
N>1 shards the same triangle's rows
round-robin over the ranks (no data-path collective; sketches are replicated by NCCL all-gather
before the timed region of `value`, inside it for `e2e`).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GREEDY_SCREEN = 89.5  # skDER default: -s (ANI cutoff 99.5 - 10)   reference bin/skder:199-204 (greedy AND dynamic mode;
                      # dynamic lowers --min-af by 20 only together with -n, bin/skder:218-219)
GREEDY_MIN_AF = 50.0  # skDER default AF cutoff                    reference bin/skder:325-329
# bounded sample of the workload the CPU arm is timed on (same generator, same clade structure): 400 genomes,
# 79,800 pairs, 1,800 survivors -- roughly 30 core-seconds of oracle work
CPU_SAMPLE_CLADES, CPU_SAMPLE_PER_CLADE = 40, 10


def workload_shape(name):
    from skder_b200 import synth

    nc, per, L, Lhi, *_ = synth.CONFIGS[name]
    return nc, per, L


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "250"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/skani_oracle.c) -- the reference's skani is not installable here
# ------------------------------------------------------------------------------------------------
def cpu_triangle_components(workload, n_clades, per_clade, threads, screen, min_af):
    """Time the oracle on a bounded sample; returns component costs per unit.  Sketching runs one genome per host
    thread; the all-vs-all (prescreen of every pair, ANI/AF of the survivors) runs inside the C library on `threads`
    pthreads (oracle/skani_oracle.c ora_triangle), so no Python per-pair overhead is in the measurement."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    from skder_b200 import synth

    O.build()
    nc, per, L, Lhi, dlo, dhi, seed = synth.CONFIGS[workload]
    with ThreadPoolExecutor(threads) as ex:
        clades = list(ex.map(lambda c: synth.one_clade(c, per_clade, L, seed, dlo, dhi, 300, Lhi), range(n_clades)))
    gens = [g for cl in clades for g in cl]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        sk = list(ex.map(O.Sketch.from_contigs, gens))
    t_sketch = time.perf_counter() - t0
    n = len(sk)
    r = O.triangle(sk, screen / 100.0, min_af / 100.0, threads)
    return {"n": n, "pairs": n * (n - 1) // 2, "survivors": r["survivors"], "edges": r["edges"], "t_sketch": t_sketch,
            "t_index": r["t_index"], "t_count": r["t_count"], "t_screen": r["t_index"] + r["t_count"], "t_ani": r["t_ani"]}


def cpu_extrapolate(comp, n_full, pairs_full, surv_full, with_sketch):
    # inverted-index prescreen: the key sort scales with the genomes, run counting with the shared markers (i.e. with
    # the surviving, within-clade pairs); ANI/AF with the surviving pairs
    t = (comp["t_count"] + comp["t_ani"]) / max(comp["survivors"], 1) * surv_full
    if with_sketch:  # end to end: sketching and the marker index are part of the step, as they are in our e2e
        t += (comp["t_sketch"] + comp["t_index"]) / comp["n"] * n_full
    return pairs_full / t, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    nc, per, L = workload_shape(args.workload)
    n_full = nc * per
    pairs_full = n_full * (n_full - 1) // 2
    surv_full = nc * per * (per - 1) // 2
    s_clades, s_per = min(nc, CPU_SAMPLE_CLADES), CPU_SAMPLE_PER_CLADE
    vals = []
    comp = None
    for it in range(args.warmup + args.steps):
        comp = cpu_triangle_components(args.workload, s_clades, s_per, threads, GREEDY_SCREEN, GREEDY_MIN_AF)
        if it >= args.warmup:
            vals.append(cpu_extrapolate(comp, n_full, pairs_full, surv_full, with_sketch=True))
    value = float(np.mean([v for v, _ in vals]))
    t_full = float(np.mean([t for _, t in vals]))
    sample = ("%d clades x %d members of %s (%d genomes, %d pairs, %d survive the screen): sketch %.2fs, inverted-index prescreen %.2fs, "
              "ANI/AF %.2fs on %d threads; per-unit costs scaled to the full workload (%d genomes, %d pairs, %d survivors)"
              % (s_clades, s_per, args.workload, comp["n"], comp["pairs"], comp["survivors"], comp["t_sketch"],
                 comp["t_screen"], comp["t_ani"], threads, n_full, pairs_full, surv_full))
    line = {
        "impl": "reference", "metric": "genome pairs/sec ANI+AF", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64/int32 + f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "screen": GREEDY_SCREEN, "min_af": GREEDY_MIN_AF,
                   "note": "oracle-CPU (C port of the published skani method), NOT the skani binary: skani is absent"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_name(w):
    nc, per, L = workload_shape(w)
    return "%s: %d synthetic %.1f Mbp genomes (%d clades x %d, 95-99.9%% ANI within clade), skDER default thresholds: greedy and dynamic mode both run `skani triangle -s 89.5 --min-af 50 -E`" % (
        w, nc * per, L / 1e6, nc, per)


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def pinned_views(packed, torch):
    """Copy every genome's 2-bit words into ONE pinned host buffer and return skb_packed views into it."""
    from skder_b200 import _lib

    total = sum(p.n_words for p in packed)
    pin = torch.empty(total, dtype=torch.int64).pin_memory()
    arr = pin.numpy().view(np.uint64)
    views, keep, off = [], [], 0
    for p in packed:
        nw = p.n_words
        arr[off:off + nw] = p.words()
        v = _lib.Packed()
        v.words = C.cast(arr[off:].ctypes.data, C.POINTER(C.c_uint64))
        v.n_words, v.n_bases, v.n_contigs = nw, p.n_bases, p.n_contigs
        lens = (C.c_int64 * max(1, p.n_contigs))(*p.contig_lens().tolist())
        v.contig_lens = C.cast(lens, C.POINTER(C.c_int64))
        v.first_name, v.n50, v.total_bases_all = None, p.n50, p.total_bases_all
        keep.append(lens)
        views.append(v)
        off += nw
    return pin, views, keep, total * 8


def run_ours(args):
    import torch
    import torch.distributed as dist

    from skder_b200 import _lib, build, engine, synth

    build.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torchrun --nproc-per-node %d" % (args.gpus, world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: bench.py has no CPU path for --impl ours")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nc, per, L = workload_shape(args.workload)
    n_full = nc * per
    pairs_full = n_full * (n_full - 1) // 2
    # every rank ingests its share of the clades (rank r: clades r, r+world, ...)
    from skder_b200 import multi

    t_gen = time.perf_counter()
    my_clades = list(range(rank, nc, world))
    packed = []
    # clade-parallel generation on host threads
    from concurrent.futures import ThreadPoolExecutor

    ncfg = synth.CONFIGS[args.workload]

    def gen(c):
        return [engine.pack_contigs(g) for g in synth.one_clade(c, per, ncfg[2], ncfg[6], ncfg[4], ncfg[5], 300, ncfg[3])]

    with ThreadPoolExecutor(max(1, min(32, (os.cpu_count() or 1) // max(world, 1)))) as ex:
        for clade in ex.map(gen, my_clades):
            packed += clade
    t_gen = time.perf_counter() - t_gen
    pin, views, keep, h2d_bytes = pinned_views(packed, torch)
    del packed
    arr = (C.POINTER(_lib.Packed) * len(views))(*[C.pointer(v) for v in views])

    eng = engine.Engine(local)
    L_ = eng._L

    phase_ms = []  # (add, replicate, index) wall ms of every sketch_all() call on this rank

    def sketch_all():
        """host packed genomes -> replicated, indexed sketch DB on every rank"""
        t0 = time.perf_counter()
        eng.clear()
        eng._ck(L_.skb_add_genomes(eng._h, len(views), arr), "skb_add_genomes")
        t1 = time.perf_counter()
        if world > 1:
            multi.replicate_sketches(eng, dist, torch)
        t2 = time.perf_counter()
        eng.index()
        t3 = time.perf_counter()
        phase_ms.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
        if os.environ.get("SKB_BENCH_DEBUG") == "1" and rank == 0:
            sys.stderr.write("add %.2f ms  replicate %.2f ms  index %.2f ms\n" % phase_ms[-1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dbg = os.environ.get("SKB_BENCH_DEBUG") == "1"

    def triangle():
        t0 = time.perf_counter()
        edges, st = eng.triangle(GREEDY_SCREEN, GREEDY_MIN_AF, part=rank, n_parts=world, to_host=world == 1)
        t1 = time.perf_counter()
        if world > 1:
            edges = multi.gather_device_edges(eng, dist, torch, sort=False)
        if dbg and rank == 0:
            sys.stderr.write("triangle %.2f ms (screen %.2f ani %.2f) gather %.2f ms\n" % (
                (t1 - t0) * 1e3, st.ms_screen, st.ms_ani, (time.perf_counter() - t1) * 1e3))
        return edges, st

    # ---- warm-up (full e2e steps)
    for _ in range(args.warmup):
        sketch_all()
        triangle()
    # ---- e2e: host packed genomes -> edges on host, every step
    with ClockSampler(local) as clk_e2e:
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            sketch_all()
            edges, st = triangle()
            d2h += edges.nbytes
        barrier()
        t_e2e = time.perf_counter() - t0
    # ---- value: sketches resident; device-timed on the library's stream
    l0 = eng.launches
    ms_ani, ms_screen, ms_anchor, n_anchor, st = [], [], [], [], None
    with ClockSampler(local) as clk:
        barrier()
        eng.timer_start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            edges, st = triangle()
            ms_ani.append(st.ms_ani)
            ms_screen.append(st.ms_screen)
            ms_anchor.append(st.ms_anchor)
            n_anchor.append(st.n_anchor_launches)
        ms_dev = eng.timer_stop()
        barrier()
        t_wall = time.perf_counter() - t0
    launches = eng.launches - l0
    if world > 1:
        tt = torch.tensor([ms_dev, t_e2e, float(np.mean(ms_ani))], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, t_e2e = float(tt[0]), float(tt[1])
        cnt = torch.tensor([st.n_pairs_screened, st.sum_query_seeds, st.sum_anchors, launches, h2d_bytes],
                           device="cuda", dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_screened, launches_all, h2d_all = int(cnt[0]), int(cnt[3]), int(cnt[4])
    else:
        tot_screened, launches_all, h2d_all = st.n_pairs_screened, launches, h2d_bytes
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_step = ms_dev / args.steps
    value = pairs_full / (ms_step / 1e3)
    e2e_value = pairs_full / (t_e2e / args.steps)
    # ---- roofline.  Algorithmic bytes per surviving pair for the whole pair stage = 32*S_q + 32*A + 20
    # (SURVEY.md section 8d / BASELINE.md section 4).  The dominant kernel is anchor_kernel; of the model's terms it
    # owns the query seed records (16*S_q), one index slot per probe (16*S_q) and the anchors written (16*A).
    # Duration = that kernel's average launch time from CUDA events the library records around every launch.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    stage_bytes = 32 * st.sum_query_seeds + 32 * st.sum_anchors + 20 * st.n_pairs_screened
    ani_ms = float(np.mean(ms_ani))
    stage_gbs = stage_bytes / (ani_ms / 1e3) / 1e9 if ani_ms > 0 else 0.0
    launches_per_step = max(1, int(np.mean(n_anchor)))
    anchor_ms = float(np.mean(ms_anchor)) / launches_per_step
    alg_bytes = (32 * st.sum_query_seeds + 16 * st.sum_anchors) / launches_per_step
    achieved = alg_bytes / (anchor_ms / 1e3) / 1e9 if anchor_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "anchor_kernel_traffic.json")))
        if tj.get("workload") == args.workload and world == 1:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N=1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        comp = cpu_triangle_components(args.workload, min(nc, CPU_SAMPLE_CLADES), CPU_SAMPLE_PER_CLADE, threads, GREEDY_SCREEN,
                                       GREEDY_MIN_AF)
        surv_full = nc * per * (per - 1) // 2
        v, t_full = cpu_extrapolate(comp, n_full, pairs_full, surv_full, with_sketch=False)
        cpu = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "%d clades x %d members of %s (%d pairs, %d survivors): prescreen run counting %.2fs, ANI/AF %.2fs on "
                         "%d threads; sketches and the marker index resident (as for `value`); per-unit costs scaled to "
                         "the full workload" % (min(nc, CPU_SAMPLE_CLADES), CPU_SAMPLE_PER_CLADE, args.workload, comp["pairs"],
                                                comp["survivors"], comp["t_count"], comp["t_ani"], threads)}
    line = {
        "metric": "genome pairs/sec ANI+AF", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64/int32 + f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "screen": GREEDY_SCREEN, "min_af": GREEDY_MIN_AF,
                   "pairs": pairs_full, "pairs_screened": tot_screened, "edges": int(len(edges)),
                   "l2": "inputs larger than L2 (sketch DB %.1f GB per rank)" % (
                       (st.sum_query_seeds and (n_full * L / 125 * 8 * 3) / 1e9) or 0.0),
                   "ms_screen": float(np.mean(ms_screen)), "ms_ani": ani_ms, "gen_s": t_gen,
                   "wall_ms_per_step": t_wall / args.steps * 1e3},
        "clocks": clk.summary(), "clocks_e2e": clk_e2e.summary(),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h // args.steps,
                "ms_per_step": t_e2e / args.steps * 1e3,
                # rank 0's host-clock phases of the timed end-to-end steps (the rest of a step is skb_triangle + gather)
                "ms_upload_sketch": float(np.mean([p[0] for p in phase_ms[-args.steps:]])),
                "ms_replicate": float(np.mean([p[1] for p in phase_ms[-args.steps:]])),
                "ms_index": float(np.mean([p[2] for p in phase_ms[-args.steps:]]))},
        "gpu_launches": launches_all,
        "roofline": {"bound": "hbm", "kernel": "anchor_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": alg_bytes,
                     "ms_per_launch": anchor_ms, "launches_per_step": launches_per_step,
                     "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                     "note": "bound by the L1 data pipe (scattered 32-byte bucket reads) and issue; reference tables are L2-resident; "
                             "see DESIGN.md section 4"},
        "roofline_stage": {"kernels": "task_setup + anchor + chain + ends + finalize", "algorithmic_bytes": stage_bytes,
                           "ms": ani_ms, "achieved": stage_gbs, "unit": "GB/s", "frac": stage_gbs / peak},
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
