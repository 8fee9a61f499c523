#!/usr/bin/env python
"""bench.py -- spectra/s and candidates scored/s of the MaxDecoy identification hot path on B200.

A step = one pass of the hot path (precursor windows -> index window search -> ModifiedPeptide filter ->
decoy generation -> fragment scoring -> top-k PSM rows [-> NCCL gather of the PSM tables]) over one batch of
synthetic spectra, against the peptide index that `digest` + `index_build` left resident in HBM (the
reference keeps that state in PostgreSQL between its `digest` and `identification` subcommands).

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA through the C ABI)
  python bench.py --impl reference [...]                     the CPU arm: the oracle port of the reference's
                                                             algorithm on the host cores (the Rust reference cannot be
                                                             built here: no cargo/PostgreSQL/Comet)
N > 1 is launched by torchrun (one rank per GPU); spectra are sharded (weak scaling: every rank gets its own
batch), the index is replicated, and the only collective is the all_gather of the PSM tables.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "max-decoy_b200"))

CONFIGS = {
    # BASELINE.json configs[0]: the reference's CPU-runnable case
    "c1": dict(n_proteins=2000, mc=1, n_spectra=1000, ppm=10, var=False, n_decoys=1000,
               text="C1: synthetic 2k-protein FASTA, trypsin MC=1, len 5..50, fixed CAM-C, 1k synthetic spectra, 10 ppm, 1000 reference-random decoys/spectrum"),
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(n_proteins=20000, mc=2, n_spectra=10000, ppm=10, var=False, n_decoys=1000,
               text="C2: synthetic human-proteome-sized FASTA (20k proteins), trypsin MC=2, len 5..50, fixed CAM-C, 10k synthetic spectra, 10 ppm, target + 1000 reference-random decoys/spectrum (reference default -d 1000), top-5 PSMs"),
    "c3": dict(n_proteins=20000, mc=2, n_spectra=10000, ppm=10, var=True, n_decoys=1000,
               text="C3: C2 + variable Met-oxidation (<=3 mods/peptide), 1000 mass-matched decoys/spectrum"),
    # BASELINE.json configs[4] on one GPU's share: open search, ~1M candidates per spectrum (no decoys: the window holds them all)
    "c5": dict(n_proteins=20000, mc=2, n_spectra=64, ppm=10, var=False, n_decoys=0, abs_da=500,
               text="C5 (one GPU's share): +-500 Da open search over the 20k-protein index, ~1M candidates per spectrum, targets only"),
}
TOP_K = 5
FRAG_TOL = 0.02


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.mark_at = 0

    def mark(self):
        """Start of the timed region: samples from here on are the ones reported (the sampler itself is started before
        the warm-up, because nvidia-smi needs longer than a short timed region to deliver its first line)."""
        self.mark_at = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[self.mark_at:]
        where = "timed region"
        if not lines:   # timed region shorter than one sampling period: the samples under load just before it (warm-up)
            lines, where = self.lines[-5:], "warm-up (timed region shorter than one sample)"
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled_during": where}


def make_workload(cfg, rank, n_spectra):
    from maxdecoy import synth
    prots = synth.synthetic_proteins(cfg["n_proteins"])
    mods = [synth.CAM, synth.OXM] if cfg["var"] else [synth.CAM]
    sp, _ = synth.synthetic_spectra(prots, n_spectra, cfg["mc"], mods=tuple(mods), seed=7 + 1000 * rank)
    return prots, mods, sp


def search_params(cfg):
    from maxdecoy import SearchParams
    a = int(cfg.get("abs_da", 0) * 1000000)
    return SearchParams(cfg["ppm"], cfg["ppm"], fragment_tolerance=FRAG_TOL, n_decoys=cfg["n_decoys"], decoy_mode=0, seed=20260101,
                        top_k=TOP_K, min_peaks=10, max_fragment_charge=3, abs_lower_uda=a, abs_upper_uda=a)


def cpu_identify(cfg, prots, mods, sp, sample, threads, repeats=1):
    """The oracle (CPU port of the reference's algorithm) on the first `sample` spectra; returns (spectra/s, pairs/s, stats)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle_engine
    import numpy as np
    e = oracle_engine(threads)
    t0 = time.time()
    e.digest(prots, cfg["mc"], 5, 50)
    e.set_modifications(mods, 3 if cfg["var"] else 0)
    e.index_build()
    t_index = time.time() - t0
    sub = sp.subset(np.arange(min(sample, len(sp))))
    best = None
    for _ in range(repeats):
        t0 = time.time()
        _, st = e.identify(sub, search_params(cfg))
        dt = time.time() - t0
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    e.close()
    return len(sub) / dt, (st["n_targets"] + st["n_decoys"]) / dt, st, t_index, len(sub)


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_sample
    prots, mods, sp = make_workload(cfg, 0, sample)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle_engine
    e = oracle_engine(threads)
    e.digest(prots, cfg["mc"], 5, 50)
    e.set_modifications(mods, 3 if cfg["var"] else 0)
    e.index_build()
    prm = search_params(cfg)
    for _ in range(args.warmup):
        e.identify(sp, prm)
    t0 = time.time()
    pairs = 0
    for _ in range(args.steps):
        _, st = e.identify(sp, prm)
        pairs += st["n_targets"] + st["n_decoys"]
    dt = time.time() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": "spectra_per_sec", "value": v, "unit": "spectra/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "candidates_per_sec": pairs / dt,
            "config": {"workload": cfg["text"], "fragment_tolerance_da": FRAG_TOL, "top_k": TOP_K},
            "cpu_baseline": {"value": v, "unit": "spectra/s", "cores": threads, "kind": "port",
                             "sample": "%d spectra of the workload per step, %d host threads; CPU oracle port of the reference's algorithm with an in-memory "
                                       "mass-sorted index (the Rust reference needs cargo + PostgreSQL + Comet, none present)" % (sample, threads)},
            "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """Keep stdout for the ONE result line: everything else any library prints to fd 1 (e.g. NCCL's version banner)
    goes to stderr."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--spectra", type=int, default=0, help="override spectra per rank")
    ap.add_argument("--decoys", type=int, default=-1, help="override decoys per spectrum")
    ap.add_argument("--cpu-sample", type=int, default=512, help="spectra of the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=128, help="spectra per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg = dict(CONFIGS[args.config])
    if args.decoys >= 0:
        cfg["n_decoys"] = args.decoys
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import maxdecoy
    from maxdecoy import _abi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_spec = args.spectra or cfg["n_spectra"]
    t0 = time.time()
    prots, mods, sp = make_workload(cfg, rank, n_spec)
    sp.spectrum_id = (np.arange(n_spec, dtype=np.uint32) + np.uint32(rank * n_spec))
    t_gen = time.time() - t0

    eng = maxdecoy.Engine(device=local_rank)
    buf, off = maxdecoy.pack_proteins(prots)
    t0 = time.time(); n_pep = eng.digest_packed(buf, off, cfg["mc"], 5, 50); t_digest = time.time() - t0
    eng.set_modifications(mods, 3 if cfg["var"] else 0)
    t0 = time.time(); eng.index_build(); t_index = time.time() - t0
    istats = eng.index_stats()
    prm = search_params(cfg)

    ext = torch.cuda.ExternalStream(eng.lib.md_stream_handle(eng.h), device=torch.device("cuda", local_rank))
    # ---- device-resident inputs (value) and pinned host inputs (e2e)
    def as_torch(a):  # unsigned 32/64-bit arrays travel as their signed views (same bytes)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        elif a.dtype == np.uint32:
            a = a.view(np.int32)
        return torch.from_numpy(a).pin_memory()
    host = {k: as_torch(getattr(sp, k)) for k in ("precursor_mz", "charge", "spectrum_id", "peak_off", "peak_mz", "peak_intensity")}
    dev = {k: v.cuda(non_blocking=False) for k, v in host.items()}
    sd = _abi.md_spectra()
    sd.n = n_spec
    for k in host:
        setattr(sd, k, dev[k].data_ptr())
    sh = _abi.md_spectra()
    sh.n = n_spec
    for k in host:
        setattr(sh, k, host[k].data_ptr())
    psm_bytes = n_spec * TOP_K * 56
    # PSM tables, double-buffered: the gather of one batch runs (NCCL's stream) while the next batch is searched
    psm_devs = [torch.empty(psm_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    psm_alls = [torch.empty(psm_bytes * world, dtype=torch.uint8, device="cuda") for _ in range(2)] if world > 1 else None
    psm_dev = psm_devs[0]
    gathers = [None, None]
    tick = [0]
    psm_host = torch.empty(psm_bytes, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    import ctypes as C
    pc = prm.to_c()

    def gather_async(k):
        """all_gather of PSM buffer k, asynchronous: the library's stream goes on with the next batch; the buffer is
        waited for before it is written again (two batches later) and at the end of the timed region."""
        gathers[k] = dist.all_gather_into_tensor(psm_alls[k], psm_devs[k], async_op=True)

    def drain_gathers():
        for k in (0, 1):
            if gathers[k] is not None:
                gathers[k].wait()
                gathers[k] = None

    def step_device():
        k = tick[0] & 1
        tick[0] += 1
        if gathers[k] is not None:
            gathers[k].wait()
            gathers[k] = None
        st = eng.identify_device(sd, prm, psm_devs[k].data_ptr())
        if world > 1:
            gather_async(k)
        return st

    def step_host():
        k = tick[0] & 1
        tick[0] += 1
        st = _abi.md_identify_stats()
        rc = eng.lib.md_identify(eng.h, C.byref(sh), C.byref(pc), C.c_void_p(psm_host.data_ptr()), C.byref(st), None, None)
        if rc != 0:
            raise RuntimeError(eng.lib.md_last_error(eng.h).decode())
        if world > 1:
            if gathers[k] is not None:
                gathers[k].wait()
                gathers[k] = None
            psm_devs[k].copy_(psm_host, non_blocking=True)
            gather_async(k)
        return eng._stats(st)

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the library's stream; L2 flushed (untimed) before each."""
        total_ms, last = 0.0, None
        with torch.cuda.stream(ext):
            for _ in range(steps):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(ext)
                last = fn()
                b.record(ext)
                b.synchronize()
                total_ms += a.elapsed_time(b)
            if world > 1:   # the gathers still in flight belong to the timed steps
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(ext)
                drain_gathers()
                b.record(ext)
                b.synchronize()
                total_ms += a.elapsed_time(b)
        return total_ms, last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local_rank)
    clocks.start()
    with torch.cuda.stream(ext):
        for _ in range(args.warmup):
            step_device()
        drain_gathers()
    barrier()
    clocks.mark()
    wall0 = time.time()
    ms_dev, st = timed(step_device, args.steps)
    barrier()
    wall = time.time() - wall0
    clk = clocks.stop()
    ms_dev = max_over_ranks(ms_dev)

    with torch.cuda.stream(ext):
        step_host()
        drain_gathers()
    barrier()
    ms_e2e, st_h = timed(step_host, args.steps)
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)

    pairs = st["n_pairs"]
    if world > 1:
        t = torch.tensor([pairs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        pairs_all = float(t.item())
    else:
        pairs_all = float(pairs)
    def shutdown():
        """Ordered teardown, then a hard exit: the library and torch/NCCL each own CUDA state whose static destructors
        must not race at interpreter exit (seen as SIGSEGV after the result line with N > 1)."""
        eng.close()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        shutdown()

    peaks, peak_kind = load_peaks()
    total_spectra = n_spec * world
    value = total_spectra * args.steps / (ms_dev / 1e3)
    ach = (st["score_bytes"] / 1e9) / (st["ms_kernel_score"] / 1e3) if st["ms_kernel_score"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k_score_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    line = {
        "metric": "spectra_per_sec", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": cfg["text"], "spectra_per_gpu": n_spec, "proteins": cfg["n_proteins"], "unique_peptides": n_pep,
                   "index_device_mb": round(istats["device_bytes"] / 1e6, 1), "fragment_tolerance_da": FRAG_TOL, "top_k": TOP_K,
                   "l2": "flushed (256 MiB write) before every timed step; flush not timed", "parallelism": "spectra sharded x%d, index replicated, PSM all_gather per batch (asynchronous, double-buffered: overlaps the next batch)" % world},
        "candidates_per_sec": pairs_all * args.steps / (ms_dev / 1e3),
        "pairs_per_step": pairs_all,
        "gpu_launches": int(st["n_kernel_launches"]) * args.steps,
        "e2e": {"value": total_spectra * args.steps / (ms_e2e / 1e3), "unit": "spectra/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(psm_bytes),
                "ms_per_step": ms_e2e / args.steps},
        "roofline": {"kernel": "k_score (fused fragment-and-score + top-k)", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "algorithmic_bytes_per_launch": int(st["score_bytes"]), "launch_ms": st["ms_kernel_score"], "pairs_per_launch": int(st["n_pairs"]),
                     "note": "algorithmic bytes = sum over scored pairs of (14 + peptide length); the kernel is integer-issue / shared-memory-gather bound, see DESIGN.md section 5"},
        "stage_ms_per_step": {"lookup": st["ms_lookup"], "decoys": st["ms_decoys"], "score": st["ms_score"], "kernel_decoy_attempts": st["ms_kernel_decoy"],
                              "kernel_score": st["ms_kernel_score"], "decoy_attempts": int(st["n_attempts"])},
        "one_time": {"digest_s": t_digest, "index_build_s": t_index, "synthetic_generation_s": t_gen},
        "clocks": clk, "wall_s_timed_region": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, pv, cst, _, ns = cpu_identify(cfg, prots, mods, sp, args.cpu_sample, os.cpu_count() or 1)
        line["cpu_baseline"] = {"value": v, "unit": "spectra/s", "cores": os.cpu_count() or 1, "kind": "port", "candidates_per_sec": pv,
                                "sample": "first %d spectra of the same workload, one pass, %d host threads over spectra; CPU oracle port with an in-memory "
                                          "mass-sorted index (favourable to the CPU: the real reference adds PostgreSQL round trips and an external Comet run)" % (ns, os.cpu_count() or 1)}
    emit(line)
    shutdown()


if __name__ == "__main__":
    main()
