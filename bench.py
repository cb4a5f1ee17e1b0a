#!/usr/bin/env python3
"""bench.py -- DCTZ hot path (block DCT-II/IDCT + adaptive binning quantiser) on 1..8 B200.

A "step" is one pass of the hot path over one slab of synthetic input: COMPRESS (statistics ->
[NCCL all-gather of 3 doubles per rank] -> fused scale+DCT+quantise+outlier compaction) followed by
DECOMPRESS (outlier scan + dequantise + IDCT + de-scale).  The metric is BASELINE.json's
"compress/decompress GB/s of input": input bytes of the slab divided by the time of the
compress+decompress round trip, summed over ranks; the two halves are also reported separately.

Default workload (config.workload = "c5-slab"): each rank owns a contiguous block slab of 2^30
doubles (8 GiB) of BASELINE.json's config[4] field (2048^3 double, error bound 1E-3, EC mode,
generated on the device by the exactly reproducible formula of SURVEY.md §8d).  At --gpus 8 the ranks
together hold exactly the 64 GiB field; fewer ranks hold its first slabs (weak scaling).
Other configs can be timed with --workload c1|c2|c3|c4 (device-resident numbers only).

`--impl reference` times the reference's own CPU implementation of the same path (the unmodified
sources compiled into oracle/_ref, FFTW replaced by the stand-in because FFTW3 is not installed) on
the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "compress+decompress round-trip GB/s of input (hot path: stats, scale, block DCT-II/IDCT, binning quantiser)"
UNIT = "GB/s"
EB = 1e-3
SEED = 20261018
HASH_DIM = 2048
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["c5-slab", "c1", "c2", "c3", "c4"], default="c5-slab")
    ap.add_argument("--slab-log2", type=int, default=30, help="elements per rank of the c5 slab (default 2^30 = 8 GiB)")
    ap.add_argument("--e2e-log2", type=int, default=27, help="elements of the slab pushed through the host-buffer API for e2e")
    ap.add_argument("--cpu-log2", type=int, default=23, help="elements per process of the CPU reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / quality legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--f32", action="store_true", help="c5-slab only: cast the slab to float (single-precision path)")
    ap.add_argument("--qt", action="store_true", help="c5-slab only, one GPU: quantiser (QT) mode instead of error-bounded (EC)")
    ap.add_argument("--noise", type=float, default=0.0, help="c5-slab only: add Gaussian noise of this std (raw units) to study the outlier path")
    ap.add_argument("--no-outlier-leg", action="store_true", help="skip the extra (reported, not headline) measurement with ~5%% outliers")
    ap.add_argument("--watchdog", type=int, default=1500, help="seconds after which a stuck run dumps every thread's Python stack to stderr and exits 3 (0 = off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = unmodified reference sources + FFTW stand-in; else the oracle port)
# ------------------------------------------------------------------------------------------------
REF_BIN = {False: os.path.join(ROOT, "oracle", "_ref", "dctz-ec-test"), True: os.path.join(ROOT, "oracle", "_ref", "dctz-qt-test")}
_T = r"([0-9.eE+-]+)\(s\)"


def _parse_ref_stdout(txt):
    """Stage timers printed under -DTIME_DEBUG (dctz-comp-lib.c:762-773, dctz-decomp-lib.c:513-528)."""
    m1 = re.search(r"sf_t=" + _T + r", dct_t=" + _T + r", zlib_t\(compress\)=" + _T, txt)
    m2 = re.search(r"sf_t=" + _T + r", idct_t=" + _T + r", zlib_t\(uncompress\)=" + _T, txt)
    if not m1 or not m2:
        raise RuntimeError("cannot parse the reference's stage timers:\n" + txt[-2000:])
    cr = re.search(r"CR = ([0-9.]+)", txt)
    return dict(comp_hot=float(m1.group(1)) + float(m1.group(2)), comp_zlib=float(m1.group(3)),
                decomp_hot=float(m2.group(1)) + float(m2.group(2)), cr=float(cr.group(1)) if cr else None)


def run_reference_cli(sample_path, n, is_double, qt, procs):
    """Run `procs` independent copies of the reference's own CLI (one per host thread; its hot path is
    single-threaded, dct.c:18-22 is not re-entrant) on the same sample; returns per-process timers."""
    tmp = tempfile.mkdtemp(prefix="dctz_ref_")
    try:
        ps = []
        for i in range(procs):
            d = os.path.join(tmp, f"p{i}")
            os.makedirs(d)
            os.symlink(sample_path, os.path.join(d, "in.bin"))
            cmd = [REF_BIN[qt], "-d" if is_double else "-f", "1E-3", "var", os.path.join(d, "in.bin"), str(n)]
            ps.append(subprocess.Popen(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate()[0] for p in ps]
        for p, o in zip(ps, outs):
            if p.returncode != 0:
                raise RuntimeError(f"reference CLI failed ({p.returncode}):\n{o[-2000:]}")
        return [_parse_ref_stdout(o) for o in outs]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def time_port(x, qt):
    """Fallback when oracle/_ref is absent: the oracle port, single thread."""
    from tests import reflib

    t0 = time.perf_counter()
    o = reflib.oracle_compress(x, EB, qt, want_coef=False)
    t1 = time.perf_counter()
    reflib.oracle_decompress(o["bin_index"], o["dc"], o["ac"], o["qtable"], x.size, EB, o["stat"]["sf"], qt, x.dtype)
    t2 = time.perf_counter()
    return dict(comp_hot=t1 - t0, decomp_hot=t2 - t1, comp_zlib=None, cr=None)


def cpu_reference(sample, qt, procs):
    """Times the CPU implementation on `sample` (numpy array); aggregate GB/s of input over `procs`
    concurrent single-threaded instances."""
    import numpy as np

    n = sample.size
    nbytes = sample.nbytes
    have_ref = os.path.exists(REF_BIN[qt])
    if have_ref:
        tmp = tempfile.mkdtemp(prefix="dctz_sample_")
        try:
            path = os.path.join(tmp, "sample.bin")
            sample.tofile(path)
            res = run_reference_cli(path, n, sample.dtype == np.float64, qt, procs)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        kind = "reference"
    else:
        procs = 1
        res = [time_port(sample, qt)]
        kind = "port"
    tc = max(r["comp_hot"] for r in res)
    td = max(r["decomp_hot"] for r in res)
    gb = procs * nbytes / 1e9
    return dict(value=gb / (tc + td), unit=UNIT, cores=procs, kind=kind, compress_gbs=gb / tc, decompress_gbs=gb / td,
                host_cpus=os.cpu_count(), cr=res[0]["cr"],
                sample=f"{procs} concurrent single-thread instance(s), each {n} elements ({nbytes / 2**20:.0f} MiB) of the workload; "
                       f"hot-path stage timers only (sf_t+dct_t, idct_t+sf_t), zlib excluded"
                       + ("; FFTW3 replaced by oracle/fftw_standin" if have_ref else ""))


def make_sample(workload, n):
    import numpy as np

    from dctz_b200 import fields

    if workload == "c5-slab":
        return fields.hash_field(0, n, HASH_DIM, SEED), False
    x, qt = make_host_field(workload)
    return np.ascontiguousarray(x[:n]), qt


def make_host_field(workload):
    import numpy as np

    from dctz_b200 import fields

    if workload == "c1":
        return fields.cesm_like(), False
    if workload == "c2":
        return fields.cesm_like(dtype=np.float32), True
    if workload == "c3":
        return fields.hurricane_like(), False
    if workload == "c4":
        return fields.nyx_like(), False
    raise ValueError(workload)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = max(1, min(os.cpu_count() or 1, 32))
    n = 1 << args.cpu_log2
    sample, qt = make_sample(args.workload, n)
    for _ in range(args.warmup):
        cpu_reference(sample, qt, procs)
    t0 = time.perf_counter()
    runs = [cpu_reference(sample, qt, procs) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    # aggregate: total bytes / total hot-path time over the K steps
    inv = sum(1.0 / r["value"] for r in runs) / len(runs)
    val = 1.0 / inv
    base = runs[-1]
    base["value"] = val
    line = dict(metric=METRIC, value=val, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * wall / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64" if sample.dtype.itemsize == 8 else "f32", data="synthetic",
                config=workload_config(args, args.gpus), cpu_baseline=base,
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    _emit(line)
    return 0


def workload_config(args, world):
    if args.workload == "c5-slab":
        n = 1 << args.slab_log2
        mode = "qt" if getattr(args, "qt", False) else "ec"
        if getattr(args, "f32", False):
            return dict(workload=f"c5-slab-f32: per-GPU slab of 2^{args.slab_log2} floats (the {HASH_DIM}^3 field cast to float), {mode.upper()}, eb 1E-3",
                        mode=mode, error_bound=EB, elements_per_gpu=n, block=64, l2="inputs larger than L2 (no flush needed)",
                        parallelism=f"slab{world}")
        return dict(workload=f"c5-slab: per-GPU contiguous slab of 2^{args.slab_log2} doubles ({n * 8 / 2**30:.0f} GiB) of the "
                             f"{HASH_DIM}^3 double field (BASELINE config[4]), {mode.upper()} mode, error bound 1E-3; "
                             f"{world} slab(s) = {world * n * 8 / 2**30:.0f} GiB",
                    mode=mode, error_bound=EB, elements_per_gpu=n, block=64, l2="inputs larger than L2 (no flush needed)",
                    parallelism=f"slab{world}")
    desc = {"c1": "config[0] CESM-ATM-shaped 1800x3600 double, EC", "c2": "config[1] 1800x3600 float, QT",
            "c3": "config[2] Hurricane-shaped 100x500x500 float, EC", "c4": "config[3] NYX-shaped 512^3 double, EC"}[args.workload]
    return dict(workload=f"{args.workload}: {desc}, error bound 1E-3, one field per GPU", error_bound=EB, block=64,
                l2="L2 flushed between timed steps (write of a 256 MiB buffer)" if args.workload in ("c1", "c2", "c3") else
                   "inputs larger than L2 (no flush needed)", parallelism=f"replica{world}")


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up (its first line takes ~0.2 s to appear); stop() keeps
    the samples whose timestamp falls inside the timed region."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime

        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as fh:
            for ln in fh:
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(p[2]), float(p[3]), float(p[4]), [nm for nm, v in zip(names, p[6:10]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
        os.unlink(self.path)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        use = inside if inside else rows[-3:]
        if not use:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = sorted(r[1] for r in use)
        reasons = sorted({x for r in use for x in r[4]})
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in use), power_w_max=max(r[3] for r in use), samples=len(use),
                    samples_in_timed_region=len(inside), reasons=reasons)


def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import dctz_b200
    from dctz_b200 import DOUBLE, FLOAT, binding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:  # run (and allocate pinned host memory) on the CPU cores / NUMA node next to this rank's GPU
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = dctz_b200.Context(local)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    # ---- workload resident in HBM -------------------------------------------------------------
    qt = False
    if args.qt and (args.workload != "c5-slab" or world > 1):
        raise SystemExit("bench.py: --qt applies to the c5-slab workload on one GPU (c2 is the QT config)")
    if args.workload == "c5-slab":
        qt = bool(args.qt)
        n = 1 << args.slab_log2
        tdt, code, es = torch.float64, DOUBLE, 8
        x = torch.empty(n, dtype=tdt, device=dev)
        ctx.fill_hash_field(x.data_ptr(), rank * n, n, HASH_DIM, SEED, sh)
        if args.noise > 0:
            gen = torch.Generator(device=dev).manual_seed(SEED + rank)
            x += args.noise * torch.randn(n, generator=gen, device=dev, dtype=torch.float64)
        if args.f32:
            x = x.float()
            tdt, code, es = torch.float32, FLOAT, 4
        n_total, first = n * world, rank == 0
    else:
        host, qt = make_host_field(args.workload)
        n = host.size
        es = host.dtype.itemsize
        tdt, code = (torch.float64, DOUBLE) if es == 8 else (torch.float32, FLOAT)
        x = torch.from_numpy(host).to(dev)
        n_total, first = n, True  # replicas: every rank compresses its own copy of the field
    nblk = (n + 63) // 64
    bins = torch.empty(n, dtype=torch.uint8, device=dev)
    dc = torch.empty(nblk, dtype=torch.float32, device=dev)
    ac = torch.empty(n, dtype=torch.float32, device=dev)
    out = torch.empty(n, dtype=tdt, device=dev)
    qtab = torch.zeros(64, dtype=tdt, device=dev)
    qraw = torch.zeros(64, dtype=tdt, device=dev)
    info_d = torch.zeros(binding.INFO_BYTES, dtype=torch.uint8, device=dev)
    stats3 = torch.zeros(3, dtype=torch.float64, device=dev)
    slabbed = args.workload == "c5-slab" and world > 1
    stats_all = torch.zeros(3 * world, dtype=torch.float64, device=dev) if slabbed else stats3
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if n * es < (200 << 20) else None

    def read_info():
        raw = info_d.cpu().numpy().tobytes()
        return binding.GpuInfo.from_buffer_copy(raw).as_dict()

    def compress(ev=None):
        ctx.stats_dev(x.data_ptr(), n, code, stats3.data_ptr(), sh)
        if slabbed:
            dist.all_gather_into_tensor(stats_all, stats3)  # the only collective: 24 bytes per rank
        if ev:
            ev[0].record(stream)
        ctx.compress_dev(x.data_ptr(), n, n_total, code, EB, qt, stats_all.data_ptr(), world if slabbed else 1, first,
                         bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), qraw.data_ptr(), info_d.data_ptr(), sh)
        if ev:
            ev[1].record(stream)
        if qt:
            ctx.qt_finish_dev(code, EB, qraw.data_ptr(), qtab.data_ptr(), ac.data_ptr(), info_d.data_ptr(), sh)

    def decompress(sf):
        ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), qtab.data_ptr(), n, code, EB, sf, qt, out.data_ptr(), sh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    compress()
    torch.cuda.synchronize()
    info = read_info()
    if info["status"] != 0:
        raise SystemExit(f"bench.py: compress failed with status {info['status']}")
    sf = info["sf"]
    p_out = info["n_outliers"] / n

    for _ in range(args.warmup):
        if flush is not None:
            flush.zero_()
        compress()
        decompress(sf)
    barrier()

    E = lambda: torch.cuda.Event(enable_timing=True)
    evs = [[E() for _ in range(5)] for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    sampler.mark_begin()
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()
        e = evs[k]
        e[0].record(stream)
        compress(ev=(e[1], e[2]))
        e[3].record(stream)
        decompress(sf)
        e[4].record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    t_c = sum(e[0].elapsed_time(e[3]) for e in evs) / 1e3
    t_d = sum(e[3].elapsed_time(e[4]) for e in evs) / 1e3
    t_k2 = sum(e[1].elapsed_time(e[2]) for e in evs) / 1e3      # k_finalize (1 thread) + k_compress
    t_k1 = sum(e[0].elapsed_time(e[1]) for e in evs) / 1e3      # k_stats (+ all-gather)
    times = torch.tensor([t_c + t_d, t_c, t_d, t_k2, t_k1], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_rt, t_c, t_d, t_k2, t_k1 = [float(v) for v in times.cpu()]
    max_err = float((out - x).abs().max().item())
    gb_all = world * n * es * args.steps / 1e9

    # ---- end to end through the host-buffer C-ABI (pinned host buffers, copies inside the timing) ----
    e2e = None
    if not args.no_e2e:
        ne = min(n, 1 << args.e2e_log2)
        nblk_e = (ne + 63) // 64
        np_dt = np.float64 if es == 8 else np.float32
        hx = binding.PinnedArray((ne,), np_dt)
        hout = binding.PinnedArray((ne,), np_dt)
        hb = binding.PinnedArray((ne,), np.uint8)
        hdc = binding.PinnedArray((nblk_e,), np.float32)
        hac = binding.PinnedArray((ne,), np.float32)
        hx.array[:] = x[:ne].cpu().numpy()
        pre = dict(bin_index=hb.array, dc=hdc.array, ac_full=hac.array)
        ksteps = max(1, min(args.steps, 5))

        def e2e_step():
            g = ctx.compress_core(hx.array, EB, qt=qt, out=pre)
            ctx.decompress_core(g["bin_index"], g["dc"], g["ac"], ne, np_dt, EB, g["sf"], qt=qt, qtable=g.get("qtable"), out=hout.array)
            return g

        g = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            g = e2e_step()
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        n_out_e = int(g["info"]["n_outliers"])
        side = ne + 4 * nblk_e + 4 * n_out_e
        e2e = dict(value=world * ne * es * ksteps / 1e9 / float(te.item()), unit=UNIT,
                   h2d_bytes_per_step=ne * es + side + (64 * es if qt else 0),
                   d2h_bytes_per_step=side + binding.INFO_BYTES + ne * es + (128 * es if qt else 0),
                   steps=ksteps, sample=f"first {ne} elements of the rank's slab per step, pinned host buffers, "
                                        f"dctz_gpu_compress_core + dctz_gpu_decompress_core (synchronous, copies included)",
                   max_abs_err=float(np.max(np.abs(hout.array - hx.array))))
        for h in (hx, hout, hb, hdc, hac):
            h.free()

    # (every rank takes part: barrier + max over ranks -- hence before the other ranks leave)
    # Reported beside the headline, never as it: compress with KNOWN statistics (those of the previous step, as a
    # time-stepping simulation would have them) -- dctz_gpu_compress_known_stats_dev reads the input once and verifies
    # the scaling factor on the fly.
    known_leg = None
    if args.workload == "c5-slab" and not qt:
        evk = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(2):
            ctx.compress_known_stats_dev(x.data_ptr(), n, n_total, code, EB, False, stats_all.data_ptr(), world if slabbed else 1, first,
                                         bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), qraw.data_ptr(), info_d.data_ptr(), sh)
        barrier()
        evk[0].record(stream)
        for _ in range(5):
            ctx.compress_known_stats_dev(x.data_ptr(), n, n_total, code, EB, False, stats_all.data_ptr(), world if slabbed else 1, first,
                                         bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), qraw.data_ptr(), info_d.data_ptr(), sh)
        evk[1].record(stream)
        barrier()
        tk = torch.tensor([evk[0].elapsed_time(evk[1]) / 5e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        ik = read_info()
        known_leg = dict(compress_ms=1e3 * float(tk.item()), compress_gbs=world * n * es / 1e9 / float(tk.item()), verified=(ik["status"] == 0),
                         note="statistics supplied by the caller (previous step) and verified in the kernel: one read of the input")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_compress) --------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    bpe_k2 = es + 1 + 4 / 64 + 4 * p_out                 # transform read + bin index + DC + outliers
    bpe_c = 2 * es + 1 + 4 / 64 + 4 * p_out              # + statistics read (SURVEY.md §8d B_c)
    bpe_d = es + 1 + 4 / 64 + 4 * p_out                  # SURVEY.md §8d B_d
    ach = bpe_k2 * n * args.steps / t_k2 / 1e9
    traffic = None  # dram__bytes_read+write of k_compress per launch, from the committed ncu capture, scaled to this slab
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get("k_compress")
        if tr and es == 8 and not qt:
            traffic = tr["bytes_per_element"] * n
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel="k_compress<%s,%s>" % ("double" if es == 8 else "float", "QT" if qt else "EC"),
                    achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                    peak_source="measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback",
                    bytes_per_element=bpe_k2, elements_per_launch=n, outlier_fraction=p_out,
                    phases=dict(
                        compress=dict(achieved=bpe_c * n * args.steps / t_c / 1e9, frac=bpe_c * n * args.steps / t_c / 1e9 / peak, bytes_per_element=bpe_c),
                        stats=dict(achieved=es * n * args.steps / t_k1 / 1e9, frac=es * n * args.steps / t_k1 / 1e9 / peak, bytes_per_element=es),
                        decompress=dict(achieved=bpe_d * n * args.steps / t_d / 1e9, frac=bpe_d * n * args.steps / t_d / 1e9 / peak, bytes_per_element=bpe_d)))

    # ---- CPU baseline + quality on a bounded sample (rank 0, N = 1 only) -------------------------
    cpu = None
    quality = dict(max_abs_err=max_err, outlier_fraction=p_out, sf=sf, n_edge=info["n_edge"], n_exact_path=info["n_exact_path"])
    if world == 1 and not args.no_cpu:
        from tests import parity, reflib

        ns = min(n, 1 << args.cpu_log2)
        sample = x[:ns].cpu().numpy()
        procs = max(1, min(os.cpu_count() or 1, 32))
        cpu = cpu_reference(sample, qt, procs)
        # quality vs the oracle on the same sample: ties, reconstruction difference, ratio with host zlib
        g = ctx.compress_core(sample, EB, qt=qt)
        o = reflib.oracle_compress(sample, EB, qt)
        rep = parity.compare_compress(g, o, sample, EB, qt)
        r_gpu = ctx.decompress_core(g["bin_index"], g["dc"], g["ac"], ns, sample.dtype, EB, g["sf"], qt=qt, qtable=g.get("qtable"))
        r_ref = reflib.oracle_decompress(o["bin_index"], o["dc"], o["ac"], o["qtable"], ns, EB, o["stat"]["sf"], qt, sample.dtype)
        zsz = 56 + sum(len(zlib.compress(a.tobytes(), -1)) for a in (g["bin_index"], g["dc"], g["ac"])) + (64 * es if qt else 0)
        quality.update(sample_elements=ns, ties=rep["ties"], bin_mismatch=rep["bin_mismatch"],
                       max_abs_err_vs_ref=float(np.max(np.abs(r_gpu.astype(np.float64) - r_ref.astype(np.float64)))),
                       max_abs_err_sample=float(np.max(np.abs(r_gpu.astype(np.float64) - sample.astype(np.float64)))),
                       max_abs_err_ref_sample=float(np.max(np.abs(r_ref.astype(np.float64) - sample.astype(np.float64)))),
                       ratio=ns * es / zsz)

    outlier_leg = None
    if world == 1 and args.workload == "c5-slab" and not args.f32 and not qt and args.noise == 0 and not args.no_outlier_leg:
        # The headline field is smooth (no AC coefficient leaves the bin range).  Reported beside it, never as the
        # headline: the same slab shape with white noise added so that ~5 % of the coefficients are outliers.
        n2 = min(n, 1 << 28)
        gen = torch.Generator(device=dev).manual_seed(SEED)
        x2 = x[:n2] + 1.3 * torch.randn(n2, generator=gen, device=dev, dtype=torch.float64)
        st2 = torch.zeros(3, dtype=torch.float64, device=dev)

        def noisy_leg(mode_qt):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            tc = td = 0.0

            def comp():
                ctx.stats_dev(x2.data_ptr(), n2, code, st2.data_ptr(), sh)
                ctx.compress_dev(x2.data_ptr(), n2, n2, code, EB, mode_qt, st2.data_ptr(), 1, True, bins.data_ptr(), dc.data_ptr(),
                                 ac.data_ptr(), qraw.data_ptr(), info_d.data_ptr(), sh)
                if mode_qt:
                    ctx.qt_finish_dev(code, EB, qraw.data_ptr(), qtab.data_ptr(), ac.data_ptr(), info_d.data_ptr(), sh)

            comp()
            torch.cuda.synchronize()
            sf2 = read_info()["sf"]
            for it in range(3 + 5):
                ev[0].record(stream)
                comp()
                ev[1].record(stream)
                ctx.decompress_dev(bins.data_ptr(), dc.data_ptr(), ac.data_ptr(), qtab.data_ptr() if mode_qt else 0, n2, code, EB, sf2, mode_qt,
                                   out.data_ptr(), sh)
                ev[2].record(stream)
                torch.cuda.synchronize()
                if it >= 3:
                    tc += ev[0].elapsed_time(ev[1]) / 1e3
                    td += ev[1].elapsed_time(ev[2]) / 1e3
            p2 = read_info()["n_outliers"] / n2
            return dict(workload=f"first 2^{n2.bit_length() - 1} elements of the slab + Gaussian noise (std 1.3), {'QT' if mode_qt else 'EC'} mode",
                        outlier_fraction=p2, value=n2 * es * 5 / 1e9 / (tc + td), compress_gbs=n2 * es * 5 / 1e9 / tc,
                        decompress_gbs=n2 * es * 5 / 1e9 / td, compress_frac=(2 * es + 1 + 4 / 64 + 4 * p2) * n2 * 5 / tc / 1e9 / peak,
                        decompress_frac=(es + 1 + 4 / 64 + 4 * p2) * n2 * 5 / td / 1e9 / peak,
                        max_abs_err=float((out[:n2] - x2).abs().max().item()))

        outlier_leg = noisy_leg(False)
        outlier_leg["qt_mode"] = noisy_leg(True)  # the quantiser mode on the same data (not strictly error bounded by design)
        del x2

    line = dict(metric=METRIC, value=gb_all / t_rt, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t_rt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64" if es == 8 else "f32", data="synthetic", config=workload_config(args, world),
                compress_gbs=gb_all / t_c, decompress_gbs=gb_all / t_d, ms_compress=1e3 * t_c / args.steps,
                ms_decompress=1e3 * t_d / args.steps, roofline=roofline, cpu_baseline=cpu, e2e=e2e, quality=quality, outlier_leg=outlier_leg, known_stats_leg=known_leg,
                gpu_launches=int(launches), clocks=clocks, impl="ours")
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line):
    """The contract is ONE JSON line on stdout: everything else any library prints goes to stderr (see __main__)."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    a = parse_args()
    # keep stdout clean: NCCL, torchrun and friends print banners on fd 1; route fd 1 to stderr and keep a private
    # handle on the real stdout for the JSON line
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    if a.watchdog > 0:  # a hung collective or kernel must not sit on the GPU box until somebody else's timeout fires
        import faulthandler

        faulthandler.dump_traceback_later(a.watchdog, exit=True)
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
