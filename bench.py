#!/usr/bin/env python
"""bench.py — filtered Gbases/s of the per-read filter/trim hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

One "step" = one pass of the whole hot path (K1 raw scan -> K3 adapter search -> K5 regions ->
[K4] -> K1 clean scan) over one batch of synthetic reads.  Workload at N=1: BASELINE config[1]
(synthetic ONT, 200 000 reads, N50 ~30 kb, planted 5' adapter, `-x ont` with adapter
auto-identify + end trim); every rank of a multi-GPU run processes its own shard of that size
(weak scaling, no data-path collective; the QC counters are combined with one NCCL allreduce
after the timed region).

`value`  : input bases / device time, inputs resident in HBM (CUDA events on the library's stream)
`e2e`    : same metric through the C-ABI with pinned HOST buffers, H2D + kernels + D2H of the results
           inside the timed region, in the input format the C++ host (src/TGSFilter.cpp) feeds:
           2-bit packed bases + Phred bytes (tgsf_submit_packed); `e2e_bytes` is the same through
           tgsf_submit with one byte per base.  Both are PCIe-bound.
`roofline`: dominant kernel (K3 k_mid_scan, INT-ALU bound per SURVEY.md §8(d)); the HBM-bound K1
           scan is reported next to it under `roofline_kernels`
`cpu_baseline`: the UNMODIFIED reference CLI (oracle/_ref/tgsfilter) on a bounded sample of the
           same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIG_READS = {1: 20_000, 2: 200_000, 3: 50_000, 4: 1_000_000, 5: 2_000_000}
CONFIG_CLI = {1: ["-x", "hifi"], 2: ["-x", "ont"], 3: ["-x", "ont", "-M", "35", "-T", "50"],
              4: ["-x", "clr", "-q", "7", "-Q", "15", "-e", "150", "-b", "1"],
              5: ["-x", "hifi", "-k", "11", "-p", "5000"]}
CONFIG_TYPE = {1: "hifi", 2: "ont", 3: "ont", 4: "clr", 5: "hifi"}
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "tgsfilter")


# ------------------------------------------------------------------------------------------------
# synthetic workload, generated on the GPU (same distributions as tgsfilter_b200.synth)
# ------------------------------------------------------------------------------------------------
def gen_workload_gpu(config: int, n_reads: int, seed: int, device):
    import torch
    from tgsfilter_b200 import synth
    from tgsfilter_b200.params import ADAPTER_LIB

    rng = np.random.default_rng(seed)
    lens = synth._lengths(config, n_reads, rng)
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    bases = torch.empty(total + 64, dtype=torch.uint8, device=device)
    quals = torch.empty(total + 64, dtype=torch.uint8, device=device)
    bases[total:] = 65
    quals[total:] = 33
    mu, sigma = synth._QUAL[config]
    mean_q = rng.normal(mu, sigma, n_reads).astype(np.float32)
    if config in (1, 5):
        low = rng.random(n_reads) < 0.05
        mean_q[low] = rng.normal(15, 4, int(low.sum())).astype(np.float32)
    d_mean = torch.from_numpy(mean_q).to(device)
    d_lens = torch.from_numpy(lens).to(device)
    group = max(1, n_reads // 32)
    for lo in range(0, n_reads, group):
        hi = min(n_reads, lo + group)
        s, e = int(offsets[lo]), int(offsets[hi])
        code = torch.randint(0, 4, (e - s,), dtype=torch.uint8, device=device, generator=g)
        # A=65 C=67 G=71 T=84
        b = 65 + 2 * code + 2 * (code >= 2).to(torch.uint8) + 11 * (code == 3).to(torch.uint8)
        bases[s:e] = b
        q = torch.repeat_interleave(d_mean[lo:hi], d_lens[lo:hi])
        q = q + 5.0 * torch.randn(e - s, dtype=torch.float32, device=device, generator=g)
        quals[s:e] = (q.round_().clamp_(1, 60) + 33).to(torch.uint8)
        del code, b, q
    d_off = torch.from_numpy(offsets).to(device)

    def plant(reads_idx, pos, variants, vlen, which):
        """bases[offsets[r] + pos + j] = variants[which][j] for j < vlen[which]"""
        if len(reads_idx) == 0:
            return
        r = torch.from_numpy(reads_idx).to(device)
        p = torch.from_numpy(pos).to(device)
        w = torch.from_numpy(which).to(device)
        V = torch.from_numpy(variants).to(device)
        VL = torch.from_numpy(vlen).to(device)
        J = torch.arange(V.shape[1], device=device)[None, :]
        mask = J < VL[w][:, None]
        dest = (d_off[r] + p)[:, None] + J
        bases[dest[mask]] = V[w][mask]

    def variant_pool(ad, err, count):
        vs = [synth.mutate(ad, err, rng) for _ in range(count)]
        ml = max(len(v) for v in vs)
        arr = np.full((count, ml), 65, dtype=np.uint8)
        for i, v in enumerate(vs):
            arr[i, :len(v)] = np.frombuffer(v, dtype=np.uint8)
        return arr, np.array([len(v) for v in vs], dtype=np.int64)

    if config in (2, 3):
        ad = ADAPTER_LIB[8]
        V, VL = variant_pool(ad, 0.10, 1024)
        sel = np.nonzero((rng.random(n_reads) < 0.80) & (lens > 200))[0]
        plant(sel, rng.integers(0, 31, len(sel)), V, VL, rng.integers(0, 1024, len(sel)))
        frac_mid, err_mid = (0.01, 0.10) if config == 2 else (0.10, 0.05)
        V2, VL2 = variant_pool(ad, err_mid, 1024)
        sel = np.nonzero((rng.random(n_reads) < frac_mid) & (lens > 1000))[0]
        pos = (400 + rng.random(len(sel)) * (lens[sel] - 900)).astype(np.int64)
        plant(sel, pos, V2, VL2, rng.integers(0, 1024, len(sel)))
    elif config in (1, 5):
        ad = ADAPTER_LIB[0]
        V, VL = variant_pool(ad, 0.03, 512)
        u = rng.random(n_reads)
        sel = np.nonzero(u < 0.015)[0]
        plant(sel, np.zeros(len(sel), np.int64), V, VL, rng.integers(0, 512, len(sel)))
        sel = np.nonzero((u >= 0.015) & (u < 0.030))[0]
        plant(sel, lens[sel] - 60, V, VL, rng.integers(0, 512, len(sel)))
        sel = np.nonzero((u >= 0.030) & (u < 0.033))[0]
        pos = (300 + rng.random(len(sel)) * (lens[sel] - 700)).astype(np.int64)
        plant(sel, pos, V, VL, rng.integers(0, 512, len(sel)))
    if config == 5:
        # 30 % of reads carry a 6-12 kb tandem repeat (unit 50-500 bp, 1 % error; substitutions only here,
        # tgsfilter_b200.synth also plants indels), so that repeatLen straddles the -p 5000 bound
        sel = np.nonzero(rng.random(n_reads) < 0.30)[0]
        unit_len = rng.integers(50, 501, len(sel))
        span = np.minimum(rng.integers(6000, 12001, len(sel)), lens[sel] - 200)
        ok = span > unit_len
        sel, unit_len, span = sel[ok], unit_len[ok], span[ok]
        pos = (100 + rng.random(len(sel)) * np.maximum(1, lens[sel] - span - 150)).astype(np.int64)
        pool = torch.randint(0, 4, (1 << 22,), dtype=torch.uint8, device=device, generator=g)
        pool = 65 + 2 * pool + 2 * (pool >= 2).to(torch.uint8) + 11 * (pool == 3).to(torch.uint8)
        unit_off = rng.integers(0, (1 << 22) - 512, len(sel))
        step = 20000
        for lo in range(0, len(sel), step):
            hi = min(len(sel), lo + step)
            sp = torch.from_numpy(span[lo:hi]).to(device)
            first = torch.cumsum(sp, 0) - sp
            rid = torch.repeat_interleave(torch.arange(hi - lo, device=device), sp)
            j = torch.arange(int(sp.sum().item()), device=device) - first[rid]
            ul = torch.from_numpy(unit_len[lo:hi]).to(device)[rid]
            uo = torch.from_numpy(unit_off[lo:hi]).to(device)[rid]
            val = pool[uo + j % ul]
            err = torch.rand(val.shape, device=device, generator=g) < 0.01
            val = torch.where(err, pool[(uo + j * 7 + 13) % (1 << 22)], val)
            dest = (d_off[torch.from_numpy(sel[lo:hi]).to(device)] + torch.from_numpy(pos[lo:hi]).to(device))[rid] + j
            bases[dest] = val
            del sp, first, rid, j, ul, uo, val, err, dest
    if config == 4:
        k = 12
        r = np.arange(n_reads)
        J = torch.arange(k, device=device)[None, :]
        at = torch.rand((n_reads, k), device=device, generator=g) < 0.70
        pick_at = torch.randint(0, 2, (n_reads, k), device=device, generator=g) * 19 + 65   # A / T
        pick_cg = torch.randint(0, 2, (n_reads, k), device=device, generator=g) * 4 + 67    # C / G
        val = torch.where(at, pick_at, pick_cg).to(torch.uint8)
        dest = d_off[:-1][:, None] + J
        mask = J < d_lens[:, None]
        bases[dest[mask]] = val[mask]
        del r
    torch.cuda.synchronize(device)
    return bases, quals, d_off, offsets, total


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def __enter__(self):
        if shutil.which("nvidia-smi") is None:
            return self
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()
        time.sleep(0.3)
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half ~ samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference on the host CPU (bounded sample)
# ------------------------------------------------------------------------------------------------
def reference_sample_fastq(config: int, sample_reads: int) -> bytes:
    from tgsfilter_b200 import synth
    batch = synth.make_config(config, sample_reads, with_names=False)
    return batch.to_fastq(), batch.n_bases


def time_reference(config: int, fq_path: str, n_bases: int, threads: int, repeats: int = 1):
    out = os.path.join(os.path.dirname(fq_path), "ref_out.fq")
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        pr = subprocess.run([REF_CLI, "-i", fq_path, "-o", out, "-t", str(threads)] + CONFIG_CLI[config],
                            stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, cwd=os.path.dirname(fq_path))
        dt = time.perf_counter() - t0
        if pr.returncode != 0:
            raise RuntimeError("reference CLI failed: " + pr.stderr.decode()[-400:])
        best = dt if best is None else min(best, dt)
    return n_bases / best / 1e9, best


def ref_threads() -> int:
    n = os.cpu_count() or 2
    return max(1, min(32, n - 1))  # the reference's own clamp, T.cpp:488-499


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = args.config
    if not os.path.exists(REF_CLI):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/tgsfilter not built"}))
        return 0
    threads = ref_threads()
    sample_reads = args.sample_reads
    tmpdir = tempfile.mkdtemp(prefix="tgsf_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fq, n_bases = reference_sample_fastq(cfg, sample_reads)
        fq_path = os.path.join(tmpdir, "sample.fq")
        with open(fq_path, "wb") as f:
            f.write(fq)
        for _ in range(args.warmup):
            time_reference(cfg, fq_path, n_bases, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            time_reference(cfg, fq_path, n_bases, threads)
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
    value = n_bases * args.steps / dt / 1e9
    sample = f"first {sample_reads} reads of the config-{cfg} generator ({n_bases} bases), FASTQ on tmpfs -> FASTQ on tmpfs"
    line = {
        "impl": "reference", "metric": "filtered Gbases/s", "value": value, "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, CONFIG_READS[cfg]), "cli": " ".join(CONFIG_CLI[cfg]),
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_name(cfg: int, n_reads: int) -> str:
    names = {1: "config[0] synthetic HiFi FASTQ ~15 kb, -x hifi",
             2: "config[1] synthetic ONT FASTQ N50 ~30 kb, planted 5' adapter, -x ont (auto-identify + end trim)",
             3: "config[2] synthetic ONT ultra-long N50 ~100 kb, planted middle adapters, -M 35 -T 50",
             4: "config[3] synthetic PacBio CLR, -q 7 -Q 15 -e 150 -b 1",
             5: "config[4] synthetic HiFi, -k 11 -p 5000"}
    return f"{names[cfg]}; {n_reads} reads per GPU"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE config index + 1 (2 = configs[1])")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (0 = the config's full size)")
    ap.add_argument("--sample-reads", type=int, default=6000, help="reads in the CPU reference sample")
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from tgsfilter_b200 import _capi, prepass
    from tgsfilter_b200.engine import FilterEngine
    from tgsfilter_b200.params import FilterParams
    from tgsfilter_b200.synth import ReadBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tgsfilter_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    cfg = args.config
    n_reads = args.reads or CONFIG_READS[cfg]
    read_type = CONFIG_TYPE[cfg]
    d_bases, d_quals, d_off, offsets, n_bases = gen_workload_gpu(cfg, n_reads, 20261017 + cfg + 1000 * rank, device)

    # host copies (pinned) for the end-to-end path and the pre-pass sampling
    h_bases = torch.empty(n_bases, dtype=torch.uint8, pin_memory=True)
    h_quals = torch.empty(n_bases, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(d_bases[:n_bases])
    h_quals.copy_(d_quals[:n_bases])
    torch.cuda.synchronize()
    host_batch = ReadBatch(h_bases.numpy(), h_quals.numpy(), offsets.astype(np.uint64))

    # parameters exactly as the CLI would resolve them for this config
    cli = CONFIG_CLI[cfg]
    p = FilterParams().apply_read_type(read_type)
    if "-q" in cli:
        p.min_q = float(cli[cli.index("-q") + 1])
    if "-Q" in cli:
        p.max_q = float(cli[cli.index("-Q") + 1])
    if "-p" in cli:
        p.min_repeat = int(cli[cli.index("-p") + 1])
    t0 = time.perf_counter()
    params, pre = prepass.run_prepass(host_batch, p, read_type, device=local_rank)
    prepass_ms = (time.perf_counter() - t0) * 1e3
    params.n_slots = 2

    eng = FilterEngine(params, device=local_rank)
    off_u64 = d_off  # int64 with the same bit pattern as uint64

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        eng.submit_device(d_bases.data_ptr(), d_quals.data_ptr(), off_u64.data_ptr(), n_reads, n_bases)
        eng.collect(want_results=False)
        return eng.last_timing()[0], eng.last_stage_ms()

    # sub-batches of the end-to-end path (host buffers; two slots -> copy / compute overlap)
    bounds = np.linspace(0, n_reads, args.e2e_chunks + 1).astype(np.int64)
    sub = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b > a:
            o = (offsets[a:b + 1] - offsets[a]).astype(np.uint64)
            o_t = torch.from_numpy(o.view(np.int64)).pin_memory()
            sub.append((int(offsets[a]), int(b - a), o_t, int(offsets[b] - offsets[a])))
    d2h_bytes = 0

    # 2-bit packed copy of every sub-batch (what the C++ host's parser produces with tgsf_pack_bases;
    # the synthetic bases are pure upper-case ACGT, so the exception list is empty)
    pk_off = [0]
    for _, _, _, nb in sub:
        pk_off.append(pk_off[-1] + ((nb + 3) // 4 + 63) // 64 * 64)
    h_packed = torch.empty(pk_off[-1] + 64, dtype=torch.uint8, pin_memory=True)
    for (s0, nr, o_t, nb), po in zip(sub, pk_off[:-1]):
        b = d_bases[s0:s0 + nb]
        if nb % 4:
            b = torch.cat([b, torch.full((4 - nb % 4,), 65, dtype=torch.uint8, device=device)])
        code = (b >> 1) & 3
        code = (code ^ (code >> 1)).view(-1, 4)          # A0 C1 G2 T3
        pk = code[:, 0] | (code[:, 1] << 2) | (code[:, 2] << 4) | (code[:, 3] << 6)
        h_packed[po:po + pk.numel()].copy_(pk)
        del b, code, pk
    torch.cuda.synchronize()

    def step_e2e_packed():
        inflight = 0
        for (s0, nr, o_t, nb), po in zip(sub, pk_off[:-1]):
            if inflight == 2:
                eng.collect()
                inflight -= 1
            eng.submit_packed_raw(h_packed.data_ptr() + po, h_quals.data_ptr() + s0, o_t.data_ptr(), nr)
            inflight += 1
        while inflight:
            eng.collect()
            inflight -= 1

    def step_e2e():
        nonlocal d2h_bytes
        d2h = 0
        inflight = 0
        for s0, nr, o_t, nb in sub:
            if inflight == 2:
                r, pcs = eng.collect()
                d2h += r.nbytes + pcs.nbytes
                inflight -= 1
            eng.submit_raw(h_bases.data_ptr() + s0, h_quals.data_ptr() + s0, o_t.data_ptr(), nr)
            inflight += 1
        while inflight:
            r, pcs = eng.collect()
            d2h += r.nbytes + pcs.nbytes
            inflight -= 1
        d2h_bytes = d2h

    # ---- warm-up
    for _ in range(args.warmup):
        step_resident()
    step_e2e()
    step_e2e_packed()
    eng.reset_counters()

    # ---- timed: resident
    launches0 = eng.launch_count()
    stage_sum = {k: 0.0 for k in _capi.STAGE_NAMES}
    barrier()
    with ClockSampler(local_rank) as clk:
        wall0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(args.steps):
            k_ms, st = step_resident()
            dev_ms += k_ms
            for k in stage_sum:
                stage_sum[k] += st[k]
        barrier()
        wall_ms = (time.perf_counter() - wall0) * 1e3
    launches = eng.launch_count() - launches0
    clocks = clk.summary()
    cnt = eng.counters()
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t_dev.item())

    # ---- timed: end to end (host buffers)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - e0
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())

    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_packed()
    barrier()
    t_pk = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_pk, op=dist.ReduceOp.MAX)
    e2e_packed_s = float(t_pk.item())
    # host packer speed (one thread), for context
    nb_probe = min(n_bases, 256 << 20)
    probe_out = np.empty(nb_probe // 4 + 16, dtype=np.uint8)
    ne = __import__("ctypes").c_uint64(0)
    tp0 = time.perf_counter()
    _capi.load().tgsf_pack_bases(h_bases.data_ptr(), nb_probe, probe_out.ctypes.data, None, None, 0,
                                 __import__("ctypes").byref(ne))
    pack_gbs = nb_probe / (time.perf_counter() - tp0) / 1e9

    # ---- final counter allreduce over NVLink (outside the timed region; reported)
    allreduce_ms = None
    if world > 1:
        ptr, nwords = eng.counters_device_ptr()

        class _Blk:  # zero-copy view of the device counter block
            __cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        blk = torch.as_tensor(_Blk(), device=device)
        warm = torch.zeros(nwords, dtype=torch.int64, device=device)  # same size: NCCL sets its channels up lazily
        dist.all_reduce(warm, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        a0 = time.perf_counter()
        dist.all_reduce(blk, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        allreduce_ms = (time.perf_counter() - a0) * 1e3

    total_bases = n_bases * world
    value = total_bases * args.steps / (dev_ms_max / 1e3) / 1e9
    e2e_value = total_bases * args.steps / e2e_s / 1e9
    e2e_packed_value = total_bases * args.steps / e2e_packed_s / 1e9

    # ---- rooflines
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    lens = np.diff(offsets)
    drop = cnt.drop_info
    # reads that reach the adapter search = all minus low-quality drops; their middle windows
    lowq_reads = int(drop[0])
    E = params.end_len
    word_cols_all = 0
    for a in params.adapters:
        q = len(a)
        nw = (q + 63) // 64
        mid = lens - 2 * E
        word_cols_all += nw * int(mid[mid >= q].sum())
    active_frac = 1.0 - (int(drop[1]) / args.steps) / max(1, n_bases)
    word_cols = word_cols_all * active_frac
    mid_ms = stage_sum["mid_scan"] / args.steps
    sm_clock = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    int_peak = 148 * 128 * sm_clock * 1e6 / 1e12  # T int32-op/s at the clock seen under load
    achieved_int = word_cols * 34 / (mid_ms / 1e3) / 1e12 if mid_ms > 0 else 0.0
    raw_ms = stage_sum["raw_scan"] / args.steps
    clean_ms = stage_sum["clean"] / args.steps
    k1_gbs = 2 * n_bases / (raw_ms / 1e3) / 1e9 if raw_ms > 0 else 0.0
    step_ms = dev_ms / args.steps if dev_ms else 0.0
    kept_bases = n_bases - int(drop[1]) // args.steps  # bases entering the clean pass (upper bound)
    clean_gbs = 2 * kept_bases / (clean_ms / 1e3) / 1e9 if clean_ms > 0 else 0.0
    kmer_ms = stage_sum["kmer"] / args.steps
    rl_mid = {"kernel": "k_mid_scan (K3 Myers HW scan, middle windows)", "bound": "int_alu",
              "achieved": achieved_int, "peak": int_peak, "unit": "Tint32op/s",
              "frac": achieved_int / int_peak if int_peak else None, "traffic": None,
              "work": "34 int32-op equivalents per 64-bit word-column (SURVEY.md §8d), "
                      f"{word_cols:.4g} word-columns per launch set",
              "peak_def": f"148 SMs x 128 lanes x {sm_clock:.0f} MHz (median SM clock under load)",
              "note": "ALU-pipe bound (ncu: pipe_alu ~92 % busy); logic ops cannot use the FMA pipe, so "
                      "the 128-lane peak is not reachable: see DESIGN.md §4",
              "ms": mid_ms, "share_of_step": mid_ms / step_ms if step_ms else None}
    rl_raw = {"kernel": "k_scan_tiles_dyn raw pass (K1)", "bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak,
              "unit": "GB/s", "frac": k1_gbs / hbm_peak, "traffic": None, "peak_source": hbm_src,
              "work": "2 B per input base", "ms": raw_ms, "share_of_step": raw_ms / step_ms if step_ms else None}
    rl_clean = {"kernel": "k_scan_tiles_dyn clean pass (K1)", "bound": "hbm", "achieved": clean_gbs,
                "peak": hbm_peak, "unit": "GB/s", "frac": clean_gbs / hbm_peak, "traffic": None,
                "peak_source": hbm_src, "work": "2 B per kept base (upper bound: bases of reads passing -q/-Q)",
                "ms": clean_ms, "share_of_step": clean_ms / step_ms if step_ms else None}
    rl_kmer = {"kernel": "k_kmer_smem (K4, k <= 12: atomics-free tag rounds over a 2-bit staged piece in shared "
                         "memory; pieces > 196 kb: shared-memory bitmap passes; k = 13: L2 bitmap; k > 13: hash)",
               "bound": "issue+barrier (one CTA per piece: ~43 k warp-instructions and ~14 barriers per 15 kb piece; "
                        "ncu: issue slots 52 % busy, 29 % of stall samples on barriers)",
               "achieved": kept_bases / (kmer_ms / 1e3) / 1e9 if kmer_ms > 0 else 0.0, "peak": None,
               "unit": "Ginserts/s", "frac": None, "traffic": None, "work": "1 insert per kept base",
               "ms": kmer_ms, "share_of_step": kmer_ms / step_ms if step_ms else None}
    # measured DRAM traffic per launch (one ncu --set full capture of this same command, committed
    # under profiles/); only valid for the workload it was captured on
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tr = json.load(f)
        if tr.get("config") == cfg and tr.get("reads_per_gpu") == n_reads:
            rl_mid["traffic"] = tr["k_mid_scan"]["dram_bytes"]
            mid_all = lens - 2 * E
            qmin = min(len(a) for a in params.adapters)
            # 1 B per middle-window column of the reads that pass -q/-Q; one pass serves both adapters
            rl_mid["algorithmic_bytes"] = int(int(mid_all[mid_all >= qmin].sum()) * active_frac)
            rl_raw["traffic"] = tr["k_scan_tiles_raw"]["dram_bytes"]
            rl_raw["algorithmic_bytes"] = 2 * n_bases
            rl_clean["traffic"] = tr["k_scan_tiles_clean"]["dram_bytes"]
    except Exception:
        pass
    by_stage = {"mid_scan": rl_mid, "raw_scan": rl_raw, "clean": rl_clean, "kmer": rl_kmer}
    dominant = max(by_stage, key=lambda k: by_stage[k]["ms"])
    roofline = by_stage[dominant]
    roofline_kernels = {k: v for k, v in by_stage.items() if k != dominant and v["ms"] > 0}
    roofline_kernels["stage_ms_per_step"] = {k: v / args.steps for k, v in stage_sum.items()}

    line = {
        "metric": "filtered Gbases/s", "value": value, "unit": "Gbases/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64",
        "data": "synthetic",
        "config": {"workload": workload_name(cfg, n_reads), "cli": " ".join(cli),
                   "reads_per_gpu": n_reads, "bases_per_gpu": n_bases,
                   "adapters": [a.decode() for a in params.adapters],
                   "head_trim": params.head_trim, "tail_trim": params.tail_trim,
                   "l2": "inputs (2 B/base, >= 0.5 GB per launch) are larger than the 126 MB L2",
                   "timing": "CUDA events on the library stream around the K1..K5 sequence, max over ranks"},
        "e2e_bytes": {"value": e2e_value, "unit": "Gbases/s", "h2d_bytes_per_step": int(2 * n_bases + 8 * (n_reads + len(sub))) * world,
                "d2h_bytes_per_step": int(d2h_bytes) * world, "chunks": len(sub), "slots": 2,
                "input_format": "byte bases + Phred bytes + offsets in pinned host memory (tgsf_submit)"},
        "e2e": {"value": e2e_packed_value, "unit": "Gbases/s",
                       "h2d_bytes_per_step": int(pk_off[-1] + n_bases + 8 * (n_reads + len(sub))) * world,
                       "d2h_bytes_per_step": int(d2h_bytes) * world, "chunks": len(sub), "slots": 2,
                       "input_format": "2-bit packed bases + Phred bytes + offsets in pinned host memory "
                                       "(tgsf_submit_packed, the path src/TGSFilter.cpp uses); packing is host "
                                       "parser work outside the timed region",
                       "host_pack_gbases_per_s_per_thread": pack_gbs},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_kernels": roofline_kernels,
        "wall_ms_per_step_resident": wall_ms / args.steps,
        "prepass_ms": prepass_ms,
        "allreduce_ms": allreduce_ms,
        "drop_info_per_step": [int(x) // args.steps for x in drop],
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_CLI):
        tmpdir = tempfile.mkdtemp(prefix="tgsf_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            fq, nb = reference_sample_fastq(cfg, args.sample_reads)
            fq_path = os.path.join(tmpdir, "sample.fq")
            with open(fq_path, "wb") as f:
                f.write(fq)
            threads = ref_threads()
            v, secs = time_reference(cfg, fq_path, nb, threads)
            line["cpu_baseline"] = {"value": v, "unit": "Gbases/s", "cores": threads, "kind": "reference",
                                    "sample": f"first {args.sample_reads} reads of the config-{cfg} generator "
                                              f"({nb} bases) through the unmodified reference CLI, -t {threads}, "
                                              f"FASTQ on tmpfs, {secs:.1f} s"}
        except Exception as exc:  # keep the GPU line even if the CPU leg fails
            line["cpu_baseline"] = {"value": None, "unit": "Gbases/s", "cores": 0, "kind": "reference",
                                    "sample": f"failed: {exc}"}
        finally:
            shutil.rmtree(tmpdir, ignore_errors=True)
    eng.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
