#!/usr/bin/env python
"""bench.py — the BASELINE.json metric: queries/sec (+ decoded-postings/sec) on the synthetic Zipfian index.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload and2|or10|tree8]

One "step" = one pass of the hot path over one batch of synthetic queries.  Default workload (configs[1]):
a batch of 1000 2-term AND queries on the 100M-doc Zipfian synthetic index, GOOGLE codec, DocumentsOnly.

  value   whole-job queries/s with the index resident in HBM, device timed with CUDA events, max over ranks
  e2e     same metric through the public C-ABI call with HOST buffers: plans H2D + every matched docID (or top-k) D2H
  roofline   of the fused k_exec_tiles kernel: algorithmic bytes (sum of the queries' term chunks + emitted bytes) / event time
  cpu_baseline   the reference's own exec_query (oracle/_ref) on the host cores, bounded sample of the same batch

N > 1: the docID space is partitioned across ranks (strong scaling: the same 100M-doc index, each rank holds the postings of
its docID range).  Docs-only results need no exchange (shard order == docID order); the top-k workload merges
per-shard top-k with ONE all-gather (NCCL) + a merge kernel.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "and2": dict(codec=0, mode=0, desc="1000 x 2-term AND, DocumentsOnly, GOOGLE codec"),
    "or10": dict(codec=1, mode=2, desc="10-term OR, BM25 top-100, LUCENE codec"),
    "tree8": dict(codec=0, mode=0, desc="8-term mixed AND/OR/NOT trees, DocumentsOnly, GOOGLE codec"),
}


def gen_queries(workload: str, nq: int, nterms: int, seed: int = 0xC0FFEE):
    """term ranks drawn ~ 1/r (query-log-like), BASELINE.md section 3"""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, nterms + 1)
    w /= w.sum()
    names = [f"t{r:04d}" for r in range(1, nterms + 1)]
    out, ranks = [], []
    for _ in range(nq):
        if workload == "and2":
            t = rng.choice(nterms, size=2, replace=False, p=w)
            out.append(f"{names[t[0]]} AND {names[t[1]]}")
        elif workload == "or10":
            t = rng.choice(nterms, size=10, replace=False, p=w)
            out.append(" OR ".join(names[i] for i in t))
        else:
            t = rng.choice(nterms, size=8, replace=False, p=w)
            n = [names[i] for i in t]
            k = int(rng.integers(0, 4))
            out.append([
                f"({n[0]} OR {n[1]}) AND ({n[2]} OR {n[3]}) AND {n[4]} NOT ({n[5]} OR {n[6]} OR {n[7]})",
                f"{n[0]} AND {n[1]} AND {n[2]} NOT {n[3]} NOT {n[4]}",
                f"({n[0]} AND {n[1]}) OR ({n[2]} AND {n[3]}) OR ({n[4]} AND {n[5]}) NOT {n[6]} NOT {n[7]}",
                f"{n[0]} AND ({n[1]} OR {n[2]} OR {n[3]}) NOT ({n[4]} AND {n[5]}) AND ({n[6]} OR {n[7]})",
            ][k])
        ranks.append(t)
    return out, ranks


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


from trinity_b200.sharded import shard_range  # noqa: E402  (host-side sharding plumbing)


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture
    of this same command (profiles/traffic.json, written from the .ncu-rep by scripts/ncu_summary.py); None if not captured"""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="and2", choices=list(WORKLOADS))
    ap.add_argument("--ndocs", type=int, default=100_000_000)
    ap.add_argument("--nterms", type=int, default=4096)
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    codec, mode = wl["codec"], wl["mode"]
    W = max(3, args.warmup) if args.impl == "ours" else args.warmup
    K = max(1, args.steps)

    if args.impl == "reference":
        return reference_arm(args, rank, world, wl, K, W)

    import torch
    import torch.distributed as dist

    import trinity_b200 as tb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — trinity_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    t0 = time.time()
    lo, hi = shard_range(args.ndocs, rank, world)
    synth = tb.SynthIndex(codec, args.ndocs, args.nterms, threads=max(1, (os.cpu_count() or 8) // world), doc_range=(lo, hi))
    build_s = time.time() - t0
    g = tb.GpuIndexSource(local_rank)
    stream = torch.cuda.current_stream()
    g.set_stream(stream.cuda_stream)
    t0 = time.time()
    g.upload(codec, np.asarray(synth.index), np.asarray(synth.terms), args.ndocs)
    upload_s = time.time() - t0
    info = g.info()

    texts, _ = gen_queries(args.workload, args.nq, args.nterms)
    tdict = tb.TermDictionary(synth.names)
    plans = [tb.parse_query(q, tdict) for q in texts]
    # global BM25 weights: df summed over shards == the unsharded df (similarity.h:209-217); identical on every rank
    if mode != tb.MODE_DOCS_ONLY:
        full_df = np.array([max(1000, args.ndocs // (2 * r)) for r in range(1, args.nterms + 1)], dtype=np.int64)
        for p in plans:
            for x in p:
                if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
                    x["weight"] = tb.bm25_idf(int(full_df[x["term"]]), args.ndocs)
    # full-scan accounting numerator (same for CPU and GPU): sum of term.documents over the UNSHARDED index
    full_df_all = np.array([max(1000, args.ndocs // (2 * r)) for r in range(1, args.nterms + 1)], dtype=np.int64)
    postings_per_batch = int(sum(int(full_df_all[x["term"]]) for p in plans for x in p if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    gathered = merged_d = merged_s = None
    if world > 1 and mode == tb.MODE_SCORED_TOPK:
        gathered_d = torch.empty((world, args.nq, args.k), dtype=torch.int32, device="cuda")
        gathered_s = torch.empty((world, args.nq, args.k), dtype=torch.float32, device="cuda")
        merged_d = torch.empty((args.nq, args.k), dtype=torch.int32, device="cuda")
        merged_s = torch.empty((args.nq, args.k), dtype=torch.float32, device="cuda")
        gathered = (gathered_d, gathered_s)

    def exchange():
        """the ONE exchange step of the sharded path: all-gather of per-shard top-k + merge kernel"""
        if gathered is None:
            return
        dptr, sptr, _ = g.last_topk_device()
        # wrap the engine's device buffers (no copy, no host round trip)
        src_d = _as_tensor(dptr, args.nq * args.k, torch.int32)
        src_s = _as_tensor(sptr, args.nq * args.k, torch.float32)
        dist.all_gather_into_tensor(gathered[0].view(-1), src_d)
        dist.all_gather_into_tensor(gathered[1].view(-1), src_s)
        g.merge_topk(gathered[0].data_ptr(), gathered[1].data_ptr(), world, args.nq, args.k, merged_d.data_ptr(), merged_s.data_ptr())

    def _as_tensor(ptr, n, dtype):
        class _Holder:
            pass
        itemsize = 4
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4" if dtype == torch.int32 else "<f4", "data": (ptr, False), "version": 3,
                                      "strides": None}
        return torch.as_tensor(h, device="cuda")

    # ---------------- warm-up (also sizes every grow-only buffer) ----------------
    packed = g.pack(plans)
    for _ in range(W):
        res = g.exec_batch(plans, mode, args.k, copy=False, packed=packed)
        exchange()
    for _ in range(W):  # the device-resident form has its own (whole-batch) buffers: warm those too
        g.exec_batch_device(plans, mode, args.k, packed=packed)
        exchange()
    matches_per_batch = int(res.match_counts.sum())
    out_bytes_per_batch = (matches_per_batch * 4) if mode == tb.MODE_DOCS_ONLY else args.nq * args.k * 8

    # ---------------- timed: device-resident ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, launches = [], 0
    barrier()
    ev0.record(stream)
    for _ in range(K):
        g.exec_batch_device(plans, mode, args.k, packed=packed)
        exchange()
    ev1.record(stream)
    barrier()
    dev_ms_total = ev0.elapsed_time(ev1)

    # ---------------- timed: end to end through the C ABI with host buffers ----------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        res = g.exec_batch(plans, mode, args.k, copy=False, packed=packed)
        exchange()
        kern_ms.append(res.exec_kernel_ms)
        launches += res.kernel_launches + (1 if gathered is not None else 0)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    times = torch.tensor([dev_ms_total / 1e3, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = float(times[0]), float(times[1])

    plan_bytes = int(sum(p.nbytes for p in plans))
    d2h = out_bytes_per_batch + (args.nq + 1) * 8 + args.nq * 8
    peak, peak_src = measured_peak()
    kernel_name = "k_exec_docs" if mode == tb.MODE_DOCS_ONLY else "k_exec_tiles"
    # the host-buffer path pipelines a set-query batch in chunks (TRN_PIPELINE_CHUNKS, default 4): one fused-kernel launch per chunk
    nlaunch = 1 if mode == tb.MODE_SCORED_TOPK else min(int(os.environ.get("TRN_PIPELINE_CHUNKS", "4")), max(1, args.nq // 8))
    k_ms_step = float(np.mean(kern_ms)) if kern_ms else None          # all fused-kernel launches of one step
    k_ms = k_ms_step / nlaunch if k_ms_step else None                 # average duration of ONE launch
    traffic = ncu_traffic(f"{kernel_name}:{args.workload}") if (world == 1 and args.ndocs == 100_000_000 and args.nq == 1000) else None
    algo_bytes = (int(res.index_bytes_touched) + out_bytes_per_batch) // nlaunch   # algorithmic bytes of ONE launch
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None

    line = {
        "metric": "queries/sec (batched 2-term AND, 100M-doc Zipfian synthetic index)" if args.workload == "and2" else f"queries/sec ({args.workload})",
        "value": args.nq * K / dev_s,
        "unit": "queries/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": dev_s * 1e3 / K,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u32",
        "data": "synthetic",
        "impl": "ours",
        "config": {
            "workload": f"{wl['desc']}; {args.ndocs} docs, {args.nterms} terms, Zipf(1) df, batch {args.nq}, seed 0xC0FFEE",
            "codec": "GOOGLE" if codec == 0 else "LUCENE",
            "docid_sharding": f"{world} x contiguous docID range" if world > 1 else "none",
            "l2": f"inputs larger than L2: {info['index_bytes'] / 1e6:.0f} MB index + {info['directory_bytes'] / 1e6:.0f} MB directory per GPU vs 126 MB L2",
            "index_build_s": round(build_s, 1), "upload_s": round(upload_s, 1),
        },
        "decoded_postings_per_s": postings_per_batch * K / dev_s,
        "matches_per_batch_rank0": matches_per_batch,
        "e2e": {"value": args.nq * K / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": plan_bytes, "d2h_bytes_per_step": d2h,
                "decoded_postings_per_s": postings_per_batch * K / e2e_s},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_ms, "launches_per_step": nlaunch,
                     "note": "achieved = SURVEY 8(d) algorithmic bytes (every referenced list once per query, skipped blocks included) / kernel time; instruction-issue and latency bound, not HBM bound; see DESIGN.md section 4"},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(synth, texts, args, mode, check=res)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def full_size_parity(res, mode, k, counts, sums, tid, tsc, n):
    """the reference run of the cpu_baseline leg doubles as the full-size checker of the GPU batch `res` (same index bytes, same
    queries): per-query match counts and docID checksums (DocumentsOnly) / counts and top-k scores (top-k mode, 1e-5 relative)"""
    import trinity_b200 as tb
    out = {"queries_checked": int(n), "match_counts_equal": bool(np.array_equal(np.asarray(res.match_counts[:n], np.uint64), counts[:n]))}
    if mode == tb.MODE_DOCS_ONLY:
        off = np.asarray(res.offsets[: n + 1], np.int64)
        ids = np.asarray(res.docids[: off[-1]], np.uint64)
        cs = np.concatenate([np.zeros(1, np.uint64), np.cumsum(ids, dtype=np.uint64)])  # (a Python 0 would promote the array to float64)
        got = cs[off[1:]] - cs[off[:-1]]
        out["docid_checksums_equal"] = bool(np.array_equal(got, sums[:n]))
    elif mode == tb.MODE_SCORED_TOPK:
        worst = 0.0
        for q in range(n):
            d, s = res.query(q)
            want = tsc[q][: len(s)]
            if len(s) != int(min(k, counts[q])):
                worst = float("inf")
                break
            if len(s):
                worst = max(worst, float(np.max(np.abs(np.asarray(s, np.float64) - want) / np.maximum(np.abs(want), 1e-30))))
        out["topk_scores_max_rel_err"] = worst
        out["topk_scores_within_1e-5"] = bool(worst <= 1e-5)
    return out


def cpu_baseline(synth, texts, args, mode, threads: int | None = None, sample: int | None = None, repeats: int = 1, check=None):
    """the reference's own exec_query (oracle/_ref == the reference compiled in place) on the host cores"""
    sys.path.insert(0, str(ROOT / "tests"))
    from refharness import RefIndex, load_ref

    ref = load_ref()
    cores = threads or (os.cpu_count() or 1)
    r = RefIndex.from_bytes(ref, synth.codec, np.asarray(synth.index), np.asarray(synth.hits), synth.names, np.asarray(synth.terms), args.ndocs, synth.sum_hits)
    n = sample or args.cpu_sample
    if not n:  # bounded sample: a probe sizes it for ~30 s of host time (the whole batch for the headline workload)
        probe = min(len(texts), 64)
        el, *_ = r.exec_batch(texts[:probe], mode != 0, args.k, cores)
        n = max(min(32, len(texts)), min(len(texts), int(30.0 * probe / max(el, 1e-9))))
    qs = texts[:n]
    best = None
    for _ in range(repeats):
        el, counts, sums, tid, tsc = r.exec_batch(qs, mode != 0, args.k, cores)
        best = el if best is None else min(best, el)
    out = {"value": n / best, "unit": "queries/s", "cores": cores, "kind": "reference",
           "sample": f"first {n} queries of the same batch, one query per host thread ({cores} threads), {best:.2f} s wall",
           "seconds": best}
    if check is not None:
        try:
            out["parity"] = full_size_parity(check, mode, args.k, counts, sums, tid, tsc, n)
        except Exception as e:  # the checker must never take the bench line down
            out["parity"] = {"error": repr(e)}
    return out


def reference_arm(args, rank, world, wl, K, W):
    """--impl reference: the reference's own CPU exec_query on the box's host cores (oracle/_ref), same config/metric."""
    if rank != 0:
        return
    import trinity_b200 as tb  # host-side index build only (bytes identical to the reference encoders'; tests/test_codecs_cpu.py)

    synth = tb.SynthIndex(wl["codec"], args.ndocs, args.nterms, threads=os.cpu_count() or 8)
    texts, _ = gen_queries(args.workload, args.nq, args.nterms)
    cores = os.cpu_count() or 1
    sys.path.insert(0, str(ROOT / "tests"))
    from refharness import RefIndex, load_ref

    r = RefIndex.from_bytes(load_ref(), synth.codec, np.asarray(synth.index), np.asarray(synth.hits), synth.names, np.asarray(synth.terms), args.ndocs, synth.sum_hits)
    # bounded sample: a probe of the first queries sizes the per-step sample so that the W + K steps together stay near 150 s of host
    # time (the whole batch at the default K/W; a prefix of it — the batch is in random order — when the driver asks for many steps)
    n = args.cpu_sample
    if not n:
        probe = min(len(texts), 128)
        el, *_ = r.exec_batch(texts[:probe], wl["mode"] != 0, args.k, cores)
        n = max(min(32, len(texts)), min(len(texts), int(150.0 * (probe / max(el, 1e-9)) / max(1, K + W))))
    qs = texts[:n]
    for _ in range(W):
        r.exec_batch(qs, wl["mode"] != 0, args.k, cores)
    tot = 0.0
    for _ in range(K):
        el, *_ = r.exec_batch(qs, wl["mode"] != 0, args.k, cores)
        tot += el
    v = n * K / tot
    print(json.dumps({
        "metric": "queries/sec (batched 2-term AND, 100M-doc Zipfian synthetic index)" if args.workload == "and2" else f"queries/sec ({args.workload})",
        "value": v, "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": tot * 1e3 / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{wl['desc']}; {args.ndocs} docs, {args.nterms} terms, Zipf(1) df, batch {args.nq}, seed 0xC0FFEE",
                   "codec": "GOOGLE" if wl["codec"] == 0 else "LUCENE"},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "reference",
                         "sample": f"each step = first {n} queries of the batch, one query per host thread ({cores} threads)"},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


if __name__ == "__main__":
    main()
