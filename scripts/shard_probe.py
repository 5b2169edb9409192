"""One docID-range shard of the N-GPU run on ONE GPU: device time of the whole and2 batch on shard `rank` of `world` (what each rank of
the scaling bench executes), to look at the small-shard behaviour of k_exec_docs without paying for N GPUs.
usage: shard_probe.py [world] [rank] [steps] [workload]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import trinity_b200 as tb  # noqa: E402
from trinity_b200.sharded import shard_range  # noqa: E402


def main():
    import torch
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    workload = sys.argv[4] if len(sys.argv) > 4 else "and2"
    ndocs, nterms, nq = 100_000_000, 4096, 1000
    lo, hi = shard_range(ndocs, rank, world)
    wl = bench.WORKLOADS[workload]
    synth = tb.SynthIndex(wl["codec"], ndocs, nterms, doc_range=(lo, hi))
    g = tb.GpuIndexSource(0)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.upload(wl["codec"], np.asarray(synth.index), np.asarray(synth.terms), ndocs)
    texts, _ = bench.gen_queries(workload, nq, nterms)
    tdict = tb.TermDictionary(synth.names)
    plans = [tb.parse_query(q, tdict) for q in texts]
    mode = tb.MODE_DOCS_COMPACT if wl["mode"] == tb.MODE_DOCS_ONLY else wl["mode"]
    packed = g.pack(plans)
    for _ in range(3):
        g.exec_batch_device(plans, mode, 100, packed=packed)
        r = g.exec_batch(plans, mode, 100, copy=False, packed=packed)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        g.exec_batch_device(plans, mode, 100, packed=packed)
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / steps
    tms = []
    for _ in range(steps):
        r = g.exec_batch(plans, mode, 100, copy=False, packed=packed)
        tms.append(g.last_timings())
    e2e = {k: round(float(np.mean([t[k] for t in tms])), 3) for k in tms[0]}
    print(json.dumps({"world": world, "rank": rank, "workload": workload, "device_ms_per_batch": round(dev_ms, 3), "qps_if_all_ranks_alike": nq / dev_ms * 1e3,
                      "e2e": e2e, "launches": int(r.kernel_launches), "result_bytes": int(r.result_bytes()) if wl["mode"] == tb.MODE_DOCS_ONLY else None}))


if __name__ == "__main__":
    main()
