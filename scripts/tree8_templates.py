"""Where the 8-term tree batch spends its time: device-resident time of the batch's queries grouped by the template they were drawn from
(bench.py gen_queries: 0 = (a|b)&(c|d)&e -(f|g|h), 1 = a&b&c -d -e, 2 = (a&b)|(c&d)|(e&f) -g -h, 3 = a&(b|c|d) -(e&f) &(g|h))."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import trinity_b200 as tb  # noqa: E402


def main():
    import torch
    ndocs, nterms, nq = 100_000_000, 4096, 1000
    synth = tb.SynthIndex(0, ndocs, nterms)
    g = tb.GpuIndexSource(0)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.upload(0, np.asarray(synth.index), np.asarray(synth.terms), ndocs)
    texts, _ = bench.gen_queries("tree8", nq, nterms)
    # the template of a query, recovered from its shape
    def tpl(t):
        if t.startswith("(") and ") AND (" in t and " NOT (" in t and t.count(" OR ") == 4:
            return 0
        if t.count(" NOT ") == 2 and "(" not in t:
            return 1
        if ") OR (" in t:
            return 2
        return 3
    tdict = tb.TermDictionary(synth.names)
    out = {}
    for k in (0, 1, 2, 3, -1):
        qs = [t for t in texts if k < 0 or tpl(t) == k]
        plans = [tb.parse_query(q, tdict) for q in qs]
        packed = g.pack(plans)
        for _ in range(2):
            g.exec_batch_device(plans, tb.MODE_DOCS_COMPACT, 100, packed=packed)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(3):
            g.exec_batch_device(plans, tb.MODE_DOCS_COMPACT, 100, packed=packed)
        ev1.record()
        torch.cuda.synchronize()
        r = g.fetch()
        out["all" if k < 0 else f"template{k}"] = {"queries": len(qs), "ms": round(ev0.elapsed_time(ev1) / 3, 2), "matches": int(r.match_counts.sum()), "example": qs[0]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
