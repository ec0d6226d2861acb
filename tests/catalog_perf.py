"""Perf harness for the catalog kernel (not collected by pytest).
    python tests/catalog_perf.py cg U I D K [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_data as synth  # noqa: E402


def run(cg, U, I, D, K, reps=3, sets=0, tn=0, asp=0, seed=11):
    from foodrec_b200 import Engine, Hyper
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(seed)
    P = torch.randn((U, 5, D), generator=g, device=dev) * 0.1
    R = torch.randn((I, D), generator=g, device=dev) * 0.1
    Cat = torch.randn((4, D), generator=g, device=dev) * 0.1
    G = torch.zeros((9, 5, D), device=dev)
    ic = synth.make_item_categories(I, seed=seed + 1)
    e = Engine(Hyper(), P, R, Cat, G, max_rows=256, item_cats=ic)
    del P, R
    e.catalog_prepare(cta_group=cg, epi_sets=sets, tile_n=tn, a_split=asp)
    info = e.catalog_info()
    e.timing_enable(True)
    for r in range(reps):
        if os.environ.get("FOODREC_CATALOG_CYCLES") and r: e.catalog_cycle_counters()
        torch.cuda.synchronize(); t0 = time.time()
        ids, sc = e.catalog_topk(K=K)
        torch.cuda.synchronize(); t1 = time.time()
        ms, npass = e.catalog_timing_read()
        flop = 2.0 * U * info["tiles"] * (info["epi_sets"] % 1000) * info["k_padded"] * (2 if info["cta_group"] >= 10 else 1)
        print(f"cg={cg} U={U} I={I} D={D} K={K} rep{r}: wall {1e3 * (t1 - t0):.2f} ms  phases {dict((k, round(v, 3)) for k, v in ms.items())} passes={npass} "
              f"gemm {flop / ms['gemm_filter'] / 1e9:.1f} TFLOP/s executed; users/s {U / (t1 - t0):.0f}", flush=True)
    if os.environ.get("FOODREC_CATALOG_CYCLES"):
        print("cycles:", {k: round(v / 1e6, 2) for k, v in e.catalog_cycle_counters().items()}, "(M cycles, last rep)", flush=True)
    print("fallback rows in last pass:", e.catalog_fallback_rows(), flush=True)
    # spot check a few users against a torch fp64 full scan (dev check only)
    ic_d = torch.as_tensor(ic, device=dev, dtype=torch.float64)
    w = ic_d / ic_d.sum(1, keepdim=True)
    a = float(np.float32(0.99)); b = float(np.float32(1) - np.float32(0.99))
    Rd = e.R.double()
    bad = 0
    for u in (0, 1, U // 2, U - 1):
        Pu = e.P[u].double()
        high = (w @ (e.Cat.double() @ Pu[0])) * a
        low = ((Rd @ Pu[1:].T) * w).sum(1) * b
        sc_all = high + low
        top = torch.topk(sc_all, K).indices.sort().values
        got = ids[u].long().sort().values
        bad += int((top != got).sum())
    print("spot-check id-set mismatches:", bad, flush=True)
    e.close()


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    run(*a)
