"""Development harness for the catalog kernel (not collected by pytest): runs the CUDA path
through the C ABI against oracle/evaluate_oracle.catalog_topk and prints mismatches + timing.
    python tests/catalog_dev.py [cg] [U] [I] [D] [K]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import evaluate_oracle, synth  # noqa: E402
from oracle.recommender_oracle import Hyper as OHyper, OracleModel  # noqa: E402


def run(cg, U, I, D, K, splits=0, sets=0, tn=0, asp=0, seed=7, check=True):
    from foodrec_b200 import Engine, Hyper
    tb = synth.make_tables(U, I, 9, D, seed=seed)
    ic = synth.make_item_categories(I, seed=seed + 1)
    e = Engine(Hyper(), tb.P, tb.R, tb.Cat, tb.G, max_rows=256, item_cats=ic)
    t0 = time.time()
    e.catalog_prepare(cta_group=cg, splits=splits, epi_sets=sets, tile_n=tn, a_split=asp)
    torch.cuda.synchronize()
    t1 = time.time()
    e.timing_enable(True)
    ids, sc = e.catalog_topk(K=K)
    torch.cuda.synchronize()
    t2 = time.time()
    ms, npass = e.catalog_timing_read()
    info = e.catalog_info()
    print(f"cg={cg} U={U} I={I} D={D} K={K} splits={splits} prepare {t1 - t0:.3f}s topk {t2 - t1:.3f}s phases(ms)={ms} passes={npass} info={info}", flush=True)
    ids = ids.cpu().numpy(); sc = sc.cpu().numpy()
    if not check:
        return True
    om = OracleModel(tb.P, tb.R, tb.Cat, tb.G, OHyper(), dtype=np.float32)
    rid, rsc = evaluate_oracle.catalog_topk(om, np.arange(U), ic, K)
    bad = int((ids != rid).sum())
    rel = float(np.abs(sc - rsc).max() / np.abs(rsc).max())
    print(f"   id mismatches {bad} / {ids.size}; max rel score err {rel:.3e}", flush=True)
    if bad:
        r = np.argwhere(ids != rid)[0]
        print("   first mismatch row", r, "gpu", ids[r[0], :8], "ref", rid[r[0], :8], sc[r[0], :4], rsc[r[0], :4])
    e.close()
    return bad == 0 and rel < 1e-12


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    if a:
        ok = run(*a)
    else:
        ok = True
        for cg, tn, asp in ((1, 256, 0), (2, 256, 0), (2, 128, 0), (2, 256, 1)):
            for sets in (1, 2, 4):
                ok &= run(cg, 300, 3000, 64, 10, sets=sets, tn=tn, asp=asp)
                ok &= run(cg, 300, 5000, 128, 100, sets=sets, tn=tn, asp=asp)
                ok &= run(cg, 1000, 20000, 128, 100, splits=3, sets=sets, tn=tn, asp=asp)
                ok &= run(cg, 700, 9000, 128, 50, splits=1, sets=sets)      # whole sweeps: users sorted by best group
    print("OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
