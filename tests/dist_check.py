"""Multi-process check of the row-sharded trainer over NCCL: launched by
tests/test_gpu_sharded.py::test_nccl_sharded_equals_unsharded (needs >= 2 GPUs) or by hand:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29641 tests/dist_check.py

Every rank steps its shard through DistRunner; rank 0 also steps an unsharded engine on the
same global batches and compares the gathered tables."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from foodrec_b200 import Engine, Hyper, sharded          # noqa: E402
from tests.util import Problem, assert_close             # noqa: E402


def main():
    rank, W = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    p = Problem(1003, 517, 9, 128, seed=71)
    h = Hyper(learner="adam", lr=0.01)
    rows, cols = np.nonzero(p.user_labels)
    cnt = np.bincount(rows, minlength=p.U)
    off = np.zeros(p.U + 1, np.int32); off[1:] = np.cumsum(cnt)
    eng = sharded.ShardedEngine(h, sharded.shard_rows(p.tb.P, rank, W), sharded.shard_rows(p.tb.R, rank, W),
                                p.tb.Cat, p.tb.G, rank, W, device=dev, max_rows=4096, adam_mode="lazy",
                                item_cats_global=p.item_cats,
                                user_label_csr_local=sharded.shard_label_csr(off, cols.astype(np.int32), rank, W, p.U),
                                max_label_entries=4096 * p.L)
    run = sharded.DistRunner(eng)
    if os.environ.get("DIST_CHECK_P2P"):
        run.enable_p2p()                     # fused gather/gradient + NVLink peer stores instead of the row all-to-alls
    single = None
    if rank == 0:
        single = Engine(h, p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, device=dev, max_rows=4096, adam_mode="lazy",
                        max_label_entries=4096 * p.L)
    B = 1500
    for s in range(4):
        f = p.bpr(B, seed=90 + s)
        ix = sharded.route_batch(f["user_input"], W)[rank]
        eng.set_batch(f["user_input"][ix] // W, f["item_input"][ix], neg_items=f["neg_item_input"][ix], global_batch=B)
        out = run.step().cpu().numpy()
        assert out[9] == 0
        if single is not None:
            single.train_step(f["user_input"], f["item_input"], categories=f["categories"], neg_items=f["neg_item_input"],
                              neg_categories=f["neg_categories"], user_one_hot_label=f["user_one_hot_label"])
            v = single.read_scalars()
            assert abs(out[0] - v[0]) <= 1e-5 * abs(v[0]), (out[0], v[0])
    eng.e.flush()
    Pl, Rl = eng.e.P.contiguous(), eng.e.R.contiguous()
    Ps = [torch.empty_like(Pl) for _ in range(W)]; Rs = [torch.empty_like(Rl) for _ in range(W)]
    dist.all_gather(Ps, Pl); dist.all_gather(Rs, Rl)
    if rank == 0:
        t = single.tables()
        assert_close(sharded.unshard_rows([x.cpu().numpy() for x in Ps], p.U), t["P"], rtol=1e-4, what="P")
        assert_close(sharded.unshard_rows([x.cpu().numpy() for x in Rs], p.I), t["R"], rtol=1e-4, what="R")
        assert_close(eng.e.G.cpu().numpy(), t["G"], rtol=1e-5, what="G")
        assert_close(eng.e.Cat.cpu().numpy(), t["Cat"], rtol=1e-4, what="Cat")
    sys.stdout.write(f"DIST_CHECK_OK_{rank}\n"); sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
