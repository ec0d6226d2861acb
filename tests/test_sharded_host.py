"""CPU tests of the host side of the row-sharded trainer: the layout helpers and -- with a
world_size-2 gloo group -- the collective wiring of DistRunner (which block goes to which
rank, in which direction, between which phases)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from foodrec_b200 import sharded
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_unshard_round_trip_and_ownership():
    x = np.arange(11 * 3, dtype=np.float32).reshape(11, 3)
    for W in (1, 2, 3, 4, 8):
        parts = [sharded.shard_rows(x, r, W) for r in range(W)]
        assert all(p.shape[0] == sharded.local_rows(11, W) for p in parts)
        np.testing.assert_array_equal(sharded.unshard_rows(parts, 11), x)
        for r, p in enumerate(parts):
            for k in range(len(range(r, 11, W))):
                np.testing.assert_array_equal(p[k], x[r + k * W])      # local row k <-> global r + k*W


def test_route_batch_is_stable_and_complete():
    users = np.array([5, 2, 9, 2, 7, 4, 5, 0])
    idx = sharded.route_batch(users, 3)
    assert sorted(np.concatenate(idx).tolist()) == list(range(8))
    for r, ix in enumerate(idx):
        assert (users[ix] % 3 == r).all() and (np.diff(ix) > 0).all()


def test_label_csr_sharding():
    off, idx = synth.make_user_label_csr(23, 9, seed=4)
    for W in (2, 4):
        for r in range(W):
            loff, lidx = sharded.shard_label_csr(off, idx, r, W, 23)
            assert loff.shape == (sharded.local_rows(23, W) + 1,)
            for k, u in enumerate(range(r, 23, W)):
                assert lidx[loff[k]:loff[k + 1]].tolist() == idx[off[u]:off[u + 1]].tolist()


WIRING = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, "__ROOT__")
    from foodrec_b200.sharded import DistRunner
    dist.init_process_group("gloo")
    r, W, cap, D = dist.get_rank(), dist.get_world_size(), 3, 2
    log = []
    class Stub:                                   # the five phases, CPU tensors, recognisable payloads; two plan slots
        rank, world, device = r, W, "cpu"
        def __init__(s):
            n = W * cap
            s.cap = cap
            s._req = [torch.empty(n, dtype=torch.int32) for _ in range(2)]; s._rreq = [torch.empty(n, dtype=torch.int32) for _ in range(2)]
            s._n_plan = s._n_apply = 0; s.batch = None; s.batches = [None, None]
            s.rows = torch.empty(n, D); s.rbuf = torch.empty(n, D)
            s.grows = torch.empty(n, D); s.rgrows = torch.empty(n, D)
            s.packed = torch.zeros(5)
        rreq = property(lambda s: s._rreq[s._n_apply & 1])
        def planned(s): return s._n_plan > s._n_apply
        def plan(s):                              # request j to owner o carries 1000*batch + 100*me + 10*o + j
            slot = s._n_plan & 1; s.batches[slot] = s.batch
            s._req[slot].copy_(torch.tensor([1000 * s.batch + 100 * r + 10 * o + j for o in range(W) for j in range(cap)], dtype=torch.int32))
            s._n_plan += 1; log.append("plan%d" % s.batch); return s._req[slot]
        def serve_prepare(s):                     # one step ahead: the requests of the step planned LAST are already here
            k = s.batches[(s._n_plan - 1) & 1]
            v = s._rreq[(s._n_plan - 1) & 1].view(W, cap)
            for src in range(W):
                assert v[src].tolist() == [1000 * k + 100 * src + 10 * r + j for j in range(cap)], v
            log.append("prep%d" % k)
        def serve(s):                             # I am the owner: every request must be addressed to me, for THIS step's batch
            k = s.batches[s._n_apply & 1]
            v = s.rreq.view(W, cap)
            for src in range(W):
                assert v[src].tolist() == [1000 * k + 100 * src + 10 * r + j for j in range(cap)], v
            s.rows.copy_((s.rreq.float() + 0.5).unsqueeze(1).expand(-1, D)); log.append("serve%d" % k)
        def forward(s):                           # rows come back in MY slot order o*cap + j
            k = s.batches[s._n_apply & 1]
            assert s.rbuf[:, 0].tolist() == [1000 * k + 100 * r + 10 * o + j + 0.5 for o in range(W) for j in range(cap)]
            s.packed.fill_(r + 1.0); log.append("forward%d" % k)
        def update(s, wp):
            assert s.packed.tolist() == [sum(range(1, W + 1))] * 5      # all-reduce SUM
            s.grows.copy_(s.rbuf * -1.0); log.append("update%d" % s.batches[s._n_apply & 1])
        def apply(s):                             # gradient rows aligned with the requests I received for this step
            assert torch.equal(s.rgrows[:, 0], -(s.rreq.float() + 0.5))
            log.append("apply%d" % s.batches[s._n_apply & 1]); s._n_apply += 1; return "done"
    g = Stub(); run = DistRunner(g)
    def batch(k):
        def f(): g.batch = k
        return f
    batch(0)()
    assert run.step() == "done" and log == ["plan0", "serve0", "forward0", "update0", "apply0"]
    # pipelined: the NEXT step's plan + request exchange are issued once this step is queued, before the next serve
    del log[:]; batch(1)()
    assert run.step(next_batch=batch(2)) == "done" and run.step(next_batch=batch(3)) == "done" and run.step() == "done"
    assert log == ["plan1", "serve1", "forward1", "update1", "apply1", "plan2", "prep2", "serve2", "forward2", "update2", "apply2",
                   "plan3", "prep3", "serve3", "forward3", "update3", "apply3"], log
    sys.stdout.write("WIRING_OK_%d\\n" % r); sys.stdout.flush()
""")


def test_dist_runner_wiring_gloo_world2(tmp_path):
    script = tmp_path / "wiring.py"
    script.write_text(WIRING.replace("__ROOT__", ROOT))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("WIRING_OK_") == 2, out.stdout


CATALOG_WIRING = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, "__ROOT__")
    from foodrec_b200.sharded import DistRunner
    dist.init_process_group("gloo")
    r, W, n, K, D = dist.get_rank(), dist.get_world_size(), 3, 2, 2
    class Stub:                       # item-sharded catalog protocol on CPU tensors with recognisable payloads
        rank, world = r, W
        def catalog_query_rows(s, users):          # my query j carries 10*me + j
            return torch.tensor([[10.0 * r + j] * D for j in range(n)]).view(n, 1, D)
        def catalog_local(s, allrows, K_):          # gathered rows arrive in rank order, block q = rank q's queries
            assert allrows.shape == (W * n, 1, D)
            assert allrows[:, 0, 0].tolist() == [10.0 * q + j for q in range(W) for j in range(n)]
            ids = torch.tensor([[1000 * r + int(v), 1000 * r + int(v) + 500] for v in allrows[:, 0, 0]], dtype=torch.int32)
            return ids, ids.double() + 0.25        # "my shard's list" for every query
        def catalog_merge(s, ids, sc):               # [W, n, K]: list w = shard w's answer for MY queries
            assert ids.shape == (W, n, K)
            for w in range(W):
                assert ids[w, :, 0].tolist() == [1000 * w + 10 * r + j for j in range(n)], ids
            assert torch.equal(sc, ids.double() + 0.25)
            return "merged"
    assert DistRunner(Stub()).catalog_topk(None, K=K) == "merged"
    sys.stdout.write("CATALOG_WIRING_OK_%d\\n" % r); sys.stdout.flush()
""")


def test_dist_runner_catalog_wiring_gloo_world2(tmp_path):
    """all-gather of query rows -> per-shard lists -> all-to-all back to the user owners."""
    script = tmp_path / "catalog_wiring.py"
    script.write_text(CATALOG_WIRING.replace("__ROOT__", ROOT))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("CATALOG_WIRING_OK_") == 2, out.stdout
