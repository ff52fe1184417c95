"""world_size-2 gloo test of the batch sharding + status gather (host logic of config C5, no GPU).
Each rank solves its slice of a small batch with the CPU oracle (a stand-in for the per-GPU batched solver),
the records are all-gathered, and the assembled table must equal a single-process run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from madqp_jl_b200 import batch, problems
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = batch.shard_range(n_items, rank, world)
    st, it, ob, ip, idu, sec = [], [], [], [], [], []
    for i in range(lo, hi):
        s = O.MPCOracle(problems.small_qp(instance=i, n=30, m=12, nnz=60), regularization=("fixed", 1e-8, -1e-8)).solve()
        st.append(s.status); it.append(s.k); ob.append(s.objective); ip.append(s.inf_pr); idu.append(s.inf_du); sec.append(s.total_time)
    table = batch.gather_records(batch.make_records(st, it, ob, ip, idu, sec), n_items)
    np.save(os.path.join(out_dir, f"table_{rank}.npy"), table)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 6])
def test_shard_and_gather_two_ranks(tmp_path, n_items):
    import torch.multiprocessing as mp
    from madqp_jl_b200 import batch, problems
    from oracle import oracle as O
    port = 29500 + os.getpid() % 500 + n_items
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "table_0.npy"), np.load(tmp_path / "table_1.npy")
    assert np.array_equal(t0, t1) and t0.shape == (n_items, len(batch.RECORD_FIELDS))
    for i in range(n_items):
        s = O.MPCOracle(problems.small_qp(instance=i, n=30, m=12, nnz=60), regularization=("fixed", 1e-8, -1e-8)).solve()
        assert batch.STATUS_NAMES[int(t0[i, 0])] == s.status and int(t0[i, 1]) == s.k
        assert t0[i, 2] == s.objective


def test_shard_range_covers_everything():
    from madqp_jl_b200 import batch
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = batch.shard_range(n, r, world)
                got += list(range(lo, hi))
            assert got == list(range(n))
