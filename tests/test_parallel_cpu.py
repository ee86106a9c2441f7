"""CPU, world_size 2, gloo: the host-side sharding and the single end-of-run gather (no data-path collective exists)."""
import os
import socket
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from morphganformer_b200 import parallel
    lo, hi = parallel.shard_range(n_items, rank, world)
    # each "job" i produces latent = i and loss = 10 i, independent of the rank that ran it
    lat = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1).repeat(1, 17, 32)
    los = torch.arange(lo, hi, dtype=torch.float32) * 10
    glat, glos = parallel.gather_results(lat, los)
    q.put((rank, lo, hi, glat[:, 0, 0].tolist(), glos.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5])
def test_shard_and_gather_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, lat0, los0), (r1, lo1, hi1, lat1, los1) = res
    assert lo0 == 0 and hi0 == lo1 and hi1 == n_items
    want = [float(i) for i in range(n_items)]
    assert lat0 == want and lat1 == want
    assert los0 == [10 * v for v in want] and los1 == los0


def test_shard_range_properties():
    from morphganformer_b200 import parallel
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_projection_schedule_matches_oracle():
    from morphganformer_b200 import projection as P
    from oracle import projection as O
    for t in (0.0, 0.01, 0.05, 0.3, 0.76, 0.99):
        assert abs(P.get_lr(t, 0.1) - O.get_lr(t, 0.1)) < 1e-12
    x = torch.randn(100, 17, 32)
    m1, s1 = P.latent_stats(x); m2, s2 = O.latent_stats(x)
    assert torch.equal(m1, m2) and float(s1) == float(s2)
