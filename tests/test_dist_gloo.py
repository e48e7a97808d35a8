"""The N>1 path on CPU: world_size 2, gloo.  Coalitions are sharded across ranks, one all-gather
brings the (correct, loss_sum) pairs back, and every rank ends with the same memo -- bit-identical
to the single-process run."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_worker(tmp, rank=None, world=None, port=None):
    env = dict(os.environ)
    out = os.path.join(tmp, f"out_{rank}.json")
    if rank is not None:
        env.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(rank))
    else:
        for k in ("RANK", "WORLD_SIZE"):
            env.pop(k, None)
    env["OMP_NUM_THREADS"] = "2"
    return subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py"), out], env=env), out


def test_two_rank_gloo_matches_single_process(tmp_path):
    tmp = str(tmp_path)
    p0, single_out = run_worker(tmp)
    assert p0.wait(timeout=600) == 0
    single = json.load(open(single_out))
    port = free_port()
    procs = [run_worker(tmp, r, 2, port) for r in range(2)]
    for p, _ in procs:
        assert p.wait(timeout=600) == 0
    ranks = [json.load(open(o)) for _, o in procs]
    assert [r["world"] for r in ranks] == [2, 2]
    # work was split: 7 coalitions -> 4 + 3
    assert sorted(r["evaluated_here"] for r in ranks) == [3, 4]
    assert single["evaluated_here"] == 7
    for r in ranks:
        assert r["counts"] == single["counts"]          # integer counts and fp64 loss sums, bit-identical
        assert r["sv"] == single["sv"]
    assert ranks[0]["bounds"] == [[0, 4], [4, 7]]
    # one coalition for two ranks: each rank scored its half of the 32 validation images, the all-reduced
    # count is exact and the fp64 loss sum equals the single-process one up to the order of the last addition
    assert [r["split"]["ranges"] for r in ranks] == [[[0, 16]], [[16, 32]]] and single["split"]["ranges"] == [None]
    for r in ranks:
        assert r["split"]["counts"][0] == single["split"]["counts"][0]
        assert r["split"]["counts"][1] == pytest.approx(single["split"]["counts"][1], rel=1e-12)
        assert r["split"]["utility"][0] == single["split"]["utility"][0]
    assert ranks[0]["split"] == ranks[1]["split"] or ranks[0]["split"]["counts"] == ranks[1]["split"]["counts"]


    # stochastic estimators: with seed None the ranks still drew the same coalitions (shared seed) and agree;
    # with an explicit seed the two-rank result is the single-process one
    assert ranks[0]["stochastic"] == ranks[1]["stochastic"]
    assert ranks[0]["stochastic"]["cc_seed5"] == single["stochastic"]["cc_seed5"]


def test_rows_digest_detects_divergent_lists():
    from shapley_vit_b200.dist import _rows_digest

    a = [[0.5, 0.5, 0.0], [0.0, 1.0, 0.0]]
    assert _rows_digest(a) == _rows_digest([list(r) for r in a])
    assert _rows_digest(a) != _rows_digest(a[::-1])
    assert 0 <= _rows_digest(a) < 2 ** 63


def test_shard_bounds_cover_everything():
    from shapley_vit_b200.dist import shard_bounds

    for n in (0, 1, 7, 8, 255, 1023):
        for ws in (1, 2, 4, 8):
            got = []
            for r in range(ws):
                lo, hi = shard_bounds(n, r, ws)
                got += list(range(lo, hi))
            assert got == list(range(n))
