"""N>1 host-side logic on CPU (gloo, world_size 2): the model blob is built once on rank 0, broadcast as bytes,
and every rank takes a byte-balanced shard of the batch; the gathered per-shard results equal the single-process
result.  The per-shard scorer here is the CPU oracle (no GPU in this test); on the GPU box the same flow runs with
Predictor.from_blob over NCCL (bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vaporetto_b200 as vb
from vpt_testlib import synth
from vpt_testlib.oracle import OraclePredictor

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, model_bytes, text, offs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. one-time model broadcast (bytes)
        if rank == 0:
            blob = torch.from_numpy(vb.build_blob(vb.Model.read(model_bytes)))
            size = torch.tensor([blob.numel()], dtype=torch.int64)
        else:
            size = torch.zeros(1, dtype=torch.int64)
        dist.broadcast(size, 0)
        if rank != 0:
            blob = torch.empty(int(size.item()), dtype=torch.uint8)
        dist.broadcast(blob, 0)
        assert bytes(blob[:7].numpy()) == b"VPTB200"
        ref = vb.build_blob(vb.Model.read(model_bytes))
        assert np.array_equal(blob.numpy(), ref)  # deterministic build: every rank could rebuild the same bytes
        # 2. shard by bytes, score the shard
        lo, hi = vb.shard_by_bytes(offs, rank, world)
        o = OraclePredictor(model_bytes)
        sc, bd, boff, st = o.predict_batch(text, offs[lo:hi + 1])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), lo=lo, hi=hi, sc=sc, bd=bd)
        # 3. the only other cross-rank step: max of the step times (bench.py) — exercise the reduction
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_broadcast(tmp_path):
    model_bytes = open(os.path.join(GOLDEN, "model.bin"), "rb").read()
    text, offs, _ = synth.gen_text(400, ragged=True)
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, model_bytes, text, offs, str(tmp_path)), nprocs=world, join=True)
    o = OraclePredictor(model_bytes)
    sc, bd, boff, st = o.predict_batch(text, offs)
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    assert int(parts[0]["lo"]) == 0 and int(parts[-1]["hi"]) == len(offs) - 1
    assert int(parts[0]["hi"]) == int(parts[1]["lo"])
    assert np.array_equal(np.concatenate([p["sc"] for p in parts]), sc)
    assert np.array_equal(np.concatenate([p["bd"] for p in parts]), bd)
    # byte balance within one sentence of the ideal
    b0 = int(offs[int(parts[0]["hi"])] - offs[0])
    assert abs(b0 - int(offs[-1]) / 2) < 600


def _lines_worker(rank, world, port, model_bytes, data, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = vb.shard_lines(data, rank, world)
        out, nl = OraclePredictor(model_bytes).tokenize_lines(data[lo:hi], wsconst="D")
        with open(os.path.join(out_dir, f"l{rank}.bin"), "wb") as f:
            f.write(out)
        n = torch.tensor([nl], dtype=torch.int64)
        dist.all_reduce(n)  # total line count over the ranks
        with open(os.path.join(out_dir, f"n{rank}.txt"), "w") as f:
            f.write(str(int(n.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_line_sharding(tmp_path):
    """The lines path (vpt_tokenize_lines) shards by bytes at line ends; per-rank outputs concatenate to the
    single-process output.  The per-shard worker here is the CPU oracle (no GPU in this test)."""
    model_bytes = open(os.path.join(GOLDEN, "model.bin"), "rb").read()
    text, offs, _ = synth.gen_text(300, ragged=True)
    lines = [text[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)]
    data = b"\r\n".join(lines[:100]) + b"\n\n" + b"\n".join(lines[100:])  # unterminated last line
    world = 2
    mp.spawn(_lines_worker, args=(world, _free_port(), model_bytes, data, str(tmp_path)), nprocs=world, join=True)
    want, nl = OraclePredictor(model_bytes).tokenize_lines(data, wsconst="D")
    got = b"".join(open(os.path.join(tmp_path, f"l{r}.bin"), "rb").read() for r in range(world))
    assert got == want
    assert int(open(os.path.join(tmp_path, "n0.txt")).read()) == nl


def test_shard_lines_properties():
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(0, 200))
        data = bytes(rng.choice(np.frombuffer(b"ab\n\r", np.uint8), size=n, p=[0.4, 0.35, 0.2, 0.05]))
        world = int(rng.integers(1, 6))
        cuts = [vb.shard_lines(data, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert b == c and a <= b
        for lo, hi in cuts[:-1]:
            assert hi == lo or hi == n or data[hi - 1:hi] == b"\n"   # a cut sits right after a line end


def test_shard_by_bytes_properties():
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(0, 40))
        lens = rng.integers(0, 100, size=n)
        offs = np.zeros(n + 1, np.uint64)
        np.cumsum(lens, out=offs[1:])
        for world in (1, 2, 3, 8):
            cuts = [vb.shard_by_bytes(offs, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
