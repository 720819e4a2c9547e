"""CPU tests of the multi-GPU host logic: contiguous sequence shards per rank, no collective on
the path, all-gather only for verification.  Two processes over gloo on 127.0.0.1; the layer is replaced by
the CPU oracle (test infrastructure) so that the sharded result can be compared with the unsharded one."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skeleton_action_recognition_b200.sharding import shard_bounds, sharded_forward


def test_shard_bounds_partition_like_dataparallel():
    for n in (0, 1, 2, 7, 8, 255, 256, 257, 65536):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert all(s >= 0 for s in sizes) and max(sizes) == -(-n // world) if n else max(sizes) == 0
            if n:       # torch.chunk (what DataParallel.scatter uses) gives the same non-empty blocks
                ref = [c.shape[0] for c in torch.arange(n).chunk(world)]
                assert [s for s in sizes if s] == ref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import virtual_radar_oracle as vro
        torch.set_num_threads(2)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n, 3, 140, 25, 1, generator=g) * 0.3          # every rank builds the same batch
        layer = vro.OracleVirtualRadar(wavelength=5e-4)
        calls = []

        def layer_fn(xs):
            calls.append(xs.shape[0])
            return layer(xs) if xs.shape[0] else torch.empty(0, 256, 140 // 16 + 1)

        full = sharded_forward(layer_fn, x, gather=True)
        local = sharded_forward(layer_fn, x, gather=False)
        lo, hi = shard_bounds(n, world, rank)
        assert calls == [hi - lo, hi - lo]                              # only this rank's sequences were computed
        assert local.shape[0] == hi - lo and torch.equal(full[lo:hi], local)
        torch.save(full, os.path.join(out_dir, "full_%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_sharded_forward_two_ranks_gloo(tmp_path, n):
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    from oracle import virtual_radar_oracle as vro
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 3, 140, 25, 1, generator=g) * 0.3
    ref = vro.OracleVirtualRadar(wavelength=5e-4)(x)
    for r in range(world):
        full = torch.load(os.path.join(str(tmp_path), "full_%d.pt" % r))
        assert full.shape == ref.shape and torch.equal(full, ref)      # sequences are independent: bit-identical
