"""Host-side frame logic without a GPU: the tile job, partitioning, and the film reduce over gloo
with two processes (the N > 1 path of bench.py / the multi-GPU renderer)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from phosphorus_mk2_b200.frame import Tiles, samples_of_rank, tiles_of_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tiles_match_the_reference_tiling():
    t = Tiles.make(1920, 1080, 32)
    assert t.size == 60 * 34  # 2040 tiles, last row 24 high
    assert t.tiles[0] == (0, 0, 32, 32) and t.tiles[-1] == (1888, 1056, 32, 24)
    assert sum(w * h for (_, _, w, h) in t.tiles) == 1920 * 1080
    t = Tiles.make(70, 50, 32)
    assert [x[2] for x in t.tiles[:3]] == [32, 32, 6] and t.tiles[-1] == (64, 32, 6, 18)


def test_cursor_hands_out_every_tile_once():
    t = Tiles.make(256, 256, 32)
    seen = []
    while True:
        c = t.next_chunk(5)
        if not c:
            break
        seen += c
    assert seen == t.tiles and t.next() is None


def test_partitions_cover_the_frame():
    tiles = Tiles.make(640, 360).tiles
    for world in (1, 2, 4, 8):
        parts = [tiles_of_rank(tiles, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(tiles)
        rng = [samples_of_rank(1024, r, world) for r in range(world)]
        assert rng[0][0] == 0 and rng[-1][1] == 1024 and all(rng[i][1] == rng[i + 1][0] for i in range(world - 1))
    assert [samples_of_rank(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_film_reduce_two_ranks_gloo(tmp_path):
    """world_size 2 over gloo: each rank fills the tiles it owns, reduce onto rank 0 rebuilds the frame."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from phosphorus_mk2_b200.frame import Tiles, tiles_of_rank, samples_of_rank, reduce_film
        dist.init_process_group("gloo")
        r, n = dist.get_rank(), dist.get_world_size()
        W, H = 96, 64
        tiles = Tiles.make(W, H).tiles
        full = np.arange(W * H * 4, dtype=np.float32).reshape(H, W, 4)
        full[..., 3] = 1.0  # alpha = 1 where rendered
        # tile-partitioned: own tiles carry the final value, the rest is zero
        film = np.zeros_like(full)
        for (x, y, w, h) in tiles_of_rank(tiles, r, n):
            film[y:y+h, x:x+w] = full[y:y+h, x:x+w]
        t = torch.from_numpy(film)
        reduce_film(t, dist, 0)
        if r == 0:
            assert np.array_equal(film, full)
        # sample-partitioned: each rank holds its weighted share of every pixel
        a, b = samples_of_rank(16, r, n)
        share = full * ((b - a) / 16.0)
        share[..., 3] = 1.0  # every rank renders every pixel: the sum of the alphas is clamped back to 1
        t = torch.from_numpy(share)
        reduce_film(t, dist, 0)
        if r == 0:
            assert np.allclose(t.numpy(), full, rtol=1e-6)
        dist.barrier()
        dist.destroy_process_group()
        print("rank%d-ok" % r, flush=True)
    """))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout
