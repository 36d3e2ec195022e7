"""Frame plumbing around the device: the tile job, the film sink and the frame state the reference
hands to ``xpu_t::start`` — plus the multi-GPU partitioning of a frame.

Mirrors ``job::tiles_t`` (src/jobs/tiles.hpp:10-90: 32 x 32 tiles, one atomic cursor shared by all
devices), ``film_t<>`` (src/film.hpp:10-16: ``add_tile(pos, size, buffer)``), ``frame_state_t``
(src/state.hpp:18-31) and the device thread of ``cpu_t::start`` / ``join`` (src/xpu/cpu.cpp:223-244).
"""
from __future__ import annotations

import threading

import numpy as np

from .device import CudaDevice, make_tiles


class Tiles:
    """job::tiles_t: precomputed tile list + a thread-safe cursor (``next``)."""

    def __init__(self, width: int, height: int, tile_size: int = 32):
        self.width, self.height, self.tile_size = width, height, tile_size
        self.tiles = make_tiles(width, height, tile_size)
        self.size = len(self.tiles)
        self._cursor = 0
        self._lock = threading.Lock()

    @staticmethod
    def make(width: int, height: int, tile_size: int = 32) -> "Tiles":
        return Tiles(width, height, tile_size)

    def next(self):
        """One tile, or None when the job is drained (tiles_t::next, tiles.hpp:40-47)."""
        got = self.next_chunk(1)
        return got[0] if got else None

    def next_chunk(self, n: int):
        """Up to n consecutive tiles with one cursor update: a GPU drains the queue in chunks so a
        wavefront covers many tiles (SURVEY.md 'tile granularity')."""
        with self._lock:
            a = self._cursor
            b = min(self.size, a + n)
            self._cursor = b
        return self.tiles[a:b]

    def remaining(self) -> int:
        with self._lock:
            return self.size - self._cursor

    def take_guided(self, peers: int = 1, lo: int = 64, hi: int = 1024):
        """The claim of integration/cuda.cpp (take_tiles): half of this device's share of what is left, at least `lo`
        tiles (a wavefront over few pixels idles in its launch tails), at most `hi` — so that the devices sharing the
        queue (session.cpp:85-99) finish together."""
        return self.next_chunk(min(hi, max(lo, self.remaining() // (2 * max(1, peers)))))


class MemoryFilm:
    """film_t<>: an in-memory sink; ``add_tile`` may be called from several device threads."""

    def __init__(self, width: int, height: int):
        self.rgba = np.zeros((height, width, 4), np.float32)
        self.tiles_added = 0
        self._lock = threading.Lock()

    def add_tile(self, pos, size, buffer: np.ndarray) -> None:
        (x, y), (w, h) = pos, size
        with self._lock:
            self.rgba[y:y + h, x:x + w, :] = buffer
            self.tiles_added += 1


class FrameState:
    """frame_state_t{sampler, tiles, film}; the sampler is reduced to what a device reads: spp and
    the seed of the counter-based generator."""

    def __init__(self, tiles: Tiles, film, spp: int, seed: int = 0):
        self.tiles, self.film, self.spp, self.seed = tiles, film, spp, seed


class DeviceThread:
    """What cpu_t::start spawns per worker, for one GPU: drain the shared tile queue in chunks,
    render each chunk as one wavefront, hand every finished tile to the film sink."""

    def __init__(self, dev: CudaDevice, frame: FrameState, chunk_tiles: int = 1024, sample_range=None, peers: int = 1):
        self.dev, self.frame, self.chunk, self.error = dev, frame, chunk_tiles, None
        self.sample_range = sample_range or (0, frame.spp)
        self.peers, self.tiles_done, self.claims = peers, 0, 0
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            f = self.frame
            W = f.tiles.width

            def claim():
                tiles = f.tiles.take_guided(self.peers, hi=self.chunk)
                if tiles:
                    self.dev.render(tiles, self.sample_range[0], self.sample_range[1], f.spp, f.seed)  # asynchronous
                return tiles

            cur = claim()
            while cur:
                y0, y1 = min(t[1] for t in cur), max(t[1] + t[3] for t in cur)
                rows = self.dev.film_read(0, y0, W, y1 - y0)  # the chunk's film rows, once (blocks until rendered)
                nxt = claim()                                  # the GPU renders the next chunk while the host slices this one
                for (x, y, w, h) in cur:
                    f.film.add_tile((x, y), (w, h), rows[y - y0:y - y0 + h, x:x + w])
                self.tiles_done += len(cur)
                self.claims += 1
                cur = nxt
        except Exception as e:  # surfaced by join(), like an exception escaping a reference worker
            self.error = e


def start(dev: CudaDevice, scene, frame: FrameState, **kw) -> DeviceThread:
    """xpu_t::start: non-blocking; returns the handle ``join`` waits on."""
    dev.film_clear()
    t = DeviceThread(dev, frame, **kw)
    t.thread.start()
    return t


def join(handle: DeviceThread) -> None:
    """xpu_t::join."""
    handle.thread.join()
    if handle.error is not None:
        raise handle.error


# ---- multi-GPU partitioning (one process per GPU, scene replicated) ------------------------------------
def tiles_of_rank(tiles, rank: int, world: int):
    """Tile-partitioned frame: rank r renders tiles r, r + world, ... (disjoint, covers all)."""
    return tiles[rank::world]


def samples_of_rank(spp: int, rank: int, world: int):
    """Sample-partitioned frame: rank r renders the contiguous sample range [r*spp/world, (r+1)*spp/world) — empty for
    some ranks when spp < world (render_partition then renders nothing on that rank: it only joins the reduce)."""
    return (spp * rank) // world, (spp * (rank + 1)) // world


def render_partition(dev: CudaDevice, tiles, sample_range, spp_total: int, seed: int = 0) -> None:
    """This rank's share of a frame into the (cleared) device film; a rank whose share is empty contributes zeros."""
    dev.film_clear()
    if tiles and sample_range[1] > sample_range[0]:
        dev.render(tiles, sample_range[0], sample_range[1], spp_total, seed)


def reduce_film(film, dist, root: int = 0):
    """The one collective of a frame: sum the per-rank films onto `root` (NCCL over NVLink for CUDA
    tensors, gloo for the CPU tests).  Disjoint tiles make the sum a gather; weighted sample ranges
    make it the average.  `film` is a torch tensor (device film wrapped zero-copy, or a host film)."""
    dist.reduce(film, dst=root, op=dist.ReduceOp.SUM)
    if dist.get_rank() == root:  # every rank wrote alpha = 1 where it rendered: a sample-partitioned sum leaves `world` there
        film.view(-1, 4)[:, 3].clamp_(max=1.0)
    return film


class _CudaArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def film_tensor(dev: CudaDevice):
    """The device film as a torch tensor sharing the library's memory (no copy)."""
    import torch
    ptr, n = dev.film_device_ptr()
    return torch.as_tensor(_CudaArray(ptr, n), device=f"cuda:{dev.device}")
