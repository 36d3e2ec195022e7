"""The N > 1 frame on real GPUs: one process per GPU (torchrun), the frame partitioned by tiles and by sample ranges, the
films reduced INSIDE the library (phos_cuda_comm_init + phos_cuda_film_reduce: one ncclReduce per frame).  Needs a box
with >= 2 GPUs (gpurun --gpus 2); on a one-GPU box the test has nothing to launch and says so (the one-rank
communicator is covered by test_gpu_render.py, the host-side partition logic by test_frame_host.py over gloo)."""
import os
import subprocess
import sys
import textwrap

import pytest

from phosphorus_mk2_b200 import lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = """
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from phosphorus_mk2_b200 import scenes
from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, make_tiles
from phosphorus_mk2_b200.frame import render_partition, samples_of_rank, tiles_of_rank
dist.init_process_group("gloo")  # only ships the 128-byte NCCL id and the barriers; the film moves through the library
r, n = dist.get_rank(), dist.get_world_size()
sc = scenes.cornell_box(128, 96)
acc = Accel(sc)
dev = CudaDevice.make(Options(8, 1, 4), int(os.environ["LOCAL_RANK"]))
dev.preprocess(sc, acc)
dev.upload_scene(sc)
dev.comm_init(dist)
tiles = make_tiles(128, 96)
whole = None
if r == 0:
    render_partition(dev, tiles, (0, 8), 8, seed=3)
    whole = dev.film_read()
dist.barrier()
# tile-partitioned: the reduce is a gather, bit-exact
render_partition(dev, tiles_of_rank(tiles, r, n), (0, 8), 8, seed=3)
dev.film_reduce(0)
got = dev.film_read()
if r == 0:
    assert np.array_equal(got, whole), "tile-partitioned frame differs"
dist.barrier()
# sample-partitioned (more ranks than samples on purpose when n > 8: empty shares contribute zeros)
render_partition(dev, tiles, samples_of_rank(8, r, n), 8, seed=3)
dev.film_reduce(0)
got = dev.film_read()
if r == 0:
    assert np.allclose(got[..., :3], whole[..., :3], rtol=2e-6, atol=1e-7), "sample-partitioned frame differs"
    assert (got[..., 3] == 1.0).all()
dist.barrier()
dev.close()
dist.destroy_process_group()
print("rank%d-ok" % r, flush=True)
"""


def test_partitioned_frame_with_the_library_film_reduce(tmp_path):
    n_gpu = lib.load().phos_cuda_device_count()
    assert n_gpu >= 1
    if n_gpu < 2:
        print("one GPU visible: the N-rank film reduce needs gpurun --gpus 2")
        return
    world = min(n_gpu, 4)
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(WORKER.format(root=ROOT)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    for r in range(world):
        assert f"rank{r}-ok" in out.stdout
