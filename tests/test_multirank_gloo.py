"""N > 1 host logic on CPU: two gloo ranks each produce their round-robin rows (rendered here by
the oracle with the keyed policy, since there is no GPU), rank 0 gathers on the host, and the
assembled frame must equal the single-rank render bit for bit."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from pt_three_ways_b200 import partition, scenefile, capi
from oracle import oracle_binding as ob

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scene = scenefile.load(os.path.join(%(root)r, "tests/golden/scenes/cornell.ptscene"))
w, h, spp = 20, 15, 2   # odd height: ranks own different numbers of rows
cam = scene.camera(w, h)
res = ob.OracleScene(scene).render(cam, ob.params_array(w, h, spp=spp, seed=3), ob.RNG_KEYED_PHILOX,
                                   row_begin=rank, row_step=world)
local = np.zeros((h, w), dtype=capi.PIXEL_DTYPE)
local["sum"] = res["sums"]
local["n"] = res["counts"]
frame = partition.gather_rows(local, rank, world)
if rank == 0:
    np.save(%(out)r, frame)
# the copy-free gather bench.py uses: every rank's rows land in ONE shared host frame
shared = partition.SharedFrame(h, w, capi.PIXEL_DTYPE, "gloo_test_%%d" %% os.getppid(), rank, dist.barrier)
shared.frame[rank::world] = local[rank::world]   # what ptb200_render's strided D2H does
assembled = shared.gathered()
if rank == 0:
    np.save(%(out)r + ".shared.npy", np.array(assembled))
shared.close()
dist.barrier()
dist.destroy_process_group()
"""


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_row_partition_and_host_gather(tmp_path, scenes, oracle, capi):
    out = str(tmp_path / "frame.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, out=out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stderr[-2000:]
    frame = np.load(out)
    scene = scenes["cornell"]
    whole = oracle.OracleScene(scene).render(scene.camera(20, 15), oracle.params_array(20, 15, spp=2, seed=3),
                                             oracle.RNG_KEYED_PHILOX)
    assert np.array_equal(frame["sum"], whole["sums"])
    assert np.array_equal(frame["n"], whole["counts"])
    shared = np.load(out + ".shared.npy")
    assert np.array_equal(shared["sum"], whole["sums"]) and np.array_equal(shared["n"], whole["counts"])


def test_rows_of_rank_cover_the_frame_once():
    from pt_three_ways_b200 import partition
    for height in (1, 7, 480, 1080):
        for world in (1, 2, 4, 8):
            seen = sorted(y for r in range(world) for y in partition.rows_of_rank(height, r, world))
            assert seen == list(range(height))
