"""GPU, >= 2 devices (skipped on a single-GPU box): the sharded inferer over NCCL - one process per GPU - must return the
single-process result for every sharding mode, with device and host outputs, labels included."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from waveformer_b200.inferers import SlidingWindowInferer
from helpers import seeded_randn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
wconv = (seeded_randn((3, 2, 3, 3, 3), 600) * 0.2).to(dev)
net = lambda p: torch.nn.functional.conv3d(p.float(), wconv, padding=1)
kw = dict(roi_size=(16, 16, 16), sw_batch_size=2, overlap=0.5, mode="gaussian", compute_dtype=torch.float32,
          channels_last=False, return_labels=True)
for case, shape, shard in (("split-volume", (1, 2, 40, 36, 30), True), ("three-volumes", (3, 2, 24, 36, 30), True),
                           ("whole-volumes", (world, 2, 24, 20, 30), True), ("interleaved", (3, 2, 24, 36, 30), "interleaved")):
    x = seeded_randn(shape, 77)
    want_inf = SlidingWindowInferer(shard=False, **kw)
    want = want_inf(x.to(dev), net)
    want_labels = want_inf.labels
    for host in (False, True):
        inf = SlidingWindowInferer(shard=shard, device="cpu" if host else None, **kw)
        y = inf(x.pin_memory() if host else x.to(dev), net)
        owned = inf.owned_volumes
        got = [None] * world
        dist.all_gather_object(got, owned)
        assert sorted(v for o in got for v in o) == list(range(shape[0])), got
        if owned:
            assert y.is_cuda != host
            err = float((y.to(dev) - want[owned]).abs().max() / want.abs().max())
            assert err < 1e-5, (case, host, err)
            flips = float((inf.labels != want_labels[owned]).float().mean())
            assert flips < 1e-3, (case, host, flips)
        else:
            assert y is None
    print("OK", case, rank, owned, flush=True)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_sharded_inferer_nccl(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    world = 2
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(script), ROOT]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("OK") == 4 * world
