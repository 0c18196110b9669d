"""Multi-rank correctness under pytest: N ranks x B users per step == 1 rank x N*B users per step.

Self-spawns `torch.distributed.run` with tools/mg_check.py when the box shows >= 2 GPUs (skips on a 1-GPU box): both
exchange formulations -- the NVLink peer-memory kernels (default; their ordering rests on "the NCCL all-reduce of the
step completes => every outbox is written") and the NCCL all-to-all path -- against the single-GPU union-batch step."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(script, world, env_extra, timeout=600):
    env = dict(os.environ); env.update(env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", script)]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return r.returncode, r.stdout


@pytest.mark.parametrize("peer", [1, 0])
def test_two_ranks_equal_one_rank_union_batch(peer):
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    rc, out = _run("mg_check.py", 2, {"POI_MG_PEER": str(peer)})
    assert rc == 0 and "MG_CHECK PASS world 2" in out, out[-3000:]


def test_all_visible_gpus_peer_exchange():
    n = _n_gpus()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    rc, out = _run("mg_check.py", n, {"POI_MG_PEER": "1"})
    assert rc == 0 and ("MG_CHECK PASS world %d" % n) in out, out[-3000:]


def test_geoie_row_sharded_two_ranks_equal_one_rank_union_batch():
    """csrc/mf_mg.cuh: 2 ranks x Bu users == 1 rank x 2 Bu users for the K-negative GeoIE mini-batch step."""
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    rc, out = _run("mg_check_geoie.py", 2, {})
    assert rc == 0 and "MG_CHECK_GEOIE PASS world 2" in out, out[-3000:]
