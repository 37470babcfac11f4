"""Multi-GPU parity (needs >= 2 GPUs on the box, otherwise skipped): the m-sharded / latitude-band-sharded
transform against the CPU oracle, with the peer-memory exchange (NVLink stores fused into the Legendre epilogue,
device-side barrier) and with the NCCL all-to-all."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("gridname,T,nf", [("O48", 47, 5), ("O160", 159, 9)])
def test_sharded_transform_two_gpus(gridname, T, nf, exchange):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29581", os.path.join(REPO, "tests", "dist_check.py"), gridname, str(T), str(nf), exchange]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "-> OK" in r.stdout, r.stdout[-3000:]
