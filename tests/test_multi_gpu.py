"""Multi-GPU parity (-m gpu, needs >= 2 GPUs on the box): z-slab runs over NCCL == the oracle on the whole box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("case", ["p8", "pwl", "p8_thin", "p8_nofuse", "p8_tall", "pwl_tall", "p8_tall_serial", "pwl_serial_thin", "user_tall", "user_nofuse"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_decomposition_matches_oracle(world, case):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29400 + world), os.path.join(HERE, "mgpu_worker.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])  # (the "multi-gpu parity ok ..." line of rank 0)
    print("\n".join(ln for ln in r.stderr.splitlines() if "OMP_NUM_THREADS" not in ln and "****" not in ln)[-3000:])
    assert r.returncode == 0
    assert "multi-gpu parity ok" in r.stdout
