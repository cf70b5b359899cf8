"""N > 1 GPUs: the strip-decomposed run (one process per GPU) is bitwise the single-GPU run (for the block-local EVP
preconditioner: bitwise the oracle in its REPRODUCIBLE build on the same 1 x P blocks).
Needs >= 2 GPUs (gpurun --gpus 2); skipped on the single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("which", ["pcsi", "pcsi22", "pcsi_plain", "chrongear", "cyclic_pcg", "gm", "pbc", "lwlim", "coupled", "evp"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_strips_are_bitwise_the_single_strip_run(which, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "mr_worker.py"),
           which]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    lines = r.stdout.splitlines()
    tail = "\n".join([l for l in lines if "FAIL" in l or "MULTIRANK" in l or "Error" in l][:20] + lines[-8:])
    assert r.returncode == 0 and "OK" in r.stdout, tail
