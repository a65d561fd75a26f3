"""GPU, multi-process: the sharded paths with REAL ranks (one process per GPU, NCCL) -- the fused band exchange
(bsplat_render_enqueue_band_p2p: peer stores from inside the rasterizer) included.  Needs >= 2 GPUs on the box
(skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multiproc.py -m gpu`).  The host logic of
the same paths is covered on CPU by tests/test_parallel_cpu.py (gloo, world size 2)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_row_bands_and_view_split_multi_process(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "mp_band_worker.py")]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MP_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
