"""GPU (needs >= 2 B200s on the box, skipped otherwise): Ulysses sequence-parallel forward == single-GPU forward.
Spawns one process per GPU with torchrun (tools/sp_check.py) for both transports."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sequence_parallel_matches_single_gpu(world, mode):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, SP_MODE=str(mode))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world + 10 * mode),
                        os.path.join(ROOT, "tools", "sp_check.py")], capture_output=True, text=True, timeout=600, env=env)
    # ranks share one stdout pipe: two records can land on one line, so records are decoded wherever the marker appears
    dec, lines, pos = json.JSONDecoder(), [], 0
    while (pos := r.stdout.find("SP_CHECK ", pos)) >= 0:
        obj, end = dec.raw_decode(r.stdout, pos + len("SP_CHECK "))
        lines.append(obj)
        pos = end
    assert r.returncode == 0 and len(lines) == world and all(l["ok"] for l in lines), (r.stdout[-2000:], r.stderr[-2000:])
