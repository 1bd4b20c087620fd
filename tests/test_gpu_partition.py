"""GPU: the mesh-partitioned driver through torchrun.  On a 1-GPU box this runs world_size 1 (the
partition machinery with a single part: sealed ghost rows, owner-masked dots, device-resident scalars,
host-driven Krylov loop); with >= 2 visible GPUs it runs world_size 2 over NCCL, where the default is
the FUSED path (peer-memory halo pushes and cross-GPU barriers inside the persistent Krylov kernels,
replicated P1 solves) and FX_DIST_FUSED=0 selects the host-driven cross-check."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cfg,port,fused", [("cfg2", 29611, "1"), ("cfg3", 29614, "1"), ("cfg4", 29612, "1"),
                                            ("cfg5", 29613, "1"), ("cfg6", 29616, "1"), ("cfg2", 29615, "0")])
def test_partitioned_matches_single_gpu(cfg, port, fused):
    """fields, energies and the boundary functionals (journal torque / load, Nusselt number) of the
    partitioned drivers against the single-GPU drivers."""
    world = 2 if torch.cuda.device_count() >= 2 else 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"),
           "10", "2", "0", cfg]
    if fused == "0" and world == 1:
        pytest.skip("world 1 always runs the host-driven loop")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, FX_DIST_FUSED=fused))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["world"] == world and out["max_rel_err"] < 1e-8
