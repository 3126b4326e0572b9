"""N>1 path on CPU: world_size-2 and -3 gloo runs of the slab protocol (chrono_b200/slab.py) with an oracle-backed
backend (tests/slab_gloo_worker.py), compared with a single-process oracle run of the same scene."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), os.path.join(HERE, "slab_gloo_worker.py")] + list(args)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SLAB GLOO PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world,lag", [(2, 0), (2, 1), (3, 1)])
def test_slab_protocol_gloo(world, lag):
    launch(world, "--lag", str(lag), "--steps", "60" if world == 2 else "170")


def test_slab_bounds_equal_counts():
    from chrono_b200 import slab
    x = np.random.default_rng(0).uniform(-3, 5, size=10001)
    for world in (1, 2, 4, 8):
        b = slab.slab_bounds(x, world)
        assert len(b) == world + 1 and b[0] == -np.inf and b[-1] == np.inf
        cnt = [((x >= b[r]) & (x < b[r + 1])).sum() for r in range(world)]
        assert sum(cnt) == len(x) and max(cnt) - min(cnt) <= 1
