"""apply_method="distributed_gpu" on every visible GPU against the CPU oracle (needs >= 2 GPUs: -m gpu on a multi-GPU
box).  One process per GPU is launched with torch.distributed.run; every rank compares ITS chunk of the amplitudes,
the allreduced norm and <Z_q>, and the distributed sampler's indices (bit for bit) with the oracle
(scripts/dist_check.py).  On a box with one GPU the test skips WITH the reason -- `bench.py --gpus N` runs the same
comparison before its timed region and prints it as `parity` -- on a box with more it runs and must pass."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world():
    from damavand_b200 import _lib
    n = _lib.load().dvd_device_count()
    w = 1
    while 2 * w <= n:
        w *= 2
    return w


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("mode", ["fused_remap_jit", "fused_remap", "fused_remap_plain_schedule", "in_place_peer_swap", "staged_nccl"])
def test_distributed_gpu_against_oracle_on_all_visible_gpus(mode):
    world = _world()
    if world < 2:
        pytest.skip("one visible GPU: the multi-rank path needs at least two (bench.py --gpus N carries the same oracle "
                    "comparison in its `parity` field)")
    env = dict(os.environ)
    env.update({"DIST_CHECK_N": "14,20,22", "DIST_CHECK_QFT_MAX": "20", "DIST_CHECK_JIT": "1" if mode == "fused_remap_jit" else "0", "DVD_JIT_MIN_QUBITS": "12"})
    if mode == "fused_remap_plain_schedule":
        env.update({"DVD_STORE_REMAP": "0", "DVD_DEFER_TAILS": "0"})   # no tail deferral, the layout restore in a pass of its own
    if mode == "in_place_peer_swap":
        env["DVD_FUSED_REMAP"] = "0"            # every global<->local swap as a k_swap_peer exchange of its own
    if mode == "staged_nccl":
        env["DVD_SWAP"] = "nccl"                # ncclSend/ncclRecv through staging buffers (the path across nodes)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    tail = (res.stdout + res.stderr)[-4000:]
    assert res.returncode == 0 and "DIST_CHECK PASS" in res.stdout, tail
    lines = [l for l in res.stdout.splitlines() if l.startswith("n=")]
    assert len(lines) == 10 and all("all_ranks_ok=True" in l for l in lines), tail
    if mode.startswith("fused_remap"):
        assert any("fused_remap_passes=" in l and "fused_remap_passes=0 " not in l for l in lines), tail
        # the swaps that restore the layout ride on the store of the last gate pass in some of the cases (remote writes)
        assert any("store_side=" in l and "store_side=0 " not in l for l in lines) == (mode != "fused_remap_plain_schedule"), tail
    else:
        assert all("fused_remap_passes=0 " in l for l in lines), tail
