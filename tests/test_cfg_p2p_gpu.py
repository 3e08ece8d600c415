"""CFG-branch exchange through peer memory (`cfg_ddim_update_p2p`, parallel.CfgPeerExchange).

The kernel waits on a flag its partner writes; a partner that never arrives ends in a device trap after ~20 s, so both
checks run in child processes (a trap must not poison this session's CUDA context)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cfg_ddim_update_p2p_loopback_bit_exact():
    """Two streams of one GPU play the two ranks: every step bit-identical to cfg_ddim_update on `[e_u; e_c]`."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cfg_p2p_loopback.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


@pytest.mark.gpu
def test_cfg_branch_split_two_gpus_p2p_matches_nccl():
    """2 GPUs: the peer-memory transport gives the same latents, bit for bit, as the NCCL all-gather transport, on both ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs of one NVLink domain")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tools", "cfg_branch_split_check.py"), "--steps", "6", "--reps", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "p2p == nccl bit for bit: True" in r.stdout
