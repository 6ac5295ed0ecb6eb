"""Two ranks on two GPUs: the peer-memory (CUDA IPC + copy engine) all-gather of ShardedTiles must give every rank the
same bytes as ncclAllGather, in the blocking and in the pipelined form.  Skipped on a one-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_p2p_allgather_matches_nccl_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "p2p_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "P2P_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"impl": ["p2p", "nccl"]' in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_peer_export_and_copy_on_one_device(mw):
    """The plumbing a one-GPU box can exercise: the exported offset follows the pointer, and mw_peer_copy is an
    asynchronous device-to-device copy on the given stream."""
    import torch
    a = torch.arange(1 << 20, device="cuda", dtype=torch.float32)
    b = torch.zeros_like(a)
    handle, off = mw.native.peer_export(a.data_ptr())
    assert len(handle) == mw.native.MW_PEER_HANDLE_BYTES and off >= 0
    _, off2 = mw.native.peer_export(a[1000:].data_ptr())
    assert off2 == off + 4000
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    mw.native.peer_copy(b.data_ptr(), a.data_ptr(), a.numel() * 4, s.cuda_stream)
    s.synchronize()
    assert torch.equal(a, b)
    with pytest.raises(mw.native.MwError):
        mw.native.peer_export(0)
