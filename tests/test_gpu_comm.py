"""C-ABI multi-GPU entry points (gw_comm_*, gw_*_sharded): NCCL bound at run time, results identical to one GPU.
One rank runs on any GPU box; the two-rank case needs two GPUs (`gpurun --gpus 2`) and is skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

from graph_embedding_b200 import _lib  # noqa: E402

WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "comm_worker.py")


def _run(nranks, tmp_path):
    idfile, out = str(tmp_path / "nccl_id"), str(tmp_path / "done")
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(nranks), idfile, out], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(nranks)]
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for k in procs:
                k.kill()
            raise
        logs.append(o)
    for r, p in enumerate(procs):
        assert p.returncode == 0 and open("%s.%d" % (out, r)).read() == "ok", "rank %d:\n%s" % (r, logs[r][-3000:])


def test_single_rank_communicator(tmp_path):
    _run(1, tmp_path)


def test_two_ranks_gather_equals_one_gpu(tmp_path):
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, tmp_path)
