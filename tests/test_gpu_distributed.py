"""N > 1 with the CUDA engine: two ranks launched through torch.distributed.run.

  * gloo plumbing, both ranks on cuda:0 -- runs on the one-GPU box of the driver: broadcast of A, product-balanced row
    blocks, the resident-block power chain on the GPU, the gathered powers bit-identical to the oracle;
  * NCCL below the C ABI (b200_comm_*), one GPU per rank -- needs two GPUs, skipped otherwise: device broadcast,
    device all-gather of the row blocks, the multi-GPU squaring chain (power_until_stable).
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _launch(mode, nproc=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(nproc):
        assert f"RANK {k} OK" in r.stdout, r.stdout[-2000:]


def test_two_ranks_cuda_engine_gloo_plumbing(gpu_ctx):
    _launch("gloo")


def test_two_ranks_nccl_below_the_abi(gpu_ctx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device)")
    _launch("nccl")


def test_compiled_host_bench_reproduces_the_readme_nnz(gpu_ctx):
    """tools/b200_bench.cpp (C++ above the C ABI, no Python): the A^2..A^7 chain of bench_repeated_exponentiation on a device-built
    operand; its own check compares nnz per power with the README column (README.md:42-47)."""
    exe = os.path.join(ROOT, "sparse_linear_algebra_tests_b200", "csrc", "b200_bench")
    assert os.path.exists(exe), "b200_bench is not built (make -C sparse_linear_algebra_tests_b200/csrc)"
    r = subprocess.run([exe, "--config", "torus30", "--iters", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "# nnz check OK" in r.stdout
    assert "torus30,1,A^7,11736555,23765080" in r.stdout
