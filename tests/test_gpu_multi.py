"""GPU, >= 2 devices: x-slab decomposition with NCCL halo exchange reproduces the single-GPU run
bit for bit (state and every dt).  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("problem,nx,ny,nsteps", [("sedov", 256, 128, 30), ("kh", 128, 64, 20), ("quad", 192, 96, 20)])
def test_decomposed_run_is_bit_identical(problem, nx, ny, nsteps):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(HERE, "multi_gpu_worker.py"), problem, str(nx), str(ny), str(nsteps)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "bit_identical=True" in res.stdout and "dt_identical=True" in res.stdout


@pytest.mark.parametrize("kind,n,split", [("dirichlet", 1024, 256), ("periodic", 512, 256), ("mixed", 1024, 512),
                                          ("xper_inhom", 512, 256)])
def test_decomposed_multigrid_is_bit_identical(kind, n, split):
    """x-slab multigrid (NCCL deep-halo exchange per blocked-smoother pass, replicated coarse levels)
    vs the single-GPU solve: identical solution bits and cycle counts"""
    ngpu = _ngpu()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29612",
           os.path.join(HERE, "multi_gpu_mg_worker.py"), kind, str(n), str(split)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "bit_identical=True" in res.stdout


@pytest.mark.parametrize("solver,problem,nx,ny,nsteps,extra", [
    ("advection", "smooth", 256, 256, 20, []), ("burgers", "test", 256, 128, 20, []),
    ("compressible", "bubble", 128, 256, 10, ["mesh.ymax=4.0", "mesh.ylboundary=hse", "mesh.yrboundary=hse"]),
    ("compressible", "advect", 128, 96, 10, ["mesh.grid_type=SphericalPolar", "mesh.xmin=1.0", "mesh.xmax=2.0", "mesh.ymin=0.523",
                                             "mesh.ymax=2.617", "mesh.xlboundary=reflect", "mesh.xrboundary=outflow",
                                             "mesh.ylboundary=outflow", "mesh.yrboundary=outflow", "compressible.riemann=CGF",
                                             "driver.fix_dt=-1.0"]),
    ("lm_atm", "bubble", 128, 128, 3, []),
    ("diffusion", "gaussian", 256, 256, 5, ["diffusion.mg_split_n=128"]),
    ("incompressible", "shear", 256, 256, 3, ["incompressible.mg_split_n=128"])])
def test_decomposed_flow_solvers_are_bit_identical(solver, problem, nx, ny, nsteps, extra):
    """advection / burgers (explicit stages on x-slabs, all-reduced dt) and the incompressible solver (plus two
    x-slab multigrid projections per step) against the single-GPU run"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(HERE, "multi_gpu_flow_worker.py"), solver, problem, str(nx), str(ny), str(nsteps)] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "bit_identical=True" in res.stdout and "dt_identical=True" in res.stdout
