"""Build + ctypes binding of libpyro2b200.so (the C ABI declared in include/pyro2b200.h).

The library is built IN-TREE (pyro2_b200/csrc/libpyro2b200.so) with
``nvcc -gencode arch=compute_100a,code=sm_100a`` -- sm_100a only, no fallbacks.  There is no CPU
path: if the shared object is missing or CUDA is unavailable the product raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# P2B_SO overrides the library path (A/B timing of kernel variants during development)
SO_PATH = os.environ.get("P2B_SO") or os.path.join(CSRC, "libpyro2b200.so")
SOURCES = ["ghost_cfl.cu", "sweep.cu", "mg.cu", "flow.cu", "bc_user.cu", "lm.cu", "slab_comm.cu"]
HEADERS = ["common.cuh", "hydro_core.cuh", "sweep_task.cuh", "sweep_args.cuh", "mg_kernels.cuh", "peer_comm.cuh", "flow_kernels.cuh", "bc_user_kernels.cuh", "lm_kernels.cuh", "../../include/pyro2b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--shared"]


def source_hash():
    """sha256 over the compiler flags and every source / header of the library: the staleness criterion (mtimes do
    not survive a snapshot to the GPU box, and the .so is git-ignored but shipped in-tree)"""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in SOURCES + HEADERS:
        path = os.path.join(CSRC, name)
        if os.path.exists(path):
            h.update(name.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def built_hash():
    """the source hash recorded when the in-tree .so was built (None: no record)"""
    try:
        with open(SO_PATH + ".hash") as fh:
            return fh.read().strip()
    except OSError:
        return None


def _stale():
    return not os.path.exists(SO_PATH) or built_hash() != source_hash()


def build(force=False, verbose=False):
    """compile every CUDA source for sm_100a into the in-tree shared object"""
    if not force and not _stale():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    tmp = SO_PATH + f".{os.getpid()}.tmp"
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + f".{os.getpid()}.o")
        res = subprocess.run([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src],
                             capture_output=True, text=True)
        return obj, res

    # one nvcc per translation unit, in parallel (the sweep and multigrid units dominate), then one link
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    log = "".join(r.stdout + r.stderr for _, r in results)
    if any(r.returncode != 0 for _, r in results):
        raise RuntimeError("nvcc failed:\n" + log)
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", tmp] + [o for o, _ in results], capture_output=True, text=True)
    for o, _ in results:
        os.remove(o)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    res.stderr = log + res.stderr
    os.replace(tmp, SO_PATH)
    with open(SO_PATH + ".hash", "w") as fh:
        fh.write(source_hash() + "\n")
    if verbose:
        print(res.stderr)
    return SO_PATH


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("ng", C.c_int), ("pitch", C.c_int),
                ("plane_stride", C.c_longlong), ("dx", C.c_double), ("dy", C.c_double)]


class CompParams(C.Structure):
    _fields_ = [("gamma", C.c_double), ("z0", C.c_double), ("z1", C.c_double), ("delta", C.c_double),
                ("cvisc", C.c_double), ("limiter", C.c_int), ("use_flattening", C.c_int),
                ("no_avisc_xhi", C.c_int), ("no_avisc_yhi", C.c_int),
                ("grav", C.c_double), ("src_flip_ylo", C.c_int), ("src_flip_yhi", C.c_int),
                ("riemann", C.c_int), ("xl_solid", C.c_int), ("yl_solid", C.c_int),
                ("heat_rate", C.c_double), ("heat_profile", C.c_void_p), ("do_sponge", C.c_int),
                ("sponge_rho_begin", C.c_double), ("sponge_rho_full", C.c_double), ("sponge_timescale", C.c_double),
                ("src_copy_yhi", C.c_int),
                ("geo_i", C.c_void_p), ("geo_j", C.c_void_p), ("geo_ni", C.c_int), ("geo_nj", C.c_int),
                ("src_flip_xlo", C.c_int), ("src_flip_xhi", C.c_int)]


BC_CODES = {"outflow": 0, "neumann": 0, "reflect-even": 1, "reflect-odd": 2, "dirichlet": 2,
            "periodic": 3, None: 4, "none": 4}

_lib = None

# every symbol include/pyro2b200.h declares: (name, restype, argtypes)
_vp, _i, _d, _ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
_PG = C.POINTER(Grid)
SIGNATURES = {
    "p2b_last_error": (C.c_char_p, []),
    "p2b_version": (_i, []),
    "p2b_device_sms": (_i, []),
    "p2b_fill_ghost_f64": (_i, [_vp, _PG, _i, C.POINTER(_i), _vp]),
    "p2b_fill_ghost_i64": (_i, [_vp, _PG, _i, C.POINTER(_i), _vp]),
    "p2b_fill_ghost_values_f64": (_i, [_vp, _PG, C.POINTER(_i), _vp, _vp, _vp, _vp, _vp]),
    "p2b_fill_hse_f64": (_i, [_vp, _PG, _d, _d, _i, _i, _vp]),
    "p2b_fill_ambient_f64": (_i, [_vp, _PG, _i, _i, _d, _vp]),
    "p2b_cfl_wavemax": (_i, [_vp, _PG, _d, _vp, _vp]),
    "p2b_compressible_sweep": (_i, [_vp, _vp, _PG, C.POINTER(CompParams), _d, _vp, _vp]),
    "p2b_sweep_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "p2b_test_fastmath": (_i, [_i, _vp, _vp, _vp, _i, _vp]),
    "p2b_sweep_uses_tensor_map": (_i, []),
    "p2b_mg_create": (_vp, [_i, C.POINTER(_i), _d, _d, _d, _d, _d, _d, _i, _i]),
    "p2b_mg_create_slab": (_vp, [_i, C.POINTER(_i), _d, _d, _d, _d, _d, _d, _i, _i, _i, _i, _i]),
    "p2b_mg_level_info": (_i, [_vp, _i, C.POINTER(_ll)]),
    "p2b_mg_destroy": (_i, [_vp]),
    "p2b_mg_set_blocking": (_i, [_vp, _i]),
    "p2b_mg_nlevels": (_i, [_vp]),
    "p2b_mg_workspace_bytes": (_ll, [_vp]),
    "p2b_mg_bind": (_i, [_vp, _vp, _ll]),
    "p2b_mg_level_ptr": (_vp, [_vp, _i, _i]),
    "p2b_mg_level_pitch": (_i, [_vp, _i]),
    "p2b_mg_set_bc_values": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "p2b_mg_smooth": (_i, [_vp, _i, _i, _vp]),
    "p2b_mg_residual": (_i, [_vp, _i, _vp]),
    "p2b_mg_restrict": (_i, [_vp, _i, _vp]),
    "p2b_mg_prolong_correct": (_i, [_vp, _i, _vp]),
    "p2b_mg_fill_bc": (_i, [_vp, _i, _vp]),
    "p2b_mg_zero_coarse": (_i, [_vp, _vp]),
    "p2b_mg_vcycle": (_i, [_vp, _vp]),
    "p2b_mg_vcycle_level": (_i, [_vp, _i, _vp]),
    "p2b_mg_tb_pass": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "p2b_mg_tb_halo": (_i, []),
    "p2b_mg_tb_iters": (_i, []),
    "p2b_mg_norm2": (_i, [_vp, _i, _i, _vp, _vp]),
    "p2b_mg_cycle_diagnostics": (_i, [_vp, _vp, _vp, _vp]),
    "p2b_mg_set_operator": (_i, [_vp, _d, _d]),
    "p2b_mg_set_peers": (_i, [_vp, C.POINTER(_vp)]),
    "p2b_mg_exchange": (_i, [_vp, _i, _i, _i, _vp]),
    "p2b_mg_set_stop": (_i, [_vp, _i, _d, _d, _i, _vp]),
    "p2b_mg_result": (_i, [_vp, C.POINTER(_d), C.POINTER(_ll), _vp]),
    "p2b_mg_control_ptr": (_vp, [_vp]),
    "p2b_slab_ctl_bytes": (_ll, []),
    "p2b_slab_create": (_vp, [_i, _i, _i, C.POINTER(_vp)]),
    "p2b_slab_destroy": (_i, [_vp]),
    "p2b_slab_register": (_i, [_vp, _i, _vp, _ll, C.POINTER(_vp)]),
    "p2b_slab_owns": (_i, [_vp, _vp]),
    "p2b_slab_exchange": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _vp]),
    "p2b_slab_allreduce_max4": (_i, [_vp, _vp, _vp]),
    "p2b_slab_error": (_i, [_vp, _vp]),
    "p2b_shared_alloc": (_vp, [_ll]),
    "p2b_shared_free": (_i, [_vp]),
    "p2b_shared_handle": (_i, [_vp, C.c_char_p]),
    "p2b_shared_open": (_vp, [C.c_char_p]),
    "p2b_shared_close": (_i, [_vp]),
    "p2b_mg_cn_rhs": (_i, [_vp, _vp, _i, _d, _vp]),
    "p2b_mg_coeff_workspace_bytes": (_ll, [_vp]),
    "p2b_mg_set_coeffs": (_i, [_vp, _vp, _ll, _vp, _i, C.POINTER(_i), _vp]),
    "p2b_mg_coeff_ptr": (_vp, [_vp, _i, _i]),
    "p2b_flow_create": (_vp, [_PG]),
    "p2b_flow_destroy": (_i, [_vp]),
    "p2b_flow_workspace_bytes": (_ll, [_vp]),
    "p2b_flow_bind": (_i, [_vp, _vp, _ll]),
    "p2b_flow_plane": (_vp, [_vp, _i]),
    "p2b_flow_interface_states": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "p2b_flow_mac_vels": (_i, [_vp, _vp]),
    "p2b_flow_mac_divergence": (_i, [_vp, _vp, _i, _vp]),
    "p2b_flow_mac_project": (_i, [_vp, _vp, _vp]),
    "p2b_flow_upwind_states": (_i, [_vp, _vp]),
    "p2b_flow_advect_update": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "p2b_flow_cc_divergence": (_i, [_vp, _vp, _vp, _vp, _i, _d, _i, _vp]),
    "p2b_flow_project": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "p2b_flow_burgers_update": (_i, [_vp, _vp, _vp, _d, _vp]),
    "p2b_flow_advection_update": (_i, [_vp, _vp, _d, _d, _d, _i, _vp]),
    "p2b_flow_maxabs": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "p2b_lm_create": (_vp, [_PG, _vp]),
    "p2b_lm_destroy": (_i, [_vp]),
    "p2b_lm_workspace_bytes": (_ll, [_vp]),
    "p2b_lm_bind": (_i, [_vp, _vp, _ll]),
    "p2b_lm_plane": (_vp, [_vp, _i]),
    "p2b_lm_coeff": (_i, [_vp, _vp, _vp, _d, _i, _i, _vp]),
    "p2b_lm_source": (_i, [_vp, _vp, _vp, _d, _vp]),
    "p2b_lm_interface_states": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "p2b_lm_mac_vels": (_i, [_vp, _vp]),
    "p2b_lm_mac_divergence": (_i, [_vp, _vp, _i, _vp]),
    "p2b_lm_mac_project": (_i, [_vp, _vp, _vp]),
    "p2b_lm_density_update": (_i, [_vp, _vp, _vp, _d, _i, _d, _vp]),
    "p2b_lm_upwind_states": (_i, [_vp, _vp]),
    "p2b_lm_advect_update": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "p2b_lm_add_source": (_i, [_vp, _vp, _d, _vp]),
    "p2b_lm_cc_divergence": (_i, [_vp, _vp, _vp, _vp, _i, _d, _i, _vp]),
    "p2b_lm_project": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp]),
    "p2b_lm_reduce": (_i, [_vp, _vp, _vp, _vp, _d, _vp, _vp]),
}


def lib():
    """the loaded C-ABI library; raises (never falls back) when it cannot be loaded"""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "pyro2_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)   # AttributeError here = header/library mismatch
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class P2BError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib().p2b_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        raise P2BError(f"libpyro2b200 error {rc}: {msg}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bc_array(rows):
    """rows: iterable of 4-tuples of BC names -> flat ctypes int array"""
    flat = [BC_CODES[b] for r in rows for b in r]
    return (C.c_int * len(flat))(*flat)
