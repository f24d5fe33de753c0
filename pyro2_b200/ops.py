"""Thin torch <-> C-ABI glue: device buffers are torch.cuda tensors, the library sees raw pointers.

State layout (DESIGN.md "Data layout in HBM"): ``planes[n, i, j]`` float64, ``planes.stride() ==
(plane_stride, pitch, 1)`` with ``pitch`` a multiple of 16 doubles so every row is 128-byte aligned;
the reference's ``data[i, j, n]`` view is ``planes.permute(1, 2, 0)[:, :qy, :]``.
"""
import ctypes as C

import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pyro2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def row_pitch(qy):
    """row pitch in elements: multiple of 16 doubles (128 B) when the row is long enough"""
    return (qy + 15) // 16 * 16 if qy >= 16 else (qy + 1) // 2 * 2


def alloc_planes(nvar, qx, qy, dtype=torch.float64, device=None):
    """zero-initialised SoA storage; returns the padded tensor (nvar, qx, pitch)"""
    require_cuda()
    return torch.zeros((nvar, qx, row_pitch(qy)), dtype=dtype, device=device or "cuda")


def grid_struct(planes, nx, ny, ng, dx=1.0, dy=1.0):
    assert planes.stride(-1) == 1
    if planes.dim() == 3:
        plane_stride, pitch = planes.stride(0), planes.stride(1)
    else:
        plane_stride, pitch = 0, planes.stride(0)
    return _lib.Grid(nx, ny, ng, pitch, plane_stride, dx, dy)


def fill_ghost(planes, nx, ny, ng, bcs):
    """planes: (nvar, qx, pitch) float64 or int64; bcs: nvar 4-tuples of BC names"""
    require_cuda()
    g = grid_struct(planes, nx, ny, ng)
    arr = _lib.bc_array(bcs)
    nvar = planes.shape[0]
    assert len(arr) == 4 * nvar
    if planes.dtype == torch.float64:
        f = _lib.lib().p2b_fill_ghost_f64
    elif planes.dtype == torch.int64:
        f = _lib.lib().p2b_fill_ghost_i64
    else:
        raise TypeError(f"ghost fill supports float64 and int64, not {planes.dtype}")
    _lib.check(f(planes.data_ptr(), C.byref(g), nvar, arr, _lib.stream_ptr()))


def fill_ghost_values(plane, nx, ny, ng, bc, dx, dy, xl=None, xr=None, yl=None, yr=None):
    """single (qx, pitch) plane with optional inhomogeneous boundary-value tensors"""
    require_cuda()
    g = grid_struct(plane, nx, ny, ng, dx, dy)
    codes = (C.c_int * 4)(*[_lib.BC_CODES[b] for b in bc])

    def p(t, code):
        return None if (t is None or code not in (0, 2)) else t.data_ptr()
    _lib.check(_lib.lib().p2b_fill_ghost_values_f64(plane.data_ptr(), C.byref(g), codes, p(xl, codes[0]),
                                                    p(xr, codes[1]), p(yl, codes[2]), p(yr, codes[3]),
                                                    _lib.stream_ptr()))


def fill_hse(planes, nx, ny, ng, dy, grav, gamma, var, side):
    """the compressible "hse" boundary for plane `var` of the 4-plane state on side 0 (ylb) / 1 (yrb)"""
    require_cuda()
    assert planes.dim() == 3 and planes.shape[0] >= 4 and planes.dtype == torch.float64
    g = grid_struct(planes, nx, ny, ng, 1.0, dy)
    _lib.check(_lib.lib().p2b_fill_hse_f64(planes.data_ptr(), C.byref(g), grav, gamma, var, side, _lib.stream_ptr()))


def fill_ambient(planes, nx, ny, ng, var, side, value):
    """the compressible "ambient" boundary: ghost rows of plane `var` beyond side 0 (ylb) / 1 (yrb) <- value"""
    require_cuda()
    g = grid_struct(planes, nx, ny, ng)
    _lib.check(_lib.lib().p2b_fill_ambient_f64(planes.data_ptr(), C.byref(g), var, side, float(value), _lib.stream_ptr()))


def new_scratch():
    require_cuda()
    return torch.zeros(8, dtype=torch.int64, device="cuda")


def cfl_wavemax(planes, nx, ny, ng, gamma, scratch):
    """max(|u|+cs), max(|v|+cs) over the full array incl. ghosts -> two python floats"""
    g = grid_struct(planes, nx, ny, ng)
    scratch[:2].zero_()
    _lib.check(_lib.lib().p2b_cfl_wavemax(planes.data_ptr(), C.byref(g), gamma, scratch.data_ptr(),
                                          _lib.stream_ptr()))
    w = scratch[:2].view(torch.float64).tolist()
    return w[0], w[1]


def comp_params(gamma=1.4, z0=0.75, z1=0.85, delta=0.33, cvisc=0.1, limiter=2, use_flattening=1,
                no_avisc_xhi=1, no_avisc_yhi=1, grav=0.0, src_flip_ylo=0, src_flip_yhi=0, riemann="HLLC",
                xl_solid=0, yl_solid=0, heat_rate=0.0, heat_profile=None, sponge=None, src_copy_yhi=0,
                geometry=None, src_flip_xlo=0, src_flip_xhi=0):
    """heat_profile: a (qx, pitch) CUDA plane laid out like one state plane (the caller keeps it alive);
    sponge: (rho_begin, rho_full, timescale) or None; geometry: (geo_i, geo_j) device tables of a SphericalPolar
    grid (mesh.patch.SphericalPolar.sweep_tables), the caller keeps them alive"""
    sp = sponge or (0.0, 0.0, 1.0)
    return _lib.CompParams(gamma, z0, z1, delta, cvisc, limiter, use_flattening, no_avisc_xhi, no_avisc_yhi,
                           grav, src_flip_ylo, src_flip_yhi, {"HLLC": 0, "CGF": 1, "HLLC_lm": 2}[riemann], xl_solid, yl_solid,
                           heat_rate, None if heat_profile is None else heat_profile.data_ptr(),
                           int(sponge is not None), sp[0], sp[1], sp[2], src_copy_yhi,
                           None if geometry is None else geometry[0].data_ptr(),
                           None if geometry is None else geometry[1].data_ptr(),
                           0 if geometry is None else geometry[0].shape[1], 0 if geometry is None else geometry[1].shape[1],
                           src_flip_xlo, src_flip_xhi)


def compressible_sweep(Uin, Uout, nx, ny, ng, dx, dy, dt, params, scratch):
    """one fused CTU step: Uin (ghosts filled) -> valid region of Uout; asynchronous"""
    assert Uin.shape == Uout.shape and Uin.stride() == Uout.stride()
    g = grid_struct(Uin, nx, ny, ng, dx, dy)
    _lib.check(_lib.lib().p2b_compressible_sweep(Uin.data_ptr(), Uout.data_ptr(), C.byref(g), C.byref(params),
                                                 dt, scratch.data_ptr(), _lib.stream_ptr()))


def sweep_info():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _lib.lib().p2b_sweep_info(C.byref(a), C.byref(b), C.byref(c))
    return {"ntasks": a.value, "resident_warps": b.value, "seglen": c.value,
            "tensor_map_tma": bool(_lib.lib().p2b_sweep_uses_tensor_map())}


class _RawDeviceMemory:
    """a device allocation owned by libpyro2b200 (p2b_shared_alloc), exposed to torch through the CUDA array interface"""

    def __init__(self, ptr, nelem):
        self.ptr, self.nelem = ptr, nelem
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def tensor_from_pointer(ptr, nelem):
    """float64 CUDA tensor over `nelem` doubles at device address `ptr` (no copy; the caller keeps the allocation alive)"""
    require_cuda()
    return torch.as_tensor(_RawDeviceMemory(ptr, nelem), device="cuda")
