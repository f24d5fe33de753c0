"""Domain decomposition across the GPUs of one box: 1-d slabs along x.

The reference is single-process (pyro/mesh/array_indexer.py:157-158: "there is only a single
grid"); this is the B200-side extension SURVEY.md 8(e) describes.  x is the slow storage axis, so a
slab's ghost rows are contiguous in every plane and can be sent / received in place, without
packing.  One process per GPU (torchrun); the exchange is torch.distributed point-to-point over
NCCL / NVLink (gloo on CPU tensors for the host-logic tests).

Why 4 rows: the dependency radius of one compressible cell update is exactly ng = 4
(SURVEY.md 9.3), so one exchange per time step reproduces the single-domain step bit for bit,
provided the artificial viscosity is applied on inter-slab faces but not on the global +x face
(SURVEY.md 9.2-13).
"""
import torch
import torch.distributed as dist


class SlabDecomposition:
    """rank r of `size` owns global rows [r * nx_local, (r + 1) * nx_local)"""

    def __init__(self, rank=None, size=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.size = dist.get_world_size(group) if size is None else size

    @property
    def is_first(self):
        return self.rank == 0

    @property
    def is_last(self):
        return self.rank == self.size - 1

    def local_nx(self, nx_global):
        if nx_global % self.size:
            raise ValueError(f"mesh.nx = {nx_global} is not divisible by the {self.size} slabs")
        return nx_global // self.size

    def ioffset(self, nx_global):
        return self.rank * self.local_nx(nx_global)

    def neighbours(self, periodic):
        """(low, high) ranks or None at a physical (non-periodic) boundary"""
        lo = self.rank - 1 if self.rank > 0 else (self.size - 1 if periodic else None)
        hi = self.rank + 1 if self.rank < self.size - 1 else (0 if periodic else None)
        if self.size == 1:
            lo = hi = None   # single slab: periodic wrap is the ordinary local ghost fill
        return lo, hi

    # ---- peer-memory transport (csrc/slab_comm.cu): planes allocated through shared_planes() are exchanged by kernels
    # that store straight into the neighbours' ghost rows; everything else goes over torch.distributed as before ----
    _peer = None

    def shared_planes(self, nvar, qx, qy, periodic):
        """collective: (nvar, qx, pitch) float64 planes in device memory every rank can write, registered with this
        decomposition's peer communicator (created on first use; `periodic`: the x direction wraps)"""
        import ctypes as C
        from . import _lib, ops
        L = _lib.lib()
        if self._peer is None:
            nbytes = L.p2b_slab_ctl_bytes()
            ctl = L.p2b_shared_alloc(nbytes)
            if not ctl:
                raise RuntimeError(L.p2b_last_error().decode())
            ptrs = self.map_peer_workspaces(ctl)
            h = L.p2b_slab_create(self.rank, self.size, int(bool(periodic)), (C.c_void_p * self.size)(*ptrs))
            if not h:
                raise RuntimeError(L.p2b_last_error().decode())
            self._peer = {"handle": h, "periodic": bool(periodic), "nbuf": 0, "keep": [ctl]}
        if self._peer["periodic"] != bool(periodic) or self._peer["nbuf"] >= 64:
            raise ValueError("peer communicator: one periodicity and at most 64 plane buffers per decomposition "
                             "(use a fresh SlabDecomposition for a problem with another x boundary type)")
        pitch = ops.row_pitch(qy)
        nelem = nvar * qx * pitch
        ptr = L.p2b_shared_alloc(nelem * 8)
        if not ptr:
            raise RuntimeError(L.p2b_last_error().decode())
        peers = self.map_peer_workspaces(ptr)
        _lib.check(L.p2b_slab_register(self._peer["handle"], self._peer["nbuf"], C.c_void_p(ptr), nelem * 8,
                                       (C.c_void_p * self.size)(*peers)))
        self._peer["nbuf"] += 1
        self._peer["keep"].append(ptr)
        return ops.tensor_from_pointer(ptr, nelem).view(nvar, qx, pitch)

    def _peer_owns(self, t, periodic):
        if self._peer is None or self._peer["periodic"] != bool(periodic) or not t.is_cuda:
            return False
        from . import _lib
        import ctypes as C
        return bool(_lib.lib().p2b_slab_owns(self._peer["handle"], C.c_void_p(t.data_ptr())))

    def exchange(self, planes, nx, ng, periodic=False):
        """fill the x ghost rows that face another slab with that slab's boundary rows.
        planes: (nvar, qx, pitch); rows of a plane are contiguous, so each plane's ng-row block is
        sent / received in place (torch.distributed) or stored by a kernel into the neighbour's ghost rows (planes
        from shared_planes())."""
        if self._peer_owns(planes, periodic):
            import ctypes as C
            from . import _lib
            _lib.check(_lib.lib().p2b_slab_exchange(self._peer["handle"], C.c_void_p(planes.data_ptr()), planes.shape[0],
                                                    planes.stride(0), planes.stride(1), nx, ng, _lib.stream_ptr()))
            return
        lo, hi = self.neighbours(periodic)
        ops = []
        # Posting order matters when both neighbours are the same rank (2 slabs, periodic): messages
        # between a pair of ranks match in posting order, so every rank posts, per plane,
        # send(top -> hi), send(bottom -> lo), recv(low ghost <- lo), recv(high ghost <- hi):
        # the peer's first send (its top rows) then lands in my low ghost rows, as it must.
        for n in range(planes.shape[0]):
            p = planes[n]
            if hi is not None:
                ops.append(dist.P2POp(dist.isend, p[nx:nx + ng], hi, self.group))            # my top valid rows
            if lo is not None:
                ops.append(dist.P2POp(dist.isend, p[ng:2 * ng], lo, self.group))             # my bottom valid rows
            if lo is not None:
                ops.append(dist.P2POp(dist.irecv, p[0:ng], lo, self.group))                  # low ghost rows
            if hi is not None:
                ops.append(dist.P2POp(dist.irecv, p[ng + nx:ng + nx + ng], hi, self.group))  # high ghost rows
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def map_peer_workspaces(self, ptr):
        """collective: every rank passes the device address of its (p2b_shared_alloc'd) multigrid workspace and gets the
        list of all ranks' workspaces as mapped into THIS process (its own address at [rank]).  Processes on one node:
        cudaIpc handles travel through torch.distributed, the mapping is opened by libpyro2b200 (p2b_shared_open)."""
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        buf = C.create_string_buffer(64)
        _lib.check(L.p2b_shared_handle(C.c_void_p(ptr), buf))
        handles = [None] * self.size
        dist.all_gather_object(handles, bytes(buf.raw), group=self.group)
        out = []
        for r, h in enumerate(handles):
            if r == self.rank:
                out.append(ptr)
                continue
            p = L.p2b_shared_open(C.create_string_buffer(h, 64))
            if not p:
                raise RuntimeError("cannot map rank %d's multigrid workspace: %s" % (r, L.p2b_last_error().decode()))
            out.append(p)
        return out

    def allreduce_max_(self, t):
        if self.size > 1:
            if self._peer is not None and t.is_cuda and t.dtype == torch.int64 and t.numel() == 4 and t.is_contiguous():
                # the sweep's scratch words (wave-speed maxima as bit patterns, status): reduced through peer slots
                import ctypes as C
                from . import _lib
                _lib.check(_lib.lib().p2b_slab_allreduce_max4(self._peer["handle"], C.c_void_p(t.data_ptr()), _lib.stream_ptr()))
                return t
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def check_peer(self):
        """raise if a wait of the peer-memory transport timed out (synchronises)"""
        if self._peer is not None:
            from . import _lib
            e = _lib.lib().p2b_slab_error(self._peer["handle"], _lib.stream_ptr())
            if e:
                raise RuntimeError(f"slab communicator: a wait on a neighbouring rank timed out (control word {e - 1})")

    def interior_sides(self, periodic):
        """(low_is_interior, high_is_interior)"""
        lo, hi = self.neighbours(periodic)
        return lo is not None, hi is not None


class LocalSlabGroup:
    """Several ranks inside ONE process (one host thread each, one stream each): the slabs of a decomposed multigrid
    then share an address space and reach each other's workspaces through plain pointers.  Used to exercise the
    peer-memory protocol on a single GPU (and by the CPU tests on the emulated device); real multi-GPU runs use one
    process per GPU (SlabDecomposition)."""

    def __init__(self, size):
        import threading
        self.size = size
        self._barrier = threading.Barrier(size)
        self._slots = [None] * size

    def member(self, rank):
        return _LocalSlab(self, rank)


class _LocalSlab(SlabDecomposition):
    def __init__(self, group, rank):
        self.group = None
        self._g = group
        self.rank, self.size = rank, group.size

    def map_peer_workspaces(self, ptr):
        g = self._g
        g._barrier.wait()
        g._slots[self.rank] = ptr
        g._barrier.wait()
        out = list(g._slots)
        g._barrier.wait()
        return out

    def barrier(self):
        self._g._barrier.wait()

    def exchange(self, planes, nx, ng, periodic=False):
        if not self._peer_owns(planes, periodic):
            raise NotImplementedError("ranks that share a process exchange through the library's peer-memory kernels only")
        super().exchange(planes, nx, ng, periodic)

    def allreduce_max_(self, t):
        if self._peer is None or t.dtype != torch.int64 or t.numel() != 4:
            raise NotImplementedError("ranks that share a process exchange through the library's peer-memory kernels only")
        return super().allreduce_max_(t)
