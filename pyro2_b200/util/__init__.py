"""Host-side utilities of the B200 build: run-time parameters, timers, messages, snapshot reading and comparison
(the counterparts of pyro/util; plotting helpers are host-side matplotlib in the reference and not built)."""
__all__ = ["runparams", "profile_pyro", "io_pyro", "compare", "msg"]


def __getattr__(name):
    # io_pyro pulls in the mesh package, which itself uses pyro2_b200.util: resolve the two readers lazily
    if name in ("read", "read_bcs"):
        from . import io_pyro   # pylint: disable=import-outside-toplevel
        return getattr(io_pyro, name)
    raise AttributeError(name)
