"""Enumerate the `-m gpu` tests (functions + their parametrisation) so that they can be run on the emulated device
(tests/emu_device.py).  TEST INFRASTRUCTURE ONLY.

As a script it rehearses every case of the given modules and prints one line per case:

    python tests/rehearse.py test_gpu_api test_gpu_flow [--max-seconds 600] [--only substring]
"""
import importlib
import itertools
import os
import sys
import time
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def cases(module_name):
    """(test id, function, kwargs) for every parametrised case of every test function of the module"""
    mod = importlib.import_module(module_name)
    out = []
    for name in sorted(vars(mod), key=lambda n: getattr(getattr(mod, n), "__code__", None).co_firstlineno
                       if hasattr(getattr(mod, n), "__code__") else 0):
        fn = getattr(mod, name)
        if not (name.startswith("test_") and callable(fn)):
            continue
        axes = []
        for mark in getattr(fn, "pytestmark", []):
            if mark.name == "parametrize":
                names = [n.strip() for n in mark.args[0].split(",")] if isinstance(mark.args[0], str) else list(mark.args[0])
                vals = [v if len(names) > 1 else (v,) for v in mark.args[1]]
                axes.append([dict(zip(names, v)) for v in vals])
        for combo in itertools.product(*axes) if axes else [()]:
            kw = {}
            for d in combo:
                kw.update(d)
            ident = name + ("[" + "-".join(str(v) for v in kw.values()) + "]" if kw else "")
            out.append((ident, fn, kw))
    return out


def run_case(fn, kw):
    import emu_device
    with emu_device.emulated_device() as dev:
        fn(**kw)
    return dev.calls


if __name__ == "__main__":
    import pytest
    args = sys.argv[1:]
    only = None
    if "--only" in args:
        only = args[args.index("--only") + 1]
        del args[args.index("--only"):args.index("--only") + 2]
    for m in args:
        for ident, fn, kw in cases(m):
            if only and only not in ident:
                continue
            t0 = time.time()
            try:
                calls = run_case(fn, kw)
                status = "PASS"
            except pytest.skip.Exception as e:
                status, calls = f"SKIP {e}", {}
            except BaseException:   # pylint: disable=broad-except
                status, calls = "FAIL", {}
                traceback.print_exc()
            print(f"{status:5s} {time.time() - t0:7.1f}s  {m}::{ident}  ({sum(calls.values())} library calls)", flush=True)
