"""Zone-by-zone comparison of two snapshots -- the mirror of pyro/util/compare.py (compare :22-66, main :69-92),
the tool behind the reference's regression tests.  The valid regions are compared on the device; only the
per-variable maxima come back to the host."""
import sys

import torch

from . import io_pyro as io

usage = """
      usage: python -m pyro2_b200.util.compare file1 file2 (rtol)

      where rtol is an (optional) relative tolerance parameter to use when
      comparing the data
"""

errors = {"gridbad": "grids don't agree",
          "namesbad": "variable lists don't agree",
          "varerr": "one or more variables don't agree"}


def compare(data1, data2, rtol=1.e-12):
    """0 if two CellCenterData2d objects agree to rtol, else a key of `errors`"""
    if not data1.grid == data2.grid:
        return "gridbad"
    if not sorted(data1.names) == sorted(data2.names):
        return "namesbad"
    print(" ")
    print("variable comparisons:")
    result = 0
    for name in data1.names:
        d1 = data1.get_var(name).v()
        d2 = data2.get_var(name).v().to(d1.device)
        diff = (d1 - d2).abs()
        abs_err = float(diff.max())
        if not bool((d2 == 0).any()):
            rel_err = float((diff / d2.abs()).max())
            print(f"{name:20s} absolute error = {abs_err:10.10g}, relative error = {rel_err:10.10g}")
        else:
            print(f"{name:20s} absolute error = {abs_err:10.10g}")
        if not torch.allclose(d1, d2, rtol=rtol):        # same default atol (1e-8) as numpy's
            result = "varerr"
    return result


def main():
    if len(sys.argv) not in (3, 4):
        print(usage)
        sys.exit(2)
    s1 = io.read(sys.argv[1])
    s2 = io.read(sys.argv[2])
    rtol = float(sys.argv[3]) if len(sys.argv) == 4 else 1.e-12
    result = compare(s1.cc_data, s2.cc_data, rtol=rtol)
    if result == 0:
        print("SUCCESS: files agree")
    else:
        print("ERROR: ", errors[result])


if __name__ == "__main__":
    main()
