"""Fourth-order finite-volume data: conversion between cell averages and cell-centre values -- the mirror
of pyro/mesh/fv.py (FV2d :8-39), used by the reference's fourth-order solvers.  Assumes dx = dy.

The two conversions are a five-point stencil each; they are written with the ArrayIndexer views and
elementwise device operations in the reference's order (every operation rounded once), so the results are
bit-identical to the reference's."""
import torch

from .patch import CellCenterData2d


class FV2d(CellCenterData2d):
    """cell-averaged data with fourth-order accurate average <-> centre conversions"""

    def to_centers(self, name, is_positive=False):
        """cell-centre values of variable `name` (stored as averages): a - dx^2 lap(a) / 24 on all but the
        outermost ghost layer; is_positive keeps the average wherever the centre value would be negative"""
        a = self.get_var(name)
        c = self.grid.scratch_array()
        ng = self.grid.ng
        c[:, :] = a.t()
        centred = a.v(buf=ng - 1) - self.grid.dx ** 2 * a.lap(buf=ng - 1) / 24.0
        if is_positive:
            centred = torch.where(centred >= 0.0, centred, a.v(buf=ng - 1))
        c.v(buf=ng - 1)[:, :] = centred
        return c

    def from_centers(self, name):
        """treat the stored data of `name` as cell-centre values and replace them by cell averages
        (ghost cells are filled first: the Laplacian needs them)"""
        self.fill_BC(name)
        a = self.get_var(name)
        a.v()[:, :] = a.v() + self.grid.dx ** 2 * a.lap() / 24.0
