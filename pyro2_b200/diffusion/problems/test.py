"""Uniform phi: the setup the reference's unit tests build their simulations on
(pyro/diffusion/problems/test.py).  It has no stock inputs file."""

DEFAULT_INPUTS = None

PROBLEM_PARAMS = {}


def init_data(my_data, rp):   # pylint: disable=unused-argument
    my_data.get_var("phi")[:, :] = 1.0


def finalize():
    pass
