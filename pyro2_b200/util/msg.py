"""Status / failure messages with the reference's conventions (pyro/util/msg.py:20-52):
``fail`` prints and exits when not interactive, raises otherwise."""
import sys


def fail(string):
    """fatal error: same contract as pyro.util.msg.fail -- SystemExit(1) unless interactive"""
    print(f"\033[1m\033[31m{string}\033[0m")
    if hasattr(sys, "ps1"):
        raise RuntimeError(string)
    sys.exit(1)


def warning(string):
    print(f"\033[35m{string}\033[0m")


def success(string):
    print(f"\033[32m{string}\033[0m")


def bold(string):
    print(f"\033[1m{string}\033[0m")
