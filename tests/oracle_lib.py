"""Loads the CPU oracle (oracle/liboracle_sgdnet.so), building it with its Makefile when needed.
Only tests, smoke() and bench.py's CPU arms go through here."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_sgdnet.so")


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "sgdnet_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "sgdnet_b200.h")
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(ORACLE_SO) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "-B", "liboracle_sgdnet.so"])
    return ORACLE_SO


def load_oracle():
    from sgdnet_b200._abi import Library
    return Library(build_oracle(), "oracle_")
