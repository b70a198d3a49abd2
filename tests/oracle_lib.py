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


REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsgdnet_ref.so")
REF_SRC = "/root/reference/src/sgdnet.cpp"


def load_reference_build():
    """oracle/_ref/libsgdnet_ref.so: the reference's own src/sgdnet.cpp (+ headers) compiled from where it lies against
    the Rcpp/Eigen stand-in of oracle/refbuild/. Built where /root/reference exists; elsewhere the prebuilt file is used.
    Returns None when neither is available."""
    if os.path.exists(REF_SRC):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ORACLE_DIR, "refbuild")])
    if not os.path.exists(REF_SO):
        return None
    from sgdnet_b200._abi import Library
    return Library(REF_SO, "ref_")
