#!/usr/bin/env python3
"""One process, one torch import: the GPU parity tests that exercise the RMI pre-filter, smoke(), then the A/B check.
Everything is appended to gpurun_out/oneshot.log as it happens (the call may be cut short)."""
import io
import os
import sys
import time
import contextlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/oneshot.log", "a", buffering=1)
T0 = time.time()


def say(*a):
    LOG.write(f"[{time.time() - T0:6.1f}s] " + " ".join(str(x) for x in a) + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


def main():
    import pytest
    say("start")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        rc = pytest.main(["tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                          "-k", "all_methods_vs_oracle or rmi_fast_search_changes_nothing"])
    say("pytest rc =", int(rc))
    say(buf.getvalue()[-3000:])
    try:
        import __graft_entry__ as e
        e.smoke()
        say("smoke ok")
    except BaseException as ex:          # noqa: BLE001 - the log is the product here
        say("smoke FAILED:", repr(ex)[:1500])
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import rmi_prefilter_check as chk
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rc2 = chk.main(sys.argv[1:])
        say("check rc =", rc2, buf.getvalue())
    except BaseException as ex:          # noqa: BLE001
        say("check FAILED:", repr(ex)[:1500])
    say("done")


if __name__ == "__main__":
    main()
