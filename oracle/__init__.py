"""CPU oracles for the rfnet_b200 parity tests.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import this package.  ``rfnet_b200`` never does: the product path has no CPU fallback.

* :mod:`oracle.port`  -- numpy front-end of ``librfnet_oracle.so`` (plain-C restatement, ``rfnet_oracle.c``).
* :mod:`oracle.ref`   -- numpy/torch front-end of ``_ref/libref_{cpu,gpu}.so``: the reference's own OpKernels and CUDA
  kernels, compiled unmodified from ``/root/reference`` against ``tf_stub`` by ``oracle/Makefile``.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("RFNET_REFERENCE_ROOT", "/root/reference")


def build(with_ref=None):
    """Compile the oracle libraries (idempotent).  ``_ref`` is only (re)built where the reference sources exist;
    on the GPU box the prebuilt ``_ref/*.so`` that travelled with the snapshot are used as they are."""
    subprocess.check_call(["make", "-s", "-C", HERE, "librfnet_oracle.so"])
    if with_ref is None:
        with_ref = os.path.isdir(REFERENCE_ROOT)
    if with_ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "REF=" + REFERENCE_ROOT])
