"""Minimal stand-in for the `vipy` package (not installed, no network).

Test infrastructure only: lets tests/golden/make_golden.py import the read-only
reference at /root/reference inside the build container to generate golden
vectors.  Covers only the handful of helpers the keyed-layer path touches.
"""
from . import util, image  # noqa: F401
