"""Shim: matplotlib is imported at module top by util.py / powerspectrum.py but never called on the hot path."""
