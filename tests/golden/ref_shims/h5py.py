"""Shim: h5py is imported at module top by util.py; never called on the hot path."""
