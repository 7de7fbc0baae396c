"""Stub for the un-vendored third-party `tme` package (only disc_chirp_tme uses it; out of scope)."""
