"""B200 implementations of the reference's `modules/` classes on the GPS hot path, registered under
the same names (see registry.py)."""
