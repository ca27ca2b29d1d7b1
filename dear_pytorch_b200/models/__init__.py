"""Model zoo used by the benchmark drivers (see models/registry.py)."""
