"""Auxiliary subsystems: checkpointing, profiling, tracing, perf models."""
