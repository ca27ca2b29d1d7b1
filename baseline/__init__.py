"""Reference arm of the benchmark (NOT product code): runs the unmodified reference from baseline/_ref."""
