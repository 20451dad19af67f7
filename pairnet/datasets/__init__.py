"""Import target of ``custom_imports``; datasets are out of scope (benchmarks use synthetic tensors)."""
