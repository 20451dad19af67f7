"""Import target of ``custom_imports`` (data pipeline is out of scope)."""
