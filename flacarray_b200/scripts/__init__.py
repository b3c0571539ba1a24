"""Command-line entry points (reference scripts/__init__.py)."""
