"""Forwarder: the full-size BASELINE configs live in tests/full_configs.py (they check against the CPU oracle, which only
tests/ may use).  Same arguments: python scripts/run_configs.py [1 2 3 4 | 5 under torchrun] [--scale F]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import full_configs  # noqa: E402

if __name__ == "__main__":
    full_configs.main(sys.argv[1:])
