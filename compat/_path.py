"""Puts the repository root on sys.path so the flat compat modules can import the `deer_b200` package."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
