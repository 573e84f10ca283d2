"""Importable alias for the package directory `whisper-diarize-rs_b200/` (a hyphen cannot be imported by name)."""
import importlib
import sys

_pkg = importlib.import_module("whisper-diarize-rs_b200")
sys.modules[__name__] = _pkg
