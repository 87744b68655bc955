"""Import helper: the package directory is named after the reference (with hyphens), which the
`import` statement cannot spell.  `from pcq_import import pcq` gives the package."""
import importlib
import sys
from pathlib import Path

_ROOT = str(Path(__file__).resolve().parent)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

pcq = importlib.import_module("adhoc-queries-pointclouds_b200")
