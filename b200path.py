"""Puts the product package directory (flat modules named like the reference's code/) on sys.path."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
_matches = sorted(glob.glob(os.path.join(ROOT, "*_b200")))
if not _matches:
    raise ImportError("package directory *_b200 not found next to b200path.py")
PKG_DIR = _matches[0]
if PKG_DIR not in sys.path:
    sys.path.insert(0, PKG_DIR)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
