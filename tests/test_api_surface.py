"""Import-path parity with d9d v0.14: every sub-package of the reference exists under ``d9d_b200`` and exports the names
the reference's ``__init__`` files export (``tests/api_surface_manifest.json``: package -> public names, names only)."""

import importlib
import json
from pathlib import Path

import pytest

MANIFEST = json.loads((Path(__file__).parent / "api_surface_manifest.json").read_text())


@pytest.mark.parametrize("package", sorted(MANIFEST), ids=lambda p: p or "<root>")
def test_reference_public_names_resolve(package):
    module = importlib.import_module("d9d_b200" + (f".{package}" if package else ""))
    missing = [name for name in MANIFEST[package] if not hasattr(module, name)]
    assert not missing, f"d9d_b200.{package} lacks {missing}"
