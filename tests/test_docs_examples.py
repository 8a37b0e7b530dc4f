"""Every fenced ``python`` block in ``docs/`` that starts with ``# runnable`` is executed here (CPU, single process),
file by file, blocks of one file sharing a namespace - so the documentation cannot drift away from the code."""

import re
from pathlib import Path

import pytest

DOCS = Path(__file__).resolve().parent.parent / "docs"
_BLOCK = re.compile(r"```python\n(# runnable\n.*?)```", re.DOTALL)


def _pages():
    return sorted(p for p in DOCS.rglob("*.md") if _BLOCK.search(p.read_text(encoding="utf-8")))


@pytest.mark.parametrize("page", _pages(), ids=lambda p: str(p.relative_to(DOCS)))
def test_documentation_examples_run(page, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    namespace = {"__name__": f"docs_example_{page.stem}", "TMP": tmp_path}
    for i, block in enumerate(_BLOCK.findall(page.read_text(encoding="utf-8"))):
        try:
            exec(compile(block, f"{page.name}[block {i}]", "exec"), namespace)  # noqa: S102
        except Exception as exc:  # noqa: BLE001
            raise AssertionError(f"{page.relative_to(DOCS)} block {i} failed: {exc!r}\n{block}") from exc
