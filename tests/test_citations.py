"""Every `file:line` citation of the reference in the ABI header, the docs, the oracle and the CUDA sources must point at
an existing file of the reference tree with at least that many lines.  Runs only where the read-only reference is mounted
(the build container); skipped elsewhere (the GPU box has no /root/reference and nothing there needs it)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAT = re.compile(r"([A-Za-z0-9_/\.]+\.(?:h|cpp|yaml|md|txt)):(\d+)(?:-(\d+))?")


def _sources():
    out = [os.path.join(ROOT, p) for p in ("DESIGN.md", "INTEGRATION.md", "include/fls_b200.h", "funny_lidar_slam_b200/shim/b200_registration.h")]
    for d, exts in (("oracle", (".h", ".cpp", ".py")), ("funny_lidar_slam_b200/csrc", (".cu", ".cuh", ".h"))):
        out += [os.path.join(ROOT, d, f) for f in sorted(os.listdir(os.path.join(ROOT, d))) if f.endswith(exts)]
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_citations_resolve():
    files = {}
    for d, _, fs in os.walk(REF):
        if "/.git" in d:
            continue
        for f in fs:
            files.setdefault(f, []).append(os.path.join(d, f))

    def n_lines(p):
        with open(p, errors="ignore") as fh:
            return sum(1 for _ in fh)

    checked, bad = 0, []
    for src in _sources():
        for m in PAT.finditer(open(src, errors="ignore").read()):
            path, last = m.group(1), int(m.group(3) or m.group(2))
            base = os.path.basename(path)
            if base.startswith(("fls_", "orc_")) or base == "b200_registration.h":
                continue  # this repository's own files
            checked += 1
            cands = files.get(base, [])
            if "/" in path:
                cands = [c for c in cands if c.endswith(path)] or cands
            if not cands:
                bad.append((os.path.relpath(src, ROOT), m.group(0), "no such file in the reference"))
            elif max(n_lines(c) for c in cands) < last:
                bad.append((os.path.relpath(src, ROOT), m.group(0), "file is shorter than the cited line"))
    assert checked > 150
    assert not bad, bad
