"""Recipe that makes the UNMODIFIED reference runnable where /root/reference is not mounted.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).

The reference is pure Python: its per-frame path (`onnx_model/`, the module tree its exporter traces),
its offline parity model (`model/`) and its public package (`package/src/dpdfnet/`) need no build step,
but the GPU box has no `/root/reference`.  This script copies those source files **verbatim** into the
git-ignored `oracle/_ref/` (same relative layout), which travels with the gpurun snapshot, and writes
`oracle/_ref/MANIFEST.json` (sha256 per file) so a run can prove the files are untouched.

    python oracle/build_ref.py            # copy if /root/reference is present, verify otherwise

`__graft_entry__.build()` calls `ensure()`; `oracle/ref_import.py` resolves the reference root as
$DPDFNET_REFERENCE, then /root/reference, then oracle/_ref.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from pathlib import Path

SRC = Path(os.environ.get("DPDFNET_REFERENCE_SRC", "/root/reference"))
DST = Path(__file__).resolve().parent / "_ref"
# (directory, glob) pairs of the reference tree that the per-frame path, the offline oracle and the public
# streaming API consist of; nothing else is copied.
PARTS = (("onnx_model", "*.py"), ("model", "*.py"), ("package/src/dpdfnet", "*.py"))


def _sha(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def ensure(verbose: bool = False) -> Path | None:
    """Copy (when the reference tree is mounted) or verify (when only the copy exists). Returns DST or None."""
    if SRC.is_dir() and (SRC / "onnx_model").is_dir():
        manifest = {}
        for sub, pat in PARTS:
            (DST / sub).mkdir(parents=True, exist_ok=True)
            for f in sorted((SRC / sub).glob(pat)):
                out = DST / sub / f.name
                if not out.exists() or _sha(out) != _sha(f):
                    shutil.copyfile(f, out)
                manifest[f"{sub}/{f.name}"] = _sha(out)
        (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "files": manifest}, indent=1))
        if verbose:
            print(f"oracle/_ref: {len(manifest)} reference files copied verbatim from {SRC}")
        return DST
    if (DST / "MANIFEST.json").exists():
        man = json.loads((DST / "MANIFEST.json").read_text())
        bad = [k for k, h in man["files"].items() if not (DST / k).exists() or _sha(DST / k) != h]
        if bad:
            raise RuntimeError(f"oracle/_ref differs from its manifest: {bad[:5]}")
        if verbose:
            print(f"oracle/_ref: {len(man['files'])} files verified against MANIFEST.json")
        return DST
    if verbose:
        print("oracle/_ref: reference tree not available here")
    return None


if __name__ == "__main__":
    sys.exit(0 if ensure(verbose=True) else 1)
