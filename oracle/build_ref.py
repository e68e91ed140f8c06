"""Recipe that stages the REAL reference's hot-path modules under oracle/_ref/ (test infrastructure, not product).

The reference is pure Python (no native sources to compile), so "building" it is staging the few modules of the
path from where they lie in /root/reference into oracle/_ref/ - git-ignored (it never enters history), but not
gpurun-ignored (it travels to the GPU box, where /root/reference does not exist).  bench.py's reference arm and
cpu_baseline then time the UNMODIFIED reference (`kind: "reference"`) instead of the oracle's restatement
(`kind: "port"`); tests use it, when present, as one more pin of the restatement.  Nothing under hgr_b200/ ever
imports it.  Run by __graft_entry__.build(); a no-op when /root/reference is absent and oracle/_ref already exists.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

REFERENCE = Path("/root/reference")
DEST = Path(__file__).resolve().parent / "_ref"
# model/multitasknet.py:8-29 and everything it imports, plus the tail / loss / metrics modules of the path
FILES = ["model/__init__.py", "model/gelan.py", "model/transformer.py", "model/multitasknet.py",
         "libs/__init__.py", "libs/utils.py", "libs/loss.py", "libs/metrics.py"]


def stage(verbose: bool = True) -> Path | None:
    if not REFERENCE.is_dir():
        if verbose:
            print(f"oracle/_ref: {REFERENCE} not present, keeping {'the staged copy' if DEST.is_dir() else 'the port only'}")
        return DEST if DEST.is_dir() else None
    for rel in FILES:
        src, dst = REFERENCE / rel, DEST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
    if verbose:
        print(f"oracle/_ref: staged {len(FILES)} reference modules from {REFERENCE}")
    return DEST


def load():
    """Imports the staged reference: returns (MultiTaskNet class, get_max_preds) or None when it is not staged."""
    if not (DEST / "model" / "multitasknet.py").exists():
        return None
    if str(DEST) not in sys.path:
        sys.path.insert(0, str(DEST))
    from model.multitasknet import MultiTaskNet  # noqa: E402  (the reference's own package name)
    from libs.utils import get_max_preds  # noqa: E402
    return MultiTaskNet, get_max_preds


if __name__ == "__main__":
    stage()
