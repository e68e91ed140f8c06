"""Generates tests/golden/pose_accuracy.npz with the REAL reference libs.metrics.pose_accuracy
(imported from /root/reference).  Run in the build container only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")
from libs.metrics import pose_accuracy  # noqa: E402  (the reference)

from tests.golden.cases import metric_cases  # noqa: E402

OUT = Path(__file__).resolve().parent

if __name__ == "__main__":
    d = {}
    for name, (prd, tgt) in metric_cases().items():
        acc, avg_acc, cnt, pred = pose_accuracy(prd, tgt)
        d["acc_" + name], d["avg_" + name], d["cnt_" + name], d["pred_" + name] = acc, np.array([avg_acc]), np.array([cnt]), pred
        print(name, avg_acc, cnt)
    np.savez_compressed(OUT / "pose_accuracy.npz", **d)
