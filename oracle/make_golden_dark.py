"""tests/golden/dark.npz from the LIVE reference's get_final_preds_v2 (src/utils/inference.py:70-87); dev container
only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_dark.py      """
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(1, REPO)
from src.utils.inference import get_final_preds_v2 as ref_v2          # noqa: E402
from oracle.golden_inputs import dark_cases                          # noqa: E402


def main():
    out = {}
    cases = dark_cases()
    for i, c in enumerate(cases):
        hm = torch.from_numpy(c["hm"].copy())          # the reference blurs its argument in place
        out[f"pred{i}"] = ref_v2(hm, c["center"], c["scale"], c["output_size"])
    out["n"] = np.array(len(cases))
    path = os.path.join(REPO, "tests", "golden", "dark.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
