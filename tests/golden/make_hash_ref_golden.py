"""Records the outputs of the REFERENCE's own CUDA kernels (oracle/_ref/_hash_encoder_ref.so, built
from /root/reference/hashencoder/src by oracle/build_ref.py) on seeded inputs.

Runs on the GPU box:   python tests/golden/make_hash_ref_golden.py   -> gpurun_out/hash_ref_*.npz
The files are then committed under tests/golden/ and pin oracle/hash_oracle.c (tests/test_oracle_hash.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests import hash_cases  # noqa: E402

CASES = {"hash_ref_full": dict(B=512, logmap=19, seed=1234), "hash_ref_small": dict(B=768, logmap=12, seed=99)}


def main():
    ref = build_ref.load_ref()
    assert ref is not None, "oracle/_ref/_hash_encoder_ref.so missing"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name, kw in CASES.items():
        c = hash_cases.make_case(**kw)
        r = hash_cases.backend_all(ref, c)
        gi, gv = hash_cases.to_coo(r["gemb"])
        hi, hv = hash_cases.to_coo(r["g2"])
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", name + ".npz"), out=r["out"].numpy(), dy_dx=r["dy_dx"].numpy(),
                            gx=r["gx"].numpy(), gg=r["gg"].numpy(), gemb_idx=gi, gemb_val=gv, g2_idx=hi, g2_val=hv,
                            meta=np.array([kw["B"], kw["logmap"], kw["seed"]]),
                            device=np.array(torch.cuda.get_device_name(0)))
        print(name, "written")


if __name__ == "__main__":
    main()
