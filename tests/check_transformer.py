"""Stand-alone transformer parity check (run in a subprocess by tests/test_coarse_gpu.py with the schedule switches of the engine set
in the environment: they are read once per process)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import loftr_oracle as lo
from tests import util, weights


def main():
    from detectorfreesfm_b200 import B200LoFTR
    sd = weights.loftr_state_dict(0)
    m = B200LoFTR(util.loftr_config()).cuda().eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    for L, S in ((300, 417), (1100, 520)):
        f0, f1 = torch.randn(1, L, 256, generator=g), torch.randn(1, S, 256, generator=g)
        r0, r1 = lo.local_feature_transformer(f0, f1, sd, "loftr_coarse", ["self", "cross"] * 4, 8)
        o0, o1 = m.transform(f0[0].cuda().clone(), f1[0].cuda().clone())
        for o, r in ((o0, r0), (o1, r1)):
            e = ((o.cpu() - r[0]).abs().max() / r[0].abs().max()).item()
            assert e < 5e-5, (L, S, e)
    print("ok")


if __name__ == "__main__":
    main()
