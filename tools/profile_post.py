"""The post-processing and image-pipeline legs of bench.py alone, for ncu captures (traffic of rs_scatter / lanczos kernels)."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from detectorfreesfm_b200 import KeypointMerger
from detectorfreesfm_b200.image_pipeline import GpuImageReader, process_resize
from tests import util

dev = torch.device("cuda", 0)
n_img, m_pair = 64, 2000
pairs = list(itertools.combinations(range(n_img), 2))
pm, names = util.synth_matches(n_img, pairs, m_pair, seed=5, dup=0.25)
merger = KeypointMerger(dev)
rows = torch.from_numpy(np.concatenate(list(pm.values()), 0)).to(dev)
counts = np.array([v.shape[0] for v in pm.values()], dtype=np.int64)
pair_off = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).to(dev)
index = {n: i for i, n in enumerate(names)}
pair_img = torch.tensor([[index[k.split(" ")[0]], index[k.split(" ")[1]]] for k in pm], dtype=torch.int32, device=dev)
for _ in range(2):
    out = merger.merge(rows, pair_off, pair_img, n_img)
torch.cuda.synchronize()
print("keypoints", int(out[0].shape[0]), "observations", 2 * int(rows.shape[0]))
rd = GpuImageReader(dev)
photo = torch.from_numpy(util.synth_photo(3000, 4000, seed=3)).to(dev)
size = process_resize(4000, 3000, (1200,), 8)
for _ in range(2):
    rd.resize_gray(photo, size)
torch.cuda.synchronize()
print("done")
