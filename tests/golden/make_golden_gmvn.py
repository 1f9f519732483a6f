"""Generate tests/golden/gmvn.npz from the UNMODIFIED reference GlobalMVN (espnet2/layers/global_mvn.py).

    python tests/golden/make_golden_gmvn.py

Stores the statistics (count / sum / sum_square), a padded batch, and the reference outputs for the four
(norm_means, norm_vars) settings.  Run in the build container only.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
from espnet2.layers.global_mvn import GlobalMVN  # noqa: E402

rng = np.random.RandomState(7)
D, N = 80, 1000
data = (rng.randn(N, D) * rng.uniform(0.5, 3.0, D) + rng.uniform(-4, 4, D)).astype(np.float64)
stats = dict(count=np.array(N, dtype=np.int64), sum=data.sum(0), sum_square=(data * data).sum(0))
x = (rng.randn(3, 11, D) * 2 + 1).astype(np.float32)
ilens = np.array([11, 6, 9], dtype=np.int64)
out = dict(x=x, ilens=ilens, **{"stats_" + k: v for k, v in stats.items()})
with tempfile.TemporaryDirectory() as td:
    f = os.path.join(td, "stats.npz")
    np.savez(f, **stats)
    for nm in (True, False):
        for nv in (True, False):
            m = GlobalMVN(f, norm_means=nm, norm_vars=nv)
            y, ol = m(torch.from_numpy(x.copy()), torch.from_numpy(ilens))
            out[f"y_m{int(nm)}_v{int(nv)}"] = y.numpy()
            out["mean"], out["std"] = m.mean.numpy(), m.std.numpy()
np.savez_compressed(os.path.join(HERE, "gmvn.npz"), **out)
print("wrote gmvn.npz", {k: getattr(v, "shape", None) for k, v in out.items()})
