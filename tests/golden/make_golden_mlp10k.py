"""Golden vectors for the tensor-core MLP path at BASELINE config 3's full batch (B = 10 000, 785-256-256-256-784),
produced by the UNMODIFIED reference ``torchcfm.models.MLP`` (torchcfm/models/models.py:4-21) fed the way
``torch_wrapper`` feeds it (torchcfm/utils.py:51-52: t as the LAST input column).

    PYTHONPATH=/root/repo/oracle:/root/reference python tests/golden/make_golden_mlp10k.py

The 10 000 x 784 float64 output is 63 MB, so the fixture keeps what pins every row and column at once without
storing them all: 96 whole rows (first / last rows of the first and last 128-row slabs, rows straddling slab
boundaries, random rows), the float64 column sums over ALL rows, and the float64 row sums of ALL rows.
Weights and inputs are regenerated from seeds by the test (default nn.Linear init under torch.manual_seed(0),
exactly how the reference would build the model).
"""
import os

import numpy as np
import torch

import ot  # noqa: F401  (oracle/ot shim, needed for `import torchcfm`)
assert "oracle" in ot.__version__
import torchcfm  # noqa: E402
from torchcfm.models.models import MLP  # noqa: E402

assert torchcfm.__file__.startswith("/root/reference"), torchcfm.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
B, DIM, W, T_VALUE = 10000, 784, 256, 0.37


def inputs():
    g = torch.Generator().manual_seed(2024)
    return torch.randn(B, DIM, generator=g)


def main():
    torch.manual_seed(0)
    mlp = MLP(dim=DIM, w=W, time_varying=True)
    x = inputs()
    inp = torch.cat([x, torch.full((B, 1), T_VALUE)], 1)
    with torch.no_grad():
        y64 = mlp.double()(inp.double()).numpy()
    rs = np.random.RandomState(5)
    rows = np.unique(np.concatenate([np.arange(0, 4), np.arange(124, 132), np.arange(252, 260),
                                     np.arange(9980, 10000), np.arange(9856, 9860),
                                     rs.randint(0, B, size=52)]))
    out = {"rows": rows.astype(np.int64), "y64_rows": y64[rows], "col_sums": y64.sum(0), "row_sums": y64.sum(1),
           "abs_max": np.array(np.abs(y64).max()), "t": np.array(T_VALUE), "w0_probe": mlp.net[0].weight[:2, :4].detach().numpy()}
    np.savez_compressed(os.path.join(HERE, "mlp10k_vectors.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
