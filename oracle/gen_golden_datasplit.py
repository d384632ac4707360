"""Generate tests/golden/datasplit/select.pt from the REAL reference (build container only):
    python oracle/gen_golden_datasplit.py
Pins the gather / mask / grid-select semantics of npf/utils/datasplit.py on given indices and masks (the random
indices themselves come from oracle/datasplit_oracle.py: numpy's Mersenne-Twister stream is not reproducible on the
device, see that file's header)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))
from npf.utils.datasplit import CntxtTrgtGetter, GetRandomIndcs, GridCntxtTrgtGetter, RandomMasker, get_all_indcs  # noqa: E402
from oracle import datasplit_oracle as D  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "datasplit", "select.pt")


def main():
    g = torch.Generator().manual_seed(0)
    cases = []
    # ---- off-grid: CntxtTrgtGetter.select on given context indices, all points as targets
    for (B, N, xd, yd, n, seed, add) in [(3, 20, 1, 1, 7, 11, False), (4, 50, 2, 3, 50, 12, False), (2, 16, 1, 2, 5, 13, True),
                                         (2, 9, 1, 1, 0, 14, False)]:
        X = torch.rand(B, N, xd, generator=g) * 2 - 1
        Y = torch.randn(B, N, yd, generator=g)
        ci = torch.from_numpy(D.random_subset(B, N, n, seed)).long()
        getter = CntxtTrgtGetter(is_add_cntxts_to_trgts=add)
        Xc, Yc, Xt, Yt = getter(X, Y, context_indcs=ci, target_indcs=get_all_indcs(B, N))
        cases.append(dict(kind="points", X=X, Y=Y, context_indcs=ci.int(), seed=seed, add=add, X_cntxt=Xc, Y_cntxt=Yc, X_trgt=Xt,
                          Y_trgt=Yt))
    # ---- grids: GridCntxtTrgtGetter.select on a given context mask, all pixels as targets
    for (B, grid, yd, n, seed, up) in [(3, (5, 7), 3, 9, 21, 1), (2, (32, 32), 1, 300, 22, 1), (2, (6, 4), 2, 24, 23, 2),
                                       (3, (40,), 2, 11, 24, 1)]:
        P = int(np.prod(grid))
        img = torch.rand(B, yd, *grid, generator=g)
        mask = torch.from_numpy(D.random_mask(B, P, n, seed)).bool().view(B, *grid, 1)
        getter = GridCntxtTrgtGetter(upscale_factor=up)
        Xc, Yc, Xt, Yt = getter(img, context_mask=mask, target_mask=torch.ones(B, *grid, 1).bool())
        cases.append(dict(kind="grid", img=img, context_mask=mask, n=n, seed=seed, upscale=up, X_cntxt=Xc, Y_cntxt=Yc, X_trgt=Xt,
                          Y_trgt=Yt))
    # ---- the host-side draw of HOW MANY points: sequence of sizes the reference produces after its set_seed(123)
    import random
    random.seed(123); np.random.seed(123); torch.manual_seed(123)
    sizes = []
    for g_ in (GetRandomIndcs(a=0.0, b=50), GetRandomIndcs(a=0.1, b=0.5), GetRandomIndcs(a=3, b=3, is_ensure_one=True)):
        sizes.append([int(g_(4, 128).shape[1]) for _ in range(12)])
    sizes.append([int(RandomMasker(a=0.0, b=0.3)(2, (32, 32)).sum()) // 2 for _ in range(6)])
    cases.append(dict(kind="sizes", seed=123, sizes=sizes))
    torch.save(cases, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
