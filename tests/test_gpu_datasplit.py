"""-m gpu: the on-device context / target split (npf_random_subset / npf_random_mask / npf_select_points /
npf_grid_select behind npf_b200.utils.datasplit) -- index and byte work, so everything is compared BIT-EXACTLY: the
random subsets with oracle/datasplit_oracle.py (same Philox draws), the gathers with vectors produced by the real
reference (tests/golden/datasplit/select.pt)."""
import os
import random

import numpy as np
import pytest
import torch

from _util import ROOT
from oracle import datasplit_oracle as D

pytestmark = pytest.mark.gpu
FIX = os.path.join(ROOT, "tests", "golden", "datasplit", "select.pt")


def _call(name, *args):
    from npf_b200 import _cabi
    _cabi.call(name, *args, torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("B,N,n,seed", [(3, 10, 4, 12345), (256, 128, 50, 1), (5, 128, 128, 7), (4, 1024, 300, (1 << 62) + 12345),
                                        (2, 12288, 1000, 3), (3, 17, 0, 5), (1, 1, 1, 9)])
def test_random_subset_and_mask_match_oracle(B, N, n, seed):
    idx = torch.full((B, max(n, 1)), -7, dtype=torch.int32, device="cuda")[:, :n].contiguous()
    _call("npf_random_subset", idx.data_ptr(), B, N, n, seed)
    want = D.random_subset(B, N, n, seed)
    assert np.array_equal(idx.cpu().numpy(), want)
    mask = torch.full((B, N), 3, dtype=torch.uint8, device="cuda")
    _call("npf_random_mask", mask.data_ptr(), B, N, n, seed)
    assert np.array_equal(mask.cpu().numpy(), D.random_mask(B, N, n, seed))


def test_subset_argument_errors():
    idx = torch.empty(2, 4, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        _call("npf_random_subset", idx.data_ptr(), 2, 3, 4, 0)           # n > N
    with pytest.raises(NotImplementedError):
        _call("npf_random_subset", idx.data_ptr(), 2, 20000, 4, 0)       # beyond one CTA's shared memory


def test_select_and_grid_select_match_reference_vectors():
    from npf_b200.utils.datasplit import CntxtTrgtGetter, GridCntxtTrgtGetter, get_all_indcs
    for c in torch.load(FIX, weights_only=False):
        if c["kind"] == "points":
            X, Y = c["X"].cuda(), c["Y"].cuda()
            getter = CntxtTrgtGetter(is_add_cntxts_to_trgts=c["add"])
            Xc, Yc, Xt, Yt = getter(X, Y, context_indcs=c["context_indcs"].cuda(), target_indcs=get_all_indcs(*X.shape[:2]))
        elif c["kind"] == "grid":
            img = c["img"].cuda()
            getter = GridCntxtTrgtGetter(upscale_factor=c["upscale"])
            Xc, Yc, Xt, Yt = getter(img, context_mask=c["context_mask"].cuda(),
                                    target_mask=torch.ones_like(c["context_mask"]).cuda())
        else:
            continue
        for got, key in ((Xc, "X_cntxt"), (Yc, "Y_cntxt"), (Xt, "X_trgt"), (Yt, "Y_trgt")):
            assert got.shape == c[key].shape and torch.equal(got.cpu(), c[key]), (c["kind"], key)


def test_getters_end_to_end_seeded():
    """Seeded like upstream's set_seed: the context size follows python's ``random``, the indices the oracle's stream under
    the key numpy hands out; the targets are the inputs themselves (no copy); a model trains on the result."""
    import npf_b200
    from npf_b200.utils import datasplit as ds
    B, N = 16, 128
    g = torch.Generator().manual_seed(0)
    X = (torch.rand(B, N, 1, generator=g) * 2 - 1).cuda()
    Y = torch.randn(B, N, 1, generator=g).cuda()
    getter = ds.CntxtTrgtGetter(contexts_getter=ds.GetRandomIndcs(a=0.0, b=50), targets_getter=ds.get_all_indcs)
    random.seed(5); np.random.seed(5)
    Xc, Yc, Xt, Yt = getter(X, Y)
    random.seed(5); np.random.seed(5)
    n = ds.GetRandomIndcs(a=0.0, b=50).n_indcs(N)
    want = D.random_subset(B, N, n, ds._draw_seed())
    assert Xc.shape == (B, n, 1) and Xt.data_ptr() == X.data_ptr() and Yt.data_ptr() == Y.data_ptr()
    bi = np.arange(B)[:, None]
    assert np.array_equal(Xc.cpu().numpy(), X.cpu().numpy()[bi, want]) and np.array_equal(Yc.cpu().numpy(), Y.cpu().numpy()[bi, want])
    model = npf_b200.ConvCNP(1, 1).cuda().train()
    loss = npf_b200.CNPFLoss()(model(Xc, Yc, Xt, Yt), Yt)
    loss.backward()
    assert torch.isfinite(loss)
    # grids: masks only (what GridConvCNP consumes), same count in every row, exactly the oracle's mask
    gg = ds.GridCntxtTrgtGetter(context_masker=ds.RandomMasker(a=0.0, b=0.3), target_masker=ds.no_masker)
    img = torch.rand(4, 3, 32, 32, generator=g).cuda()
    random.seed(6); np.random.seed(6)
    m_c, img_last, m_t, _ = gg(img, is_return_masks=True)
    random.seed(6); np.random.seed(6)
    n = ds.RandomMasker(a=0.0, b=0.3).n_indcs(1024)
    want = D.random_mask(4, 1024, n, ds._draw_seed())
    assert m_c.dtype == torch.bool and m_c.shape == (4, 32, 32, 1) and m_t.shape == (4, 32, 32, 1) and bool(m_t.all())
    assert np.array_equal(m_c.view(4, -1).cpu().numpy(), want.astype(bool))
    gm = npf_b200.GridConvCNP(1, 3).cuda().train()
    loss = npf_b200.CNPFLoss()(gm(m_c, img_last, m_t, img_last), img_last)
    loss.backward()
    assert torch.isfinite(loss)
    # ragged masks are an error, not silently truncated rows
    bad = m_c.clone()
    bad[1].view(-1)[:] = False
    with pytest.raises(ValueError):
        gg.select(img_last, None, bad)
