"""CPU: pins oracle/datasplit_oracle.py (the checker of the on-device context / target split) --
  * Philox-4x32-10 against its published known-answer vectors,
  * the subset sampler against the properties the reference's collate guarantees (duplicate-free subsets of the requested
    size, independent rows, uniform inclusion and uniform order),
  * the __host__ build of the very shuffle routine the CUDA kernel runs (oracle/_ref/check_datasplit_host, compiled by
    __graft_entry__.build()) against the oracle, index for index,
  * gather / mask / grid-select against vectors produced by the REAL reference (oracle/gen_golden_datasplit.py),
  * the host-side draw of the context SIZE against the sequence the reference produces after the same seeding."""
import os
import random
import subprocess

import numpy as np
import pytest
import torch

from _util import ROOT
from oracle import datasplit_oracle as D

FIX = os.path.join(ROOT, "tests", "golden", "datasplit", "select.pt")
HOST_BIN = os.path.join(ROOT, "oracle", "_ref", "check_datasplit_host")


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(D.philox4x32_10(ctr, key)) == want


@pytest.mark.parametrize("B,N,n", [(5, 1, 1), (4, 10, 0), (6, 10, 10), (7, 33, 5)])
def test_subsets_are_valid(B, N, n):
    idx = D.random_subset(B, N, n, seed=99)
    assert idx.shape == (B, n) and idx.dtype == np.int32
    for row in idx:
        assert len(set(row.tolist())) == n and all(0 <= v < N for v in row)
    if n == N:
        assert all(sorted(r.tolist()) == list(range(N)) for r in idx)
    m = D.random_mask(B, N, n, seed=99)
    assert (m.sum(1) == n).all() and all(set(np.flatnonzero(m[b]).tolist()) == set(idx[b].tolist()) for b in range(B))
    assert not np.array_equal(D.random_subset(B, N, n, seed=100), idx) or n in (0,) or N == 1   # the seed matters
    if B > 1 and 0 < n < N:
        assert any(not np.array_equal(idx[0], idx[b]) for b in range(1, B))                      # rows are independent


def test_subset_uniformity():
    """4000 rows of 5-out-of-16: inclusion counts and first-position counts against the uniform law (chi-square, 15 dof:
    the 0.999 quantile is 37.7)."""
    B, N, n = 4000, 16, 5
    idx = D.random_subset(B, N, n, seed=2024)
    incl = np.bincount(idx.reshape(-1), minlength=N)
    first = np.bincount(idx[:, 0], minlength=N)
    chi_incl = ((incl - B * n / N) ** 2 / (B * n / N * (1 - n / N))).sum()   # inclusion indicators have variance p(1-p)
    chi_first = ((first - B / N) ** 2 / (B / N)).sum()
    assert chi_incl < 37.7 and chi_first < 37.7, (chi_incl, chi_first)


@pytest.mark.parametrize("B,N,n,seed", [(3, 10, 4, 12345), (5, 128, 128, 7), (2, 1024, 300, (1 << 62) + 12345)])
def test_cuda_source_host_build_matches_oracle(B, N, n, seed):
    if not os.path.exists(HOST_BIN):
        pytest.skip("oracle/_ref/check_datasplit_host not built (python __graft_entry__.py build)")
    out = subprocess.run([HOST_BIN, str(B), str(N), str(n), str(seed)], capture_output=True, text=True, check=True).stdout
    got = np.array([[int(v) for v in line.split()] for line in out.strip().splitlines()], dtype=np.int32)
    assert np.array_equal(got, D.random_subset(B, N, n, seed))


def test_select_and_grid_select_match_reference_vectors():
    cases = torch.load(FIX, weights_only=False)
    n_checked = 0
    for c in cases:
        if c["kind"] == "points":
            B, N = c["X"].shape[:2]
            ci = c["context_indcs"].numpy()
            assert np.array_equal(ci, D.random_subset(B, N, ci.shape[1], c["seed"]))
            Xc, Yc = D.select_points(c["X"].numpy(), c["Y"].numpy(), ci)
            assert np.array_equal(Xc, c["X_cntxt"].numpy()) and np.array_equal(Yc, c["Y_cntxt"].numpy())
            assert np.array_equal(c["X_trgt"].numpy(), c["X"].numpy()) and np.array_equal(c["Y_trgt"].numpy(), c["Y"].numpy())
        elif c["kind"] == "grid":
            img = c["img"].numpy()
            img_last = np.moveaxis(img, 1, -1)
            mask = c["context_mask"].numpy()[..., 0]
            B = mask.shape[0]
            assert np.array_equal(mask.reshape(B, -1), D.random_mask(B, mask[0].size, c["n"], c["seed"]).astype(bool))
            Xc, Yc = D.grid_select(mask, img_last, c["upscale"])
            assert np.array_equal(Xc, c["X_cntxt"].numpy()) and np.array_equal(Yc, c["Y_cntxt"].numpy())
            Xt, Yt = D.grid_select(np.ones_like(mask), img_last, c["upscale"])
            assert np.array_equal(Xt, c["X_trgt"].numpy()) and np.array_equal(Yt, c["Y_trgt"].numpy())
        else:
            continue
        n_checked += 1
    assert n_checked == 8


def test_context_sizes_follow_the_reference_sequence():
    """After the same seeding, the HOST part of the getters (how many points) yields the reference's sequence: python's
    ``random`` is consumed exactly as upstream consumes it (the device key comes from numpy's RNG)."""
    from npf_b200.utils import datasplit as ds
    c = [c for c in torch.load(FIX, weights_only=False) if c["kind"] == "sizes"][0]
    random.seed(c["seed"]); np.random.seed(c["seed"])
    getters = [ds.GetRandomIndcs(a=0.0, b=50), ds.GetRandomIndcs(a=0.1, b=0.5), ds.GetRandomIndcs(a=3, b=3, is_ensure_one=True)]
    for g, want in zip(getters, c["sizes"][:3]):
        got = []
        for _ in want:
            got.append(g.n_indcs(128))
            ds._draw_seed()
        assert got == want
    m = ds.RandomMasker(a=0.0, b=0.3)
    got = []
    for _ in c["sizes"][3]:
        got.append(m.n_indcs(32 * 32))
        ds._draw_seed()
    assert got == c["sizes"][3]
    with pytest.raises(RuntimeError):
        ds.GetRandomIndcs()(2, 10, device="cpu")           # no CPU path
    with pytest.raises(ValueError):
        ds.ratio_to_int(-0.1, 10)
