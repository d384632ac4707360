"""CPU oracle for the on-device context / target split  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy integer arithmetic, the algorithm of ``neural-process-family_b200/csrc/datasplit.cu`` (which in turn
replaces the host-side collate of the reference, npf/utils/datasplit.py:108-145, 246-278, 423-452): the first ``n``
entries of a per-row partial Fisher-Yates shuffle driven by Philox-4x32-10.  Index / byte work: the CUDA kernels must
match these functions bit for bit (tests/test_gpu_datasplit.py).

Pinned by (tests/test_datasplit_oracle.py):
  * the published known-answer vectors of Philox-4x32-10 (Random123 ``kat_vectors``),
  * the properties the reference's collate guarantees -- every row a duplicate-free subset of range(N) of the requested
    size, rows independent, every index equally likely (chi-square against the reference's own ``GetRandomIndcs``
    run in the build container is not possible bit-wise: numpy's Mersenne-Twister stream cannot be reproduced on the
    device, so parity with the reference is distributional by construction; the gather / mask / grid-select steps ARE
    compared exactly with the reference's ``CntxtTrgtGetter.select`` / ``RandomMasker`` / ``GridCntxtTrgtGetter.select``
    semantics through fixtures generated from it, tests/golden/datasplit/select.pt).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (Salmon et al., SC'11; 10 rounds)."""
    c = [np.uint64(int(v) & 0xFFFFFFFF) for v in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for r in range(10):
        if r > 0:
            k0 = (k0 + W0) & 0xFFFFFFFF
            k1 = (k1 + W1) & 0xFFFFFFFF
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & MASK32, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & MASK32]
    return [int(v) for v in c]


def random_subset(B, N, n, seed):
    """[B, n] int32: per row the first n entries of the Philox-driven partial Fisher-Yates shuffle of arange(N)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((B, n), dtype=np.int32)
    for b in range(B):
        perm = list(range(N))
        r = None
        for i in range(n):
            if i % 4 == 0:
                r = philox4x32_10((i >> 2, b, 0, 0), key)
            j = i + ((r[i % 4] * (N - i)) >> 32)
            perm[i], perm[j] = perm[j], perm[i]
        out[b] = perm[:n]
    return out


def random_mask(B, P, n, seed):
    """[B, P] uint8 with ones at random_subset(B, P, n, seed) (RandomMasker, datasplit.py:259-278)."""
    m = np.zeros((B, P), dtype=np.uint8)
    idx = random_subset(B, P, n, seed)
    for b in range(B):
        m[b, idx[b]] = 1
    return m


def select_points(X, Y, indcs):
    """CntxtTrgtGetter.select, datasplit.py:246-255."""
    b = np.arange(X.shape[0])[:, None]
    return X[b, indcs], Y[b, indcs]


def grid_select(mask, img, upscale=1.0):
    """GridCntxtTrgtGetter.select, datasplit.py:423-452.  mask [B, *grid] (bool / bytes), img [B, *grid, y]
    -> X [B, n, n_grid_dim] (fp32, the same three roundings per coordinate as the in-place torch ops), Y [B, n, y]."""
    B, *grid = mask.shape
    y = img.shape[-1]
    n = int(np.count_nonzero(mask[0]))
    Xs, Ys = [], []
    for b in range(B):
        nz = np.argwhere(mask[b] != 0)[:n].astype(np.float32)  # row-major order, like Tensor.nonzero()
        for d, size in enumerate(grid):
            nz[:, d] = nz[:, d] * np.float32(2 / (size - 1))
            nz[:, d] = nz[:, d] - np.float32(1)
        nz = nz * np.float32(upscale)
        Xs.append(nz)
        Ys.append(img[b][mask[b] != 0][:n].reshape(-1, y))
    return np.stack(Xs), np.stack(Ys)
