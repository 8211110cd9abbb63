"""Hierarchical block permutation keys (reference: keynet/blockpermute.py:6-79).

Integer host logic.  The permutation must consume numpy's global legacy RNG in exactly the
reference's order (two draws per permuted level: row order, then column order; crops visited
row-major; recursion depth-first) so that the same seed yields the same key.
"""
import numpy as np

from .util import find_closest_positive_divisor


def block_permute(img, cropshape, seed=None):
    """Scatter every non-overlapping (ch,cw) crop of an HxWxC image to a random crop position; rows and
    columns of the crop grid are permuted independently (blockpermute.py:6-20)."""
    (H, W) = img.shape[0:2]
    (ch, cw) = cropshape
    assert H % ch == 0 and W % cw == 0, "Blocksize must be evenly divisible with image shape"
    if seed is not None:
        np.random.seed(seed)
    dst_rows = np.random.permutation(np.arange(0, H, ch)) // ch     # crop-grid row a lands on row dst_rows[a]
    dst_cols = np.random.permutation(np.arange(0, W, cw)) // cw
    grid = img.reshape(H // ch, ch, W // cw, cw, -1)
    out = np.empty_like(grid)
    tmp = np.empty_like(grid)
    tmp[dst_rows] = grid
    out[:, :, dst_cols] = tmp
    return out.reshape(img.shape)


def hierarchical_block_permute(img, blockshape, permute_at_level, min_blocksize=8, seed=None, twist=False, strict=True):
    """Top-down recursive block shuffle (or 90-degree twist) of an HxWxC image (blockpermute.py:23-68)."""
    levels = list(np.asarray(permute_at_level).reshape(-1))
    if len(levels) == 0 or tuple(blockshape) == tuple(img.shape):
        return np.copy(img)
    if img.shape[0] % blockshape[0] != 0 and img.shape[1] % blockshape[1] != 0:
        if strict:
            raise ValueError("Recursive image size %s and block layout %s must be divisible" % (str(img.shape[0:2]), str(blockshape)))
        blockshape = (find_closest_positive_divisor(img.shape[0], blockshape[0]), find_closest_positive_divisor(img.shape[1], blockshape[1]))
    cropshape = (img.shape[0] // blockshape[0], img.shape[1] // blockshape[1])
    out = np.copy(img)
    if seed is not None:
        np.random.seed(seed)
    if 0 in levels:
        if twist:
            out = np.rot90(out, k=1 if np.random.rand() > 0.5 else 3)
        else:
            out = block_permute(out, cropshape, seed=None)
    if len(levels) == 1 and levels[0] == 0:
        return out
    deeper = max(levels) > 0
    out = np.copy(out)   # rot90 returns a view
    for i in range(0, img.shape[0], cropshape[0]):
        for j in range(0, img.shape[1], cropshape[1]):
            if min(cropshape) >= min_blocksize and deeper:
                sub = out[i:i + cropshape[0], j:j + cropshape[1]]
                out[i:i + cropshape[0], j:j + cropshape[1]] = hierarchical_block_permute(
                    sub, blockshape, permute_at_level=np.array(levels) - 1, seed=None, min_blocksize=min_blocksize, twist=twist)
            elif deeper:
                raise ValueError('Recursive blockshape=%s < minimum blockshape=%d' % (str(cropshape), min_blocksize))
    return out


def hierarchical_block_permutation_matrix(imgshape, blockshape, permute_at_level, min_blocksize=8, seed=None, twist=False, withinverse=False, strict=True):
    """Key P with P.dot(img.flatten()).reshape(imgshape) == hierarchical_block_permute(img, ...) for an
    HxWxC image (blockpermute.py:71-79), returned as a MonomialKey."""
    from .sparse import MonomialKey
    idx = np.arange(int(np.prod(imgshape))).reshape(imgshape)
    P = MonomialKey(hierarchical_block_permute(idx, blockshape, permute_at_level, min_blocksize, seed=seed, twist=twist, strict=strict).flatten())
    return P if not withinverse else (P, P.transpose())
