"""The conflict-code algorithm of wave_deps_kernel (sgdnet_b200/csrc/saga_sparse.cu), restated in numpy: one table of hashed
feature buckets x staged-row bits answers "which earlier rows of my window MAY hold this feature" with two lookups; the
candidates, nearest first, are confirmed by search in the sorted index runs. The test checks the restatement against the
brute-force definition (nearest earlier row within the window that holds the feature, and the feature's position there) -
i.e. that false positives of the table cost only a search and false negatives cannot happen."""
import numpy as np

BUCKETS = 8192


def h1(j):
    return ((np.uint32(j) * np.uint32(2654435761)) & np.uint32(0xFFFFFFFF)) >> np.uint32(19)


def h2(j):
    return ((np.uint32(j) * np.uint32(0x85EBCA6B) + np.uint32(0x27D4EB2F)) & np.uint32(0xFFFFFFFF)) >> np.uint32(19)


def codes_by_table(rows, window):
    """rows: list of sorted int arrays (staged rows, in sequence order). Returns per row a list of (distance, position) or None."""
    table = np.zeros(BUCKETS, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for r, idx in enumerate(rows):
            for j in idx:
                table[h1(j)] |= np.uint32(1 << r)
                table[h2(j)] |= np.uint32(1 << r)
        out = []
        for me, idx in enumerate(rows):
            lo = max(0, me - window)
            wmask = ((1 << me) - 1) & ~((1 << lo) - 1)
            res = []
            for j in idx:
                cand = int(table[h1(j)] & table[h2(j)]) & wmask
                hit = None
                while cand:
                    pr = cand.bit_length() - 1          # nearest predecessor first
                    cand &= ~(1 << pr)
                    pos = int(np.searchsorted(rows[pr], j))
                    if pos < len(rows[pr]) and rows[pr][pos] == j:
                        hit = (me - pr, pos)
                        break
                res.append(hit)
            out.append(res)
    return out


def codes_brute_force(rows, window):
    out = []
    for me, idx in enumerate(rows):
        res = []
        for j in idx:
            hit = None
            for d in range(1, min(window, me) + 1):
                where = np.nonzero(rows[me - d] == j)[0]
                if where.size:
                    hit = (d, int(where[0]))
                    break
            res.append(hit)
        out.append(res)
    return out


def test_bucket_table_equals_brute_force():
    rng = np.random.default_rng(0)
    for p, nnz, window in ((300, 40, 7), (5000, 100, 7), (100000, 100, 11), (64, 30, 3)):
        for trial in range(6):
            rows = [np.sort(rng.choice(p, size=min(p, int(rng.integers(0, nnz + 1))), replace=False)).astype(np.int64) for _ in range(32)]
            if trial == 0:
                rows[5] = rows[4].copy()            # a repeated row: every feature conflicts at distance 1
            got, ref = codes_by_table(rows, window), codes_brute_force(rows, window)
            assert got == ref
