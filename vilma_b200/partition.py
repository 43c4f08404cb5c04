"""Sharding of LD blocks (and the SNPs they cover) across ranks.

The P x P coupling of the model is per SNP, so every cohort's block containing SNP i must
live on the rank that owns i: shard on connected components of the union of the cohorts'
block partitions, balanced by LD bytes with LPT.  SNPs in no block of
any cohort carry no LD work and are dealt out to even up SNP counts.
"""
import numpy as np


def block_cost(n, r):
    """Bytes one mat-vec reads for a block (packed dense vs two factor passes)."""
    dense = 4 * n * (n + 1) if n <= 65528 else 8 * n * n      # symmetric-packed (csrc/ld_kernels.cuh)
    return float(min(dense, 16 * n * r))


def _components(block_lists, M):
    """Connected components of the union of the cohorts' block partitions: root[i] = smallest SNP index
    of i's component.  Min-label propagation, vectorised per sweep (np.minimum.reduceat over the
    blocks of a cohort); cohorts that share a block file converge in two sweeps."""
    root = np.arange(M, dtype=np.int64)
    packed = []
    for blocks in block_lists:
        blocks = [np.asarray(s, dtype=np.int64) for s, _ in blocks if len(s)]
        if not blocks:
            continue
        idx = np.concatenate(blocks)
        starts = np.concatenate([[0], np.cumsum([len(b) for b in blocks])[:-1]]).astype(np.int64)
        lens = np.array([len(b) for b in blocks], dtype=np.int64)
        packed.append((idx, starts, lens))
    while True:
        changed = False
        for idx, starts, lens in packed:
            low = np.minimum.reduceat(root[idx], starts)
            new = np.repeat(low, lens)
            if np.any(new < root[idx]):
                root[idx] = np.minimum(root[idx], new)
                changed = True
        # pointer jumping: labels always point at a SNP whose own label is <= theirs
        while True:
            nxt = root[root]
            if np.array_equal(nxt, root):
                break
            root = nxt
            changed = True
        if not changed:
            return root


def partition_snps(block_lists, M, world):
    """block_lists[p] = list of (snp_index_array, cost) for cohort p.

    Returns a list (one per rank) of sorted global SNP index arrays.
    """
    if world == 1:
        return [np.arange(M, dtype=np.int64)]
    roots = _components(block_lists, M)
    cost = np.zeros(M)
    in_block = np.zeros(M, dtype=bool)
    for blocks in block_lists:
        firsts = np.array([int(s[0]) for s, _ in blocks if len(s)], dtype=np.int64)
        costs = np.array([c for s, c in blocks if len(s)], dtype=np.float64)
        np.add.at(cost, roots[firsts], costs)
        for snps, _ in blocks:
            in_block[np.asarray(snps, dtype=np.int64)] = True
    comp_roots = np.unique(roots[in_block])
    # LPT: heaviest component first onto the lightest rank (ties -> lowest rank, deterministic)
    order = comp_roots[np.argsort(-cost[comp_roots], kind='stable')]
    load = np.zeros(world)
    owner_of_root = np.full(M, -1, dtype=np.int64)
    for root in order:
        r = int(np.argmin(load))
        owner_of_root[root] = r
        load[r] += cost[root]
    owner = np.where(in_block, owner_of_root[roots], -1)
    # SNPs in no block: dealt out to even up the SNP counts (lowest rank first on ties)
    counts = np.bincount(owner[in_block], minlength=world).astype(np.int64)
    free = np.where(~in_block)[0]
    if len(free):
        seq = _deal_sequence(counts, len(free))
        owner[free] = seq
    return [np.where(owner == r)[0].astype(np.int64) for r in range(world)]


def _deal_sequence(counts, n):
    """Rank of each of n SNPs dealt one by one to the rank that currently has the fewest (ties ->
    lowest rank): vectorised by levels instead of a Python loop per SNP."""
    c = np.asarray(counts, dtype=np.int64).copy()
    out = np.empty(n, dtype=np.int64)
    pos = 0
    while pos < n:
        lo = c.min()
        at_lo = np.where(c == lo)[0]
        higher = c[c > lo]
        levels = int(higher.min() - lo) if len(higher) else (n - pos + len(at_lo) - 1) // len(at_lo)
        levels = max(1, min(levels, (n - pos + len(at_lo) - 1) // len(at_lo)))
        chunk = np.tile(at_lo, levels)[:n - pos]
        out[pos:pos + len(chunk)] = chunk
        np.add.at(c, chunk, 1)
        pos += len(chunk)
    return out


def host_block_lists(ld_mats):
    """(snps, cost) per block for host BlockDiagonalMatrix objects."""
    out = []
    for ld in ld_mats:
        blocks = []
        for b, m in enumerate(ld.matrices):
            snps = np.asarray(ld.perm[ld.starts[b]:ld.starts[b + 1]], dtype=np.int64)
            n = m.shape[0]
            r = m.u.shape[1] if getattr(m, 'factorized', True) else n      # (lazy dense blocks: full rank assumed)
            blocks.append((snps, block_cost(n, r)))
        out.append(blocks)
    return out


def local_blocks(ld, snps, M):
    """Blocks of `ld` owned by the rank holding `snps`; returns (block_ids, perm_local)."""
    g2l = np.full(M, -1, dtype=np.int64)
    g2l[snps] = np.arange(len(snps))
    ids, perm = [], []
    for b in range(len(ld.matrices)):
        idx = np.asarray(ld.perm[ld.starts[b]:ld.starts[b + 1]], dtype=np.int64)
        loc = g2l[idx]
        if loc[0] >= 0:
            if np.any(loc < 0):
                raise RuntimeError('LD block split across ranks')
            ids.append(b)
            perm.append(loc)
    perm = np.concatenate(perm) if perm else np.zeros(0, dtype=np.int64)
    return ids, perm
