"""Sharding of LD blocks (and the SNPs they cover) across ranks.

The P x P coupling of the model is per SNP, so every cohort's block containing SNP i must
live on the rank that owns i: shard on connected components of the union of the cohorts'
block partitions, balanced by LD bytes with LPT.  SNPs in no block of
any cohort carry no LD work and are dealt out to even up SNP counts.
"""
import numpy as np


def block_cost(n, r):
    """Bytes one mat-vec reads for a block (packed dense vs two factor passes)."""
    dense = 4 * n * (n + 1) if n <= 2816 else 8 * n * n
    return float(min(dense, 16 * n * r))


def _find(parent, i):
    root = i
    while parent[root] != root:
        root = parent[root]
    while parent[i] != root:
        parent[i], i = root, parent[i]
    return root


def partition_snps(block_lists, M, world):
    """block_lists[p] = list of (snp_index_array, cost) for cohort p.

    Returns a list (one per rank) of sorted global SNP index arrays.
    """
    if world == 1:
        return [np.arange(M, dtype=np.int64)]
    parent = np.arange(M, dtype=np.int64)
    for blocks in block_lists:
        for snps, _ in blocks:
            if len(snps) == 0:
                continue
            r0 = _find(parent, int(snps[0]))
            for i in snps[1:]:
                ri = _find(parent, int(i))
                if ri != r0:
                    parent[ri] = r0
    roots = np.array([_find(parent, i) for i in range(M)], dtype=np.int64)
    cost = np.zeros(M)
    in_block = np.zeros(M, dtype=bool)
    for blocks in block_lists:
        for snps, c in blocks:
            if len(snps):
                cost[roots[int(snps[0])]] += c
                in_block[snps] = True
    comp_roots = np.unique(roots[in_block])
    # LPT: heaviest component first onto the lightest rank (ties -> lowest rank, deterministic)
    order = comp_roots[np.argsort(-cost[comp_roots], kind='stable')]
    load = np.zeros(world)
    owner_of_root = {}
    for root in order:
        r = int(np.argmin(load))
        owner_of_root[int(root)] = r
        load[r] += cost[root]
    owner = np.full(M, -1, dtype=np.int64)
    for i in np.where(in_block)[0]:
        owner[i] = owner_of_root[int(roots[i])]
    counts = np.array([(owner == r).sum() for r in range(world)], dtype=np.int64)
    for i in np.where(~in_block)[0]:
        r = int(np.argmin(counts))
        owner[i] = r
        counts[r] += 1
    return [np.where(owner == r)[0].astype(np.int64) for r in range(world)]


def host_block_lists(ld_mats):
    """(snps, cost) per block for host BlockDiagonalMatrix objects."""
    out = []
    for ld in ld_mats:
        blocks = []
        for b, m in enumerate(ld.matrices):
            snps = np.asarray(ld.perm[ld.starts[b]:ld.starts[b + 1]], dtype=np.int64)
            blocks.append((snps, block_cost(m.u.shape[0], m.u.shape[1])))
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
