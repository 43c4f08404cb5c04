"""Loading of variants, annotations, summary statistics and LD schemas.

Host-side restatement of ``vilma.load`` (/root/reference/src/vilma/load.py) so that
``vilma fit`` stays a drop-in: same file formats, same matching /
allele-flip / denylist semantics, same return values.  These are the *inputs to* the hot
path, not the hot path: pandas on the host, as in the reference.  (``--mmap`` is accepted
and ignored: LD factors live in HBM, never on disk.)
"""
import logging
from pathlib import Path

import numpy as np
import pandas as pd

from .matrix_structures import BlockDiagonalMatrix, LowRankMatrix


def _read_table(path, **kwargs):
    return pd.read_csv(path, sep=r'\s+', **kwargs)


def load_variant_list(variant_filename):
    """Read the list of variants to analyse (load.py:21-39)."""
    variants = _read_table(variant_filename, header=0).drop_duplicates(ignore_index=True)
    if 'ID' not in variants.columns:
        raise ValueError('Variant file must contain a column labeled ID')
    if 'A1' not in variants.columns:
        raise ValueError('Variant file must contain a column labeled A1')
    if 'A2' not in variants.columns:
        if 'REF' not in variants.columns or 'ALT' not in variants.columns:
            raise ValueError('Variant file must contain a column labeled A2')
        variants['A2'] = variants['REF'].copy()
        flip = variants['A1'] == variants['REF']
        variants.loc[flip, 'A2'] = variants.loc[flip, 'ALT'].copy()
    return variants[['ID', 'A1', 'A2']].copy()


def load_annotations(annotations_filename, variants):
    """One-hot annotations aligned to `variants`, plus the un-annotated rows (load.py:42-68)."""
    if not annotations_filename:
        return np.ones((variants.shape[0], 1)), []
    dframe = _read_table(annotations_filename, header=0)
    if 'ID' not in dframe.columns:
        raise ValueError('Annotation file must contain a column labeled ID')
    if 'ANNOTATION' not in dframe.columns:
        raise ValueError('Annotation file must contain a column labeled ANNOTATION')
    dframe = pd.merge(variants, dframe, on='ID', how='left')
    dframe = pd.DataFrame(dframe['ANNOTATION'])
    n_missing = int(dframe['ANNOTATION'].isna().sum())
    if n_missing > 0:
        logging.warning('%d out of %d total variants are missing annotations. '
                        'These will get set to having the first annotation!',
                        n_missing, dframe.shape[0])
    denylist = np.where(dframe['ANNOTATION'].isna())[0].tolist()
    dframe.loc[dframe['ANNOTATION'].isna(), 'ANNOTATION'] = 0
    return pd.get_dummies(dframe['ANNOTATION'], dummy_na=False).to_numpy(), denylist


def load_sumstats(sumstats_filename, variants):
    """GWAS summary statistics aligned to `variants` (load.py:71-139).

    Missing or allele-mismatched variants get BETA 0 / SE 1 and are reported; flipped alleles
    negate BETA; an OR column is converted to a log odds ratio."""
    header = _read_table(sumstats_filename, nrows=1, header=0)
    if 'ID' not in header.columns:
        raise ValueError('Summary Statistics File must contain a column labeled ID')
    if 'A1' not in header.columns:
        raise ValueError('Summary Statistics File must contain a column labeled A1')
    a2_cols = ['A2']
    if 'A2' not in header.columns:
        a2_cols = ['REF', 'ALT']
        if 'REF' not in header.columns or 'ALT' not in header.columns:
            raise ValueError('If summary statistics file does not contain a column labeled '
                             'A2, then it must contain REF and ALT columns.')
    if 'SE' not in header.columns:
        raise ValueError('Summary Statistics File must contain a column labeled SE')
    effect_col = 'BETA'
    if 'BETA' not in header.columns:
        effect_col = 'OR'
        if 'OR' not in header.columns:
            raise ValueError('Summary stat file needs to contain eitherBETA or OR filed.')

    sumstats = _read_table(sumstats_filename, header=0,
                           usecols=['ID', 'A1', 'SE', effect_col] + a2_cols)
    sumstats = sumstats[sumstats.ID.isin(variants.ID)].reset_index(drop=True)
    if 'A2' not in sumstats.columns:
        sumstats['A2'] = sumstats['REF'].copy()
        flip = sumstats['A1'] == sumstats['REF']
        sumstats.loc[flip, 'A2'] = sumstats.loc[flip, 'ALT'].copy()
    if 'BETA' not in sumstats.columns:
        sumstats['BETA'] = np.log(sumstats.OR)
    sumstats['BETA'] = sumstats['BETA'].astype(np.float64)
    sumstats['SE'] = sumstats['SE'].astype(np.float64)

    sumstats = pd.merge(variants, sumstats, on='ID', how='left')
    stay_allele = ((sumstats.A1_x == sumstats.A1_y) & (sumstats.A2_x == sumstats.A2_y))
    flip_allele = ((sumstats.A1_x == sumstats.A2_y) & (sumstats.A1_y == sumstats.A2_x))
    missing = (sumstats.BETA.isna() | sumstats.SE.isna() | ((~stay_allele) & (~flip_allele)))
    logging.warning('%d out of %d total variants are missing sumstats',
                    missing.sum(), sumstats.shape[0])
    logging.warning('%d alleles have been flipped', (flip_allele).sum())
    sumstats.loc[missing, 'BETA'] = 0.
    sumstats.loc[missing, 'SE'] = 1.
    sumstats.loc[flip_allele, 'BETA'] = -sumstats.loc[flip_allele, 'BETA']
    return sumstats, np.where(missing)[0].tolist()


def schema_iterator(schema_path):
    """Yield (.var path, .npy path) per LD block of a schema manifest (load.py:142-163)."""
    schema_path = Path(schema_path)
    with open(schema_path, 'r') as schema:
        for line in schema:
            if not line.strip():
                continue
            snp_path, ld_path = line.split()
            yield Path(schema_path.parents[0], snp_path), Path(schema_path.parents[0], ld_path)


def load_ld_mat(ld_path, variant_indices=None, mismatch=None, signs=None):
    """One block of an LD schema as a dense matrix restricted to the kept SNPs (load.py:166-234).

    A square .npy is the correlation matrix itself; a tall (n+1) x r one stacks the
    eigenvectors (rows 0..n-1) on the eigenvalues (last row); a 0-d one is a 1x1 block."""
    ld_matrix = np.load(ld_path)
    if signs is not None and not np.allclose(np.asarray(signs)**2, 1):
        raise ValueError('signs must be a vector consisting entirely of +1s and -1s.')
    if len(ld_matrix.shape) == 0:
        return ld_matrix[None, None]
    num_snps = ld_matrix.shape[0]
    if ld_matrix.shape[0] > ld_matrix.shape[1]:
        num_snps -= 1
    if variant_indices is None:
        variant_indices = np.ones(num_snps, dtype=bool)
    if mismatch is None:
        mismatch = np.zeros(variant_indices.sum(), dtype=bool)
    if signs is None:
        signs = np.ones(int(variant_indices.sum()))
    if ld_matrix.shape[0] == ld_matrix.shape[1]:
        sub = np.copy(ld_matrix[np.ix_(variant_indices, variant_indices)])
        sub = sub * np.outer(signs, signs)
        return sub[np.ix_(~mismatch, ~mismatch)]
    if ld_matrix.shape[0] < ld_matrix.shape[1]:
        raise ValueError('Bad LD matrix.')
    if num_snps != variant_indices.shape[0]:
        raise ValueError('Bad LD matrix.')
    u_mat = np.copy(ld_matrix[0:num_snps])
    s_vec = np.copy(ld_matrix[num_snps])
    u_mat = u_mat[variant_indices, :]
    u_mat = signs.reshape((-1, 1)) * u_mat
    u_mat = np.copy(u_mat[~mismatch])
    return (u_mat * s_vec).dot(u_mat.T)


_MMAP_CHARS = None


def _consume_mmap_rng():
    """With --mmap the reference names two HDF5 datasets per block with 100 random characters
    each, drawn from the global NumPy stream (matrix_structures.py:31-35, :124-132).  We store
    nothing on disk but draw the same numbers so later seeded draws stay aligned."""
    global _MMAP_CHARS
    if _MMAP_CHARS is None:
        import string
        _MMAP_CHARS = list(string.ascii_letters + string.digits)
    np.random.choice(_MMAP_CHARS, size=100)
    np.random.choice(_MMAP_CHARS, size=100)


def _make_block(accepted, ldthresh):
    return LowRankMatrix(accepted, ldthresh, lazy=ldthresh >= 1.0)


def load_ld_from_schema(schema_path, variants, denylist, ldthresh, mmap=False):
    """Block-diagonal LD of a schema, matched and oriented to `variants` (load.py:237-354).

    Returns (BlockDiagonalMatrix in the order of `variants`, list of positions without LD)."""
    from ._pool import OrderedPipeline
    pipe = OrderedPipeline()        # one eigh per block, run concurrently (results in block order)
    perm = []
    var_reidx = variants.set_index('ID')
    var_reidx['old_idx'] = np.arange(var_reidx.shape[0])
    var_a1 = variants['A1'].to_numpy()
    var_a2 = variants['A2'].to_numpy()
    total_flipped = 0
    try:
        for snp_path, ld_path in schema_iterator(schema_path):
            snp_metadata = _read_table(snp_path, header=None,
                                       names=['ID', 'CHROM', 'BP', 'CM', 'A1', 'A2'])
            logging.info('LD matrix shape: %s', ((snp_metadata.shape[0], snp_metadata.shape[0]),))
            variant_indices = np.array(snp_metadata.ID.isin(variants.ID).to_numpy(), dtype=bool)
            if np.sum(variant_indices) == 0:
                continue
            kept_ids = snp_metadata.ID[variant_indices]
            idx = np.array(var_reidx.loc[kept_ids].old_idx.to_numpy()).flatten()
            keep = np.isin(idx, denylist, invert=True)
            to_change = np.where(variant_indices)[0][~keep]
            variant_indices[to_change] = False
            logging.info('Proportion of variant indices being used: %e', np.mean(variant_indices))
            idx = idx[keep]
            if len(idx) == 0:
                continue
            ld_a1 = snp_metadata['A1'].to_numpy()[variant_indices]
            ld_a2 = snp_metadata['A2'].to_numpy()[variant_indices]
            stay = np.array([(x1 == y1) and (x2 == y2) for x1, y1, x2, y2 in
                             zip(var_a1[idx], ld_a1, var_a2[idx], ld_a2)], dtype=bool)
            flip = np.array([(x1 == y2) and (x2 == y1) for x1, y1, x2, y2 in
                             zip(var_a1[idx], ld_a1, var_a2[idx], ld_a2)], dtype=bool)
            total_flipped += flip.sum()
            mismatch = np.logical_and(~flip, ~stay)
            if len(idx[~mismatch]) == 0:
                continue
            signs = np.ones(len(idx))
            signs[flip] = -1
            accepted = load_ld_mat(ld_path, variant_indices, mismatch, signs)
            perm.append(idx[~mismatch])
            if mmap:
                _consume_mmap_rng()
            # --ldthresh 1 (no truncation): the eigendecomposition is postponed -- the GPU set-up
            # (BlockDiagonalMatrix.device_setup) only needs it for blocks that are not safely full rank
            pipe.submit(_make_block, accepted, ldthresh)

        svds = pipe.results()
    finally:
        pipe.close()               # also on a bad schema: workers and the BLAS thread limit are released
    num = variants.shape[0]
    perm = np.concatenate(perm) if len(perm) > 0 else np.array([], dtype=np.int64)
    list_of_missing = sorted(set(range(num)) - set(perm.tolist()))
    missing = np.array(list_of_missing, dtype=np.int64)
    logging.info('Loaded a total of %d variants.', num)
    logging.warning('Missing LD info for %d variants. They will be ignored during '
                    'optimization.', len(missing))
    logging.warning('The alleles did not match for %d variants. They were flipped',
                    total_flipped)
    perm = np.concatenate([perm, missing]).astype(np.int64)
    if not np.all(perm == np.arange(len(perm))):
        logging.warning('The variants in the extract file and the variants in the LD matrix '
                        'were not in the same order.  The variants in the LD matrix have been '
                        'reordered to match the extract file.')
    return BlockDiagonalMatrix(svds, perm=perm, missing=missing), list_of_missing
