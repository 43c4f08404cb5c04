import os, sys
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from _fixtures import build_ld, load_case, vi_kwargs
from vilma_b200.matrix_structures import BlockDiagonalMatrix, LowRankMatrix
from vilma_b200.variational_inference import MultiPopVI
for name in ['syn_p5', 'syn_p3']:
    fx = load_case(name)
    vi = MultiPopVI(ld_mats=build_ld(fx, LowRankMatrix, BlockDiagonalMatrix), **vi_kwargs(fx))
    np.random.seed(int(fx['seed']))
    mu, delta, hyper = vi._initialize()
    ref = fx['init_vi_delta']
    bad = ~np.isclose(delta, ref, rtol=1e-7, atol=1e-300)
    print(name, os.environ.get('VILMA_B200_LIB', 'default'), os.environ.get('VILMA_B200_OPTIONS', ''),
          'nan', np.isnan(delta).sum(), 'bad', bad.sum(), 'of', bad.size)
    idx = np.argwhere(bad)[:8]
    for i, k in idx:
        print('   ', i, k, delta[i, k], ref[i, k], 'row max', ref[i].max())
