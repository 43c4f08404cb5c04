"""TEST-ONLY stand-in for vilma_b200.engine.CudaEngine built on the oracle's NumPy numerics.

Lets the CPU test-suite drive the product's *host* logic (control flow of MultiPopVI, SNP
sharding, the all-reduce plumbing under gloo) without a GPU.  It is injected through the
``engine_factory=`` hook; the product never constructs it.
"""
import numpy as np

from oracle import numerics_np as nm
from vilma_b200.partition import local_blocks


class NumpyShardEngine:
    def __init__(self, vi, snps, pieces):
        self.K, self.P, self.M, self.A = pieces['K'], pieces['P'], pieces['M'], pieces['A']
        self.adj, self.se, self.sld = pieces['adj'], pieces['se'], pieces['sld']
        self.scal, self.ann = pieces['scalings'], pieces['annotations']
        self.prec = pieces['mixture_prec'][..., None]       # [K,P,P,1]
        self.log_det = pieces['log_det']
        self.tau = np.ones(self.P)
        self.blocks = []                                     # per cohort: [(block, local idx)]
        for ld in vi.ld_mats:
            ids, perm_local = local_blocks(ld, snps, vi.num_loci)
            out, off = [], 0
            for b in ids:
                n = ld.matrices[b].shape[0]
                out.append((ld.matrices[b], perm_local[off:off + n]))
                off += n
            self.blocks.append(out)
        self.cur = None
        self.trial = None
        self.prev = np.zeros((self.P, self.M))
        self.ckpt = np.zeros((self.P, self.M))

    # ---- small inputs
    def set_tau(self, tau):
        self.tau = np.array(tau, dtype=float)

    def set_hyper(self, hyper):
        self.hyper = np.array(hyper, dtype=float)

    def set_delta_grad(self, table):
        self.gtable = np.array(table, dtype=float)

    def set_params(self, mu, delta):
        self.cur = self._evaluate(np.array(mu), np.array(delta), need_stats=False)
        self.trial = None

    def get_params(self):
        return self.cur['mu'].copy(), self.cur['delta'].copy()

    # ---- maths
    def _cov(self):
        K, P, M = self.K, self.P, self.M
        lam = np.zeros((K, P, P, M))
        idx = np.arange(P)
        lam[:, idx, idx, :] = self.sld / self.tau[:, None]
        lam += self.prec
        S = nm.vi_sigma_inv(lam)
        c = nm.vi_sigma_log_det(S)                            # [K,M]
        match = np.einsum('kpq,kqpi->ik', self.prec[..., 0], S)
        return lam, S, c, self.log_det - c.T + match

    def _matvec(self, p, x):
        y = np.zeros(self.M)
        for blk, idx in self.blocks[p]:
            y[idx] = blk._host_dot(x[idx])
        return y

    def _evaluate(self, mu, delta, need_stats=True):
        lam, S, c, sigsum = self._cov()
        pm = nm.posterior_mean(mu, delta)
        pv = nm.pmv(pm, mu, delta, np.einsum('kppi->kpi', S))
        z = pm / self.se
        linked = np.stack([self._matvec(p, z[p]) for p in range(self.P)])
        st = np.zeros(3 * self.P + 3)
        P = self.P
        st[0:P] = (pm * self.adj).sum(axis=1)
        st[P:2 * P] = (self.sld * pv).sum(axis=1)
        st[2 * P:3 * P] = (z * linked).sum(axis=1)
        st[3 * P] = nm.delta_kl(delta, self.hyper, self.ann)
        st[3 * P + 1] = nm.inner_product_comp(mu, self.prec, delta)
        st[3 * P + 2] = nm.beta_kl(sigsum, delta)
        return dict(mu=mu, delta=delta, pm=pm, pv=pv, z=z, linked=linked, stats=st)

    def _softmax(self, mu, eta, c):
        return nm.invert_nat_vi_delta(mu, eta, c.T, self.gtable[self.ann])

    def eval(self):
        self.cur = self._evaluate(self.cur['mu'], self.cur['delta'])
        return self.cur['stats']

    def beta_trial(self, step):
        lam, S, c, _ = self._cov()
        cur = self.cur
        eta_old = np.einsum('kpqi,kqi->kpi', lam, cur['mu'])
        lk = cur['linked'] / self.se - cur['pm'] * self.sld
        g = (self.adj - lk) / self.tau[:, None]
        eta = nm.sum_betas(eta_old, g[None], step)
        mu = nm.nat_inner_product(eta, S)
        self.trial = self._evaluate(mu, self._softmax(mu, eta, c))
        return self.trial['stats']

    def refresh_delta(self):
        lam, S, c, _ = self._cov()
        mu = self.cur['mu']
        eta = np.einsum('kpqi,kqi->kpi', lam, mu)
        self.trial = self._evaluate(mu, self._softmax(mu, eta, c))
        return self.trial['stats']

    def accept(self):
        self.cur, self.trial = self.trial, None

    def sum_annotations(self):
        return nm.sum_annotations(self.cur['delta'], self.ann, self.A).reshape(-1)

    def posterior(self):
        return self.cur['pm'].copy(), self.cur['pv'].copy()

    def pm_diff(self, atol, rtol):
        v = self.cur['pm'] * self.scal
        o, c = self.prev, self.ckpt
        d, dc = np.abs(v - o), np.abs(v - c)
        out = np.array([np.sum(~(d <= atol + rtol * np.abs(o))), d.sum(), (d * d).sum(),
                        dc.sum(), (dc * dc).sum(), np.abs(v).max(),
                        np.abs((v - o) / (o + 1e-100)).max(), d.max(),
                        np.abs((v - c) / (c + 1e-100)).max(), dc.max()], dtype=float)
        self.prev = v.copy()
        return out

    def pm_mark(self, which):
        v = self.cur['pm'] * self.scal
        if which == 0:
            self.prev = v.copy()
        else:
            self.ckpt = v.copy()

    def vi_sigma(self, k0=0, k1=None, out=None):
        S = self._cov()[1][k0:k1]
        if out is None:
            return S
        out[...] = S
        return out

    @staticmethod
    def _host_array(shape):
        return np.empty(shape)
