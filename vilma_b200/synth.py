"""Synthetic benchmark inputs of BASELINE.json's shapes.

Deterministic (seeded) generator of HapMap3-like problems: LD block sizes, per-block dense
LD (sample correlation of AR(1) haplotype-like columns), and summary statistics drawn from
the reference's own generative model (sim.py:97-156: beta_hat = S R S^-1 beta + S R^1/2 eps).
Runs on any torch device: CUDA for the B200 arm (blocks are born in HBM and never touch
the host), CPU for the bounded sample the CPU baseline is timed on.

This is input scaffolding, not the fit path: torch library calls (randn, matmul, cholesky)
are used freely here and nowhere in the timed loop.
"""
import numpy as np


def block_sizes(M_ld, B, cv=0.6, seed=1234):
    """n_b ~ round(LogNormal(mean M/B, CV)), clipped to [32, 4 M/B], rescaled to sum to M_ld."""
    rng = np.random.default_rng(seed)
    mean = M_ld / B
    sigma2 = np.log(1 + cv * cv)
    raw = rng.lognormal(np.log(mean) - 0.5 * sigma2, np.sqrt(sigma2), size=B)
    raw = np.clip(raw, 32, 4 * mean)
    n = np.maximum(32, np.round(raw * (M_ld / raw.sum()))).astype(np.int64)
    # fix the rounding drift on the largest blocks
    diff = int(M_ld - n.sum())
    order = np.argsort(-n)
    i = 0
    while diff != 0:
        step = 1 if diff > 0 else -1
        n[order[i % B]] += step
        diff -= step
        i += 1
    return n


def assign_blocks(n, world):
    """LPT by n^2: list of block-id arrays per rank (deterministic)."""
    load = np.zeros(world)
    owner = np.zeros(len(n), dtype=np.int64)
    for b in np.argsort(-n, kind='stable'):
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += float(n[b]) ** 2
    return [np.where(owner == r)[0] for r in range(world)]


def make_block(n, seed, device, rho=0.95, noise=0.1, n_ref_factor=2.0):
    """(R [n,n], Gn [n_ref,n]) with R = Gn^T Gn the sample correlation; fp64 torch tensors."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    n_ref = max(2, int(np.ceil(n_ref_factor * n)))
    e = torch.randn((n_ref, n), generator=gen, device=device, dtype=torch.float64)
    idx = torch.arange(n, device=device, dtype=torch.float64)
    expo = idx[None, :] - idx[:, None]                       # j - i
    c = float(np.sqrt(1 - rho * rho))
    t = torch.where(expo >= 0, torch.pow(torch.tensor(rho, device=device, dtype=torch.float64),
                                         expo.clamp(min=0)), torch.zeros((), device=device,
                                                                         dtype=torch.float64))
    t[1:, :] *= c                                            # row 0 carries e_0 unscaled
    g = e @ t
    g += np.sqrt(noise) * torch.randn((n_ref, n), generator=gen, device=device,
                                      dtype=torch.float64)
    g -= g.mean(dim=0, keepdim=True)
    g /= torch.sqrt((g * g).sum(dim=0, keepdim=True))
    r = g.T @ g
    r = 0.5 * (r + r.T)
    r.fill_diagonal_(1.0)
    return r, g, gen


def block_se(n, seed, b, N_gwas, device):
    """Standard errors of block b's SNPs: SE = 1 / sqrt(N 2f(1-f)), f ~ U(0.05, 0.5)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 7919 + int(b) + 17)
    f = 0.05 + 0.45 * torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    return 1.0 / torch.sqrt(N_gwas * 2 * f * (1 - f))


def make_block_sumstats(n, seed, b, se, M, device, h2=0.3, n_ref_factor=2.0):
    """One block: LD R, GWAS estimates beta_hat (given its SEs) and its Cholesky factor."""
    import torch
    r, g, gen = make_block(n, seed * 1000003 + int(b), device, n_ref_factor=n_ref_factor)
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    var = torch.zeros(n, device=device, dtype=torch.float64)
    var = torch.where(u > 0.95, torch.full_like(var, 1e-6), var)
    var = torch.where(u > 0.99, torch.full_like(var, 1e-5), var)
    var = torch.where(u > 0.999, torch.full_like(var, 1e-4), var)
    # h2-normalise: E[sum_i beta_i^2 2f(1-f)] over M SNPs = h2  (E[2f(1-f)] = 0.3175 for U(.05,.5))
    scale = h2 / (M * 0.3175 * (0.04 * 1e-6 + 0.009 * 1e-5 + 0.001 * 1e-4))
    beta = torch.sqrt(var * scale) * torch.randn(n, generator=gen, device=device,
                                                 dtype=torch.float64)
    xi = torch.randn(g.shape[0], generator=gen, device=device, dtype=torch.float64)
    beta_hat = se * (r @ (beta / se)) + se * (g.T @ xi)
    return r, beta_hat


def precompute_block_full_rank(r, beta_hat, se, prior):
    """Set-up quantities of VIScheme.__init__ (:236-252) for a FULL-RANK dense block, where the
    pseudo-inverse is the inverse: adj = z/se, chi = z^T R^-1 z, ridge start via Cholesky."""
    import torch
    z = beta_hat / se
    L = torch.linalg.cholesky(r)
    sol = torch.cholesky_solve(z[:, None], L)[:, 0]
    chi = float(z @ sol)
    adj = z / se
    Lr = torch.linalg.cholesky(r + torch.diag(se * se / prior))
    inv_betas = torch.cholesky_solve(z[:, None], Lr)[:, 0] * se
    return chi, adj, inv_betas


def mixture_grid_single(betas, std_errs, K):
    """vi_options.py:208-229 + _make_simple for one cohort (no RNG consumed)."""
    b = np.abs(betas)
    s = std_errs
    psi = 1. / len(b)
    with np.errstate(over='ignore'):
        probs = 1. / (1. + ((1. - psi) / psi * np.sqrt(b**2 / s**2)
                            * np.exp(-0.5 * b**2 / s**2 + 0.5)))
    ebayes = np.maximum(b**2 - s**2, 1e-10)
    raw = b / (1. + s**2 / ebayes**2)
    mx = np.max(probs * raw)**2
    mn = np.nanpercentile(betas[betas**2 > 0]**2, 2.5)
    diag = [mn * 1e-6] + [mn * np.exp(np.log(mx / mn) / K * k) for k in range(K + 1)]
    return [np.array([[d]]) for d in diag]


# ------------------------------------------------------------------------------------------
# multi-cohort / low-rank workloads (BASELINE.json configs[2] and configs[4])
# ------------------------------------------------------------------------------------------
def lowrank_factors(g, ldthresh):
    """Eigen-factors (U [n,r], s [r]) of R = g^T g kept by LowRankMatrix(t=ldthresh)
    (matrix_structures.py:15-28: eigenvalues > 1 - sqrt(t), and > 1e-12 max for t = 1), from the
    small Gram matrix g g^T (n_ref x n_ref) instead of the n x n block."""
    import torch
    lam, w = torch.linalg.eigh(g @ g.T)
    cut = max(1.0 - float(np.sqrt(ldthresh)), 1e-12 * float(lam.max()))
    keep = lam > cut
    lam, w = lam[keep], w[:, keep]
    u = (g.T @ w) / torch.sqrt(lam)[None, :]
    return u.contiguous(), lam.contiguous()


def shared_effects(n, seed, b, P, M, device, h2=0.3, corr=0.8):
    """True effects of block b for P cohorts: the 4-component scale mixture of sim.py:97-156 with a
    shared causal indicator and cross-cohort correlation `corr`.  [P, n]."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 15485863 + int(b) * 31 + 5)
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    var = torch.zeros(n, device=device, dtype=torch.float64)
    var = torch.where(u > 0.95, torch.full_like(var, 1e-6), var)
    var = torch.where(u > 0.99, torch.full_like(var, 1e-5), var)
    var = torch.where(u > 0.999, torch.full_like(var, 1e-4), var)
    scale = h2 / (M * 0.3175 * (0.04 * 1e-6 + 0.009 * 1e-5 + 0.001 * 1e-4))
    shared = torch.randn(n, generator=gen, device=device, dtype=torch.float64)
    own = torch.randn((P, n), generator=gen, device=device, dtype=torch.float64)
    return torch.sqrt(var * scale)[None, :] * (np.sqrt(corr) * shared[None, :] + np.sqrt(1 - corr) * own)


def cohort_block(n, seed, b, p, se, beta, device, n_ref_factor, ldthresh):
    """One cohort's view of block b: LD factors (or the dense block), GWAS estimates and the set-up
    values of VIScheme.__init__ (:226-252) that need the block's (pseudo-)inverse.
    Returns dict(R or (U, s), beta_hat, chi, adj, ld_diag, rank) -- torch tensors on `device`."""
    import torch
    r, g, gen = make_block(n, (seed * 1000003 + int(b)) * 8 + p, device, n_ref_factor=n_ref_factor)
    xi = torch.randn(g.shape[0], generator=gen, device=device, dtype=torch.float64)
    full_rank = g.shape[0] > n and ldthresh >= 1.0
    if full_rank:
        beta_hat = se * (r @ (beta / se)) + se * (g.T @ xi)
        z = beta_hat / se
        L = torch.linalg.cholesky(r)
        chi = float(z @ torch.cholesky_solve(z[:, None], L)[:, 0])
        return dict(R=r, U=None, s=None, beta_hat=beta_hat, chi=chi, adj=z / se,
                    ld_diag=torch.ones_like(se), rank=n)
    u, s = lowrank_factors(g, ldthresh)
    rb = (u * s[None, :]) @ (u.T @ (beta / se))
    beta_hat = se * rb + se * (g.T @ xi)
    z = beta_hat / se
    a = u.T @ z
    chi = float((a * a / s).sum())
    adj = (u @ a) / se                        # R R^+ z / se
    ld_diag = (u * u) @ s
    return dict(R=None, U=u, s=s, beta_hat=beta_hat, chi=chi, adj=adj, ld_diag=ld_diag,
                rank=int(s.shape[0]))


def ridge_start(blk, se, prior):
    """inverse_betas of one block (reference :246-252 -> ridge_inverse_dot,
    matrix_structures.py:349-387): (R + diag(se^2/prior))^-1 (adj se) * se."""
    import torch
    lam = se * se / prior
    b = blk['adj'] * se
    if blk['R'] is not None:
        Lr = torch.linalg.cholesky(blk['R'] + torch.diag(lam))
        return torch.cholesky_solve(b[:, None], Lr)[:, 0] * se
    u, s = blk['U'], blk['s']
    li = 1.0 / lam
    c = torch.diag(1.0 / s) + (u.T * li[None, :]) @ u           # Woodbury core, r x r
    t = torch.cholesky_solve((u.T @ (li * b))[:, None], torch.linalg.cholesky(c))[:, 0]
    return (li * b - li * (u @ t)) * se
