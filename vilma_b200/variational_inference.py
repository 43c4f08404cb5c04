"""Variational-inference engine of `vilma fit`, B200-native.

Host mirror of ``vilma.variational_inference`` (/root/reference/src/vilma/
variational_inference.py): the same classes (``VIScheme`` :27-564, ``MultiPopVI``
:567-889), constructor arguments, method names, return layouts, exceptions and
control-flow constants -- but the parameters live in HBM for the whole of
``optimize()`` and every array operation of the loop is a CUDA kernel
(``csrc/snp_kernels.cuh``, ``csrc/ld_kernels.cuh``) reached through the C ABI.

What changes relative to the reference (none of it changes a result beyond rounding):

* each parameter state is evaluated ONCE (one fused per-SNP pass + one LD mat-vec per
  cohort); the objective, the mat-vec output and the sufficient statistics are cached with
  the state, which removes the reference's 3-4x re-evaluation of unchanged states
  (:766, :829, :442) and the K-fold broadcast of the gradient (:817-821);
* ``vi_sigma`` / ``nat_sigma`` / log-dets / traces (:712-733) are recomputed in registers per
  (k, SNP) instead of being stored as three [K,P,P,M] arrays;
* LD blocks are sharded over ranks; one small all-reduce per evaluated state.

The methods that take / return host ``params`` tuples (``elbo``, ``_update_beta``,
``_nat_grad_step``, ``_optimize_step``, ...) keep the reference semantics (inputs are never
mutated, new arrays are returned) by uploading, running the device path and downloading.
"""
import logging

import numpy as np

from . import numerics
from . import matrix_structures
from .dist import default_comm
from .partition import host_block_lists, local_blocks, partition_snps

L_MAX = 1e12        # sets minimum natural gradient stepsize is 1/L_MAX   (:18)
REL_TOL = 1e-6      # (:19)
ABS_TOL = 1e-6      # (:20)
ELBO_TOL = 0.1      # (:21)
EM_TOL = 10         # (:22)
ELBO_MOMENTUM = 0.5  # (:23)
MAX_NUM_ITERS = 20  # (:24)
INIT_HOST_BYTES = 2**31   # _initialize runs on the device when its host arrays would exceed this


class DeviceBlockDiagonalMatrix():
    """An LD operator that already lives in HBM (this rank's shard), for callers that build
    LD on the device.  Carries no host factors, so the constructor's host-side set-up
    (pseudo-inverse, ridge) cannot run on it: pass ``precomputed=`` to the VI class."""

    def __init__(self, device_ld, shape):
        self.device_ld = device_ld
        self.shape = tuple(shape)


class VIScheme():
    """Parent class for VI models of GWAS summary statistics (reference :27-564)."""

    def __init__(self,
                 marginal_effects=None,
                 std_errs=None,
                 ld_mats=None,
                 annotations=None,
                 mixture_covs=None,
                 checkpoint=True,
                 checkpoint_freq=5,
                 scaled=False,
                 scale_se=False,
                 output='vilma_output',
                 gwas_N=None,
                 init_hg=None,
                 num_its=None,
                 comm=None,
                 device=None,
                 precomputed=None,
                 local_snps=None,
                 engine_factory=None,
                 context=None):
        for val, name in [(init_hg, 'init_hg'), (gwas_N, 'gwas_N'),
                          (marginal_effects, 'marginal_effects'), (std_errs, 'std_errs'),
                          (ld_mats, 'ld_mats'), (annotations, 'annotations'),
                          (mixture_covs, 'mixture_covs'), (num_its, 'num_its')]:
            if val is None:
                raise ValueError('%s must be specified when calling VIScheme()' % name)
        marginal_effects = np.asarray(marginal_effects, dtype=np.float64)
        std_errs = np.asarray(std_errs, dtype=np.float64)
        if not np.all(np.isfinite(marginal_effects)):
            raise ValueError('Encountered an infinite or NaN value in the GWAS effect '
                             'size estimates')
        if not np.all(np.isfinite(std_errs)):
            raise ValueError('Encountered an infinity or NaN value in the GWAS standard '
                             'errors')
        self.scaled = scaled
        self.scale_se = scale_se
        self.error_scaling = np.ones(marginal_effects.shape[0])
        self.checkpoint = checkpoint
        self.checkpoint_path = '%s-checkpoint' % output
        self.num_pops, self.num_loci = marginal_effects.shape
        if len(ld_mats) != self.num_pops:
            raise ValueError('Fewer LD matrices than populations.')
        for ld in ld_mats:
            if not isinstance(ld, (matrix_structures.BlockDiagonalMatrix,
                                   DeviceBlockDiagonalMatrix)):
                raise ValueError('LD Matrices must be of type BlockDiagonalMatrix.')
        host_ld = all(isinstance(ld, matrix_structures.BlockDiagonalMatrix) for ld in ld_mats)
        if not host_ld and precomputed is None:
            raise ValueError('device-resident LD operators need precomputed= set-up values')
        for ld in ld_mats:
            if host_ld and ld.shape != (self.num_loci, self.num_loci):
                raise ValueError('LD matrix shape does not match GWAS marginal effect '
                                 'size shape.')
        annotations = np.asarray(annotations)
        if not np.allclose(annotations.sum(axis=1), 1):
            raise ValueError('Some SNPs are either missing annotations or have more than '
                             'one annotation.')
        self.num_annotations = annotations.shape[1]
        if annotations.shape[0] != self.num_loci:
            raise ValueError('annotations dimension does not match GWAS marginal effect '
                             'size shape.')

        self.use_native_loop = True      # C++ outer iteration (vb_fit_iteration) when possible
        self.init_on_device = None       # _initialize on the device: None = when the host form is too big
        self.speculate_next_trial = True  # queue the next iteration's first trial behind the refresh
        self._comm = comm if comm is not None else default_comm()
        self._device = device
        self._engine_factory = engine_factory
        self._ctx = context
        self._owns_ctx = False

        self.marginal_effects = np.copy(marginal_effects)
        if self.scaled:
            self.marginal_effects = self.marginal_effects / (std_errs + numerics.EPSILON)
            self.std_errs = np.ones_like(std_errs)
            self.scalings = (std_errs + numerics.EPSILON)
        else:
            self.std_errs = np.copy(std_errs)
            self.scalings = np.ones_like(std_errs)
        self.ld_mats = ld_mats
        self.annotations = np.copy(np.where(annotations)[1])
        self.annotation_counts = annotations.sum(axis=0)
        self.checkpoint_freq = checkpoint_freq
        self.init_hg = init_hg
        self.gwas_N = gwas_N
        self.num_its = num_its
        self.param_names = None

        if precomputed is not None:
            # set-up values supplied by the caller (device-built LD): same meaning as below
            self.ld_diags = np.array(precomputed['ld_diags'], dtype=np.float64)
            self.adj_marginal_effects = np.array(precomputed['adj_marginal_effects'],
                                                 dtype=np.float64)
            self.chi_stat = np.array(precomputed['chi_stat'], dtype=np.float64)
            self.ld_ranks = np.array(precomputed['ld_ranks'], dtype=np.float64)
            self.inverse_betas = np.array(precomputed['inverse_betas'], dtype=np.float64)
        else:
            self.ld_diags = np.concatenate(
                [ld.diag().reshape((1, -1)) for ld in ld_mats], axis=0)
        self.scaled_ld_diags = self.std_errs**-2 * self.ld_diags

        # which SNPs this rank owns: decided BEFORE the set-up below, so that with several ranks each
        # one factors, solves and uploads only its own LD blocks (1/N of the LAPACK work and of the HBM)
        if local_snps is not None:
            self._snps = np.asarray(local_snps, dtype=np.int64)
        elif self._comm.world == 1:
            self._snps = np.arange(self.num_loci, dtype=np.int64)
        else:
            parts = partition_snps(host_block_lists(ld_mats), self.num_loci, self._comm.world)
            self._snps = parts[self._comm.rank]
        sharded = host_ld and len(self._snps) != self.num_loci
        self._local_ld = [ld.restrict(self._snps) for ld in ld_mats] if sharded else list(ld_mats)

        if precomputed is None:
            # adjusted marginal effects S^-1 X X^+ S^-1 beta_hat, chi statistic, LD rank and the
            # LDpred-inf style ridge start (reference :226-252) over this rank's blocks; the ranks'
            # pieces are disjoint, so one sum assembles the global vectors.  Pseudo-inverse and ridge
            # solve are one-off host LAPACK; the R @ mle product is the GPU operator.
            snps = self._snps
            self.adj_marginal_effects = np.zeros_like(self.marginal_effects)
            self.chi_stat = np.zeros(self.num_pops)
            self.ld_ranks = np.zeros(self.num_pops)
            self.inverse_betas = np.zeros_like(self.marginal_effects)
            self.setup_report = []
            for p in range(self.num_pops):
                ld = self._local_ld[p]
                se = self.std_errs[p, snps]
                z_scores = self.marginal_effects[p, snps] / se
                prior = (2 * gwas_N[p] * init_hg[p] / (self.std_errs[p, :]**-2).sum())
                if any(m._X is not None and not m.factorized for m in ld.matrices) \
                        and self._engine_factory is None:
                    # lazily loaded dense blocks: pseudo-inverse product, R mle and the ridge start on
                    # the GPU (two Cholesky factorisations per block), exact host path for the rest
                    res = ld.device_setup(z_scores, se**2 / prior, ctx=self._context())
                    self.chi_stat[p] = res['chi']
                    this_adj_marg = res['rmle'] / se
                    self.ld_ranks[p] = res['rank']
                    inv_z_scores = res['ridge']
                    self.setup_report.append((res['gpu_blocks'], res['host_blocks']))
                else:
                    mle = ld.inverse.dot(z_scores)
                    self.chi_stat[p] = z_scores.dot(mle)
                    this_adj_marg = ld.dot(np.copy(mle), ctx=self._context())
                    this_adj_marg = this_adj_marg / se
                    self.ld_ranks[p] = ld.get_rank()
                    inv_z_scores = ld.ridge_inverse_dot(this_adj_marg * se, se**2 / prior)
                self.adj_marginal_effects[p, snps] = this_adj_marg
                self.inverse_betas[p, snps] = inv_z_scores * se
            if sharded:
                self.adj_marginal_effects = self._comm.sum(self.adj_marginal_effects)
                self.inverse_betas = self._comm.sum(self.inverse_betas)
                self.chi_stat = self._comm.sum(self.chi_stat)
                self.ld_ranks = self._comm.sum(self.ld_ranks)

        if not np.allclose(self.adj_marginal_effects[np.isclose(self.ld_diags, 0)], 0):
            raise ValueError('Some SNPs that are missing in the LD matrix are not being '
                             'treated as missing.')

    # ------------------------------------------------------------------ device plumbing
    def _context(self):
        if self._engine_factory is not None:
            return None
        if self._ctx is None:
            # a vb_ctx holds ONE fit state, so every model gets its own context and frees it
            # (close() / __del__) unless the caller passed one in with context=
            from .engine import DeviceContext
            self._ctx = DeviceContext(self._device if self._device is not None
                                      else _current_device())
            self._owns_ctx = True
        return self._ctx

    def close(self):
        """Release the device memory of this model: the fit state and, when the model created its
        own context, that context with the LD operators uploaded into it.  Idempotent; also run
        when the object is garbage collected."""
        eng = self.__dict__.get('_eng')
        if eng is not None and hasattr(eng, 'close'):
            eng.close()
        if self.__dict__.get('_owns_ctx') and self._ctx is not None:
            for ld in (self.__dict__.get('ld_mats') or []) + (self.__dict__.get('_local_ld') or []):
                if hasattr(ld, 'release_device'):
                    ld.release_device()
            self._ctx.close()
        self._eng = None
        self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dump_info(self, num_its, diff):
        """Log information about convergence (reference :292-331), from device reductions."""
        n = self.num_pops * self.num_loci
        logging.info('Completed iteration %d', num_its + 1)
        logging.info('Maximum posterior mean beta: %e', diff[5])
        logging.info('SE scaling is: %r', self.error_scaling)
        logging.info('Max relative difference is: %e', diff[6])
        logging.info('Max absolute difference is: %e', diff[7])
        logging.info('Mean absolute difference is: %e', diff[1] / n)
        logging.info('RMSE difference is: %e', np.sqrt(diff[2] / n))
        logging.info('Max relative difference (checkpoint iterations) is: %e', diff[8])
        logging.info('Max absolute difference (checkpoint iterations) is: %e', diff[9])
        logging.info('Mean absolute difference (checkpoint iterations) is: %e', diff[3] / n)
        logging.info('RMSE difference (checkpoint iterations) is: %e', np.sqrt(diff[4] / n))

    def create_dump_dict(self, params):
        dump_dict = dict(zip(self.param_names, params))
        dump_dict['error_scaling'] = self.error_scaling
        dump_dict['scalings'] = self.scalings
        return dump_dict

    # ------------------------------------------------------------------ the loop
    def optimize(self, loaded_checkpoint=None):
        """Initialize params and optimize objective function (reference :340-394)."""
        if loaded_checkpoint is None:
            params = self._initialize()
        else:
            params = [np.asarray(loaded_checkpoint[p_name]) for p_name in self.param_names]
            try:
                self.error_scaling = np.array(loaded_checkpoint['error_scaling'],
                                              dtype=np.float64)
            except KeyError:
                logging.warning('Did not find "error_scaling" in the loaded '
                                'checkpoint. That is okay, but we will have '
                                'to assume that the error scalings are 1.')
            self._set_state(params)
        state = self.begin_loop(params)
        try:
            state = self.run_loop(state, self.num_its, fresh=loaded_checkpoint is None)
        finally:
            self._finish_checkpoints()
        num_its = state['num_its']
        if num_its == self.num_its:
            logging.warning('Failed to converge')
        logging.info('Optimization ran for %d iterations', num_its)
        self.num_its_run = num_its
        return self._download()

    # -- periodic checkpoints (reference :362-367) are snapshotted synchronously (device -> fresh host
    # arrays) and written by a background thread while the next iterations run on the device; at most
    # one write is in flight, so the host holds at most two snapshots.
    def _write_checkpoint(self, fname, dump_dict):
        if self._comm.rank != 0:
            return
        import threading
        self._finish_checkpoints()
        dump_dict = {k: np.array(v) if k in ('error_scaling',) else v for k, v in dump_dict.items()}
        err = []

        def work():
            try:
                np.savez(fname, **dump_dict)
            except BaseException as exc:
                err.append(exc)
        th = threading.Thread(target=work, name='checkpoint-writer', daemon=True)
        th.start()
        self._ckpt_writer = (th, err)

    def _finish_checkpoints(self):
        pending = getattr(self, '_ckpt_writer', None)
        if pending is None:
            return
        self._ckpt_writer = None
        pending[0].join()
        if pending[1]:
            raise pending[1][0]

    def begin_loop(self, params):
        """Make `params` resident and set up the loop state of optimize() (reference :353-360)."""
        elbo = self.elbo(params)        # uploads; the state is resident from here on
        self._eng.pm_mark(0)            # post_mean
        self._eng.pm_mark(1)            # ckp_post_mean
        self.trajectory = {'elbo': [], 'L0': [], 'trials': [], 'running': []}
        return {'elbo': elbo, 'running': None, 'num_its': 0, 'L': np.ones(5),
                'converged': False}

    def run_loop(self, state, max_its, fresh=True):
        """The while-loop of optimize() (reference :361-389) on the device-resident state, until
        convergence or `max_its` total iterations.  Returns the updated loop state."""
        eng = self._eng
        L, elbo, running_elbo_delta = state['L'], state['elbo'], state['running']
        num_its, converged = state['num_its'], state['converged']
        want_info = logging.getLogger().isEnabledFor(logging.INFO)
        native = (self.use_native_loop and not want_info
                  and hasattr(eng, 'iteration') and self._engine_factory is None)
        if native:
            try:
                return self._run_loop_native(state, max_its, fresh)
            finally:
                self._finish_checkpoints()
        while num_its < max_its and not converged:
            if num_its % self.checkpoint_freq == 0 and self.checkpoint:
                eng.pm_mark(1)
                fname = '{}.{}'.format(self.checkpoint_path, num_its)
                dump_dict = self.create_dump_dict(self._download())
                self._write_checkpoint(fname, dump_dict)
            trials0 = self.n_trials
            L, elbo, running_elbo_delta = self._optimize_step_dev(
                L, elbo, 2., running_elbo_delta)

            diff_dev = eng.pm_diff(ABS_TOL, REL_TOL)
            diff = self._comm.sum(diff_dev)
            converged = diff[0] == 0
            converged = converged or bool(np.isclose(running_elbo_delta, 0,
                                                     atol=ELBO_TOL, rtol=0))
            if num_its < 10 and fresh:
                converged = False
            if want_info:
                if self._comm.world > 1:
                    diff[5:] = self._comm.max(diff_dev)[5:]
                self._dump_info(num_its, diff)
            self.trajectory['elbo'].append(float(elbo))
            self.trajectory['L0'].append(float(L[0]))
            self.trajectory['trials'].append(self.n_trials - trials0)
            self.trajectory['running'].append(float(running_elbo_delta))
            num_its += 1
        self._finish_checkpoints()
        return {'elbo': elbo, 'running': running_elbo_delta, 'num_its': num_its, 'L': L,
                'converged': converged}

    def _run_loop_native(self, state, max_its, fresh):
        """run_loop with each outer iteration executed by the library's C++ loop
        (vb_fit_iteration): same thresholds and decisions, one ctypes call per iteration."""
        from ._lib import StepIO
        eng = self._eng
        if not eng.native_ready:
            eng.set_constants(self.chi_stat, self.ld_ranks, self.annotation_counts,
                              self.log_det, self.scale_se)
        L, elbo, running = state['L'], state['elbo'], state['running']
        num_its, converged = state['num_its'], state['converged']
        tau = np.ascontiguousarray(self.error_scaling, dtype=np.float64).copy()
        hyper = np.ascontiguousarray(self._hyper, dtype=np.float64).copy()
        stats = np.ascontiguousarray(self._res_stats, dtype=np.float64).copy()
        io = StepIO()
        io.line_search_rate = 2.
        io.atol, io.rtol = ABS_TOL, REL_TOL
        io.do_diff = 1
        io.speculate = 1 if self.speculate_next_trial else 0
        io.obj = self._res_obj
        for i in range(5):
            io.L[i] = L[i]
        # The interpreter's cyclic collector is paused for the duration of the loop: a collection is a
        # 0.3 ms (young generation) to 40 ms (full) stall of this rank's launch thread, and with N ranks in
        # lock step every rank waits for whichever one is collecting -- measurable at 0.5 ms per iteration.
        # Nothing in the loop creates reference cycles; everything it allocates is freed by refcount.
        import gc
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            return self._native_iterations(eng, io, tau, hyper, stats, elbo, running, num_its, converged,
                                           max_its, fresh)
        finally:
            if gc_was_on:
                gc.enable()

    def _native_iterations(self, eng, io, tau, hyper, stats, elbo, running, num_its, converged, max_its,
                           fresh):
        while num_its < max_its and not converged:
            if num_its % self.checkpoint_freq == 0 and self.checkpoint:
                eng.pm_mark(1)
                self._sync_from_native(tau, hyper, stats, io)
                fname = '{}.{}'.format(self.checkpoint_path, num_its)
                dump_dict = self.create_dump_dict(self._download())
                self._write_checkpoint(fname, dump_dict)
            io.speculate = 1 if (self.speculate_next_trial and num_its + 1 < max_its) else 0
            io.has_running = 0 if running is None else 1
            io.running_elbo_delta = 0. if running is None else running
            eng.iteration(io, tau, hyper, stats)
            self._resident = None
            elbo_change = io.elbo_delta
            elbo = elbo + elbo_change                       # reference :405
            if running is None:
                running = elbo_change
            running *= ELBO_MOMENTUM
            running += (1 - ELBO_MOMENTUM) * np.maximum(elbo_change, 0)
            self.n_trials += io.trials
            self.n_evals += io.evals
            self.n_rejects += io.rejects
            converged = io.diff[0] == 0
            converged = converged or abs(running) <= ELBO_TOL     # np.isclose(running, 0, atol=ELBO_TOL, rtol=0)
            if num_its < 10 and fresh:
                converged = False
            self.trajectory['elbo'].append(float(elbo))
            self.trajectory['L0'].append(float(io.L[0]))
            self.trajectory['trials'].append(int(io.trials))
            self.trajectory['running'].append(float(running))
            num_its += 1
        self._sync_from_native(tau, hyper, stats, io)
        return {'elbo': elbo, 'running': running, 'num_its': num_its,
                'L': np.array([io.L[i] for i in range(5)]), 'converged': converged}

    def _sync_from_native(self, tau, hyper, stats, io):
        """Mirror the state the C++ loop advanced back into the Python object."""
        self.error_scaling = tau.copy()
        self._hyper = hyper.copy()
        self._gtable = numerics.vi_delta_grad_table(self._hyper, self.log_det)
        self._res_stats = stats.copy()
        self._res_obj = float(io.obj)
        self._res_valid = True
        self._resident = None

    def _optimize_step_dev(self, L, curr_elbo, line_search_rate, running_elbo_delta):
        logging.info('Current ELBO = %f and L = %f,%f,%f,%f,%f',
                     curr_elbo, L[0], L[1], L[2], L[3], L[4])
        L_new, elbo_change = self._nat_grad_step_dev(L, line_search_rate, running_elbo_delta)
        elbo = curr_elbo + elbo_change
        if running_elbo_delta is None:
            running_elbo_delta = elbo_change
        running_elbo_delta *= ELBO_MOMENTUM
        running_elbo_delta += (1 - ELBO_MOMENTUM) * np.maximum(elbo_change, 0)
        return L_new, elbo, running_elbo_delta

    def _optimize_step(self, params, L, curr_elbo, line_search_rate=1.25,
                       running_elbo_delta=None):
        """Update each set of params and tally up improvement in ELBo (reference :396-410)."""
        self._make_resident(params)
        L_new, elbo, running = self._optimize_step_dev(L, curr_elbo, line_search_rate,
                                                       running_elbo_delta)
        return self._download(), L_new, elbo, running

    def elbo(self, params):
        """ELBo of the state `params` (reference :412-417; annotation KL is zero)."""
        self._make_resident(params)
        return self._res_obj

    def _nat_grad_step(self, params, L, line_search_rate, running_elbo_delta=None):
        """One iteration of updating each set of (hyper)parameters (reference :419-450)."""
        self._make_resident(params)
        L, delta = self._nat_grad_step_dev(L, line_search_rate, running_elbo_delta)
        return self._download(), L, delta

    def _nat_grad_step_dev(self, L, line_search_rate, running_elbo_delta=None):
        updates = [self._update_beta_dev, self._update_hyper_delta_dev,
                   self._update_annotation_dev]
        conv_tol = (float('inf') if running_elbo_delta is None
                    else 0.1 * running_elbo_delta)
        new_elbo_delta = 0
        for idx, update in enumerate(updates):
            orig_obj = None
            for update_iter in range(MAX_NUM_ITERS):
                L[idx] = max([1., L[idx] / 1.25])
                logging.info('...Updating paramset %d, L=%f', idx, L[idx])
                L, orig_obj, new_obj = update(orig_obj, L, idx, line_search_rate)
                new_elbo_delta += new_obj - orig_obj
                with np.errstate(invalid='ignore'):
                    small = np.isclose(new_obj - orig_obj, 0, atol=conv_tol, rtol=0)
                if small or L[idx] == 1 or L[idx] > L_MAX:
                    break
                orig_obj = new_obj

        if self.scale_se and new_elbo_delta < EM_TOL:
            orig_obj = self._res_obj
            self._update_error_scaling_dev()
            new_obj = self._refresh_delta_dev()
            new_elbo_delta += new_obj - orig_obj
            logging.info('...Updating error_scaling, old ELBo=%f, new ELBo=%f',
                         orig_obj, new_obj)
        return L, new_elbo_delta

    # -- objective pieces from the reduced statistics of the resident state
    def _objective(self, stats):
        P = self.num_pops
        A_, C_, B_ = stats[0:P], stats[P:2 * P], stats[2 * P:3 * P]
        per_pop = (-0.5 * (C_ + B_) + A_) - 0.5 * self.chi_stat       # numerics.py:39-44
        loglik = (per_pop / self.error_scaling
                  - 0.5 * self.ld_ranks * np.log(self.error_scaling)).sum()
        kl = stats[3 * P] + stats[3 * P + 1] + stats[3 * P + 2]
        return float(loglik), float(kl)

    def _log_likelihood(self, params):
        self._make_resident(params)
        return self._objective(self._res_stats)[0]

    def _beta_objective(self, params):
        return self.elbo(params)

    def _update_error_scaling_dev(self):
        """tau_p = [chi_p - 2 pm.adj + z^T R z + sum sld pv] / rank_p   (reference :472-486)."""
        P = self.num_pops
        s = self._res_stats
        self.error_scaling = (self.chi_stat - 2 * s[0:P] + s[2 * P:3 * P] + s[P:2 * P]) \
            / self.ld_ranks
        self._eng.set_tau(self.error_scaling)
        self._res_valid = False

    def _update_error_scaling(self, params):
        self._make_resident(params)
        self._update_error_scaling_dev()


def _current_device():
    import torch
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


class MultiPopVI(VIScheme):
    """Standard VI scheme for GWAS across one or more populations (reference :567-889)."""

    def __init__(self, mixture_covs=None, **kwargs):
        num_pops = np.asarray(kwargs['marginal_effects']).shape[0]
        for mc in mixture_covs:
            if np.asarray(mc).shape != (num_pops, num_pops):
                raise ValueError('Mixture component has a covariance matrix of the wrong '
                                 'shape.')
        signs, _ = np.linalg.slogdet(np.array(mixture_covs))
        if not np.all(signs == 1):
            raise ValueError('Mixture component has a non-positive definite covariance '
                             'matrix.')
        if num_pops >= 3:
            # The reference only tests the determinant's sign (:610-613), which also lets through
            # matrices with an even number of negative eigenvalues; it then inverts them with LAPACK
            # and produces negative "variances".  The P >= 3 kernels factor Lambda = Prec_k + diag(.)
            # as L D L^T and take logs of the pivots, so such a grid is rejected here, loudly, instead
            # of being fitted to garbage (documented deviation).
            evals = np.linalg.eigvalsh(np.array(mixture_covs, dtype=np.float64))
            if not np.all(evals > 0):
                raise ValueError('Mixture component has a non-positive definite covariance '
                                 'matrix.')
        self.num_mix = len(mixture_covs)
        VIScheme.__init__(self, mixture_covs=mixture_covs, **kwargs)
        self.param_names = ['vi_mu', 'vi_delta', 'hyper_delta']

        covs = np.array(mixture_covs, dtype=np.float64)
        self.mixture_prec = numerics.small_inverse(covs)[:, :, :, None]     # [K,P,P,1]
        self.log_det = np.copy(numerics.small_log_det(covs))

        self._gtable = None             # nat_grad_vi_delta as an [A,K-1] table
        self._resident = None           # host arrays the device state corresponds to
        self._resident_fp = None
        self._res_valid = False
        self._res_obj = None
        self._res_stats = None
        self._hyper = None
        self.n_trials = 0
        self.n_evals = 0
        self.n_rejects = 0              # line-search step sizes rejected (native loop)
        self._build_engine()

    # ------------------------------------------------------------------ engine
    def _build_engine(self):
        snps = self._snps
        M = self.num_loci
        if len(snps) == 0:
            raise ValueError('rank %d owns no SNPs; use fewer ranks' % self._comm.rank)
        pieces = dict(K=self.num_mix, P=self.num_pops, M=len(snps), A=self.num_annotations,
                      adj=self.adj_marginal_effects[:, snps], se=self.std_errs[:, snps],
                      sld=self.scaled_ld_diags[:, snps], scalings=self.scalings[:, snps],
                      annotations=self.annotations[snps],
                      mixture_prec=self.mixture_prec[..., 0], log_det=self.log_det)
        if self._engine_factory is not None:
            self._eng = self._engine_factory(self, snps, pieces)
        else:
            from .engine import CudaEngine
            ctx = self._context()
            lds = []
            for ld in self._local_ld:
                if isinstance(ld, DeviceBlockDiagonalMatrix):
                    lds.append(ld.device_ld)
                else:
                    # (the rank's own blocks only; already resident if the set-up above used them)
                    lds.append(ld.to_device(ctx))
            self._eng = CudaEngine(ctx, lds, **pieces)
            self._eng.init_comm(self._comm)
            if len(snps) != M:
                self._eng.set_shard(snps, M)
        self._eng.set_tau(self.error_scaling)

    # ------------------------------------------------------------------ hidden state
    @property
    def nat_grad_vi_delta(self):
        if self._gtable is None:
            return None
        return self._gtable[self.annotations]

    @nat_grad_vi_delta.setter
    def nat_grad_vi_delta(self, value):
        if value is None:
            self._gtable = None
            return
        value = np.asarray(value, dtype=np.float64)
        table = np.zeros((self.num_annotations, self.num_mix - 1))
        for a in range(self.num_annotations):
            rows = np.where(self.annotations == a)[0]
            if len(rows):
                table[a] = value[rows[0]]
        if not np.array_equal(table[self.annotations], value):
            raise ValueError('nat_grad_vi_delta must be constant within an annotation')
        self._set_gtable(table)

    def _set_gtable(self, table):
        self._gtable = np.array(table, dtype=np.float64)
        self._eng.set_delta_grad(self._gtable)
        self._res_valid = False

    def _covariances(self):
        """(vi_sigma [K,P,P,M], Lambda) on the host from the device kernel -- outputs/tests only."""
        return self._comm.gather_snp_axis(self._eng.vi_sigma(), self._snps, self.num_loci, 3)

    def _covariance_slice(self, k0, k1, out=None):
        """vi_sigma[k0:k1] (all SNPs) on the host; `out` = this rank's landing buffer.  The streamed
        `.npz` writer (outputs.save_fit_npz) walks the components with this so that the [K,P,P,M] array
        never exists whole."""
        local = self._eng.vi_sigma(k0, k1, out=out)
        return self._comm.gather_snp_axis(local, self._snps, self.num_loci, 3)

    @property
    def vi_sigma(self):
        """S_ki = (Prec_k + diag(sld_i/tau))^-1, [K,P,P,M]   (reference :712-724)."""
        self._eng.set_tau(self.error_scaling)
        return self._covariances()

    @property
    def nat_sigma(self):
        K, P, M = self.num_mix, self.num_pops, self.num_loci
        lam = np.zeros((K, P, P, M))
        idx = np.arange(P)
        lam[:, idx, idx, :] = self.scaled_ld_diags / self.error_scaling.reshape((-1, 1))
        lam += self.mixture_prec
        return -0.5 * lam

    @property
    def vi_sigma_log_det(self):
        S = np.transpose(self.vi_sigma, (0, 3, 1, 2))        # [K,M,P,P]
        return np.linalg.slogdet(S)[1]

    @property
    def vi_sigma_matches(self):
        return np.einsum('kpq,kqpi->ik', self.mixture_prec[..., 0], self.vi_sigma)

    @property
    def sigma_summary(self):
        return self.log_det - self.vi_sigma_log_det.T + self.vi_sigma_matches

    def _set_vi_sigma(self):
        """vi_sigma is a function of tau only and is recomputed on the device on the fly."""
        self._eng.set_tau(self.error_scaling)
        self._res_valid = False

    # ------------------------------------------------------------------ host <-> device
    @staticmethod
    def _fingerprint(params):
        """Cheap content check of a host state (a strided sample of every array): catches in-place
        edits of arrays the identity-keyed cache would otherwise take for the resident state."""
        out = []
        for a in params:
            flat = np.asarray(a).reshape(-1)
            out.append(float(flat[::max(1, flat.size // 509)].sum()) if flat.size else 0.0)
        return tuple(out)

    def _same(self, params):
        if self._resident is None or not self._res_valid:
            return False
        if not all(a is b for a, b in zip(params, self._resident)):
            return False
        fp = self._fingerprint(params)
        return all(x == y or (x != x and y != y) for x, y in zip(fp, self._resident_fp))

    def _make_resident(self, params):
        """Upload a host state (unless it is the resident one) and evaluate it."""
        if self._same(params):
            return
        vi_mu, vi_delta, hyper_delta = (np.asarray(x, dtype=np.float64) for x in params)
        self._upload(vi_mu, vi_delta, hyper_delta)
        stats = self._comm.sum(self._eng.eval())
        self.n_evals += 1
        self._set_result(stats, tuple(params))

    def _upload(self, vi_mu, vi_delta, hyper_delta):
        snps = self._snps
        self._hyper = np.array(hyper_delta, dtype=np.float64)
        self._eng.set_hyper(self._hyper)
        self._eng.set_tau(self.error_scaling)
        if self._gtable is not None:
            self._eng.set_delta_grad(self._gtable)
        if len(snps) == self.num_loci:
            self._eng.set_params(vi_mu, vi_delta)
        else:
            # this rank's shard only crosses PCIe: whole LD blocks are contiguous SNP ranges, so the
            # shard is cut out on the host with a few slice copies
            eng = self._eng
            if hasattr(eng, 'set_params_shard'):
                eng.set_params_shard(vi_mu, vi_delta)      # (native: run copies into page-locked staging)
            else:
                from .dist import take_runs
                eng.set_params(take_runs(vi_mu, snps, 2), take_runs(vi_delta, snps, 0))

    def _set_result(self, stats, resident):
        """Cache the reduced statistics / objective of the (new) accepted device state.
        `resident` is the host tuple it corresponds to, or None if it only lives in HBM."""
        self._res_stats = stats
        ll, kl = self._objective(stats)
        self._res_obj = ll - kl
        self._resident = resident
        self._resident_fp = self._fingerprint(resident) if resident is not None else None
        self._res_valid = True

    def _download(self):
        """Resident state -> host tuple (new arrays, reference layouts)."""
        if self._resident is not None:
            return self._resident
        eng, comm = self._eng, self._comm
        got = None
        if comm.world > 1 and hasattr(comm, 'shared_arrays'):
            K, P, M = self.num_mix, self.num_pops, self.num_loci
            got = comm.shared_arrays([(K, P, M), (M, K)])
        if got is not None:
            # one node-shared host mapping; this rank fills in its own SNPs -- 1/N of the bytes over PCIe.
            # Where the driver can page-lock the mapping the GPU scatters them there itself.
            (mu, delta), entry = got
            if hasattr(eng, 'get_params_shard'):
                if 'registered' not in entry:
                    entry['registered'] = eng.host_register(entry['address'], entry['nbytes'])
                    if entry['registered']:
                        from . import _lib
                        addr = entry['address']
                        entry['release'] = lambda: _lib.load().vb_host_unregister(_lib.C.c_void_p(addr))
                direct = entry['registered']
            else:
                direct = False
            if direct:
                eng.get_params_shard(mu, delta)
            else:
                comm.fill_shared([mu, delta], eng.get_params(), self._snps, [2, 0])
            comm.barrier()                   # every shard is in place
        else:
            mu, delta = eng.get_params()
            if comm.world > 1:
                mu = comm.gather_snp_axis(mu, self._snps, self.num_loci, 2)
                delta = comm.gather_snp_axis(delta, self._snps, self.num_loci, 0)
        self._resident = (mu, delta, np.array(self._hyper))
        for arr in self._resident:
            # the cache is keyed on object identity (_same): an in-place edit by the caller would go
            # unnoticed, so the arrays are handed out read-only (copy them to modify)
            arr.setflags(write=False)
        self._resident_fp = self._fingerprint(self._resident)
        return self._resident

    # ------------------------------------------------------------------ moments
    def real_posterior_mean(self, vi_mu, vi_delta, hyper_delta):
        return self._posterior_mean(vi_mu, vi_delta, hyper_delta) * self.scalings

    def real_posterior_variance(self, vi_mu, vi_delta, hyper_delta):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        pv = self._gather_pp(self._eng.posterior()[1])
        return pv * (self.scalings**2)

    def _posterior_mean(self, vi_mu, vi_delta, hyper_delta):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        return self._gather_pp(self._eng.posterior()[0])

    def _posterior_marginal_variance(self, mean, vi_mu, vi_delta, hyper_delta):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        return self._gather_pp(self._eng.posterior()[1])

    def _gather_pp(self, local):
        return self._comm.gather_snp_axis(local, self._snps, self.num_loci, 1)

    # ------------------------------------------------------------------ initialisation
    def _initialize(self):
        """Starting values of the variational parameters (reference :643-700).

        One-off and seed dependent: consumes np.random.normal exactly as the reference does,
        on the host; the covariances it needs come from the device kernel."""
        real_mu = self.inverse_betas
        logging.info('Largest inverse_beta is %f', np.max(np.abs(real_mu)))
        missing = np.isclose(self.ld_diags, 0)
        fake_mu = np.copy(real_mu)
        fake_mu = np.random.normal(loc=fake_mu, scale=1e-3 * self.std_errs,
                                   size=fake_mu.shape)
        fake_mu[missing] = np.nan
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', category=RuntimeWarning)
            mu_fill = np.tile(np.nanmean(fake_mu, axis=0), [fake_mu.shape[0], 1])
        fake_mu[missing] = mu_fill[missing]
        fake_mu[np.isnan(fake_mu)] = 0.
        on_device = self.init_on_device
        if on_device is None:
            # the host form below holds three [K,P,P,M] arrays at once
            on_device = 24. * self.num_mix * self.num_pops**2 * self.num_loci > INIT_HOST_BYTES
        if on_device and hasattr(self._eng, 'init_delta'):
            return self._initialize_device(fake_mu)
        vi_sigma = self.vi_sigma
        matches = np.einsum('kpq,kqpi->ik', self.mixture_prec[..., 0], vi_sigma)
        probs = np.einsum('pi,oi,kpo->ik', 1.6 * fake_mu, 1.6 * fake_mu,
                          self.mixture_prec[..., 0])
        probs += matches
        probs -= self.log_det
        probs = np.exp(-0.5 * (probs - np.min(probs, axis=1, keepdims=True)))
        vi_delta = np.maximum(probs / probs.sum(axis=1, keepdims=True), numerics.EPSILON)
        real_hyper_delta = numerics.sum_annotations_host(vi_delta, self.annotations,
                                                         self.num_annotations)
        real_hyper_delta += 1.
        real_hyper_delta /= np.sum(real_hyper_delta, axis=1, keepdims=True)
        real_hyper_delta = np.maximum(real_hyper_delta, numerics.EPSILON)
        self._set_gtable(numerics.vi_delta_grad_table(real_hyper_delta, self.log_det))
        avg_mats = np.einsum('kpqi,ik->ipq', vi_sigma, vi_delta)
        inv_avg_mats = np.linalg.inv(avg_mats)
        temp_nat_mu = np.einsum('pi,iqp->qi', fake_mu, inv_avg_mats)
        vi_mu = np.einsum('kqpi,pi->kqi', vi_sigma, temp_nat_mu)
        _, vi_delta, _ = self._nat_to_not_vi_delta((vi_mu, vi_delta, real_hyper_delta))
        return vi_mu, vi_delta, real_hyper_delta

    def _initialize_device(self, fake_mu):
        """The rest of _initialize (reference :660-679) on the device: S_ki is recomputed per
        (k, SNP) in registers by vb_init_delta_kernel / vb_init_mu_kernel instead of being held
        as [K,P,P,M] host arrays.  The state stays resident; the host tuple is its download."""
        eng = self._eng
        eng.set_tau(self.error_scaling)
        eng.init_delta(fake_mu[:, self._snps])
        sums = np.array(self._comm.sum(eng.sum_annotations()), dtype=np.float64)
        real_hyper_delta = sums.reshape(self.num_annotations, self.num_mix) + 1.
        real_hyper_delta /= np.sum(real_hyper_delta, axis=1, keepdims=True)
        real_hyper_delta = np.maximum(real_hyper_delta, numerics.EPSILON)
        self._set_gtable(numerics.vi_delta_grad_table(real_hyper_delta, self.log_det))
        self._hyper = np.array(real_hyper_delta)
        eng.set_hyper(self._hyper)
        eng.init_mu()
        self._refresh_delta_dev()
        self._resident = None
        return self._download()

    def _set_state(self, params):
        """Set internal values given parameter values (reference :702-710)."""
        vi_mu, vi_delta, hyper_delta = params
        self._set_vi_sigma()
        self._set_gtable(numerics.vi_delta_grad_table(
            np.asarray(hyper_delta, dtype=np.float64), self.log_det))

    # ------------------------------------------------------------------ updates (device)
    def _refresh_delta_dev(self):
        """delta := softmax from the resident mu under the current gtable / tau; accept."""
        stats = self._comm.sum(self._eng.refresh_delta())
        self.n_evals += 1
        self._eng.accept()
        self._set_result(stats, None)
        return self._res_obj

    def _nat_to_not_vi_delta(self, params):
        """Convert natural parameterization of delta to delta (reference :632-641)."""
        vi_mu, vi_delta, hyper_delta = params
        vi_mu = np.asarray(vi_mu, dtype=np.float64)
        self._upload(vi_mu, np.asarray(vi_delta, dtype=np.float64),
                     np.asarray(hyper_delta, dtype=np.float64))
        self._refresh_delta_dev()
        _, delta = self._eng.get_params()
        delta = self._comm.gather_snp_axis(delta, self._snps, self.num_loci, 0)
        out = (vi_mu, delta, hyper_delta)
        self._resident = out
        self._resident_fp = self._fingerprint(out)
        return out

    def _update_beta_dev(self, orig_obj, L, idx, lsr):
        """Natural-gradient step on (vi_mu, vi_delta) with backtracking (reference :762-802)."""
        if orig_obj is None:
            orig_obj = self._res_obj
        if self._gtable is None:
            raise RuntimeError('nat_grad_vi_delta must always be set prior to running '
                               '_update_beta')
        while True:
            step_size = 1. / L[idx]
            stats = self._comm.sum(self._eng.beta_trial(step_size))
            self.n_trials += 1
            self.n_evals += 1
            ll, kl = self._objective(stats)
            new_obj = ll - kl
            logging.info('...Old objective = %f, new objective = %f', orig_obj, new_obj)
            if new_obj >= orig_obj - REL_TOL * np.abs(orig_obj) - ABS_TOL:
                if L[idx] > L_MAX:
                    if not np.isclose(orig_obj, new_obj):
                        raise RuntimeError('Encountered a numerical error.')
                break
            if L[idx] > L_MAX:
                if not np.isclose(orig_obj, new_obj):
                    raise RuntimeError('Encountered a numerical error.')
                return L, orig_obj, orig_obj
            L[idx] *= lsr
        self._eng.accept()
        self._set_result(stats, None)
        return L, orig_obj, new_obj

    def _update_beta(self, vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        L, orig_obj, new_obj = self._update_beta_dev(orig_obj, L, idx, lsr)
        return self._download(), L, orig_obj, new_obj

    def _update_hyper_delta_dev(self, orig_obj, L, idx, lsr):
        """Mixture-weight update (reference :825-860): per-annotation mean of delta."""
        if orig_obj is None:
            orig_obj = self._res_obj
        sums = self._comm.sum(self._eng.sum_annotations()).reshape(
            self.num_annotations, self.num_mix)
        new_hyper_delta = np.maximum(
            sums / (self.annotation_counts.reshape((-1, 1)) + numerics.EPSILON),
            numerics.EPSILON)
        new_hyper_delta /= new_hyper_delta.sum(axis=1, keepdims=True)
        self._hyper = new_hyper_delta
        self._eng.set_hyper(new_hyper_delta)
        self._set_gtable(numerics.vi_delta_grad_table(new_hyper_delta, self.log_det))
        new_obj = self._refresh_delta_dev()
        logging.info('...Old objective = %f, new objective = %f', orig_obj, new_obj)
        return L, orig_obj, new_obj

    def _update_hyper_delta(self, vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        L, orig_obj, new_obj = self._update_hyper_delta_dev(orig_obj, L, idx, lsr)
        return self._download(), L, orig_obj, new_obj

    def _update_annotation_dev(self, orig_obj, L, idx, lsr):
        """In this VI scheme, this does nothing (reference :862-866)."""
        return L, 0., 0.

    def _update_annotation(self, vi_mu, vi_delta, hyper_delta, orig_obj, L, idx, lsr):
        return (vi_mu, vi_delta, hyper_delta), L, 0., 0.

    # ------------------------------------------------------------------ KL pieces
    def _delta_KL(self, vi_mu, vi_delta, hyper_delta):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        return float(self._res_stats[3 * self.num_pops])

    def _beta_KL(self, vi_mu, vi_delta, hyper_delta):
        self._make_resident((vi_mu, vi_delta, hyper_delta))
        return self._objective(self._res_stats)[1]

    def _annotation_KL(self, *params):
        return 0.
