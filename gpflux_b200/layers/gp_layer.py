"""``GPLayer``: host-side mirror of ``gpflux.layers.GPLayer`` (reference
gpflux/layers/gp_layer.py:43-384) whose arithmetic runs on the B200 through the C-ABI
(``gpfx_layer_forward`` / ``gpfx_layer_backward``).

Same constructor arguments, attributes (``q_mu [M,P]``, ``q_sqrt [P,M,M]``, ``kernel``,
``inducing_variable``, ``mean_function``, ``num_data``, ``whiten``, ``num_samples``, ``full_cov``,
``full_output_cov``, ``num_latent_gps``) and methods (``predict``, ``call``, ``prior_kl``,
``sample``) as the reference.  Differences forced by the host framework (PyTorch instead of
Keras/TFP): ``call`` returns a small ``MultivariateNormalDiag`` object that already carries the
reparameterised sample drawn by the fused kernel; losses are exposed as ``layer.losses`` (a list,
replaced on every call - Keras' ``add_loss`` semantics asserted by reference
tests/gpflux/layers/test_gp_layer.py:159-180).
"""
from __future__ import annotations

import warnings
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import ops
from ..exceptions import GPLayerIncompatibilityException
from ..gpflow_compat import (
    Identity,
    InducingPoints,
    MeanFunction,
    MultioutputInducingVariables,
    MultioutputKernel,
    Parameter,
    Stationary,
    default_float,
    default_jitter,
)
from ..runtime_checks import verify_compatibility
from .. import _cabi


class MultivariateNormalDiag:
    """Stand-in for ``tfp.distributions.MultivariateNormalDiag(loc, scale_diag)`` as returned by
    the reference's ``_make_distribution_fn`` (gp_layer.py:339-341).  ``variance()`` is stored (not
    the scale) so nothing is recomputed; ``sample()`` returns the draw the fused layer kernel
    already made when there is one."""

    def __init__(self, loc: torch.Tensor, var: torch.Tensor, sample: Optional[torch.Tensor] = None):
        self.loc = loc
        self._var = var
        self._sample = sample

    def mean(self) -> torch.Tensor:
        return self.loc

    def variance(self) -> torch.Tensor:
        return self._var

    @property
    def scale_diag(self) -> torch.Tensor:
        return torch.sqrt(self._var)

    @property
    def shape(self):
        return self.loc.shape

    @property
    def dtype(self):
        return self.loc.dtype

    def sample(self, sample_shape=(), eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        sample_shape = tuple(sample_shape)
        if not sample_shape and eps is None and self._sample is not None:
            return self._sample
        shape = sample_shape + tuple(self.loc.shape)
        if eps is None:
            eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        loc = self.loc.expand(shape).contiguous()
        var = self._var.expand(shape).contiguous()
        return ops.reparam_sample(loc.detach(), var.detach(), eps)

    def _tensor(self) -> torch.Tensor:
        """Keras coerces the distribution to a tensor via convert_to_tensor_fn (a sample)."""
        return self.sample()


class MultivariateNormalTriL:
    """Stand-in for ``tfp.distributions.MultivariateNormalTriL(loc, scale_tril)`` as returned by the
    reference's ``_make_distribution_fn`` for ``full_cov`` (loc [Q,N], scale [Q,N,N]) and
    ``full_output_cov`` (loc [N,Q], scale [N,Q,Q]) - gp_layer.py:329-338."""

    def __init__(self, loc: torch.Tensor, scale_tril: torch.Tensor):
        self.loc = loc
        self.scale_tril = scale_tril

    def mean(self) -> torch.Tensor:
        return self.loc

    def covariance(self) -> torch.Tensor:
        return ops.gemm(self.scale_tril, self.scale_tril, transB=True)

    def variance(self) -> torch.Tensor:
        """Marginal variances (tfp's MultivariateNormalTriL.variance()): row sums of scale_tril^2."""
        return (self.scale_tril * self.scale_tril).sum(-1)

    @property
    def shape(self):
        return self.loc.shape

    @property
    def dtype(self):
        return self.loc.dtype

    def sample(self, sample_shape=(), eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        """loc + scale_tril @ eps, eps ~ N(0, I) of shape sample_shape + loc.shape."""
        sample_shape = tuple(sample_shape)
        S = 1
        for d in sample_shape:
            S *= int(d)
        B, n = self.loc.shape
        if eps is None:
            eps = torch.randn(sample_shape + (B, n), dtype=self.loc.dtype, device=self.loc.device)
        e = eps.reshape(S, B, n).permute(1, 2, 0).contiguous()  # [B, n, S]
        out = ops.gemm(self.scale_tril, e, tri=_cabi.TRI_LOWER_A)  # [B, n, S]
        out = out.permute(2, 0, 1) + self.loc
        return out.reshape(sample_shape + (B, n))


def unwrap_dist(x):
    """Mirrors gpflux/types.py:27-35."""
    return x


class GPLayer(nn.Module):
    def __init__(
        self,
        kernel: MultioutputKernel,
        inducing_variable: MultioutputInducingVariables,
        num_data: int,
        mean_function: Optional[MeanFunction] = None,
        *,
        num_samples: Optional[int] = None,
        full_cov: bool = False,
        full_output_cov: bool = False,
        num_latent_gps: int = None,
        whiten: bool = True,
        name: Optional[str] = None,
        verbose: bool = True,
    ):
        super().__init__()
        self.name = name
        self.kernel = kernel
        self.inducing_variable = inducing_variable
        self.num_data = num_data

        if mean_function is None:
            mean_function = Identity()
            if verbose:
                warnings.warn(
                    "Beware, no mean function was specified in the construction of the `GPLayer` "
                    "so the default `gpflow.mean_functions.Identity` is being used. "
                    "This mean function will only work if the input dimensionality "
                    "matches the number of latent Gaussian processes in the layer."
                )
        self.mean_function = mean_function
        self.full_output_cov = full_output_cov
        self.full_cov = full_cov
        self.whiten = whiten
        self.verbose = verbose

        try:
            num_inducing, self.num_latent_gps = verify_compatibility(kernel, mean_function, inducing_variable)
        except GPLayerIncompatibilityException as e:
            if num_latent_gps is None:
                raise e
            if verbose:
                warnings.warn(
                    "Could not verify the compatibility of the `kernel`, `inducing_variable` "
                    "and `mean_function`. We advise using `gpflux.helpers.construct_*` to create "
                    "compatible kernels and inducing variables. As "
                    f"`num_latent_gps={num_latent_gps}` has been specified explicitly, this will "
                    "be used to create the `q_mu` and `q_sqrt` parameters."
                )
            num_inducing, self.num_latent_gps = inducing_variable.num_inducing, num_latent_gps

        self.q_mu = Parameter(
            np.zeros((num_inducing, self.num_latent_gps)),
            name=f"{self.name}_q_mu" if self.name else "q_mu",
        )  # [num_inducing, num_latent_gps]
        self.q_sqrt = Parameter(
            np.stack([np.eye(num_inducing) for _ in range(self.num_latent_gps)]),
            transform="triangular",
            name=f"{self.name}_q_sqrt" if self.name else "q_sqrt",
        )  # [num_latent_gps, num_inducing, num_inducing]
        self.num_samples = num_samples
        # B200 addition: how A = L^-1 Kuf is evaluated - "gemm" (explicit-inverse DMMA GEMM, default) or
        # "trsm" (blocked forward substitution, the reference-faithful form for badly conditioned Kuu);
        # ``choose_solver()`` picks from a condition estimate of Kuu
        self.solver = "gemm"
        self.losses = []
        self.metrics = {}
        self._noise_seed = None
        self._noise_calls = 0

    def seed_noise(self, seed: Optional[int]) -> None:
        """Opt in to (``None``: out of) the library's counter-based generator for the layer's
        reparameterisation noise: call c then draws eps = ops.randn_philox((N, P), seed, offset=c) in one
        kernel of ours instead of torch.randn.  Equal in distribution to the reference's TFP draw, not
        stream-identical (nothing is: SURVEY §8c); pass ``eps=`` to ``call`` for parity runs.  The
        offset advances on the host, so a captured CUDA graph would replay one fixed draw - keep the
        default generator (whose state torch registers with the graph) under graph capture."""
        self._noise_seed = None if seed is None else int(seed)
        self._noise_calls = 0

    def _draw_eps(self, shape, like):
        if self._noise_seed is None or not like.is_cuda:
            return torch.randn(shape, dtype=like.dtype, device=like.device)
        eps = ops.randn_philox(shape, self._noise_seed, self._noise_calls, device=like.device)
        self._noise_calls += 1
        return eps

    # ------------------------------------------------------------------ parameter packing
    def _latent_kernels(self):
        if isinstance(self.kernel, MultioutputKernel):
            return tuple(self.kernel.latent_kernels)
        # KernelWithFeatureDecomposition wraps the exact kernel the conditional uses
        return (getattr(self.kernel, "_kernel", self.kernel),)

    def _inducing_points(self):
        if isinstance(self.inducing_variable, MultioutputInducingVariables):
            return tuple(self.inducing_variable.inducing_variables)
        return (self.inducing_variable,)

    def _packed(self, input_dim: int):
        """-> kind, Z [K,M,D], lengthscales [K,D], variance [K] with K = 1 (everything shared) or P."""
        kernels, ivs = self._latent_kernels(), self._inducing_points()
        for k in kernels:
            if not isinstance(k, Stationary):
                raise NotImplementedError(f"kernel {type(k).__name__} is not supported by the B200 SVGP path")
        kinds = {k.kind for k in kernels}
        if len(kinds) != 1:
            raise NotImplementedError("mixed kernel families are evaluated group by group (_family_groups)")
        for iv in ivs:
            if not isinstance(iv, InducingPoints):
                raise NotImplementedError("only InducingPoints inducing variables are supported")
        K = max(len(kernels), len(ivs))
        if K > 1 and K != self.num_latent_gps:
            raise GPLayerIncompatibilityException(
                f"{K} separate kernels / inducing variables for {self.num_latent_gps} latent GPs"
            )
        Zs = [iv.Z.value() for iv in ivs]
        ls = [k.lengthscales_vector(input_dim) for k in kernels]
        var = [k.variance.value().reshape(()) for k in kernels]
        if K == 1:
            # shared kernel + shared inducing points: views only, no stack / copy kernels
            return kinds.pop(), Zs[0].unsqueeze(0), ls[0].unsqueeze(0), var[0].reshape(1)
        if len(Zs) == 1:
            Zs = Zs * K
        if len(ls) == 1:
            ls, var = ls * K, var * K
        return kinds.pop(), torch.stack(Zs), torch.stack(ls), torch.stack(var)

    def _family_groups(self):
        """SeparateIndependent with MIXED kernel families (e.g. [SquaredExponential, Matern32], reference
        tests/gpflux/layers/basis_functions/fourier_features/test_random.py:209-228): the device entry
        points take one kernel family per call, so the latent GPs are grouped by family and each group is
        evaluated as its own (smaller) separate-kernel layer.  -> None if the layer has a single family,
        else [(kind, [p indices])] in first-appearance order."""
        kernels = self._latent_kernels()
        if len({k.kind for k in kernels if isinstance(k, Stationary)}) <= 1:
            return None
        groups = {}
        for p, k in enumerate(kernels):
            groups.setdefault(k.kind, []).append(p)
        return list(groups.items())

    def _forward_mixed(self, inputs, eps, groups):
        """One fused layer call per kernel family; outputs are stitched back in latent-GP order."""
        kernels, ivs = self._latent_kernels(), self._inducing_points()
        D = inputs.shape[1]
        P = self.num_latent_gps
        if len(kernels) != P:
            raise GPLayerIncompatibilityException(f"{len(kernels)} separate kernels for {P} latent GPs")
        for k in kernels:
            if not isinstance(k, Stationary):
                raise NotImplementedError(f"kernel {type(k).__name__} is not supported by the B200 SVGP path")
        mean_add = self._mean_add(inputs)
        q_mu, q_sqrt = self.q_mu.value(), self._q_sqrt_raw()
        N = inputs.shape[0]
        cols = [None] * P
        kl_total = None
        for kind, idx in groups:
            ii = torch.as_tensor(idx, device=inputs.device)
            Z = torch.stack([(ivs[p] if len(ivs) > 1 else ivs[0]).Z.value() for p in idx])
            ls = torch.stack([kernels[p].lengthscales_vector(D) for p in idx])
            var = torch.stack([kernels[p].variance.value().reshape(()) for p in idx])
            out = ops.svgp_layer(
                inputs, Z, ls, var, q_mu.index_select(1, ii), q_sqrt.index_select(0, ii),
                mean_add=None if mean_add is None else mean_add.index_select(1, ii),
                eps=None if eps is None else eps.index_select(1, ii),
                kernel=kind, whiten=self.whiten, jitter=default_jitter(), solver=self.solver,
            )
            for j, p in enumerate(idx):
                cols[p] = tuple(o[:, j] for o in out[:3])
            kl_total = out[3] if kl_total is None else kl_total + out[3]
        fsample, fmean, fvar = (torch.stack([c[i] for c in cols], dim=1) for i in range(3))
        if N == 0:
            fsample = fmean = fvar = torch.zeros((0, P), dtype=inputs.dtype, device=inputs.device)
        return fsample, fmean, fvar, kl_total

    # ------------------------------------------------------------------ reference API
    def _check_inputs(self, inputs):
        if hasattr(inputs, "_tensor"):
            inputs = inputs._tensor()
        if not isinstance(inputs, torch.Tensor):
            if hasattr(inputs, "__dlpack__") or type(inputs).__name__ == "PyCapsule":
                # any DLPack producer: validated (device / dtype / strides / alignment) and viewed zero-copy
                from ..dlpack import from_dlpack

                inputs = from_dlpack(inputs, ndim=2, what="inputs")
            else:
                inputs = torch.as_tensor(inputs)
        if inputs.dtype != default_float():
            raise ValueError(
                f"inputs need to have dtype {default_float()} (gpflow.default_float()), got {inputs.dtype}"
            )
        if inputs.dim() != 2:
            raise ValueError(f"inputs must be [N, D], got shape {tuple(inputs.shape)}")
        return inputs

    def _status_buffer(self, device):
        buf = getattr(self, "_status", None)
        if buf is None or buf.device != device:
            buf = torch.zeros((max(self.num_latent_gps, 1),), dtype=torch.int32, device=device)
            self._status = buf
        return buf

    def check_status(self) -> None:
        """Raise ``GpfxError`` (status GPFX_E_NOT_PD) if the last evaluation of this layer met a Kuu that is
        not positive definite - where the reference raises InvalidArgumentError from tf.linalg.cholesky.
        The status is copied to a small device buffer by every forward (no sync on the hot path); this
        call is the one device->host read.  ``TrainingModel.fit`` / ``NatGradModel`` call it when a loss
        comes back non-finite; call it yourself after ``predict``."""
        buf = getattr(self, "_status", None)
        if buf is None:
            return
        info = buf.cpu()
        bad = torch.nonzero(info)
        if bad.numel():
            k = int(bad[0])
            raise _cabi.GpfxError(
                f"gpfx status {_cabi.E_NOT_PD} (GPFX_E_NOT_PD): Kuu[{k}] of layer {self.name or ''!r} is not positive "
                f"definite (first non-positive pivot {int(info[k]) - 1})")

    def _forward_fused(self, inputs, eps):
        with ops.status_sink(self._status_buffer(inputs.device)):
            return self._forward_fused_impl(inputs, eps)

    def _forward_fused_impl(self, inputs, eps):
        pending = getattr(self, "_pending_factor", None)
        self._pending_factor = None
        groups = self._family_groups()
        if groups is not None and inputs.shape[0] > 0:
            return self._forward_mixed(inputs, eps, groups)
        if inputs.shape[0] == 0:
            # empty minibatch: nothing to condition on; only the parameter-only KL is evaluated
            empty = torch.zeros((0, self.num_latent_gps), dtype=inputs.dtype, device=inputs.device)
            if pending is not None:
                if pending[7] is not None:  # the KL was computed on the side stream
                    torch.cuda.current_stream().wait_stream(pending[7])
                kl = pending[6][2]
            else:
                kl = self.prior_kl()
            return empty, empty.clone(), empty.clone(), kl
        if pending is not None:
            kind, Z, ls, var, q_mu, q_sqrt, factor, stream = pending
            if Z.shape[2] != inputs.shape[1]:
                raise ValueError(f"layer expects input dim {Z.shape[2]}, got {inputs.shape[1]}")
            mean_add = self._mean_add(inputs)
            if stream is not None:
                torch.cuda.current_stream().wait_stream(stream)
            fsample, fmean, fvar = ops.svgp_cond(
                inputs, Z, ls, var, q_mu, q_sqrt, factor,
                mean_add=mean_add, eps=eps, kernel=kind, whiten=self.whiten, jitter=default_jitter(),
                solver=self.solver,
            )
            return fsample, fmean, fvar, factor[2].to(default_float())
        kind, Z, ls, var = self._packed(inputs.shape[1])
        mean_add = self._mean_add(inputs)
        return ops.svgp_layer(
            inputs, Z, ls, var, self.q_mu.value(), self._q_sqrt_raw(),
            mean_add=mean_add, eps=eps, kernel=kind, whiten=self.whiten, jitter=default_jitter(),
            solver=self.solver,
        )

    def choose_solver(self, threshold: float = 1e-10) -> str:
        """Pick ``self.solver`` from a condition estimate of Kuu: cond(Kuu) ~ (max diag L / min diag L)^2
        from the Cholesky factor; the explicit-inverse GEMM and the substitution differ at O(cond * eps),
        so "trsm" is selected when cond * eps > ``threshold``.  One device->host read: call it outside the
        training loop (e.g. once per epoch); the choice is not re-evaluated automatically."""
        ivs = self._inducing_points()
        D = int(ivs[0].Z.shape[1])
        with torch.no_grad():
            kind, Z, ls, var = self._packed(D)
            Kuu = ops.kernel_matrix(kind, Z, Z, ls, var, jitter=default_jitter())
            L, _, info = ops.cholesky_inverse(Kuu)
            if int(info.abs().max()) != 0:
                raise _cabi.GpfxError("Kuu is not positive definite (gpfx status E_NOT_PD)")
            d = torch.diagonal(L, dim1=-2, dim2=-1)
            cond = float(((d.max(dim=-1).values / d.min(dim=-1).values) ** 2).max())
        self.solver = "trsm" if cond * 2.220446049250313e-16 > threshold else "gemm"
        return self.solver

    def _mean_add(self, inputs):
        mean_add = self.mean_function(inputs)
        if mean_add is not None and tuple(mean_add.shape) != (inputs.shape[0], self.num_latent_gps):
            raise ValueError(
                f"mean function output has shape {tuple(mean_add.shape)}, expected "
                f"{(inputs.shape[0], self.num_latent_gps)} (Identity needs input_dim == num_latent_gps)"
            )
        return mean_add

    def _q_sqrt_raw(self):
        # the kernels read only the lower triangle of q_sqrt and return a lower-triangular gradient,
        # so the triangular() transform needs no separate tril pass on the hot path
        return self.q_sqrt.unconstrained_variable if self.q_sqrt.transform == "triangular" else self.q_sqrt.value()

    def factorize(self, stream: Optional["torch.cuda.Stream"] = None) -> None:
        """Run the parameter-only half of the layer now (tril(q_sqrt), KL, Kuu, Cholesky, L^-1 -
        ``gpfx_layer_factor_forward``), optionally on a side ``stream`` so that it overlaps the data
        path of other layers.  The next ``call``/``predict`` consumes it.  ``DeepGP`` does this for
        all its layers before evaluating the first one."""
        if self._family_groups() is not None:
            return  # mixed kernel families: evaluated group by group inside the call
        ivs = self._inducing_points()
        D = int(ivs[0].Z.shape[1])
        kind, Z, ls, var = self._packed(D)
        q_mu, q_sqrt = self.q_mu.value(), self._q_sqrt_raw()
        kw = dict(kernel=kind, whiten=self.whiten, jitter=default_jitter())
        sink = ops.status_sink(self._status_buffer(Z.device))
        if stream is None:
            with sink:
                factor = ops.svgp_factor(Z, ls, var, q_mu, q_sqrt, **kw)
        else:
            main = torch.cuda.current_stream()
            stream.wait_stream(main)
            # the side stream reads tensors the caching allocator handed out on the main stream
            for t in (Z, ls, var, q_mu, q_sqrt):
                t.record_stream(stream)
            with torch.cuda.stream(stream), sink:
                factor = ops.svgp_factor(Z, ls, var, q_mu, q_sqrt, **kw)
            for t in (factor[0], factor[2]):
                t.record_stream(main)
        self._pending_factor = (kind, Z, ls, var, q_mu, q_sqrt, factor, stream)

    def predict(self, inputs, *, full_cov: bool = False, full_output_cov: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Posterior mean [N,Q] and marginal variance [N,Q] at ``inputs`` including the mean function
        (reference gp_layer.py:226-268).  ``full_cov`` -> cov [Q,N,N]; ``full_output_cov`` -> [N,Q,Q];
        both -> [N,Q,N,Q] (gp_layer.py:238-246)."""
        inputs = self._check_inputs(inputs)
        if full_cov:
            mean, cov = self._predict_full_cov(inputs)  # [N,Q], [Q,N,N]
            if full_output_cov:  # gpflow expand_independent_outputs: [N,Q,N,Q], block diagonal in Q
                Q, N = cov.shape[0], cov.shape[1]
                eye = torch.eye(Q, dtype=cov.dtype, device=cov.device)
                cov = cov.permute(1, 2, 0)[:, None, :, :] * eye[None, :, None, :]  # [N,Q,N,Q']
                cov = cov.permute(0, 1, 2, 3).contiguous()
            return mean, cov
        _, mean, var, _ = self._forward_fused(inputs, None)
        if full_output_cov:  # independent latent GPs: [N,Q,Q] diagonal (gpflow expand_independent_outputs)
            return mean, torch.diag_embed(var)
        return mean, var

    def _predict_full_cov(self, inputs) -> Tuple[torch.Tensor, torch.Tensor]:
        """``conditional(..., full_cov=True)`` (gp_layer.py:257-266 -> base_conditional):
        cov_p = Knn - A^T A + B_p^T B_p with A = L^-1 Kuf (then L^-T A when not whitened) and
        B_p = tril(q_sqrt_p)^T A, composed from the C-ABI GEMM / kernel-matrix / Cholesky entry points.
        Not on the training path (the reference uses it to draw correlated samples at a few
        hundred inputs): forward only, no gradients."""
        from ..conditionals import full_cov_conditional

        D = inputs.shape[1]
        kind, Z, ls, var = self._packed(D)
        mean, cov = full_cov_conditional(inputs, kind, Z, ls, var, self.q_mu.value(), self.q_sqrt.value(),
                                         whiten=self.whiten, jitter=default_jitter())
        mean_add = self._mean_add(inputs)
        if mean_add is not None:
            mean = mean + mean_add
        return mean, cov

    def call(self, inputs, *args, training: Optional[bool] = None, eps: Optional[torch.Tensor] = None, **kwargs):
        """Reference gp_layer.py:270-301.  ``eps`` ([N,P]) fixes the reparameterisation noise (parity
        runs); by default it is drawn with torch.randn on the device."""
        if self.full_cov or self.full_output_cov:
            if self.full_cov and self.full_output_cov:
                raise NotImplementedError(
                    "The combination of both `full_cov` and `full_output_cov` is not permitted."
                )
            return self._call_full(self._check_inputs(inputs), training, eps)
        inputs = self._check_inputs(inputs)
        N, P = inputs.shape[0], self.num_latent_gps
        if eps is None and self.num_samples is None:
            eps = self._draw_eps((N, P), inputs)
        fsample, fmean, fvar, kl = self._forward_fused(inputs, eps if self.num_samples is None else None)
        if self.num_samples is not None:
            dist = MultivariateNormalDiag(fmean, fvar)
            dist._sample = _reparam_autograd(fmean, fvar, (self.num_samples, N, P), eps)
        else:
            dist = MultivariateNormalDiag(fmean, fvar, fsample)

        if training:
            loss_per_datapoint = self._loss_per_datapoint(kl)
        else:
            loss_per_datapoint = torch.zeros((), dtype=default_float(), device=inputs.device)
        self.losses = [loss_per_datapoint]
        name = f"{self.name}_prior_kl" if self.name else "prior_kl"
        self.metrics = {name: loss_per_datapoint.detach()}
        return dist

    forward = call

    def _loss_per_datapoint(self, kl):
        """(KL - sum of log prior densities) / num_data (reference gp_layer.py:286-289).  Parameters
        without a prior contribute exactly 0 and launch nothing."""
        params = self.kernel.trainable_parameters if hasattr(self.kernel, "trainable_parameters") else []
        priors = [p.log_prior_density() for p in params if p.prior is not None]
        if priors:
            kl = kl - sum(priors[1:], priors[0])
        return kl / self.num_data

    def _call_full(self, inputs, training, eps):
        """``full_cov`` / ``full_output_cov`` branches of _make_distribution_fn / _convert_to_tensor_fn
        (gp_layer.py:329-338, 347-368) with math._cholesky_with_jitter (math.py:25-39).  Differentiable like the
        reference: every step is an autograd Function over C-ABI calls (conditionals.py)."""
        jitter = default_jitter()
        mean, cov = self.predict(inputs, full_cov=self.full_cov, full_output_cov=self.full_output_cov)
        if self.full_cov:  # loc [Q,N], scale [Q,N,N]
            n = cov.shape[-1]
            covj = cov + jitter * torch.eye(n, dtype=cov.dtype, device=cov.device)
            scale, _, info = ops.cholesky_inverse(covj)
            dist = MultivariateNormalTriL(mean.T.contiguous(), scale)
        else:  # loc [N,Q], scale [N,Q,Q]: the covariance is diagonal, so is its Cholesky factor
            var = torch.diagonal(cov, dim1=-2, dim2=-1)
            dist = MultivariateNormalTriL(mean, torch.diag_embed(torch.sqrt(var + jitter)))
        shape = () if self.num_samples is None else (self.num_samples,)
        samples = dist.sample(shape, eps=eps)
        if self.full_cov:
            samples = samples.transpose(-1, -2)  # [S,N,Q] or [N,Q]
        dist._sample = samples
        dist._tensor = lambda: samples
        if training:
            kl = self.prior_kl()
            loss_per_datapoint = self._loss_per_datapoint(kl)
        else:
            loss_per_datapoint = torch.zeros((), dtype=default_float(), device=inputs.device)
        self.losses = [loss_per_datapoint]
        name = f"{self.name}_prior_kl" if self.name else "prior_kl"
        self.metrics = {name: loss_per_datapoint.detach()}
        return dist

    def prior_kl(self) -> torch.Tensor:
        """KL[q(v) || p(v)] (reference gp_layer.py:303-311); whitened representation only."""
        if not self.whiten:
            # KL[q(u) || N(0, Kuu)] needs the Cholesky factor: run the parameter-only half of the layer
            ivs = self._inducing_points()
            kind, Z, ls, var = self._packed(int(ivs[0].Z.shape[1]))
            return ops.svgp_factor(Z, ls, var, self.q_mu.value(), self._q_sqrt_raw(), kernel=kind, whiten=False,
                                   jitter=default_jitter())[2].to(default_float())
        return _KLWhiteFn.apply(self.q_mu.value(), self.q_sqrt.value())

    def _convert_to_tensor_fn(self, distribution: MultivariateNormalDiag) -> torch.Tensor:
        """Reference gp_layer.py:347-368 (the full_cov draw is transposed back to [.., N, Q])."""
        if isinstance(distribution, MultivariateNormalTriL):
            return distribution._tensor()
        return distribution.sample()

    def sample(self):
        """Reference gp_layer.py:370-384: efficient_sample(...) + mean_function."""
        from ..sampling.sample import efficient_sample

        return efficient_sample(
            self.inducing_variable, self.kernel, self.q_mu.value(), q_sqrt=self.q_sqrt.value(), whiten=self.whiten
        ) + self.mean_function


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, var, eps):
        ctx.save_for_backward(var, eps)
        return ops.reparam_sample(mean, var, eps)

    @staticmethod
    def backward(ctx, g):
        var, eps = ctx.saved_tensors
        # d/dvar sqrt(var) eps = eps / (2 sqrt(var)); tiny [S,N,P] op outside the fused path
        return g, g * eps / (2.0 * torch.sqrt(var)), None


def _reparam_autograd(mean, var, shape, eps):
    if eps is None:
        eps = torch.randn(shape, dtype=mean.dtype, device=mean.device)
    S = shape[0]
    out = _ReparamFn.apply(mean.expand(shape).contiguous(), var.expand(shape).contiguous(), eps)
    return out.reshape((S,) + tuple(mean.shape))


class _KLWhiteFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_mu, q_sqrt):
        ctx.save_for_backward(q_mu, q_sqrt)
        return ops.kl_white(q_mu, q_sqrt)

    @staticmethod
    def backward(ctx, g):
        q_mu, q_sqrt = ctx.saved_tensors
        return ops.kl_white_bwd(q_mu, q_sqrt, g)
