"""The library baseline SURVEY §8(d) asks for beside the CPU one: a bench.py workload (c2 or c3) written
with stock torch-CUDA float64 operators - cuBLAS matmul, cuSOLVER / MAGMA cholesky and triangular solve,
elementwise kernels, autograd for the backward pass - un-fused: what a straight GPU translation of the
reference's TensorFlow graph gives.  It shares no code with the product's kernels or with oracle/ (the
equations are those of SURVEY §2a); the parameters are read from the same model object the product arm
times.  Measurement infrastructure only: imported by bench.py's ``library_baseline`` leg."""
import math

import torch


def _sqdist(A, B):
    return (A * A).sum(-1)[:, :, None] + (B * B).sum(-1)[:, None, :] - 2.0 * A @ B.transpose(-1, -2)


def layer_forward(X, p, eps):
    """X [N,D]; p: Z [K,M,D], ls [K,D], var [K], q_mu [M,P], q_sqrt [P,M,M], identity (mean function)."""
    Z, ls, var, q_mu, q_sqrt = p["Z"], p["ls"], p["var"], p["q_mu"], p["q_sqrt"]
    K, M, _ = Z.shape
    P = q_mu.shape[1]
    Zs, Xs = Z / ls[:, None, :], X[None] / ls[:, None, :]
    eye = torch.eye(M, dtype=X.dtype, device=X.device)
    Kuu = var[:, None, None] * torch.exp(-0.5 * _sqdist(Zs, Zs).clamp_min(0.0)) + 1e-6 * eye
    L = torch.linalg.cholesky(Kuu)
    Kuf = var[:, None, None] * torch.exp(-0.5 * _sqdist(Zs, Xs).clamp_min(0.0))  # [K,M,N]
    A = torch.linalg.solve_triangular(L, Kuf, upper=False)  # [K,M,N]
    Ap = A if K == P else A.expand(P, -1, -1)
    mean = (Ap * q_mu.T[:, :, None]).sum(1).T  # [N,P]
    if p["identity"]:
        mean = mean + X
    Lq = torch.tril(q_sqrt)
    B = Lq.transpose(-1, -2) @ Ap  # [P,M,N]
    varp = var if K == P else var.expand(P)
    fvar = varp[None, :] - (Ap * Ap).sum(1).T + (B * B).sum(1).T
    diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
    kl = 0.5 * ((q_mu ** 2).sum() - M * P - torch.log(diag * diag).sum() + (Lq * Lq).sum())
    f = mean if eps is None else mean + torch.sqrt(fvar) * eps
    return f, mean, fvar, kl


def loss_fn(params, noise, X, Y, eps, num_data):
    f, kls = X, []
    for i, p in enumerate(params):
        f, mean, fvar, kl = layer_forward(f, p, eps[i])
        kls.append(kl)
    ve = -0.5 * math.log(2.0 * math.pi) - 0.5 * torch.log(noise) - 0.5 * ((Y - mean) ** 2 + fvar) / noise
    return -ve.mean() + sum(kls) / num_data


def params_from_model(model, device):
    """Leaf tensors (requires_grad) holding the constrained parameter values of a gpflux_b200 DeepGP."""
    from gpflux_b200 import gpflow_compat as gc

    out = []
    for layer in model.f_layers:
        D = int(layer._inducing_points()[0].Z.shape[1])
        kind, Z, ls, var = layer._packed(D)
        if kind != "se" or not layer.whiten:
            raise NotImplementedError("library baseline: squared-exponential, whitened layers only")
        leaf = lambda t: t.detach().to(device).clone().requires_grad_(True)  # noqa: E731
        out.append({"Z": leaf(Z), "ls": leaf(ls), "var": leaf(var), "q_mu": leaf(layer.q_mu.value()),
                    "q_sqrt": leaf(layer.q_sqrt.value()), "identity": isinstance(layer.mean_function, gc.Identity)})
    noise = model.likelihood_layer.likelihood.variance.value().detach().to(device).clone().requires_grad_(True)
    return out, noise


def step_ms(params, noise, X, Y, num_data, steps=3, warmup=1):
    """Device time (CUDA events) of one eager fwd+bwd step on the rows given -> mean ms."""
    leaves = [t for p in params for t in p.values() if isinstance(t, torch.Tensor)] + [noise]
    ts = []
    for i in range(warmup + steps):
        eps = [torch.randn(X.shape[0], p["q_mu"].shape[1], dtype=X.dtype, device=X.device) for p in params[:-1]] + [None]
        for t in leaves:
            t.grad = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = loss_fn(params, noise, X, Y, eps, num_data)
        loss.backward()
        b.record()
        b.synchronize()
        if i >= warmup:
            ts.append(a.elapsed_time(b))
        del loss
    return sum(ts) / len(ts)
