"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def params_from(g, prefix="param/"):
    return {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith(prefix)}


def t(x):
    return torch.from_numpy(np.array(x))


def assert_close(actual, expected, rtol, atol, what=""):
    a = actual.detach().cpu().double().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual, dtype=np.float64)
    e = expected.detach().cpu().double().numpy() if isinstance(expected, torch.Tensor) else np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    err = np.abs(a - e)
    tol = atol + rtol * np.abs(e)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {e[i]!r} "
                             f"(|err|={err[i]:.3e}, tol={tol[i]:.3e}); max|err|={err.max():.3e}")


def assert_close_scaled(actual, expected, tol=2e-5, what="", floor=2e-7):
    """|a - e| <= tol * (|e| + max|e|): the fp32 bar (north_star: ~1e-5 relative), with the absolute part
    tied to the tensor's own scale so near-zero entries of a gradient are judged against its magnitude."""
    a = actual.detach().cpu().double().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual, dtype=np.float64)
    e = expected.detach().cpu().double().numpy() if isinstance(expected, torch.Tensor) else np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    assert np.all(np.isfinite(a)), f"{what}: non-finite values"
    scale = np.abs(e).max() if e.size else 0.0
    bound = tol * (np.abs(e) + scale) + floor
    err = np.abs(a - e)
    if not np.all(err <= bound):
        i = np.unravel_index(np.argmax(err / bound), err.shape)
        raise AssertionError(f"{what}: at {i} got {a[i]!r} want {e[i]!r} |err|={err[i]:.3e} bound={bound[i]:.3e} "
                             f"(scale {scale:.3e})")


def load_into(module, params):
    """Copy a reference-keyed parameter dict into a product module (state_dict layouts are identical)."""
    sd = {k: (v.clone() if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in params.items()}
    missing, unexpected = module.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return module


def _dead_biases(module):
    """Biases of Linear layers feeding a BatchNorm1d that runs on batch statistics: BN subtracts the batch mean, so
    their true gradient is exactly zero and every implementation (torch included) returns rounding noise."""
    dead = set()
    for name, seq in module.named_modules():
        if isinstance(seq, torch.nn.Sequential):
            mods = list(seq)
            for i in range(len(mods) - 1):
                if isinstance(mods[i], torch.nn.Linear) and isinstance(mods[i + 1], torch.nn.BatchNorm1d) \
                        and mods[i + 1].training:
                    dead.add(f"{name}.{i}.bias")
    return dead


def dead_bias_names(module):
    return _dead_biases(module)


def check_grads(module, expected, tol, what=""):
    """Compare every parameter gradient with `expected[name]`.  The absolute allowance is tied to the largest
    expected gradient entry; mathematically-zero gradients (see _dead_biases) are only required to be noise of
    the same order as the reference's own noise."""
    exp = {k: (v.detach().cpu().double().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64))
           for k, v in expected.items()}
    gmax = max(float(np.abs(v).max()) for v in exp.values() if v.size)
    dead = _dead_biases(module)
    for k, prm in module.named_parameters():
        assert prm.grad is not None, f"{what}: no gradient for {k}"
        if k in dead:
            a = prm.grad.detach().cpu().double().numpy()
            assert np.all(np.isfinite(a)), f"{what} grad {k}: non-finite"
            noise = max(float(np.abs(exp[k]).max()), 1e-7 * gmax)
            assert float(np.abs(a).max()) <= 20 * noise + 1e-3 * gmax + 1e-12, \
                f"{what} grad {k}: |{np.abs(a).max():.3e}| is not rounding noise (reference noise {noise:.3e})"
            continue
        assert_close_scaled(prm.grad, exp[k], tol, f"{what} grad {k}", floor=2e-6 * gmax + 1e-12)
