"""Laplacian factories of the oracle (TEST INFRASTRUCTURE).

reference: src/deepqmc/physics.py:144-156 (reverse_forward_laplacian: linearize(grad f),
loop over the 3N unit vectors).  Two independent implementations so the oracle checks itself.
"""
import torch


def laplacian_hessian(f, x):
    """(trace of Hessian, gradient) of scalar f at x[3N] via torch.func."""
    g = torch.func.grad(f)(x)
    h = torch.func.hessian(f)(x)
    return torch.diagonal(h).sum(), g


def laplacian_jvp_loop(f, x):
    """reference loop: acc += jvp(grad f)(e_i)[i] for i in range(3N)."""
    grad_f = torch.func.grad(f)
    g = grad_f(x)
    acc = torch.zeros((), dtype=x.dtype)
    eye = torch.eye(len(x), dtype=x.dtype)
    for i in range(len(x)):
        _, t = torch.func.jvp(grad_f, (x,), (eye[i],))
        acc = acc + t[i]
    return acc, g
