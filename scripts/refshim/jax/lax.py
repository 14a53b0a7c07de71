import torch


def stop_gradient(x):
    return x.detach() if isinstance(x, torch.Tensor) else x


def optimization_barrier(x):
    return x


def erfc(x):
    return torch.erfc(x)


def fori_loop(lower, upper, body, init):
    val = init
    for i in range(int(lower), int(upper)):
        val = body(i, val)
    return val


def scan(f, init, xs, length=None):
    from . import tree

    carry = init
    ys = []
    n = length if xs is None else tree.leaves(xs)[0].shape[0]
    for i in range(n):
        carry, y = f(carry, None if xs is None else tree.map(lambda t: t[i], xs))
        ys.append(y)
    return carry, (tree.map(lambda *t: torch.stack(t), *ys) if ys else ys)
