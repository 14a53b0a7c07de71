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
    carry = init
    ys = []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, (torch.stack(ys) if ys and isinstance(ys[0], torch.Tensor) else ys)
