"""jax.tree.map / leaves over dicts (sorted keys), lists, tuples and NamedTuples."""
import dataclasses


def _is_leaf(x):
    return not isinstance(x, (dict, list, tuple)) and not (dataclasses.is_dataclass(x) and not isinstance(x, type))


def map(fn, t, *rest):  # noqa: A001
    if isinstance(t, dict):
        return {k: map(fn, t[k], *[r[k] for r in rest]) for k in t}
    if isinstance(t, tuple) and hasattr(t, "_fields"):
        return type(t)(*[map(fn, x, *[r[i] for r in rest]) for i, x in enumerate(t)])
    if isinstance(t, (list, tuple)):
        return type(t)(map(fn, x, *[r[i] for r in rest]) for i, x in enumerate(t))
    if dataclasses.is_dataclass(t) and not isinstance(t, type):
        return dataclasses.replace(t, **{f.name: map(fn, getattr(t, f.name), *[getattr(r, f.name) for r in rest])
                                         for f in dataclasses.fields(t)})
    return fn(t, *rest)


def leaves(t):
    if isinstance(t, dict):
        out = []
        for k in sorted(t):
            out.extend(leaves(t[k]))
        return out
    if isinstance(t, (list, tuple)):
        out = []
        for x in t:
            out.extend(leaves(x))
        return out
    if dataclasses.is_dataclass(t) and not isinstance(t, type):
        out = []
        for f in dataclasses.fields(t):
            out.extend(leaves(getattr(t, f.name)))
        return out
    return [t]
