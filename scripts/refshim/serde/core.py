import dataclasses


def field(*args, **kwargs):
    for k in ("skip", "rename", "serializer", "deserializer", "flatten", "alias", "skip_if", "skip_if_false",
              "skip_if_default"):
        kwargs.pop(k, None)
    return dataclasses.field(*args, **kwargs)
