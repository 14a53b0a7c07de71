import typing

ArrayLike = typing.Any
DTypeLike = typing.Any
