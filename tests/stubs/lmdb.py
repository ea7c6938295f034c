"""Test stub: the reference's utils/dataset.py imports lmdb at module level (utils/dataset.py:5); the package is not
installable on a no-network box and the drop-in test never opens a database."""


def open(*_a, **_k):  # noqa: A001
    raise RuntimeError("lmdb stub: no database access in the drop-in test")
