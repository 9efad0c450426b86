"""Test stub of `patchify` (imported by the reference's data/patches.py; the integration test trains with --use_patches '')."""


def patchify(*a, **k):
    raise NotImplementedError('patchify stub')
