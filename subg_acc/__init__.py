"""Top-level `subg_acc` module: the name the reference imports (`from subg_acc import gset_sampler, walk_sampler`,
sampler/random_walks.py:18; method table subg_acc/subg_acc.c:1036-1043).  With this repository on sys.path the
reference's sampler/ package resolves to the B200 implementation without editing a line; the functions are those of
surel_plus_b200.subg_acc (same keywords, return lists and exception classes)."""
from surel_plus_b200.subg_acc import batch_sampler, gset_sampler, walk_join, walk_sampler  # noqa: F401

__all__ = ["gset_sampler", "walk_sampler", "walk_join", "batch_sampler"]
