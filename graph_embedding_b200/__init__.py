"""graph_embedding_b200 — B200-native random-walk engine behind the reference's own API.

  node2vec   drop-in for node2vec/src/node2vec.py (Graph, alias_setup, alias_draw)
  main       drop-in CLI for node2vec/src/main.py (--input/--output/--p/--q/--walk-length/--num-walks ...)
  simrank    Python mirror of the Java TopSim surface (Graph, SingleRandomWalk, SimRank, Print, Eval)
  _lib       ctypes binding of libgraphwalk.so (include/graphwalk.h), the C-ABI boundary
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "node2vec", "main", "simrank", "dist"]
