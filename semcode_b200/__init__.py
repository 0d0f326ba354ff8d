"""semcode_b200 -- B200-native IVF_FLAT engine behind semcode's Milvus wrapper.

Public surface:
  * ``semcode_b200.storage.MilvusVectorStore`` -- drop-in for reference
    src/semcode/storage/milvus_store.py (connect / upsert_embeddings / search)
  * ``semcode_b200.IVFFlatIndex`` -- one index on one GPU (train / add / search, numpy or torch)
  * ``semcode_b200.ShardedIVFFlat`` -- row-sharded index, one process per GPU over torch.distributed
  * ``include/semcode_ivf.h`` / ``libsemcode_ivf.so`` -- the C ABI everything above calls

There is no CPU fallback: importing is cheap, but every compute call needs the built shared
library and an sm_100 GPU.
"""

from ._capi import METRIC_IP, METRIC_L2, NativeError  # noqa: F401
from .index import IVFFlatIndex, merge_topk  # noqa: F401

__all__ = ["IVFFlatIndex", "merge_topk", "METRIC_IP", "METRIC_L2", "NativeError", "build"]


def build(force: bool = False) -> str:
    """Compile libsemcode_ivf.so in-tree (nvcc, sm_100a)."""
    from ._build import build as _b

    return _b(force=force)


def __getattr__(name):
    if name == "ShardedIVFFlat":
        from .sharded import ShardedIVFFlat

        return ShardedIVFFlat
    raise AttributeError(name)
