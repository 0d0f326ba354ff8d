"""Storage layer drop-in: mirrors reference src/semcode/storage/__init__.py (exports MilvusVectorStore)."""

from .milvus_store import (  # noqa: F401
    Entity,
    GpuCollection,
    Hit,
    Hits,
    MilvusVectorStore,
    SearchResult,
    drop_collection,
    has_collection,
)

__all__ = ["MilvusVectorStore", "GpuCollection", "Hit", "Hits", "Entity", "SearchResult", "has_collection",
           "drop_collection"]
